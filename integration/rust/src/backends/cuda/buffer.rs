//! `src/backends/cuda/buffer.rs` — the device-buffer type: f32 storage resident in HBM with
//! pinned host staging, so chains of ops stay on the device (precedent in the reference:
//! `GpuCommandBatch::upload/read`, src/backends/gpu/batch.rs:140-200, and trueno-gpu's
//! `GpuBuffer<T>`, trueno-gpu/src/driver/memory.rs:53-164).
//!
//! Ops on `DeviceBuffer`s are stream-ordered on the backend's stream and do not synchronise;
//! `to_vec()` / `read_scalar()` are the synchronisation points.
use super::{check, sys};
use crate::TruenoError;
use core::ptr;

pub struct DeviceBuffer {
    raw: *mut sys::trn_buf,
}

// The handle is an owning pointer to device memory; the C library is thread-safe.
unsafe impl Send for DeviceBuffer {}
unsafe impl Sync for DeviceBuffer {}

impl DeviceBuffer {
    pub fn new(len: usize) -> Result<Self, TruenoError> {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::trn_buf_alloc(len, &mut raw) })?;
        Ok(Self { raw })
    }
    pub fn from_slice(data: &[f32]) -> Result<Self, TruenoError> {
        let buf = Self::new(data.len())?;
        check(unsafe { sys::trn_buf_upload(buf.raw, data.as_ptr(), data.len()) })?;
        Ok(buf)
    }
    pub fn len(&self) -> usize { unsafe { sys::trn_buf_len(self.raw) } }
    pub fn is_empty(&self) -> bool { self.len() == 0 }
    pub(crate) fn ptr(&self) -> *mut f32 { unsafe { sys::trn_buf_ptr(self.raw) } }

    pub fn to_vec(&self) -> Result<Vec<f32>, TruenoError> {
        let mut out = vec![0.0f32; self.len()];
        check(unsafe { sys::trn_buf_download(self.raw, out.as_mut_ptr(), out.len()) })?;
        Ok(out)
    }

    // ---- resident ops: results stay in HBM ------------------------------------------------------
    pub fn add(&self, other: &Self) -> Result<Self, TruenoError> {
        let out = Self::new(self.len())?;
        check(unsafe { sys::trn_add_f32_dev(self.ptr(), self.len(), other.ptr(), other.len(), out.ptr(), ptr::null_mut()) })?;
        Ok(out)
    }
    pub fn mul(&self, other: &Self) -> Result<Self, TruenoError> {
        let out = Self::new(self.len())?;
        check(unsafe { sys::trn_mul_f32_dev(self.ptr(), self.len(), other.ptr(), other.len(), out.ptr(), ptr::null_mut()) })?;
        Ok(out)
    }
    pub fn gelu(&self) -> Result<Self, TruenoError> {
        let out = Self::new(self.len())?;
        check(unsafe { sys::trn_gelu_f32_dev(self.ptr(), self.len(), out.ptr(), ptr::null_mut()) })?;
        Ok(out)
    }
    pub fn sigmoid(&self) -> Result<Self, TruenoError> {
        let out = Self::new(self.len())?;
        check(unsafe { sys::trn_sigmoid_f32_dev(self.ptr(), self.len(), out.ptr(), ptr::null_mut()) })?;
        Ok(out)
    }
    pub fn softmax_rows(&self, rows: usize, cols: usize) -> Result<Self, TruenoError> {
        let out = Self::new(self.len())?;
        check(unsafe { sys::trn_softmax_rows_f32_dev(self.ptr(), out.ptr(), rows, cols, ptr::null_mut()) })?;
        Ok(out)
    }
    /// C[m x n] = self[m x k] * other[k x n], all three resident.
    pub fn matmul(&self, other: &Self, m: usize, k: usize, n: usize) -> Result<Self, TruenoError> {
        let out = Self::new(m * n)?;
        check(unsafe { sys::trn_matmul_f32_dev(self.ptr(), m, k, other.ptr(), k, n, out.ptr(), ptr::null_mut()) })?;
        Ok(out)
    }
    /// Scalar reductions land in a one-element device buffer; read it back with `to_vec()`.
    pub fn dot(&self, other: &Self) -> Result<Self, TruenoError> {
        let out = Self::new(1)?;
        check(unsafe { sys::trn_dot_f32_dev(self.ptr(), self.len(), other.ptr(), other.len(), out.ptr(), ptr::null_mut()) })?;
        Ok(out)
    }
    pub fn sum(&self) -> Result<Self, TruenoError> {
        let out = Self::new(1)?;
        check(unsafe { sys::trn_sum_f32_dev(self.ptr(), self.len(), out.ptr(), ptr::null_mut()) })?;
        Ok(out)
    }
}

impl Drop for DeviceBuffer {
    fn drop(&mut self) {
        // stream-ordered work that still uses the buffer is finished by cudaFree's implicit sync
        unsafe { sys::trn_buf_free(self.raw) };
    }
}
