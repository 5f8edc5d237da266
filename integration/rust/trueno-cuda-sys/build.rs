//! Compiles the CUDA sources of the B200 backend with nvcc for sm_100a only and links the
//! resulting shared library.  There is deliberately no fallback: if nvcc is missing the build
//! fails, it never produces a CPU stub.
//!
//! Layout expected (this repository): `<repo>/trueno_b200/csrc/*.cu`, `<repo>/include/trueno_cuda.h`.
//! Override the source root with TRUENO_CUDA_SRC and the compiler with NVCC.
use std::env;
use std::path::PathBuf;
use std::process::Command;

// Every translation unit under trueno_b200/csrc — the same list, in the same order, as trueno_b200/build.py
// (tests/test_build_recipes_cpu.py compares the two with the directory listing).
const SOURCES: &[&str] = &[
    "context.cu", "reduce.cu", "map.cu", "softmax.cu", "gemm_simt.cu", "gemm_tc.cu", "batch.cu", "peer.cu", "conv.cu",
    "attention.cu", "eigen.cu", "gather.cu", "api.cu",
];
// Headers every object depends on.
const HEADERS: &[&str] = &["common.cuh", "tcgen05.cuh"];

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let root = env::var("TRUENO_CUDA_SRC")
        .map(PathBuf::from)
        .unwrap_or_else(|_| manifest.join("../../.."));
    let csrc = root.join("trueno_b200/csrc");
    let include = root.join("include");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());

    let mut objs = Vec::new();
    for src in SOURCES {
        let path = csrc.join(src);
        println!("cargo:rerun-if-changed={}", path.display());
        let obj = out.join(src.replace(".cu", ".o"));
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"])
            .args(["-Xcompiler", "-fPIC,-fvisibility=hidden,-O2", "-cudart", "shared"])
            .arg("-I").arg(&include).arg("-I").arg(&csrc)
            .arg("-c").arg(&path).arg("-o").arg(&obj)
            .status()
            .expect("nvcc not found: the trueno CUDA backend targets sm_100a and has no fallback build");
        assert!(status.success(), "nvcc failed for {}", src);
        objs.push(obj);
    }
    for h in HEADERS {
        println!("cargo:rerun-if-changed={}", csrc.join(h).display());
    }
    println!("cargo:rerun-if-changed={}", include.join("trueno_cuda.h").display());

    let lib = out.join("libtrueno_cuda.so");
    let status = Command::new(&nvcc)
        .args(["-shared", "-cudart", "shared", "-o"]).arg(&lib).args(&objs)
        // a source missing from SOURCES must fail HERE, not when the library is loaded
        .args(["-Xlinker", "-rpath,/usr/local/cuda/lib64", "-Xlinker", "--no-undefined"])
        .status()
        .expect("nvcc link step failed to start");
    assert!(status.success(), "linking libtrueno_cuda.so failed");

    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=trueno_cuda");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
}
