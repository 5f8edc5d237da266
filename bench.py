#!/usr/bin/env python
"""bench.py — the judged benchmark for the trueno hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline metric (BASELINE.json): f32 matmul TFLOP/s at 8192^2 — Matrix::matmul 8192x8192x8192
(config 2, "3xTF32 tcgen05").  One step = one pass of the hot path over one batch: C = A * B with
A, B resident in HBM (value) or in pinned host memory through the host-slice C-ABI call (e2e).
At N > 1 the path shards by output row blocks: every rank owns an 8192-row block
of A and C and a replica of B — weak scaling, no data-path collective.  The reduction half of the
BASELINE metric (dot / sum / argmax / norm_l2 GB/s on 2^30 f32, sliced across the ranks with an
NCCL all-reduce of the partials) and the config-5 row kernels are measured after the timed region
and reported under "secondary"; they are explanatory, not the headline value.

--impl reference times the CPU restatement of trueno's own AVX2 path (oracle/, all host threads,
the `parallel`-feature partitioning) on the same workload — the reference itself is Rust and
cannot be built in this image (no cargo), see DESIGN.md.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M = K = N = 8192
FLOP_PER_STEP = 2.0 * M * K * N
WORKLOAD = "Matrix::matmul 8192x8192x8192 f32 (BASELINE.json configs[1])"


line_holder: list[str] = []   # the JSON line, emitted by main() once fd 1 points at the real stdout again


def load_peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def splitmix_u01_torch(torch, seed: int, n: int, device):
    """x = u01(splitmix64(seed ^ idx)) — same generator as tests/kats.py, on device."""
    idx = torch.arange(n, dtype=torch.int64, device=device)
    z = (idx ^ seed) + (-7046029254386353131)            # 0x9E3779B97F4A7C15 as int64
    z = (z ^ (z >> 30 & 0x3FFFFFFFF)) * (-4658895280553007687)   # 0xBF58476D1CE4E5B9
    z = (z ^ (z >> 27 & 0x1FFFFFFFFF)) * (-7723592293110705685)  # 0x94D049BB133111EB
    z = z ^ (z >> 31 & 0x1FFFFFFFF)
    return ((z >> 40) & 0xFFFFFF).to(torch.float32) * (1.0 / (1 << 24))


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.rows, self.proc, self.dev = [], None, device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, region: tuple[float, float] | None = None, load_window: tuple[float, float] | None = None) -> dict:
        """Summarises the samples that fell inside the timed region; a region shorter than three
        sampling periods falls back to the surrounding window in which the same kernel was running
        back to back (warm-up + timed + roofline loops) and says so in `window`."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        window = "all samples"
        rows = [r for _, r in self.rows]
        if region:
            inside = [r for t, r in self.rows if region[0] <= t <= region[1]]
            if len(inside) >= 3:
                rows, window = inside, "timed region"
            elif load_window:
                rows = [r for t, r in self.rows if load_window[0] <= t <= load_window[1]] or rows
                window = "timed region < 3 samples: warm-up + timed + roofline loops (same kernel, back to back)"
        sm, smax, reasons, power = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                                  ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "window": window,
                "reasons": sorted(reasons)}


# ==================================================================================================
# reference arm: the CPU port of trueno's AVX2 matmul path, all host threads
# ==================================================================================================
def cpu_matmul_sample(rows: int, threads: int | None = None, reps: int = 1):
    """Times `rows` output rows of the 8192^3 product with the oracle's blocked AVX2 path
    (matmul_simd, src/matrix.rs:912-1401, `parallel` feature = 256-row blocks over all threads)."""
    import numpy as np
    import oracle
    orc = oracle.get()
    if threads:
        orc.set_threads(threads)
    rng = np.random.default_rng(0x5EED0001)
    A = rng.random((rows, K), dtype=np.float32)
    B = rng.random((K, N), dtype=np.float32)
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        orc.matmul_simd(A, B, rows, K, N, parallel=True)
        best = min(best, time.perf_counter() - t0)
    return 2.0 * rows * K * N / best / 1e12, best, orc.num_threads()


def run_reference(args) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    import oracle
    orc = oracle.get()
    cores = os.cpu_count() or 1
    orc.set_threads(cores)
    # bounded sample: one 256-row block (the unit the reference's rayon path schedules,
    # src/matrix.rs:962-1011) per host thread, at most the 32 blocks the full 8192-row product has;
    # same k and n as the full workload.
    rows = 256 * max(1, min(32, cores))
    rng = np.random.default_rng(0x5EED0001)
    A = rng.random((rows, K), dtype=np.float32)
    B = rng.random((K, N), dtype=np.float32)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        orc.matmul_simd(A, B, rows, K, N, parallel=True)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    per_step = sum(times) / len(times)
    value = 2.0 * rows * K * N / per_step / 1e12
    line = {
        "impl": "reference", "metric": "f32 matmul TFLOP/s (8192^2)", "value": value, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{rows} of 8192 output rows per step", "timing": "host clock"},
        "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": min(cores, orc.num_threads()), "kind": "port",
                         "sample": f"{rows}x{K}x{N} slice of the 8192^3 product per step, oracle matmul_simd "
                                   f"(AVX2 4x1 microkernel, 256-row blocks over OpenMP threads)"},
        "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line_holder.append(json.dumps(line))
    return 0


# ==================================================================================================
# our arm
# ==================================================================================================
def run_ours(args) -> int:
    import numpy as np
    import torch
    import torch.distributed as dist

    import trueno_b200 as trn
    from trueno_b200 import parallel as par

    rank, local_rank, world = par.init_distributed()
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    trn.check(trn.lib.trn_cuda_init(local_rank))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    L = trn.lib
    peaks = load_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bf16_peak = float(peaks.get("bf16_tflops", 1590.0))
    bf16_sustained = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured" if peaks else "fallback"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs: this rank's 8192-row block of A, the replica of B; U[0,1) from the counter generator
    a = splitmix_u01_torch(torch, 0x5EED0001 + 7919 * rank, M * K, dev).view(M, K)   # distinct row block per rank
    b = splitmix_u01_torch(torch, 0x5EED0002, K * N, dev).view(K, N)
    c = torch.empty(M, N, dtype=torch.float32, device=dev)

    def step_dev():
        trn.check(L.trn_matmul_f32_dev(a.data_ptr(), M, K, b.data_ptr(), K, N, c.data_ptr(), st))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)   # nvidia-smi start-up
    t_load0 = time.time()
    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()

    # ---- timed region: K steps, device events, barrier + sync on both sides, clocks sampled during.
    # The library brackets the dominant kernel of every call with its own CUDA events on the launching stream
    # (trn_profile_*: three event records per step, no host synchronisation), so the roofline numerator is
    # measured live INSIDE the timed region.
    launches0 = trn.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    L.trn_profile_enable(1)
    t_region0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    barrier()
    t_region1 = time.time()
    L.trn_profile_enable(0)
    elapsed_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = trn.launch_count() - launches0
    ms_per_step = elapsed_ms / args.steps
    value = FLOP_PER_STEP * world / (ms_per_step * 1e-3) / 1e12

    # ---- roofline of the dominant kernel: mean duration over the (last <= 64) steps of the timed region
    p_ms, k_ms = C.c_float(), C.c_float()
    trn.check(L.trn_profile_last_gemm(C.byref(p_ms), C.byref(k_ms)))
    kern_ms, pre_ms = [k_ms.value], [p_ms.value]
    torch.cuda.synchronize()
    clocks = sampler.stop((t_region0, t_region1), (t_load0, time.time())) if rank == 0 else {}
    kernel_ms = sum(kern_ms) / len(kern_ms)
    achieved = FLOP_PER_STEP / (kernel_ms * 1e-3) / 1e12
    tf32x3_peak = bf16_peak / 6.0   # TF32 runs at half the bf16 rate; 3 TF32 MMAs per f32 product
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {
        "bound": "tensor", "kernel": "gemm_tf32x3_pair_kernel (tcgen05.mma.cta_group::2 kind::tf32, 3xTF32)", "achieved": achieved,
        "peak": tf32x3_peak, "unit": "TFLOP/s", "frac": achieved / tf32x3_peak, "traffic": traffic,
        "peak_note": f"{peak_src} bf16 burst {bf16_peak} TF/s / 2 (TF32 rate) / 3 (3xTF32); sustained figure "
                     f"{bf16_sustained / 6.0:.1f}",
        "frac_of_sustained": achieved / (bf16_sustained / 6.0),
        "kernel_ms": kernel_ms, "prepass_ms": sum(pre_ms) / len(pre_ms),
        "algorithmic_flop_per_launch": FLOP_PER_STEP,
    }

    # ---- e2e: the host-slice C-ABI call with pinned HOST buffers; H2D of A and B and D2H of C inside
    e2e = None
    if True:
        ha, hb, hc = trn.pinned_empty(M * K), trn.pinned_empty(K * N), trn.pinned_empty(M * N)
        ha[:] = a.view(-1).cpu().numpy()
        hb[:] = b.view(-1).cpu().numpy()
        e2e_steps = max(2, min(args.steps, 5))

        def step_host():
            trn.check(L.trn_matmul_f32(ha.ctypes.data, M, K, hb.ctypes.data, K, N, hc.ctypes.data))

        step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        if world > 1:
            dist.barrier()
        e2e = {"value": FLOP_PER_STEP * world * e2e_steps / dt / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": 4 * (M * K + K * N), "d2h_bytes_per_step": 4 * M * N,
               "ms_per_step": dt / e2e_steps * 1e3, "steps": e2e_steps,
               "api": "trn_matmul_f32 (host slices, pinned)", "checksum": float(hc[:1024].sum())}
        del ha, hb, hc

    # ---- secondary: reductions on 2^30 f32 sliced over the ranks (+ all-reduce), config-5 row kernels
    secondary = []
    del a, b, c
    torch.cuda.empty_cache()

    def timed(fn, iters=20):
        for _ in range(3):
            fn()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(iters):
            fn()
        s1.record(stream)
        barrier()
        return max_over_ranks(s0.elapsed_time(s1)) / iters

    n_total = 1 << 30
    sh = par.shard_range(n_total, rank, world, align=4)
    x = (splitmix_u01_torch(torch, 0x5EED0005 + rank, sh.count, dev) * 2 - 1)
    y = (splitmix_u01_torch(torch, 0x5EED0006 + rank, sh.count, dev) * 2 - 1)
    # N > 1: the exchange step runs INSIDE the slice kernel over NVLink peer memory (csrc/peer.cu); the
    # kernel + NCCL variant is timed beside it for comparison
    comm = par.PeerComm() if world > 1 else None
    variants = [("", comm)] + ([(" [kernel + NCCL exchange]", None)] if world > 1 else [])
    for suffix, cm in variants:
        vx, vy = par.ShardedVector(x, sh, cm), par.ShardedVector(y, sh, cm)
        for name, fn, nbytes in (("dot", lambda: vx.dot(vy), 8 * n_total), ("sum", lambda: vx.sum(), 4 * n_total),
                                 ("argmax", lambda: vx.argmax(), 4 * n_total), ("norm_l2", lambda: vx.norm_l2(), 4 * n_total)):
            ms = timed(fn)
            gbs = nbytes / ms / 1e6
            secondary.append({"metric": f"{name} 2^30 f32 GB/s{suffix}", "value": gbs, "ms": ms,
                              "roofline_frac": gbs / (hbm_peak * world), "bound": "hbm",
                              "exchange": "none (1 GPU)" if world == 1 else ("fused P2P over NVLink" if cm else "NCCL")})
    del x, y, vx, vy
    if comm is not None:
        torch.cuda.synchronize()
        dist.barrier()
    rows_total, cols = 4096, 32000
    rsh = par.shard_range(rows_total, rank, world)
    logits = torch.randn(rsh.count, cols, device=dev) * 4
    out = torch.empty_like(logits)
    for name, fn in (("softmax", L.trn_softmax_rows_f32_dev), ("log_softmax", L.trn_log_softmax_rows_f32_dev)):
        ms = timed(lambda: trn.check(fn(logits.data_ptr(), out.data_ptr(), rsh.count, cols, st)))
        gbs = 8.0 * rows_total * cols / ms / 1e6
        secondary.append({"metric": f"{name} 4096x32000 f32 GB/s", "value": gbs, "ms": ms,
                          "roofline_frac": gbs / (hbm_peak * world), "bound": "hbm"})
    ms = timed(lambda: trn.check(L.trn_gelu_f32_dev(logits.data_ptr(), logits.numel(), out.data_ptr(), st)))
    gbs = 8.0 * rows_total * cols / ms / 1e6
    secondary.append({"metric": "gelu 4096x32000 f32 GB/s", "value": gbs, "ms": ms,
                      "roofline_frac": gbs / (hbm_peak * world), "bound": "hbm"})
    del logits, out

    # ---- secondary: BASELINE config 3 (Q K^T, batch 8 x heads 32, seq 2048, head_dim 128, heads sharded over the
    # ranks, no collective) and the same heads through the fused attention kernel (scores never leave the SM)
    B3, H3, S3, D3 = 8, 32, 2048, 128
    hsh = par.shard_range(B3 * H3, rank, world)
    q3 = torch.randn(hsh.count * S3 * D3, device=dev)
    k3 = torch.randn(hsh.count * S3 * D3, device=dev)
    v3 = torch.randn(hsh.count * S3 * D3, device=dev)
    kt3 = k3.view(hsh.count, S3, D3).transpose(1, 2).contiguous().view(-1)
    c3 = torch.empty(hsh.count * S3 * S3, device=dev)
    ms = timed(lambda: trn.check(L.trn_batched_matmul_4d_f32_dev(q3.data_ptr(), q3.numel(), kt3.data_ptr(), kt3.numel(),
                                                                 c3.data_ptr(), 1, hsh.count, S3, D3, S3, st)), iters=10)
    secondary.append({"metric": "batched_matmul_4d Q K^T 8x32x2048x128 TFLOP/s", "value": 2.0 * B3 * H3 * S3 * S3 * D3 / ms / 1e9,
                      "ms": ms, "roofline_frac": 2.0 * B3 * H3 * S3 * S3 * D3 / ms / 1e9 / (tf32x3_peak * world), "bound": "tensor"})
    del c3, kt3
    o3 = torch.empty_like(q3)
    for causal in (0, 1):
        flop = 4.0 * B3 * H3 * S3 * S3 * D3 * (0.5 if causal else 1.0)
        ms = timed(lambda: trn.check(L.trn_attention_f32_dev(q3.data_ptr(), q3.numel(), k3.data_ptr(), k3.numel(), v3.data_ptr(),
                                                             v3.numel(), o3.data_ptr(), hsh.count, S3, D3, 1.0 / D3 ** 0.5, causal, st)),
                   iters=10)
        secondary.append({"metric": f"fused attention 256 heads x 2048 x 128{' causal' if causal else ''} TFLOP/s",
                          "value": flop / ms / 1e9, "ms": ms, "roofline_frac": flop / ms / 1e9 / (tf32x3_peak * world),
                          "bound": "tensor"})
    del q3, k3, v3, o3

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on the host cores, bounded sample
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        rows = 256 * max(1, min(32, cores))
        tf, secs, nthreads = cpu_matmul_sample(rows, threads=cores, reps=2)
        tf1, secs1, _ = cpu_matmul_sample(256, threads=1, reps=1)
        cpu_baseline = {"value": tf, "unit": "TFLOP/s", "cores": min(cores, nthreads), "kind": "port",
                        "sample": f"{rows} of 8192 output rows of the same 8192^3 product ({secs:.2f} s), oracle "
                                  f"matmul_simd with the `parallel`-feature partitioning over all host threads",
                        "single_thread_value": tf1,
                        "single_thread_sample": f"256 output rows, 1 thread ({secs1:.2f} s) — `cargo bench` default features"}

    # HBM-bound lines are reported against the measured copy bandwidth AND the nominal 8 TB/s
    for entry in secondary:
        if entry.get("bound") == "hbm":
            entry["roofline_frac_nominal_8TBs"] = entry["value"] / (8000.0 * world)

    if rank == 0:
        line = {
            "metric": "f32 matmul TFLOP/s (8192^2)", "value": value, "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "engine": "tcgen05 cta_group::2 3xTF32, two-level accumulation",
                       "sharding": f"C/A row blocks of {M} rows per GPU, B replicated, no collective" if world > 1 else "single GPU",
                       "l2": "inputs (2 x 256 MiB + 1 GiB split scratch) exceed the 126 MB L2",
                       "generator": "u01(splitmix64(seed ^ idx)), seeds 0x5EED0001/2"},
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "clocks": clocks, "secondary": secondary,
            "device": trn.device_info()["name"],
        }
        line_holder.append(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: libraries that write to fd 1 behind Python's back (NCCL prints its
    # version banner there) are pointed at stderr for the duration of the run.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line_holder.clear()
        rc = run_reference(args) if args.impl == "reference" else run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for line in line_holder:
        print(line)
    sys.stdout.flush()
    return rc


if __name__ == "__main__":
    sys.exit(main())
