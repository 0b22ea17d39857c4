#!/usr/bin/env python3
"""Benchmark of the hot path BASELINE.json names: BLS12-381 G1 Pippenger MSM, points/s at 2^26.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log-n 26]

A "step" is one full MSM (scalars -> Jacobian result) over one batch of synthetic input:
bases P_i = [start + i]G synthesized on the device once (they are the constant half of the
workload, like a Groth16 proving key), scalars uniform in [0, r).

* ``value``  : points/s with the scalars already resident in HBM (ozl_msm_device_async).
* ``e2e``    : same metric through the reference-facing C-ABI call ``ozl_msm`` with HOST (pinned)
               scalars: H2D of the step's scalars and D2H of the result inside the timed region.
* ``roofline``: for the dominant kernel (k_accumulate), algorithmic bytes (128 B per point: 32 B
               scalar + 96 B affine base) / its CUDA-event duration vs the measured HBM peak; the
               companion ``fma_pipe`` block gives the binding roofline (field multiplications/s
               vs the measured multiplier peak).
* ``cpu_baseline`` / ``--impl reference``: the C++ restatement of ark-ec 0.3.0's
               VariableBaseMSM (oracle/c/ozl_oracle.cpp; the Rust reference cannot be built in this
               image) on the box's host cores, windows in parallel like rayon, bounded sample.

N > 1 (torchrun): weak scaling -- every rank owns 2^log_n contiguous points of a global
N * 2^log_n MSM, runs the full single-GPU pipeline, and the 144-byte Jacobian partials are
exchanged with one NCCL all-gather and summed (EC addition is not an NCCL reduction op).
"""
from __future__ import annotations

import argparse
import math
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R381 = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
# DRAM traffic of the dominant kernel from the committed ncu --set full capture of this exact
# configuration, keyed by (log2 n, window bits, base copies); other configurations report null.
NCU_TRAFFIC = {(26, 22, 12): 157.800895e9 + 1.624909e9}
MUL_EQUIV_PER_MADD = 6 + 1.5 + 2 * 222.0 / 288.0    # 6 plain products, one fused pair (3/2), two dedicated squarings
METRIC = "bls12_381_g1_msm_points_per_sec"
UNIT = "points/s"


def _p2(v: int) -> str:
    return f"2^{v.bit_length() - 1}" if v > 0 and v & (v - 1) == 0 else str(v)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi in the background during the timed region)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# synthetic scalars: uniform in [0, r) by mask-and-reject, generated on the device
# ------------------------------------------------------------------------------------------
def device_scalars(n: int, modulus: int, seed: int, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    bits = modulus.bit_length()
    top_mask = (1 << (bits - 192)) - 1
    mod = [(modulus >> (64 * i)) & ((1 << 64) - 1) for i in range(4)]
    out = torch.empty((n, 4), dtype=torch.int64, device=device)
    filled = 0
    sign = 1 << 63

    def to_signed_key(x):  # order-preserving map u64 -> i64 for comparisons
        return x ^ torch.tensor(-sign, dtype=torch.int64, device=device)

    mod_keys = [((m ^ sign) - (1 << 64)) if (m ^ sign) >= sign else (m ^ sign) for m in mod]
    while filled < n:
        m = min(n - filled, 1 << 24)
        k = int(m * 1.25) + 1024
        hi = torch.randint(0, 1 << 32, (k, 4), dtype=torch.int64, device=device, generator=g)
        lo = torch.randint(0, 1 << 32, (k, 4), dtype=torch.int64, device=device, generator=g)
        cand = (hi << 32) | lo
        del hi, lo
        cand[:, 3] &= top_mask
        lt = torch.zeros(k, dtype=torch.bool, device=device)
        eq = torch.ones(k, dtype=torch.bool, device=device)
        for limb in (3, 2, 1, 0):
            key = to_signed_key(cand[:, limb])
            lt |= eq & (key < mod_keys[limb])
            eq &= key == mod_keys[limb]
        good = cand[lt][:m]
        out[filled:filled + good.shape[0]] = good
        filled += good.shape[0]
    return out


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the C++ restatement of ark's MSM on host cores
# ------------------------------------------------------------------------------------------
def cpu_msm_throughput(log_sample: int, steps: int, warmup: int, seed: int = 1):
    from oracle import cbind
    from tests.util import random_scalars
    n = 1 << log_sample
    bases = cbind.bases_seq("bls12_381_g1", 1, n)
    scalars = random_scalars(n, R381, seed)
    c = cbind.window_bits(n)
    windows = (255 + c - 1) // c
    # one thread per window like rayon's cfg_into_iter!(window_starts); when the box has one core fewer than
    # windows (17 windows on 16 cores) still start them all, so the last window does not cost a second round
    hw = cbind.hw_threads()
    threads = windows if hw >= windows - 1 else max(1, min(windows, hw))
    for _ in range(warmup):
        cbind.msm("bls12_381_g1", bases, scalars, threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        cbind.msm("bls12_381_g1", bases, scalars, threads=threads)
    dt = (time.perf_counter() - t0) / steps
    return n / dt, dt, threads, c


def _stdout_to_stderr():
    """Point file descriptor 1 at stderr for the duration of a run: libraries that write to stdout
    from native code (NCCL prints its version line there when NCCL_DEBUG is set) must not share it
    with the ONE JSON line the driver parses.  Returns the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def _restore_stdout(saved):
    sys.stdout.flush()
    os.dup2(saved, 1)
    os.close(saved)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    log_sample = args.cpu_log_n
    v, dt, threads, c = cpu_msm_throughput(log_sample, args.steps, args.warmup)
    sample = (f"2^{log_sample} of the 2^{args.log_n} points per step (ark window rule c={c}, "
              f"{threads} threads = one per window up to nproc)")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (Montgomery, 381-bit)", "data": "synthetic",
        "config": {"workload": f"BLS12-381 G1 MSM 2^{args.log_n} (CPU arm runs a bounded 2^{log_sample} sample per step)",
                   "what": "C++17 restatement of ark-ec 0.3.0 VariableBaseMSM (reference Rust crate cannot be built here)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def msm_measure(args, world, rank, local_rank, n, steps, warmup, precompute, full):
    """One MSM configuration: n points per rank (bases [start, start + n) of a world * n point MSM),
    device-timed steps, the end-to-end C-ABI leg with host scalars, verification against the known
    discrete logs.  full = also the un-precomputed reference point, the pipelined submit variant and the
    per-stage timers.  Returns a dict; every device buffer is released before returning."""
    import torch
    import torch.distributed as dist
    import openzl_b200 as ozl
    dev = torch.device("cuda", local_rank)
    curve = ozl.BLS12_381_G1
    ctx = ozl.Context(local_rank)
    ctx.use_torch_stream()
    # precompute 0 = measured best on B200 (profiles/msm_sweep_full_precompute_r01.jsonl): a FULL set
    # of shifted copies -- one per window, so a single bucket set and no Horner tail -- whenever it
    # fits in ~90 GB (2^26 points: c = 22, 12 copies, 72 GiB); otherwise as many copies as fit in ~80 GB
    # under the 31-bit index limit (n * copies < 2^31).  The planner picks the window width for the factor.
    if precompute == 0:
        if n * 96 * 13 <= 90e9:
            precompute = 32                          # >= the number of windows: one copy per window
        else:
            precompute = max(1, min(4, int(80e9 // (n * 96)), ((1 << 31) - 1) // n))
    if args.window_bits:
        ctx.set_window_bits(args.window_bits)
    start = 1 + rank * n
    bases = ctx.generate_bases(curve, start, n)
    value_plain, tpre = None, 0.0
    d_scalars = device_scalars(n, R381, seed=1234 + rank, device=dev)
    h_scalars = torch.empty((n, 4), dtype=torch.int64, pin_memory=True)
    h_scalars.copy_(d_scalars)
    d_out = torch.zeros(18, dtype=torch.int64, device=dev)

    # N > 1: the library's own NCCL communicator (ozl_comm_*): shard MSM, one ncclAllGather of the
    # 144-byte partials and the device-side sum, all on the context's stream; torch.distributed is
    # only the rendezvous, the barrier and the max-over-ranks of the timings.
    comm = None
    if world > 1:
        from openzl_b200.multi_gpu import Comm
        comm = Comm.from_torch(ctx)

    def step():
        if comm is not None:
            comm.msm_sharded_device(bases, d_scalars.data_ptr(), n, d_out.data_ptr())
        else:
            bases.msm_device(d_scalars.data_ptr(), n, d_out.data_ptr())

    def combine():
        return d_out.cpu().numpy().view(np.uint64)      # the combined point on every rank when N > 1

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if precompute > 1:
        if full:
            # one untimed reference point without precomputed copies (same kernels, W bucket sets)
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for _ in range(2):
                step()
            p1.record()
            torch.cuda.synchronize()
            value_plain = world * n / (p0.elapsed_time(p1) / 2 * 1e-3)
        tpre = time.perf_counter()
        bases.precompute(precompute)
        ctx.synchronize()
        tpre = time.perf_counter() - tpre
    for _ in range(warmup):
        step()
        combine()
    torch.cuda.synchronize()

    # --- timed region: exactly K steps, device events, max over ranks ------------------------
    sampler = ClockSampler(local_rank)
    ctx.enable_timing(full)
    stage_acc = {}
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches_before = ctx.launch_count
    e0.record()
    for _ in range(steps):
        step()
        result = combine()
        if full:
            for name, ms, _l in ctx.stage_times():       # synchronizes the stream (once per ~0.3 s step)
                stage_acc.setdefault(name, []).append(ms)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches_timed = ctx.launch_count - launches_before
    clocks = sampler.stop()
    ctx.enable_timing(False)
    ms_per_step = max_over_ranks(e0.elapsed_time(e1)) / steps
    value = world * n / (ms_per_step * 1e-3)

    # --- end-to-end through the C ABI with host buffers --------------------------------------
    out_host = np.zeros(18, dtype=np.uint64)
    torch.cuda.synchronize()

    def e2e_call():
        if comm is not None:      # ozl_msm_sharded: host scalars of the shard in, combined point out
            ctx._check(ctx._lib.ozl_msm_sharded(ctx._h, comm._h, bases.handle, h_scalars.data_ptr(), n, out_host.ctypes.data),
                       "ozl_msm_sharded")
        else:
            bases.msm_host_ptr(h_scalars.data_ptr(), n, out_host)

    e2e_call()                                              # warm the staging buffers of the batched path
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_call()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / steps)
    e2e_value = world * n / e2e_s
    e2e_combined = out_host.copy()                          # with N > 1 ozl_msm_sharded returns the COMBINED point
    pipelined = None
    pageable = None
    if full and world == 1:
        # the same call with PAGEABLE host scalars (what a Rust Vec is): seven point-range batches, the blocking
        # copies of batch j + 1 run while batch j is accumulated
        hp = np.empty((n, 4), dtype=np.uint64)
        hp[:] = h_scalars.numpy().view(np.uint64)
        out_p = np.zeros(18, dtype=np.uint64)
        bases.msm_host_ptr(hp.ctypes.data, n, out_p)
        t0 = time.perf_counter()
        for _ in range(max(2, steps // 4)):
            bases.msm_host_ptr(hp.ctypes.data, n, out_p)
        pg_s = (time.perf_counter() - t0) / max(2, steps // 4)
        same = bool((ctx.jacobian_to_affine(curve, out_p)[0] == ctx.jacobian_to_affine(curve, e2e_combined)[0]).all())
        pageable = {"value": n / pg_s, "ms_per_step": pg_s * 1e3, "results_identical": same,
                    "api": "ozl_msm with pageable (malloc) host scalars, 7 batches"}
        del hp
    if full:
        # pipelined variant: K back-to-back submissions, H2D of step i+1 overlapping the kernels of step i
        outs = torch.zeros((steps, 18), dtype=torch.int64).pin_memory()
        bases.msm_submit(h_scalars.data_ptr(), n, outs[0].data_ptr())
        ctx.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            bases.msm_submit(h_scalars.data_ptr(), n, outs[i].data_ptr())
        ctx.synchronize()
        pipe_s = max_over_ranks((time.perf_counter() - t0) / steps)
        # Jacobian representatives depend on the (atomic) order of points inside a bucket: compare affine forms
        pipe_affine = [ctx.jacobian_to_affine(curve, o)[0] for o in outs.numpy().view(np.uint64)]
        d_part = torch.zeros(18, dtype=torch.int64, device=dev)
        bases.msm_device(d_scalars.data_ptr(), n, d_part.data_ptr())     # this rank's partial
        ctx.synchronize()
        ref_affine = ctx.jacobian_to_affine(curve, d_part.cpu().numpy().view(np.uint64))[0]
        pipelined = {"value": world * n / pipe_s, "ms_per_step": pipe_s * 1e3,
                     "results_identical": bool(all((a == ref_affine).all() for a in pipe_affine)),
                     "api": "ozl_msm_submit x K + ozl_ctx_synchronize (H2D of step i+1 overlaps step i)"}

    # --- verification outside the timed region: sum s_i [start+i]G == [sum s_i (start+i)]G ----
    verified = None
    if not args.no_verify:
        from oracle import cbind
        hs = h_scalars.numpy().view(np.uint64)
        k = cbind.dot_mod_r("bls12_381_fr", hs, np.arange(start, start + n, dtype=np.uint64))
        if world > 1:
            ks = [None] * world
            dist.all_gather_object(ks, k)
            k = sum(ks) % R381
        exp, _ = cbind.to_affine("bls12_381_g1", cbind.gen_mul("bls12_381_g1", k))
        got, _ = ctx.jacobian_to_affine(curve, result)           # device-scalar path (timed `value`)
        got2, _ = ctx.jacobian_to_affine(curve, e2e_combined)    # host-scalar path (timed `e2e`)
        verified = bool((got == exp).all() and (got2 == exp).all())

    info = bases.info(n)
    res = dict(n=n, value=value, ms_per_step=ms_per_step, e2e_value=e2e_value, e2e_s=e2e_s, pipelined=pipelined,
               verified=verified, clocks=clocks, pageable=pageable, launches_timed=int(launches_timed), info=info, tpre=tpre,
               value_plain=value_plain, stage_acc=stage_acc, precompute=precompute,
               mul_peak=ctx.bench_field_mul(0, 4000) if full else None)
    if comm is not None:
        comm.close()
    bases.free()
    del d_scalars, h_scalars, d_out
    ctx.close()
    torch.cuda.empty_cache()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    n = 1 << args.log_n
    strong_only = bool(args.total_log_n)
    if strong_only:
        # strong scaling as the headline (BASELINE config 5): a fixed 2^total_log_n-point MSM split by point range
        assert (1 << args.total_log_n) % world == 0
        n = (1 << args.total_log_n) // world

    m = msm_measure(args, world, rank, local_rank, n, args.steps, args.warmup, args.precompute, full=True)

    # BASELINE config 5 beside the weak-scaling headline: ONE 2^28-point MSM split by point range over the
    # ranks (strong scaling), so that the driver's 1/2/4/8 runs record it with verification and e2e.
    strong = None
    if not strong_only and args.strong_log_n and (1 << args.strong_log_n) % world == 0 and args.log_n == 26:
        ns = (1 << args.strong_log_n) // world
        sm = msm_measure(args, world, rank, local_rank, ns, args.strong_steps, 1, 0, full=False)
        si = sm["info"]
        strong = {"metric": METRIC, "value": sm["value"], "unit": UNIT, "scaling": "strong", "n_gpus": world,
                  "total_points": 1 << args.strong_log_n, "points_per_gpu": ns, "steps": args.strong_steps, "warmup": 1,
                  "ms_per_step": sm["ms_per_step"],
                  "config": {"workload": f"BLS12-381 G1 MSM 2^{args.strong_log_n} sharded by point range across {world} GPU(s), "
                                         f"window c={si['c']} ({si['windows']} windows in {si['bucket_sets']} bucket sets), {si['factor']} base copies"},
                  "e2e": {"value": sm["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": ns * 32, "d2h_bytes_per_step": 144,
                          "ms_per_step": sm["e2e_s"] * 1e3, "api": "ozl_msm_sharded" if world > 1 else "ozl_msm"},
                  "gpu_launches": sm["launches_timed"], "verified_vs_known_dlog": sm["verified"], "clocks": sm["clocks"]}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    stage_acc, info = m["stage_acc"], m["info"]
    ms_per_step = m["ms_per_step"]
    acc = float(np.mean(stage_acc.get("accumulate", [float("nan")])))
    alg_bytes = n * 128.0
    achieved = alg_bytes / (acc * 1e-3) / 1e9
    c, W, Wc = info["c"], info["windows"], info["bucket_sets"]
    madds_per_s = n * W / (acc * 1e-3)
    line = {
        "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong_only else "weak", "vs_baseline": None,
        "dtype": "u32 limbs (Montgomery, 381-bit)", "data": "synthetic",
        "config": {"workload": f"BLS12-381 G1 Pippenger MSM, {_p2(n)} points per GPU ({_p2(world * n)} total), window c={c} ({W} windows in {Wc} bucket sets, signed digits)",
                   "bases": "P_i=[start+i]G generated on device, resident (constant across steps like a proving key)",
                   "precompute_factor": info["factor"], "precompute_s": m["tpre"],
                   "value_without_precompute": m["value_plain"],
                   "scalars": "uniform in [0,r), mask-and-reject", "l2": "inputs (8 GiB per step) are far larger than L2; no flush needed",
                   "parallelism": f"point-range shards x{world}, one ncclAllGather of 144 B partials + device-side sum on the MSM's stream (ozl_msm_sharded)" if world > 1 else "single GPU"},
        "clocks": m["clocks"],
        "e2e": {"value": m["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": 144,
                "ms_per_step": m["e2e_s"] * 1e3,
                "api": ("ozl_msm_sharded" if world > 1 else "ozl_msm") + " (C ABI, pinned host scalars, bases resident; the scalars cross PCIe in "
                       "point-range batches that overlap the accumulation of earlier batches)",
                "pipelined": m["pipelined"], "pageable": m["pageable"]},
        "gpu_launches": m["launches_timed"],
        "roofline": {"kernel": "k_accumulate", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": NCU_TRAFFIC.get((int(math.log2(n)), c, info["factor"])),
                     "traffic_unit": "bytes per launch",
                     "traffic_source": "bytes per launch, ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r02c_ncu_k_accumulate_2p26.csv (one gather of 96 B per point per window, at 64-byte DRAM granularity)",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": acc,
                     "share_of_step": acc / ms_per_step},
        "fma_pipe": {"note": "binding roofline: 381-bit Montgomery multiplications on the integer fma pipe.  A mixed addition is 8 products + "
                             "2 squarings; with y3 = r(q - x3) - y p3 fused into one reduction (3 N^2 instead of 4 N^2 wide multiplies) and the "
                             "dedicated squaring (222 instead of 288) that is 9.04 multiplication-equivalents; ncu of the same kernel: "
                             "sm__pipe_fmaheavy_cycles_active 90.7 % (profiles/r02c_ncu_k_accumulate_2p26.csv)",
                     "mul_equivalents_per_mixed_add": MUL_EQUIV_PER_MADD,
                     "field_mul_per_s": madds_per_s * MUL_EQUIV_PER_MADD, "mixed_adds_per_s": madds_per_s,
                     "measured_mul_peak_per_s": m["mul_peak"], "frac": madds_per_s * MUL_EQUIV_PER_MADD / m["mul_peak"]},
        "stages_ms": {k: float(np.mean(v)) for k, v in stage_acc.items()},
        "verified_vs_known_dlog": m["verified"],
    }
    if strong is not None:
        line["strong_2p28"] = strong
    if world == 1 and not args.no_cpu_baseline:
        v, dt, threads, cc = cpu_msm_throughput(args.cpu_baseline_log_n, 1, 0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"one MSM over 2^{args.cpu_baseline_log_n} of the workload's points (ark window c={cc}, "
                                          f"{(255 + cc - 1) // cc} windows; 2^26 would use c=19, 14 windows), {dt:.1f} s"}
    # the second BASELINE metric and config 3, in the same line (single-GPU runs only: proofs and NTTs are replicas across GPUs)
    if world == 1 and not args.no_groth16:
        try:
            line["groth16"] = groth16_block(args, local_rank)
        except Exception as exc:      # never lose the MSM line to the second metric
            line["groth16"] = {"error": repr(exc)}
    if world == 1 and not args.no_ntt:
        try:
            line["ntt"] = ntt_block(args, 24, local_rank, 10)
        except Exception as exc:
            line["ntt"] = {"error": repr(exc)}
        try:
            line["g2_msm"] = g2_msm_block(args, 22, local_rank)
        except Exception as exc:
            line["g2_msm"] = {"error": repr(exc)}
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def groth16_cpu_baseline(pk, z_canonical: np.ndarray, log_n: int):
    """One sample of the reference-shaped CPU prover core on the host cores: the 7 NTTs of
    R1CStoQAP::witness_map (serial radix-2, as ark without `parallel` inside one transform) and the 5
    MSMs of create_proof (C++ restatement of ark's Pippenger, one thread per window), on the SAME
    proving key (bases read back from the device) and the same assignment.  Mat-vec and proof assembly
    are left out, which flatters the CPU."""
    import openzl_b200 as ozl
    from openzl_b200.context import Bases
    from oracle import cbind
    r254 = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    from tests.util import random_scalars
    n = 1 << log_n
    hw = cbind.hw_threads()
    t_ntt = 0.0
    x = random_scalars(n, r254, 11)
    for inverse, coset in ((True, False),) * 3 + ((False, True),) * 3 + ((True, True),):
        t0 = time.perf_counter()
        cbind.ntt("bn254_fr", x, inverse=inverse, coset=coset)
        t_ntt += time.perf_counter() - t0
    t_msm = 0.0
    ni = pk.r1cs.n_instance
    per = {}
    for name, curve, cname, sl in (("a", ozl.BN254_G1, "bn254_g1", slice(0, None)), ("b1", ozl.BN254_G1, "bn254_g1", slice(0, None)),
                                   ("b2", ozl.BN254_G2, "bn254_g2", slice(0, None)), ("l", ozl.BN254_G1, "bn254_g1", slice(ni, None)),
                                   ("h", ozl.BN254_G1, "bn254_g1", None)):
        h, cnt = pk.query_handles[name]
        bases = Bases(pk.ctx, h, curve, cnt).download()
        inf = np.packbits(~bases.any(axis=1), bitorder="little")
        sc = z_canonical[sl] if sl is not None else random_scalars(cnt, r254, 12)
        c = cbind.window_bits(cnt)
        windows = (254 + c - 1) // c
        threads = windows if hw >= windows - 1 else max(1, min(windows, hw))
        t0 = time.perf_counter()
        cbind.msm(cname, bases, np.ascontiguousarray(sc[:cnt]), inf=inf, threads=threads)
        per[name] = time.perf_counter() - t0
        t_msm += per[name]
        del bases
    total = t_ntt + t_msm
    return {"value": 1.0 / total, "unit": "proofs/s", "cores": min(hw, 17), "kind": "port",
            "sample": f"one proof's 7 NTTs ({t_ntt:.2f} s, serial radix-2) + 5 MSMs ({t_msm:.2f} s, one thread per window; "
                      f"a {per['a']:.2f} b1 {per['b1']:.2f} b2 {per['b2']:.2f} l {per['l']:.2f} h {per['h']:.2f}) on the same key; "
                      "mat-vec and assembly excluded"}


def groth16_block(args, device_index: int = 0):
    """Second headline metric: Groth16 proofs/s at 2^20 constraints (BN254, Poseidon hash chain).
    A step is one `ozl_groth16_prove`: host witness in (H2D inside the call), device witness map
    (3 SpMV + 7 NTT + pointwise) + 4 G1 MSMs + 1 G2 MSM + assembly, three proof points out.  The last
    proof is VERIFIED at this size outside the timed region, with real pairings, twice: by the product's
    host verifier (openzl_b200.pairing, ate) and by the oracle's (oracle/pairing.py, Tate)."""
    import random
    import torch
    import openzl_b200 as ozl
    from openzl_b200.circuits import PoseidonChain
    from openzl_b200.groth16 import Groth16, Trapdoor, ints_to_limbs, PAIRINGS
    p = PAIRINGS["bn254"]["r"]
    links = args.links
    t0 = time.perf_counter()
    ch = PoseidonChain(links)
    r1 = ch.r1cs()
    z = ch.assignment(1234567, 7654321)
    z_m = ints_to_limbs(z, p, mont=True)
    t_circuit = time.perf_counter() - t0
    ctx = ozl.Context(device_index)
    rnd = random.Random(2026)
    td = Trapdoor(*[rnd.randrange(2, p) for _ in range(5)])
    if args.window_bits:
        ctx.set_window_bits(args.window_bits)
    t0 = time.perf_counter()
    pk, vk = Groth16.compile(ctx, "bn254", r1, td, precompute=args.g16_precompute)
    t_setup = time.perf_counter() - t0
    zt = torch.from_numpy(z_m.view(np.int64)).pin_memory()
    z_pinned = zt.numpy().view(np.uint64)
    r, s = rnd.randrange(p), rnd.randrange(p)
    steps, warmup = args.g16_steps, max(3, args.warmup)
    for _ in range(warmup):
        proof = Groth16.prove_with_randomness(pk, z_pinned, r, s)
    sampler = ClockSampler(device_index)
    launches0 = ctx.launch_count
    torch.cuda.synchronize()
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(steps):
        proof = Groth16.prove_with_randomness(pk, z_pinned, r, s)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    # stage times from separate, untimed proofs (the per-stage events serialise nothing but add host work)
    ctx.enable_timing(True)
    stage_acc = {}
    for _ in range(3):
        Groth16.prove_with_randomness(pk, z_pinned, r, s)
        per = {}
        for name, ms, _l in ctx.stage_times():
            per[name] = per.get(name, 0.0) + ms
        for k, v in per.items():
            stage_acc.setdefault(k, []).append(v)
        timeline = [[name, round(t0_, 3), round(t1_, 3)] for name, t0_, t1_, _l in ctx.stage_spans()]
    ctx.enable_timing(False)
    stages = {k: float(np.mean(v)) for k, v in stage_acc.items()}
    # --- verification AT THIS SIZE, outside the timed region -----------------------------------
    verified, verify_s = None, None
    if not args.no_verify:
        t0 = time.perf_counter()
        ok_product = Groth16.verify(vk, [z[1]], proof) and not Groth16.verify(vk, [(z[1] + 1) % p], proof)
        from oracle import pairing as opair
        from openzl_b200 import pairing as ppair
        E = ppair.ENGINES["bn254"]
        ic_pts = [ppair.g1_from_limbs(E, row) for row in vk.gamma_abc_g1]
        o_g2 = lambda limbs: ppair.g2_from_limbs(E, limbs)      # same ((x0, x1), (y0, y1)) tuples as the oracle's curves
        ok_oracle = opair.groth16_verify("bn254", ppair.g1_from_limbs(E, vk.alpha_g1), o_g2(vk.beta_g2), o_g2(vk.gamma_g2),
                                         o_g2(vk.delta_g2), ic_pts, [z[1]],
                                         (ppair.g1_from_limbs(E, proof.a), o_g2(proof.b), ppair.g1_from_limbs(E, proof.c)))
        verified = bool(ok_product and ok_oracle)
        verify_s = time.perf_counter() - t0
    # throughput with several independent provers in flight (one context + proving key each, one host
    # thread each; ctypes releases the GIL): fills the latency-bound tails of one proof with another's work
    conc = None
    if args.concurrency > 1:
        import threading
        workers = [(ctx, pk)]
        for _ in range(args.concurrency - 1):
            c2 = ozl.Context(device_index)
            workers.append((c2, Groth16.compile(c2, "bn254", r1, td, precompute=args.g16_precompute)[0]))
        def run(w, k):
            for _ in range(k):
                Groth16.prove_with_randomness(w[1], z_pinned, r, s)
        for w in workers:
            run(w, 2)
        torch.cuda.synchronize()
        k = max(steps, 4)
        ths = [threading.Thread(target=run, args=(w, k)) for w in workers]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        torch.cuda.synchronize()
        conc = {"provers": args.concurrency, "proofs_per_s": args.concurrency * k / (time.perf_counter() - t0)}
        for w in workers[1:]:
            w[1].free()
            w[0].close()
    cpu = None
    if not args.no_cpu_baseline:
        cpu = groth16_cpu_baseline(pk, ints_to_limbs(z), pk.domain_size.bit_length() - 1)
    block = {
        "metric": "groth16_proofs_per_sec", "value": 1.0 / dt, "unit": "proofs/s", "n_gpus": 1, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (Montgomery, 254-bit)", "data": "synthetic",
        "config": {"workload": f"Groth16 prove, BN254, Poseidon arity-2 hash chain, {links} links = {r1.n_constraints} constraints, "
                               f"{r1.n_vars} variables, domain 2^{pk.domain_size.bit_length() - 1}",
                   "setup_s": t_setup, "circuit_and_witness_s": t_circuit},
        "clocks": clocks,
        "e2e": {"value": 1.0 / dt, "unit": "proofs/s", "h2d_bytes_per_step": int(z_m.nbytes) + 64,
                "d2h_bytes_per_step": 64 + 128 + 64, "api": "ozl_groth16_prove (C ABI, pinned host witness)"},
        "gpu_launches": int(launches), "stages_ms": stages, "timeline_ms": timeline, "verified": verified,
        "verified_how": "pairing equation e(A,B) = e(alpha,beta) e(IC,gamma) e(C,delta) on the timed proof at this size, product verifier (ate) "
                        "and oracle verifier (Tate) both accept it and the product rejects a wrong public input",
        "verify_s": verify_s, "concurrent": conc, "cpu_baseline": cpu,
    }
    pk.free()
    ctx.close()
    return block


def g2_msm_block(args, log_n: int = 22, device_index: int = 0, steps: int = 4):
    """BLS12-381 G2 MSM (the b_g2_query MSM of a BLS12-381 Groth16 proof), device-resident scalars, one
    shifted copy of the bases per window, verified against the known discrete logs."""
    import torch
    import openzl_b200 as ozl
    from oracle import cbind
    n = 1 << log_n
    dev = torch.device("cuda", device_index)
    ctx = ozl.Context(device_index)
    ctx.use_torch_stream()
    curve = ozl.BLS12_381_G2
    h = ctx.generate_bases(curve, 1, n)
    h.precompute(32)
    sc = device_scalars(n, R381, 99, dev)
    out = torch.zeros(36, dtype=torch.int64, device=dev)
    for _ in range(2):
        h.msm_device(sc.data_ptr(), n, out.data_ptr())
    torch.cuda.synchronize()
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        h.msm_device(sc.data_ptr(), n, out.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = ctx.launch_count - l0
    verified = None
    if not args.no_verify:
        k = cbind.dot_mod_r("bls12_381_fr", sc.cpu().numpy().view(np.uint64), np.arange(1, n + 1, dtype=np.uint64))
        exp, _ = cbind.to_affine("bls12_381_g2", cbind.gen_mul("bls12_381_g2", k))
        got, _ = ctx.jacobian_to_affine(curve, out.cpu().numpy().view(np.uint64))
        verified = bool((got == exp).all())
    info = h.info(n)
    h.free()
    del sc, out
    ctx.close()
    torch.cuda.empty_cache()
    return {"metric": "bls12_381_g2_msm_points_per_sec", "value": n / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": 2,
            "ms_per_step": ms, "config": {"workload": f"BLS12-381 G2 Pippenger MSM, 2^{log_n} points, window c={info['c']} ({info['windows']} windows "
                                                      f"in {info['bucket_sets']} bucket sets), {info['factor']} base copies, scalars resident"},
            "gpu_launches": int(launches), "verified_vs_known_dlog": verified}


def run_groth16(args):
    emit(groth16_block(args))


def ntt_block(args, log_n: int = 24, device_index: int = 0, steps: int = 10):
    """BASELINE config 3: BN254 Fr radix-2 NTT, forward in-order in place, 2^log_n elements."""
    import torch
    import openzl_b200 as ozl
    r254 = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    n = 1 << log_n
    dev = torch.device("cuda", device_index)
    ctx = ozl.Context(device_index)
    ctx.use_torch_stream()
    x = device_scalars(n, r254, 3, dev)
    h = torch.empty((n, 4), dtype=torch.int64, pin_memory=True)
    h.copy_(x)
    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        ctx.ntt_device(ozl.BN254_FR, x.data_ptr(), log_n, False, False)
    torch.cuda.synchronize()
    sampler = ClockSampler(device_index)
    l0 = ctx.launch_count
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ctx.ntt_device(ozl.BN254_FR, x.data_ptr(), log_n, False, False)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = ctx.launch_count - l0
    ms = e0.elapsed_time(e1) / steps
    passes = launches // steps
    hv = h.numpy().view(np.uint64)
    ctx.ntt(ozl.BN254_FR, hv)                      # warm the staging buffer
    t0 = time.perf_counter()
    for _ in range(max(2, steps // 2)):
        ctx.ntt(ozl.BN254_FR, hv)
    e2e_s = (time.perf_counter() - t0) / max(2, steps // 2)
    # size-independent check outside the timed region: inverse(forward(x)) == x on the device
    y = x.clone()
    ctx.ntt_device(ozl.BN254_FR, y.data_ptr(), log_n, False, False)
    ctx.ntt_device(ozl.BN254_FR, y.data_ptr(), log_n, True, False)
    torch.cuda.synchronize()
    round_trip = bool(torch.equal(x, y))
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import cbind
        sample_log = min(log_n, 22)
        xs = hv[: 1 << sample_log].copy()
        t0 = time.perf_counter()
        cbind.ntt("bn254_fr", xs)
        dtc = time.perf_counter() - t0
        cpu = {"value": (1 << sample_log) / dtc, "unit": "elements/s", "cores": 1, "kind": "port",
               "sample": f"one serial radix-2 NTT of 2^{sample_log} elements (ark without `parallel`), {dtc:.2f} s"}
    peaks, peak_src = measured_peaks()
    mul_peak = ctx.bench_field_mul(1, 4000)
    achieved = n * 64 / (ms * 1e-3) / 1e9
    block = {
        "metric": "bn254_fr_ntt_elements_per_sec", "value": n / (ms * 1e-3), "unit": "elements/s", "n_gpus": 1,
        "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 limbs (Montgomery, 254-bit)", "data": "synthetic",
        "config": {"workload": f"BN254 Fr forward NTT 2^{log_n}, natural order in/out, in place, {passes} kernel passes",
                   "l2": f"data {n * 32 >> 20} MiB (+ equal scratch, + {n * 16 >> 20} MiB twiddles) vs 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": n / e2e_s, "unit": "elements/s", "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": n * 32,
                "ms_per_step": e2e_s * 1e3, "api": "ozl_ntt (C ABI, pinned host buffer)"},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "k_ntt_tile", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_transform": n * 64, "passes": passes, "avg_launch_ms": ms / max(passes, 1)},
        "fma_pipe": {"note": "binding roofline: one 254-bit Montgomery multiplication per butterfly",
                     "field_mul_per_s": (n / 2) * log_n / (ms * 1e-3), "measured_mul_peak_per_s": mul_peak,
                     "frac": (n / 2) * log_n / (ms * 1e-3) / mul_peak},
        "verified_round_trip": round_trip, "cpu_baseline": cpu,
    }
    del x, y, h
    ctx.close()
    return block


def run_ntt(args):
    emit(ntt_block(args, args.log_n if args.log_n <= 28 and args.log_n != 26 else 24, 0, args.steps))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="msm", choices=["msm", "groth16", "ntt"])
    ap.add_argument("--links", type=int, default=3013)
    ap.add_argument("--concurrency", type=int, default=2, help="groth16: independent provers in flight for the throughput figure")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=26)
    ap.add_argument("--total-log-n", type=int, default=0, help="strong scaling: total points = 2^k split across ranks")
    ap.add_argument("--cpu-log-n", type=int, default=20, help="reference arm: points per step of the CPU sample")
    ap.add_argument("--cpu-baseline-log-n", type=int, default=22, help="our arm: size of the one-shot cpu_baseline sample")
    ap.add_argument("--window-bits", type=int, default=0)
    ap.add_argument("--precompute", type=int, default=0, help="shifted base copies kept in HBM (1 = none, 0 = best measured for the size)")
    ap.add_argument("--g16-precompute", type=int, default=32, help="groth16: shifted copies of each proving-key query")
    ap.add_argument("--g16-steps", type=int, default=10, help="timed proofs of the groth16 block")
    ap.add_argument("--no-groth16", action="store_true", help="msm workload: skip the Groth16 @2^20 block")
    ap.add_argument("--no-ntt", action="store_true", help="msm workload: skip the BN254 Fr NTT 2^24 block")
    ap.add_argument("--strong-log-n", type=int, default=28, help="msm workload: total size of the strong-scaling block (BASELINE config 5); 0 = skip")
    ap.add_argument("--strong-steps", type=int, default=3)
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    global _SAVED_STDOUT
    _SAVED_STDOUT = _stdout_to_stderr()
    try:
        if args.workload == "groth16":
            run_groth16(args)
        elif args.workload == "ntt":
            run_ntt(args)
        elif args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        if _SAVED_STDOUT is not None:
            _restore_stdout(_SAVED_STDOUT)
            _SAVED_STDOUT = None


_SAVED_STDOUT = None


def emit(line: dict):
    """Print the JSON line on the real stdout (restored first if a run redirected it)."""
    global _SAVED_STDOUT
    if _SAVED_STDOUT is not None:
        _restore_stdout(_SAVED_STDOUT)
        _SAVED_STDOUT = None
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
