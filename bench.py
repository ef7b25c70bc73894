#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Nystrom spectral filter.

Metric (BASELINE.json): enhance MP/s (p=1600, k=50).  A "step" = train + enhance of ONE image
(NLEFilter::trainFilter + enhance on the L channel, filter.cpp:480-502 and 426-436).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

N=1 workload = BASELINE.json configs[2]: synthetic 1024x1024 gray image, 40x40=1600 samples,
hx=500 hy=30, 20 Sinkhorn iterations, k=50, weights 2 3 4 1 (SURVEY.md 8d "S-gray-1024").
N>1 = weak scaling: the image grows to (1024*N) x 1024 and is sharded by image rows, one slab per
rank; NCCL carries the p-vector Sinkhorn sums, one p x p Gram and the k-vector V^T z.

The workload is a gray image in 3 equal BGR channels, as the reference's CLI would read it; the filter's input is the L
channel of its 8-bit BGR2Lab conversion (filter.cpp:463-466).
value : MP/s with that luminance slab already resident in HBM (nle_b200_train_u8_dev +
        nle_b200_enhance_luminance_u8_dev), timed with CUDA events, max over ranks.
e2e   : MP/s through the image-level host-pointer C ABI (nle_b200_train_bgr_u8 + nle_b200_enhance_bgr_u8): BGR image
        in pinned host memory in, BGR image out, colour conversion on the device, H2D and D2H inside the timed region.
--impl reference : the CPU restatement of the reference (oracle/nle_oracle.py, NumPy/SciPy FP64,
        all host threads) on a bounded crop of the same workload.  The reference binary itself
        needs Eigen + OpenCV C++ which this image does not have.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv:
    # the CPU arm uses every host core it is allowed to: torchrun exports OMP_NUM_THREADS=1 for its children, which
    # must be undone BEFORE NumPy / OpenBLAS load
    _n = str(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = _n

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---- workload -----------------------------------------------------------------------------------
BASE_ROWS, COLS = 1024, 1024
GRID = (40, 40)
HX, HY = 500.0, 30.0
T_SINK, K_EIG = 20, 50
WEIGHTS = [2.0, 3.0, 4.0, 1.0]
CPU_CROP = 384            # cpu_baseline / reference arm: CPU_CROP x CPU_CROP crop, same grid/k/T (~10 s of CPU work, ~8 GB dense)


def synth_luminance(rows, cols, seed=1234):
    """SURVEY.md 8(d) S-gray generator: smooth periodic structure + 5x5-box-smoothed noise, 8-bit."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:rows, 0:cols].astype(np.float64)
    noise = rng.standard_normal((rows + 4, cols + 4))
    cs = np.cumsum(np.cumsum(np.pad(noise, ((1, 0), (1, 0))), axis=0), axis=1)
    box = (cs[5:, 5:] - cs[:-5, 5:] - cs[5:, :-5] + cs[:-5, :-5]) / 25.0
    img = 128.0 + 60.0 * np.sin(2 * np.pi * x / 97.0) * np.cos(2 * np.pi * y / 61.0) + 25.0 * 5.0 * box
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def workload_images(rows, cols):
    """The bench workload as the reference's CLI sees it: the S-gray image in 3 equal BGR channels, and the L channel of
    its 8-bit BGR2Lab conversion (what getLuminanceChannel, filter.cpp:460-469, feeds the filter).  Host OpenCV, setup only."""
    import cv2
    gray = synth_luminance(rows, cols)
    bgr = np.ascontiguousarray(np.repeat(gray[:, :, None], 3, axis=2))
    lum = np.ascontiguousarray(cv2.cvtColor(bgr, cv2.COLOR_BGR2Lab)[:, :, 0])
    return bgr, lum


def nproc_used():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---- clocks sampler -----------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self, first_line=0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines[first_line:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- reference / CPU arm ------------------------------------------------------------------------
def cpu_step(lum_crop):
    from oracle import nle_oracle as O
    flt = O.train_dense(lum_crop.astype(np.float64), GRID[0], GRID[1], HX, HY, T_SINK, K_EIG)
    out = O.enhance_luminance(flt, lum_crop, WEIGHTS)
    return out


def cpu_sample_desc():
    return (f"{CPU_CROP}x{CPU_CROP} top-left crop of the {BASE_ROWS}x{COLS} workload, same 40x40 grid "
            f"(p=1600), T={T_SINK}, k={K_EIG}; dense FP64 NumPy/SciPy restatement of filter.cpp "
            f"(oracle/nle_oracle.py), OpenBLAS threads = all host cores; full image needs ~80 GB dense")


def run_reference(args, rank, world):
    if rank != 0:
        return
    lum = workload_images(BASE_ROWS, COLS)[1][:CPU_CROP, :CPU_CROP].copy()
    for _ in range(args.warmup):
        cpu_step(lum)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(lum)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    mp = CPU_CROP * CPU_CROP / 1e6
    val = mp / dt
    line = {
        "impl": "reference", "metric": "enhance MP/s (p=1600,k=50)", "value": val, "unit": "MP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "MP/s", "cores": nproc_used(), "kind": "port", "sample": cpu_sample_desc()},
        "e2e": {"value": val, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n):
    return {"workload": f"synthetic {BASE_ROWS * n}x{COLS} gray image (S-gray generator, seed 1234) in 3 equal BGR channels, L of 8-bit BGR2Lab, "
                        f"{GRID[0]}x{GRID[1]}=1600 Nystrom samples, hx={HX:g} hy={HY:g}, T={T_SINK} Sinkhorn iters, "
                        f"k={K_EIG}, weights {WEIGHTS}",
            "baseline_config": "BASELINE.json configs[2]",
            "rows": BASE_ROWS * n, "cols": COLS, "p": 1600, "k": K_EIG, "sinkhorn_iters": T_SINK,
            "parallelism": f"row-sharded x{n}" if n > 1 else "single GPU",
            "l2_policy": "each step streams >1.5 GB of scratch (per-cell histograms, Gram partials, V) through the 126 MB L2; "
                         "inputs are re-uploaded / re-read every step",
            "lab_conversion": "gray image as 3 equal BGR channels; value: L = 8-bit BGR2Lab resident in HBM; e2e: BGR in/out, "
                              "BGR<->Lab on the device (byte-exact with cv::cvtColor)"}


# ---- B200 arm -----------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from nonlocal_image_edit_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    lib.nle_b200_set_keep_stages(0)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    rows = BASE_ROWS * world
    # the workload as the reference's CLI sees it: a gray image in 3 equal BGR channels; the filter's input is the L
    # channel of its 8-bit BGR2Lab conversion (filter.cpp:463-466).  The device-resident arm is fed that L channel, the
    # e2e arm the BGR image itself (colour conversion on the device, byte-exact with cv::cvtColor: csrc/lab.cu).
    bgr, lum = workload_images(rows, COLS)
    lab_dev = np.empty_like(bgr)
    _lib.check(lib.nle_b200_bgr_to_lab_u8(C.c_void_p(bgr.ctypes.data), rows * COLS, C.c_void_p(lab_dev.ctypes.data)))
    assert np.array_equal(lab_dev[:, :, 0], lum), "device BGR2Lab differs from cv2"
    del lab_dev
    row0, row1 = rank * BASE_ROWS, (rank + 1) * BASE_ROWS
    nloc = (row1 - row0) * COLS
    weights = (C.c_double * len(WEIGHTS))(*WEIGHTS)

    # sample luminances (p bytes, host) -- needed by every rank
    p = C.c_int(0)
    _lib.check(lib.nle_b200_sample_count(rows, COLS, GRID[0], GRID[1], C.byref(p)))

    def axis(n, k):
        step = n // k
        off = (step - 1 + (n - step * k)) // 2
        r = np.arange(n)
        return r[(r >= off) & (r <= n - off) & ((r - off) % step == 0)]
    ys = np.ascontiguousarray(lum[np.ix_(axis(rows, GRID[0]), axis(COLS, GRID[1]))].ravel())
    assert ys.size == p.value

    class _Arr:   # zero-copy view of a raw device pointer for torch
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}

    def allreduce(buf, count, stream, user):
        try:
            t = torch.as_tensor(_Arr(buf, count), device=dev)
            dist.all_reduce(t)
            return 0
        except Exception:
            import traceback
            traceback.print_exc()
            return 1
    cb = _lib.ALLREDUCE_FN(allreduce) if world > 1 else C.cast(None, _lib.ALLREDUCE_FN)

    pinned_in = torch.from_numpy(bgr).pin_memory()
    pinned_out = torch.empty(nloc * 3, dtype=torch.uint8).pin_memory()
    d_slab = torch.from_numpy(lum[row0:row1].copy()).to(dev)
    d_out = torch.empty(nloc, dtype=torch.uint8, device=dev)

    def step_dev():
        h = C.c_void_p()
        _lib.check(lib.nle_b200_train_u8_dev(C.c_void_p(d_slab.data_ptr()), rows, COLS, row0, row1,
                                             ys.ctypes.data_as(C.c_void_p), GRID[0], GRID[1], HX, HY, T_SINK, K_EIG,
                                             cb, None, C.byref(h)))
        _lib.check(lib.nle_b200_enhance_luminance_u8_dev(h, C.c_void_p(d_slab.data_ptr()), weights, len(WEIGHTS),
                                                         C.c_void_p(d_out.data_ptr())))
        return h

    def step_host():
        h = C.c_void_p()
        # trainForEnhancement(BGR) + enhance(BGR) -> BGR, host buffers in and out (filter.cpp:514-519, 412-443)
        _lib.check(lib.nle_b200_train_bgr_u8(C.c_void_p(pinned_in.data_ptr()), rows, COLS, row0, row1, GRID[0], GRID[1],
                                             HX, HY, T_SINK, K_EIG, cb, None, C.byref(h)))
        _lib.check(lib.nle_b200_enhance_bgr_u8(h, C.c_void_p(pinned_in.data_ptr() + 3 * row0 * COLS), weights,
                                               len(WEIGHTS), C.c_void_p(pinned_out.data_ptr())))
        return h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_ms = {}

    def timed(fn, steps, collect=None, tag="dev"):
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for i in range(steps):
            h = fn()
            if collect is not None:
                collect(h)
            lib.nle_b200_free(h)
            ev[i + 1].record()
        barrier()
        ms = ev[0].elapsed_time(ev[steps])
        step_ms[tag] = [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(steps)]
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # nvidia-smi takes ~1 s to start and stalls CUDA calls of this process while it initialises: start it
    # before the warm-up and keep only the samples taken inside the timed region.
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        t_wait = time.time()
        while not sampler.lines and time.time() - t_wait < 8.0:
            time.sleep(0.05)
    for _ in range(max(3, args.warmup)):
        lib.nle_b200_free(step_dev())
    for _ in range(2):
        lib.nle_b200_free(step_host())

    stage_ms = []
    infos = []

    def collect(h):
        out = (C.c_double * 8)()
        size = C.c_size_t(0)
        lib.nle_b200_get_stage(h, 8, out, 8, C.byref(size))
        stage_ms.append(list(out))
        inf = _lib.Info()
        lib.nle_b200_filter_info(h, C.byref(inf))
        infos.append(inf)

    lib.nle_b200_launch_count(1)
    first_line = len(sampler.lines)
    ms_dev = timed(step_dev, args.steps, collect)
    launches = int(lib.nle_b200_launch_count(0))
    clocks = sampler.stop(first_line) if rank == 0 else None
    ms_host = timed(step_host, args.steps, tag="e2e")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_mp = rows * COLS / 1e6
    ms_step = ms_dev / args.steps
    value = total_mp / (ms_step * 1e-3)
    e2e_val = total_mp / (ms_host / args.steps * 1e-3)
    st = np.median(np.array(stage_ms), axis=0)
    inf = infos[-1]

    # roofline of the dominant kernel: gram_cells_kernel (FP64 tensor pipe, DMMA).  The Gram is contracted over
    # the non-empty (image row, luminance level) cells of this rank's slab: algorithmic work per
    # launch = one fused multiply-add per (cell, sample pair i<=j) = K_cells*p*(p+1) flops.  SURVEY.md 8d's
    # figure for the same quantity on the pixel axis, N*p*(p+1), is reported next to it: their ratio is the
    # work the re-association removes, not a roofline fraction.  launch_ms is the CUDA-event time of the whole
    # Gram stage (cell sort + per-cell histograms + gram_cells_kernel + split reduce), i.e. conservative.
    pp = inf.p
    slab = lum[row0:row1]
    k_cells = int(sum(np.unique(r).size for r in slab))
    gram_flops = float(k_cells) * pp * (pp + 1)
    survey_flops = float(nloc) * pp * (pp + 1)
    peak = lib.nle_b200_fp64_fma_peak_tflops()
    achieved = gram_flops / (st[7] * 1e-3) * 1e-12 if st[7] > 0 else None
    roofline = {"kernel": "gram_cells_kernel (register-generated affinity fragments + FP64 DMMA over (row, level) cells)",
                "bound": "tensor", "pipe": "FP64 tensor pipe (mma.sync m8n8k4.f64, SASS DMMA); not tcgen05: there is no FP64 tcgen05 MMA",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if (achieved and peak) else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of one gram_cells_kernel launch at this configuration
                # (ncu --set full, profiles/r1g_gram_cells_kernel_full.md); algorithmic HBM bytes are Hh once
                # (cells x nC(nC+1)/2 x 8 B = 1.11e9) plus the split partials: nothing is re-read.
                "traffic": 1.2391e9 if world == 1 else None,
                "traffic_source": "profiles/r1g_gram_cells_kernel_full.md" if world == 1 else None,
                "launch_ms": float(st[7]),
                "peak_source": "measured in this run by nle_b200_fp64_fma_peak_tflops (register-resident DFMA "
                               "microbenchmark); MEASURED_PEAKS.json has no FP64 figure",
                "algorithmic_flops_per_launch": gram_flops, "cells": k_cells,
                "pixel_axis_flops_survey_8d": survey_flops,
                "pixel_axis_equivalent_tflops": survey_flops / (st[7] * 1e-3) * 1e-12 if st[7] > 0 else None}

    # CPU baseline on a bounded crop (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        crop = lum[:CPU_CROP, :CPU_CROP].copy()
        t0 = time.perf_counter()
        cpu_step(crop)
        dt = time.perf_counter() - t0
        cpu = {"value": CPU_CROP * CPU_CROP / 1e6 / dt, "unit": "MP/s", "cores": nproc_used(), "kind": "port",
               "sample": cpu_sample_desc(), "seconds": dt}

    line = {
        "metric": "enhance MP/s (p=1600,k=50)", "value": value, "unit": "MP/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(world),
        "e2e": {"value": e2e_val, "unit": "MP/s", "h2d_bytes_per_step": int(2 * 3 * nloc + 3 * ys.size) * world,
                "d2h_bytes_per_step": int(3 * nloc + ys.size) * world, "ms_per_step": ms_host / args.steps,
                "api": "nle_b200_train_bgr_u8 + nle_b200_enhance_bgr_u8: BGR host image in, BGR host image out"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "stage_ms": {"setup_tables_Ka": float(st[0]), "eig_Ka": float(st[1]), "sinkhorn_passes": float(st[2]),
                     "gram": float(st[3]), "small_algebra_2eigs": float(st[4]), "extension": float(st[5]),
                     "train_total": float(st[6]), "gram_kernel_only": float(st[7])},
        "step_ms": step_ms,
        "filter": {"p": inf.p, "r": inf.r, "r2": inf.r2, "k": inf.k, "eig_sweeps": list(inf.eig_sweeps)},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
