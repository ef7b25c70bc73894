#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Nystrom spectral filter.

Metric (BASELINE.json): enhance MP/s (p=1600, k=50).  A "step" = train + enhance of ONE image
(NLEFilter::trainFilter + enhance on the L channel, filter.cpp:480-502 and 426-436).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

N=1 workload = BASELINE.json configs[2]: synthetic 1024x1024 gray image, 40x40=1600 samples,
hx=500 hy=30, 20 Sinkhorn iterations, k=50, weights 2 3 4 1 (SURVEY.md 8d "S-gray-1024").
N>1 = weak scaling: the image grows to (1024*N) x 1024 and is sharded by image rows, one slab per
rank; the library's own communicator (csrc/nccl_comm.cu) carries the p-vector Sinkhorn sums and the k-vector V^T z
(peer_allreduce_kernel over NVLink peer memory) and one p x p Gram (ncclAllReduce).  At N>1 the line also carries "multi_gpu_parity": every rank's enhanced slab against the SAME image
trained unsharded on rank 0 (outside the timed region; the run fails if they differ by more than 1 LSB).
Extra keys measured after the main timed region: "enhance_only" (train once, 50 enhance calls: the HBM-bound apply
pass), "c5_strong" (BASELINE configs[4]: 4096x4096 BGR, p=2500, k=100, strong-scaled over the N ranks, BGR in/out) and
"c4_strong" (configs[3]: full-resolution rock2, p=2500, k=100, T=50), "roofline_all" (every stage >= 5 % of the step).

The workload is a gray image in 3 equal BGR channels, as the reference's CLI would read it; the filter's input is the L
channel of its 8-bit BGR2Lab conversion (filter.cpp:463-466).
value : MP/s with that luminance slab already resident in HBM (nle_b200_train_u8_dev +
        nle_b200_enhance_luminance_u8_dev), timed with CUDA events, max over ranks.
e2e   : MP/s through the image-level host-pointer C ABI (nle_b200_train_bgr_u8 + nle_b200_enhance_bgr_u8): BGR image
        in pinned host memory in, BGR image out, colour conversion on the device, H2D and D2H inside the timed region.
--impl reference : the CPU restatement of the reference (oracle/nle_oracle.py, NumPy/SciPy FP64,
        all host threads) on a bounded crop of the same workload -- its `config` names the crop it ran (rows/cols of
        the crop, not of the GPU workload).  The reference binary itself needs Eigen + OpenCV C++ which this image
        does not have.
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


def synth_luminance(rows, cols, seed=1234, period_x=97.0, period_y=61.0):
    """SURVEY.md 8(d) S-gray generator: smooth periodic structure + 5x5-box-smoothed noise, 8-bit."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:rows, 0:cols].astype(np.float64)
    noise = rng.standard_normal((rows + 4, cols + 4))
    cs = np.cumsum(np.cumsum(np.pad(noise, ((1, 0), (1, 0))), axis=0), axis=1)
    box = (cs[5:, 5:] - cs[:-5, 5:] - cs[5:, :-5] + cs[:-5, :-5]) / 25.0
    img = 128.0 + 60.0 * np.sin(2 * np.pi * x / period_x) * np.cos(2 * np.pi * y / period_y) + 25.0 * 5.0 * box
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synth_bgr(rows, cols):
    """SURVEY.md 8(d) S-rgb generator (BASELINE configs[4]): the S-gray generator per channel, seeds 1234/1235/1236 and
    different periods, interleaved BGR."""
    chans = [synth_luminance(rows, cols, 1234 + c, 97.0 + 14.0 * c, 61.0 + 9.0 * c) for c in range(3)]
    return np.ascontiguousarray(np.stack(chans, axis=2))


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
def cpu_step(lum_crop, stages=None):
    """One train + enhance of the dense FP64 restatement (oracle/nle_oracle.py::train_dense, line by line filter.cpp:480-502)
    on `lum_crop`.  stages (optional dict) receives the wall-clock split: seconds in the three p x p eigensolves
    (p^3 work, independent of the image size) and in everything else (work proportional to the pixel count)."""
    from oracle import nle_oracle as O
    t_eig = [0.0]
    real = O.eigen_decomposition

    def timed_eig(M, eps=O.EPS):
        t0 = time.perf_counter()
        try:
            return real(M, eps)
        finally:
            t_eig[0] += time.perf_counter() - t0
    t0 = time.perf_counter()
    O.eigen_decomposition = timed_eig
    try:
        flt = O.train_dense(lum_crop.astype(np.float64), GRID[0], GRID[1], HX, HY, T_SINK, K_EIG)
        out = O.enhance_luminance(flt, lum_crop, WEIGHTS)
    finally:
        O.eigen_decomposition = real
    if stages is not None:
        total = time.perf_counter() - t0
        stages["eigensolves_s"] = stages.get("eigensolves_s", 0.0) + t_eig[0]
        stages["pixel_scaled_s"] = stages.get("pixel_scaled_s", 0.0) + (total - t_eig[0])
        stages["steps"] = stages.get("steps", 0) + 1
    return out


def cpu_sample_desc():
    return (f"{CPU_CROP}x{CPU_CROP} top-left crop of the {BASE_ROWS}x{COLS} workload, same 40x40 grid "
            f"(p=1600), T={T_SINK}, k={K_EIG}; dense FP64 NumPy/SciPy restatement of filter.cpp "
            f"(oracle/nle_oracle.py), OpenBLAS threads = all host cores; the full image needs ~80 GB dense")


def cpu_extrapolation(stages):
    """Estimate for the FULL 1024x1024 image from the crop's stage split: the eigensolves cost the same at any image
    size (p is fixed), the rest scales with the pixel count.  Reported next to the measured crop number, never as it."""
    n = max(1, stages.get("steps", 1))
    t_eig, t_pix = stages["eigensolves_s"] / n, stages["pixel_scaled_s"] / n
    scale = (BASE_ROWS * COLS) / float(CPU_CROP * CPU_CROP)
    t_full = t_eig + t_pix * scale
    return {"crop_seconds_eigensolves": t_eig, "crop_seconds_pixel_scaled": t_pix,
            "full_image_seconds_estimate": t_full, "full_image_mp_s_estimate": BASE_ROWS * COLS / 1e6 / t_full,
            "note": "estimate = eigensolve seconds + pixel-scaled seconds x (1024^2 / crop pixels); the crop itself "
                    "over-weights the p^3 eigensolves per pixel by that factor"}


def crop_config():
    return {"workload": f"{CPU_CROP}x{CPU_CROP} top-left crop of the synthetic {BASE_ROWS}x{COLS} gray image (S-gray generator, "
                        f"seed 1234), L of 8-bit BGR2Lab, {GRID[0]}x{GRID[1]}=1600 Nystrom samples, hx={HX:g} hy={HY:g}, "
                        f"T={T_SINK} Sinkhorn iters, k={K_EIG}, weights {WEIGHTS}",
            "baseline_config": "bounded CPU sample of BASELINE.json configs[2] (NOT the full 1024x1024 image: its dense "
                               "intermediates need ~80 GB and ~minutes per step)",
            "rows": CPU_CROP, "cols": CPU_CROP, "p": 1600, "k": K_EIG, "sinkhorn_iters": T_SINK,
            "parallelism": "host cores (OpenBLAS threads)",
            "gpu_arm_rows": BASE_ROWS, "gpu_arm_cols": COLS}


def run_reference(args, rank, world):
    if rank != 0:
        return
    lum = workload_images(BASE_ROWS, COLS)[1][:CPU_CROP, :CPU_CROP].copy()
    for _ in range(args.warmup):
        cpu_step(lum)
    stages = {}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(lum, stages)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    mp = CPU_CROP * CPU_CROP / 1e6
    val = mp / dt
    line = {
        "impl": "reference", "metric": "enhance MP/s (p=1600,k=50)", "value": val, "unit": "MP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": crop_config(),
        "cpu_baseline": {"value": val, "unit": "MP/s", "cores": nproc_used(), "kind": "port", "sample": cpu_sample_desc(),
                         "extrapolated": cpu_extrapolation(stages)},
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


def hbm_peak_gbs():
    """HBM roofline denominator: the driver-written MEASURED_PEAKS.json when it travelled with the repo, else the
    fallback /opt/skills/guides/B200_PROFILING.md states for this pool's B200s."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6547.5, "fallback (B200_PROFILING.md: measured copy bandwidth of this pool's B200s; MEASURED_PEAKS.json absent)"


def sample_axis(n, k):
    step = n // k
    off = (step - 1 + (n - step * k)) // 2
    r = np.arange(n)
    return r[(r >= off) & (r <= n - off) & ((r - off) % step == 0)]


# ---- B200 arm -----------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from nonlocal_image_edit_b200 import _lib
    from nonlocal_image_edit_b200.sharding import LibraryComm, row_slab
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    lib.nle_b200_set_keep_stages(0)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = LibraryComm(dev)          # all-reduces enqueued by the library itself (csrc/nccl_comm.cu)
    cb = comm.callback if comm else C.cast(None, _lib.ALLREDUCE_FN)
    user = comm.user if comm else None

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
    ys = np.ascontiguousarray(lum[np.ix_(sample_axis(rows, GRID[0]), sample_axis(COLS, GRID[1]))].ravel())
    assert ys.size == p.value

    pinned_in = torch.from_numpy(bgr).pin_memory()
    pinned_out = torch.empty(nloc * 3, dtype=torch.uint8).pin_memory()
    d_slab = torch.from_numpy(lum[row0:row1].copy()).to(dev)
    d_out = torch.empty(nloc, dtype=torch.uint8, device=dev)

    def step_dev():
        h = C.c_void_p()
        _lib.check(lib.nle_b200_train_u8_dev(C.c_void_p(d_slab.data_ptr()), rows, COLS, row0, row1,
                                             ys.ctypes.data_as(C.c_void_p), GRID[0], GRID[1], HX, HY, T_SINK, K_EIG,
                                             cb, user, C.byref(h)))
        _lib.check(lib.nle_b200_enhance_luminance_u8_dev(h, C.c_void_p(d_slab.data_ptr()), weights, len(WEIGHTS),
                                                         C.c_void_p(d_out.data_ptr())))
        return h

    def step_host():
        h = C.c_void_p()
        # trainForEnhancement(BGR) + enhance(BGR) -> BGR, host buffers in and out (filter.cpp:514-519, 412-443)
        _lib.check(lib.nle_b200_train_bgr_u8(C.c_void_p(pinned_in.data_ptr()), rows, COLS, row0, row1, GRID[0], GRID[1],
                                             HX, HY, T_SINK, K_EIG, cb, user, C.byref(h)))
        _lib.check(lib.nle_b200_enhance_bgr_u8(h, C.c_void_p(pinned_in.data_ptr() + 3 * row0 * COLS), row1 - row0, COLS, 3,
                                               weights, len(WEIGHTS), C.c_void_p(pinned_out.data_ptr())))
        return h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    step_ms = {}

    def timed(fn, steps, collect=None, tag="dev", free=True):
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for i in range(steps):
            h = fn()
            if collect is not None:
                collect(h)
            if free and h is not None:
                lib.nle_b200_free(h)
            ev[i + 1].record()
        barrier()
        ms = ev[0].elapsed_time(ev[steps])
        if tag:
            step_ms[tag] = [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(steps)]
        return max_over_ranks(ms)

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
        out = (C.c_double * 16)()
        size = C.c_size_t(0)
        lib.nle_b200_get_stage(h, 8, out, 16, C.byref(size))
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

    # ---- N > 1: the sharded result against the same image trained UNSHARDED on rank 0 (outside the timed region)
    parity = None
    if world > 1:
        h = step_dev()
        S_sh = (C.c_double * K_EIG)()
        inf_sh = _lib.Info()
        lib.nle_b200_filter_info(h, C.byref(inf_sh))
        _lib.check(lib.nle_b200_eigenvalues(h, S_sh))
        lib.nle_b200_free(h)
        mine = d_out.clone()
        full = torch.empty(rows * COLS, dtype=torch.uint8, device=dev)
        S_one = torch.zeros(K_EIG + 4, dtype=torch.float64, device=dev)
        if rank == 0:
            d_full = torch.from_numpy(lum).to(dev)
            h1 = C.c_void_p()
            _lib.check(lib.nle_b200_train_u8_dev(C.c_void_p(d_full.data_ptr()), rows, COLS, 0, rows, ys.ctypes.data_as(C.c_void_p),
                                                 GRID[0], GRID[1], HX, HY, T_SINK, K_EIG, C.cast(None, _lib.ALLREDUCE_FN), None,
                                                 C.byref(h1)))
            _lib.check(lib.nle_b200_enhance_luminance_u8_dev(h1, C.c_void_p(d_full.data_ptr()), weights, len(WEIGHTS),
                                                             C.c_void_p(full.data_ptr())))
            inf1 = _lib.Info()
            lib.nle_b200_filter_info(h1, C.byref(inf1))
            s1 = (C.c_double * K_EIG)()
            _lib.check(lib.nle_b200_eigenvalues(h1, s1))
            lib.nle_b200_free(h1)
            S_one[:inf1.k] = torch.tensor(list(s1)[:inf1.k], dtype=torch.float64)
            S_one[K_EIG:] = torch.tensor([inf1.r, inf1.r2, inf1.k, 0], dtype=torch.float64)
            del d_full
        torch.cuda.synchronize()
        dist.broadcast(full, src=0)
        dist.broadcast(S_one, src=0)
        ref_slab = full[row0 * COLS:row1 * COLS]
        diff = (mine.to(torch.int16) - ref_slab.to(torch.int16)).abs()
        stats = torch.tensor([float(diff.max().item()), float((diff == 0).sum().item()), float((diff <= 1).sum().item())],
                             dtype=torch.float64, device=dev)
        mx = stats[:1].clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        cnt = stats[1:].clone()
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        k1 = int(S_one[K_EIG + 2].item())
        s_sh = np.array(list(S_sh)[:inf_sh.k])
        s_1 = S_one[:k1].cpu().numpy()
        same_ranks = (inf_sh.r, inf_sh.r2, inf_sh.k) == (int(S_one[K_EIG].item()), int(S_one[K_EIG + 1].item()), k1)
        sq_rel = float(np.abs(s_sh - s_1).max() / np.abs(s_1).max()) if same_ranks else None
        parity = {"max_abs_diff": int(mx.item()), "frac_identical": float(cnt[0].item() / (rows * COLS)),
                  "frac_within_1": float(cnt[1].item() / (rows * COLS)), "Sq_rel": sq_rel,
                  "ranks_sharded": [inf_sh.r, inf_sh.r2, inf_sh.k],
                  "ranks_single_gpu": [int(S_one[K_EIG].item()), int(S_one[K_EIG + 1].item()), k1],
                  "what": f"enhanced L of the {rows}x{COLS} image: {world} row slabs over NCCL vs the same image unsharded on rank 0"}
        del full, mine

    # ---- train once / enhance many (n4): the HBM-bound apply pass (K5) on a trained handle
    h_keep = step_dev()
    inf_keep = _lib.Info()
    lib.nle_b200_filter_info(h_keep, C.byref(inf_keep))
    wsets = [(C.c_double * 4)(1.0 + 0.05 * i, 2.0 + 0.03 * i, 3.0 - 0.02 * i, 1.0) for i in range(50)]
    it = {"i": 0}

    def enhance_dev():
        w = wsets[it["i"] % 50]
        it["i"] += 1
        _lib.check(lib.nle_b200_enhance_luminance_u8_dev(h_keep, C.c_void_p(d_slab.data_ptr()), w, 4, C.c_void_p(d_out.data_ptr())))

    def enhance_host():
        w = wsets[it["i"] % 50]
        it["i"] += 1
        _lib.check(lib.nle_b200_enhance_bgr_u8(h_keep, C.c_void_p(pinned_in.data_ptr() + 3 * row0 * COLS), row1 - row0, COLS, 3,
                                               w, 4, C.c_void_p(pinned_out.data_ptr())))
    for _ in range(3):
        enhance_dev()
        enhance_host()
    ms_en_dev = timed(enhance_dev, 50, tag=None, free=False) / 50
    ms_en_host = timed(enhance_host, 50, tag=None, free=False) / 50
    lib.nle_b200_free(h_keep)
    hbm_peak, hbm_src = hbm_peak_gbs()
    k5_bytes = 16.0 * nloc * inf_keep.k + 2.0 * nloc           # V read twice (V^T z, then V g) + L in + L out
    k5_gbs = k5_bytes / (ms_en_dev * 1e-3) * 1e-9
    enhance_only = {
        "what": "one trained handle, 50 enhance calls with 50 different weight sets (filter.hpp:44: enhance is const)",
        "device_resident": {"ms_per_call": ms_en_dev, "mp_s": rows * COLS / 1e6 / (ms_en_dev * 1e-3)},
        "bgr_host_in_out": {"ms_per_call": ms_en_host, "mp_s": rows * COLS / 1e6 / (ms_en_host * 1e-3),
                            "h2d_bytes_per_call": 3 * nloc * world, "d2h_bytes_per_call": 3 * nloc * world},
        "roofline": {"kernel": "vtz_kernel + recompose_kernel (the whole device-resident enhance call: two passes over V, "
                               "k-vector all-reduce, weights upload)",
                     "bound": "hbm", "achieved": k5_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": k5_gbs / hbm_peak,
                     "algorithmic_bytes_per_call": k5_bytes, "peak_source": hbm_src, "traffic": None}}

    # ---- the north_star target configurations, strong-scaled over the N ranks, BGR host in / BGR host out
    def strong(bgr_img, grid, hx, hy, T, k, w, steps=3):
        R, Wd = bgr_img.shape[:2]
        r0, r1 = row_slab(R, rank, world)
        pin = torch.from_numpy(bgr_img).pin_memory()
        pout = torch.empty((r1 - r0) * Wd * 3, dtype=torch.uint8).pin_memory()
        wv = (C.c_double * len(w))(*w)
        st, last = [], {}

        def one():
            h = C.c_void_p()
            _lib.check(lib.nle_b200_train_bgr_u8(C.c_void_p(pin.data_ptr()), R, Wd, r0, r1, grid[0], grid[1], hx, hy, T, k,
                                                 cb, user, C.byref(h)))
            _lib.check(lib.nle_b200_enhance_bgr_u8(h, C.c_void_p(pin.data_ptr() + 3 * r0 * Wd), r1 - r0, Wd, 3, wv, len(w),
                                                   C.c_void_p(pout.data_ptr())))
            out = (C.c_double * 16)()
            size = C.c_size_t(0)
            lib.nle_b200_get_stage(h, 8, out, 16, C.byref(size))
            st.append(list(out))
            inf = _lib.Info()
            lib.nle_b200_filter_info(h, C.byref(inf))
            last["inf"] = inf
            return h
        for _ in range(2):                 # the second call after a change of image shape is the steady state (the
            lib.nle_b200_free(one())       # temporaries' arena is coalesced into one chunk on the call after it grew)
        st.clear()
        ms = timed(one, steps, tag=f"strong_{R}x{Wd}") / steps
        m = np.median(np.array(st), axis=0)
        inf = last["inf"]
        return {"rows": R, "cols": Wd, "p": inf.p, "r": inf.r, "r2": inf.r2, "k": inf.k, "sinkhorn_iters": T,
                "n_gpus": world, "steps": steps, "ms_per_image": ms, "mp_s": R * Wd / 1e6 / (ms * 1e-3),
                "api": "nle_b200_train_bgr_u8 + nle_b200_enhance_bgr_u8, BGR host in / BGR host out, rows sharded over the ranks",
                "stage_ms_rank0": {"eig_Ka": m[1], "sinkhorn_passes": m[2], "gram": m[3], "small_algebra_2eigs": m[4],
                                   "extension": m[5], "train_total": m[6], "tridiag_3_solves": m[8], "divide_conquer": m[9],
                                   "back_transform": m[10]}}
    c5 = c4 = None
    if not args.no_targets:
        c5 = strong(synth_bgr(4096, 4096), (50, 50), 500.0, 30.0, 20, 100, [2.0, 3.0, 4.0, 1.0])
        c5["config"] = "BASELINE.json configs[4]: synthetic 4096x4096 BGR (S-rgb generator), 50x50=2500 samples, k=100, hx=500 hy=30, T=20"
        rock = os.path.join(ROOT, "tests", "golden", "rock2_input.png")
        if os.path.exists(rock):
            import cv2
            c4 = strong(cv2.imread(rock), (50, 50), 500.0, 10.0, 50, 100, [4.0, 3.0, 4.0, 1.0])
            c4["config"] = "BASELINE.json configs[3]: full-resolution data/rock2.jpg (584x876), 50x50=2500 samples, k=100, hx=500 hy=10, T=50"

    if rank != 0:
        if comm:
            comm.close()
        if world > 1:
            dist.destroy_process_group()
        return

    comm_info = comm.info() if comm else None
    total_mp = rows * COLS / 1e6
    ms_step = ms_dev / args.steps
    value = total_mp / (ms_step * 1e-3)
    e2e_val = total_mp / (ms_host / args.steps * 1e-3)
    st = np.median(np.array(stage_ms), axis=0)
    inf = infos[-1]

    # ---- rooflines.  Algorithmic work per launch, CUDA-event time of the stage on the library's stream.
    pp, nR, nC, kk = inf.p, inf.n_row_samples_eff, inf.n_col_samples_eff, inf.k
    slab = lum[row0:row1]
    k_cells = int(sum(np.unique(r).size for r in slab))
    pk = (C.c_double * 8)()
    _lib.check(lib.nle_b200_measured_peaks(pk, 8))            # csrc/peaks.cu, outside every timed region
    dfma_peak, dmma_peak = float(pk[0]), float(pk[1])
    measured_peaks = {"fp64_fma_tflops": pk[0], "fp64_dmma_tflops": pk[1], "fp32_fma_tflops": pk[2], "mufu_ex2_gops": pk[3],
                      "shared_load_gbs": pk[4], "shared_load_bytes_per_sm_clk": pk[7], "l2_read_gbs_32mb": pk[5], "hbm_copy_gbs": pk[6],
                      "source": "microbenchmarks of csrc/peaks.cu run on this GPU in this process (best of 5 launches each)"}
    nominal = 148 * 64 * 2 * 1.965e9 * 1e-12
    peak_src = ("measured in this run: register-resident DFMA microbenchmark (nle_b200_fp64_fma_peak_tflops) and back-to-back "
                "mma.sync.m8n8k4.f64 microbenchmark (nle_b200_fp64_dmma_peak_tflops); nominal 148 SM x 64 FP64 lanes x 2 x "
                f"1.965 GHz = {nominal:.1f} TFLOP/s; MEASURED_PEAKS.json has no FP64 figure")

    def tensor_entry(kernel, flops, ms, note, peak=None):
        peak = peak or dmma_peak
        ach = flops / (ms * 1e-3) * 1e-12 if ms > 0 else None
        return {"kernel": kernel, "bound": "tensor", "pipe": "FP64 tensor pipe (mma.sync m8n8k4.f64, SASS DMMA); tcgen05 has no FP64 kind",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": (ach / peak) if (ach and peak) else None,
                "launch_ms": float(ms), "share_of_step": float(ms / ms_step), "algorithmic_flops_per_launch": flops,
                "peak_dfma_tflops": dfma_peak, "peak_dmma_tflops": dmma_peak, "peak_nominal_tflops": nominal, "note": note}
    gram_flops = float(k_cells) * pp * (pp + 1)
    gram = tensor_entry("gram_cells_kernel (+ cell sort, per-cell histograms, split reduce)", gram_flops, st[7],
                        f"K_cells*p*(p+1), K_cells={k_cells}; SURVEY 8d's pixel-axis figure N*p*(p+1) = {float(nloc) * pp * (pp + 1):.3e} "
                        "is the work the cell re-association removes, not a roofline numerator")
    gram["traffic"] = 1.2391e9 if world == 1 else None
    gram["traffic_source"] = "profiles/r1g_gram_cells_kernel_full.md (dram read + write of one launch; = Hh once + partials)" if world == 1 else None
    gram["cells"] = k_cells
    trd_flops = 4.0 / 3.0 * (float(pp) ** 3 + float(inf.r) ** 3 + float(inf.r2) ** 3)
    trd = tensor_entry("tridiag_cluster_kernel x3 (Householder tridiagonalisation of Ka, Wa, Q-block)", trd_flops, st[8],
                       "(4/3)(p^3 + r^3 + r2^3) flop; BLAS-2 with one grid-wide exchange per Householder column: bound by the "
                       "latency of that exchange (TMA multicast landing + validation, profiles/r2zb_trd_multicast.md) and of the two block reductions behind it, not by the FP64 pipes", peak=dfma_peak)
    trd["bound"] = "latency"
    trd["pipe"] = "FP64 FMA pipe (DFMA) + shared memory; per-step exchange of flagged cells through L2, landed by TMA multicast"
    nrows_loc = row1 - row0
    sk_flops = 2.0 * T_SINK * (2 * 2.0 * nrows_loc * 256 * pp + 4.0 * nloc * nC)
    sk = tensor_entry("Sinkhorn half-iterations x2T (sk_dot_gemm + sk_pix_cells + sk_reduce_gemm + sample-side passes)", sk_flops, st[2],
                      "2T x (two level-table GEMMs rows x 256 x p + 4 N nC pixel-pass flop); the cell pass is L2-latency bound "
                      "(profiles/r1g_sk_pix_cells_kernel_full.md)")
    ext_flops = 2.0 * k_cells * pp * kk + 2.0 * nloc * nC * kk
    ext = tensor_entry("extension (ext_fx_kernel + ext_pix_kernel + small GEMMs)", ext_flops, st[5],
                       "2 K_cells p k' + 2 N nC k' flop; writes V = 8 N k' bytes")
    ext["algorithmic_bytes_written"] = 8.0 * nloc * kk
    ext["hbm_gbs_on_V_write"] = 8.0 * nloc * kk / (st[5] * 1e-3) * 1e-9 if st[5] > 0 else None
    others = [{"kernel": "dc_* (divide & conquer, 3 solves)", "launch_ms": float(st[9]), "share_of_step": float(st[9] / ms_step)},
              {"kernel": "bt_wy_kernel / bt_tfactor_kernel (back-transformation, 3 solves)", "launch_ms": float(st[10]),
               "share_of_step": float(st[10] / ms_step)}]
    cand = [gram, trd, sk, ext]
    roofline_all = sorted([c for c in cand if c["share_of_step"] >= 0.05] + [o for o in others if o["share_of_step"] >= 0.05],
                          key=lambda c: -c["launch_ms"])
    roofline = dict(max(cand, key=lambda c: c["launch_ms"]))
    roofline["peak_source"] = peak_src

    # CPU baseline on a bounded crop (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        crop = lum[:CPU_CROP, :CPU_CROP].copy()
        stages = {}
        t0 = time.perf_counter()
        cpu_step(crop, stages)
        dt = time.perf_counter() - t0
        cpu = {"value": CPU_CROP * CPU_CROP / 1e6 / dt, "unit": "MP/s", "cores": nproc_used(), "kind": "port",
               "sample": cpu_sample_desc(), "seconds": dt, "rows": CPU_CROP, "cols": CPU_CROP,
               "extrapolated": cpu_extrapolation(stages)}

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
        "roofline_all": roofline_all,
        "measured_peaks": measured_peaks,
        "cpu_baseline": cpu,
        "stage_ms": {"setup_tables_Ka": float(st[0]), "eig_Ka": float(st[1]), "sinkhorn_passes": float(st[2]),
                     "gram": float(st[3]), "small_algebra_2eigs": float(st[4]), "extension": float(st[5]),
                     "train_total": float(st[6]), "gram_kernel_only": float(st[7]), "tridiag_3_solves": float(st[8]),
                     "divide_conquer_3_solves": float(st[9]), "back_transform_3_solves": float(st[10])},
        "step_ms": step_ms,
        "filter": {"p": inf.p, "r": inf.r, "r2": inf.r2, "k": inf.k, "eig_sweeps": list(inf.eig_sweeps),
                   "eig_fallbacks": inf.eig_fallbacks, "topk_products": inf.topk_products},
        "multi_gpu_parity": parity,
        "enhance_only": enhance_only,
        "c5_strong": c5,
        "c4_strong": c4,
        "collectives": ({"owner": "library-owned communicator (csrc/nccl_comm.cu)",
                         "small_messages": ("peer_allreduce_kernel: flagged cells stored into every rank's inbox over NVLink peer memory "
                                            "(CUDA IPC), summed in rank order, one launch per reduction") if comm_info["peer_path"]
                         else "ncclAllReduce (peer-memory path unavailable: %s)" % comm_info["why"],
                         "gram": "ncclAllReduce (p x p doubles)", **comm_info} if world > 1 else None),
    }
    print(json.dumps(line), flush=True)
    if comm:
        comm.close()
    if world > 1:
        dist.destroy_process_group()
    if parity is not None:
        bad = parity["max_abs_diff"] > 1 or parity["Sq_rel"] is None or parity["Sq_rel"] > 1e-7
        if bad:
            raise SystemExit(f"bench.py: multi-GPU result differs from the single-GPU result: {parity}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-targets", action="store_true", help="skip the c5_strong / c4_strong target configurations")
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
