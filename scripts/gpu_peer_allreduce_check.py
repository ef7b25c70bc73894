#!/usr/bin/env python
"""Correctness and latency of the library's peer-memory all-reduce (csrc/nccl_comm.cu, peer_allreduce_kernel).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
      scripts/gpu_peer_allreduce_check.py [--iters 300] [--json out.json]

Every rank creates the library's communicator (LibraryComm), then
  * `iters` reductions of random lengths (1 ... 8192 doubles, the peer path) back to back on one stream -- the
    double-buffered inboxes are reused every second call -- each compared BIT FOR BIT with the sum in rank order of the
    all-gathered inputs (what the kernel promises: identical on every rank);
  * one message above the limit (-> ncclAllReduce) against torch.distributed's all_reduce to rounding;
  * latency of a 1600-double reduction between two dependent kernels: peer path, the library's ncclAllReduce on the same
    buffer (NLE_B200_PEER_AR=off in a second communicator is not possible in one process, so the NCCL figure is
    torch.distributed.all_reduce on the same stream), CUDA events, max over ranks.
Rank 0 prints one JSON line; exit code 1 on any mismatch."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from nonlocal_image_edit_b200 import _lib
    from nonlocal_image_edit_b200.sharding import LibraryComm

    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    comm = LibraryComm(dev)
    info0 = comm.info()
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    def lib_allreduce(t):
        rc = lib.nle_b200_comm_allreduce(C.c_void_p(t.data_ptr()), t.numel(), sp, comm.handle)
        if rc != 0:
            raise RuntimeError(f"nle_b200_comm_allreduce rc={rc}: {lib.nle_b200_last_error().decode()}")

    g = torch.Generator(device="cpu"); g.manual_seed(1234)            # same lengths on every rank
    gv = torch.Generator(device="cpu"); gv.manual_seed(99 + rank)      # different values
    bad = 0
    worst = 0.0
    # all inputs first, reductions back to back afterwards (no host sync between them: exercises the parity reuse)
    lens = [int(torch.randint(1, 8193, (1,), generator=g)) for _ in range(args.iters)]
    lens[:4] = [1, 8192, 1600, 50]
    xs = [(torch.randn(n, generator=gv, dtype=torch.float64) * 10.0 ** float(torch.randint(-6, 7, (1,), generator=gv))).to(dev)
          for n in lens]
    gathered = []
    for x in xs:
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x)
        gathered.append(parts)
    torch.cuda.synchronize()
    ys = [x.clone() for x in xs]
    for y in ys:
        lib_allreduce(y)
    torch.cuda.synchronize()
    for y, parts in zip(ys, gathered):
        ref = parts[0].clone()
        for r in range(1, world):
            ref = ref + parts[r]                       # rank order, as the kernel sums
        if not torch.equal(y, ref):
            bad += 1
            worst = max(worst, float((y - ref).abs().max()))
    # above the limit: NCCL path of the same entry point
    big = torch.randn(20000, generator=gv, dtype=torch.float64).to(dev)
    big_ref = big.clone()
    dist.all_reduce(big_ref)
    lib_allreduce(big)
    torch.cuda.synchronize()
    big_err = float((big - big_ref).abs().max() / big_ref.abs().max())

    # latency between two dependent kernels
    v = torch.randn(1600, dtype=torch.float64, device=dev)
    def timed(fn, n=200):
        for _ in range(20):
            v.mul_(1.0); fn(v); v.mul_(1.0)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            v.mul_(1.0 / world); fn(v); v.mul_(1.0)   # 1/world keeps the values bounded
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    def base(_):
        pass
    us_base = timed(base)
    us_lib = timed(lib_allreduce)
    us_torch = timed(lambda t: dist.all_reduce(t))
    info1 = comm.info()
    t_bad = torch.tensor([bad], device=dev); dist.all_reduce(t_bad)
    out = {"world": world, "peer_path": info0["peer_path"], "why": info0["why"], "reductions_checked": args.iters,
           "mismatching_reductions_all_ranks": int(t_bad), "worst_abs_diff_rank0": worst, "nccl_path_rel_err": big_err,
           "us_two_kernels_alone": us_base, "us_with_library_allreduce_1600": us_lib, "us_with_torch_nccl_allreduce_1600": us_torch,
           "library_allreduce_us": us_lib - us_base, "torch_nccl_allreduce_us": us_torch - us_base,
           "peer_calls": info1["peer_calls"], "nccl_calls": info1["nccl_calls"]}
    if rank == 0:
        print(json.dumps(out), flush=True)
        if args.json:
            with open(args.json, "w") as f:
                json.dump(out, f, indent=1)
    comm.close()
    dist.destroy_process_group()
    if int(t_bad) != 0 or big_err > 1e-14:
        sys.exit(1)


if __name__ == "__main__":
    main()
