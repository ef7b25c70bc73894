"""CPU test (-m "not gpu") of the N>1 path's host logic: two gloo ranks, each owning a row slab, run the
factor-form pipeline with the SAME all-reduce adapter the CUDA path uses (sharding.torch_allreduce) and
must reproduce the single-rank result.  What crosses the wire is exactly what NCCL carries on the GPU
box: p-vectors, one p x p Gram."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist
    from nle_testlib import synth_lum
    from nonlocal_image_edit_b200.sharding import row_slab, torch_allreduce
    from oracle import nle_oracle as O
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        L = synth_lum(48, 64).astype(np.float64)
        args = (6, 8, 20.0, 25.0, 5, 6)
        reduce_fn = torch_allreduce(device=None)
        sent = []

        def allreduce(a):
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
            sent.append(a.size)
            reduce_fn(a.ctypes.data, a.size)
            return a
        slab = row_slab(48, rank, world)
        flt = O.train_streaming(L, *args, slab=slab, allreduce=allreduce)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), V=flt.eigvecs, S=flt.eigvals, slab=np.array(slab),
                 sent=np.array(sent), p=flt.stages["p"])
    finally:
        dist.destroy_process_group()


def test_two_rank_row_sharding_reproduces_single_rank(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, HERE)
    from nle_testlib import synth_lum
    from oracle import nle_oracle as O
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    full = O.train_streaming(synth_lum(48, 64).astype(np.float64), 6, 8, 20.0, 25.0, 5, 6)
    V = np.vstack([parts[0]["V"], parts[1]["V"]])
    assert tuple(parts[0]["slab"]) == (0, 24) and tuple(parts[1]["slab"]) == (24, 48)
    assert np.allclose(parts[0]["S"], full.eigvals, rtol=1e-10) and np.allclose(parts[1]["S"], full.eigvals, rtol=1e-10)
    # eigenvectors up to sign
    for j in range(full.eigvals.size):
        s = np.sign(np.dot(V[:, j], full.eigvecs[:, j]))
        assert np.allclose(s * V[:, j], full.eigvecs[:, j], atol=1e-9)
    p = int(parts[0]["p"])
    sent = parts[0]["sent"]
    assert set(sent.tolist()) <= {p, p * p}, "only p-vectors and one p x p Gram may cross ranks"
    assert (sent == p * p).sum() == 1
