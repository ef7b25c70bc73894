"""GPU: the two large BASELINE.json configurations on one GPU, through size-independent properties (the dense oracle
needs ~60 GB / >2 TB for them):
  configs[3]  full-resolution data/rock2.jpg, 50x50 = 2500 samples, k = 100          (reference OOMs per its README)
  configs[4]  synthetic 4096x4096 (16.7 MP), 2500 samples, k = 100                      (one channel; one GPU here)
Checks: sample count, eigenvalues sorted and S[0] ~ 1 (W is doubly stochastic after Sinkhorn), V f(S) V^T with f = 1 is
an (approximately orthogonal) projector, enhance is affine in the weights, a two-slab sharded run equals the single-slab
run (rock2)."""
import os
import threading

import cv2
import numpy as np
import pytest

import nonlocal_image_edit_b200 as nb
from nle_testlib import GOLDEN

pytestmark = pytest.mark.gpu


def _properties(f, L, k_req):
    inf = f.info()
    assert inf.p == 2500 and 1 <= inf.k <= k_req and inf.r2 <= inf.r <= inf.p
    S = f.eigvals
    assert np.all(np.diff(S) <= 1e-12) and abs(S[0] - 1.0) < 5e-3 and S[-1] > 0
    z = L.astype(np.float64)
    ones = np.ones(S.size)
    p1 = f.apply(z, ones)
    p2 = f.apply(p1, ones)
    assert np.abs(p2 - p1).max() <= 2e-2 * np.abs(p1).max()
    a = f.apply(z, nb.transformEigenValues(S, [4.0, 3.0, 4.0, 1.0]))
    b = f.apply(z, nb.transformEigenValues(S, [1.0, 1.0, 1.0, 1.0]))
    c = f.apply(z, nb.transformEigenValues(S, [7.0, 5.0, 7.0, 1.0]))
    assert np.allclose(c, 2 * a - b, atol=1e-6 * max(1.0, np.abs(a).max()))
    out = f.enhanceLuminance(L, [4.0, 3.0, 4.0, 1.0])
    assert out.shape == L.shape and out.dtype == np.uint8
    ref = np.clip(np.rint(a), 0, 255).astype(np.uint8).reshape(L.shape)
    assert (np.abs(out.astype(int) - ref.astype(int)) <= 1).all()           # fused clamp/round == host clamp/round (ties aside)


def test_rock2_full_resolution_p2500_k100():
    img = cv2.imread(os.path.join(GOLDEN, "rock2_input.png"))
    L = np.ascontiguousarray(cv2.cvtColor(img, cv2.COLOR_BGR2Lab)[:, :, 0])
    assert L.shape == (584, 876)
    args = (50, 50, 500.0, 10.0, 50, 100)
    f = nb.NLEFilter().trainFilter(L, *args)
    _properties(f, L, 100)
    # the same image as two row slabs reduced through the all-reduce hook (host threads, one GPU)
    import torch
    bar = threading.Barrier(2)
    stash = [None, None]
    res, err = [None, None], []

    class Arr:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 3}

    def run(rank):
        try:
            def allreduce(ptr, count, stream):
                t = torch.as_tensor(Arr(ptr, count), device="cuda")
                stash[rank] = t.clone()
                bar.wait()
                tot = stash[0] + stash[1]
                bar.wait()
                t.copy_(tot)
                torch.cuda.synchronize()
                bar.wait()
            g = nb.NLEFilter().trainFilter(L, *args, shard=((0, 300, allreduce) if rank == 0 else (300, 584, allreduce)))
            res[rank] = (g.eigvals, g.enhanceLuminance(L[:300] if rank == 0 else L[300:], [4.0, 3.0, 4.0, 1.0]), g)
        except Exception as e:   # pragma: no cover
            err.append(e)
            bar.abort()
    th = [threading.Thread(target=run, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not err, err
    assert np.allclose(res[0][0], f.eigvals, rtol=1e-7) and np.allclose(res[1][0], f.eigvals, rtol=1e-7)
    out = np.vstack([res[0][1], res[1][1]])
    full = f.enhanceLuminance(L, [4.0, 3.0, 4.0, 1.0])
    d = np.abs(out.astype(int) - full.astype(int))
    assert d.max() <= 1 and (d != 0).mean() <= 1e-3


def test_synthetic_16mp_p2500_k100():
    from bench import synth_luminance
    L = synth_luminance(4096, 4096)
    f = nb.NLEFilter().trainFilter(L, 50, 50, 500.0, 30.0, 20, 100)
    _properties(f, L, 100)
