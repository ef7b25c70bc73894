"""GPU (>= 2 devices): the library's peer-memory all-reduce (csrc/nccl_comm.cu, peer_allreduce_kernel) against the sum in
rank order of the all-gathered inputs, bit for bit, over a few hundred back-to-back reductions of random lengths; the NCCL
path of the same entry point for a message above the limit.  Runs scripts/gpu_peer_allreduce_check.py under torchrun on two
ranks; skipped on a single-GPU box (bench.py --gpus N checks the sharded result against the unsharded one there)."""
import json
import os
import subprocess
import sys

import pytest

from nle_testlib import ROOT

pytestmark = pytest.mark.gpu


def test_peer_allreduce_is_bit_identical_to_the_rank_ordered_sum(nb, tmp_path):
    if nb.load().nle_b200_device_count() < 2:
        pytest.skip("needs two GPUs")
    out = tmp_path / "peer.json"
    env = {k: v for k, v in os.environ.items() if not k.startswith("NLE_B200_")}
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "gpu_peer_allreduce_check.py"), "--iters", "200", "--json", str(out)]
    subprocess.run(cmd, env=env, check=True, timeout=600)
    res = json.loads(out.read_text())
    assert res["mismatching_reductions_all_ranks"] == 0
    assert res["nccl_path_rel_err"] <= 1e-14
    if res["peer_path"]:
        assert res["peer_calls"] >= 200
