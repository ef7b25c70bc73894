#!/bin/bash
# N-GPU check of the peer-memory all-reduce: correctness + latency script, then the bench line (weak scaling, parity vs unsharded).
N=${1:-2}
mkdir -p gpurun_out
NLE_B200_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   scripts/gpu_peer_allreduce_check.py --iters 300 --json gpurun_out/peer_check_n$N.json > gpurun_out/peer_check_n$N.log 2>&1
echo "peer check rc=$?"; tail -5 gpurun_out/peer_check_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
   bench.py --gpus $N > gpurun_out/bench_r2zj_n$N.json 2> gpurun_out/bench_r2zj_n$N.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_r2zj_n$N.err
python scripts/show_bench.py gpurun_out/bench_r2zj_n$N.json | head; python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2zj_n$N.json").read().strip().splitlines()[-1])
print(d.get("collectives")); print(d.get("multi_gpu_parity"))
PY
if [ -n "$SKIP_NCCL_ARM" ]; then exit 0; fi
NLE_B200_PEER_AR=off timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
   bench.py --gpus $N --no-targets > gpurun_out/bench_r2zj_n${N}_nccl.json 2> gpurun_out/bench_r2zj_n${N}_nccl.err
echo "bench (NCCL small messages) rc=$?"
python scripts/show_bench.py gpurun_out/bench_r2zj_n${N}_nccl.json | head -3
