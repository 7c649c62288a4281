#!/bin/bash
# round 2, second 2-GPU call: the NCCL world-2 parity test (fixed inputs) and the stdout contract of the sharded bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dist_gpu.py -q -s -m gpu > gpurun_out/r2q_dist_gpu.log 2>&1; echo "dist test rc=$?"; tail -4 gpurun_out/r2q_dist_gpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2q_2gpu.json 2> gpurun_out/r2q_2gpu.err; echo "2gpu rc=$? stdout lines: $(wc -l < gpurun_out/r2q_2gpu.json)"
python -c "
import json; d=json.load(open('gpurun_out/r2q_2gpu.json')); print('sharded: ms/step', round(d['ms_per_step'],2), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
