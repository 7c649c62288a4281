#!/bin/bash
# round 2, last 2-GPU call: NCCL world-2 parity test and the sharded bench (NCCL scatter / gather inside the timed region,
# pipelined D2H on rank 0 in the end-to-end leg) on the final build; reference arm under torchrun (rank 0 only works)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dist_gpu.py -q -s -m gpu > gpurun_out/r3c_dist_gpu.log 2>&1; echo "dist test rc=$?"; tail -3 gpurun_out/r3c_dist_gpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r3c_2gpu.json 2> gpurun_out/r3c_2gpu.err; echo "2gpu rc=$? stdout lines: $(wc -l < gpurun_out/r3c_2gpu.json)"
python -c "
import json; d=json.load(open('gpurun_out/r3c_2gpu.json')); print('sharded: ms/step', round(d['ms_per_step'],2), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'per rank', [(r['rank'], round(r['ms_per_step'],2), r['valid_frames']) for r in d['per_rank']])" || tail -20 gpurun_out/r3c_2gpu.err
timeout 200 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('1 gpu same box: ms/step', round(d['ms_per_step'],2), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r3c_ref_2gpu.json 2> gpurun_out/r3c_ref_2gpu.err; echo "reference arm under torchrun rc=$? lines $(wc -l < gpurun_out/r3c_ref_2gpu.json)"; head -c 300 gpurun_out/r3c_ref_2gpu.json; echo
