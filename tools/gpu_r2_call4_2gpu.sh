#!/bin/bash
# round 2, 2-GPU call: NCCL scatter/gather test and the sharded bench (weak + strong) under torchrun
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests/test_dist_gpu.py -q -s -m gpu > gpurun_out/r2_dist_gpu.log 2>&1; echo "dist test rc=$?"; tail -3 gpurun_out/r2_dist_gpu.log
for mode in sharded replicas; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --dist $mode > gpurun_out/r2d_2gpu_$mode.json 2> gpurun_out/r2d_2gpu_$mode.err; echo "2gpu $mode rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2d_2gpu_$mode.json"))
    print("$mode: ms/step", round(d["ms_per_step"],2), "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["config"]["parallelism"]); print(d["per_rank"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2d_2gpu_$mode.err").read()[-1500:])
PY
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --workload multilingual_v3_b1024x128 > gpurun_out/r2d_2gpu_strong.json 2> gpurun_out/r2d_2gpu_strong.err; echo "2gpu strong rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2d_2gpu_strong.json')); print('strong: ms/step', round(d['ms_per_step'],2), 'value', round(d['value'],1), d['scaling'], d['per_rank'])" || tail -c 1500 gpurun_out/r2d_2gpu_strong.err
timeout 200 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('1 gpu same box: ms/step', round(d['ms_per_step'],2), 'value', round(d['value'],1))"
