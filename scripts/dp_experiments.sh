#!/bin/bash
# NCCL / bucket knobs of the data-parallel step on N GPUs (default 8):  bash scripts/dp_experiments.sh [N]
N=${1:-8}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 "$@" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d[\"value\"], d[\"ms_per_step\"], d[\"dp\"])"; }
echo base; run
echo nch4; NCCL_MAX_NCHANNELS=4 run
echo nch8; NCCL_MAX_NCHANNELS=8 run
echo no-overlap; run --dp-overlap 0
echo bucket100; run --dp-bucket-mb 100
echo n1; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-cuda --no-expressive 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d[\"value\"], d[\"ms_per_step\"])"
