run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 "$@" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d[\"value\"], d[\"ms_per_step\"], d[\"dp\"])"; }
echo base; run
echo no-overlap; run --dp-overlap 0
echo no-overlap-1bucket; run --dp-overlap 0 --dp-bucket-mb 400
echo overlap-nch2; NCCL_MAX_NCHANNELS=2 run
echo overlap-nch4; NCCL_MAX_NCHANNELS=4 run
echo n1; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-cuda --no-expressive 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d[\"value\"], d[\"ms_per_step\"])"
