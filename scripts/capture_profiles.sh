#!/bin/bash
# One GPU-box pass that produces everything profiles/README.md cites (run from the repo root under gpurun; 1 GPU).
# Bench numbers are taken WITHOUT a profiler attached; the ncu passes only provide shares, traffic and counters.
set -x
O=gpurun_out
mkdir -p $O
python __graft_entry__.py smoke > $O/r2_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench_final.json 2> $O/r2_bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err; echo "reference rc=$?"
timeout 300 python scripts/bench_gemm.py $O/r2_gemm_bench.json > $O/r2_gemm_bench.txt 2>&1
timeout 300 python scripts/bench_bert_attn.py > $O/r2_bert_attn_bench.txt 2>&1
timeout 300 python scripts/bench_gru.py > $O/r2_gru_bench.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2_step_launches_final.csv python bench.py --profile-step > $O/ncu_a.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:gemm_tma --csv --log-file $O/traffic_step_gemm.csv python bench.py --profile-step > $O/ncu_b.log 2>&1; echo "gemm traffic rc=$?"
for g in gwnet_fwd gwnet_fwdbwd xattn_fwd xattn_fwdbwd; do
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/traffic_$g.csv python scripts/traffic_probe.py $g 2 > $O/ncu_c.log 2>&1; echo "traffic $g rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tma -s 4 -c 2 -o $O/r2_gemm_final -f python scripts/bench_gemm.py --only=out_proj > $O/ncu_d.log 2>&1; echo "gemm full rc=$?"
