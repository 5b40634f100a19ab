#!/bin/bash
# Graph-WaveNet microbenchmark grid (forward and forward+backward, both precisions), one JSON line per point.
#   bash scripts/c4_grid.sh > gpurun_out/r2_c4_grid.jsonl
for prec in bf16 fp32; do
  for B in 32 128 512 1024; do
    for C in 32 64 128 256; do
      for V in 9 10 42 43; do
        # keep the sweep bounded: the large-batch corners only for the HOP channel count
        if [ $B -ge 512 ] && [ $C -ne 64 ]; then continue; fi
        if [ $prec = fp32 ] && [ $B -ge 512 ]; then continue; fi
        timeout 120 python scripts/bench_ops.py gwnet --precision $prec --B $B --V $V --C $C --iters 10 2>/dev/null | tail -2
      done
    done
  done
done
