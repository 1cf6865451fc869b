#!/bin/bash
# Round-end evidence run (one GPU): bench line, ncu launch list of the same command (eager), ncu --set full of the
# dominant kernel on the four C3 GEMM shapes, per-CTA GEMM timelines.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r01z}
python bench.py --steps 20 --warmup 3 2>gpurun_out/${tag}_bench.err | tail -1 > gpurun_out/${tag}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 2800 -c 2400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${tag}_launches.log 2>&1
for shape in qkv lr2 "dx " dWqkv; do
  s=$(echo $shape | tr -d ' ')
  ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 5 -c 1 -f -o gpurun_out/${tag}_gemm_$s \
      python tools/bench_gemm.py "$shape" 4 > /dev/null 2>&1
  # keep the raw metric page (small); the .ncu-rep with sources is ~36 MB and gpurun_out/ is capped at 64 MiB
  ncu -i gpurun_out/${tag}_gemm_$s.ncu-rep --page raw --csv > gpurun_out/${tag}_gemm_$s.raw.csv 2>/dev/null
  rm -f gpurun_out/${tag}_gemm_$s.ncu-rep
done
python tools/trace_gemm.py > gpurun_out/${tag}_gemm_trace.txt 2>&1
ls -la gpurun_out | tail -12
