#!/bin/bash
# Evidence run (one GPU): bench line, ncu launch list of the same command (eager), `ncu --set full` of every fused tcgen05
# kernel (one encoder layer forward + backward at C3 size, one scaler convolution), per-CTA pipeline timelines.
# Outputs under gpurun_out/<tag>_*; summarise into profiles/ with tools/summarize_launches.py and tools/ncu_summary.py.
tag=${1:-r02z}
python bench.py --steps 20 --warmup 5 2>gpurun_out/${tag}_bench.err | tail -1 > gpurun_out/${tag}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 1400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-parts > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'enc_|conv3x3_tc|conv_split|gemm_tc' -s 10 -c 14 -f \
    -o gpurun_out/${tag}_fused python tools/run_layer_once.py > gpurun_out/${tag}_fused.log 2>&1
ncu -i gpurun_out/${tag}_fused.ncu-rep --page raw --csv > gpurun_out/${tag}_fused.raw.csv 2>/dev/null
rm -f gpurun_out/${tag}_fused.ncu-rep      # ~100 MB with sources; gpurun_out/ is capped at 64 MiB -- the raw page is what gets summarised
python tools/trace_fused.py --warm > gpurun_out/${tag}_trace_fwd.txt 2>&1
python tools/trace_fused.py --bwd > gpurun_out/${tag}_trace_bwd.txt 2>&1
ls -la gpurun_out | tail -12
