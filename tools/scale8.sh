#!/bin/bash
# 8-GPU lines for profiles/ (one box, 8 ranks): C3 weak + strong; with "all" also C4 and C5 at the reference's per-GPU batch.
N=${1:-8}
what=${2:-c3}
run() { # name, extra args
  name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-parts "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  tail -n 1 gpurun_out/$name.json | cut -c1-400
}
run r02_bench_c3_n${N}_weak --config c3 --scaling weak
run r02_bench_c3_n${N}_strong --config c3 --scaling strong
if [ "$what" = "all" ]; then
  run r02_bench_c4_n${N} --config c4
  run r02_bench_c5_n${N} --config c5
fi
