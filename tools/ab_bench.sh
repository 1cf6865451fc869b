#!/bin/bash
# A/B a tuning environment variable in the full benchmark:  tools/ab_bench.sh VAR v1 v2 ...
var=$1; shift
for v in "$@" "$@"; do
  env $var=$v python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/ab.json
  python - "$var" "$v" <<'PY'
import json, sys
d = json.load(open('/tmp/ab.json'))
print(sys.argv[1], sys.argv[2], round(d["ms_per_step"], 3), "ms/step  e2e", round(d["e2e"]["value"] / 1e6, 2), "M/s",
      [(k["kernel"], k["ms_per_step"]) for k in d["kernels"][:3]])
PY
done
