#!/bin/bash
# A/B whole-bench runs over "ENV=VAL,ENV=VAL" settings:  tools/ab_variants.sh "A=1,B=2" "A=0" ...
for rep in 1 2; do
for v in "$@"; do
  env $(echo "$v" | tr ',' ' ') python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/ab.json
  python - "$v" <<'PY'
import json, sys
d = json.load(open('/tmp/ab.json'))
print(sys.argv[1], round(d["ms_per_step"], 3), "ms/step  e2e", round(d["e2e"]["value"] / 1e6, 2), "M/s")
PY
done
done
