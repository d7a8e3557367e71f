#!/bin/bash
# usage: tools/sweep_env.sh VAR v1 v2 ...   -> bench value / ms per step / stage times for every value of the env var
VAR=$1; shift
for V in "$@"; do
  env $VAR=$V timeout 300 python bench.py --steps 20 --warmup 3 2>/dev/null > /tmp/_sweep.json
  python - "$VAR=$V" <<'PY'
import json, sys
d = json.load(open("/tmp/_sweep.json"))
print(sys.argv[1], round(d["value"], 1), "GB/s", round(d["ms_per_step"], 4), "ms", d["roofline"]["stage_ms"])
PY
done
