#!/bin/bash
# CLIP bench over tower chunk sizes (development aid)
for c in "$@"; do
  timeout -s KILL 120 python bench.py --steps 40 --warmup 5 --no-cpu --chunk $c 2>/dev/null | tail -1 > /tmp/_line.json
  python - "$c" <<'PY'
import json, sys
d = json.loads(open('/tmp/_line.json').read())
print(sys.argv[1], round(d["value"]), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], round(d["roofline"]["frac"], 3), d["roofline"].get("eager_ms_per_step_by_kernel"))
PY
done
