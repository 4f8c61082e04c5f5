#!/bin/bash
# tuning sweep over env knobs / bench configs; usage: scripts/sweep.sh "<bench args>" "ENV=1 ENV2=2" ...   ("-" = no env)
BARGS=$1; shift
for e in "$@"; do
  [ "$e" = "-" ] && e="X=0"
  OUT=$(env $e timeout 300 python bench.py $BARGS --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1)
  python - "$OUT" "$BARGS $e" <<'PY'
import sys, json
try:
    d = json.loads(sys.argv[1])
    print(f"{sys.argv[2]:70s} kernel_ms={d['roofline']['kernel_ms']:8.3f} ms_step={d['ms_per_step']:8.3f} qps={d['value']:10.1f} frac={d['roofline']['frac']:.3f} ref={d['config']['refined_queries']} fb={d['config']['fallback_queries']} clk={d['clocks']['sm_mhz']}")
except Exception as e:
    print(sys.argv[2], "FAILED", sys.argv[1][-300:])
PY
done
