#!/bin/bash
# usage: epi_sweep2.sh "cfg cfg ..." "ENV=1 ENV2=2" "ENV=..." ...   (each further argument is one env set; "-" = none)
run() {
  OUT=$(env $2 timeout 300 python bench.py --config $1 --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1)
  python - "$OUT" "$1 $2" <<'PY'
import sys, json
try:
    d = json.loads(sys.argv[1])
    print(f"{sys.argv[2]:60s} kernel_ms={d['roofline']['kernel_ms']:8.3f} ms_step={d['ms_per_step']:8.3f} qps={d['value']:10.1f} frac={d['roofline']['frac']:.3f} ref={d['config']['refined_queries']} fb={d['config']['fallback_queries']} clk={d['clocks']['sm_mhz']}")
except Exception as e:
    print(sys.argv[2], "FAILED", sys.argv[1][-300:])
PY
}
CFGS=$1; shift
for CFG in $CFGS; do
  for e in "$@"; do
    if [ "$e" = "-" ]; then run $CFG "X=0"; else run $CFG "$e"; fi
  done
done
