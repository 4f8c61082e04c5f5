#!/bin/bash
# usage: scripts/sweep.sh "<bench args>" "ENV=.. ENV2=.." "ENV=.." ...   (one bench run per env set; "-" = no extra env)
# prints per run: env, ms/step, kernel ms, search ms, QPS, roofline fraction (of burst / sustained), SM clock, parity
ARGS="$1"; shift
for E in "$@"; do
  [ "$E" = "-" ] && E=""
  OUT=$(env $E python bench.py $ARGS --no-cpu-baseline --no-also 2>/dev/null | tail -1)
  python - "$E" <<PY
import json,sys
try:
    d=json.loads('''$OUT'''); r=d["roofline"]
    print("%-48s ms/step %8.3f kernel %8.3f search %8.3f qps %10.0f e2e %10.0f burst %.3f sust %.3f clk %s %s par=%s ref=%s fb=%s" % (sys.argv[1] or "-", d["ms_per_step"], r["kernel_ms"], r["search_device_ms"], d["value"], d["e2e"]["value"], r.get("frac_of_burst",0), r.get("frac_of_sustained",0), d["clocks"]["sm_mhz"], ",".join(d["clocks"]["reasons"]), d.get("parity",{}).get("ok"), d["details"]["refined_queries"], d["details"]["fallback_queries"]), flush=True)
except Exception as e:
    print(sys.argv[1], "FAILED", e, flush=True)
PY
done
