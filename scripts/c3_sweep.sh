#!/bin/bash
# tuning sweep on C3: unit order x slice count x threshold publication; records kernel ms, clocks, power
run() {
  ( nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits -lms 100 > /tmp/pw.csv & echo $! > /tmp/pw.pid )
  OUT=$(env "$@" timeout 300 python bench.py --config ${CFG:-c3} --steps 4 --warmup 2 --no-cpu-baseline 2>&1 | tail -1)
  kill $(cat /tmp/pw.pid)
  python - "$OUT" "$*" <<'PY'
import sys, json, numpy as np
d = json.loads(sys.argv[1])
rows = [l.split(',') for l in open('/tmp/pw.csv') if l.count(',') == 1]
clk = np.array([float(r[0]) for r in rows]); pw = np.array([float(r[1]) for r in rows])
busy = pw > 0.6 * pw.max()
print(f"{sys.argv[2]:55s} kernel_ms={d['roofline']['kernel_ms']:8.3f} qps={d['value']:9.1f} clk_busy={np.median(clk[busy]):6.0f} pw_busy={np.median(pw[busy]):6.0f} fb={d['config']['fallback_queries']}")
PY
}
run FENIX_TC_ORDER=0 FENIX_TC_PUBLISH=0
run FENIX_TC_ORDER=0 FENIX_TC_PUBLISH=1
run FENIX_TC_ORDER=1 FENIX_TC_PUBLISH=1
run FENIX_TC_ORDER=1 FENIX_TC_PUBLISH=0
run FENIX_TC_ORDER=0 FENIX_TC_PUBLISH=0 FENIX_TC_SLICES=37
run FENIX_TC_ORDER=1 FENIX_TC_PUBLISH=1 FENIX_TC_SLICES=37
run FENIX_TC_ORDER=0 FENIX_TC_PUBLISH=0 FENIX_TC_SLICES=4
run FENIX_TC_ORDER=1 FENIX_TC_PUBLISH=1 FENIX_TC_SLICES=4
