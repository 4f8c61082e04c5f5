#!/bin/bash
run() {
  OUT=$(env "$@" timeout 300 python bench.py --config ${CFG:-c2} --steps 4 --warmup 2 --no-cpu-baseline 2>&1 | tail -1)
  python - "$OUT" "$*" <<'PY'
import sys, json
d = json.loads(sys.argv[1])
print(f"{sys.argv[2]:55s} kernel_ms={d['roofline']['kernel_ms']:8.3f} ms_step={d['ms_per_step']:8.3f} qps={d['value']:9.1f} fb={d['config']['fallback_queries']}")
PY
}
run FENIX_TC_ORDER=0 FENIX_TC_PUBLISH=0
run FENIX_TC_ORDER=0 FENIX_TC_PUBLISH=1
run FENIX_TC_ORDER=1 FENIX_TC_PUBLISH=1
run FENIX_TC_ORDER=1 FENIX_TC_PUBLISH=0
run FENIX_TC_ORDER=0 FENIX_TC_PUBLISH=0 FENIX_TC_SLICES=2
run FENIX_TC_ORDER=0 FENIX_TC_PUBLISH=0 FENIX_TC_SLICES=5
run FENIX_TC_ORDER=0 FENIX_TC_PUBLISH=0 FENIX_TC_SLICES=1
