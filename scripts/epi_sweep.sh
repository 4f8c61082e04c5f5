#!/bin/bash
# epilogue tuning sweep: library variants (fenix_b200/variants/*.so) x configs; usage: epi_sweep.sh "c2 c4s" variant...
run() {
  OUT=$(env "$@" timeout 300 python bench.py --config $CFG --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1)
  python - "$OUT" "$CFG $*" <<'PY'
import sys, json
try:
    d = json.loads(sys.argv[1])
    print(f"{sys.argv[2]:70s} kernel_ms={d['roofline']['kernel_ms']:8.3f} ms_step={d['ms_per_step']:8.3f} qps={d['value']:10.1f} frac={d['roofline']['frac']:.3f} ref={d['config']['refined_queries']} fb={d['config']['fallback_queries']} clk={d['clocks']['sm_mhz']}")
except Exception as e:
    print(sys.argv[2], "FAILED", sys.argv[1][-300:])
PY
}
V=fenix_b200/variants
CFGS=$1; shift
for CFG in $CFGS; do
  run X=0
  for v in "$@"; do
    run FENIX_KNN_LIB=$PWD/$V/$v.so
  done
done
