#!/bin/bash
# (filter launches alternate sample prepass / main pass: launch index 5 is the main pass of the third search)
# One GPU call that refreshes the round's evidence: GPU tests, smoke, the default bench line, ncu captures of the dominant
# kernels (C3 CTA-pair main pass, one C4 shard, C5 batch 8), the C1 Flight / latency numbers. Outputs under gpurun_out/.
# usage: profile_pass.sh <prefix> <stage>   stage = runs | ncu1 | ncu2 | all   (one gpurun call brings back at most 64 MiB:
# the four ncu reports are ~20 MB each, so they travel in two calls)
R=${1:-r02}
S=${2:-all}
if [ "$S" = runs ] || [ "$S" = all ]; then
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${R}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err; echo "bench rc=$?"
fi
B="python bench.py --steps 2 --warmup 2 --no-also --no-cpu-baseline --no-parity"
if [ "$S" = ncu1 ] || [ "$S" = all ]; then
ncu --set full --clock-control none --import-source on -k regex:knn_tc_filter_kernel -s 5 -c 1 -o gpurun_out/${R}_prof_c3 $B > gpurun_out/${R}_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:knn_rq_filter_kernel -s 5 -c 1 -o gpurun_out/${R}_prof_c4s $B --config c4s > gpurun_out/${R}_ncu_c4s.log 2>&1; echo "ncu c4s rc=$?"
fi
if [ "$S" = ncu2 ] || [ "$S" = all ]; then
ncu --set full --clock-control none --import-source on -k regex:knn_tc_filter_kernel -s 5 -c 1 -o gpurun_out/${R}_prof_c5_8 $B --config c5_8 > gpurun_out/${R}_ncu_c5_8.log 2>&1; echo "ncu c5_8 rc=$?"
ncu --set full --cache-control none --clock-control none --import-source on -k regex:knn_direct -s 4 -c 1 -o gpurun_out/${R}_prof_c1_direct python scripts/ubench/c1_one.py > gpurun_out/${R}_ncu_c1.log 2>&1; echo "ncu c1 rc=$?"
fi
if [ "$S" = runs ] || [ "$S" = all ]; then
timeout 300 python scripts/bench_flight.py > gpurun_out/${R}_flight_c1.txt 2> gpurun_out/${R}_flight_c1.err; echo "flight rc=$?"
timeout 300 python scripts/latency_c1.py --c5 > gpurun_out/${R}_latency_c1.txt 2>&1; echo "latency rc=$?"
# launch lists (every kernel of a short bench run, serialised): the dominant kernel's share of the step
L="ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv"
$L --log-file gpurun_out/${R}_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-also --no-cpu-baseline --no-parity > /dev/null 2>&1; echo "launches c3 rc=$?"
$L --log-file gpurun_out/${R}_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-also --no-cpu-baseline --no-parity --config c2 > /dev/null 2>&1; echo "launches c2 rc=$?"
$L --log-file gpurun_out/${R}_launches_c1.csv python scripts/ubench/c1_one.py > /dev/null 2>&1; echo "launches c1 rc=$?"
# batched IVF
timeout 300 python scripts/bench_ivf.py > gpurun_out/${R}_ivf_batched.json 2>&1; echo "ivf rc=$?"
fi
ls -la gpurun_out/${R}_*
