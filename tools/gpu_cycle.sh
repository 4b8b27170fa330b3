#!/bin/bash
# one GPU round trip: parity tests, a short bench, then (only if both exited 0) one ncu capture of the step kernel
# usage: bash tools/gpu_cycle.sh <tag> [workload]
tag=${1:-x}; wl=${2:-pick}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$tag.log 2>&1; rc=$?; tail -3 gpurun_out/pytest_$tag.log
[ $rc -ne 0 ] && exit $rc
timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-her 2>gpurun_out/bench_$tag.err | tail -1 > gpurun_out/bench_$tag.json || exit 1
python -c "import json; d=json.load(open('gpurun_out/bench_$tag.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])" || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mcb_env_kernel -s 150 -c 1 -o gpurun_out/prof_$tag -f python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-her > gpurun_out/ncu_$tag.log 2>&1 || { tail -5 gpurun_out/ncu_$tag.log; exit 1; }
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_raw_$tag.csv
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/prof_src_$tag.csv
echo profiled
