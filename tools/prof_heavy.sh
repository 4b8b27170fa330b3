#!/bin/bash
# ncu --set full captures of the contact-rich regime: the common-layout kernel of the push workload (full lockstep) and the
# last-tier kernel of the IK workload (every launch() issues the three tier kernels in order, so the skip count mod 3 picks the tier;
# 300 = past the reset, the autotune and the pre-roll); summaries via tools/ncu_summary.py
cd "$(dirname "$0")/.."
tag=${1:-r02f}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mcb_env_kernel -s 300 -c 1 -o gpurun_out/prof_${tag}_push -f python bench.py --workload push --steps 3 --warmup 3 --no-cpu-baseline --no-her > gpurun_out/ncu_${tag}_push.log 2>&1 || tail -5 gpurun_out/ncu_${tag}_push.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mcb_env_kernel -s 302 -c 1 -o gpurun_out/prof_${tag}_ik2 -f python bench.py --workload ik --steps 3 --warmup 3 --no-cpu-baseline --no-her > gpurun_out/ncu_${tag}_ik2.log 2>&1 || tail -5 gpurun_out/ncu_${tag}_ik2.log
for k in push ik2; do ncu -i gpurun_out/prof_${tag}_$k.ncu-rep --page raw --csv > gpurun_out/prof_raw_${tag}_$k.csv; done
ls -la gpurun_out/prof_${tag}_*.ncu-rep
