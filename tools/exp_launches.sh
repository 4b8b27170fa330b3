#!/bin/bash
# per-tier kernel durations of one bench step per workload (ncu launch list; serialised and cold-cache, shares only)
cd "$(dirname "$0")/.."
for wl in ${@:-push grasp}; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mcb_env_kernel -c 24 --csv --log-file gpurun_out/launch_$wl.csv python bench.py --workload $wl --steps 2 --warmup 6 --no-cpu-baseline --no-her > gpurun_out/launch_$wl.log 2>&1
  python - "$wl" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(f"gpurun_out/launch_{sys.argv[1]}.csv")) if len(r) > 5 and r[0].isdigit()]
print(sys.argv[1], [(r[4].split("<")[1][:1], r[-1]) for r in rows][-9:])
PY
done
