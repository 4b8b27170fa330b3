#!/bin/bash
# one GPU call that refreshes the round's evidence: tests, bench lines of every workload (incl. the reference arm and the grasp
# workload), the launch list of the default bench command and one ncu --set full capture of the step kernel
# usage: bash tools/round_artifacts.sh <tag>
tag=${1:-r02x}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_$tag.log 2>&1; tail -3 gpurun_out/pytest_$tag.log
timeout 600 python bench.py > gpurun_out/bench_${tag}_pick.json 2> gpurun_out/bench_${tag}_pick.err
for wl in push reach ik mocap grasp; do
  timeout 600 python bench.py --workload $wl --no-her --no-cpu-baseline 2>gpurun_out/bench_${tag}_$wl.err | tail -1 > gpurun_out/bench_${tag}_$wl.json
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 | tail -1 > gpurun_out/bench_${tag}_reference.json
python - <<PYEOF
import json
for wl in ["pick", "push", "reach", "ik", "mocap", "grasp"]:
    try:
        d = json.loads(open(f"gpurun_out/bench_${tag}_{wl}.json").read().strip().splitlines()[-1])
        print(wl, "lockstep", d["config"].get("lockstep_warps"), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4),
              "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"]), "fallback", d["episode_stats"].get("fallback_envs_last_step"), d["episode_stats"].get("last_tier_envs_last_step"),
              "iters/substep", d["episode_stats"]["solver_iters_per_substep"])
    except Exception as e:
        print(wl, "failed", e)
d = json.loads(open("gpurun_out/bench_${tag}_reference.json").read().strip().splitlines()[-1])
print("reference arm", round(d["value"]), d["cpu_baseline"]["sample"])
PYEOF
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --preroll 0 --no-cpu-baseline > gpurun_out/ncu_launches_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mcb_env_kernel -s 150 -c 1 -o gpurun_out/prof_$tag -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-her > gpurun_out/ncu_$tag.log 2>&1 || tail -5 gpurun_out/ncu_$tag.log
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_raw_$tag.csv
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/prof_src_$tag.csv
echo artifacts done
