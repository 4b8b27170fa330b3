#!/bin/bash
# bench lines of every workload -> gpurun_out/bench_<tag>_<workload>.json
tag=${1:-x}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for wl in pick push reach ik mocap; do
  extra="--no-her"; [ $wl = pick ] && extra=""
  timeout 600 python bench.py --workload $wl $extra 2>gpurun_out/bench_${tag}_$wl.err | tail -1 > gpurun_out/bench_${tag}_$wl.json
  python - <<PYEOF
import json
d = json.load(open('gpurun_out/bench_${tag}_$wl.json'))
print('$wl', 'lockstep', d['config'].get('lockstep_warps'), round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'], 4),
      d['cpu_baseline'] and round(d['cpu_baseline']['value']), d['episode_stats']['row_overflows'], 'fallback', d['episode_stats'].get('fallback_envs_last_step'), d['episode_stats'].get('last_tier_envs_last_step'), d.get('her_relabel') and round(d['her_relabel']['ms'], 4))
PYEOF
done
