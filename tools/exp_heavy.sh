cd /root/repo
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for lw in 1 16; do
for wl in grasp push mocap ik; do
  MCB_LOCKSTEP=$lw MCB_GRASP_NOISE=0.002 python bench.py --workload $wl --no-her --no-cpu-baseline --steps 10 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl', 'lockstep', d['config']['lockstep_warps'], round(d['value']), round(d['ms_per_step'],2), d['episode_stats']['fallback_envs_last_step'], d['episode_stats']['last_tier_envs_last_step'])"
done; done
