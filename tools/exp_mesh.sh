cd /root/repo
bash tools/exp_bench.sh pick 2>&1 | tail -3
for wl in push mocap; do for mc in 0 1; do
  MCB_MESH=$mc python bench.py --workload $wl --no-her --no-cpu-baseline --steps 10 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl', 'mesh', $mc, round(d['value']), round(d['ms_per_step'],2), d['episode_stats']['fallback_envs_last_step'], d['episode_stats']['last_tier_envs_last_step'], d['episode_stats']['row_overflows'])"
done; done
