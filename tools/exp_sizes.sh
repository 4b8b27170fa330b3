cd /root/repo
for n in 1 17 1000 262144 1048576; do
  timeout 600 python bench.py --envs-per-gpu $n --steps 3 --warmup 3 --preroll 5 --no-cpu-baseline --no-her 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($n, round(d['value']), round(d['ms_per_step'],3), d['episode_stats']['env_steps'], d['episode_stats']['row_overflows'])" || echo "$n failed"
done
