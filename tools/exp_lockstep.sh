cd /root/repo
for wl in push mocap ik; do
for lw in 16 8 4 2; do
  MCB_LOCKSTEP=$lw python bench.py --workload $wl --no-her --no-cpu-baseline --steps 10 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl', 'lockstep', d['config']['lockstep_warps'], round(d['value']), round(d['ms_per_step'],2))"
done; done
