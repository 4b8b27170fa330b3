cd /root/repo
python -m pytest tests/test_gpu_parity.py tests/test_gpu_canary.py -x -q -m gpu 2>&1 | tail -2
for wl in ik mocap push grasp pick; do for lw in 0 1 16; do
  MCB_LOCKSTEP=$lw python bench.py --workload $wl --no-her --no-cpu-baseline --steps 10 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl', 'requested', $lw, 'lockstep', d['config']['lockstep_warps'], round(d['value']), round(d['ms_per_step'],2))"
done; done
