#!/bin/bash
# env-steps/s per lockstep grouping (MCB_LOCKSTEP) and workload
cd "$(dirname "$0")/.."
for wl in ${WLS:-push ik mocap pick}; do for lw in ${LWS:-2 4 8 16}; do
  MCB_LOCKSTEP=$lw python bench.py --workload $wl --no-her --no-cpu-baseline --steps 10 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl', 'requested', $lw, 'lockstep', d['config']['lockstep_warps'], round(d['value']), round(d['ms_per_step'],2))"
done; done
