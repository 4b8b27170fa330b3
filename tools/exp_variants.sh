#!/bin/bash
# env-steps/s of every experimental build in exp_build/ (MCB_LIB override) and the default library, per workload, at MCB_LOCKSTEP=$LW
cd "$(dirname "$0")/.."
shopt -s nullglob
for wl in ${WLS:-push mocap ik pick}; do for lib in "" exp_build/*.so; do
  if [ -n "$lib" ]; then export MCB_LIB=$PWD/$lib; else unset MCB_LIB; fi
  MCB_LOCKSTEP=${LW:-16} timeout 300 python bench.py --workload $wl --no-her --no-cpu-baseline --steps 10 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl', '${lib:-default}', 'lockstep', d['config']['lockstep_warps'], round(d['value']), round(d['ms_per_step'],2))"
done; done
