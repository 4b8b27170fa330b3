#!/bin/bash
# bench every experimental build in exp_build/ (MCB_LIB override) plus the default library; prints env-steps/s
cd "$(dirname "$0")/.."
wl=${1:-pick}
shopt -s nullglob
for lib in "" exp_build/*.so; do
  if [ -n "$lib" ]; then export MCB_LIB=$PWD/$lib; else unset MCB_LIB; fi
  v=$(timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-her 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']))")
  echo "${lib:-default} $v"
done
