#!/bin/bash
# Same-box A/B of k_trace variants: "name | kernel times of the last of 4 bench frames". Run on the GPU box: bash scripts/ab_trace.sh name=ENV,ENV...
run() { name=$1; shift; out=$(env "$@" python scripts/profile_frame.py 4 2>&1 | tail -1); echo "$name | $(echo "$out" | python -c "
import sys,re
s=sys.stdin.read()
g=lambda k: re.search(r\"'%s': ([0-9.]+)\" % k, s)
print(' '.join('%s=%.2f' % (k, float(g(k).group(1))) if g(k) else k+'=?' for k in ('trace_ms','shade_ms','tail_ms','primary_ms','rt_pass_ms','tlas_update_ms')))")"; }
A=$PWD/build_ab
for spec in "$@"; do
  name=${spec%%=*}; envs=${spec#*=}
  IFS=',' read -ra E <<< "$envs"
  EV=()
  for e in "${E[@]}"; do EV+=("${e//@A/$A}"); done
  run "$name" "${EV[@]}"
done
