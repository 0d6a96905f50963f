#!/bin/bash
# Same-box A/B of the 4K 64-spp frame (config 5 on one GPU): bash scripts/ab_4k.sh name=ENV,ENV...
for spec in "$@"; do
  name=${spec%%=*}; envs=${spec#*=}
  IFS=',' read -ra E <<< "$envs"
  out=$(env "${E[@]}" python scripts/profile_frame.py 3 0 64 3840 2160 2>&1 | tail -1)
  echo "$name | $(echo "$out" | python -c "
import sys,re
s=sys.stdin.read()
g=lambda k: re.search(r\"'%s': ([0-9.]+)\" % k, s)
print(' '.join('%s=%.2f' % (k, float(g(k).group(1))) if g(k) else k+'=?' for k in ('trace_ms','shade_ms','tail_ms','primary_ms','rt_pass_ms','tlas_update_ms','wavefront_iterations')))")"
done
