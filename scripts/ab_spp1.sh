#!/bin/bash
# Same-box A/B of the 1-spp frame (config 3's rt pass): bash scripts/ab_spp1.sh name=ENV,ENV...
for spec in "$@"; do
  name=${spec%%=*}; envs=${spec#*=}
  IFS=',' read -ra E <<< "$envs"
  out=$(env "${E[@]}" python scripts/profile_frame.py 6 0 1 2>&1 | tail -1)
  echo "$name | $(echo "$out" | python -c "
import sys,re
s=sys.stdin.read()
g=lambda k: re.search(r\"'%s': ([0-9.]+)\" % k, s)
print(' '.join('%s=%.3f' % (k, float(g(k).group(1))) if g(k) else k+'=?' for k in ('trace_ms','shade_ms','tail_ms','primary_ms','rt_pass_ms','tlas_update_ms','reproject_pass_ms','compose_pass_ms','wavefront_iterations')))")"
done
