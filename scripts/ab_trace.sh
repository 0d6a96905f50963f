#!/bin/bash
# Same-box A/B of k_trace variants: prints "name | trace_ms shade_ms rt_pass_ms" for each (env/lib) combination. Run on the GPU box.
run() { name=$1; shift; out=$(env "$@" python scripts/profile_frame.py 4 2>&1 | tail -1); echo "$name | $(echo "$out" | python -c "
import sys,re
s=sys.stdin.read()
g=lambda k: re.search(r\"'%s': ([0-9.]+)\" % k, s)
print(' '.join('%s=%.2f' % (k, float(g(k).group(1))) if g(k) else k+'=?' for k in ('trace_ms','shade_ms','tail_ms','primary_ms','rt_pass_ms','tlas_update_ms')))")"; }
A=$PWD/build_ab
run "n4 base"            GKCU_NODE_FORMAT=4
run "n8 tmin"            GKCU_NODE_FORMAT=8
run "n8 notmin"          GKCU_NODE_FORMAT=8 GKCU_LIB=$A/n8notmin/libgknext_cuda.so
run "n8 tmin leaf3"      GKCU_NODE_FORMAT=8 GKCU_LEAF_MAX=3
run "n8 tmin leaf4"      GKCU_NODE_FORMAT=8 GKCU_LEAF_MAX=4
run "n8 tmin refill12"   GKCU_NODE_FORMAT=8 GKCU_REFILL=12
run "n8 tmin refill20"   GKCU_NODE_FORMAT=8 GKCU_REFILL=20
run "n4 rep2"            GKCU_NODE_FORMAT=4 GKCU_LIB=$A/rep2/libgknext_cuda.so
run "n8 rep2"            GKCU_NODE_FORMAT=8 GKCU_LIB=$A/rep2/libgknext_cuda.so
run "n4 base again"      GKCU_NODE_FORMAT=4
