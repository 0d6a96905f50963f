"""Condenses one bench.py JSON line from stdin (also appends the full line to gpurun_out/bench_lines.jsonl)."""
import json, os, sys
l = sys.stdin.read().strip().splitlines()
l = l[-1] if l else ""
try:
    d = json.loads(l)
except Exception:
    print("NOT JSON:", l[-1200:]); sys.exit(0)
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/bench_lines.jsonl", "a").write(l + "\n")
mg = d.get("multi_gpu") or {}
print(f"n={d['n_gpus']} {d['config'].get('width')}x{d['config'].get('height')} spp/gpu={d['config'].get('spp_per_gpu')} value={d['value']:.1f} {d['unit']} ms/step={d['ms_per_step']:.3f} "
      f"e2e={d['e2e']['value']:.1f} exchange+post={mg.get('exchange_resolve_post_ms_per_frame_rank0', 0):.3f} ms [{mg.get('exchange', '')[:40]}] "
      f"rt={d['passes_ms_per_frame_rank0']['rt_pass_ms']:.2f} clocks={d['clocks']['sm_mhz'] if d.get('clocks') else None}")
