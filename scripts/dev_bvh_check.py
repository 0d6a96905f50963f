import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gknextengine_b200 import CudaPathTracingRenderer, host
from gknextengine_b200.bvh_check import validate
for builder in ("lbvh", "sah"):
    os.environ["GKCU_BLAS_BUILDER"] = builder
    for name in ("CornellBox.proc", "RTIO.proc", "LegoBricks.synth:2000"):
        s = host.Scene(name); s.update_nodes()
        r = CudaPathTracingRenderer(0); r.CreateSwapChain((64, 64)); r.OnPostLoadScene(s); r.PreRender()
        rep = validate(r)
        print(builder, name)
        for k, v in rep.items():
            if k != "info": print("   ", k, v)
        r.close()
