"""dielectric world (material arrays everywhere): fused E+H stepping vs the two-pass path"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from alis_b200 import capi, synthetic
n = (1024, 1024, 384)
w = synthetic.BoxWorld(*n)
e = w.make_engine(capi.load_engine())
eps = [np.full((n[0] + (c == 0), n[1] + (c == 1), n[2] + (c == 2)), 1.0 / 2.25, np.float32) for c in range(3)]
e.set_inv_eps(*eps)
del eps
e.run(5); e.synchronize()
t0 = time.perf_counter(); e.run(40); e.synchronize(); dt = time.perf_counter() - t0
e.set_profiling(True); e.run(10); e.synchronize()
avg, nl, by = e.profile_dominant()
print("dominant kernel: %.3f ms per launch, %d launches, %.2f GB 'two-pass algorithmic' per launch -> %.0f GB/s" % (avg, nl, by*1e-9*84/72, by*84/72/(avg*1e-3)*1e-9))
print("fuse env", os.environ.get("ALISB_FUSE", "default"), "ms/step %.3f" % (dt/40*1e3), "Gcells/s %.1f" % (w.cells()*40/dt*1e-9), "fused launches", e.fused_launch_count())
