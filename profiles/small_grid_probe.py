"""steady-state cost per step of a small (launch-bound) grid through the engine C ABI"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alis_b200 import capi, synthetic
lib = capi.load_engine()
sizes = [int(a) for a in sys.argv[1:]] or [76, 128, 192, 256, 384, 512]
for shape in [(n, n, n) for n in sizes]:
    w = synthetic.BoxWorld(*shape)
    e = w.make_engine(lib)
    e.run(100); e.synchronize()
    l0 = e.launch_count(); t0 = time.perf_counter()
    e.run(500); 
    t_enq = time.perf_counter() - t0
    e.synchronize()
    dt = time.perf_counter() - t0
    print(shape, "us/step %.1f (host enqueue %.1f)" % (dt/500*1e6, t_enq/500*1e6), "launches/step %.1f" % ((e.launch_count()-l0)/500), "Mcells/s %.0f" % (w.cells()*500/dt*1e-6), "fused launches", e.fused_launch_count())
