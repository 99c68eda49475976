"""one grid size, a few steps: used under ncu to list per-launch times of a mid-size cube"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alis_b200 import capi, synthetic
n = int(sys.argv[1]); steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = synthetic.BoxWorld(n, n, n)
e = w.make_engine(capi.load_engine())
e.run(steps); e.synchronize()
