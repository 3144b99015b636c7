import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
import numpy as np, torch
import depthmatch as dm
rng = np.random.default_rng(0)
flt = dm.getFilter(dm.Geometry(layers=[[3, 5, 5, 8], [4, 16, 16, 10]]), rng)
x = torch.rand((2, 3, 360, 640), device="cuda")
flt.forward(x); torch.cuda.synchronize()
dm.default_context().set_option("volume_debug", 9); dm.default_context().set_option("conv", 2)
flt.forward(x); torch.cuda.synchronize()
