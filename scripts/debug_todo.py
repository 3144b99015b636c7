"""Which pixels of the benchmark pair go through the exact thresholded pass, and why?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import depthmatch as dm
from synth import make_pair
in1, in2, flow = make_pair(10, 360, 640, 33, 33, seed=1234, noise=0.05)
ctx = dm.Context(0)
ctx.set_option("debug_todo", 1)
got = dm.match_extract(torch.from_numpy(in1).cuda(), torch.from_numpy(in2).cuda(), 33, 33, want=("index", "pmax", "score_thr", "index_thr", "min_ssd"), ctx=ctx)
sc = got["score_thr"].cpu().numpy(); pm = got["pmax"].cpu().numpy(); ms = got["min_ssd"].cpu().numpy()
todo = sc == -1
print("todo pixels", todo.sum(), "of", todo.size)
ys, xs = np.nonzero(todo)
print("x mod 4 histogram", np.bincount(xs % 4), " x mod 128:", np.bincount(xs % 128)[:8], "...")
print("y mod 15 histogram", np.bincount(ys % 15))
print("flow dy of todo:", np.bincount(flow[0][todo] + 16, minlength=33))
print("flow dx of todo:", np.bincount(flow[1][todo] + 16, minlength=33))
print("flow dx of all :", np.bincount(flow[1].reshape(-1) + 16, minlength=33))
print("pmax of todo: min %.6f max %.6f" % (pm[todo].min(), pm[todo].max()), " min_ssd of todo: mean %.4f" % ms[todo].mean())
