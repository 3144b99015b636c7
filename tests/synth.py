"""Seeded synthetic inputs shared by tests, smoke() and bench.py.

Frame 2 is N(0,1) noise; frame 1 is frame 2 displaced by a smooth integer flow
(|flow| <= window/2 - 1) plus N(0, noise^2), so the true match is well separated
and the planted flow is the known answer (the reference's
cartesian_groundtruth_cc_testme idea, radial/radial_opticalflow_groundtruth.lua:170-210).
"""
import numpy as np


def smooth_integer_flow(h, w, maxdy, maxdx, seed=4321):
    """Piecewise-smooth integer flow field (2,h,w) with |dy|<=maxdy, |dx|<=maxdx."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
    ph = rng.uniform(0, 2 * np.pi, 4)
    fy = maxdy * np.sin(2 * np.pi * (0.7 * yy + 0.4 * xx) + ph[0]) * np.cos(1.3 * np.pi * xx + ph[1])
    fx = maxdx * np.cos(2 * np.pi * (0.5 * xx - 0.6 * yy) + ph[2]) * np.sin(1.1 * np.pi * yy + ph[3])
    return np.stack([np.rint(fy), np.rint(fx)]).astype(np.int64)


def make_pair(C, H2, W2, maxh, maxw, seed=1234, noise=0.05, flow_seed=4321):
    """Returns (in1 [C,H1,W1], in2 [C,H2,W2], flow [2,H1,W1]) with
    in1[:, y, x] = in2[:, y + cy-1 + fy, x + cx-1 + fx] + noise, i.e. the reference's
    winner index decodes to (fy, fx) (centre = ceil(max/2), opticalflow_model.lua:208-212)."""
    rng = np.random.default_rng(seed)
    in2 = rng.standard_normal((C, H2, W2), dtype=np.float32)
    H1, W1 = H2 - maxh + 1, W2 - maxw + 1
    cy, cx = (maxh + 1) // 2, (maxw + 1) // 2
    # displacement index range: dy in [0, maxh) -> flow in [-(cy-1), maxh-cy]
    my = max(0, min(cy - 1, maxh - cy) - 1) if maxh > 2 else 0
    mx = max(0, min(cx - 1, maxw - cx) - 1) if maxw > 2 else 0
    flow = smooth_integer_flow(H1, W1, my, mx, flow_seed)
    yy, xx = np.meshgrid(np.arange(H1), np.arange(W1), indexing="ij")
    sy = yy + cy - 1 + flow[0]
    sx = xx + cx - 1 + flow[1]
    in1 = in2[:, sy, sx] + noise * rng.standard_normal((C, H1, W1), dtype=np.float32)
    return np.ascontiguousarray(in1, np.float32), in2, flow
