"""Parity accounting of the fused CUDA path against the CPU oracle, shared by the GPU tests and by
bench.py's `parity` block (test infrastructure: imports the oracle, never the other way round).

Bars (BASELINE.json north_star): integer indices bit-exact except where the oracle's top-2
soft-max scores differ by less than 1e-5 relative (those pixels are counted and logged); fp32
scores within 1e-4 relative.
"""
import math

import numpy as np

RTOL = 1e-4       # fp32 scores
NEAR_TIE = 1e-5   # relative top-2 gap below which an index may differ


def oracle_pair(oracle, in1, in2, maxh, maxw, thr=0.11, canvas=None, nthreads=0):
    """The reference path on one pair: SpatialMatching -> Minus -> SoftMax -> argmax + tie rule ->
    extractOutput(thr) -> canvas (opticalflow_model.lua:93-109,153-169,201-252)."""
    K = maxh * maxw
    shp = tuple(in1.shape[1:])
    vol = oracle.spatial_matching(in1, in2, maxh, maxw, nthreads=nthreads)
    prob = oracle.neg_softmax(vol, nthreads=nthreads)
    middle = (math.ceil(maxh / 2) - 1) * maxw + math.ceil(maxw / 2)
    idx, pmax = oracle.argmax_tie(prob, K, middle)
    ret, sc, _ = oracle.extract_output(prob.reshape(shp + (K,)), thr)
    gap = oracle.top2_relgap(prob, K).reshape(shp)
    out = dict(index=idx.reshape(shp), pmax=pmax.reshape(shp), index_thr=ret, score_thr=sc, gap=gap,
               min_ssd=vol.reshape(shp + (K,)).min(-1), middle=middle)
    # probabilities within 4e-5 relative of the threshold: extractOutput's decision may flip
    p3 = prob.reshape(shp + (K,))
    out["thr_edge"] = (np.abs(p3 - thr) < thr * 4e-5).any(-1)
    if canvas is not None:
        out["flow_full"] = oracle.flow_canvas(out["index"], shp[0], shp[1], maxh, maxw, canvas[0], canvas[1])
    return out


def _maxrel(got, want, floor=1e-30):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    return float((np.abs(got - want) / np.maximum(np.abs(want), floor)).max()) if want.size else 0.0


def parity_report(got, want):
    """Counts for one pair.  `got` holds whichever of index / pmax / score_thr / index_thr /
    flow_full / min_ssd the CUDA call produced."""
    tie = want["gap"] < NEAR_TIE
    rep = {"pixels": int(tie.size), "near_tie": int(tie.sum())}
    if "index" in got:
        diff = np.asarray(got["index"]) != want["index"]
        rep["bit_exact"] = int((~diff).sum())
        rep["mismatched"] = int((diff & ~tie).sum())
        rep["near_tie_differing"] = int((diff & tie).sum())
    if "pmax" in got:
        rep["pmax_max_rel"] = _maxrel(got["pmax"], want["pmax"])
    if "min_ssd" in got:
        # relative to max(SSD, 1e-3): a perfect match reads ~0 in both
        rep["min_ssd_max_rel"] = _maxrel(got["min_ssd"], want["min_ssd"], floor=1e-3)
    ok = ~(want["thr_edge"] | tie)
    if "score_thr" in got:
        rep["score_max_rel"] = _maxrel(np.asarray(got["score_thr"])[ok], want["score_thr"][ok], floor=1e-6)
    if "index_thr" in got:
        rep["index_thr_mismatched"] = int((np.asarray(got["index_thr"])[ok] != want["index_thr"][ok]).sum())
        rep["thr_edge"] = int(want["thr_edge"].sum())
    if "flow_full" in got and "flow_full" in want:
        # the canvas decodes `index`: it may differ exactly where the index does
        d = (np.asarray(got["flow_full"]) != want["flow_full"]).any(0)
        rep["flow_full_differing_px"] = int(d.sum())
    return rep


def assert_parity(rep, what=""):
    assert rep.get("mismatched", 0) == 0, "%s: %d index mismatches outside near-ties (%r)" % (
        what, rep["mismatched"], rep)
    assert rep.get("pmax_max_rel", 0.0) <= RTOL, "%s: pmax off by %.3g relative (%r)" % (
        what, rep["pmax_max_rel"], rep)
    assert rep.get("score_max_rel", 0.0) <= RTOL, "%s: score_thr off by %.3g relative (%r)" % (
        what, rep["score_max_rel"], rep)
    assert rep.get("index_thr_mismatched", 0) == 0, "%s: %r" % (what, rep)
    assert rep.get("min_ssd_max_rel", 0.0) <= RTOL, "%s: %r" % (what, rep)
    if "flow_full_differing_px" in rep:
        assert rep["flow_full_differing_px"] <= rep.get("near_tie_differing", 0), "%s: %r" % (what, rep)
