"""Mirror of the reference's model / calibration files on top of torch7io:
opticalflow_model_io.lua:98-220 (saveModel / loadModel / loadWeightsFrom, file version 9),
radial/radial_opticalflow_network.lua:120-158 (saveNetwork / loadTesterNetwork, version 1),
radial/generate_calibration_file.lua:106-114 (calibration tables).

What a version-9 file holds (opticalflow_model_io.lua:147-159): version, getModel / getKernels /
getFilter (Lua closures, kept opaque), model_descr, weights = model:getWeights() -- the
convolution *weights* only: biases and the SpatialConvolutionMap connection table are not in
the file (the reference re-draws the table on load) --, geometry, learning, score.
"""
import os

import numpy as np

from . import torch7io
from . import api


def loadCalibration(filename):
    """torch.load(opt.calibration_file) (radial/train_radial_opticalflow.lua:81): wImg, hImg, K
    (3x3), distortion (5), sfm {...}, bad_image_threshold."""
    cal = torch7io.load(filename)
    for k in ("wImg", "hImg", "K", "distortion"):
        if k not in cal:
            raise api.DepthMatchError(api._lib.DM_ERR_INVALID, "%s: no field %r" % (filename, k))
    return cal


def _geometry_of(table):
    g = api.Geometry()
    for k, v in table.items():
        g[k] = [list(x) if isinstance(x, (list, tuple)) else x for x in v] if isinstance(v, list) else v
    return g


def _copy_weights(dst, src, strict_shapes=True):
    n = 0
    for k, v in src.items():
        if k in dst:
            if strict_shapes and tuple(dst[k].shape) != tuple(v.shape):
                raise api.DepthMatchError(api._lib.DM_ERR_INVALID, "weights[%r]: file has %s, model %s"
                                          % (k, tuple(v.shape), tuple(dst[k].shape)))
            dst[k][...] = v
            n += 1
    return n


def modelDirectory(dir, geometry, learning):
    """The directory saveModel builds (opticalflow_model_io.lua:98-146)."""
    if not dir.endswith("/"):
        dir += "/"
    kernel = "_".join("x".join(str(v) for v in l) for l in geometry.layers)
    kernel += "-%sx%s-" % (geometry.maxhHR, geometry.maxwHR)
    if geometry.L2Pooling:
        kernel += "_l2"
    if geometry.output_extraction_method == "mean":
        kernel += "_mean"
    if geometry.share_filters:
        kernel += "_sf"
    if geometry.multiscale:
        kernel += "".join("-%d" % r for r in geometry.ratios)
    train_cascad = ""
    if geometry.multiscale:
        train_cascad = "_tcw" if geometry.cascad_trainable_weights else "_ntcw"
        if geometry.single_beta:
            train_cascad += "_sb"
    lr = learning
    params = "%sx%s-r%s_rd%s_wd%s" % (geometry.maxhGT, geometry.maxwGT, lr["rate"], lr["rate_decay"], lr["weight_decay"])
    if lr.get("soft_targets"):
        params += "_st%s" % lr["st_sigma2"]
    if lr.get("renew_train_set"):
        params += "_renew"
    if lr.get("groundtruth") == "liu":
        params += "_liu"
    params += train_cascad
    images = "%s_%s_%s" % (lr["first_image"], lr["delta"], lr["first_image"] + lr["delta"] * (lr["num_images"] - 1))
    if geometry.motion_correction:
        images += "_mc"
    return dir + kernel + "/" + params + "/" + images


def saveModel(dir, basefilename, geometry, learning, model, nEpochs, score=None):
    """Writes a version-9 table loadWeightsFrom (Lua and here) and loadModel (here) read.  The Lua
    closures of the original (getModel, getKernels, getFilter) cannot be produced without a Lua
    VM and are left out: the reference's own loadModel needs them, its loadWeightsFrom does not."""
    modeldir = modelDirectory(dir, geometry, learning)
    os.makedirs(modeldir, exist_ok=True)
    tosave = {"version": 9, "model_descr": type(model).__name__,
              "weights": {k: np.asarray(v, np.float32) for k, v in model.getWeights().items()},
              "geometry": {k: v for k, v in geometry.items() if v is not None},
              "learning": dict(learning), "score": score}
    path = "%s/%s_e%06d" % (modeldir, basefilename, nEpochs)
    torch7io.save(path, tosave)
    return path


def loadModel(filename, full_output=None, prefilter=None, wImg=None, hImg=None, fused=False, rng=None, ctx=None):
    loaded = torch7io.load(filename)
    if not isinstance(loaded, dict) or loaded.get("version", 0) < 9:
        raise api.DepthMatchError(api._lib.DM_ERR_INVALID,
                                  "loadModel: can't load before version 9 (structure has changed too much)")
    ret = {}
    g = _geometry_of(loaded["geometry"])
    if wImg:
        g.wImg = wImg
    if hImg:
        g.hImg = hImg
    g.training_mode = not full_output
    ret["geometry"] = g
    ret["score"] = loaded.get("score")
    ret["getKernels"] = loaded.get("getKernels")
    if prefilter:
        flt = api.getFilter(g, rng, ctx)
        _copy_weights(flt.getWeights(), loaded["weights"])
        flt.reset_weights()
        ret["filter"] = api.getMultiscalePrefilter(g, flt, ctx) if g.multiscale else flt
        ret["model"] = api.getModel(g, full_output, True, fused=fused, ctx=ctx)
    else:
        if g.multiscale:
            raise api.DepthMatchError(api._lib.DM_ERR_UNSUPPORTED, "loadModel: the multiscale model runs "
                                      "prefiltered here (pass prefilter=True)")
        ret["model"] = api.getModel(g, full_output, False, fused=fused, rng=rng, ctx=ctx)
        _copy_weights(ret["model"].getWeights(), loaded["weights"])
        ret["model"].filter.reset_weights()
    return ret


def loadWeightsFrom(model, filename):
    loaded = torch7io.load(filename)
    if loaded.get("version", 0) < 9:
        raise api.DepthMatchError(api._lib.DM_ERR_INVALID, "Can't load weights from file before version 9")
    _copy_weights(model.getWeights(), loaded["weights"])
    f = getattr(model, "filter", model)
    if hasattr(f, "reset_weights"):
        f.reset_weights()


# ---------------------------------------------------------------- radial (version 1)
def _radial_weights(flt):
    convs = [m for m in flt.modules if hasattr(m, "weight")]
    return [m.weight for m in convs], [m.bias for m in convs]


def saveNetwork(dir, iEpoch, networkp, flt):
    if not dir.endswith("/"):
        dir += "/"
    os.makedirs(dir, exist_ok=True)
    w, b = _radial_weights(flt)
    path = dir + "model_%s" % iEpoch
    torch7io.save(path, {"version": 1, "networkp": dict(networkp), "weights": [list(w), list(b)]})
    return path


def loadTesterNetwork(filename, rng=None, ctx=None):
    """-> (filter, matcher, networkp): getTesterNetwork's parts (the filter on both frames, frame 1
    cropped by hWin-1 rows, then nn.SpatialRadialMatching(hWin))."""
    loaded = torch7io.load(filename)
    if loaded.get("version") != 1:
        raise api.DepthMatchError(api._lib.DM_ERR_INVALID, "Input file has version %s but is required to "
                                  "have version 1" % loaded.get("version"))
    networkp = dict(loaded["networkp"])
    networkp["layers"] = [l if isinstance(l, str) else list(l) for l in networkp["layers"]]
    flt = api.getRadialFilter(networkp, rng, ctx)
    w, b = _radial_weights(flt)
    sw, sb = loaded["weights"]
    if len(sw) != len(w) or len(sb) != len(b):
        raise api.DepthMatchError(api._lib.DM_ERR_INVALID, "loadTesterNetwork: layer count differs")
    for d, s in zip(list(w) + list(b), list(sw) + list(sb)):
        if tuple(d.shape) != tuple(s.shape):
            raise api.DepthMatchError(api._lib.DM_ERR_INVALID, "loadTesterNetwork: weight shape differs")
        d[...] = s
    flt.reset_weights()
    return flt, api.SpatialRadialMatching(networkp["hWin"], ctx=ctx), networkp
