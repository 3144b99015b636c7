"""Mirror of depth_estimation_api.lua -- the surface the robot host calls
(ardrone/ardrone_api.cpp:32-88 embeds one Lua VM and calls nextFrameDepth()).

The camera and OpenCV parts stay with the host (frame grab, sfm2.undistortImage,
sfm2.getEgoMotion, image.scale: depth_estimation_api.lua:63-72,137-144); everything from the
scaled frame on runs on the GPU: removeEgoMotion of the previous frame's *features*
(:147), filter:forward (:149), prepareInput -> model:forward -> processOutput (:164-168),
enlargeMask and the mask composition (:171-186).  The previous frame's features stay on the
device between calls, as `last_filtered` does in the reference (:68-72,188-190), so a stream
uploads every frame once.
"""
import math

import numpy as np

from . import api

try:
    import torch
except ImportError:  # pragma: no cover
    torch = None


def removeEgoMotion(im, K, R, mode="bilinear", inverse=False, ctx=None):
    """sfm2.removeEgoMotion(im, K, R[, mode]) (out-of-tree sfm2): the image (or feature maps)
    [C,H,W] seen through the camera rotated by R: a gather through the homography K R K^-1
    (K^-1... R^-1 when inverse=True).  Returns (warped, mask); mask is 0 where the source pixel
    falls outside the frame."""
    if mode != "bilinear":
        raise api.DepthMatchError(api._lib.DM_ERR_UNSUPPORTED, "removeEgoMotion: only 'bilinear'")
    K = np.asarray(K.cpu() if api._is_torch(K) else K, np.float64).reshape(3, 3)
    R = np.asarray(R.cpu() if api._is_torch(R) else R, np.float64).reshape(3, 3)
    Hm = K @ (np.linalg.inv(R) if inverse else R) @ np.linalg.inv(K)
    return warpHomography(im, Hm, ctx=ctx)


def warpHomography(im, Hm, hd=None, wd=None, ctx=None):
    args = api._Args(ctx)
    sptr, s = args.inp(im)
    c = args.ctx_for(s)
    ch, hs, ws = s.shape
    hd, wd = hd or hs, wd or ws
    hm = np.ascontiguousarray(Hm, np.float64).reshape(9)
    dptr, dst = args.out((ch, hd, wd), np.float32, like=s)
    mptr, mask = args.out((hd, wd), np.float32, like=s)
    api.check(c._lib.dm_warp_homography(c.handle, sptr, ch, hs, ws,
                                        hm.ctypes.data_as(api.C.POINTER(api.C.c_double)), hd, wd, dptr, mptr))
    return dst, mask


class DepthEstimationAPI:
    """The state depth_estimation_api.lua keeps in file-level locals.

    geometry: wImg, hImg, maxh, maxw, layers (as loadModel returns it); filter: a Filter
    (getFilter / loadModel(...)['filter']); K: the 3x3 camera matrix of the *full-size* frame --
    the feature warp uses Khalf = K/2 with K[3][3] = 1 like the reference (:55-56)."""

    bad_image_threshold = 0.2  # nInliers / nFound (:158)

    def __init__(self, geometry, filter, K=None, first_frame=None, fused=True, ctx=None):
        self.geometry = api.Geometry(geometry)
        self.geometry.prefilter = True
        self.geometry.output_extraction_method = "mean"       # :31
        self.filter, self.ctx = filter, ctx
        # fused: one kernel gives the soft mean and the marginal confidence (no volume in HBM);
        # fused=False walks the reference's graph: probability volume, then processOutput on it
        self.model = api.DenseMatch(self.geometry, ctx=ctx) if fused else \
            api._MatchModel(api.Geometry(self.geometry, output_extraction_method="max"), ctx=ctx)
        self.Khalf = None
        if K is not None:
            self.Khalf = np.asarray(K, np.float64).reshape(3, 3) * 0.5
            self.Khalf[2, 2] = 1.0
        self.last_filtered = None
        self.last_im_scaled = None
        if first_frame is not None:
            self.reset(first_frame)

    def reset(self, im_scaled):
        """:63-72 -- the first frame only primes last_filtered."""
        self.last_im_scaled = im_scaled
        self.last_filtered = self.filter.forward(im_scaled)

    def nextFrameDepth(self, im_scaled, R=None, nFound=1, nInliers=1):
        """:134-200.  im_scaled: the undistorted frame scaled to geometry.wImg x hImg, [3,h,w]
        (numpy, or a torch CUDA tensor to keep everything on the device); R, nFound, nInliers:
        what the host's sfm2.getEgoMotion returned (R None: no rotation compensation).
        Returns (im_scaled, x-flow [hImg,wImg], mask [hImg,wImg]) like the reference."""
        g = self.geometry
        if self.last_filtered is None:
            raise api.DepthMatchError(api._lib.DM_ERR_INVALID, "nextFrameDepth: call reset(first_frame) first")
        on_dev = api._is_torch(im_scaled)
        last = self.last_filtered
        if R is not None and self.Khalf is not None:
            last, mask = removeEgoMotion(last, self.Khalf, R, ctx=self.ctx)              # :147
        else:
            h, w = last.shape[-2:]
            mask = torch.ones((h, w), device=last.device) if api._is_torch(last) else np.ones((h, w), np.float32)
        filtered = self.filter.forward(im_scaled)                                      # :149
        if nInliers / max(nFound, 1) < self.bad_image_threshold:                       # :158-161
            zeros = (lambda *s: torch.zeros(s, device=im_scaled.device)) if on_dev else \
                (lambda *s: np.zeros(s, np.float32))
            mask, output = zeros(g.hImg, g.wImg), zeros(2, g.hImg, g.wImg)
        else:
            inp = api.prepareInput(g, last, filtered)                                  # :163
            moutput = self.model.forward(inp)                                          # :165
            poutput = api.processOutput(g, moutput, True, None, ctx=self.ctx)          # :167
            output = poutput["full"]
            hy, wy = poutput["y"].shape[-2:]
            mask = api.enlargeMask(mask if on_dev else np.ascontiguousarray(mask),     # :171-173
                                   math.ceil((g.wImg - wy) / 2), math.ceil((g.hImg - hy) / 2), ctx=self.ctx)
            mh, mw = mask.shape
            mask2 = torch.zeros((g.hImg, g.wImg), device=mask.device) if api._is_torch(mask) else \
                np.zeros((g.hImg, g.wImg), np.float32)
            # narrow(1, floor((hImg-mh)/2), mh): 1-based start in Lua; 0 would be an error there,
            # so the feature map is always smaller than the frame
            oy, ox = max((g.hImg - mh) // 2 - 1, 0), max((g.wImg - mw) // 2 - 1, 0)
            mask2[oy:oy + mh, ox:ox + mw] = mask                                        # :175-179
            conf = poutput["full_confidences"]
            mask = mask2 * (conf.to(mask2.dtype) if api._is_torch(conf) else conf.astype(np.float32))  # :181
        self.last_im_scaled = im_scaled                                                # :187-189
        self.last_filtered = filtered
        return im_scaled, output[1], mask
