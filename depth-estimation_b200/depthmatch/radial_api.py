"""Mirror of the radial pipeline's forward pass, radial/test_radial_opticalflow.lua:183-221:
polar remap of both frames around the epipole, the shared filter (getTesterNetwork: the previous
frame loses its last hWin-1 polar rows first, radial_opticalflow_network.lua:52-74), the 1-D
radial matcher, argmin, back to cartesian, flow -> depth.  The ego-motion estimate (epipole e2,
rotation R) and the image rescale stay with the host (sfm2 / OpenCV)."""
import math

import numpy as np

from . import api


def polarGeometry(networkp, e2, alpha_polar=1.0):
    """rmax and the padded C2P parameters of :183-189."""
    rmax = api.getRMax(networkp["hImg"], networkp["wImg"], e2)
    lp = math.floor((networkp["wKernel"] - 1) / 2)
    rp = math.ceil((networkp["wKernel"] - 1) / 2)
    return dict(wdst=networkp["wInput"], hdst=networkp["hInput"], xcenter=e2[0], ycenter=e2[1],
                lpadding=lp, rpadding=rp, rmax=rmax, alpha=alpha_polar)


class RadialTester:
    """getTesterNetwork(networkp) + the surrounding remaps.  filter: getRadialFilter(networkp) or
    what loadTesterNetwork returns."""

    def __init__(self, networkp, filter, alpha_polar=1.0, use_masks=False, ctx=None):
        self.networkp, self.filter, self.alpha, self.ctx = dict(networkp), filter, alpha_polar, ctx
        self.use_masks = use_masks   # True: build the reference's LUTs and warp through them
        self.matcher = api.SpatialRadialMatching(networkp["hWin"], ctx=ctx)

    def toPolar(self, img, e2):
        pg = polarGeometry(self.networkp, e2, self.alpha)
        if self.use_masks:
            n = self.networkp
            mask = api.getC2PMask(n["wImg"], n["hImg"], n["wInput"], n["hInput"], e2[0], e2[1],
                                  pg["lpadding"], pg["rpadding"], pg["rmax"], self.alpha, ctx=self.ctx)
            return api.cartesian2polar(img, mask, ctx=self.ctx)
        return api.cartesian2polar(img, ctx=self.ctx, **pg)

    def forward(self, prev_warped, img_scaled, e2, kinfty=0.65):
        """-> dict(polar_flow [hOut,wInput], cart_flow, depth, confs) as :196-221 computes them."""
        n = self.networkp
        polar_img = self.toPolar(img_scaled, e2)
        polar_prev = self.toPolar(prev_warped, e2)
        crop = polar_prev[:, :polar_prev.shape[1] - n["hWin"] + 1]         # SpatialPadding(0,0,0,-hWin+1)
        crop = crop.contiguous() if api._is_torch(crop) else np.ascontiguousarray(crop)
        f_prev = self.filter.forward(crop)
        f_img = self.filter.forward(polar_img)
        idx, _ = self.matcher.argmin_flow([f_prev, f_img])                 # output:min(3), idx - 1
        if self.use_masks:
            cart = api.cartesian2polar(idx, api.getP2CMaskOF(n, e2, self.alpha, ctx=self.ctx), ctx=self.ctx)
        else:
            hPolar = n["hInput"] - n["hKernel"] - n["hWin"] + 2
            k = hPolar / n["hInput"]
            cart = api.polar2cartesian(idx, n["wImg"] * k, n["hImg"] * k, e2[0] * k, e2[1] * k,
                                       api.getRMax(n["hImg"], n["wImg"], e2) * k, self.alpha, ctx=self.ctx)
        kout = api.getKOutput(n)
        depth, confs = api.flow2depth(n, cart, (e2[0] * kout, e2[1] * kout), kinfty, ctx=self.ctx)
        return dict(polar_flow=idx, cart_flow=cart, depth=depth, confs=confs)
