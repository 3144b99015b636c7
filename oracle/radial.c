/*
 * radial.c -- oracle restatement of the polar remap used by the radial path.
 * TEST INFRASTRUCTURE ONLY (see dm_oracle.h).  Build with -ffp-contract=off.
 *
 * Sources: radial/cartesian2polar.lua:4-49 (getC2PMask), :51-89 (getP2CMask),
 * :91-93 (cartesian2polar = image.warp bilinear, absolute coordinates),
 * radial/radial_opticalflow_polar.lua:4-10 (getRMax),
 * radial/radial_opticalflow_display.lua:6-58 (flow2depth).
 * The LUT builders are restated from in-tree inline C (float variables, double
 * libm calls, exactly as a C compiler evaluates those expressions).
 * image.warp is out-of-tree (Torch7 `image`, 2012): PARITY UNPINNED.
 */
#include "dm_oracle.h"

#include <math.h>
#include <stddef.h>
#include <string.h>

/* radial/cartesian2polar.lua:4-49 */
void orc_c2p_mask(int wdst, int hdst, double xcenter_d, double ycenter_d, int lpad, int rpad,
                  double rmax, double alpha_d, float *mask) {
  const int wp = wdst + lpad + rpad;
  const size_t plane = (size_t)hdst * wp;
  /* Lua side (:12-13): kr = rmax/(hdst^alpha), ktheta = 2*pi/wdst in double, then
   * every argument lands in a C float (:19-24). */
  const float xcenter = (float)xcenter_d, ycenter = (float)ycenter_d;
  const float kr = (float)(rmax / pow((double)hdst, alpha_d));
  const float ktheta = (float)(2.0 * M_PI / (double)wdst);
  const float alpha = (float)alpha_d;
  for (int i = 0; i < hdst; ++i)
    for (int j = 0; j < wdst; ++j) {
      const float r = (float)((double)kr * pow((double)(float)i, (double)alpha));
      const float theta = ktheta * (float)j;
      mask[(size_t)i * wp + lpad + j] = (float)((double)r * sin((double)theta) + (double)ycenter);
      mask[plane + (size_t)i * wp + lpad + j] =
          (float)((double)r * cos((double)theta) + (double)xcenter);
    }
  for (int c = 0; c < 2; ++c)
    for (int i = 0; i < hdst; ++i) {
      float *row = mask + c * plane + (size_t)i * wp;
      for (int j = 0; j < lpad; ++j) row[j] = row[lpad + wdst - lpad + j];      /* :41-43 */
      for (int j = 0; j < rpad; ++j) row[lpad + wdst + j] = row[lpad + j];      /* :44-46 */
    }
}

/* radial/cartesian2polar.lua:51-89 */
void orc_p2c_mask(int wsrc, int hsrc, int wdst, int hdst, double xcenter_d, double ycenter_d,
                  double rmax, double alpha_d, float *mask) {
  const size_t plane = (size_t)hdst * wdst;
  const double pi2d = 2.0 * M_PI;
  const float xcenter = (float)xcenter_d, ycenter = (float)ycenter_d;
  const float kx = (float)((double)wsrc / pi2d);
  const float ky = (float)((double)hsrc / pow(rmax, 1.0 / alpha_d));
  const float pi2 = (float)pi2d;
  const float invalpha = (float)(1.0 / alpha_d) * 0.5f;
  for (int i = 0; i < hdst; ++i)
    for (int j = 0; j < wdst; ++j) {
      const float x = (float)j - xcenter;
      const float y = (float)i - ycenter;
      const float x2 = x * x, y2 = y * y;
      const float n2 = x2 + y2;
      mask[(size_t)i * wdst + j] = (float)(pow((double)n2, (double)invalpha) * (double)ky);
      mask[plane + (size_t)i * wdst + j] =
          (float)(fmod(atan2((double)y, (double)x) + (double)pi2, (double)pi2) * (double)kx);
    }
}

/* radial/radial_opticalflow_polar.lua:4-10 */
double orc_get_rmax(int h, int w, double ex, double ey) {
  const double a = ex * ex + ey * ey, b = (w - ex) * (w - ex) + ey * ey;
  const double c = ex * ex + (h - ey) * (h - ey), d = (w - ex) * (w - ex) + (h - ey) * (h - ey);
  return floor(sqrt(fmax(fmax(a, b), fmax(c, d))));
}

/* image.warp(src, field, 'bilinear', false): Torch7 image/generic/image.c
 * Main_warp as of 2012 (recollection).  PARITY UNPINNED. */
void orc_warp_bilinear(const float *src, int C, int hs, int ws, const float *field, int hd,
                       int wd, float *dst) {
  const size_t fplane = (size_t)hd * wd;
  for (int y = 0; y < hd; ++y)
    for (int x = 0; x < wd; ++x) {
      float iy = field[(size_t)y * wd + x];
      float ix = field[fplane + (size_t)y * wd + x];
      ix = ix > 0 ? ix : 0;
      ix = ix < (float)(ws - 1) ? ix : (float)(ws - 1);
      iy = iy > 0 ? iy : 0;
      iy = iy < (float)(hs - 1) ? iy : (float)(hs - 1);
      const long x0 = (long)floorf(ix), y0 = (long)floorf(iy);
      const long x1 = x0 + 1, y1 = y0 + 1;
      const float wnw = ((float)x1 - ix) * ((float)y1 - iy);
      const float wne = (ix - (float)x0) * ((float)y1 - iy);
      const float wsw = ((float)x1 - ix) * (iy - (float)y0);
      const float wse = (ix - (float)x0) * (iy - (float)y0);
      const long x1c = x1 < ws - 1 ? x1 : ws - 1, y1c = y1 < hs - 1 ? y1 : hs - 1;
      for (int k = 0; k < C; ++k) {
        const float *s = src + (size_t)k * hs * ws;
        const float a = s[y0 * ws + x0] * wnw;
        const float b = s[y0 * ws + x1c] * wne;
        const float c = s[y1c * ws + x0] * wsw;
        const float d = s[y1c * ws + x1c] * wse;
        dst[((size_t)k * hd + y) * wd + x] = ((a + b) + c) + d;
      }
    }
}

/* radial/radial_opticalflow_display.lua:6-58 (the inline C body; the caller then
 * divides by infty, :56) */
void orc_flow2depth(const float *flow, int h, int w, float xcenter, float ycenter,
                    float infty, float *depth, float *confs) {
  for (int i = 0; i < h; ++i)
    for (int j = 0; j < w; ++j) {
      const size_t p = (size_t)i * w + j;
      depth[p] = 0.0f;
      confs[p] = 1.0f;
      const float dx = (float)j - xcenter, dy = (float)i - ycenter;
      const float d = (float)sqrt((double)(dx * dx + dy * dy));
      if (d > 10.0f) {
        const float f = flow[p];
        depth[p] = f < 0.1f ? infty : d / f;
      } else {
        confs[p] = 0.0f;
      }
    }
}

/* sfm2.removeEgoMotion (out-of-tree sfm2; call sites depth_estimation_api.lua:147,
 * radial/test_radial_opticalflow.lua:192), restated as a homography gather: destination (x, y)
 * samples the source at hmat * (x, y, 1), bilinear like image.warp; outside the frame 0 and
 * mask = 0.  PARITY UNPINNED (the library is not in the tree). */
void orc_warp_homography(const float *src, int C, int hs, int ws, const double *hmat, int hd, int wd,
                         float *dst, float *mask) {
  for (int y = 0; y < hd; ++y)
    for (int x = 0; x < wd; ++x) {
      const double X = hmat[0] * x + hmat[1] * y + hmat[2];
      const double Y = hmat[3] * x + hmat[4] * y + hmat[5];
      const double Z = hmat[6] * x + hmat[7] * y + hmat[8];
      const float ix = (float)(X / Z), iy = (float)(Y / Z);
      const int inside = Z > 0.0 && ix >= 0.0f && ix <= (float)(ws - 1) && iy >= 0.0f && iy <= (float)(hs - 1);
      if (mask) mask[(size_t)y * wd + x] = inside ? 1.0f : 0.0f;
      if (!inside) {
        for (int k = 0; k < C; ++k) dst[((size_t)k * hd + y) * wd + x] = 0.0f;
        continue;
      }
      const long x0 = (long)floorf(ix), y0 = (long)floorf(iy);
      const long x1 = x0 + 1, y1 = y0 + 1;
      const float wnw = ((float)x1 - ix) * ((float)y1 - iy);
      const float wne = (ix - (float)x0) * ((float)y1 - iy);
      const float wsw = ((float)x1 - ix) * (iy - (float)y0);
      const float wse = (ix - (float)x0) * (iy - (float)y0);
      const long x1c = x1 < ws - 1 ? x1 : ws - 1, y1c = y1 < hs - 1 ? y1 : hs - 1;
      for (int k = 0; k < C; ++k) {
        const float *s = src + (size_t)k * hs * ws;
        const float a = s[y0 * ws + x0] * wnw;
        const float b = s[y0 * ws + x1c] * wne;
        const float c = s[y1c * ws + x0] * wsw;
        const float d = s[y1c * ws + x1c] * wse;
        dst[((size_t)k * hd + y) * wd + x] = ((a + b) + c) + d;
      }
    }
}
