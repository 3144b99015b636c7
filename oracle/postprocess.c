/*
 * postprocess.c -- oracle restatement of the steps right after the matching path
 * (SURVEY.md 8f "next" rows 1 and 2).  TEST INFRASTRUCTURE ONLY (see dm_oracle.h).
 *
 * Sources: postProcessImage inline C (opticalflow_model.lua:323-472), enlargeMask
 * (depth_estimation_api.lua:76-132), radial() (test_opticalflow.lua:143-216),
 * computeDepthMapFromFlow (ardrone/ardrone_api.cpp:99-140).
 * The first three are PINNED against the reference's own inline C, extracted verbatim and
 * compiled into oracle/_ref (oracle/extract_inline.py); computeDepthMapFromFlow needs OpenCV and
 * the drone class to compile, it is restated only (parity unpinned).
 */
#include "dm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static int cmp_float(const void *a_, const void *b_) {
  const float a = *(const float *)a_, b = *(const float *)b_;
  return a == b ? 0 : (a < b ? -1 : 1);
}

/* opticalflow_model.lua:388-434 (fmed): k x k window anchored at (i,j), i < h-k, j < w-k, masked
 * entries only, median = sorted[n/2] (0 when the window holds no masked pixel).  ret must be
 * zero-filled by the caller like the Lua does (:324). */
void orc_pp_median(const float *flow, const float *mask, int h, int w, int k, float *ret) {
  const size_t plane = (size_t)h * w;
  const int halfk = k / 2;
  float ty[32], tx[32];
  for (int i = 0; i < h - k; ++i)
    for (int j = 0; j < w - k; ++j) {
      memset(ty, 0, sizeof(ty));
      memset(tx, 0, sizeof(tx));
      int n = 0;
      for (int a = i; a < i + k; ++a)
        for (int b = j; b < j + k; ++b)
          if (mask[(size_t)a * w + b] != 0.0f && n < 32) {
            ty[n] = flow[(size_t)a * w + b];
            tx[n] = flow[plane + (size_t)a * w + b];
            ++n;
          }
      qsort(ty, n, sizeof(float), cmp_float);
      qsort(tx, n, sizeof(float), cmp_float);
      ret[(size_t)(i + halfk) * w + (j + halfk)] = ty[n / 2];
      ret[plane + (size_t)(i + halfk) * w + (j + halfk)] = tx[n / 2];
    }
}

/* opticalflow_model.lua:342-386 (fmax): mode of v = vx + 16*vy over the masked window (values
 * truncated to int; the reference indexes a 256-bin histogram without a range check, entries
 * outside 0..255 are skipped here), first maximum wins. */
void orc_pp_mode(const float *flow, const float *mask, int h, int w, int k, float *ret) {
  const size_t plane = (size_t)h * w;
  const int halfk = k / 2;
  int hist[256];
  for (int i = 0; i < h - k; ++i)
    for (int j = 0; j < w - k; ++j) {
      memset(hist, 0, sizeof(hist));
      for (int a = i; a < i + k; ++a)
        for (int b = j; b < j + k; ++b)
          if (mask[(size_t)a * w + b] != 0.0f) {
            const int vx = (int)flow[plane + (size_t)a * w + b];
            const int vy = (int)flow[(size_t)a * w + b];
            const int v = vx + 16 * vy;
            if (v >= 0 && v < 256) ++hist[v];
          }
      int im = 0;
      for (int l = 0; l < 256; ++l)
        if (hist[l] > hist[im]) im = l;
      ret[plane + (size_t)(i + halfk) * w + (j + halfk)] = (float)(im % 16);
      ret[(size_t)(i + halfk) * w + (j + halfk)] = (float)(im / 16);
    }
}

/* postProcessImage (opticalflow_model.lua:323-327,435-443): 'max' rounds, shifts by the global
 * minimum so that the histogram index is non-negative, and adds it back EVERYWHERE (the border
 * that the filter never writes ends up at m, like the reference's `output+m`). */
void orc_post_process_image(const float *input, const float *mask, int h, int w, int k, int method_max,
                            float *output) {
  const size_t n2 = (size_t)2 * h * w;
  memset(output, 0, n2 * sizeof(float));
  if (!method_max) {
    orc_pp_median(input, mask, h, w, k, output);
    return;
  }
  float *r = (float *)malloc(n2 * sizeof(float));
  float m = INFINITY;
  for (size_t i = 0; i < n2; ++i) {
    r[i] = floorf(input[i] + 0.5f);
    if (r[i] < m) m = r[i];
  }
  for (size_t i = 0; i < n2; ++i) r[i] = r[i] - m;
  orc_pp_mode(r, mask, h, w, k, output);
  for (size_t i = 0; i < n2; ++i) output[i] = output[i] + m;
  free(r);
}

/* depth_estimation_api.lua:76-132, in place */
void orc_enlarge_mask(float *mask, int h, int w, int ix, int iy) {
  for (int i = 0; i < h; ++i) {
    float *row = mask + (size_t)i * w;
    for (int j = 0; j < w; ++j)
      if (row[j] > 0.5f) {
        for (int k = j; k < (j + ix < w ? j + ix : w); ++k) row[k] = 0.0f;
        break;
      }
    for (int j = w - 1; j >= 0; --j)
      if (row[j] > 0.5f) {
        for (int k = j; k >= (j - ix + 1 > 0 ? j - ix + 1 : 0); --k) row[k] = 0.0f;
        break;
      }
  }
  for (int j = 0; j < w; ++j) {
    for (int i = 0; i < h; ++i)
      if (mask[(size_t)i * w + j] > 0.5f) {
        for (int k = i; k < (i + iy < h ? i + iy : h); ++k) mask[(size_t)k * w + j] = 0.0f;
        break;
      }
    for (int i = h - 1; i >= 0; --i)
      if (mask[(size_t)i * w + j] > 0.5f) {
        for (int k = i; k >= (i - iy + 1 > 0 ? i - iy + 1 : 0); --k) mask[(size_t)k * w + j] = 0.0f;
        break;
      }
  }
}

/* test_opticalflow.lua:143-193 (radial): depth = |p - c| / |flow|; note the reference's
 * `px*dx+dy*dy` (sic).  ret/conf start at zero (:146-147). */
void orc_radial_depth(const float *flow, int h, int w, float mh, float mw, float infty, float *ret,
                      float *conf) {
  const size_t plane = (size_t)h * w;
  memset(ret, 0, plane * sizeof(float));
  memset(conf, 0, plane * sizeof(float));
  for (int i = 0; i < h; ++i)
    for (int j = 0; j < w; ++j) {
      const size_t o = (size_t)i * w + j;
      const float py = (float)i - mh, px = (float)j - mw;
      const float pn = (float)sqrt((double)(px * px + py * py));
      const float dy = flow[o], dx = flow[plane + o];
      const float dn = (float)sqrt((double)(dx * dx + dy * dy));
      if (dn >= 0.2f) {
        const float q = pn / dn;
        ret[o] = q < infty ? q : infty;
        if (px * dx + dy * dy > 0.125f) conf[o] = 1.0f;
      } else {
        conf[o] = 1.0f;
        ret[o] = infty;
      }
    }
}

/* ardrone/ardrone_api.cpp:99-140: mode filter of the rounded x-flow over the masked
 * [i-3,i+3) x [j-3,j+3) window (20 bins, flow+8; out-of-range flows are skipped here, the
 * reference would index out of bounds), then depth = m*|j - w/2| / |flow|, 100 when |flow| < 1.1.
 * Unmasked pixels: depth 0 (uninitialised in the reference), confidence 0. */
void orc_depth_from_xflow(const float *xflow, const float *mask, int h, int w, float m, float *depth,
                          float *conf) {
  const int k = 3;
  float *fp = (float *)calloc((size_t)h * w, sizeof(float));
  for (int i = 0; i < w; ++i)
    for (int j = 0; j < h; ++j)
      if (mask[(size_t)j * w + i] != 0.0f) {
        int values[20] = {0};
        for (int i2 = (i - k > 0 ? i - k : 0); i2 < (i + k < w ? i + k : w); ++i2)
          for (int j2 = (j - k > 0 ? j - k : 0); j2 < (j + k < h ? j + k : h); ++j2)
            if (mask[(size_t)j2 * w + i2] != 0.0f) {
              const int f = (int)round((double)xflow[(size_t)j2 * w + i2]);
              if (f + 8 >= 0 && f + 8 < 20) ++values[f + 8];
            }
        int mx = 0, im = 0;
        for (int iv = 0; iv < 20; ++iv)
          if (values[iv] > mx) {
            mx = values[iv];
            im = iv - 8;
          }
        fp[(size_t)j * w + i] = (float)im;
      }
  const int middlex = w / 2;
  for (int i = 0; i < h; ++i)
    for (int j = 0; j < w; ++j) {
      const size_t o = (size_t)i * w + j;
      depth[o] = 0.0f;
      if (mask[o] > 0.5f && j - middlex != 0) {
        const float a = fabsf(fp[o]);
        depth[o] = a < 1.1f ? 100.0f : m * (float)abs(j - middlex) / a;
        conf[o] = 1.0f;
      } else {
        conf[o] = 0.0f;
      }
    }
  free(fp);
}
