/*
 * matching.c -- oracle restatement of the single-scale matching path.
 * TEST INFRASTRUCTURE ONLY (see dm_oracle.h).  Build with -ffp-contract=off.
 *
 * PARITY UNPINNED for everything in this file except orc_flow_canvas: the
 * arithmetic lives in Torch7 nn/nnx of 2012 (not vendored, no version pin in the
 * reference); it is restated from the reference's call sites and tests.
 */
#include "dm_oracle.h"

#include <math.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int pick_threads(int nthreads) {
#ifdef _OPENMP
  return nthreads > 0 ? nthreads : omp_get_max_threads();
#else
  (void)nthreads;
  return 1;
#endif
}

/* opticalflow_model.lua:93 (nn.SpatialMatching), tests/test_multiscale.lua:149-166:
 * the per-pixel minimum over (dy,dx) of this volume is the brute-force SSD match. */
void orc_spatial_matching(const float *in1, const float *in2, int C, int H1, int W1,
                          int H2, int W2, int maxh, int maxw, float *out, int nthreads) {
  const size_t plane1 = (size_t)H1 * W1, plane2 = (size_t)H2 * W2;
  const int nt = pick_threads(nthreads);
  (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static)
  for (int y = 0; y < H1; ++y) {
    for (int x = 0; x < W1; ++x) {
      float *o = out + ((size_t)y * W1 + x) * maxh * maxw;
      for (int dy = 0; dy < maxh; ++dy) {
        for (int dx = 0; dx < maxw; ++dx) {
          float acc = 0.0f;
          for (int k = 0; k < C; ++k) {
            const float d = in1[k * plane1 + (size_t)y * W1 + x] -
                            in2[k * plane2 + (size_t)(y + dy) * W2 + (x + dx)];
            const float sq = d * d;
            acc = acc + sq;
          }
          o[dy * maxw + dx] = acc;
        }
      }
    }
  }
}

/* radial/radial_opticalflow_network.lua:32-34,56-74 */
void orc_radial_matching(const float *in1, const float *in2, int C, int H1, int W,
                         int H2, int hWin, float *out, int nthreads) {
  orc_spatial_matching(in1, in2, C, H1, W, H2, W, hWin, 1, out, nthreads);
}

/* 2012 Torch7 TH/THGeneral.c THExpMinusApprox (recollection): exp(-x), x >= 0 */
static double th_exp_minus_approx(double x) {
  if (x < 13.0) {
    double y = 1.0 + x * (0.125 + x * (0.0078125 + x * (0.00032552083 + x * 1.0172526e-5)));
    y *= y;
    y *= y;
    y *= y;
    return 1.0 / y;
  }
  return 0.0;
}

/* opticalflow_model.lua:94-109: Minus, reshape to rows x K, SoftMax, reshape back */
void orc_neg_softmax(const float *vol, int64_t rows, int K, int exp_mode, float *out,
                     int nthreads) {
  const int nt = pick_threads(nthreads);
  (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    const float *v = vol + r * K;
    float *o = out + r * K;
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) {
      const float neg = -v[k];
      if (neg >= mx) mx = neg;
    }
    double sum = 0.0;
    for (int k = 0; k < K; ++k) {
      const float neg = -v[k];
      float z;
      if (exp_mode == 1)
        z = (float)th_exp_minus_approx((double)(mx - neg));
      else
        z = expf(neg - mx);
      o[k] = z;
      sum += z;
    }
    const double inv = 1.0 / sum; /* "output_data[d] *= 1/sum" with a double sum */
    for (int k = 0; k < K; ++k) o[k] = (float)((double)o[k] * inv);
  }
}

/* opticalflow_model.lua:153-161 */
void orc_argmax_tie(const float *prob, int64_t rows, int K, int middle, int64_t *idx,
                    float *maxval) {
  for (int64_t r = 0; r < rows; ++r) {
    const float *p = prob + r * K;
    int best = 0;
    float m = p[0];
    for (int k = 1; k < K; ++k)
      if (p[k] > m) {
        m = p[k];
        best = k;
      }
    int64_t out = best + 1;
    if (middle > 0 && m == p[middle - 1]) out = middle;
    idx[r] = out;
    if (maxval) maxval[r] = m;
  }
}

/* radial/radial_opticalflow_groundtruth.lua:88-95, radial/test_radial_opticalflow.lua:205 */
void orc_argmin_tie(const float *vol, int64_t rows, int K, int middle, int64_t *idx,
                    float *minval) {
  for (int64_t r = 0; r < rows; ++r) {
    const float *p = vol + r * K;
    int best = 0;
    float m = p[0];
    for (int k = 1; k < K; ++k)
      if (p[k] < m) {
        m = p[k];
        best = k;
      }
    int64_t out = best + 1;
    if (middle > 0 && m == p[middle - 1]) out = middle;
    idx[r] = out;
    if (minval) minval[r] = m;
  }
}

void orc_top2_relgap(const float *prob, int64_t rows, int K, float *relgap) {
  for (int64_t r = 0; r < rows; ++r) {
    const float *p = prob + r * K;
    float a = -INFINITY, b = -INFINITY; /* a >= b: two largest */
    for (int k = 0; k < K; ++k) {
      if (p[k] > a) {
        b = a;
        a = p[k];
      } else if (p[k] > b) {
        b = p[k];
      }
    }
    const float den = fabsf(a) > 0.0f ? fabsf(a) : 1.0f;
    relgap[r] = (K > 1) ? (a - b) / den : 1.0f;
  }
}

/* OutputExtractor.lua:21-35, opticalflow_model.lua:171-185 */
void orc_soft_mean(const float *prob, int64_t rows, int maxh, int maxw, float *ymean,
                   float *xmean) {
  for (int64_t r = 0; r < rows; ++r) {
    const float *p = prob + r * (int64_t)maxh * maxw;
    double sx = 0.0, sy = 0.0;
    for (int i = 0; i < maxh; ++i)
      for (int j = 0; j < maxw; ++j) {
        const float v = p[i * maxw + j];
        const float px = v * (float)(j + 1);
        const float py = v * (float)(i + 1);
        sx += px;
        sy += py;
      }
    xmean[r] = (float)sx;
    ymean[r] = (float)sy;
  }
}

/* opticalflow_model.lua:191 */
void orc_marginal_x(const float *prob, int64_t rows, int maxh, int maxw, float *pm) {
  for (int64_t r = 0; r < rows; ++r) {
    const float *p = prob + r * (int64_t)maxh * maxw;
    for (int i = 0; i < maxh; ++i) {
      double s = 0.0;
      for (int j = 0; j < maxw; ++j) s += p[i * maxw + j];
      pm[r * maxh + i] = (float)s;
    }
  }
}

/* opticalflow_model.lua:16-25 (x2yx), :208-212 (centre), :227-250 (canvas) */
void orc_flow_canvas(const int64_t *idx, int h1, int w1, int maxh, int maxw, int hImg,
                     int wImg, float *full) {
  const int cy = (maxh + 1) / 2, cx = (maxw + 1) / 2; /* math.ceil(max/2) */
  const int hoff = (hImg - h1) / 2, woff = (wImg - w1) / 2;
  memset(full, 0, sizeof(float) * 2 * (size_t)hImg * wImg);
  for (int y = 0; y < h1; ++y)
    for (int x = 0; x < w1; ++x) {
      const double xd = (double)idx[(size_t)y * w1 + x] - 1.0;
      const double row0 = floor(xd / maxw);
      const double col0 = xd - row0 * maxw;
      const double row = floor(row0 + 1.5), col = floor(col0 + 1.5);
      full[(size_t)(y + hoff) * wImg + (x + woff)] = (float)(row - cy);
      full[(size_t)hImg * wImg + (size_t)(y + hoff) * wImg + (x + woff)] = (float)(col - cx);
    }
}
