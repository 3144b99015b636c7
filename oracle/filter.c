/*
 * filter.c -- CPU oracle (TEST INFRASTRUCTURE ONLY) for the feature extractor in front of the
 * matching path: getFilter (opticalflow_model.lua:45-79, radial/radial_opticalflow_network.lua:6-31)
 * and getMultiscalePrefilter (opticalflow_model_multiscale.lua:134-173).
 *
 * nn.SpatialConvolution / nn.SpatialConvolutionMap / nn.Tanh / nn.SpatialZeroPadding are Torch7
 * nn (out-of-tree, un-pinned): restated from the published algorithm (valid cross-correlation,
 * TH conv2Dmv 'V','X': bias fill, then for every connected input plane the per-pixel tap sum in
 * (ky,kx) order is added).  PARITY UNPINNED numerically; cross-checked against torch's conv2d
 * and the reference's identity "patch-unfolding" weights (tests/test_multiscale.lua:44-55) in
 * tests/test_oracle_invariants.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "dm_oracle.h"
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

void orc_conv_layer(const float *in, int n_in, int h, int w, const float *weight, const float *bias,
                    int n_out, int kh, int kw, const int32_t *conn, int n_conn, int pad_l, int pad_r,
                    int pad_t, int pad_b, int tanh_after, float *out, int nthreads) {
  const int hp = h + pad_t + pad_b, wp = w + pad_l + pad_r;
  const int ho = hp - kh + 1, wo = wp - kw + 1;
  if (ho <= 0 || wo <= 0) return;
  /* nn.SpatialZeroPadding (multiscale.lua:146) */
  float *p = (float *)calloc((size_t)n_in * hp * wp, sizeof(float));
  for (int c = 0; c < n_in; ++c)
    for (int y = 0; y < h; ++y)
      memcpy(p + ((size_t)c * hp + y + pad_t) * wp + pad_l, in + ((size_t)c * h + y) * w, sizeof(float) * w);
  const int full = conn == NULL;
  const int nc = full ? n_in * n_out : n_conn;
  const int nt = pick_threads(nthreads);
  (void)nt;
#pragma omp parallel for num_threads(nt) schedule(dynamic)
  for (int o = 0; o < n_out; ++o) {
    float *dst = out + (size_t)o * ho * wo;
    for (int i = 0; i < ho * wo; ++i) dst[i] = bias[o];
    for (int e = 0; e < nc; ++e) {
      /* full: weight[o][i][kh][kw]; map: conn rows are 1-based (from, to), weight[e][kh][kw] */
      const int from = full ? e % n_in : conn[2 * e] - 1;
      const int to = full ? e / n_in : conn[2 * e + 1] - 1;
      if (to != o) continue;
      const float *wk = weight + (size_t)e * kh * kw;
      const float *src = p + (size_t)from * hp * wp;
      for (int y = 0; y < ho; ++y)
        for (int x = 0; x < wo; ++x) {
          float sum = 0.0f;
          for (int ky = 0; ky < kh; ++ky)
            for (int kx = 0; kx < kw; ++kx) sum += src[(size_t)(y + ky) * wp + x + kx] * wk[ky * kw + kx];
          dst[(size_t)y * wo + x] += sum;
        }
    }
    if (tanh_after)
      for (int i = 0; i < ho * wo; ++i) dst[i] = tanhf(dst[i]);
  }
  free(p);
}
