/*
 * multiscale.c -- oracle restatement of the multiscale index encoding, the
 * cascade and the ring extraction.  TEST INFRASTRUCTURE ONLY (see dm_oracle.h).
 *
 * Sources: opticalflow_model_multiscale.lua:10-52 (yx2xMulti), :83-132
 * (x2yxMultiNumber), :293-333 (ring join), x2yxMulti2.c:1-95 (vectorised decode,
 * with its divergences), CascadingAddTable.lua:108-135, common.lua (round).
 * The bug-compatible decode is PINNED against the reference C compiled into
 * oracle/_ref/; the Lua-spec functions are pinned by the reference's own
 * round-trip test (tests/test_multiscale.lua:57-80) restated in tests/.
 */
#include "dm_oracle.h"

#include <math.h>
#include <stddef.h>
#include <string.h>

static double lua_round(double v) { return floor(v + 0.5); } /* common.lua round() */

static int ring_border(int maxw, int r, int rprev) {
  return (int)lua_round((double)maxw * (double)(r - rprev) / (2.0 * (double)r));
}

static int ring_len(int maxh, int maxw, int d) { return 2 * d * maxw + 2 * (maxh - 2 * d) * d; }

static int is_in(double size, double v) { return v >= -ceil(size / 2) + 1 && v <= floor(size / 2); }

int orc_multiscale_length(int maxh, int maxw, const int *ratios, int nratios) {
  int L = maxh * maxw;
  for (int i = 1; i < nratios; ++i) L += ring_len(maxh, maxw, ring_border(maxw, ratios[i], ratios[i - 1]));
  return L;
}

/* opticalflow_model_multiscale.lua:10-52 */
int64_t orc_yx2x_multi(int maxh, int maxw, const int *ratios, int nratios, double y, double x) {
  x = lua_round(x);
  y = lua_round(y);
  int i = 0;
  double tx = 0, ty = 0;
  for (; i < nratios; ++i) {
    if (is_in((double)maxw * ratios[i], x) && is_in((double)maxh * ratios[i], y)) {
      tx = ceil(x / ratios[i]) + ceil(maxw / 2.0);
      ty = ceil(y / ratios[i]) + ceil(maxh / 2.0);
      break;
    }
  }
  if (i == nratios) return 0; /* assert(i <= #ratios) */
  const int64_t X = (int64_t)tx, Y = (int64_t)ty;
  if (i == 0) return (Y - 1) * maxw + X;
  const int d = ring_border(maxw, ratios[i], ratios[i - 1]);
  int64_t it;
  if (Y <= d)
    it = (Y - 1) * maxw + X;
  else if (Y > maxh - d)
    it = (int64_t)d * maxw + 2 * (maxh - 2 * d) * d + (Y - (maxh - d) - 1) * maxw + X;
  else if (X <= d)
    it = (int64_t)d * maxw + (Y - d - 1) * d + X;
  else if (X > maxw - d)
    it = (int64_t)d * maxw + (maxh - 2 * d) * d + (Y - d - 1) * d + X - (maxw - d);
  else
    return 0; /* assert(false) */
  return (int64_t)maxw * maxh + (int64_t)(i - 1) * ring_len(maxh, maxw, d) + it;
}

/* opticalflow_model_multiscale.lua:83-132 */
int orc_x2yx_multi_number(int maxh, int maxw, const int *ratios, int nratios, int64_t x,
                          int64_t *outy, int64_t *outx) {
  const int64_t cy = (maxh + 1) / 2, cx = (maxw + 1) / 2; /* math.ceil(max/2) */
  if (x <= (int64_t)maxh * maxw) {
    *outy = (x - 1) / maxw + 1 - cy; /* x >= 1 on this branch */
    *outx = (x - 1) % maxw + 1 - cx;
    return x >= 1 ? 0 : -1;
  }
  x -= (int64_t)maxh * maxw;
  for (int i = 1; i < nratios; ++i) {
    const int d = ring_border(maxw, ratios[i], ratios[i - 1]);
    const int len = ring_len(maxh, maxw, d);
    const int64_t side = (int64_t)(maxh - 2 * d) * d;
    if (x > len) {
      x -= len;
      continue;
    }
    int64_t ty, tx;
    if (x <= (int64_t)d * maxw) { /* top rows */
      ty = (x - 1) / maxw + 1;
      tx = (x - 1) % maxw + 1;
    } else if ((x -= (int64_t)d * maxw) <= side) { /* left columns */
      ty = (x - 1) / d + 1 + d;
      tx = (x - 1) % d + 1;
    } else if ((x -= side) <= side) { /* right columns */
      ty = (x - 1) / d + 1 + d;
      tx = (x - 1) % d + 1 + maxw - d;
    } else if ((x -= side) <= (int64_t)d * maxw) { /* bottom rows */
      ty = (x - 1) / maxw + 1 + maxh - d;
      tx = (x - 1) % maxw + 1;
    } else {
      return -1;
    }
    *outy = (ty - cy) * ratios[i];
    *outx = (tx - cx) * ratios[i];
    return 0;
  }
  return -1;
}

/* x2yxMulti2.c:1-95, bug for bug.  The Lua table is read with keys 0..n-1
 * (:15-19) so the C sees {0, ratios[0], ..., ratios[n-2]}; ceil() is applied to
 * an integer quotient (:24-25); lengths lack the factor d on the first term
 * (:41); the scale-1 test and the row tests are strict (:50,:61,:79); when no
 * segment matches the entry is left untouched and the mutated x carries over to
 * the next ratio (:57-85). */
void orc_x2yx_multi2_bugcompat(const int64_t *xim, int h, int w, int maxh, int maxw,
                               const int *ratios, int nratios, int64_t *retx,
                               int64_t *rety) {
  int rc[16], border[16], length[16];
  if (nratios > 10) nratios = 10; /* N_MAX_RATIOS */
  for (int i = 0; i < nratios; ++i) rc[i] = i == 0 ? 0 : ratios[i - 1];
  const int chh = maxh / 2, chw = maxw / 2;
  const int area = maxh * maxw;
  for (int i = 1; i < nratios; ++i) {
    border[i] = (int)round((float)maxw * ((float)rc[i] - (float)rc[i - 1]) / (2.0f * (float)rc[i]));
    length[i] = 2 * maxw + 2 * (maxh - 2 * border[i]) * border[i];
  }
  for (int64_t p = 0; p < (int64_t)h * w; ++p) {
    long x = (long)xim[p];
    if (x < area) {
      rety[p] = (x - 1) / maxw + 1 - chh; /* C division truncates toward zero */
      retx[p] = (x - 1) % maxw + 1 - chw; /* C remainder keeps the dividend's sign */
      continue;
    }
    x -= area;
    for (int k = 1; k < nratios; ++k) {
      const int d = border[k];
      const int mH = (maxh - 2 * d) * d;
      if (x > length[k]) {
        x -= length[k];
        continue;
      }
      if (x < (long)d * maxw) {
        rety[p] = ((x - 1) / maxw + 1 - chh) * rc[k];
        retx[p] = ((x - 1) % maxw + 1 - chw) * rc[k];
        break;
      }
      x -= (long)d * maxw;
      if (d == 0) continue; /* the reference divides by zero here (SIGFPE); never reached with valid ratios */
      if (x <= mH) {
        rety[p] = ((x - 1) / d + 1 + d - chh) * rc[k];
        retx[p] = ((x - 1) % d + 1 - chw) * rc[k];
        break;
      }
      x -= mH;
      if (x <= mH) {
        rety[p] = ((x - 1) / d + 1 + d - chh) * rc[k];
        retx[p] = ((x - 1) % d + 1 + maxw - d - chw) * rc[k];
        break;
      }
      x -= mH;
      if (x < (long)d * maxw) {
        rety[p] = ((x - 1) / maxw + 1 + maxh - d - chh) * rc[k];
        retx[p] = ((x - 1) % maxw + 1 - chw) * rc[k];
        break;
      }
    }
  }
}

/* CascadingAddTable.lua:108-135 (forward; normalisers commented out :29,:46,:61) */
void orc_cascade_add(const float *in, int64_t rows, int Kh, int Kw, const int *ratios,
                     int nratios, float *out) {
  const size_t per = (size_t)rows * Kh * Kw;
  memcpy(out + (nratios - 1) * per, in + (nratios - 1) * per, per * sizeof(float));
  for (int i = nratios - 2; i >= 0; --i) {
    const int r = ratios[i], r2 = ratios[i + 1];
    const int dh = Kh * (r2 - r) / (2 * r2), dw = Kw * (r2 - r) / (2 * r2);
    const int f = r2 / r;
    const float *coarse = out + (size_t)(i + 1) * per;
    const float *fine = in + (size_t)i * per;
    float *o = out + (size_t)i * per;
    for (int64_t p = 0; p < rows; ++p)
      for (int a = 0; a < Kh; ++a)
        for (int b = 0; b < Kw; ++b) {
          const size_t dst = ((size_t)p * Kh + a) * Kw + b;
          const size_t src = ((size_t)p * Kh + (dh + a / f)) * Kw + (dw + b / f);
          o[dst] = fine[dst] + coarse[src];
        }
  }
}

/* opticalflow_model_multiscale.lua:293-333 */
void orc_ring_join(const float *casc, int64_t rows, int maxh, int maxw, const int *ratios,
                   int nratios, float *outvec) {
  const int L = orc_multiscale_length(maxh, maxw, ratios, nratios);
  const size_t per = (size_t)rows * maxh * maxw;
  for (int64_t p = 0; p < rows; ++p) {
    float *o = outvec + (size_t)p * L;
    const float *s0 = casc + (size_t)p * maxh * maxw;
    int n = 0;
    for (int k = 0; k < maxh * maxw; ++k) o[n++] = s0[k];
    for (int i = 1; i < nratios; ++i) {
      const int d = ring_border(maxw, ratios[i], ratios[i - 1]);
      const float *s = casc + (size_t)i * per + (size_t)p * maxh * maxw;
      for (int a = 0; a < d; ++a)
        for (int b = 0; b < maxw; ++b) o[n++] = s[a * maxw + b];
      for (int a = d; a < maxh - d; ++a)
        for (int b = 0; b < d; ++b) o[n++] = s[a * maxw + b];
      for (int a = d; a < maxh - d; ++a)
        for (int b = maxw - d; b < maxw; ++b) o[n++] = s[a * maxw + b];
      for (int a = maxh - d; a < maxh; ++a)
        for (int b = 0; b < maxw; ++b) o[n++] = s[a * maxw + b];
    }
  }
}

/* nn.SpatialDownSampling(r, r): mean over r x r blocks (multiscale.lua:145;
 * out-of-tree nnx, PARITY UNPINNED: sum in fp32 then one multiply by 1/(r*r)) */
void orc_downsample_avg(const float *in, int C, int H, int W, int r, float *out) {
  const int h = H / r, w = W / r;
  const float norm = 1.0f / (float)(r * r);
  for (int c = 0; c < C; ++c)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        float s = 0.0f;
        for (int a = 0; a < r; ++a)
          for (int b = 0; b < r; ++b) s = s + in[((size_t)c * H + (y * r + a)) * W + (x * r + b)];
        out[((size_t)c * h + y) * w + x] = s * norm;
      }
}

/* nearest upsampling of an h x w map of K-vectors to (h*r) x (w*r) */
void orc_upsample_nearest_rows(const float *in, int h, int w, int K, int r, float *out) {
  const int H = h * r, W = w * r;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      memcpy(out + ((size_t)y * W + x) * K, in + ((size_t)(y / r) * w + (x / r)) * K,
             (size_t)K * sizeof(float));
}
