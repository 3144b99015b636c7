// radial.cu -- K4: the polar remap of the radial optical-flow path.
//
// The reference builds a look-up table on the CPU (getC2PMask / getP2CMask,
// radial/cartesian2polar.lua:4-89) and gathers through image.warp(..., 'bilinear',
// false).  Here the sampling coordinates are computed analytically inside the
// gather kernel -- same float/double expression tree as the reference's inline C,
// so the coordinates round identically -- and no LUT ever touches HBM.  The LUT
// builders and the LUT-driven warp are kept as separate entry points for callers
// that hold masks (radial/radial_opticalflow_groundtruth.lua:127-141).
#include "dm_common.cuh"

namespace dm {

// image.warp bilinear tap (Torch7 image package, 2012): clamp to the border, 4
// neighbours, MIN on the +1 neighbours.  Products and sums round separately.
__device__ __forceinline__ void bilinear_setup(float iy, float ix, int hs, int ws, long long *o00,
                                               long long *o01, long long *o10, long long *o11,
                                               float *wnw, float *wne, float *wsw, float *wse) {
  ix = ix > 0.0f ? ix : 0.0f;
  ix = ix < (float)(ws - 1) ? ix : (float)(ws - 1);
  iy = iy > 0.0f ? iy : 0.0f;
  iy = iy < (float)(hs - 1) ? iy : (float)(hs - 1);
  const int x0 = (int)floorf(ix), y0 = (int)floorf(iy);
  const int x1 = x0 + 1, y1 = y0 + 1;
  *wnw = __fmul_rn(__fsub_rn((float)x1, ix), __fsub_rn((float)y1, iy));
  *wne = __fmul_rn(__fsub_rn(ix, (float)x0), __fsub_rn((float)y1, iy));
  *wsw = __fmul_rn(__fsub_rn((float)x1, ix), __fsub_rn(iy, (float)y0));
  *wse = __fmul_rn(__fsub_rn(ix, (float)x0), __fsub_rn(iy, (float)y0));
  const int x1c = x1 < ws - 1 ? x1 : ws - 1, y1c = y1 < hs - 1 ? y1 : hs - 1;
  *o00 = (long long)y0 * ws + x0;
  *o01 = (long long)y0 * ws + x1c;
  *o10 = (long long)y1c * ws + x0;
  *o11 = (long long)y1c * ws + x1c;
}

__device__ __forceinline__ float bilinear_tap(const float *s, long long o00, long long o01,
                                              long long o10, long long o11, float wnw, float wne,
                                              float wsw, float wse) {
  const float a = __fmul_rn(__ldg(s + o00), wnw), b = __fmul_rn(__ldg(s + o01), wne);
  const float c = __fmul_rn(__ldg(s + o10), wsw), d = __fmul_rn(__ldg(s + o11), wse);
  return __fadd_rn(__fadd_rn(__fadd_rn(a, b), c), d);
}

struct C2P {
  float xcenter, ycenter, kr, ktheta, alpha;
  int wdst, hdst, lpad, rpad;
};
struct P2C {
  float xcenter, ycenter, kx, ky, pi2, invalpha;
  int wdst, hdst;
};

// radial/cartesian2polar.lua:31-38
__device__ __forceinline__ void c2p_coord(const C2P &m, int i, int j, float *y, float *x) {
  const float r = (float)((double)m.kr * pow((double)(float)i, (double)m.alpha));
  const float theta = __fmul_rn(m.ktheta, (float)j);
  *y = (float)((double)r * sin((double)theta) + (double)m.ycenter);
  *x = (float)((double)r * cos((double)theta) + (double)m.xcenter);
}
// radial/cartesian2polar.lua:77-84
__device__ __forceinline__ void p2c_coord(const P2C &m, int i, int j, float *row, float *col) {
  const float x = __fsub_rn((float)j, m.xcenter), y = __fsub_rn((float)i, m.ycenter);
  const float n2 = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
  *row = (float)(pow((double)n2, (double)m.invalpha) * (double)m.ky);
  *col = (float)(fmod(atan2((double)y, (double)x) + (double)m.pi2, (double)m.pi2) * (double)m.kx);
}

__global__ void c2p_mask_kernel(C2P m, float *mask) {
  const int wp = m.wdst + m.lpad + m.rpad;
  const long long plane = (long long)m.hdst * wp;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < plane;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / wp), jp = (int)(t % wp);
    int j = jp - m.lpad;  // circular padding (cartesian2polar.lua:41-46)
    if (j < 0) j += m.wdst;
    if (j >= m.wdst) j -= m.wdst;
    float y, x;
    c2p_coord(m, i, j, &y, &x);
    mask[t] = y;
    mask[plane + t] = x;
  }
}

__global__ void p2c_mask_kernel(P2C m, float *mask) {
  const long long plane = (long long)m.hdst * m.wdst;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < plane;
       t += (long long)gridDim.x * blockDim.x) {
    float row, col;
    p2c_coord(m, (int)(t / m.wdst), (int)(t % m.wdst), &row, &col);
    mask[t] = row;
    mask[plane + t] = col;
  }
}

// mode 0: analytic cartesian->polar, 1: analytic polar->cartesian, 2: LUT
__global__ void warp_kernel(int mode, C2P c2p, P2C p2c, const float *field, const float *src, int C,
                            int hs, int ws, int hd, int wd, float *dst) {
  const long long plane = (long long)hd * wd;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < plane;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / wd), jp = (int)(t % wd);
    float iy, ix;
    if (mode == 0) {
      int j = jp - c2p.lpad;
      if (j < 0) j += c2p.wdst;
      if (j >= c2p.wdst) j -= c2p.wdst;
      c2p_coord(c2p, i, j, &iy, &ix);
    } else if (mode == 1) {
      p2c_coord(p2c, i, jp, &iy, &ix);
    } else {
      iy = field[t];
      ix = field[plane + t];
    }
    long long o00, o01, o10, o11;
    float wnw, wne, wsw, wse;
    bilinear_setup(iy, ix, hs, ws, &o00, &o01, &o10, &o11, &wnw, &wne, &wsw, &wse);
    for (int k = 0; k < C; ++k)
      dst[(long long)k * plane + t] =
          bilinear_tap(src + (long long)k * hs * ws, o00, o01, o10, o11, wnw, wne, wsw, wse);
  }
}

// sfm2.removeEgoMotion(im, K, R) (out-of-tree sfm2; call sites depth_estimation_api.lua:147,
// radial/test_radial_opticalflow.lua:192): the previous frame (or its feature maps) seen through
// the rotated camera, i.e. a homography gather.  Destination pixel (x, y) samples the source at
// Hm * (x, y, 1) (double arithmetic, one division), bilinear like image.warp; pixels whose
// source falls outside the frame read 0 and clear the mask.
struct Homography {
  double m[9];
};
__global__ void homography_kernel(Homography Hm, const float *src, int C, int hs, int ws, int hd, int wd,
                                  float *dst, float *mask) {
  const long long plane = (long long)hd * wd;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < plane;
       t += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(t / wd), x = (int)(t % wd);
    const double X = Hm.m[0] * x + Hm.m[1] * y + Hm.m[2];
    const double Y = Hm.m[3] * x + Hm.m[4] * y + Hm.m[5];
    const double Z = Hm.m[6] * x + Hm.m[7] * y + Hm.m[8];
    const float ix = (float)(X / Z), iy = (float)(Y / Z);
    const bool inside = Z > 0.0 && ix >= 0.0f && ix <= (float)(ws - 1) && iy >= 0.0f && iy <= (float)(hs - 1);
    long long o00 = 0, o01 = 0, o10 = 0, o11 = 0;
    float wnw = 0.0f, wne = 0.0f, wsw = 0.0f, wse = 0.0f;
    if (inside) bilinear_setup(iy, ix, hs, ws, &o00, &o01, &o10, &o11, &wnw, &wne, &wsw, &wse);
    for (int k = 0; k < C; ++k)
      dst[(long long)k * plane + t] =
          inside ? bilinear_tap(src + (long long)k * hs * ws, o00, o01, o10, o11, wnw, wne, wsw, wse) : 0.0f;
    if (mask) mask[t] = inside ? 1.0f : 0.0f;
  }
}

// radial/radial_opticalflow_display.lua:28-53
__global__ void flow2depth_kernel(const float *flow, int h, int w, float xc, float yc, float infty,
                                  float *depth, float *confs) {
  const long long total = (long long)h * w;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / w), j = (int)(t % w);
    const float dx = __fsub_rn((float)j, xc), dy = __fsub_rn((float)i, yc);
    const float d = (float)sqrt((double)__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    float out = 0.0f, conf = 1.0f;
    if (d > 10.0f) {
      const float f = flow[t];
      out = f < 0.1f ? infty : __fdiv_rn(d, f);
    } else {
      conf = 0.0f;
    }
    depth[t] = out;
    confs[t] = conf;
  }
}

static int blocks_for(dm_ctx *ctx, long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)ctx->num_sms * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

static C2P make_c2p(int wdst, int hdst, double xc, double yc, int lpad, int rpad, double rmax,
                    double alpha) {
  C2P m;
  m.xcenter = (float)xc;
  m.ycenter = (float)yc;
  m.kr = (float)(rmax / pow((double)hdst, alpha));  // cartesian2polar.lua:12
  m.ktheta = (float)(2.0 * M_PI / (double)wdst);    // :13
  m.alpha = (float)alpha;
  m.wdst = wdst;
  m.hdst = hdst;
  m.lpad = lpad;
  m.rpad = rpad;
  return m;
}

static P2C make_p2c(int wsrc, int hsrc, int wdst, int hdst, double xc, double yc, double rmax,
                    double alpha) {
  P2C m;
  const double pi2 = 2.0 * M_PI;
  m.xcenter = (float)xc;
  m.ycenter = (float)yc;
  m.kx = (float)((double)wsrc / pi2);                  // cartesian2polar.lua:58
  m.ky = (float)((double)hsrc / pow(rmax, 1.0 / alpha));  // :59
  m.pi2 = (float)pi2;
  m.invalpha = (float)(1.0 / alpha) * 0.5f;  // :69
  m.wdst = wdst;
  m.hdst = hdst;
  return m;
}

}  // namespace dm

using namespace dm;

extern "C" {

int dm_c2p_mask(dm_ctx *ctx, int wdst, int hdst, double xcenter, double ycenter, int lpad, int rpad,
                double rmax, double alpha, float *mask) {
  DM_REQUIRE(ctx && mask, "dm_c2p_mask: NULL argument");
  DM_REQUIRE(wdst >= 1 && hdst >= 1 && lpad >= 0 && rpad >= 0 && lpad <= wdst && rpad <= wdst,
             "dm_c2p_mask: bad shape");
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  void *d;
  const long long plane = (long long)hdst * (wdst + lpad + rpad);
  DM_CHECK(call.out(mask, (size_t)plane * 2 * 4, &d));
  c2p_mask_kernel<<<blocks_for(ctx, plane), 256, 0, ctx->stream>>>(
      make_c2p(wdst, hdst, xcenter, ycenter, lpad, rpad, rmax, alpha), static_cast<float *>(d));
  count_launch(ctx);
  return call.finish();
}

int dm_p2c_mask(dm_ctx *ctx, int wsrc, int hsrc, int wdst, int hdst, double xcenter, double ycenter,
                double rmax, double alpha, float *mask) {
  DM_REQUIRE(ctx && mask, "dm_p2c_mask: NULL argument");
  DM_REQUIRE(wdst >= 1 && hdst >= 1 && wsrc >= 1 && hsrc >= 1, "dm_p2c_mask: bad shape");
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  void *d;
  const long long plane = (long long)hdst * wdst;
  DM_CHECK(call.out(mask, (size_t)plane * 2 * 4, &d));
  p2c_mask_kernel<<<blocks_for(ctx, plane), 256, 0, ctx->stream>>>(
      make_p2c(wsrc, hsrc, wdst, hdst, xcenter, ycenter, rmax, alpha), static_cast<float *>(d));
  count_launch(ctx);
  return call.finish();
}

static int warp_common(dm_ctx *ctx, int mode, const C2P &c2p, const P2C &p2c, const float *field,
                       const float *src, int c, int hs, int ws, int hd, int wd, float *dst) {
  DM_REQUIRE(c >= 1 && hs >= 1 && ws >= 1 && hd >= 1 && wd >= 1, "warp: bad shape");
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  const void *dsrc, *dfield = nullptr;
  void *ddst;
  DM_CHECK(call.in(src, (size_t)c * hs * ws * 4, &dsrc));
  if (field) DM_CHECK(call.in(field, (size_t)2 * hd * wd * 4, &dfield));
  DM_CHECK(call.out(dst, (size_t)c * hd * wd * 4, &ddst));
  warp_kernel<<<blocks_for(ctx, (long long)hd * wd), 256, 0, ctx->stream>>>(
      mode, c2p, p2c, static_cast<const float *>(dfield), static_cast<const float *>(dsrc), c, hs, ws,
      hd, wd, static_cast<float *>(ddst));
  count_launch(ctx);
  return call.finish();
}

int dm_polar_remap(dm_ctx *ctx, const float *src, int c, int hsrc, int wsrc, double xcenter,
                   double ycenter, double rmax, double alpha, int lpad, int rpad, float *dst,
                   int hdst, int wdst) {
  DM_REQUIRE(ctx && src && dst, "dm_polar_remap: NULL argument");
  DM_REQUIRE(lpad >= 0 && rpad >= 0 && lpad <= wdst && rpad <= wdst, "dm_polar_remap: bad padding");
  P2C none;
  memset(&none, 0, sizeof(none));
  return warp_common(ctx, 0, make_c2p(wdst, hdst, xcenter, ycenter, lpad, rpad, rmax, alpha), none,
                     nullptr, src, c, hsrc, wsrc, hdst, wdst + lpad + rpad, dst);
}

int dm_polar_unmap(dm_ctx *ctx, const float *src, int c, int hsrc, int wsrc, double xcenter,
                   double ycenter, double rmax, double alpha, float *dst, int hdst, int wdst) {
  DM_REQUIRE(ctx && src && dst, "dm_polar_unmap: NULL argument");
  C2P none;
  memset(&none, 0, sizeof(none));
  return warp_common(ctx, 1, none, make_p2c(wsrc, hsrc, wdst, hdst, xcenter, ycenter, rmax, alpha),
                     nullptr, src, c, hsrc, wsrc, hdst, wdst, dst);
}

int dm_warp_bilinear(dm_ctx *ctx, const float *src, int c, int hs, int ws, const float *field, int hd,
                     int wd, float *dst) {
  DM_REQUIRE(ctx && src && field && dst, "dm_warp_bilinear: NULL argument");
  C2P a;
  P2C b;
  memset(&a, 0, sizeof(a));
  memset(&b, 0, sizeof(b));
  return warp_common(ctx, 2, a, b, field, src, c, hs, ws, hd, wd, dst);
}

int dm_warp_homography(dm_ctx *ctx, const float *src, int c, int hs, int ws, const double *hmat, int hd,
                       int wd, float *dst, float *mask) {
  DM_REQUIRE(ctx && src && hmat && dst, "dm_warp_homography: NULL argument");
  DM_REQUIRE(c >= 1 && hs >= 1 && ws >= 1 && hd >= 1 && wd >= 1, "dm_warp_homography: bad shape");
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  const void *ds;
  void *dd, *dmask = nullptr;
  DM_CHECK(call.in(src, (size_t)c * hs * ws * 4, &ds));
  DM_CHECK(call.out(dst, (size_t)c * hd * wd * 4, &dd));
  if (mask) DM_CHECK(call.out(mask, (size_t)hd * wd * 4, &dmask));
  Homography Hm;
  for (int i = 0; i < 9; ++i) Hm.m[i] = hmat[i];  // host pointer: 9 doubles, read now
  const long long plane = (long long)hd * wd;
  int blocks = (int)((plane + 255) / 256);
  if (blocks > ctx->num_sms * 16) blocks = ctx->num_sms * 16;
  homography_kernel<<<blocks, 256, 0, ctx->stream>>>(Hm, static_cast<const float *>(ds), c, hs, ws, hd, wd,
                                                     static_cast<float *>(dd), static_cast<float *>(dmask));
  count_launch(ctx);
  return call.finish();
}

int dm_flow2depth(dm_ctx *ctx, const float *flow, int h, int w, float xcenter, float ycenter,
                  float infty, float *depth, float *confs) {
  DM_REQUIRE(ctx && flow && depth && confs, "dm_flow2depth: NULL argument");
  DM_REQUIRE(h >= 1 && w >= 1, "dm_flow2depth: bad shape");
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  const void *dflow;
  void *dd, *dc;
  const long long total = (long long)h * w;
  DM_CHECK(call.in(flow, (size_t)total * 4, &dflow));
  DM_CHECK(call.out(depth, (size_t)total * 4, &dd));
  DM_CHECK(call.out(confs, (size_t)total * 4, &dc));
  flow2depth_kernel<<<blocks_for(ctx, total), 256, 0, ctx->stream>>>(
      static_cast<const float *>(dflow), h, w, xcenter, ycenter, infty, static_cast<float *>(dd),
      static_cast<float *>(dc));
  count_launch(ctx);
  return call.finish();
}

}  // extern "C"
