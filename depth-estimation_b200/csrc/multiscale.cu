// multiscale.cu -- K3: the multiscale cascade.
//
//   dm_x2yx_multi        ring index -> (dy,dx)       x2yxMulti2.c / x2yxMultiNumber
//   dm_cascade_add       nn.CascadingAddTable forward (CascadingAddTable.lua:108-135)
//   dm_multiscale_extract  per-scale matching + softmax, then per full-resolution pixel the
//                        cascade, the ring join, the argmax with the middle-index tie rule
//                        and the decode, fused: the L-vector (opticalflow_model_multiscale.lua
//                        :293-333) is never stored.
//   dm_downsample_avg    nn.SpatialDownSampling(r, r)
#include "dm_common.cuh"

namespace dm {

int match_volume_on(Call &call, const dm_pair *in, int maxh, int maxw, int mode, float *out);

constexpr int kMaxRatios = 10;  // N_MAX_RATIOS of x2yxMulti2.c:1

struct Ratios {
  int n;
  int r[kMaxRatios];
  int d[kMaxRatios];    // ring border of scale i (Lua spec), d[0] unused
  int len[kMaxRatios];  // ring length of scale i
  // bug-compatible tables of x2yxMulti2.c
  int rc[kMaxRatios], bd[kMaxRatios], bl[kMaxRatios];
};

static int lua_round_i(double v) { return (int)floor(v + 0.5); }

static Ratios make_ratios(int maxh, int maxw, const int *ratios, int n) {
  Ratios R;
  memset(&R, 0, sizeof(R));
  R.n = n;
  for (int i = 0; i < n; ++i) R.r[i] = ratios[i];
  for (int i = 1; i < n; ++i) {
    R.d[i] = lua_round_i((double)maxw * (ratios[i] - ratios[i - 1]) / (2.0 * ratios[i]));
    R.len[i] = 2 * R.d[i] * maxw + 2 * (maxh - 2 * R.d[i]) * R.d[i];
  }
  for (int i = 0; i < n; ++i) R.rc[i] = i == 0 ? 0 : ratios[i - 1];  // x2yxMulti2.c:15-19
  for (int i = 1; i < n; ++i) {
    R.bd[i] = R.rc[i] ? (int)round((float)maxw * ((float)R.rc[i] - (float)R.rc[i - 1]) /
                                   (2.0f * (float)R.rc[i]))
                      : 0;
    R.bl[i] = 2 * maxw + 2 * (maxh - 2 * R.bd[i]) * R.bd[i];  // x2yxMulti2.c:41
  }
  return R;
}

// x2yxMultiNumber (opticalflow_model_multiscale.lua:83-132); false where the Lua asserts
__device__ __forceinline__ bool decode_spec(const Ratios &R, int maxh, int maxw, long long x,
                                            long long *oy, long long *ox) {
  const long long cy = (maxh + 1) / 2, cx = (maxw + 1) / 2;
  const long long area = (long long)maxh * maxw;
  if (x <= area) {
    if (x < 1) return false;
    *oy = (x - 1) / maxw + 1 - cy;
    *ox = (x - 1) % maxw + 1 - cx;
    return true;
  }
  x -= area;
  for (int i = 1; i < R.n; ++i) {
    const int d = R.d[i];
    if (x > R.len[i]) {
      x -= R.len[i];
      continue;
    }
    const long long side = (long long)(maxh - 2 * d) * d, top = (long long)d * maxw;
    long long ty, tx;
    if (x <= top) {
      ty = (x - 1) / maxw + 1;
      tx = (x - 1) % maxw + 1;
    } else if (x - top <= side) {
      x -= top;
      ty = (x - 1) / d + 1 + d;
      tx = (x - 1) % d + 1;
    } else if (x - top - side <= side) {
      x -= top + side;
      ty = (x - 1) / d + 1 + d;
      tx = (x - 1) % d + 1 + maxw - d;
    } else if (x - top - 2 * side <= top) {
      x -= top + 2 * side;
      ty = (x - 1) / maxw + 1 + maxh - d;
      tx = (x - 1) % maxw + 1;
    } else {
      return false;
    }
    *oy = (ty - cy) * R.r[i];
    *ox = (tx - cx) * R.r[i];
    return true;
  }
  return false;
}

// x2yxMulti2.c:46-91 with its divergences; false = the C falls through (entry untouched)
__device__ __forceinline__ bool decode_bugcompat(const Ratios &R, int maxh, int maxw, long long x,
                                                 long long *oy, long long *ox) {
  const int chh = maxh / 2, chw = maxw / 2;
  const long long area = (long long)maxh * maxw;
  if (x < area) {
    *oy = (x - 1) / maxw + 1 - chh;
    *ox = (x - 1) % maxw + 1 - chw;
    return true;
  }
  x -= area;
  for (int k = 1; k < R.n; ++k) {
    const int d = R.bd[k];
    const long long mH = (long long)(maxh - 2 * d) * d;
    if (x > R.bl[k]) {
      x -= R.bl[k];
      continue;
    }
    if (x < (long long)d * maxw) {
      *oy = ((x - 1) / maxw + 1 - chh) * R.rc[k];
      *ox = ((x - 1) % maxw + 1 - chw) * R.rc[k];
      return true;
    }
    x -= (long long)d * maxw;
    if (d == 0) continue;
    if (x <= mH) {
      *oy = ((x - 1) / d + 1 + d - chh) * R.rc[k];
      *ox = ((x - 1) % d + 1 - chw) * R.rc[k];
      return true;
    }
    x -= mH;
    if (x <= mH) {
      *oy = ((x - 1) / d + 1 + d - chh) * R.rc[k];
      *ox = ((x - 1) % d + 1 + maxw - d - chw) * R.rc[k];
      return true;
    }
    x -= mH;
    if (x < (long long)d * maxw) {
      *oy = ((x - 1) / maxw + 1 + maxh - d - chh) * R.rc[k];
      *ox = ((x - 1) % maxw + 1 - chw) * R.rc[k];
      return true;
    }
  }
  return false;
}

__global__ void x2yx_multi_kernel(const long long *x, long long total, int maxh, int maxw, Ratios R,
                                  int bug_compat, long long *rety, long long *retx) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    long long oy, ox;
    const bool ok = bug_compat ? decode_bugcompat(R, maxh, maxw, x[t], &oy, &ox)
                               : decode_spec(R, maxh, maxw, x[t], &oy, &ox);
    if (ok) {
      rety[t] = oy;
      retx[t] = ox;
    } else if (!bug_compat) {
      rety[t] = 0;
      retx[t] = 0;
    }
  }
}

// CascadingAddTable forward: value of cascade output `i` at window entry (a,b) of
// row `row`, given the per-scale inputs laid out [scale][rows][kh][kw].
// out[n-1] = in[n-1]; out[i] = in[i] + up(crop(out[i+1])) -- summed coarse to fine.
__device__ __forceinline__ float cascade_value(const float *in, long long rows, long long row, int kh,
                                               int kw, const Ratios &R, int i, int a, int b) {
  const long long per = rows * kh * kw;
  long long off[kMaxRatios];
  int depth = 0;
  for (int s = i;; ++s) {
    off[depth++] = (long long)s * per + (row * kh + a) * kw + b;
    if (s + 1 >= R.n) break;
    const int r = R.r[s], r2 = R.r[s + 1];
    const int f = r2 / r;
    a = kh * (r2 - r) / (2 * r2) + a / f;
    b = kw * (r2 - r) / (2 * r2) + b / f;
  }
  float v = in[off[depth - 1]];
  for (int j = depth - 2; j >= 0; --j) v = __fadd_rn(in[off[j]], v);
  return v;
}

__global__ void cascade_add_kernel(const float *in, long long rows, int kh, int kw, Ratios R,
                                   float *out) {
  const long long per = rows * kh * kw, total = per * R.n;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / per);
    const long long e = t - (long long)i * per;
    const int b = (int)(e % kw), a = (int)((e / kw) % kh);
    const long long row = e / ((long long)kw * kh);
    out[t] = cascade_value(in, rows, row, kh, kw, R, i, a, b);
  }
}

struct ScaleMaps {
  const float *p[kMaxRatios];  // per-scale softmax maps [h/r][w/r][maxh*maxw]
};

// Entry l of the joined vector (opticalflow_model_multiscale.lua:293-324) -> its cascade chain
// (CascadingAddTable.lua:108-135): off[l * n + s] = window offset a*maxw+b to read in scale s's
// map, or -1 when scale s is not on the chain.  Built on the host once per call.
static void build_chain_table(const Ratios &R, int maxh, int maxw, int L, std::vector<int> *tab) {
  const int n = R.n, K = maxh * maxw;
  tab->assign((size_t)L * n, -1);
  for (int l = 0; l < L; ++l) {
    int i = 0, a, b, e = l;
    if (e < K) {
      a = e / maxw;
      b = e - a * maxw;
    } else {
      e -= K;
      i = 1;
      while (e >= R.len[i]) e -= R.len[i++];
      const int d = R.d[i], top = d * maxw, side = (maxh - 2 * d) * d;
      if (e < top) {
        a = e / maxw;
        b = e - a * maxw;
      } else if (e < top + side) {
        e -= top;
        a = d + e / d;
        b = e % d;
      } else if (e < top + 2 * side) {
        e -= top + side;
        a = d + e / d;
        b = maxw - d + e % d;
      } else {
        e -= top + 2 * side;
        a = maxh - d + e / maxw;
        b = e % maxw;
      }
    }
    for (int s = i;; ++s) {
      (*tab)[(size_t)l * n + s] = a * maxw + b;
      if (s + 1 >= n) break;
      const int rs = R.r[s], r2 = R.r[s + 1], f = r2 / rs;
      a = maxh * (r2 - rs) / (2 * r2) + a / f;
      b = maxw * (r2 - rs) / (2 * r2) + b / f;
    }
  }
}

// One warp per full-resolution pixel, lanes over the L entries of the joined vector.  Scale s
// reads the nearest-upsampled map, i.e. pixel (Y / r_s, X / r_s); the chain is summed coarse to
// fine like CascadingAddTable does (out[n] = in[n]; out[i] = in[i] + up(out[i+1])).
__global__ void __launch_bounds__(256)
ring_argmax_kernel(ScaleMaps maps, const int *chain, int h, int w, int maxh, int maxw, Ratios R, int L,
                   int middle, long long *index, long long *flow_y, long long *flow_x) {
  const int lane = threadIdx.x & 31;
  const int K = maxh * maxw, n = R.n;
  const long long npx = (long long)h * w;
  // a lane sees the same entries l = lane + 32 e for every pixel: their per-scale offsets live
  // in registers (L <= 256 and at most three scales, which covers the reference's geometries: L = 112, 160)
  constexpr int kLaneEntries = 8, kRegRatios = 3;
  const bool in_regs = L <= 32 * kLaneEntries && n <= kRegRatios;
  int offs[kLaneEntries][kRegRatios];
  if (in_regs) {
#pragma unroll
    for (int e = 0; e < kLaneEntries; ++e)
#pragma unroll
      for (int s = 0; s < kRegRatios; ++s) {
        const int l = lane + 32 * e;
        offs[e][s] = (l < L && s < n) ? __ldg(chain + (size_t)l * n + s) : -1;
      }
  }
  // a warp takes kBatch consecutive pixels: one gather + shuffle reduction per pixel, then the
  // first kBatch lanes decode and store the results (coalesced 8-byte stores).  Small batches keep
  // the last wave of the grid short (a 640x360 frame is only ~2 batches of 32 per resident warp).
  constexpr int kBatch = 8;
  for (long long px0 = ((long long)blockIdx.x * 8 + __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0)) * kBatch; px0 < npx;
       px0 += (long long)gridDim.x * 8 * kBatch) {
   long long mywin = 0;
   // per-scale offset of the window of pixel px0 + lane, computed by its lane (the divisions
   // run once per 32 pixels) and broadcast inside the loop
   int myoff[kRegRatios];
   {
     const long long pxl = px0 + lane < npx ? px0 + lane : npx - 1;
     const int Yl = (int)(pxl / w), Xl = (int)(pxl - (long long)Yl * w);
#pragma unroll
     for (int s = 0; s < kRegRatios; ++s) {
       const int rs = s < n ? R.r[s] : 1;
       myoff[s] = ((Yl / rs) * (w / rs) + (Xl / rs)) * K;
     }
   }
   for (int j = 0; j < kBatch && px0 + j < npx; ++j) {
    const long long px = px0 + j;
    const float *base[kMaxRatios];
    if (in_regs) {
#pragma unroll
      for (int s = 0; s < kRegRatios; ++s) base[s] = maps.p[s] + __shfl_sync(0xffffffffu, myoff[s], j);
    } else {
      const int Y = (int)(px / w), X = (int)(px % w);
#pragma unroll
      for (int s = 0; s < kMaxRatios; ++s)
        if (s < n) {
          const int rs = R.r[s];
          base[s] = maps.p[s] + ((long long)(Y / rs) * (w / rs) + (X / rs)) * K;
        }
    }
    float best = -__int_as_float(0x7f800000), vmid = 0.0f;
    int lb = 0x7fffffff;
    if (in_regs) {
      const int ne = (L + 31) >> 5;
#pragma unroll
      for (int e = 0; e < kLaneEntries; ++e) {
        const int l = lane + 32 * e;
        if (e < ne && l < L) {
          float v = 0.0f;
          bool first = true;
#pragma unroll
          for (int s = kRegRatios - 1; s >= 0; --s)
            if (s < n && offs[e][s] >= 0) {
              const float x = __ldg(base[s] + offs[e][s]);
              v = first ? x : __fadd_rn(x, v);
              first = false;
            }
          if (l + 1 == middle) vmid = v;
          if (v > best) {
            best = v;
            lb = l;
          }
        }
      }
    } else
    for (int l = lane; l < L; l += 32) {
      const int *off = chain + (size_t)l * n;
      float v = 0.0f;
      bool first = true;
#pragma unroll
      for (int s = kMaxRatios - 1; s >= 0; --s)
        if (s < n) {
          const int o = __ldg(off + s);
          if (o >= 0) {
            const float x = __ldg(base[s] + o);
            v = first ? x : __fadd_rn(x, v);
            first = false;
          }
        }
      if (l + 1 == middle) vmid = v;
      if (v > best) {
        best = v;
        lb = l;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int ol = __shfl_xor_sync(0xffffffffu, lb, o);
      vmid += __shfl_xor_sync(0xffffffffu, vmid, o);  // exactly one lane holds it
      if (ob > best || (ob == best && ol < lb)) {
        best = ob;
        lb = ol;
      }
    }
    long long win = lb + 1;
    if (vmid == best) win = middle;  // opticalflow_model.lua:157-159 with yx2xMulti(0,0)
    if (lane == j) mywin = win;      // every lane holds the reduced values
   }
   if (lane < kBatch && px0 + lane < npx) {
     long long oy = 0, ox = 0;
     decode_spec(R, maxh, maxw, mywin, &oy, &ox);
     if (index) index[px0 + lane] = mywin;
     if (flow_y) flow_y[px0 + lane] = oy;
     if (flow_x) flow_x[px0 + lane] = ox;
   }
  }
}

__global__ void downsample_avg_kernel(const float *in, int C, int H, int W, int r, float *out) {
  const int h = H / r, w = W / r;
  const long long total = (long long)C * h * w;
  const float norm = 1.0f / (float)(r * r);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(t % w), y = (int)((t / w) % h), c = (int)(t / ((long long)w * h));
    float s = 0.0f;
    for (int a = 0; a < r; ++a)
      for (int b = 0; b < r; ++b)
        s = __fadd_rn(s, in[((long long)c * H + (y * r + a)) * W + (x * r + b)]);
    out[t] = __fmul_rn(s, norm);
  }
}

static int flat_blocks(dm_ctx *ctx, long long n, int block) {
  long long b = (n + block - 1) / block;
  const long long cap = (long long)ctx->num_sms * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

static int check_ratios(int maxh, int maxw, const int *ratios, int nratios) {
  DM_REQUIRE(ratios && nratios >= 1 && nratios <= kMaxRatios, "ratios: need 1..%d entries", kMaxRatios);
  DM_REQUIRE(maxh >= 1 && maxw >= 1, "bad window %dx%d", maxh, maxw);
  for (int i = 0; i < nratios; ++i) {
    DM_REQUIRE(ratios[i] >= 1, "ratios must be positive");
    if (i) DM_REQUIRE(ratios[i] > ratios[i - 1], "ratios must be increasing");
  }
  return DM_OK;
}

}  // namespace dm

using namespace dm;

extern "C" {

int dm_x2yx_multi(dm_ctx *ctx, const int64_t *x, int h, int w, int maxh, int maxw,
                  const int *ratios, int nratios, int bug_compat, int64_t *rety, int64_t *retx) {
  DM_REQUIRE(ctx && x && rety && retx, "dm_x2yx_multi: NULL argument");
  DM_REQUIRE(h >= 0 && w >= 0, "dm_x2yx_multi: bad shape");
  DM_CHECK(check_ratios(maxh, maxw, ratios, nratios));
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  const long long total = (long long)h * w;
  if (total == 0) return call.finish();
  const void *dx;
  void *dry, *drx;
  DM_CHECK(call.in(x, (size_t)total * 8, &dx));
  DM_CHECK(call.out(rety, (size_t)total * 8, &dry, bug_compat != 0));
  DM_CHECK(call.out(retx, (size_t)total * 8, &drx, bug_compat != 0));
  const Ratios R = make_ratios(maxh, maxw, ratios, nratios);
  x2yx_multi_kernel<<<flat_blocks(ctx, total, 256), 256, 0, ctx->stream>>>(
      static_cast<const long long *>(dx), total, maxh, maxw, R, bug_compat,
      static_cast<long long *>(dry), static_cast<long long *>(drx));
  count_launch(ctx);
  return call.finish();
}

int dm_cascade_add(dm_ctx *ctx, const float *in, int64_t rows, int kh, int kw, const int *ratios,
                   int nratios, float *out) {
  DM_REQUIRE(ctx && in && out, "dm_cascade_add: NULL argument");
  DM_CHECK(check_ratios(kh, kw, ratios, nratios));
  DM_REQUIRE(rows >= 0, "dm_cascade_add: bad rows");
  for (int i = 0; i + 1 < nratios; ++i) {
    const int r = ratios[i], r2 = ratios[i + 1];
    // CascadingAddTable.lua:121-124
    DM_REQUIRE((kh * (r2 - r)) % (2 * r2) == 0 && (kw * (r2 - r)) % (2 * r2) == 0 && r2 % r == 0,
               "nn.CascadingAddTable: ratios and input sizes not compatible");
  }
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  if (rows == 0) return call.finish();
  const size_t bytes = (size_t)nratios * rows * kh * kw * sizeof(float);
  const void *din;
  void *dout;
  DM_CHECK(call.in(in, bytes, &din));
  DM_CHECK(call.out(out, bytes, &dout));
  const Ratios R = make_ratios(kh, kw, ratios, nratios);
  const long long total = (long long)nratios * rows * kh * kw;
  cascade_add_kernel<<<flat_blocks(ctx, total, 256), 256, 0, ctx->stream>>>(
      static_cast<const float *>(din), rows, kh, kw, R, static_cast<float *>(dout));
  count_launch(ctx);
  return call.finish();
}

int dm_downsample_avg(dm_ctx *ctx, const float *in, int c, int h, int w, int r, float *out) {
  DM_REQUIRE(ctx && in && out, "dm_downsample_avg: NULL argument");
  DM_REQUIRE(c >= 1 && h >= r && w >= r && r >= 1, "dm_downsample_avg: bad shape");
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  const void *din;
  void *dout;
  DM_CHECK(call.in(in, (size_t)c * h * w * 4, &din));
  const long long total = (long long)c * (h / r) * (w / r);
  DM_CHECK(call.out(out, (size_t)total * 4, &dout));
  downsample_avg_kernel<<<flat_blocks(ctx, total, 256), 256, 0, ctx->stream>>>(
      static_cast<const float *>(din), c, h, w, r, static_cast<float *>(dout));
  count_launch(ctx);
  return call.finish();
}

int dm_multiscale_extract(dm_ctx *ctx, const float *const *in1, const float *const *in2,
                          int channels, int h, int w, int maxh, int maxw, const int *ratios,
                          int nratios, int64_t *index, int64_t *flow_y, int64_t *flow_x) {
  DM_REQUIRE(ctx && in1 && in2, "dm_multiscale_extract: NULL argument");
  DM_CHECK(check_ratios(maxh, maxw, ratios, nratios));
  DM_REQUIRE(ratios[0] == 1, "getModelMultiscale: ratios[1] must be 1");  // multiscale.lua:182
  const int rmax = ratios[nratios - 1];
  DM_REQUIRE(h % rmax == 0 && w % rmax == 0, "frame %dx%d is not a multiple of the largest ratio %d",
             h, w, rmax);
  for (int i = 0; i + 1 < nratios; ++i) {
    const int r = ratios[i], r2 = ratios[i + 1];
    DM_REQUIRE((maxh * (r2 - r)) % (2 * r2) == 0 && (maxw * (r2 - r)) % (2 * r2) == 0 && r2 % r == 0,
               "nn.CascadingAddTable: ratios and input sizes not compatible");
  }
  const Ratios R = make_ratios(maxh, maxw, ratios, nratios);
  const int K = maxh * maxw;
  int L = K;
  for (int i = 1; i < nratios; ++i) L += R.len[i];
  // one Call for everything: the per-scale soft-max maps and the chain table live in its arena
  size_t total = 0;
  for (int i = 0; i < nratios; ++i) total += (size_t)(h / ratios[i]) * (w / ratios[i]) * K;
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  void *mp = nullptr, *tp = nullptr;
  DM_CHECK(call.alloc(&mp, total * sizeof(float)));
  float *maps_dev = static_cast<float *>(mp);
  std::vector<int> table;
  build_chain_table(R, maxh, maxw, L, &table);
  DM_CHECK(call.alloc(&tp, table.size() * sizeof(int)));
  DM_CUDA(cudaMemcpyAsync(tp, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice,
                          ctx->stream));  // pageable source: staged before the call returns
  ScaleMaps maps;
  memset(&maps, 0, sizeof(maps));
  size_t off = 0;
  int rc = DM_OK;
  // the scales are independent until the ring join: scale 0 stays on the context's stream, the
  // others alternate over two side streams forked from it (their small grids overlap with the
  // full-resolution scale instead of queueing behind it)
  cudaStream_t main_stream = ctx->stream;
  const bool fork = nratios > 1;
  if (fork && !ctx->aux[0]) {
    for (int i = 0; i < 2; ++i) {
      DM_CUDA(cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking));
      DM_CUDA(cudaEventCreateWithFlags(&ctx->aux_join[i], cudaEventDisableTiming));
    }
    DM_CUDA(cudaEventCreateWithFlags(&ctx->aux_fork, cudaEventDisableTiming));
  }
  if (fork) {
    DM_CUDA(cudaEventRecord(ctx->aux_fork, main_stream));
    for (int i = 0; i < 2; ++i) DM_CUDA(cudaStreamWaitEvent(ctx->aux[i], ctx->aux_fork, 0));
  }
  bool used[2] = {false, false};
  for (int i = 0; i < nratios && rc == DM_OK; ++i) {
    dm_pair pr;
    memset(&pr, 0, sizeof(pr));
    pr.in1 = in1[i];
    pr.in2 = in2[i];
    pr.n_pairs = 1;
    pr.channels = channels;
    pr.h1 = h / ratios[i];
    pr.w1 = w / ratios[i];
    pr.h2 = pr.h1 + maxh - 1;
    pr.w2 = pr.w1 + maxw - 1;
    maps.p[i] = maps_dev + off;
    if (i > 0) {
      ctx->stream = ctx->aux[(i - 1) & 1];
      used[(i - 1) & 1] = true;
    }
    rc = match_volume_on(call, &pr, maxh, maxw, DM_VOLUME_NEG_SOFTMAX, maps_dev + off);
    ctx->stream = main_stream;
    off += (size_t)pr.h1 * pr.w1 * K;
  }
  for (int i = 0; i < 2; ++i)
    if (used[i]) {
      cudaEventRecord(ctx->aux_join[i], ctx->aux[i]);
      cudaStreamWaitEvent(main_stream, ctx->aux_join[i], 0);
    }
  if (rc == DM_OK) {
    const long long npx = (long long)h * w;
    void *didx = nullptr, *dfy = nullptr, *dfx = nullptr;
    if (index) rc = call.out(index, (size_t)npx * 8, &didx);
    if (rc == DM_OK && flow_y) rc = call.out(flow_y, (size_t)npx * 8, &dfy);
    if (rc == DM_OK && flow_x) rc = call.out(flow_x, (size_t)npx * 8, &dfx);
    if (rc == DM_OK) {
      // middle index = yx2xMulti(0, 0): zero flow lives in the scale-1 block
      const int middle = ((maxh + 1) / 2 - 1) * maxw + (maxw + 1) / 2;
      long long blocks = (npx + 63) / 64;  // 8 warps x 8 pixels per pass
      const long long cap = (long long)ctx->num_sms * 3;  // 80 registers: three CTAs per SM
      if (blocks > cap) blocks = cap;
      ring_argmax_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(
          maps, static_cast<const int *>(tp), h, w, maxh, maxw, R, L, middle,
          static_cast<long long *>(didx), static_cast<long long *>(dfy), static_cast<long long *>(dfx));
      count_launch(ctx);
    }
  }
  const int rf = call.finish();
  return rc != DM_OK ? rc : rf;
}

}  // extern "C"
