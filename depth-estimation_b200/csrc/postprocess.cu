// postprocess.cu -- the steps right after the matching path ("next" rows 1-2):
//   dm_post_process_image  postProcessImage: masked k x k median / mode filter of the flow
//                          (opticalflow_model.lua:323-472, inline C compiled at run time there)
//   dm_enlarge_mask        enlargeMask (depth_estimation_api.lua:76-132)
//   dm_radial_depth        radial(): depth = |p - c| / |flow| (test_opticalflow.lua:143-193)
//   dm_depth_from_xflow    ARdroneAPI::computeDepthMapFromFlow (ardrone/ardrone_api.cpp:99-140)
// Small stencil / element-wise kernels, one thread per output pixel; bit-exact with the
// reference's C (integer and comparison work, fp32 divisions with IEEE rounding).
#include "dm_common.cuh"

namespace dm {

constexpr int kMaxWin = 5;  // the reference's scratch arrays hold 32 values: k*k <= 25

__device__ __forceinline__ void insertion_sort(float *v, int n) {
  for (int a = 1; a < n; ++a) {
    const float x = v[a];
    int b = a - 1;
    while (b >= 0 && v[b] > x) {
      v[b + 1] = v[b];
      --b;
    }
    v[b + 1] = x;
  }
}

// fmed (opticalflow_model.lua:388-434)
__global__ void pp_median_kernel(const float *flow, const float *mask, int h, int w, int k, float *ret) {
  const long long plane = (long long)h * w;
  const int hh = h - k, ww = w - k, halfk = k / 2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < (long long)hh * ww;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / ww), j = (int)(t % ww);
    float ty[kMaxWin * kMaxWin], tx[kMaxWin * kMaxWin];
    int n = 0;
    for (int a = i; a < i + k; ++a)
      for (int b = j; b < j + k; ++b)
        if (mask[(long long)a * w + b] != 0.0f) {
          ty[n] = flow[(long long)a * w + b];
          tx[n] = flow[plane + (long long)a * w + b];
          ++n;
        }
    float my = 0.0f, mx = 0.0f;  // an empty window reads the zero-filled scratch
    if (n > 0) {
      insertion_sort(ty, n);
      insertion_sort(tx, n);
      my = ty[n / 2];
      mx = tx[n / 2];
    }
    ret[(long long)(i + halfk) * w + (j + halfk)] = my;
    ret[plane + (long long)(i + halfk) * w + (j + halfk)] = mx;
  }
}

// global minimum of floor(x + 0.5) (opticalflow_model.lua:436-437)
__device__ __forceinline__ int float_to_ordered(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void round_min_kernel(const float *in, long long n, int *gmin) {
  int m = 0x7fffffff;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n;
       t += (long long)gridDim.x * blockDim.x)
    m = min(m, float_to_ordered(floorf(__fadd_rn(in[t], 0.5f))));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMin(gmin, m);
}

// fmax (opticalflow_model.lua:342-386) on round(input) - m, result + m everywhere (:438-439)
__global__ void pp_mode_kernel(const float *in, const float *mask, int h, int w, int k, const int *gmin,
                               float *out) {
  const long long plane = (long long)h * w;
  const float m = ordered_to_float(*gmin);
  const int hh = h - k, ww = w - k, halfk = k / 2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < plane;
       t += (long long)gridDim.x * blockDim.x) {
    const int Y = (int)(t / w), X = (int)(t % w);
    const int i = Y - halfk, j = X - halfk;
    float oy = 0.0f, ox = 0.0f;
    if (i >= 0 && i < hh && j >= 0 && j < ww) {
      int v[kMaxWin * kMaxWin];
      int n = 0;
      for (int a = i; a < i + k; ++a)
        for (int b = j; b < j + k; ++b)
          if (mask[(long long)a * w + b] != 0.0f) {
            const int vy = (int)__fsub_rn(floorf(__fadd_rn(in[(long long)a * w + b], 0.5f)), m);
            const int vx = (int)__fsub_rn(floorf(__fadd_rn(in[plane + (long long)a * w + b], 0.5f)), m);
            const int c = vx + 16 * vy;
            if (c >= 0 && c < 256) v[n++] = c;
          }
      // mode: the histogram scan keeps the first (smallest) index among equal counts
      int best = 0, bestc = 0;
      for (int a = 0; a < n; ++a) {
        int c = 0;
        for (int b = 0; b < n; ++b) c += v[b] == v[a];
        if (c > bestc || (c == bestc && v[a] < best)) {
          bestc = c;
          best = v[a];
        }
      }
      oy = (float)(best / 16);
      ox = (float)(best % 16);
    }
    out[t] = __fadd_rn(oy, m);
    out[plane + t] = __fadd_rn(ox, m);
  }
}

// enlargeMask: rows first, then columns on the result (depth_estimation_api.lua:92-125)
__global__ void enlarge_rows_kernel(float *mask, int h, int w, int ix) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < h; i += gridDim.x * blockDim.x) {
    float *row = mask + (long long)i * w;
    for (int j = 0; j < w; ++j)
      if (row[j] > 0.5f) {
        for (int k = j; k < min(j + ix, w); ++k) row[k] = 0.0f;
        break;
      }
    for (int j = w - 1; j >= 0; --j)
      if (row[j] > 0.5f) {
        for (int k = j; k >= max(j - ix + 1, 0); --k) row[k] = 0.0f;
        break;
      }
  }
}
__global__ void enlarge_cols_kernel(float *mask, int h, int w, int iy) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < w; j += gridDim.x * blockDim.x) {
    for (int i = 0; i < h; ++i)
      if (mask[(long long)i * w + j] > 0.5f) {
        for (int k = i; k < min(i + iy, h); ++k) mask[(long long)k * w + j] = 0.0f;
        break;
      }
    for (int i = h - 1; i >= 0; --i)
      if (mask[(long long)i * w + j] > 0.5f) {
        for (int k = i; k >= max(i - iy + 1, 0); --k) mask[(long long)k * w + j] = 0.0f;
        break;
      }
  }
}

// radial() (test_opticalflow.lua:165-187), including its `px*dx+dy*dy`
__global__ void radial_depth_kernel(const float *flow, int h, int w, float mh, float mw, float infty,
                                    float *ret, float *conf) {
  const long long plane = (long long)h * w;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < plane;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / w), j = (int)(t % w);
    const float py = __fsub_rn((float)i, mh), px = __fsub_rn((float)j, mw);
    const float pn = (float)sqrt((double)__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)));
    const float dy = flow[t], dx = flow[plane + t];
    const float dn = (float)sqrt((double)__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    float r = 0.0f, c = 0.0f;
    if (dn >= 0.2f) {
      r = fminf(__fdiv_rn(pn, dn), infty);
      if (__fadd_rn(__fmul_rn(px, dx), __fmul_rn(dy, dy)) > 0.125f) c = 1.0f;
    } else {
      c = 1.0f;
      r = infty;
    }
    ret[t] = r;
    conf[t] = c;
  }
}

// computeDepthMapFromFlow (ardrone/ardrone_api.cpp:99-140)
__global__ void depth_from_xflow_kernel(const float *xflow, const float *mask, int h, int w, float m,
                                        float *depth, float *conf) {
  const long long plane = (long long)h * w;
  const int middlex = w / 2, k = 3;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < plane;
       t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t / w), i = (int)(t % w);  // (row j, column i) like the reference
    float d = 0.0f, c = 0.0f;
    if (mask[t] != 0.0f) {
      int values[20];
      for (int v = 0; v < 20; ++v) values[v] = 0;
      for (int i2 = max(0, i - k); i2 < min(w, i + k); ++i2)
        for (int j2 = max(0, j - k); j2 < min(h, j + k); ++j2)
          if (mask[(long long)j2 * w + i2] != 0.0f) {
            const int f = (int)round((double)xflow[(long long)j2 * w + i2]);
            if (f + 8 >= 0 && f + 8 < 20) ++values[f + 8];
          }
      int mx = 0, im = 0;
      for (int v = 0; v < 20; ++v)
        if (values[v] > mx) {
          mx = values[v];
          im = v - 8;
        }
      if (mask[t] > 0.5f && i - middlex != 0) {
        const float a = fabsf((float)im);
        d = a < 1.1f ? 100.0f : __fdiv_rn(__fmul_rn(m, (float)abs(i - middlex)), a);
        c = 1.0f;
      }
    }
    depth[t] = d;
    conf[t] = c;
  }
}

static int blocks_for(dm_ctx *ctx, long long n, int bs = 256) {
  long long b = (n + bs - 1) / bs;
  const long long cap = (long long)ctx->num_sms * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace dm

using namespace dm;

extern "C" {

int dm_post_process_image(dm_ctx *ctx, const float *input, const float *mask, int h, int w, int winsize,
                          int method_max, float *output) {
  DM_REQUIRE(ctx && input && mask && output, "dm_post_process_image: NULL argument");
  DM_REQUIRE(h >= 1 && w >= 1 && winsize >= 1 && winsize <= kMaxWin,
             "dm_post_process_image: window %d outside 1..%d (the reference's scratch holds 32 values)",
             winsize, kMaxWin);
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  const size_t plane = (size_t)h * w;
  const void *din, *dmask;
  void *dout;
  DM_CHECK(call.in(input, 2 * plane * 4, &din));
  DM_CHECK(call.in(mask, plane * 4, &dmask));
  DM_CHECK(call.out(output, 2 * plane * 4, &dout));
  if (method_max) {
    void *gmin;
    DM_CHECK(call.alloc(&gmin, 4));
    DM_CUDA(cudaMemsetAsync(gmin, 0x7f, 4, ctx->stream));  // 0x7f7f7f7f: above any finite input
    round_min_kernel<<<blocks_for(ctx, 2 * (long long)plane), 256, 0, ctx->stream>>>(
        static_cast<const float *>(din), 2 * (long long)plane, static_cast<int *>(gmin));
    pp_mode_kernel<<<blocks_for(ctx, plane), 256, 0, ctx->stream>>>(
        static_cast<const float *>(din), static_cast<const float *>(dmask), h, w, winsize,
        static_cast<const int *>(gmin), static_cast<float *>(dout));
    count_launch(ctx, 2);
  } else {
    DM_CUDA(cudaMemsetAsync(dout, 0, 2 * plane * 4, ctx->stream));
    if (h > winsize && w > winsize) {
      pp_median_kernel<<<blocks_for(ctx, (long long)(h - winsize) * (w - winsize)), 256, 0, ctx->stream>>>(
          static_cast<const float *>(din), static_cast<const float *>(dmask), h, w, winsize,
          static_cast<float *>(dout));
      count_launch(ctx);
    }
  }
  return call.finish();
}

int dm_enlarge_mask(dm_ctx *ctx, float *mask, int h, int w, int ix, int iy) {
  DM_REQUIRE(ctx && mask, "dm_enlarge_mask: NULL argument");
  DM_REQUIRE(h >= 1 && w >= 1, "dm_enlarge_mask: bad shape");
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  void *d;
  DM_CHECK(call.out(mask, (size_t)h * w * 4, &d, true));
  enlarge_rows_kernel<<<(h + 127) / 128, 128, 0, ctx->stream>>>(static_cast<float *>(d), h, w, ix);
  enlarge_cols_kernel<<<(w + 127) / 128, 128, 0, ctx->stream>>>(static_cast<float *>(d), h, w, iy);
  count_launch(ctx, 2);
  return call.finish();
}

int dm_radial_depth(dm_ctx *ctx, const float *flow, int h, int w, float mh, float mw, float infty,
                    float *ret, float *conf) {
  DM_REQUIRE(ctx && flow && ret && conf, "dm_radial_depth: NULL argument");
  DM_REQUIRE(h >= 1 && w >= 1, "dm_radial_depth: bad shape");
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  const size_t plane = (size_t)h * w;
  const void *df;
  void *dr, *dc;
  DM_CHECK(call.in(flow, 2 * plane * 4, &df));
  DM_CHECK(call.out(ret, plane * 4, &dr));
  DM_CHECK(call.out(conf, plane * 4, &dc));
  radial_depth_kernel<<<blocks_for(ctx, plane), 256, 0, ctx->stream>>>(
      static_cast<const float *>(df), h, w, mh, mw, infty, static_cast<float *>(dr), static_cast<float *>(dc));
  count_launch(ctx);
  return call.finish();
}

int dm_depth_from_xflow(dm_ctx *ctx, const float *xflow, const float *mask, int h, int w, float m,
                        float *depth, float *conf) {
  DM_REQUIRE(ctx && xflow && mask && depth && conf, "dm_depth_from_xflow: NULL argument");
  DM_REQUIRE(h >= 1 && w >= 1, "dm_depth_from_xflow: bad shape");
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  const size_t plane = (size_t)h * w;
  const void *dx, *dm_;
  void *dd, *dc;
  DM_CHECK(call.in(xflow, plane * 4, &dx));
  DM_CHECK(call.in(mask, plane * 4, &dm_));
  DM_CHECK(call.out(depth, plane * 4, &dd));
  DM_CHECK(call.out(conf, plane * 4, &dc));
  depth_from_xflow_kernel<<<blocks_for(ctx, plane), 256, 0, ctx->stream>>>(
      static_cast<const float *>(dx), static_cast<const float *>(dm_), h, w, m, static_cast<float *>(dd),
      static_cast<float *>(dc));
  count_launch(ctx);
  return call.finish();
}

}  // extern "C"
