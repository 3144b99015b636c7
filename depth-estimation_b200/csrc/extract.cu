// extract.cu -- K2 as stand-alone, module-level operators over a volume that
// already sits in memory: Minus+SoftMax, argmax with the zero-flow tie rule,
// extractOutput / extractOutputMarginalized, the OutputExtractor soft mean, the
// x-marginal and the flow canvas.  All are streaming kernels (HBM bound, one warp
// per pixel, lanes over window entries so every load is a coalesced 128-byte row).
#include "dm_common.cuh"

namespace dm {

constexpr int kRowWarps = 8;

__device__ __forceinline__ long long warp_row(long long rows) {
  (void)rows;
  return (long long)blockIdx.x * kRowWarps + __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
}

// nn.Minus + nn.SoftMax (opticalflow_model.lua:94-109)
__global__ void __launch_bounds__(kRowWarps * 32) neg_softmax_kernel(const float *vol, long long rows,
                                                                     int K, float *out) {
  const int lane = threadIdx.x & 31;
  for (long long r = warp_row(rows); r < rows; r += (long long)gridDim.x * kRowWarps) {
    const float *v = vol + r * K;
    float vmin = __int_as_float(0x7f800000);
    for (int k = lane; k < K; k += 32) vmin = fminf(vmin, v[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    double sum = 0.0;
    for (int k = lane; k < K; k += 32) sum += (double)expf(vmin - v[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const double inv = 1.0 / sum;
    float *o_ = out + r * K;
    for (int k = lane; k < K; k += 32) o_[k] = (float)((double)expf(vmin - v[k]) * inv);
  }
}

// getOutputConfidences, threshold == nil (opticalflow_model.lua:153-161) and the
// argmin variant of the ground-truth generators
__global__ void __launch_bounds__(kRowWarps * 32)
argmax_tie_kernel(const float *vol, long long rows, int K, int middle, int take_min, long long *index,
                  float *value) {
  const int lane = threadIdx.x & 31;
  const float sgn = take_min ? -1.0f : 1.0f;
  for (long long r = warp_row(rows); r < rows; r += (long long)gridDim.x * kRowWarps) {
    const float *v = vol + r * K;
    float best = -__int_as_float(0x7f800000);
    int kb = 0x7fffffff;
    for (int k = lane; k < K; k += 32) {
      const float x = sgn * v[k];
      if (x > best) {
        best = x;
        kb = k;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int ok = __shfl_xor_sync(0xffffffffu, kb, o);
      if (ob > best || (ob == best && ok < kb)) {
        best = ob;
        kb = ok;
      }
    }
    if (lane == 0) {
      long long win = kb + 1;
      if (middle > 0 && sgn * v[middle - 1] == best) win = middle;
      index[r] = win;
      if (value) value[r] = sgn * best;
    }
  }
}

__device__ __constant__ unsigned char eNet4[5][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}, {1, 2}};
__device__ __constant__ unsigned char eNet8[19][2] = {
    {0, 1}, {2, 3}, {4, 5}, {6, 7}, {0, 2}, {1, 3}, {4, 6}, {5, 7}, {1, 2}, {5, 6},
    {0, 4}, {3, 7}, {1, 5}, {2, 6}, {1, 4}, {3, 6}, {2, 4}, {3, 5}, {3, 4}};

// extractOutput / extractOutputMarginalized (extract_output.cpp:63-155, :157-255)
__global__ void __launch_bounds__(kRowWarps * 32)
extract_output_kernel(const float *input, long long rows, int n, double threshold, int M,
                      int marginalized, double threshold_acc, long long *ret, float *scores,
                      long long *retgd, unsigned long long *n_untouched) {
  const int lane = threadIdx.x & 31;
  for (long long r = warp_row(rows); r < rows; r += (long long)gridDim.x * kRowWarps) {
    const float *v = input + r * n;
    float val[8], pos[8];
    for (int j = 0; j < 8; ++j) val[j] = pos[j] = 0.0f;
    int got = 0;
    for (int k0 = 0; k0 < n && got < M; k0 += 32) {
      const int k = k0 + lane;
      const float x = k < n ? v[k] : 0.0f;
      unsigned hit = __ballot_sync(0xffffffffu, k < n && (double)x > threshold);
      while (hit && got < M) {
        const int src = __ffs(hit) - 1;
        hit &= hit - 1;
        val[got] = __shfl_sync(0xffffffffu, x, src);
        pos[got] = (float)(k0 + src + 1);
        ++got;
      }
    }
    if (lane != 0) continue;
    if (marginalized) retgd[r] = 0;
    if (!(val[0] > 0.0f)) {  // extract_output.cpp:120 -- untouched
      if (n_untouched) atomicAdd(n_untouched, 1ull);
      continue;
    }
    const int nex = M == 4 ? 5 : 19;
    for (int e = 0; e < nex; ++e) {
      const int a = M == 4 ? eNet4[e][0] : eNet8[e][0];
      const int b = M == 4 ? eNet4[e][1] : eNet8[e][1];
      if (val[b] > val[a]) {
        float t = val[a]; val[a] = val[b]; val[b] = t;
        t = pos[a]; pos[a] = pos[b]; pos[b] = t;
      }
    }
    ret[r] = (long long)pos[0];
    for (int k = 1; k < M; ++k) val[k] = __fadd_rn(val[k], val[k - 1]);
    double acc = 0.0;
    for (int k = 0; k < M; ++k) acc += (double)val[k];
    if (marginalized) {
      if (acc >= threshold_acc) retgd[r] = 1;
    } else {
      scores[r] = (float)acc;
    }
  }
}

// nn.OutputExtractor (OutputExtractor.lua:21-35)
__global__ void __launch_bounds__(kRowWarps * 32)
soft_mean_kernel(const float *prob, long long rows, int maxh, int maxw, float *ymean, float *xmean) {
  const int lane = threadIdx.x & 31;
  const int K = maxh * maxw;
  for (long long r = warp_row(rows); r < rows; r += (long long)gridDim.x * kRowWarps) {
    const float *p = prob + r * K;
    double sx = 0.0, sy = 0.0;
    for (int k = lane; k < K; k += 32) {
      const int i = k / maxw, j = k - i * maxw;
      const float v = p[k];
      sx += (double)(v * (float)(j + 1));
      sy += (double)(v * (float)(i + 1));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
    }
    if (lane == 0) {
      xmean[r] = (float)sx;
      ymean[r] = (float)sy;
    }
  }
}

// input:reshape(h,w,maxh,maxw):sum(4) (opticalflow_model.lua:191)
__global__ void marginal_x_kernel(const float *prob, long long total, int maxw, float *pm) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const float *p = prob + t * maxw;
    double s = 0.0;
    for (int j = 0; j < maxw; ++j) s += (double)p[j];
    pm[t] = (float)s;
  }
}

// x2yx + centre + canvas (opticalflow_model.lua:16-25,208-212,227-250)
__global__ void flow_canvas_kernel(const long long *index, int h1, int w1, int maxh, int maxw,
                                   int h_img, int w_img, float *full) {
  const int cy = (maxh + 1) / 2, cx = (maxw + 1) / 2;
  const int hoff = (h_img - h1) / 2, woff = (w_img - w1) / 2;
  const long long plane = (long long)h_img * w_img;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < plane;
       t += (long long)gridDim.x * blockDim.x) {
    const int Y = (int)(t / w_img), X = (int)(t % w_img);
    const int y = Y - hoff, x = X - woff;
    float fy = 0.0f, fx = 0.0f;
    if (y >= 0 && y < h1 && x >= 0 && x < w1) {
      const long long k = index[(long long)y * w1 + x] - 1;
      const long long row = k >= 0 ? k / maxw : -((-k + maxw - 1) / maxw);  // floor
      const long long col = k - row * maxw;
      fy = (float)(row + 1 - cy);
      fx = (float)(col + 1 - cx);
    }
    full[t] = fy;
    full[plane + t] = fx;
  }
}

static int row_grid(dm_ctx *ctx, long long rows) {
  long long b = (rows + kRowWarps - 1) / kRowWarps;
  const long long cap = (long long)ctx->num_sms * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
static int flat_grid(dm_ctx *ctx, long long n, int block) {
  long long b = (n + block - 1) / block;
  const long long cap = (long long)ctx->num_sms * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace dm

using namespace dm;

extern "C" {

int dm_neg_softmax(dm_ctx *ctx, const float *vol, int64_t rows, int k, float *out) {
  DM_REQUIRE(ctx && vol && out, "dm_neg_softmax: NULL argument");
  DM_REQUIRE(rows >= 0 && k >= 1, "dm_neg_softmax: bad shape %lld x %d", (long long)rows, k);
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  if (rows == 0) return call.finish();
  const size_t bytes = (size_t)rows * k * sizeof(float);
  const void *din;
  void *dout;
  DM_CHECK(call.in(vol, bytes, &din));
  if (out == vol && classify(vol) == PtrKind::Device) {
    dout = out;
  } else {
    DM_CHECK(call.out(out, bytes, &dout));
  }
  neg_softmax_kernel<<<row_grid(ctx, rows), kRowWarps * 32, 0, ctx->stream>>>(
      static_cast<const float *>(din), rows, k, static_cast<float *>(dout));
  count_launch(ctx);
  return call.finish();
}

int dm_argmax_tie(dm_ctx *ctx, const float *vol, int64_t rows, int k, int middle, int take_min,
                  int64_t *index, float *value) {
  DM_REQUIRE(ctx && vol && index, "dm_argmax_tie: NULL argument");
  DM_REQUIRE(rows >= 0 && k >= 1 && middle <= k, "dm_argmax_tie: bad shape/middle");
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  if (rows == 0) return call.finish();
  const void *din;
  void *didx, *dval = nullptr;
  DM_CHECK(call.in(vol, (size_t)rows * k * sizeof(float), &din));
  DM_CHECK(call.out(index, (size_t)rows * 8, &didx));
  if (value) DM_CHECK(call.out(value, (size_t)rows * 4, &dval));
  argmax_tie_kernel<<<row_grid(ctx, rows), kRowWarps * 32, 0, ctx->stream>>>(
      static_cast<const float *>(din), rows, k, middle, take_min, static_cast<long long *>(didx),
      static_cast<float *>(dval));
  count_launch(ctx);
  return call.finish();
}

static int extract_common(dm_ctx *ctx, const float *input, int h, int w, int n, double threshold,
                          int marginalized, double threshold_acc, int64_t *ret, float *scores,
                          int64_t *retgd, int64_t *n_untouched) {
  DM_REQUIRE(h >= 0 && w >= 0 && n >= 1, "extractOutput: bad shape %dx%dx%d", h, w, n);
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  const long long rows = (long long)h * w;
  if (rows == 0) {
    if (n_untouched) *n_untouched = 0;
    return call.finish();
  }
  const void *din;
  void *dret, *dsc = nullptr, *dgd = nullptr, *dcnt = nullptr;
  DM_CHECK(call.in(input, (size_t)rows * n * sizeof(float), &din));
  // untouched pixels keep the caller's values: preload host outputs
  DM_CHECK(call.out(ret, (size_t)rows * 8, &dret, true));
  if (scores) DM_CHECK(call.out(scores, (size_t)rows * 4, &dsc, true));
  if (retgd) DM_CHECK(call.out(retgd, (size_t)rows * 8, &dgd, false));
  if (n_untouched) {
    DM_CHECK(call.alloc(&dcnt, 8));
    DM_CUDA(cudaMemsetAsync(dcnt, 0, 8, ctx->stream));
  }
  const int M = threshold < 0.2 ? 8 : 4;
  extract_output_kernel<<<row_grid(ctx, rows), kRowWarps * 32, 0, ctx->stream>>>(
      static_cast<const float *>(din), rows, n, threshold, M, marginalized, threshold_acc,
      static_cast<long long *>(dret), static_cast<float *>(dsc), static_cast<long long *>(dgd),
      static_cast<unsigned long long *>(dcnt));
  count_launch(ctx);
  if (n_untouched) {
    DM_CUDA(cudaMemcpyAsync(n_untouched, dcnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->call_has_host = true;
  }
  return call.finish();
}

int dm_extract_output(dm_ctx *ctx, const float *input, int h, int w, int n, double threshold,
                      int64_t *ret, float *scores, int64_t *n_untouched) {
  DM_REQUIRE(ctx && input && ret && scores, "dm_extract_output: NULL argument");
  return extract_common(ctx, input, h, w, n, threshold, 0, 0.0, ret, scores, nullptr, n_untouched);
}

int dm_extract_output_marginalized(dm_ctx *ctx, const float *input, int h, int w, int n,
                                   double threshold, double threshold_acc, int64_t *ret,
                                   int64_t *retgd) {
  DM_REQUIRE(ctx && input && ret && retgd, "dm_extract_output_marginalized: NULL argument");
  return extract_common(ctx, input, h, w, n, threshold, 1, threshold_acc, ret, nullptr, retgd,
                        nullptr);
}

int dm_soft_mean(dm_ctx *ctx, const float *prob, int64_t rows, int maxh, int maxw, float *ymean,
                 float *xmean) {
  DM_REQUIRE(ctx && prob && ymean && xmean, "dm_soft_mean: NULL argument");
  DM_REQUIRE(rows >= 0 && maxh >= 1 && maxw >= 1, "dm_soft_mean: bad shape");
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  if (rows == 0) return call.finish();
  const void *din;
  void *dy, *dx;
  DM_CHECK(call.in(prob, (size_t)rows * maxh * maxw * sizeof(float), &din));
  DM_CHECK(call.out(ymean, (size_t)rows * 4, &dy));
  DM_CHECK(call.out(xmean, (size_t)rows * 4, &dx));
  soft_mean_kernel<<<row_grid(ctx, rows), kRowWarps * 32, 0, ctx->stream>>>(
      static_cast<const float *>(din), rows, maxh, maxw, static_cast<float *>(dy),
      static_cast<float *>(dx));
  count_launch(ctx);
  return call.finish();
}

int dm_marginal_x(dm_ctx *ctx, const float *prob, int64_t rows, int maxh, int maxw, float *pm) {
  DM_REQUIRE(ctx && prob && pm, "dm_marginal_x: NULL argument");
  DM_REQUIRE(rows >= 0 && maxh >= 1 && maxw >= 1, "dm_marginal_x: bad shape");
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  if (rows == 0) return call.finish();
  const void *din;
  void *dpm;
  DM_CHECK(call.in(prob, (size_t)rows * maxh * maxw * sizeof(float), &din));
  DM_CHECK(call.out(pm, (size_t)rows * maxh * 4, &dpm));
  const long long total = (long long)rows * maxh;
  marginal_x_kernel<<<flat_grid(ctx, total, 256), 256, 0, ctx->stream>>>(
      static_cast<const float *>(din), total, maxw, static_cast<float *>(dpm));
  count_launch(ctx);
  return call.finish();
}

int dm_flow_canvas(dm_ctx *ctx, const int64_t *index, int h1, int w1, int maxh, int maxw,
                   int h_img, int w_img, float *full) {
  DM_REQUIRE(ctx && index && full, "dm_flow_canvas: NULL argument");
  DM_REQUIRE(h1 >= 1 && w1 >= 1 && h_img >= h1 && w_img >= w1 && maxh >= 1 && maxw >= 1,
             "dm_flow_canvas: bad shape");
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  const void *din;
  void *dfull;
  DM_CHECK(call.in(index, (size_t)h1 * w1 * 8, &din));
  DM_CHECK(call.out(full, (size_t)2 * h_img * w_img * 4, &dfull));
  const long long plane = (long long)h_img * w_img;
  flow_canvas_kernel<<<flat_grid(ctx, plane, 256), 256, 0, ctx->stream>>>(
      static_cast<const long long *>(din), h1, w1, maxh, maxw, h_img, w_img,
      static_cast<float *>(dfull));
  count_launch(ctx);
  return call.finish();
}

}  // extern "C"
