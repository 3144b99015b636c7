// match_generic.cu -- untiled matching kernels: one warp per output pixel, lanes
// over window entries, frame 2 read through L2.  Used for shapes the tiled TMA
// kernels do not cover (more than 16 channels, e.g. the ground-truth generators'
// unfolded patches, groundtruth_opticalflow.lua:53-74) and for the radial search
// (window width 1, radial/radial_opticalflow_network.lua:32-34), where there is no
// dx reuse to tile for.  Three passes per pixel (min, sum, normalised outputs)
// mirror the reference's Minus -> SoftMax -> max/extract order literally.
#include "dm_common.cuh"

namespace dm {

constexpr int kGWarps = 4;


__device__ __forceinline__ float ssd_at(const GenericParams &P, const float *a, const float *b,
                                        bool exact) {
  float acc = 0.0f;
  for (int c = 0; c < P.C; ++c) {
    const float d = __ldg(a + c * P.s1c) - __ldg(b + c * P.s2c);
    acc = exact ? __fadd_rn(acc, __fmul_rn(d, d)) : fmaf(d, d, acc);
  }
  return acc;
}

__device__ __constant__ unsigned char gNet4[5][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}, {1, 2}};
__device__ __constant__ unsigned char gNet8[19][2] = {
    {0, 1}, {2, 3}, {4, 5}, {6, 7}, {0, 2}, {1, 3}, {4, 6}, {5, 7}, {1, 2}, {5, 6},
    {0, 4}, {3, 7}, {1, 5}, {2, 6}, {1, 4}, {3, 6}, {2, 4}, {3, 5}, {3, 4}};

template <bool VOLUME>
__global__ void __launch_bounds__(kGWarps * 32) generic_kernel(const GenericParams P) {
  const int lane = threadIdx.x & 31;
  const long long npx = (long long)P.N * P.H1 * P.W1;
  const int K = P.maxh * P.maxw;
  const bool exact = P.flags & DM_FLAG_EXACT_SSD;
  const long long nwork = P.list ? (long long)*P.nlist : npx;
  for (long long it = (long long)blockIdx.x * kGWarps + __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0); it < nwork;
       it += (long long)gridDim.x * kGWarps) {
    const long long px = P.list ? (long long)P.list[it] : it;
    const int x = (int)(px % P.W1);
    const int y = (int)((px / P.W1) % P.H1);
    const int n = (int)(px / ((long long)P.W1 * P.H1));
    const float *a = P.in1 + n * P.s1n + y * P.s1y + x;
    const float *b0 = P.in2 + n * P.s2n + y * P.s2y + x;
    // pass 1: minimum, first occurrence
    float vbest = __int_as_float(0x7f800000);
    int kbest = 0x7fffffff;
    for (int k = lane; k < K; k += 32) {
      const int dy = k / P.maxw, dx = k - dy * P.maxw;
      const float v = ssd_at(P, a, b0 + dy * P.s2y + dx, exact);
      if (VOLUME && P.mode == DM_VOLUME_SSD) P.vol[px * K + k] = v;
      if (v < vbest) {
        vbest = v;
        kbest = k;
      }
    }
    if (VOLUME && P.mode == DM_VOLUME_SSD) continue;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, vbest, o);
      const int ok = __shfl_xor_sync(0xffffffffu, kbest, o);
      if (ov < vbest || (ov == vbest && ok < kbest)) {
        vbest = ov;
        kbest = ok;
      }
    }
    // pass 2: sum of exp(min - v) in double (TH sums in accreal)
    double sum = 0.0;
    for (int k = lane; k < K; k += 32) {
      const int dy = k / P.maxw, dx = k - dy * P.maxw;
      const float v = ssd_at(P, a, b0 + dy * P.s2y + dx, exact);
      sum += (double)expf(vbest - v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const double inv = 1.0 / sum;
    const float pbest = (float)(1.0 * inv);
    // pass 3: normalised probabilities -> soft mean, thresholded list, tie rule, volume
    double sx = 0.0, sy = 0.0;
    float pmid = -1.0f;
    float cval[8], cpos[8];
    int got = 0;
    for (int j = 0; j < 8; ++j) cval[j] = cpos[j] = 0.0f;
    const bool want_thr = !VOLUME && (P.index_thr || P.score_thr || P.n_untouched);
    for (int k0 = 0; k0 < K; k0 += 32) {
      const int k = k0 + lane;
      float pk = 0.0f;
      if (k < K) {
        const int dy = k / P.maxw, dx = k - dy * P.maxw;
        const float v = ssd_at(P, a, b0 + dy * P.s2y + dx, exact);
        pk = (float)((double)expf(vbest - v) * inv);
        if (VOLUME) P.vol[px * K + k] = pk;
        sx += (double)(pk * (float)(dx + 1));
        sy += (double)(pk * (float)(dy + 1));
        if (k + 1 == P.middle) pmid = pk;
      }
      if (want_thr) {
        unsigned hit = __ballot_sync(0xffffffffu, k < K && (double)pk > P.thr);
        while (hit && got < P.M) {  // scan order: lower lanes are earlier entries
          const int src = __ffs(hit) - 1;
          hit &= hit - 1;
          cval[got] = __shfl_sync(0xffffffffu, pk, src);
          cpos[got] = (float)(k0 + src + 1);
          ++got;
        }
      }
    }
    if (VOLUME) continue;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      pmid = fmaxf(pmid, __shfl_xor_sync(0xffffffffu, pmid, o));
    }
    float cmarg = 0.0f;
    if (!VOLUME && P.conf_marginal) {
      // opticalflow_model.lua:191-196: marginal over dx (TH sum: double accumulate, float result),
      // then extractOutput(pm, thr) > 0, i.e. some row's marginal exceeds the threshold
      for (int dy = 0; dy < P.maxh; ++dy) {
        double rs = 0.0;
        for (int dx = lane; dx < P.maxw; dx += 32) {
          const float v = ssd_at(P, a, b0 + dy * P.s2y + dx, exact);
          rs += (double)(float)((double)expf(vbest - v) * inv);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
        if ((double)(float)rs > P.thr) cmarg = 1.0f;
      }
    }
    if (lane != 0) continue;
    if (!VOLUME && P.conf_marginal) P.conf_marginal[px] = cmarg;
    int win = kbest + 1;
    if ((P.flags & DM_FLAG_TIE_MIDDLE) && pmid == pbest) win = P.middle;
    if (P.index) P.index[px] = win;
    if (P.min_ssd) P.min_ssd[px] = vbest;
    if (P.pmax) P.pmax[px] = pbest;
    if (P.radial_flow) P.radial_flow[px] = (float)kbest;
    if (P.soft_yx) {
      const size_t plane = (size_t)P.H1 * P.W1;
      const size_t so = (size_t)n * 2 * plane + (size_t)y * P.W1 + x;
      P.soft_yx[so] = (float)sy;
      P.soft_yx[so + plane] = (float)sx;
    }
    if (P.flow_full) {
      const int row = (win - 1) / P.maxw + 1, col = (win - 1) % P.maxw + 1;
      const size_t plane = (size_t)P.h_img * P.w_img;
      const size_t fo = (size_t)n * 2 * plane + (size_t)(y + P.hoff) * P.w_img + (x + P.woff);
      P.flow_full[fo] = (float)(row - P.cy);
      P.flow_full[fo + plane] = (float)(col - P.cx);
    }
    if (want_thr) {
      long long ret = 0;
      float score = 0.0f;
      if (got > 0) {
        const int nex = P.M == 4 ? 5 : 19;
        for (int e = 0; e < nex; ++e) {
          const int ia = P.M == 4 ? gNet4[e][0] : gNet8[e][0];
          const int ib = P.M == 4 ? gNet4[e][1] : gNet8[e][1];
          if (cval[ib] > cval[ia]) {
            float t = cval[ia]; cval[ia] = cval[ib]; cval[ib] = t;
            t = cpos[ia]; cpos[ia] = cpos[ib]; cpos[ib] = t;
          }
        }
        ret = (long long)cpos[0];
        for (int k = 1; k < P.M; ++k) cval[k] = __fadd_rn(cval[k], cval[k - 1]);
        double acc = 0.0;
        for (int k = 0; k < P.M; ++k) acc += (double)cval[k];
        score = (float)acc;
      } else if (P.n_untouched) {
        atomicAdd(P.n_untouched + n, 1ull);
      }
      if (P.index_thr) P.index_thr[px] = ret;
      if (P.score_thr) P.score_thr[px] = score;
    }
  }
}

static int fill_inputs(Call &call, const dm_pair *in, int maxh, int maxw, GenericParams *P) {
  DM_REQUIRE(in && in->in1 && in->in2, "input pointers are NULL");
  DM_REQUIRE(in->n_pairs >= 1 && in->channels >= 1, "n_pairs and channels must be >= 1");
  DM_REQUIRE(maxh >= 1 && maxw >= 1, "window must be at least 1x1 (got %dx%d)", maxh, maxw);
  DM_REQUIRE(in->h1 >= 1 && in->w1 >= 1, "empty frame-1 map (%dx%d)", in->h1, in->w1);
  DM_REQUIRE(in->h2 >= in->h1 + maxh - 1 && in->w2 >= in->w1 + maxw - 1,
             "frame 2 (%dx%d) smaller than frame 1 (%dx%d) + window (%dx%d) - 1", in->h2, in->w2,
             in->h1, in->w1, maxh, maxw);
  memset(P, 0, sizeof(*P));
  P->N = in->n_pairs;
  P->C = in->channels;
  P->H1 = in->h1;
  P->W1 = in->w1;
  P->maxh = maxh;
  P->maxw = maxw;
  P->s1y = in->in1_stride_y ? in->in1_stride_y : in->w1;
  P->s1c = in->in1_stride_c ? in->in1_stride_c : (long long)in->h1 * P->s1y;
  P->s1n = in->in1_stride_n ? in->in1_stride_n : (long long)in->channels * P->s1c;
  P->s2y = in->in2_stride_y ? in->in2_stride_y : in->w2;
  P->s2c = in->in2_stride_c ? in->in2_stride_c : (long long)in->h2 * P->s2y;
  P->s2n = in->in2_stride_n ? in->in2_stride_n : (long long)in->channels * P->s2c;
  const size_t span1 =
      (size_t)((P->N - 1) * P->s1n + (P->C - 1) * P->s1c + (P->H1 - 1) * P->s1y + P->W1);
  const size_t span2 =
      (size_t)((P->N - 1) * P->s2n + (P->C - 1) * P->s2c + (in->h2 - 1) * P->s2y + in->w2);
  const void *p = nullptr;
  DM_CHECK(call.in(in->in1, span1 * sizeof(float), &p));
  P->in1 = static_cast<const float *>(p);
  DM_CHECK(call.in(in->in2, span2 * sizeof(float), &p));
  P->in2 = static_cast<const float *>(p);
  return DM_OK;
}

// extractOutput on the RAW SSD volume (radial/radial_opticalflow_groundtruth.lua:105, version2/groundtruth.lua:103:
// threshold 0.21 -> M = 4) without the volume: one warp per pixel walks the window in scan order, 32 entries at
// a time, SSD with separately rounded multiply and add (the CPU path's arithmetic), keeps the FIRST M values
// above the threshold with their 1-based positions (extract_output.cpp:96-114) and stops as soon as it has
// them -- with a threshold below the typical SSD that is the first chunk.  Sorting network, ret and score as in
// extract_output.cpp:17-61,123-129.  A pixel with no value above the threshold keeps ret = 0, score = 0 and is
// counted (the reference leaves the caller's buffers untouched there).
__global__ void __launch_bounds__(kGWarps * 32) raw_ssd_extract_kernel(const GenericParams P) {
  const int lane = threadIdx.x & 31;
  const long long npx = (long long)P.N * P.H1 * P.W1;
  const int K = P.maxh * P.maxw;
  for (long long px = (long long)blockIdx.x * kGWarps + __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0); px < npx;
       px += (long long)gridDim.x * kGWarps) {
    const int x = (int)(px % P.W1);
    const int y = (int)((px / P.W1) % P.H1);
    const int n = (int)(px / ((long long)P.W1 * P.H1));
    const float *a = P.in1 + n * P.s1n + y * P.s1y + x;
    const float *b0 = P.in2 + n * P.s2n + y * P.s2y + x;
    float cval[8], cpos[8];
    int got = 0;
    for (int j = 0; j < 8; ++j) cval[j] = cpos[j] = 0.0f;
    for (int k0 = 0; k0 < K && got < P.M; k0 += 32) {
      const int k = k0 + lane;
      float v = 0.0f;
      if (k < K) {
        const int dy = k / P.maxw, dx = k - dy * P.maxw;
        v = ssd_at(P, a, b0 + dy * P.s2y + dx, true);
      }
      unsigned hit = __ballot_sync(0xffffffffu, k < K && (double)v > P.thr);
      while (hit && got < P.M) {  // scan order: lower lanes are earlier entries
        const int src = __ffs(hit) - 1;
        hit &= hit - 1;
        cval[got] = __shfl_sync(0xffffffffu, v, src);
        cpos[got] = (float)(k0 + src + 1);
        ++got;
      }
    }
    if (lane != 0) continue;
    long long ret = 0;
    float score = 0.0f;
    if (got > 0) {
      const int nex = P.M == 4 ? 5 : 19;
      for (int e = 0; e < nex; ++e) {
        const int ia = P.M == 4 ? gNet4[e][0] : gNet8[e][0];
        const int ib = P.M == 4 ? gNet4[e][1] : gNet8[e][1];
        if (cval[ib] > cval[ia]) {
          float t = cval[ia]; cval[ia] = cval[ib]; cval[ib] = t;
          t = cpos[ia]; cpos[ia] = cpos[ib]; cpos[ib] = t;
        }
      }
      ret = (long long)cpos[0];
      for (int k = 1; k < P.M; ++k) cval[k] = __fadd_rn(cval[k], cval[k - 1]);
      double acc = 0.0;
      for (int k = 0; k < P.M; ++k) acc += (double)cval[k];
      score = (float)acc;
    } else if (P.n_untouched) {
      atomicAdd(P.n_untouched + n, 1ull);
    }
    P.index_thr[px] = ret;
    P.score_thr[px] = score;
  }
}

static int launch_generic(dm_ctx *ctx, const GenericParams &P, bool volume) {
  const long long npx = (long long)P.N * P.H1 * P.W1;
  long long blocks = (npx + kGWarps - 1) / kGWarps;
  const long long cap = (long long)ctx->num_sms * 16;
  if (blocks > cap) blocks = cap;
  if (volume)
    generic_kernel<true><<<(int)blocks, kGWarps * 32, 0, ctx->stream>>>(P);
  else
    generic_kernel<false><<<(int)blocks, kGWarps * 32, 0, ctx->stream>>>(P);
  DM_CUDA(cudaGetLastError());
  count_launch(ctx);
  return DM_OK;
}

// The pixels the tiled sweep handed over (ExtractParams::resc): every requested output of theirs,
// entry by entry in the reference's order of operations, SSD with separately rounded multiply
// and add like the CPU path.  The list length is only known on the device.
int generic_rescore(dm_ctx *ctx, GenericParams P, const int *list, const unsigned *nlist) {
  P.list = list;
  P.nlist = nlist;
  P.flags |= DM_FLAG_EXACT_SSD;
  generic_kernel<false><<<ctx->num_sms * 2, kGWarps * 32, 0, ctx->stream>>>(P);
  DM_CUDA(cudaGetLastError());
  count_launch(ctx);
  return DM_OK;
}

int generic_match_extract(Call &call, const dm_pair *in, int maxh, int maxw, unsigned flags,
                          double thr, int h_img, int w_img, const dm_extract_out *out) {
  dm_ctx *ctx = call.ctx;
  GenericParams P;
  DM_CHECK(fill_inputs(call, in, maxh, maxw, &P));
  const size_t npx = (size_t)P.N * P.H1 * P.W1;
  P.flags = flags;
  P.thr = thr;
  P.M = thr < 0.2 ? 8 : 4;
  P.cy = (maxh + 1) / 2;
  P.cx = (maxw + 1) / 2;
  P.middle = (P.cy - 1) * maxw + P.cx;
  P.h_img = h_img;
  P.w_img = w_img;
  P.hoff = (h_img - P.H1) / 2;
  P.woff = (w_img - P.W1) / 2;
  void *p;
#define DM_OUT(field, type, bytes)                 \
  if (out->field) {                                \
    DM_CHECK(call.out(out->field, (bytes), &p));   \
    P.field = static_cast<type *>(p);              \
  }
  DM_OUT(index, long long, npx * 8)
  DM_OUT(min_ssd, float, npx * 4)
  DM_OUT(pmax, float, npx * 4)
  DM_OUT(flow_full, float, (size_t)P.N * 2 * h_img * w_img * 4)
  DM_OUT(index_thr, long long, npx * 8)
  DM_OUT(score_thr, float, npx * 4)
  DM_OUT(soft_yx, float, npx * 2 * 4)
#undef DM_OUT
  if (out->n_untouched) {
    DM_CHECK(call.out(out->n_untouched, (size_t)P.N * 8, &p));
    P.n_untouched = static_cast<unsigned long long *>(p);
    DM_CUDA(cudaMemsetAsync(p, 0, (size_t)P.N * 8, ctx->stream));
  }
  if (P.flow_full)
    DM_CUDA(cudaMemsetAsync(P.flow_full, 0, (size_t)P.N * 2 * h_img * w_img * 4, ctx->stream));
  return launch_generic(ctx, P, false);
}

int generic_match_volume(Call &call, const dm_pair *in, int maxh, int maxw, int mode, bool exact,
                         float *out) {
  GenericParams P;
  DM_CHECK(fill_inputs(call, in, maxh, maxw, &P));
  P.flags = exact ? DM_FLAG_EXACT_SSD : 0u;
  P.mode = mode;
  P.thr = 2.0;
  P.M = 8;
  void *p = nullptr;
  DM_CHECK(call.out(out, (size_t)P.N * P.H1 * P.W1 * maxh * maxw * sizeof(float), &p));
  P.vol = static_cast<float *>(p);
  return launch_generic(call.ctx, P, true);
}

// nn.SpatialRadialMatching + `output:min(3)` (radial/radial_opticalflow_network.lua:32-34,
// radial/test_radial_opticalflow.lua:204-207) need no soft-max: one thread per pixel walks the
// hWin rows below it, every load coalesced along the polar angle; SSD with separately rounded
// multiply and add like the CPU path, strict < (first occurrence).
__global__ void radial_argmin_kernel(const GenericParams P) {
  const long long npx = (long long)P.N * P.H1 * P.W1;
  for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < npx;
       px += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(px % P.W1);
    const int y = (int)((px / P.W1) % P.H1);
    const int n = (int)(px / ((long long)P.W1 * P.H1));
    const float *a = P.in1 + n * P.s1n + y * P.s1y + x;
    const float *b0 = P.in2 + n * P.s2n + y * P.s2y + x;
    float vbest = __int_as_float(0x7f800000);
    int dbest = 0;
    for (int d = 0; d < P.maxh; ++d) {
      const float v = ssd_at(P, a, b0 + d * P.s2y, true);
      if (v < vbest) {
        vbest = v;
        dbest = d;
      }
    }
    P.radial_flow[px] = (float)dbest;
    if (P.min_ssd) P.min_ssd[px] = vbest;
  }
}

}  // namespace dm

using namespace dm;

extern "C" int dm_radial_match_extract(dm_ctx *ctx, const dm_pair *in, int h_win, float *flow,
                                       float *min_ssd) {
  DM_REQUIRE(ctx && in && flow, "dm_radial_match_extract: NULL argument");
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  GenericParams P;
  DM_CHECK(fill_inputs(call, in, h_win, 1, &P));
  P.flags = DM_FLAG_EXACT_SSD;  // the radial search is tiny: keep it bit-exact with the CPU path
  P.thr = 2.0;
  P.M = 8;
  const size_t npx = (size_t)P.N * P.H1 * P.W1;
  void *p = nullptr;
  DM_CHECK(call.out(flow, npx * sizeof(float), &p));
  P.radial_flow = static_cast<float *>(p);
  if (min_ssd) {
    DM_CHECK(call.out(min_ssd, npx * sizeof(float), &p));
    P.min_ssd = static_cast<float *>(p);
  }
  {
    long long blocks = ((long long)npx + 127) / 128;
    if (blocks > (long long)ctx->num_sms * 32) blocks = (long long)ctx->num_sms * 32;
    prof_begin(ctx);
    radial_argmin_kernel<<<(int)blocks, 128, 0, ctx->stream>>>(P);
    prof_end(ctx);
    count_launch(ctx);
  }
  return call.finish();
}

namespace dm {
static int raw_ssd_extract_on(Call &call, const dm_pair *in, int maxh, int maxw, double threshold, int64_t *ret,
                              float *scores, int64_t *n_untouched) {
  dm_ctx *ctx = call.ctx;
  GenericParams P;
  DM_CHECK(fill_inputs(call, in, maxh, maxw, &P));
  const size_t npx = (size_t)P.N * P.H1 * P.W1;
  P.thr = threshold;
  P.M = threshold < 0.2 ? 8 : 4;
  void *p = nullptr;
  DM_CHECK(call.out(ret, npx * sizeof(long long), &p));
  P.index_thr = static_cast<long long *>(p);
  DM_CHECK(call.out(scores, npx * sizeof(float), &p));
  P.score_thr = static_cast<float *>(p);
  if (n_untouched) {
    DM_CHECK(call.out(n_untouched, (size_t)P.N * 8, &p));
    P.n_untouched = static_cast<unsigned long long *>(p);
    DM_CUDA(cudaMemsetAsync(p, 0, (size_t)P.N * 8, ctx->stream));
  }
  long long blocks = ((long long)npx + kGWarps - 1) / kGWarps;
  const long long cap = (long long)ctx->num_sms * 16;
  if (blocks > cap) blocks = cap;
  raw_ssd_extract_kernel<<<(int)blocks, kGWarps * 32, 0, ctx->stream>>>(P);
  DM_CUDA(cudaGetLastError());
  count_launch(ctx);
  return DM_OK;
}
}  // namespace dm

extern "C" int dm_match_extract_raw_ssd(dm_ctx *ctx, const dm_pair *in, int maxh, int maxw, double threshold,
                                        int64_t *ret, float *scores, int64_t *n_untouched) {
  DM_REQUIRE(ctx && in && ret && scores, "dm_match_extract_raw_ssd: NULL argument");
  DM_CUDA(cudaSetDevice(ctx->device));
  dm::Call call(ctx);
  const int rc = dm::raw_ssd_extract_on(call, in, maxh, maxw, threshold, ret, scores, n_untouched);
  const int rf = call.finish();
  return rc != DM_OK ? rc : rf;
}
