// filter.cu -- the feature extractor in front of the matching path:
// getFilter (opticalflow_model.lua:45-79, radial/radial_opticalflow_network.lua:6-31):
// nn.SpatialConvolution / nn.SpatialConvolutionMap(nn.tables.random) layers with nn.Tanh
// between them, after nn.SpatialZeroPadding for the multiscale prefilter
// (opticalflow_model_multiscale.lua:134-173).
//
// Direct fp32 convolution on the CUDA cores: the features feed an exact SSD argmin, so the
// layer keeps fp32 products and accumulation (tf32/bf16 tensor-core operands lose the 1e-4
// parity bar).  One CTA owns a TW x TH output tile of every output plane; the input tile of all
// input planes (halo included, zero padding materialised) sits in shared memory.  A warp owns
// one output plane of a 32 x 16 sub-tile, a thread a 4 x 4 block of it: per input row and
// 4-tap chunk it reads 8 floats (two conflict-free LDS.128) and feeds 64 FMAs; the 4-tap weight
// vectors are warp-uniform (one broadcast LDG.128 each, L1-resident).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dm_common.cuh"
#include "filter_tc.cuh"

struct dm_filter {
  struct Layer {
    int n_in, n_out, kh, kw, n_conn, tanh_after;
    int taps;     // taps per chunk: 5 when kw == 5 (one chunk, no padding), else 4
    int nchunks;  // chunks per kernel row
    int rowlen;   // float2 pairs per packed weight row (16-byte multiple)
    int max_conn_per_out;
    float *weight;  // [n_conn][kh + 1][rowlen][2]: pairs {w[ky][kx], w[ky-1][kx]}, zero outside;
                    // connections grouped by output plane
    float *bias;    // [n_out]
    int *row_ptr;   // [n_out + 1]
    int *from;      // [n_conn]
    dm::TcPlan tc;  // tensor-core path (filter_tc.cu): plan and packed hi / lo weight operands
    float *tcB;
  };
  int device;
  std::vector<Layer> layers;
  void *blob = nullptr;
};

namespace dm {

constexpr int kSubW = 32, kSubH = 16;  // warp sub-tile: 8 x 4 lanes of 4 x 4 outputs
constexpr int kMaxConvWarps = 20;

struct ConvArgs {
  const float *in;
  float *out;
  const float *weight, *bias;
  const int *row_ptr, *from;
  int n_in, n_out, kh, nchunks, rowlen, tanh_after;
  int h, w, hout, wout, pad_t, pad_l;
  int tw, th, pitch, rows;
  int group;       // output planes per weight phase
  int tile_floats; // offset of the weight stage in shared memory
};

// One input row of one 4-tap (5-tap) chunk: 8 window floats against the packed weight rows
// r (output rows 0,1) and r-2 (output rows 2,3).
template <int J, bool A, bool B>
__device__ __forceinline__ void conv_step(const float *row, const float2 *wa, const float2 *wb,
                                          float2 (&acc01)[4], float2 (&acc23)[4]) {
  const float4 lo = *reinterpret_cast<const float4 *>(row);
  const float4 hi = *reinterpret_cast<const float4 *>(row + 4);
  const float win[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  if (A) {
    float2 w[6];
    *reinterpret_cast<float4 *>(&w[0]) = *reinterpret_cast<const float4 *>(wa);
    *reinterpret_cast<float4 *>(&w[2]) = *reinterpret_cast<const float4 *>(wa + 2);
    if (J == 5) w[4] = wa[4];
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
      for (int p = 0; p < 4; ++p) acc01[p] = __ffma2_rn(w[j], make_float2(win[p + j], win[p + j]), acc01[p]);
  }
  if (B) {
    float2 w[6];
    *reinterpret_cast<float4 *>(&w[0]) = *reinterpret_cast<const float4 *>(wb);
    *reinterpret_cast<float4 *>(&w[2]) = *reinterpret_cast<const float4 *>(wb + 2);
    if (J == 5) w[4] = wb[4];
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
      for (int p = 0; p < 4; ++p) acc23[p] = __ffma2_rn(w[j], make_float2(win[p + j], win[p + j]), acc23[p]);
  }
}

template <int J>
__global__ void conv_tile_kernel(const ConvArgs a) {
  extern __shared__ __align__(16) float tile[];  // [n_in][rows][pitch] | weight stage
  float2 *wsm = reinterpret_cast<float2 *>(tile + a.tile_floats);
  const int img = blockIdx.z;
  const int ox0 = blockIdx.x * a.tw, oy0 = blockIdx.y * a.th;
  const int plane_s = a.rows * a.pitch;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;   // shuffle: warp-uniform for the compiler
  {  // input footprint of every plane, zero padding materialised (cp.async zero-fill); one
     // warp per row, every copy in flight before anybody waits
    const float *src = a.in + (size_t)img * a.n_in * a.h * a.w;
    const int nrows = a.n_in * a.rows;
    for (int row = warp; row < nrows; row += nwarps) {
      const int c = row / a.rows, r = row - c * a.rows;
      const int iy = oy0 - a.pad_t + r;
      const bool row_ok = iy >= 0 && iy < a.h;
      const float *s = src + ((size_t)c * a.h + (row_ok ? iy : 0)) * a.w;
      const uint32_t d = smem_u32(tile + row * a.pitch);
      for (int x = lane; x < a.pitch; x += 32) {
        const int ix = ox0 - a.pad_l + x;
        const bool ok = row_ok && ix >= 0 && ix < a.w;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d + 4 * x), "l"(s + (ok ? ix : 0)),
                     "r"(ok ? 4 : 0)
                     : "memory");
      }
    }
  }

  const int lx = lane & 7, ly = lane >> 3;
  const int subs_x = a.tw / kSubW, subs = subs_x * (a.th / kSubH);
  const int wrow = a.rowlen;                // float2 per packed weight row
  const int wconn = (a.kh + 1) * wrow;      // float2 per connection
  for (int g0 = 0; g0 < a.n_out; g0 += a.group) {
    const int g1 = min(g0 + a.group, a.n_out);
    const int c0 = __ldg(a.row_ptr + g0), c1 = __ldg(a.row_ptr + g1);
    __syncthreads();  // the previous phase is done with the weight stage
    {
      const float4 *src = reinterpret_cast<const float4 *>(a.weight) + (size_t)c0 * (wconn / 2);
      const uint32_t dst = smem_u32(wsm);
      const int n4 = (c1 - c0) * (wconn / 2);
      for (int i = threadIdx.x; i < n4; i += blockDim.x)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * i), "l"(src + i) : "memory");
      asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");  // the tile's copies too
    }
    __syncthreads();
    const int n_items = (g1 - g0) * subs;
    for (int item = warp; item < n_items; item += nwarps) {
      const int to = g0 + item / subs, sub = item % subs;
      const int x0 = (sub % subs_x) * kSubW + lx * 4, y0 = (sub / subs_x) * kSubH + ly * 4;
      // acc01[p] = output rows (0, 1) of column p, acc23[p] = rows (2, 3): one FFMA2 each
      float2 acc01[4], acc23[4];
      const float b = __ldg(a.bias + to);
#pragma unroll
      for (int p = 0; p < 4; ++p) acc01[p] = acc23[p] = make_float2(b, b);
      const int e1 = __ldg(a.row_ptr + to + 1);
      for (int e = __ldg(a.row_ptr + to); e < e1; ++e) {
        const float2 *wbase = wsm + (e - c0) * wconn;
        const float *tp = tile + __ldg(a.from + e) * plane_s + y0 * a.pitch + x0;
        for (int c = 0; c < a.nchunks; ++c) {
          // input row r feeds output row t through tap row r - t; the packed weight row r
          // holds {w[r][j], w[r-1][j]}: rows (0,1) use row r, rows (2,3) row r - 2
          const float *row = tp + c * 4;
          const float2 *wq = wbase + c * J;
          conv_step<J, true, false>(row, wq, wq, acc01, acc23);
          conv_step<J, true, false>(row + a.pitch, wq + wrow, wq, acc01, acc23);
          row += 2 * a.pitch;
#pragma unroll 2
          for (int r = 2; r <= a.kh; ++r, row += a.pitch, wq += wrow)
            conv_step<J, true, true>(row, wq + 2 * wrow, wq, acc01, acc23);
          conv_step<J, false, true>(row, wq, wq, acc01, acc23);
          conv_step<J, false, true>(row + a.pitch, wq, wq + wrow, acc01, acc23);
        }
      }
      float *dst = a.out + ((size_t)img * a.n_out + to) * a.hout * a.wout;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int oy = oy0 + y0 + t;
        if (oy >= a.hout) break;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int ox = ox0 + x0 + p;
          const float2 pr = t < 2 ? acc01[p] : acc23[p];
          const float v = (t & 1) ? pr.y : pr.x;
          if (ox < a.wout) dst[(size_t)oy * a.wout + ox] = a.tanh_after ? tanhf(v) : v;
        }
      }
    }
  }
}

static int launch_layer(dm_ctx *ctx, const dm_filter::Layer &L, const float *in, float *out, int n_img,
                        int h, int w, int pad_l, int pad_r, int pad_t, int pad_b) {
  ConvArgs a{};
  a.in = in;
  a.out = out;
  a.weight = L.weight;
  a.bias = L.bias;
  a.row_ptr = L.row_ptr;
  a.from = L.from;
  a.n_in = L.n_in;
  a.n_out = L.n_out;
  a.kh = L.kh;
  a.nchunks = L.nchunks;
  a.rowlen = L.rowlen;
  a.tanh_after = L.tanh_after;
  a.h = h;
  a.w = w;
  a.hout = h + pad_t + pad_b - L.kh + 1;
  a.wout = w + pad_l + pad_r - L.kw + 1;
  a.pad_t = pad_t;
  a.pad_l = pad_l;
  const void *kernel = L.taps == 5 ? (const void *)conv_tile_kernel<5> : (const void *)conv_tile_kernel<4>;
  // Tile shape and output planes per weight phase: the combination that wastes the least of
  // the machine (partial last wave, outputs past the edge, idle warps in a round, too few
  // resident warps, halo re-reads, weight phases), among those whose input footprint + weight
  // stage fit in shared memory.  Weights of the terms fitted on c1's two layers (DM_CONV_TILE
  // sweeps, scripts/bench_filter.py).
  const int cand[4][2] = {{64, 32}, {64, 16}, {32, 32}, {32, 16}};
  const size_t out_bytes = (size_t)L.max_conn_per_out * (L.kh + 1) * L.rowlen * 8;  // weights of one plane
  double best = -1.0;
  int pick_tw = 0, pick_th = 0, pick_warps = 0, pick_group = 0;
  size_t pick_smem = 0, pick_tile = 0;
  // option conv_tile = "<candidate 0-3>,<CTAs per SM 1-2>": tuning only
  const int force_target = ctx->opt.conv_target, force_tile = force_target ? ctx->opt.conv_tile : -1;
  for (int i = 0; i < 4; ++i) {
    for (int per_sm_target = 2; per_sm_target >= 1; --per_sm_target) {
      if (force_tile >= 0 && (i != force_tile || per_sm_target != force_target)) continue;
      const int tw = cand[i][0], th = cand[i][1];
      const size_t tile_bytes = (size_t)L.n_in * (th + L.kh - 1) * (tw + 4 * L.nchunks) * 4;
      const size_t budget = per_sm_target == 1 ? ctx->smem_optin : (ctx->smem_optin + 1024) / 2 - 1024;
      if (tile_bytes + out_bytes > budget) continue;
      const int group = (int)std::min<size_t>(L.n_out, (budget - tile_bytes) / out_bytes);
      const size_t bytes = tile_bytes + group * out_bytes;
      const int subs = (tw / kSubW) * (th / kSubH);
      const int items = group * subs;
      const int min_rounds = (items + kMaxConvWarps - 1) / kMaxConvWarps;
      if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
        continue;
      for (int rounds = min_rounds; rounds <= min_rounds + 2 && rounds <= items; ++rounds) {
        const int nwarps = (items + rounds - 1) / rounds;
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, nwarps * 32, bytes) != cudaSuccess ||
            per_sm < 1)
          continue;
        const long long ctas = (long long)((a.wout + tw - 1) / tw) * ((a.hout + th - 1) / th) * n_img;
        const long long slots = (long long)per_sm * ctx->num_sms;
        const long long waves = (ctas + slots - 1) / slots;
        const double fill = (double)ctas / (double)(waves * slots);                      // last wave
        const double edge = (double)a.wout * a.hout * n_img / ((double)ctas * tw * th);  // past the edge
        int total_rounds = 0;
        for (int g0 = 0; g0 < L.n_out; g0 += group)
          total_rounds += (std::min(group, L.n_out - g0) * subs + nwarps - 1) / nwarps;
        const double lanes = (double)L.n_out * subs / ((double)total_rounds * nwarps);   // idle warps
        const double resident = std::min(1.0, per_sm * nwarps / 16.0);                   // latency hiding
        const double overlap = per_sm >= 2 ? 1.0 : 0.97;  // a second CTA computes during the tile load
        const double halo = (double)tw * th / ((double)(tw + 4 * L.nchunks) * (th + L.kh - 1));
        const int phases = (L.n_out + group - 1) / group;  // each one: barrier + weight reload
        const double score = fill * edge * lanes * resident * overlap * (0.75 + 0.25 * halo) *
                             (1.0 - 0.04 * (phases - 1));
        if (score > best * 1.03) {
          best = score;
          pick_tw = tw;
          pick_th = th;
          pick_warps = nwarps;
          pick_group = group;
          pick_smem = bytes;
          pick_tile = tile_bytes;
        }
      }
    }
  }
  DM_REQUIRE(best > 0, "filter layer %d x %d x %d needs more than %zu bytes of shared memory per tile", L.n_in,
             L.kh, L.kw, (size_t)ctx->smem_optin);
  a.tw = pick_tw;
  a.th = pick_th;
  a.pitch = a.tw + 4 * L.nchunks;
  a.rows = a.th + L.kh - 1;
  a.group = pick_group;
  a.tile_floats = (int)(pick_tile / 4);
  DM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pick_smem));
  dim3 grid((a.wout + a.tw - 1) / a.tw, (a.hout + a.th - 1) / a.th, n_img);
  if (L.taps == 5)
    conv_tile_kernel<5><<<grid, pick_warps * 32, pick_smem, ctx->stream>>>(a);
  else
    conv_tile_kernel<4><<<grid, pick_warps * 32, pick_smem, ctx->stream>>>(a);
  count_launch(ctx);
  return DM_OK;
}

}  // namespace dm

using namespace dm;

extern "C" {

int dm_filter_create(dm_ctx *ctx, const dm_conv_layer *layers, int n_layers, dm_filter **out) {
  DM_REQUIRE(ctx && layers && out, "dm_filter_create: NULL argument");
  DM_REQUIRE(n_layers >= 1 && n_layers <= 16, "dm_filter_create: %d layers", n_layers);
  DM_CUDA(cudaSetDevice(ctx->device));
  // pack everything into one host image, then one device blob
  std::vector<char> host;
  auto reserve = [&](size_t bytes) {
    const size_t at = (host.size() + 255) & ~size_t(255);
    host.resize(at + bytes, 0);
    return at;
  };
  struct Off {
    size_t weight, bias, row_ptr, from;
  };
  std::vector<Off> offs(n_layers);
  dm_filter *f = new dm_filter;
  f->device = ctx->device;
  int prev_out = -1;
  for (int i = 0; i < n_layers; ++i) {
    const dm_conv_layer &l = layers[i];
    if (!(l.n_in >= 1 && l.n_out >= 1 && l.kh >= 1 && l.kw >= 1 && l.weight && l.bias) ||
        (l.n_conn > 0 && !l.conn) || (prev_out >= 0 && l.n_in != prev_out)) {
      delete f;
      DM_REQUIRE(false, "dm_filter_create: layer %d is inconsistent (n_in %d after %d outputs, kernel %d x %d)",
                 i + 1, l.n_in, prev_out, l.kh, l.kw);
    }
    const bool full = l.n_conn <= 0;
    const int nc = full ? l.n_in * l.n_out : l.n_conn;
    dm_filter::Layer L{};
    L.n_in = l.n_in;
    L.n_out = l.n_out;
    L.kh = l.kh;
    L.kw = l.kw;
    L.taps = l.kw == 5 ? 5 : 4;
    L.nchunks = L.taps == 5 ? 1 : (l.kw + 3) / 4;
    L.rowlen = L.taps == 5 ? 6 : 4 * L.nchunks;
    L.n_conn = nc;
    L.tanh_after = l.tanh_after;
    offs[i].weight = reserve((size_t)nc * (L.kh + 1) * L.rowlen * 2 * 4);
    offs[i].bias = reserve((size_t)L.n_out * 4);
    offs[i].row_ptr = reserve((size_t)(L.n_out + 1) * 4);
    offs[i].from = reserve((size_t)nc * 4);
    float *wdst = reinterpret_cast<float *>(host.data() + offs[i].weight);
    int *rp = reinterpret_cast<int *>(host.data() + offs[i].row_ptr);
    int *fr = reinterpret_cast<int *>(host.data() + offs[i].from);
    memcpy(host.data() + offs[i].bias, l.bias, (size_t)L.n_out * 4);
    int slot = 0;
    for (int o = 0; o < L.n_out; ++o) {  // group by output plane, table order kept inside
      rp[o] = slot;
      for (int e = 0; e < nc; ++e) {
        const int from = full ? e % l.n_in : l.conn[2 * e] - 1;
        const int to = full ? e / l.n_in : l.conn[2 * e + 1] - 1;
        if (from < 0 || from >= l.n_in || to < 0 || to >= l.n_out) {
          delete f;
          DM_REQUIRE(false, "dm_filter_create: layer %d connection %d = (%d, %d) outside %d x %d", i + 1,
                     e + 1, from + 1, to + 1, l.n_in, l.n_out);
        }
        if (to != o) continue;
        fr[slot] = from;
        const float *wsrc = l.weight + (size_t)e * L.kh * L.kw;
        for (int ky = 0; ky <= L.kh; ++ky)
          for (int kx = 0; kx < L.kw; ++kx) {
            float *pair = wdst + (((size_t)slot * (L.kh + 1) + ky) * L.rowlen + kx) * 2;
            pair[0] = ky < L.kh ? wsrc[ky * L.kw + kx] : 0.0f;
            pair[1] = ky >= 1 ? wsrc[(ky - 1) * L.kw + kx] : 0.0f;
          }
        ++slot;
      }
    }
    rp[L.n_out] = slot;
    for (int o = 0; o < L.n_out; ++o) L.max_conn_per_out = std::max(L.max_conn_per_out, rp[o + 1] - rp[o]);
    if (L.max_conn_per_out < 1) L.max_conn_per_out = 1;
    prev_out = l.n_out;
    L.tcB = nullptr;
    if (dm::tc_plan_layer(ctx, l.n_in, l.n_out, l.kh, l.kw, &L.tc)) {
      std::vector<float> packed;
      dm::tc_pack_weights(L.tc, l.n_in, l.n_out, l.kh, l.kw, l.n_conn, l.conn, l.weight, &packed);
      cudaError_t e2 = cudaMalloc(&L.tcB, packed.size() * sizeof(float));
      if (e2 == cudaSuccess) e2 = cudaMemcpy(L.tcB, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice);
      if (e2 != cudaSuccess) {
        cudaGetLastError();
        if (L.tcB) cudaFree(L.tcB);
        L.tcB = nullptr;
        L.tc.ok = 0;   // the CUDA-core kernel still serves the layer
      }
    }
    f->layers.push_back(L);
  }
  cudaError_t e = cudaMalloc(&f->blob, host.size());
  if (e == cudaSuccess) e = cudaMemcpy(f->blob, host.data(), host.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (f->blob) cudaFree(f->blob);
    delete f;
    DM_CUDA(e);
  }
  char *base = static_cast<char *>(f->blob);
  for (int i = 0; i < n_layers; ++i) {
    f->layers[i].weight = reinterpret_cast<float *>(base + offs[i].weight);
    f->layers[i].bias = reinterpret_cast<float *>(base + offs[i].bias);
    f->layers[i].row_ptr = reinterpret_cast<int *>(base + offs[i].row_ptr);
    f->layers[i].from = reinterpret_cast<int *>(base + offs[i].from);
  }
  *out = f;
  return DM_OK;
}

int dm_filter_destroy(dm_filter *f) {
  if (!f) return DM_OK;
  cudaSetDevice(f->device);
  if (f->blob) cudaFree(f->blob);
  for (auto &L : f->layers)
    if (L.tcB) cudaFree(L.tcB);
  delete f;
  return DM_OK;
}

int dm_filter_output_size(const dm_filter *f, int h, int w, int pad_l, int pad_r, int pad_t, int pad_b,
                          int *channels, int *hout, int *wout) {
  DM_REQUIRE(f && channels && hout && wout, "dm_filter_output_size: NULL argument");
  h += pad_t + pad_b;
  w += pad_l + pad_r;
  for (const auto &L : f->layers) {
    h -= L.kh - 1;
    w -= L.kw - 1;
  }
  *channels = f->layers.back().n_out;
  *hout = h;
  *wout = w;
  DM_REQUIRE(h >= 1 && w >= 1, "dm_filter_output_size: the input is smaller than the filter's footprint");
  return DM_OK;
}

int dm_filter_forward(dm_ctx *ctx, const dm_filter *f, const float *in, int n_img, int h, int w, int pad_l,
                      int pad_r, int pad_t, int pad_b, float *out) {
  DM_REQUIRE(ctx && f && in && out, "dm_filter_forward: NULL argument");
  DM_REQUIRE(f->device == ctx->device, "dm_filter_forward: the filter lives on device %d, the context on %d",
             f->device, ctx->device);
  DM_REQUIRE(n_img >= 1 && h >= 1 && w >= 1 && pad_l >= 0 && pad_r >= 0 && pad_t >= 0 && pad_b >= 0,
             "dm_filter_forward: bad shape");
  int co, ho, wo;
  DM_CHECK(dm_filter_output_size(f, h, w, pad_l, pad_r, pad_t, pad_b, &co, &ho, &wo));
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  const void *din;
  void *dout;
  DM_CHECK(call.in(in, (size_t)n_img * f->layers[0].n_in * h * w * 4, &din));
  DM_CHECK(call.out(out, (size_t)n_img * co * ho * wo * 4, &dout));
  const float *cur = static_cast<const float *>(din);
  int ch = h, cw = w;
  prof_begin(ctx);
  for (size_t i = 0; i < f->layers.size(); ++i) {
    const auto &L = f->layers[i];
    const int pl = i == 0 ? pad_l : 0, pr = i == 0 ? pad_r : 0, pt = i == 0 ? pad_t : 0, pb = i == 0 ? pad_b : 0;
    const int nh = ch + pt + pb - L.kh + 1, nw = cw + pl + pr - L.kw + 1;
    float *dst;
    if (i + 1 == f->layers.size()) {
      dst = static_cast<float *>(dout);
    } else {
      void *tmp;
      DM_CHECK(call.alloc(&tmp, (size_t)n_img * L.n_out * nh * nw * 4));
      dst = static_cast<float *>(tmp);
    }
    // option conv = 2: the tcgen05 tensor-core kernel (filter_tc.cu) where the layer fits it.  It is
    // parity-green but measured slower than the CUDA-core kernel on B200 (three-term
    // tf32 split, connection-table zeros and the per-row operand staging eat the tensor advantage on
    // these 3..16-plane layers), so it is opt-in.
    if (L.tc.ok && ctx->opt.conv == 2)
      DM_CHECK(dm::tc_launch_layer(ctx, L.tc, L.tcB, L.bias, L.n_in, L.n_out, L.kh, L.kw, L.tanh_after, cur, dst, n_img,
                                   ch, cw, pl, pr, pt, pb));
    else
      DM_CHECK(launch_layer(ctx, L, cur, dst, n_img, ch, cw, pl, pr, pt, pb));
    cur = dst;
    ch = nh;
    cw = nw;
  }
  prof_end(ctx);
  return call.finish();
}

}  // extern "C"
