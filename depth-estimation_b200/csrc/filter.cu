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
#include <vector>

#include "dm_common.cuh"

struct dm_filter {
  struct Layer {
    int n_in, n_out, kh, kw, kw4, n_conn, tanh_after;
    float *weight;  // [n_conn][kh][kw4], connections grouped by output plane, rows zero padded
    float *bias;    // [n_out]
    int *row_ptr;   // [n_out + 1]
    int *from;      // [n_conn]
  };
  int device;
  std::vector<Layer> layers;
  void *blob = nullptr;
};

namespace dm {

constexpr int kSubW = 32, kSubH = 16;  // warp sub-tile: 8 x 4 lanes of 4 x 4 outputs

struct ConvArgs {
  const float *in;
  float *out;
  const float *weight, *bias;
  const int *row_ptr, *from;
  int n_in, n_out, kh, kw4, tanh_after;
  int h, w, hout, wout, pad_t, pad_l;
  int tw, th, pitch, rows;
};

__global__ void conv_tile_kernel(const ConvArgs a) {
  extern __shared__ __align__(16) float tile[];  // [n_in][rows][pitch]
  const int img = blockIdx.z;
  const int ox0 = blockIdx.x * a.tw, oy0 = blockIdx.y * a.th;
  const int plane_s = a.rows * a.pitch;
  {
    const float *src = a.in + (size_t)img * a.n_in * a.h * a.w;
    const int total = a.n_in * plane_s;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int c = idx / plane_s, rem = idx - c * plane_s;
      const int r = rem / a.pitch, x = rem - r * a.pitch;
      const int iy = oy0 - a.pad_t + r, ix = ox0 - a.pad_l + x;
      float v = 0.0f;
      if (iy >= 0 && iy < a.h && ix >= 0 && ix < a.w) v = __ldg(src + ((size_t)c * a.h + iy) * a.w + ix);
      tile[idx] = v;
    }
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int lx = lane & 7, ly = lane >> 3;
  const int subs_x = a.tw / kSubW, subs = subs_x * (a.th / kSubH);
  const int n_items = a.n_out * subs;
  const int wstride = a.kh * a.kw4;
  for (int item = warp; item < n_items; item += nwarps) {
    const int to = item / subs, sub = item - to * subs;
    const int x0 = (sub % subs_x) * kSubW + lx * 4, y0 = (sub / subs_x) * kSubH + ly * 4;
    float acc[4][4];
    const float b = __ldg(a.bias + to);
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[t][p] = b;
    const int e1 = __ldg(a.row_ptr + to + 1);
    for (int e = __ldg(a.row_ptr + to); e < e1; ++e) {
      const float *wp = a.weight + (size_t)e * wstride;
      const float *tp = tile + __ldg(a.from + e) * plane_s + y0 * a.pitch + x0;
      for (int r = 0; r < a.kh + 3; ++r) {
        const float *row = tp + r * a.pitch;
        for (int kx0 = 0; kx0 < a.kw4; kx0 += 4) {
          const float4 lo = *reinterpret_cast<const float4 *>(row + kx0);
          const float4 hi = *reinterpret_cast<const float4 *>(row + kx0 + 4);
          const float win[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int ky = r - t;  // input row r feeds output row t through tap row r - t
            if ((unsigned)ky < (unsigned)a.kh) {
              const float4 wv = __ldg(reinterpret_cast<const float4 *>(wp + ky * a.kw4 + kx0));
#pragma unroll
              for (int p = 0; p < 4; ++p) {
                acc[t][p] = fmaf(win[p], wv.x, acc[t][p]);
                acc[t][p] = fmaf(win[p + 1], wv.y, acc[t][p]);
                acc[t][p] = fmaf(win[p + 2], wv.z, acc[t][p]);
                acc[t][p] = fmaf(win[p + 3], wv.w, acc[t][p]);
              }
            }
          }
        }
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
        if (ox < a.wout) dst[(size_t)oy * a.wout + ox] = a.tanh_after ? tanhf(acc[t][p]) : acc[t][p];
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
  a.kw4 = L.kw4;
  a.tanh_after = L.tanh_after;
  a.h = h;
  a.w = w;
  a.hout = h + pad_t + pad_b - L.kh + 1;
  a.wout = w + pad_l + pad_r - L.kw + 1;
  a.pad_t = pad_t;
  a.pad_l = pad_l;
  // the largest tile whose input footprint fits; prefer two CTAs per SM and a grid that fills
  // the machine
  const int cand[4][2] = {{64, 32}, {64, 16}, {32, 16}, {32, 16}};
  size_t smem = 0;
  int pick = -1;
  for (int i = 0; i < 3; ++i) {
    const int tw = cand[i][0], th = cand[i][1];
    const size_t bytes = (size_t)L.n_in * (th + L.kh - 1) * (tw + L.kw4) * 4;
    const long long ctas = (long long)((a.wout + tw - 1) / tw) * ((a.hout + th - 1) / th) * n_img;
    const bool last = i == 2;
    if (bytes > ctx->smem_optin) continue;
    if (!last && (bytes > 110 * 1024 || ctas < 2LL * ctx->num_sms)) continue;
    pick = i;
    smem = bytes;
    break;
  }
  DM_REQUIRE(pick >= 0, "filter layer %d x %d x %d needs more than %zu bytes of shared memory per tile",
             L.n_in, L.kh, L.kw, (size_t)ctx->smem_optin);
  a.tw = cand[pick][0];
  a.th = cand[pick][1];
  a.pitch = a.tw + L.kw4;
  a.rows = a.th + L.kh - 1;
  const int items = L.n_out * (a.tw / kSubW) * (a.th / kSubH);
  const int rounds = (items + 15) / 16;
  const int nwarps = (items + rounds - 1) / rounds;
  DM_CUDA(cudaFuncSetAttribute(conv_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((a.wout + a.tw - 1) / a.tw, (a.hout + a.th - 1) / a.th, n_img);
  conv_tile_kernel<<<grid, nwarps * 32, smem, ctx->stream>>>(a);
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
    L.kw4 = (l.kw + 3) & ~3;
    L.n_conn = nc;
    L.tanh_after = l.tanh_after;
    offs[i].weight = reserve((size_t)nc * L.kh * L.kw4 * 4);
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
        for (int ky = 0; ky < L.kh; ++ky)
          memcpy(wdst + ((size_t)slot * L.kh + ky) * L.kw4, l.weight + ((size_t)e * L.kh + ky) * L.kw,
                 (size_t)L.kw * 4);
        ++slot;
      }
    }
    rp[L.n_out] = slot;
    prev_out = l.n_out;
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
    DM_CHECK(launch_layer(ctx, L, cur, dst, n_img, ch, cw, pl, pr, pt, pb));
    cur = dst;
    ch = nh;
    cw = nw;
  }
  prof_end(ctx);
  return call.finish();
}

}  // extern "C"
