// match_kernels.cuh -- the tiled SSD sweep shared by the fused extract kernel and
// the volume kernel.
//
// Work decomposition (one CTA = 8 warps = a tile of TH=8 output rows x TW=128
// output columns of one frame pair):
//   * warp w owns output row y0+w; lane l owns P=4 consecutive pixels x0+4l..x0+4l+3
//     whose C feature values (frame 1) stay in registers for the whole sweep;
//   * the sweep runs dy = 0..maxh-1.  At step dy warp w needs frame-2 row
//     y0+w+dy, columns x0..x0+TW+maxw-2, all C channels: one "row slab"
//     [C][WB] floats.  Slabs stream through a ring of NSLOT shared-memory slots
//     filled by TMA (cp.async.bulk.tensor 4-D box {WB,1,C,1}, zero fill outside
//     the frame) and signalled through one mbarrier per slot, so a CTA holds
//     only TH+4 rows of the halo at any time, whatever the window height;
//   * inside a step the dx range is cut in blocks of R=8 (last block up to 9
//     wide): a thread turns P+R-1 <= 12 floats of the slab (3 LDS.128) into
//     P*R SSD partial sums per channel -- 64..72 FSUB+FFMA per 3 loads.
// The per-block epilogue is supplied by the caller (extract or volume).
#pragma once

#include "dm_common.cuh"

namespace dm {

constexpr int kP = 4;              // pixels per thread
constexpr int kWarps = 8;          // output rows per tile (one per warp)
constexpr int kThreads = kWarps * 32;
constexpr int kTW = 32 * kP;       // 128 output columns per tile
constexpr int kTH = kWarps;
constexpr int kR = 8;              // displacement block width
constexpr int kRT = 9;             // tail block width (masked)
constexpr int kNB = 12;            // floats of the slab a thread reads per channel (3 x float4)
constexpr int kPrefetch = 4;       // rows in flight beyond the TH the warps are reading
constexpr int kNSlot = kTH + kPrefetch;
constexpr int kMaxC = 16;          // channels supported by the tiled kernels

// Block schedule for a window width: nfull blocks of 8 then one tail of 1..9.
__host__ __device__ inline void block_schedule(int maxw, int *nfull, int *tailw) {
  int nf = maxw > kRT ? (maxw - kRT + kR - 1) / kR : 0;
  *nfull = nf;
  *tailw = maxw - nf * kR;
}

// slab width in floats: covers x0 .. x0 + TW + (nfull*8 + 12) - 4, multiple of 4
__host__ __device__ inline int slab_width(int maxw) {
  int nfull, tailw;
  block_schedule(maxw, &nfull, &tailw);
  return kTW - kP + nfull * kR + kNB;
}

struct SweepGeom {
  int N, C, Cin, H1, W1, H2, W2, maxh, maxw;  // C: channels of the slab box (>= Cin, zero-filled)
  int tiles_x, tiles_y, ntiles;
  int WB, nfull, tailw;
  const float *in1;
  long long s1n, s1c, s1y;
};

// One channel-complete SSD block: acc[p][r] = sum_k (a[k][p] - slab[k][p + r])^2.
// EXACT keeps multiply and add separate (bit-exact with the non-contracting CPU path).
template <int CT, int R, bool EXACT>
__device__ __forceinline__ void ssd_block(const float (&a)[CT][kP], const float *bsrc, int WB,
                                          float (&acc)[kP][R]) {
#pragma unroll
  for (int p = 0; p < kP; ++p)
#pragma unroll
    for (int r = 0; r < R; ++r) acc[p][r] = 0.0f;
#pragma unroll
  for (int k = 0; k < CT; ++k) {
    float b[kNB];
    const float4 *src = reinterpret_cast<const float4 *>(bsrc + k * WB);
#pragma unroll
    for (int j = 0; j < kNB / 4; ++j) {
      const float4 t = src[j];
      b[4 * j + 0] = t.x;
      b[4 * j + 1] = t.y;
      b[4 * j + 2] = t.z;
      b[4 * j + 3] = t.w;
    }
#pragma unroll
    for (int p = 0; p < kP; ++p)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float d = a[k][p] - b[p + r];
        if (EXACT)
          acc[p][r] = __fadd_rn(acc[p][r], __fmul_rn(d, d));
        else
          acc[p][r] = fmaf(d, d, acc[p][r]);
      }
  }
}

// Runs the whole sweep for the tiles of this CTA.  `Epi` supplies:
//   void tile_begin(n, y, x0)                       per tile, after `a` is loaded
//   void block<R>(acc, dy, dxb, rvalid)             per (dy, dx-block), all threads
//   void tile_end(n, y, x0)                         per tile
template <int CT, bool EXACT, class Epi>
__device__ __forceinline__ void run_sweep(const CUtensorMap *tmap, const SweepGeom &g, float *ring,
                                          uint64_t *full, Epi &epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_total = kTH + g.maxh - 1;
  const uint32_t slab_bytes = (uint32_t)(g.C * g.WB * sizeof(float));
  const int slab_floats = g.C * g.WB;

  if (threadIdx.x == 0) {
    prefetch_tmap(tmap);
    for (int s = 0; s < kNSlot; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();

  uint32_t g0 = 0;  // sequence number of row 0 of the current tile
  for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
    const int tx = tile % g.tiles_x;
    const int ty = (tile / g.tiles_x) % g.tiles_y;
    const int n = tile / (g.tiles_x * g.tiles_y);
    const int y0 = ty * kTH, xt = tx * kTW;
    const int y = y0 + warp, x0 = xt + lane * kP;

    auto issue_row = [&](int j) {
      const uint32_t seq = g0 + (uint32_t)j;
      const int slot = (int)(seq % kNSlot);
      if (y0 + j < g.H2) {
        mbar_arrive_expect_tx(&full[slot], slab_bytes);
        tma_load_4d(ring + slot * slab_floats, tmap, &full[slot], xt, y0 + j, 0, n);
      } else {
        mbar_arrive(&full[slot]);  // row below the frame: only masked pixels read it
      }
    };
    if (threadIdx.x == 0) {
      const int pro = rows_total < kNSlot ? rows_total : kNSlot;
      for (int j = 0; j < pro; ++j) issue_row(j);
    }

    float a[CT][kP];
    {
      const bool rowok = y < g.H1;
      const float *src = g.in1 + (long long)n * g.s1n + (long long)(rowok ? y : 0) * g.s1y;
#pragma unroll
      for (int k = 0; k < CT; ++k)
#pragma unroll
        for (int p = 0; p < kP; ++p)
          a[k][p] = (rowok && k < g.Cin && x0 + p < g.W1)
                        ? __ldg(src + (long long)k * g.s1c + x0 + p)
                        : 0.0f;
    }
    epi.tile_begin(n, y, x0);

    for (int dy = 0; dy < g.maxh; ++dy) {
      const uint32_t seq = g0 + (uint32_t)(warp + dy);
      const int slot = (int)(seq % kNSlot);
      mbar_wait(&full[slot], (seq / kNSlot) & 1u);
      const float *brow = ring + slot * slab_floats + lane * kP;
      for (int blk = 0; blk < g.nfull; ++blk) {
        float acc[kP][kR];
        ssd_block<CT, kR, EXACT>(a, brow + blk * kR, g.WB, acc);
        epi.template block<kR>(acc, dy, blk * kR, kR);
      }
      {
        float acc[kP][kRT];
        ssd_block<CT, kRT, EXACT>(a, brow + g.nfull * kR, g.WB, acc);
        epi.template block<kRT>(acc, dy, g.nfull * kR, g.tailw);
      }
      __syncthreads();  // every warp is done with row y0+dy: its slot can be refilled
      if (threadIdx.x == 0 && dy + kNSlot < rows_total) issue_row(dy + kNSlot);
    }
    epi.tile_end(n, y, x0);
    g0 += (uint32_t)rows_total;
  }
}

}  // namespace dm
