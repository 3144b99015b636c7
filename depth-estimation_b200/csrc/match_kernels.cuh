// match_kernels.cuh -- the tiled SSD sweep shared by the fused extract kernel and
// the volume kernel.
//
// Work decomposition (one CTA = 15 consumer warps (+1 TMA producer warp) = a tile of TH=15 output rows x TW=128
// output columns of one frame pair):
//   * warp w owns output row y0+w; lane l owns P=4 consecutive pixels x0+4l..x0+4l+3
//     whose C feature values (frame 1) stay in registers for the whole sweep;
//   * the sweep runs dy = 0..maxh-1.  At step dy warp w needs frame-2 row
//     y0+w+dy, columns x0..x0+TW+maxw-2, all C channels: one "row slab"
//     [C][WB] floats.  Slabs stream through a ring of NSLOT shared-memory slots
//     filled by TMA (cp.async.bulk.tensor 4-D box {WB,1,C,1}, zero fill outside
//     the frame) and signalled through one mbarrier per slot, so a CTA holds
//     only TH+4 rows of the halo at any time, whatever the window height;
//   * inside a step the dx range is cut in blocks of R=8 (the last one masked when
//     the width is not a multiple of 8; a width of 8n+1 ends with one single
//     column instead): a thread turns P+R-1 = 11 floats of the slab (3 LDS.128)
//     into P*R = 32 SSD partial sums per channel -- 64 FSUB+FFMA per 3 loads.
// The per-block epilogue is supplied by the caller (extract or volume).
#pragma once

#include "dm_common.cuh"

namespace dm {

constexpr int kP = 4;              // pixels per thread
constexpr int kWarps = 15;         // consumer warps = output rows per tile (one per warp)
constexpr int kCThreads = kWarps * 32;    // consumer threads
constexpr int kThreads = kCThreads + 32;  // + one producer warp that issues the TMA loads
constexpr int kTW = 32 * kP;       // 128 output columns per tile
constexpr int kTH = kWarps;
constexpr int kR = 8;              // displacement block width
constexpr int kNB = 12;            // floats of the slab a thread reads per channel (3 x float4)
constexpr int kPrefetch = 8;       // rows in flight beyond the TH the warps are reading
constexpr int kNSlot = kTH + kPrefetch;
constexpr int kMaxC = 16;          // channels supported by the tiled kernels

// Block schedule for a window width: n8 blocks of 8 columns, the last of which has
// `last_valid` (1..8) real columns, then `single` (0/1) trailing single column.
struct BlockSchedule {
  int n8, last_valid, single;
  __host__ __device__ int per_row() const { return n8 + single; }
};
__host__ __device__ inline BlockSchedule block_schedule(int maxw) {
  BlockSchedule s;
  const int rem = maxw % kR;
  if (rem == 1 && maxw > 1) {
    s.n8 = maxw / kR;
    s.last_valid = kR;
    s.single = 1;
  } else {
    s.n8 = (maxw + kR - 1) / kR;
    s.last_valid = rem ? rem : kR;
    s.single = 0;
  }
  return s;
}

// slab width in floats (multiple of 4): lane 31 reads up to 124 + (n8-1)*8 + 12
__host__ __device__ inline int slab_width(int maxw) { return kTW + block_schedule(maxw).n8 * kR; }

struct SweepGeom {
  int N, C, Cin, H1, W1, H2, W2, maxh, maxw;  // C: channels of the slab box (>= Cin, zero-filled)
  int tiles_x, tiles_y, ntiles;
  int WB;
  BlockSchedule bs;
  int slab_floats;  // floats between ring slots: C*WB rounded up to 128 bytes
  const float *in1;
  long long s1n, s1c, s1y;
};

template <bool EXACT>
__device__ __forceinline__ float sq_first(float d) {
  return EXACT ? __fmul_rn(d, d) : d * d;
}
template <bool EXACT>
__device__ __forceinline__ float sq_acc(float d, float acc) {
  return EXACT ? __fadd_rn(acc, __fmul_rn(d, d)) : fmaf(d, d, acc);
}

// One channel-complete SSD block: acc[p][r] = sum_k (a[k][p] - slab[k][p + r])^2.
// EXACT keeps multiply and add separate (bit-exact with the non-contracting CPU path).
template <int CT, bool EXACT>
__device__ __forceinline__ void ssd_block(const float (&a)[CT][kP], const float *bsrc, int WB,
                                          float (&acc)[kP][kR]) {
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
      for (int r = 0; r < kR; ++r) {
        const float d = a[k][p] - b[p + r];
        acc[p][r] = k == 0 ? sq_first<EXACT>(d) : sq_acc<EXACT>(d, acc[p][r]);
      }
  }
}

// A single displacement column: acc[p] = sum_k (a[k][p] - slab[k][p])^2
template <int CT, bool EXACT>
__device__ __forceinline__ void ssd_column(const float (&a)[CT][kP], const float *bsrc, int WB,
                                           float (&acc)[kP]) {
#pragma unroll
  for (int k = 0; k < CT; ++k) {
    const float4 t = *reinterpret_cast<const float4 *>(bsrc + k * WB);
    const float b[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int p = 0; p < kP; ++p) {
      const float d = a[k][p] - b[p];
      acc[p] = k == 0 ? sq_first<EXACT>(d) : sq_acc<EXACT>(d, acc[p]);
    }
  }
}

// Runs the whole sweep for the tiles of this CTA.  `Epi` supplies:
//   void tile_begin(n, y, x0)                       per tile, after `a` is loaded
//   void block(acc[kP][8], dy, blk, rvalid)         per (dy, 8-wide dx-block), all threads
//   void column(acc[kP], dy, blk)                   per (dy, trailing single column)
//   void tile_end(n, y, x0)                         per tile
//
// Pipeline: rows of all the CTA's tiles form one sequence; row `seq` lives in ring slot
// seq % NSLOT.  The producer warp (one elected lane) waits for the slot's previous tenant to
// be released (empty[slot], one arrival per consumer warp), then issues the TMA load that
// completes full[slot].  Every consumer warp walks every row of its tile in order: wait
// full, compute if the row is inside its own window (row - warp in [0, maxh)), release.
// No CTA-wide barrier: warps drift apart by up to NSLOT - TH rows, across tile borders too.
template <int CT, bool EXACT, class Epi>
__device__ __forceinline__ void run_sweep(const CUtensorMap *tmap, const SweepGeom &g, float *ring,
                                          uint64_t *bars, Epi &epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_total = kTH + g.maxh - 1;
  const uint32_t slab_bytes = (uint32_t)(g.C * g.WB * sizeof(float));
  const int slab_floats = g.slab_floats;
  const int n8 = g.bs.n8;
  uint64_t *full = bars, *empty = bars + kNSlot;

  if (threadIdx.x == 0) {
    prefetch_tmap(tmap);
    for (int s = 0; s < kNSlot; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kWarps) {
    // ---------------- producer warp
    if (lane == 0) {
      uint32_t seq = 0;
      for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
        const int tx = tile % g.tiles_x;
        const int ty = (tile / g.tiles_x) % g.tiles_y;
        const int n = tile / (g.tiles_x * g.tiles_y);
        const int y0 = ty * kTH, xt = tx * kTW;
        for (int j = 0; j < rows_total; ++j, ++seq) {
          const int slot = (int)(seq % kNSlot);
          const uint32_t inst = seq / kNSlot;
          if (inst > 0) mbar_wait_backoff(&empty[slot], (inst - 1) & 1u);
          if (y0 + j < g.H2) {
            mbar_arrive_expect_tx(&full[slot], slab_bytes);
            tma_load_4d(ring + slot * slab_floats, tmap, &full[slot], xt, y0 + j, 0, n);
          } else {
            mbar_arrive(&full[slot]);  // row below the frame: only masked pixels read it
          }
        }
      }
    }
    return;
  }

  // ---------------- consumer warps
  uint32_t g0 = 0;  // sequence number of row 0 of the current tile
  for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
    const int tx = tile % g.tiles_x;
    const int ty = (tile / g.tiles_x) % g.tiles_y;
    const int n = tile / (g.tiles_x * g.tiles_y);
    const int y0 = ty * kTH, xt = tx * kTW;
    const int y = y0 + warp, x0 = xt + lane * kP;

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

#pragma unroll 1
    for (int j = 0; j < rows_total; ++j) {
      const uint32_t seq = g0 + (uint32_t)j;
      const int slot = (int)(seq % kNSlot);
      mbar_wait(&full[slot], (seq / kNSlot) & 1u);
      const int dy = j - warp;
      if (dy >= 0 && dy < g.maxh) {
        const float *brow = ring + slot * slab_floats + lane * kP;
#pragma unroll 1
        for (int blk = 0; blk < n8; ++blk) {
          float acc[kP][kR];
          ssd_block<CT, EXACT>(a, brow + blk * kR, g.WB, acc);
          epi.block(acc, dy, blk, blk == n8 - 1 ? g.bs.last_valid : kR);
        }
        if (g.bs.single) {
          float acc[kP];
          ssd_column<CT, EXACT>(a, brow + n8 * kR, g.WB, acc);
          epi.column(acc, dy, n8);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);  // this warp is done with the row
    }
    epi.tile_end(n, y, x0);
    g0 += (uint32_t)rows_total;
  }
}

}  // namespace dm
