// match_kernels.cuh -- the tiled SSD sweep shared by the fused extract kernel and
// the volume kernel.
//
// Work decomposition (one CTA = NW consumer warps + 1 TMA producer warp = a tile of TH = NW output
// rows x TW = 128 output columns of one frame pair; NW = 15 or 5 for the fused kernel, 12 for the
// volume kernel, see SweepCfg):
//   * warp w owns output row y0+w; lane l owns P=4 consecutive pixels x0+4l..x0+4l+3
//     whose C feature values (frame 1) stay in registers for the whole sweep;
//   * the sweep runs dy = 0..maxh-1.  At step dy warp w needs frame-2 row
//     y0+w+dy, columns x0..x0+TW+maxw-2, all C channels: one "row slab"
//     [C][WB] floats.  Slabs stream through a ring of NSLOT shared-memory slots
//     filled by TMA (cp.async.bulk.tensor 4-D box {WB,1,C,1}, zero fill outside
//     the frame) and signalled through one mbarrier per slot, so a CTA holds
//     only NSLOT rows of the halo at any time, whatever the window height;
//   * inside a step the dx range is cut in skewed blocks of R=8 (BlockSchedule: the last one
//     masked, or 2 wide when at most two columns are left): a thread turns 12 floats of the
//     slab (3 LDS.128) into P*R = 32 SSD partial sums per channel with packed fp32 -- 16 FADD2 +
//     16 FFMA2, or 16 FFMA2 alone in the dot form.
// The per-block epilogue is supplied by the caller (extract or volume).
#pragma once

#include "dm_common.cuh"

namespace dm {

constexpr int kP = 4;              // pixels per thread
constexpr int kTW = 32 * kP;       // 128 output columns per tile
constexpr int kR = 8;              // displacement block width
constexpr int kNB = 12;            // floats of the slab a thread reads per channel (3 x float4)
constexpr int kMaxC = 16;          // channels supported by the tiled kernels

// CTA shape of a sweep kernel: NW consumer warps (= output rows per tile, one per warp) plus
// one producer warp; PF rows in flight beyond the NW the warps are reading.
template <int NW, int PF>
struct SweepCfg {
  static constexpr int kWarps = NW;
  static constexpr int kCThreads = NW * 32;        // consumer threads
  static constexpr int kThreads = kCThreads + 32;  // + the TMA producer warp
  static constexpr int kTH = NW;
  static constexpr int kNSlot = NW + PF;
};
using ExtractCfg = SweepCfg<15, 8>;  // fused extraction: registers allow 16 warps of 128
// the same kernel for calls too small to fill the machine with 15-row tiles (one robot-size
// frame pair): 5-row tiles, 6 warps, two or three CTAs per SM
using ExtractCfgSmall = SweepCfg<5, 8>;
using VolumeCfg = SweepCfg<12, 4>;   // volume output: leaves shared memory for the store staging

// Block schedule for a window width.  Displacements are swept in "skewed" blocks: in block
// `blk` an even pixel (x even) covers dx = 8*blk + r and an odd pixel dx = 8*blk - 1 + r,
// r = 0..7, so that the pair (even pixel, r) / (odd pixel, r) reads the SAME frame-2 column
// and one packed FADD2/FFMA2 with a broadcast operand serves both.  n8 = maxw / 8 full blocks
// are followed by one tail block (index n8) of `tail_r` columns: 2 when at most two columns
// are left (maxw % 8 <= 1), else 8.  Entries outside [0, maxw) are masked by the epilogue.
struct BlockSchedule {
  int n8, tail_r;
  __host__ __device__ int per_row() const { return n8 + 1; }
};
__host__ __device__ inline BlockSchedule block_schedule(int maxw) {
  BlockSchedule s;
  s.n8 = maxw / kR;
  s.tail_r = (maxw - s.n8 * kR + 1 <= 2) ? 2 : kR;
  return s;
}

// slab width in floats (multiple of 4): lane 31 reads 124 + 8*n8 + (12 or 4)
__host__ __device__ inline int slab_width(int maxw) {
  const BlockSchedule s = block_schedule(maxw);
  return kTW - kP + s.n8 * kR + (s.tail_r == kR ? kNB : kP);
}

struct SweepGeom {
  int N, C, Cin, H1, W1, H2, W2, maxh, maxw;  // C: channels of the slab box (>= Cin, zero-filled)
  int tiles_x, tiles_y, ntiles;
  int WB;
  BlockSchedule bs;
  int slab_floats;  // floats between ring slots: C*WB rounded up to 128 bytes (+ the norm row)
  int nb_off;       // kDot: offset of the |b|^2 row inside a slot (128-byte aligned)
  int nslot;        // ring slots actually used (<= Cfg::kNSlot, >= Cfg::kTH + 2): what fits
  const float *in1;
  long long s1n, s1c, s1y;
};

// How a sweep kernel forms the SSD.
//   kFma   (a-b)^2 accumulated with FMA (packed FADD2 + FFMA2), the default for the volume
//   kExact multiply and add rounded separately: bit-exact with the non-contracting CPU path
//   kDot   |a|^2 + |b|^2 - 2 a.b: one packed FFMA2 per two terms instead of FADD2 + FFMA2.  The
//          frame-2 norms come from a small pre-pass through a second TMA box per row; the
//          absolute error is a few ulp of |a|^2 + |b|^2, so the host only allows it when the
//          largest norms keep that below the parity bar (see match_extract_impl).
enum SsdMode { kFma = 0, kExact = 1, kDot = 2 };

// kDot block: acc2[pp][j] = nb[col] + sum_k (-2 a[k]) * b[k][col] = v - |a|^2; `a2` holds -2a.  |a|^2
// is constant per pixel: minima, their order and the soft-max differences are those of this sum
// itself, so it never enters the loop (the epilogue adds it back where an SSD is reported).  The
// first FFMA2 takes the norm as a broadcast addend, so a block is C packed operations per two
// window entries, nothing else.  Same operation sequence for every (pixel, displacement):
// identical inputs give bit-identical sums, so the ties of flat image regions stay ties.
template <int CT, int JW>
__device__ __forceinline__ void dot_block2(const float2 (&a2)[CT][2], const float *bsrc, const float *nbsrc, int WB,
                                           float2 (&acc2)[2][JW]) {
  constexpr int NBF = JW == 2 ? kP : JW + kP;   // slab floats a block reads per channel: 4 (tail), 12, 20
  float nb[NBF];
  {
    const float4 *src = reinterpret_cast<const float4 *>(nbsrc);
#pragma unroll
    for (int j = 0; j < NBF / 4; ++j) {
      const float4 t = src[j];
      nb[4 * j + 0] = t.x;
      nb[4 * j + 1] = t.y;
      nb[4 * j + 2] = t.z;
      nb[4 * j + 3] = t.w;
    }
  }
#pragma unroll
  for (int k = 0; k < CT; ++k) {
    float b[NBF];
    const float4 *src = reinterpret_cast<const float4 *>(bsrc + k * WB);
#pragma unroll
    for (int j = 0; j < NBF / 4; ++j) {
      const float4 t = src[j];
      b[4 * j + 0] = t.x;
      b[4 * j + 1] = t.y;
      b[4 * j + 2] = t.z;
      b[4 * j + 3] = t.w;
    }
#pragma unroll
    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
      for (int j = 0; j < JW; ++j) {
        const float2 addend = k == 0 ? make_float2(nb[2 * pp + j], nb[2 * pp + j]) : acc2[pp][j];
        acc2[pp][j] = __ffma2_rn(a2[k][pp], make_float2(b[2 * pp + j], b[2 * pp + j]), addend);
      }
  }
}

// One channel-complete SSD block in packed fp32 (sm_100 FADD2 / FFMA2 / FMUL2):
//   acc2[pp][j].x = sum_k (a[k][2pp]   - slab[k][2pp + j])^2     even pixel, dx = 8*blk + j
//   acc2[pp][j].y = sum_k (a[k][2pp+1] - slab[k][2pp + j])^2     odd pixel,  dx = 8*blk + j - 1
// EXACT keeps multiply and add separate (bit-exact with the non-contracting CPU path).  ptxas
// 12.9 fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with explicit .rn, so the EXACT path
// packs only the subtraction and squares / accumulates with the scalar __fmul_rn / __fadd_rn,
// which are never contracted.
template <int CT, bool EXACT, int JW>
__device__ __forceinline__ void ssd_block2(const float2 (&a2)[CT][2], const float *bsrc, int WB,
                                           float2 (&acc2)[2][JW]) {
  constexpr int NBF = JW == 2 ? kP : JW + kP;   // slab floats a block reads per channel: 4 (tail), 12, 20
#pragma unroll
  for (int k = 0; k < CT; ++k) {
    float b[NBF];
    const float4 *src = reinterpret_cast<const float4 *>(bsrc + k * WB);
#pragma unroll
    for (int j = 0; j < NBF / 4; ++j) {
      const float4 t = src[j];
      b[4 * j + 0] = t.x;
      b[4 * j + 1] = t.y;
      b[4 * j + 2] = t.z;
      b[4 * j + 3] = t.w;
    }
#pragma unroll
    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
      for (int j = 0; j < JW; ++j) {
        const float nb = -b[2 * pp + j];
        const float2 d = __fadd2_rn(a2[k][pp], make_float2(nb, nb));
        if (EXACT) {
          const float sx = __fmul_rn(d.x, d.x), sy = __fmul_rn(d.y, d.y);
          acc2[pp][j] = k == 0 ? make_float2(sx, sy)
                               : make_float2(__fadd_rn(acc2[pp][j].x, sx), __fadd_rn(acc2[pp][j].y, sy));
        } else {
          acc2[pp][j] = k == 0 ? __fmul2_rn(d, d) : __ffma2_rn(d, d, acc2[pp][j]);
        }
      }
  }
}

template <int JW>
__device__ __forceinline__ void unpack_block(const float2 (&acc2)[2][JW], float (&acc)[kP][JW]) {
#pragma unroll
  for (int pp = 0; pp < 2; ++pp)
#pragma unroll
    for (int j = 0; j < JW; ++j) {
      acc[2 * pp][j] = acc2[pp][j].x;
      acc[2 * pp + 1][j] = acc2[pp][j].y;
    }
}

// Runs the whole sweep for the tiles of this CTA.  `Epi` supplies:
//   void tile_begin(n, y, x0)                       per tile, after `a` is loaded
//   void block<R>(acc[kP][R], dy, blk)              per (dy, skewed dx-block), R = 8 or 2:
//                                                   acc[p][r] is dx = 8*blk - (p & 1) + r
//   void tile_end(n, y, x0)                         per tile
//
// Pipeline: rows of all the CTA's tiles form one sequence; row `seq` lives in ring slot
// seq % NSLOT.  The producer warp (one elected lane) waits for the slot's previous tenant to
// be released (empty[slot], one arrival per consumer warp), then issues the TMA load that
// completes full[slot].  Every consumer warp walks every row of its tile in order: wait
// full, compute if the row is inside its own window (row - warp in [0, maxh)), release.
// No CTA-wide barrier: warps drift apart by up to NSLOT - TH rows, across tile borders too.
template <class Cfg, int CT, int MODE, class Epi>
__device__ __forceinline__ void run_sweep(const CUtensorMap *tmap, const CUtensorMap *tmap_nb,
                                          const SweepGeom &g, float *ring, uint64_t *bars, Epi &epi) {
  constexpr bool EXACT = MODE == kExact;
  constexpr int kWarps = Cfg::kWarps, kTH = Cfg::kTH;
  const int kNSlot = g.nslot;
  // the shuffle tells the compiler that the warp index is warp-uniform: the row tests below become uniform
  // branches and the slot arithmetic moves to the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int rows_total = kTH + g.maxh - 1;
  const uint32_t slab_bytes = (uint32_t)((g.C + (MODE == kDot ? 1 : 0)) * g.WB * sizeof(float));
  const int slab_floats = g.slab_floats;
  const int n8 = g.bs.n8;
  const bool wide_tail = g.bs.tail_r == kR;
  uint64_t *full = bars, *empty = bars + Cfg::kNSlot;

  if (threadIdx.x == 0) {
    prefetch_tmap(tmap);
    if (MODE == kDot) prefetch_tmap(tmap_nb);
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
            if (MODE == kDot)
              tma_load_4d(ring + slot * slab_floats + g.nb_off, tmap_nb, &full[slot], xt, y0 + j, 0, n);
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

    float2 a2[CT][2];
    {
      const bool rowok = y < g.H1;
      const float *src = g.in1 + (long long)n * g.s1n + (long long)(rowok ? y : 0) * g.s1y;
#pragma unroll
      for (int k = 0; k < CT; ++k) {
        float a[kP];
#pragma unroll
        for (int p = 0; p < kP; ++p)
          a[p] = (rowok && k < g.Cin && x0 + p < g.W1) ? __ldg(src + (long long)k * g.s1c + x0 + p)
                                                        : 0.0f;
        a2[k][0] = make_float2(a[0], a[1]);
        a2[k][1] = make_float2(a[2], a[3]);
      }
    }
    if (MODE == kDot) {   // a <- -2a (exact)
#pragma unroll
      for (int k = 0; k < CT; ++k) {
        a2[k][0] = make_float2(-2.0f * a2[k][0].x, -2.0f * a2[k][0].y);
        a2[k][1] = make_float2(-2.0f * a2[k][1].x, -2.0f * a2[k][1].y);
      }
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
        const int nwide = n8 + (wide_tail ? 1 : 0);
#pragma unroll 1
        for (int blk = 0; blk < nwide; ++blk) {
          float2 acc2[2][kR];
          if constexpr (MODE == kDot)
            dot_block2<CT, kR>(a2, brow + blk * kR, brow + g.nb_off + blk * kR, g.WB, acc2);
          else
            ssd_block2<CT, EXACT, kR>(a2, brow + blk * kR, g.WB, acc2);
          float acc[kP][kR];
          unpack_block<kR>(acc2, acc);
          epi.template block<kR>(acc, dy, blk);
        }
        if (!wide_tail) {
          float2 acc2[2][2];
          if constexpr (MODE == kDot)
            dot_block2<CT, 2>(a2, brow + n8 * kR, brow + g.nb_off + n8 * kR, g.WB, acc2);
          else
            ssd_block2<CT, EXACT, 2>(a2, brow + n8 * kR, g.WB, acc2);
          float acc[kP][2];
          unpack_block<2>(acc2, acc);
          epi.template block<2>(acc, dy, n8);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);  // this warp is done with the row
    }
    if constexpr (MODE == kDot) epi.template tile_rescore<CT>(a2, n, y, x0);  // a2 = -2a here
    epi.tile_end(n, y, x0);
    g0 += (uint32_t)rows_total;
  }
}

}  // namespace dm
