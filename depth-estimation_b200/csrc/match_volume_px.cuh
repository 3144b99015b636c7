// match_volume_px.cuh -- volume mode through whole-pixel streams ("strip kernel").
//
// The volume [N][H1][W1][maxh*maxw] (the nn.SpatialMatching output layout,
// /root/reference/opticalflow_model.lua:61-66 builds the module, :153-169 reads the volume) gives
// every pixel one contiguous stream of K = maxh*maxw floats, and the streams of the pixels of an
// image row follow each other: 16 pixels of a row are ONE contiguous run of 16*K floats, 16-byte
// aligned when the first pixel index is a multiple of 4.  The tiled sweep (VolumeEpi) produces 8
// entries of 128 different streams per block and has to push them out sector by sector through
// the LSU (128 store wavefronts per block: the kernel is L1TEX-bound at 37 % of HBM).  This
// kernel turns the loops around:
//
//   * a CTA walks DOWN a strip of 16 pixel columns; a step = one output row of the strip = the
//     complete streams of 16 pixels, staged in shared memory in exactly the global layout and
//     written by the copy engine of the SM with ONE bulk copy per step
//     (cp.async.bulk.global.shared::cta): no store instruction, no partial sector, ever;
//   * frame-2 rows live in a ring of maxh + 3 slots ([C][48 columns] each, one TMA box per step),
//     so a step reads 1.9 KB of new input for 70 KB of output;
//   * the work of a step is cut into warp items: 4 window rows x 2 displacement blocks x 4 pixel
//     quads = 32 lanes, each lane the usual 4-pixel x 8-displacement register block
//     (dot_block2 / ssd_block2 of match_kernels.cuh).  The lane order makes both shared-memory
//     sides conflict-free: a quarter-warp reads blocks b and b+2 of one slab row (128 contiguous
//     bytes per LDS.128 wavefront), and the 32 lanes of a scalar staging store hit 32 different
//     banks (bank = 4*quad + row + 16*(which block) when K % 32 == 1 and maxw % 32 == 1, the
//     33x33 window);
//   * items are dealt round-robin over the compute warps ACROSS steps (measured: a fixed, per-warp
//     balanced assignment is 28 % slower -- what has to balance is the load per SM sub-partition,
//     and the rotation does that for free); the two staging buffers are handed over by mbarriers
//     (no CTA-wide barrier): a warp that finishes its share of step s starts on step s+1 while the
//     copy engine drains step s-1.
//
// Same arithmetic as the tiled kernels (the same block functions in the same order per entry), so
// the two kernels agree bit for bit; tests/test_gpu_parity.py compares both with the oracle.
#pragma once

#include "match_kernels.cuh"

namespace dm {

constexpr int kPxW = 16;                           // pixels of a strip (4 quads of kP)
constexpr int kPxMaxWarps = 8;                     // compute warps (6 / 10 / 12 / 14 measured: 0.39 / 0.31 / 0.30 / 0.30 ms against 0.29)
constexpr int kPxMaxThreads = (kPxMaxWarps + 2) * 32;   // + TMA loader warp + store warp
constexpr int kPxAhead = 3;                        // ring slots beyond the window height
constexpr int kPxMaxSlot = 72;                     // barrier array size
constexpr int kPxKRegs = 37;                       // fused soft-max: window entries per lane held in registers (K <= 1184)
constexpr int kPxARing = 4;                        // steps of frame-1 values in flight (= kPxAhead + 1)

struct PxGeom {
  int WBs;      // slab columns a step reads: 12 + 8*n8 + (wide tail ? 12 : 4)
  int pitch;    // floats between ring slots (multiple of 32)
  int nb_off;   // kDot: offset of the |b|^2 row inside a slot
  int nslot;    // maxh + kPxAhead
  int strips, band, nbands, units;
  int nwide;    // full-width blocks per window row (n8 + wide tail)
  int nbp;      // block pairs per window row
  int ndg;      // groups of four window rows
  int nfull;    // ndg * nbp: items of full-width blocks per step
  int items;    // nfull + items of narrow tail blocks (8 window rows each)
  int ncw;      // compute warps
  int rot;      // rotation of the item deal per step (>= items, coprime to ncw)
  int nrow;     // 1: the maxh % 4 (1 or 2) rows left over are one item of their own (four blocks wide)
};

__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// the two blocks of pair j of a window row: (b, b + 2) inside complete groups of four blocks, what
// is left over in a last incomplete group as it comes.  -1 = no block.
__device__ __forceinline__ void px_block_pair(int j, int nwide, int *b0, int *b1) {
  const int g4 = j >> 1, l = j & 1, base = 4 * g4, rem = nwide - base;
  if (rem >= 4) {
    *b0 = base + l;
    *b1 = base + l + 2;
  } else if (rem == 3) {
    *b0 = base + l;
    *b1 = l == 0 ? base + 2 : -1;
  } else if (rem == 2) {
    *b0 = base;
    *b1 = base + 1;
  } else {
    *b0 = base;
    *b1 = -1;
  }
}

// soft-max of two pixel streams of K floats in shared memory, in place, by one warp: values in registers
// (K / 32 per lane and pixel; NF full chunks of 32 without a bounds test, the rest guarded), minimum and sum
// by four partial chains + shuffles, one read and one write.  7 instructions per entry.
template <int NF>
__device__ __forceinline__ void px_softmax_two(float *base0, float *base1, int K, int lane) {
  constexpr int NT = NF == 0 ? kPxKRegs : 3;          // guarded chunks after the NF full ones
  float v0[NF + NT], v1[NF + NT];
#pragma unroll
  for (int j = 0; j < NF + NT; ++j) {
    const int idx = lane + 32 * j;
    const bool in = j < NF || idx < K;
    v0[j] = in ? base0[idx] : 3.0e38f;
    v1[j] = in ? base1[idx] : 3.0e38f;
  }
  float m0[4] = {3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f}, m1[4] = {3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f};
#pragma unroll
  for (int j = 0; j < NF + NT; ++j) {
    m0[j & 3] = fminf(m0[j & 3], v0[j]);
    m1[j & 3] = fminf(m1[j & 3], v1[j]);
  }
  float a0 = fminf(fminf(m0[0], m0[1]), fminf(m0[2], m0[3])), a1 = fminf(fminf(m1[0], m1[1]), fminf(m1[2], m1[3]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 = fminf(a0, __shfl_xor_sync(0xffffffffu, a0, o));
    a1 = fminf(a1, __shfl_xor_sync(0xffffffffu, a1, o));
  }
  const float mL0 = a0 * kLog2e, mL1 = a1 * kLog2e;
  float s0[4] = {0.0f, 0.0f, 0.0f, 0.0f}, s1[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int j = 0; j < NF + NT; ++j) {
    // a padding value (3e38) gives ex2(-huge) = 0: no test needed here
    v0[j] = ex2_approx(fmaf(v0[j], -kLog2e, mL0));
    v1[j] = ex2_approx(fmaf(v1[j], -kLog2e, mL1));
    s0[j & 3] += v0[j];
    s1[j & 3] += v1[j];
  }
  float t0 = (s0[0] + s0[1]) + (s0[2] + s0[3]), t1 = (s1[0] + s1[1]) + (s1[2] + s1[3]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    t0 += __shfl_xor_sync(0xffffffffu, t0, o);
    t1 += __shfl_xor_sync(0xffffffffu, t1, o);
  }
  const float i0 = 1.0f / t0, i1 = 1.0f / t1;
#pragma unroll
  for (int j = 0; j < NF + NT; ++j) {
    const int idx = lane + 32 * j;
    if (j < NF || idx < K) {
      base0[idx] = v0[j] * i0;
      base1[idx] = v1[j] * i1;
    }
  }
}

template <int CT, int MODE, bool FSM>
__global__ void __launch_bounds__(kPxMaxThreads, 1)
match_volume_px_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_nb,
                       const VolumeParams P, const PxGeom X) {
  if (P.stats) {  // twin launch: the norm bound picks the dot or the difference form
    const bool dot_ok = __uint_as_float(P.stats[0]) + __uint_as_float(P.stats[1]) <= P.dot_limit;
    if (dot_ok != (MODE == kDot)) return;
  }
  constexpr bool EXACT = MODE == kExact;
  // FSM: soft-max volume without a statistics sweep -- the staging buffer holds the complete streams of the
  // step's pixels, so minimum, sum and exp(min - v) / sum are taken there before the copy engine gets the
  // buffer.  A template parameter: as a run-time flag its branches cost the plain SSD volume 7 %.
  constexpr bool fused_softmax = FSM;
  const SweepGeom &g = P.g;
  const int maxh = g.maxh, maxw = g.maxw, K = maxh * maxw;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *ring = reinterpret_cast<float *>(smem_raw);
  float *stage = ring + (size_t)X.nslot * X.pitch;                 // [2][kPxW * K], global layout
  // frame-1 side of a step: [CT][16] values (times -2 in the dot form), then 16 x min*log2e, 16 x 1/sum
  constexpr int kAFloats = CT * kPxW + 2 * kPxW;
  float *aring = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(stage + 2 * (size_t)kPxW * K) + 15) & ~uintptr_t(15));
  uint64_t *full = reinterpret_cast<uint64_t *>(aring + kPxARing * kAFloats);
  uint64_t *empty = full + kPxMaxSlot, *sdone = empty + kPxMaxSlot, *sfree = sdone + 2, *afull = sfree + 2, *sraw = afull + kPxARing;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  constexpr int kPxWarps = kPxMaxWarps;   // compile-time: the item rotation and barrier counts fold

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap);
    if (MODE == kDot) prefetch_tmap(&tmap_nb);
    for (int s = 0; s < X.nslot; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kPxWarps);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sdone[b], kPxWarps);
      mbar_init(&sfree[b], 1);
    }
    for (int b = 0; b < kPxARing; ++b) mbar_init(&afull[b], 1);
    for (int b = 0; b < 2; ++b) mbar_init(&sraw[b], kPxWarps);
    fence_mbar_init();
  }
  __syncthreads();

  const uint32_t slab_bytes = (uint32_t)((CT + (MODE == kDot ? 1 : 0)) * X.WBs * sizeof(float));
  uint32_t rbase = 0, sbase = 0;   // rows / steps of the units before this one (all roles count alike)
  for (int unit = blockIdx.x; unit < X.units; unit += gridDim.x) {
    int u = unit;
    const int yb = u % X.nbands; u /= X.nbands;
    const int xs = u % X.strips;
    const int n = u / X.strips;
    const int y0 = yb * X.band, y1 = min(g.H1, y0 + X.band);
    const int steps = y1 - y0, rows = steps + maxh - 1;
    const int x0 = xs * kPxW;

    if (warp == kPxWarps) {
      // ---------------- loader: frame-2 row y0 + j -> slot (rbase + j) % nslot (TMA, lane 0), and the
      // frame-1 values of the step that needs this row first, t = j - (maxh - 1), into the a-ring (all
      // lanes, plain loads: frame 1 may have any strides).  The wait on the row's slot also covers the
      // a-ring: the previous tenant of the slot is row j - nslot, released after step t - kPxARing.
      for (int j = 0; j < rows; ++j) {
        const uint32_t seq = rbase + (uint32_t)j;
        const int slot = (int)(seq % (uint32_t)X.nslot);
        const uint32_t inst = seq / (uint32_t)X.nslot;
        // frame-1 values first: the loads do not depend on the slot, their latency passes under the wait
        const int t = j - (maxh - 1);
        float v[(kAFloats + 31) / 32];
        if (t >= 0) {
          const int y = y0 + t;
          const float *src = g.in1 + (long long)n * g.s1n + (long long)y * g.s1y;
          const size_t orow = ((size_t)n * g.H1 + y) * g.W1;
#pragma unroll
          for (int e = 0; e < (kAFloats + 31) / 32; ++e) {
            const int idx = lane + 32 * e, k = idx / kPxW, x = x0 + (idx % kPxW);
            float t0 = 0.0f;
            if (idx < CT * kPxW) {
              if (k < g.Cin && x < g.W1) t0 = __ldg(src + (long long)k * g.s1c + x) * (MODE == kDot ? -2.0f : 1.0f);
            } else if (idx < kAFloats) {
              const bool second = idx >= CT * kPxW + kPxW;
              t0 = second ? 1.0f : 0.0f;
              if (P.mode == DM_VOLUME_NEG_SOFTMAX && !fused_softmax && x < g.W1)
                t0 = second ? __ldg(P.vinv + orow + x) : __ldg(P.vmin + orow + x) * kLog2e;
            }
            v[e] = t0;
          }
        }
        if (lane == 0) {
          if (inst > 0) mbar_wait_backoff(&empty[slot], (inst - 1) & 1u);
          mbar_arrive_expect_tx(&full[slot], slab_bytes);
          tma_load_4d(ring + (size_t)slot * X.pitch, &tmap, &full[slot], x0, y0 + j, 0, n);
          if (MODE == kDot) tma_load_4d(ring + (size_t)slot * X.pitch + X.nb_off, &tmap_nb, &full[slot], x0, y0 + j, 0, n);
        }
        __syncwarp();
        if (t >= 0) {
          const uint32_t sg = sbase + (uint32_t)t;
          float *dst = aring + (sg % kPxARing) * kAFloats;
#pragma unroll
          for (int e = 0; e < (kAFloats + 31) / 32; ++e)
            if (lane + 32 * e < kAFloats) dst[lane + 32 * e] = v[e];
          __syncwarp();
          if (lane == 0) mbar_arrive(&afull[sg % kPxARing]);
        }
      }
    } else if (warp == kPxWarps + 1) {
      // ---------------- store warp: one bulk copy per step, the streams of the strip's pixels
      if (lane == 0) {
        const int npx = min(kPxW, g.W1 - x0);
        const uint32_t bytes = (uint32_t)((size_t)npx * K * sizeof(float));
        for (int s = 0; s < steps; ++s) {
          const uint32_t sg = sbase + (uint32_t)s, b = sg & 1u, use = sg >> 1;
          mbar_wait_backoff(&sdone[b], use & 1u);
          float *dst = P.out + (((size_t)n * g.H1 + (y0 + s)) * g.W1 + x0) * (size_t)K;
          if (!(P.debug & 1)) {
            bulk_store(dst, stage + (size_t)b * kPxW * K, bytes);
            bulk_commit();
            bulk_wait_read0();               // the copy engine has read the buffer
          }
          mbar_arrive(&sfree[b]);
        }
      }
    } else {
      // ---------------- compute warps
      const int q = lane & 3, sl = lane >> 2, ddy = sl >> 1, hb = sl & 1;
      for (int s = 0; s < steps; ++s) {
        const uint32_t sg = sbase + (uint32_t)s, b = sg & 1u, use = sg >> 1;
        const int y = y0 + s;
        // every warp observes every row's arrival once, in order
        if (s == 0) {
          for (int j = 0; j < maxh - 1; ++j) {
            const uint32_t seq = rbase + (uint32_t)j;
            mbar_wait(&full[seq % (uint32_t)X.nslot], (seq / (uint32_t)X.nslot) & 1u);
          }
        }
        {
          const uint32_t seq = rbase + (uint32_t)(s + maxh - 1);
          mbar_wait(&full[seq % (uint32_t)X.nslot], (seq / (uint32_t)X.nslot) & 1u);
        }
        // the copy engine is done with this step's staging buffer.  EVERY warp waits, also one without
        // items in this step: its arrival on sdone must not run a phase ahead of the slowest warp
        if (use > 0) mbar_wait(&sfree[b], (use - 1) & 1u);
        const int rs = (int)((rbase + (uint32_t)s) % (uint32_t)X.nslot);   // slot of window row 0
        // this warp's items of the step: item i of step sg goes to warp (sg * rot + i) % warps.  rot >= items
        // is coprime to the warp count, so every warp meets every residue in turn: with rot = 22 and 8
        // warps the even warps kept the heavier residues and two SM sub-partitions ran 12 % longer
        const int i0 = (int)(((uint32_t)warp + kPxWarps - (uint32_t)(((unsigned long long)sg * (unsigned)X.rot) % kPxWarps)) % kPxWarps);
        if (i0 < X.items) {
          float2 a2[CT][2];
          float mL[kP], inv[kP];
          {
            mbar_wait(&afull[sg % kPxARing], (sg / kPxARing) & 1u);
            const float4 *src = reinterpret_cast<const float4 *>(aring + (sg % kPxARing) * kAFloats) + q;
#pragma unroll
            for (int k = 0; k < CT; ++k) {
              const float4 a = src[k * (kPxW / 4)];
              a2[k][0] = make_float2(a.x, a.y);
              a2[k][1] = make_float2(a.z, a.w);
            }
            const float4 m4 = src[CT * (kPxW / 4)], i4 = src[(CT + 1) * (kPxW / 4)];
            mL[0] = m4.x, mL[1] = m4.y, mL[2] = m4.z, mL[3] = m4.w;
            inv[0] = i4.x, inv[1] = i4.y, inv[2] = i4.z, inv[3] = i4.w;
          }
          float *stg = stage + (size_t)b * kPxW * K + (size_t)(q * kP) * K;
          for (int i = i0; i < X.items; i += kPxWarps) {
            if (i < X.nfull + X.nrow) {
              int dy, blk;
              if (i < X.nfull) {   // four window rows x a pair of blocks
                const int dg = i / X.nbp, j = i - dg * X.nbp;
                int b0, b1;
                px_block_pair(j, X.nwide, &b0, &b1);
                dy = 4 * dg + ddy;
                blk = hb ? b1 : b0;
              } else {             // the one or two window rows left over by the groups of four x four blocks
                dy = 4 * X.ndg + (sl >> 2);
                blk = (sl & 1) * 2 + ((sl >> 1) & 1);
              }
              const bool active = dy < maxh && blk >= 0;
              int slot = rs + (active ? dy : 0);
              slot -= slot >= X.nslot ? X.nslot : 0;
              const int dx0 = (active ? blk : 0) * kR;     // first displacement of the block (even pixels)
              const float *brow = ring + (size_t)slot * X.pitch + q * kP + dx0;
              float2 acc2[2][kR];
              if constexpr (MODE == kDot)
                dot_block2<CT, kR>(a2, brow, brow + X.nb_off, X.WBs, acc2);
              else
                ssd_block2<CT, EXACT, kR>(a2, brow, X.WBs, acc2);
              float acc[kP][kR];
              unpack_block<kR>(acc2, acc);
              if (P.mode == DM_VOLUME_NEG_SOFTMAX && !fused_softmax) {
#pragma unroll
                for (int p = 0; p < kP; ++p)
#pragma unroll
                  for (int r = 0; r < kR; ++r) acc[p][r] = ex2_approx(fmaf(acc[p][r], -kLog2e, mL[p])) * inv[p];
              }
              if (active) {
#pragma unroll
                for (int p = 0; p < kP; ++p) {
                  float *dst = stg + p * K + dy * maxw + dx0 - (p & 1);
#pragma unroll
                  for (int r = 0; r < kR; ++r) {
                    const int dx = dx0 - (p & 1) + r;
                    if (dx >= 0 && dx < maxw) dst[r] = acc[p][r];
                  }
                }
              }
            } else {
              // narrow tail blocks (2 columns): 8 window rows x 4 quads
              const int dy = 8 * (i - X.nfull - X.nrow) + sl, blk = g.bs.n8;
              const bool active = dy < maxh;
              int slot = rs + (active ? dy : 0);
              slot -= slot >= X.nslot ? X.nslot : 0;
              const float *brow = ring + (size_t)slot * X.pitch + q * kP + blk * kR;
              float2 acc2[2][2];
              if constexpr (MODE == kDot)
                dot_block2<CT, 2>(a2, brow, brow + X.nb_off, X.WBs, acc2);
              else
                ssd_block2<CT, EXACT, 2>(a2, brow, X.WBs, acc2);
              float acc[kP][2];
              unpack_block<2>(acc2, acc);
              if (P.mode == DM_VOLUME_NEG_SOFTMAX && !fused_softmax) {
#pragma unroll
                for (int p = 0; p < kP; ++p)
#pragma unroll
                  for (int r = 0; r < 2; ++r) acc[p][r] = ex2_approx(fmaf(acc[p][r], -kLog2e, mL[p])) * inv[p];
              }
              if (active) {
#pragma unroll
                for (int p = 0; p < kP; ++p) {
                  float *dst = stg + p * K + dy * maxw + blk * kR - (p & 1);
#pragma unroll
                  for (int r = 0; r < 2; ++r) {
                    const int dx = blk * kR - (p & 1) + r;
                    if (dx >= 0 && dx < maxw) dst[r] = acc[p][r];
                  }
                }
              }
            }
          }
          fence_proxy_async();   // the staging stores of this thread -> visible to the copy engine
        }
        __syncwarp();
        if constexpr (FSM) {
          // soft-max volume without a statistics sweep: once every warp's entries of the step are in the buffer
          // (sraw), warp w turns the streams of pixels 2w, 2w + 1 into probabilities in place.  Measured
          // alternatives: three extra warps doing this under the next step's arithmetic are starved by the
          // compute warps (0.61 ms); with the barrier the eight compute warps do it at full width
          if (lane == 0) mbar_arrive(&sraw[b]);
          mbar_wait(&sraw[b], use & 1u);
          const int npx = min(kPxW, g.W1 - x0);
          for (int px = 2 * warp; px < npx; px += 2 * kPxWarps) {
            float *base0 = stage + (size_t)b * kPxW * K + (size_t)px * K;
            float *base1 = px + 1 < npx ? base0 + K : base0;      // an odd last pixel is simply done twice
            if (K >= 34 * 32 && K <= 37 * 32)
              px_softmax_two<34>(base0, base1, K, lane);          // 33x33 and neighbours: 34 chunks unguarded
            else
              px_softmax_two<0>(base0, base1, K, lane);
          }
          fence_proxy_async();
          __syncwarp();
        }
        if (lane == 0) {
          mbar_arrive(&sdone[b]);
          // window row 0 of this step is not read again; the last step releases the rest
          mbar_arrive(&empty[rs]);
          if (s == steps - 1)
            for (int j = 1; j < maxh; ++j) {
              int slot = rs + j;
              slot -= slot >= X.nslot ? X.nslot : 0;
              mbar_arrive(&empty[slot]);
            }
        }
      }
    }
    rbase += (uint32_t)rows;
    sbase += (uint32_t)steps;
  }
  if (warp == kPxWarps + 1 && lane == 0) bulk_wait0();   // global writes complete before the CTA retires
}

}  // namespace dm
