// filter_tc.cu -- the feature extractor's convolution layers on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM), SURVEY.md 8f row 3 / VERDICT r1 item 7.
//
// getFilter (opticalflow_model.lua:45-79): out[o][y][x] = b[o] + sum_{c,ky,kx} in[c][y+ky][x+kx] * w[o][c][ky][kx]
// (nn.SpatialConvolution / SpatialConvolutionMap: unconnected (c,o) pairs carry zero weights).
//
// As a GEMM.  The im2col operand A[x][(c,ky,kx)] is a Hankel matrix (a window advancing by 4 bytes)
// that no UMMA operand layout addresses in place.  Moving kx to the N side removes the expansion:
//     P[x'][(o,kx)] = sum_{(c,ky)} in[c][y+ky][x'] * w[o][c][ky][kx]        one plain GEMM per output row
//     out[o][y][x]  = b[o] + sum_kx P[x+kx][(o,kx)]                          a shifted-diagonal sum
// with M = 128 input columns x' (TMEM lanes), N = outputs-per-group * kW (padded to 16), K = n_in * kH
// (kH padded to 8 per input plane with zero weights).  Every product of P is used by exactly one
// output, so apart from the kW-1 columns at the tile edge no MAC is wasted.
//
// Precision.  The features feed an exact SSD argmin; the parity bar is 1e-4 relative on fp32
// features, a single tf32 pass (2^-11) misses it.  Three-term split: x = hi + lo with hi = the
// tf32 truncation the hardware applies to an fp32 operand anyway, lo = x - hi (exact in fp32):
//     A*B ~= A_hi*B_hi + A_hi*B_lo + A_lo*B_hi          (the dropped lo*lo term is 2^-22 relative)
// Three MMAs per K step into the same TMEM accumulator; measured on B200 against a double
// reference (scripts/umma_probe.cu): 6e-7 of sum |a||b|.
//
// Operands.  B (weights, hi and lo, K-major no-swizzle core matrices: element (n,k) at (k%4) + 4*(n%8)
// + SBO*(n/8) + LBO*(k/4)) is packed by the host at dm_filter_create and stays resident in shared
// memory.  A goes registers -> TMEM: thread m (= TMEM lane m = input column x0+m) writes hi / lo of
// its column of the window with tcgen05.st, and the MMA reads only B from shared memory (with both
// operands in shared memory a 128 x 80 x 8 tf32 MMA would need 115 bytes per clock of the 128 the
// SM has).
//
// The window slides: the A operand of output row y+1 is the one of row y shifted by one kernel row.
// Re-staging 128 x K values per output row through the CUDA cores made the gather, not the MMA, the
// bound (first versions: 0.82 and 0.44 ms per 640x360 pair against 0.26 for the CUDA-core kernel).
// So A is a RING that is never shifted: a unit walks the output rows of ONE residue class
// y = y0 + b + 4t (b = 0..3 fixed per unit), and the A column of plane c, slot j holds the image
// row R with (R - y0 - b) mod KP == j.  Advancing t by one replaces exactly one 16-byte chunk of
// four slots per plane (four new image rows), and rotates the pairing with the kernel rows by
// one chunk: slot j meets ky = (j - 4t) mod KP.  K-major B keeps k in 16-byte chunks of four, so
// that rotation is a different START CHUNK of the same resident weights -- the descriptor of the
// K step over slots [8s, 8s+8) of plane c points at chunk (2s - t) mod (KP/4) of that plane (the
// host appends a copy of chunk 0 after the last chunk so that the pair never wraps).  Per output
// row a gather thread now loads 4 x planes/2 values and issues 2 x planes/2 tcgen05.st.x4.
//
// Schedule.  One CTA per SM, persistent over units (output group, image, column tile, row band,
// residue class), warp-specialised, three roles connected by mbarriers only:
//   * 8 gather warps (two sets of four: set h owns half h of the input planes = half h of K): the
//     four new image rows of a step are prefetched one step ahead (global -> registers), then
//     written hi / lo into the TMEM ring: a_full[h].
//   * 1 issuing thread: per step and half, waits a_full[h], issues 3 x K/16 MMAs into accumulator
//     buffer t & 1, commits to a_empty[h] (the set may overwrite its chunk while the other half's
//     MMAs still run), and after both halves commits to d_full.
//   * 8 epilogue warps (two per TMEM quadrant, splitting the output planes): tcgen05.ld of the
//     accumulator row into shared memory, d_empty, then the diagonal sum + bias + tanh +
//     coalesced store -- under the next step's MMAs.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "dm_common.cuh"
#include "filter_tc.cuh"

namespace dm {

constexpr int kTcM = 128;          // input columns per tile = TMEM lanes
constexpr int kTcColD = 0;         // TMEM columns: two accumulator buffers of 128
constexpr int kTcColAhi = 256;     //               A hi (K_tot <= 128 columns)
constexpr int kTcColAlo = 384;     //               A lo
constexpr int kTcMaxN = 128, kTcMaxK = 128;
constexpr int kTcGatherWarps = 8, kTcEpiWarps = 8;
constexpr int kTcThreads = (kTcGatherWarps + kTcEpiWarps + 1) * 32;

struct TcArgs {
  const float *in;
  float *out;
  const float *B;      // [ngroups][2][Npad * Ktot] packed hi / lo
  const float *bias;
  int n_img, n_in, n_out, kh, kw, tanh_after;
  int h, w, hout, wout, pad_t, pad_l;
  int KP, Ktot, KtotB, G, ngroups, Npad;
  int twv, col_tiles, band, nbands;
  int units;
  long long *prof;   // tuning (option volume_debug = 9): cycles of CTA 0 per role: [0] gather wait a_empty, [1] gather work,
                     // [2] issuer wait d_empty, [3] issuer wait a_full, [4] issuer issue, [5] epilogue wait d_full, [6] epilogue work, [7] steps
};

__device__ __forceinline__ uint64_t tc_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell); no swizzle
  return d;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// executed by a whole (converged) warp; one elected lane commits
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tc_named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(kTcThreads, 1) conv_tc_kernel(const TcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // the shuffle makes the warp index warp-uniform for the compiler: role code then runs on the uniform
  // datapath, which is what lets the issuer's tcgen05.mma operands stay in uniform registers
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int nk = a.Npad * a.KtotB;                      // floats of one B term
  const int NP1 = a.Npad + 1;
  float *sB = reinterpret_cast<float *>(smem);          // [2][nk]
  float *sP = sB + 2 * nk;                              // [128][Npad + 1]
  uint64_t *bars = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(sP + kTcM * NP1) + 15) & ~uintptr_t(15));
  uint64_t *a_full = bars, *a_empty = bars + 2, *d_full = bars + 4, *d_empty = bars + 6;
  uint32_t *tptr = reinterpret_cast<uint32_t *>(bars + 8);

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 128);
      mbar_init(&a_empty[i], 1);
      mbar_init(&d_full[i], 1);
      mbar_init(&d_empty[i], kTcEpiWarps * 32);
    }
    fence_mbar_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tptr)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's quadrant of lanes
  const int m = (warp & 3) * 32 + lane;                              // TMEM lane = input column of the tile
  // halves of K: input planes [0, c_split) and [c_split, n_in)
  const int c_split = (a.n_in + 1) / 2;

  // roles
  const bool is_gather = warp < kTcGatherWarps;
  const bool is_epi = warp >= kTcGatherWarps && warp < kTcGatherWarps + kTcEpiWarps;
  const bool is_issuer = warp == kTcGatherWarps + kTcEpiWarps;
  const int gset = warp >> 2;                        // gather: which half
  const int esub = (warp - kTcGatherWarps) >> 2;     // epilogue: which share of the output planes

  // running counts of rows handled (all roles walk the same sequence): parities of the barriers
  uint32_t rows_done = 0;
  int loaded_group = -1;
  for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
    // unit -> (group, image, column tile, row band); the group is the slowest index so that a CTA
    // re-loads the weights as rarely as possible
    int u = unit;
    const int cls = u & 3; u >>= 2;            // residue class of the output rows: y = y_begin + cls + 4t
    const int rb = u % a.nbands; u /= a.nbands;
    const int tx = u % a.col_tiles; u /= a.col_tiles;
    const int n = u % a.n_img; u /= a.n_img;
    const int g = u;
    const int y_begin = rb * a.band + cls, y_end = min(a.hout, rb * a.band + a.band);
    const int x0 = tx * a.twv;                 // first output column of the tile = first window column
    const int gout0 = g * a.G, gcount = min(a.G, a.n_out - gout0);
    const int nrows = y_end > y_begin ? (y_end - y_begin + 3) / 4 : 0;   // steps of this unit
    const int CP = a.KP / 4;                   // 16-byte chunks (four kernel rows) per input plane

    if (g != loaded_group) {                   // (CTA-uniform) new weights: everybody stops at the unit border
      // the issuer's last MMAs read sB: they are complete once the epilogue saw d_full of the last row,
      // and the epilogue warps pass this barrier only after that
      __syncthreads();
      const float4 *src = reinterpret_cast<const float4 *>(a.B + (size_t)g * 2 * nk);
      float4 *dst = reinterpret_cast<float4 *>(sB);
      for (int i = tid; i < 2 * nk / 4; i += kTcThreads) dst[i] = __ldg(src + i);
      loaded_group = g;
      fence_proxy_async();                     // generic-proxy writes -> visible to the tensor core
      __syncthreads();
    }

    if (is_gather) {
      // ---------------- A producers: the TMEM ring, one chunk of four image rows per plane and step
      const int c0 = gset == 0 ? 0 : c_split, c1 = gset == 0 ? c_split : a.n_in;
      const int xin = x0 + m - a.pad_l;        // this thread's input column
      const bool xok = xin >= 0 && xin < a.w;
      const float *img = a.in + (size_t)n * a.n_in * a.h * a.w;
      auto fetch = [&](int r, int c) -> float {  // window row r (output-row coordinates) of plane c
        const int rin = r - a.pad_t;
        return (xok && rin >= 0 && rin < a.h) ? __ldg(img + ((size_t)c * a.h + rin) * a.w + xin) : 0.0f;
      };
      auto put4 = [&](int c, int chunk, const float (&v)[4]) {  // hi / lo of four slots of plane c
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float t = __uint_as_float(__float_as_uint(v[j]) & 0xffffe000u);   // what the MMA reads of v
          hi[j] = __float_as_uint(v[j]);
          lo[j] = __float_as_uint(v[j] - t);
        }
        const uint32_t cc = (uint32_t)(c * a.KP + 4 * chunk);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tlane + kTcColAhi + cc), "r"(hi[0]),
                     "r"(hi[1]), "r"(hi[2]), "r"(hi[3])
                     : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tlane + kTcColAlo + cc), "r"(lo[0]),
                     "r"(lo[1]), "r"(lo[2]), "r"(lo[3])
                     : "memory");
      };
      constexpr int kPl = 4;                   // planes per set (the plan guarantees <= 4); unrolled: registers, no local memory
      const size_t plane = (size_t)a.h * a.w;
      const float *col0 = img + (size_t)c0 * plane + (xok ? xin : 0);
      const int npl = c1 - c0;
      auto fetch4 = [&](int r0, float (&v)[kPl][4]) {   // window rows r0 .. r0 + 3 of this set's planes
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rin = r0 + j - a.pad_t;
          const bool ok = xok && rin >= 0 && rin < a.h;
          const float *p0 = col0 + (size_t)(ok ? rin : 0) * a.w;
#pragma unroll
          for (int c = 0; c < kPl; ++c) v[c][j] = (ok && c < npl) ? __ldg(p0 + c * plane) : 0.0f;
        }
      };
      float nxt[kPl][4];                       // the chunk of the NEXT step, prefetched
      for (int i = 0; i < nrows; ++i) {
        const uint32_t it = rows_done + (uint32_t)i;
        const int y = y_begin + 4 * i;
        // the MMAs of the previous step that read this half are done
        const long long tg0 = clock64();
        mbar_wait_backoff(&a_empty[gset], (it & 1u) ^ 1u);   // suspended, not polling: a spinning warp steals the issuer's slots
        tc_fence_after();
        const long long tg1 = clock64();
        if (i == 0) {
          // first step of the unit: the whole window, rows y .. y + KP - 1 -> slots 0 .. KP - 1
          for (int q = 0; q < CP; ++q) {
            fetch4(y + 4 * q, nxt);
#pragma unroll
            for (int c = 0; c < kPl; ++c)
              if (c < npl) put4(c0 + c, q, nxt[c]);
          }
        } else {
          // rows y + KP - 4 .. y + KP - 1 replace the chunk that held rows y - 4 .. y - 1
          const int chunk = (i - 1) % CP;
#pragma unroll
          for (int c = 0; c < kPl; ++c)
            if (c < npl) put4(c0 + c, chunk, nxt[c]);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(&a_full[gset]);
        if (i + 1 < nrows) fetch4(y + a.KP, nxt);   // prefetch the next step's four rows under this step's MMAs
        if (a.prof && blockIdx.x == 0 && tid == 0) {
          a.prof[0] += tg1 - tg0;
          a.prof[1] += clock64() - tg1;
          a.prof[7] += 1;
        }
      }
    } else if (is_issuer) {
      // ---------------- MMA issuer: A_hi*B_hi + A_hi*B_lo + A_lo*B_hi per half
      // The WHOLE warp walks the loop with warp-uniform operands and the MMA is predicated on elect.sync
      // inside the asm block: ptxas keeps descriptors and TMEM addresses in uniform registers and emits
      // UMOV/UIADD3 + UTCHMMA.  Issued from one divergent lane (`if (lane == 0)`), every MMA paid an
      // ELECT + 6 x R2UR.BROADCAST waterfall: 188 cycles per MMA against 59 (SS) / 83 (TS) this way
      // (scripts/umma_rate.cu, profiles/r02_umma_probes.txt).
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.Npad >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
      const uint32_t lbo = 128, sbo = 32u * (uint32_t)a.KtotB;  // bytes: k-chunks adjacent, 8-row groups KtotB/4 chunks apart
      const uint64_t d_hi = tc_desc(smem_u32(sB), lbo, sbo), d_lo = tc_desc(smem_u32(sB + nk), lbo, sbo);
      const int hcp = CP / 2;
      for (int i = 0; i < nrows; ++i) {
        const uint32_t it = rows_done + (uint32_t)i;
        const uint32_t buf = it & 1u, use = it >> 1;
        const long long ti0 = clock64();
        mbar_wait(&d_empty[buf], (use & 1u) ^ 1u);   // the epilogue has read this accumulator buffer
        long long twait = 0, ti1 = clock64();
        const uint32_t td = tmem + kTcColD + buf * 128u;
        // slot j meets kernel row (j - 4 i) mod KP: chunk q = (2 sp - rot) mod CP of the packed weights
        const int rot = i % CP;
        uint32_t acc = 0;
        for (int hsel = 0; hsel < 2; ++hsel) {
          const long long tw0 = clock64();
          mbar_wait(&a_full[hsel], it & 1u);
          tc_fence_after();
          twait += clock64() - tw0;
          const int pc0 = hsel == 0 ? 0 : c_split, pc1 = hsel == 0 ? c_split : a.n_in;
          for (int term = 0; term < 3; ++term) {          // hi*hi, hi*lo, lo*hi
            const uint32_t acol = tmem + (term == 2 ? kTcColAlo : kTcColAhi);
            const uint64_t dterm = term == 1 ? d_lo : d_hi;
            for (int c = pc0; c < pc1; ++c) {
#pragma unroll 2
              for (int sp = 0; sp < hcp; ++sp) {
                int q = 2 * sp - rot;
                q += q < 0 ? CP : 0;
                const uint64_t db = dterm + (uint64_t)((c * (CP + 1) + q) * 8);
                asm volatile(
                    "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(td),
                    "r"(acol + (uint32_t)(c * a.KP + 8 * sp)), "l"(db), "r"(idesc), "r"(acc)
                    : "memory");
                acc = 1;
              }
            }
          }
          tc_commit(&a_empty[hsel]);           // this half of A may be overwritten once these MMAs retire
        }
        tc_commit(&d_full[buf]);
        if (a.prof && blockIdx.x == 0 && lane == 0) {
          a.prof[2] += ti1 - ti0;
          a.prof[3] += twait;
          a.prof[4] += clock64() - ti1 - twait;
        }
      }
    } else if (is_epi) {
      // ---------------- epilogue: accumulator row -> shared memory -> diagonal sums -> global
      const int og0 = esub == 0 ? 0 : (gcount + 1) / 2, og1 = esub == 0 ? (gcount + 1) / 2 : gcount;  // this warp set's planes
      const int ncol0 = og0 * a.kw, ncol1 = og1 * a.kw;                 // its accumulator columns
      for (int i = 0; i < nrows; ++i) {
        const int y = y_begin + 4 * i;
        const uint32_t it = rows_done + (uint32_t)i;
        const uint32_t buf = it & 1u, use = it >> 1;
        const long long te0 = clock64();
        mbar_wait_backoff(&d_full[buf], use & 1u);
        tc_fence_after();
        const long long te1 = clock64();
        tc_named_sync(1, kTcEpiWarps * 32);      // the previous row's readers of sP are done
        for (int n0 = ncol0 & ~15; n0 < ncol1; n0 += 16) {
          uint32_t v[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
              : "r"(tlane + kTcColD + buf * 128u + (uint32_t)n0));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + j >= ncol0 && n0 + j < ncol1) sP[m * NP1 + n0 + j] = __uint_as_float(v[j]);
        }
        tc_fence_before();
        mbar_arrive(&d_empty[buf]);              // the accumulator buffer is free for row y + 2
        tc_named_sync(2, kTcEpiWarps * 32);      // sP complete (both warp sets wrote their columns)
        // out[o][y][x0 + m] = b[o] + sum_kx P[m + kx][(o, kx)]
        const int x = x0 + m;
        if (m < a.twv && x < a.wout) {
          for (int og = og0; og < og1; ++og) {
            float sum = a.bias[gout0 + og];
            const float *p = sP + m * NP1 + og * a.kw;
            for (int kx = 0; kx < a.kw; ++kx) sum += p[kx * NP1 + kx];
            if (a.tanh_after) sum = tanhf(sum);
            a.out[(((size_t)n * a.n_out + gout0 + og) * a.hout + y) * a.wout + x] = sum;
          }
        }
        if (a.prof && blockIdx.x == 0 && warp == kTcGatherWarps && lane == 0) {
          a.prof[5] += te1 - te0;
          a.prof[6] += clock64() - te1;
        }
      }
    }
    rows_done += (uint32_t)nrows;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// ---------------------------------------------------------------- host side
int tc_plan_layer(const dm_ctx *ctx, int n_in, int n_out, int kh, int kw, TcPlan *p) {
  memset(p, 0, sizeof(*p));
  if (kh > 32 || kw > 64 || kw > kTcM / 2) return 0;
  p->KP = (kh + 7) / 8 * 8;
  p->Ktot = n_in * p->KP;
  p->KtotB = n_in * (p->KP + 4);   // per plane: KP/4 chunks + the copy of chunk 0
  if (p->Ktot > kTcMaxK || (n_in + 1) / 2 > 4) return 0;   // a gather thread keeps <= 4 planes x 4 prefetched rows in registers
  for (int G = n_out; G >= 1; --G) {
    const int Npad = (G * kw + 15) / 16 * 16;
    if (Npad > kTcMaxN) continue;
    size_t smem = (size_t)2 * Npad * p->KtotB * 4 + (size_t)kTcM * (Npad + 1) * 4 + 160;
    if (smem > ctx->smem_optin) continue;
    // the largest group that fits decides how many groups there are; the planes are then spread evenly
    // (an MMA costs its padded N whatever the number of real planes in the group)
    p->ngroups = (n_out + G - 1) / G;
    p->G = (n_out + p->ngroups - 1) / p->ngroups;
    p->Npad = (p->G * kw + 15) / 16 * 16;
    smem = (size_t)2 * p->Npad * p->KtotB * 4 + (size_t)kTcM * (p->Npad + 1) * 4 + 160;
    p->smem = smem;
    p->ok = 1;
    return 1;
  }
  return 0;
}

// weights [conn][kh][kw] (+ connection table) -> per group, per term (hi, lo), the K-major
// no-swizzle operand image of B[n = (og, kx)][k = (c, ky)]
void tc_pack_weights(const TcPlan &p, int n_in, int n_out, int kh, int kw, int n_conn, const int *conn, const float *weight,
                     std::vector<float> *out) {
  const size_t nk = (size_t)p.Npad * p.KtotB;
  out->assign((size_t)p.ngroups * 2 * nk, 0.0f);
  const int CP = p.KP / 4;
  const bool full = n_conn <= 0;
  const int nc = full ? n_in * n_out : n_conn;
  for (int e = 0; e < nc; ++e) {
    const int c = full ? e % n_in : conn[2 * e] - 1;
    const int o = full ? e / n_in : conn[2 * e + 1] - 1;
    const int g = o / p.G, og = o % p.G;
    for (int ky = 0; ky < kh; ++ky)
      for (int kx = 0; kx < kw; ++kx) {
        const int n = og * kw + kx;
        const size_t chunk = (size_t)c * (CP + 1) + ky / 4;   // chunk index along K (the plane's copy of chunk 0 is filled below)
        const size_t at = (size_t)(ky % 4) + 4 * (n % 8) + (size_t)(8 * p.KtotB) * (n / 8) + 32 * chunk;
        // several table entries may connect the same (c, o): their kernels add up
        const float wv = weight[((size_t)e * kh + ky) * kw + kx];
        float *hi = out->data() + (size_t)g * 2 * nk + at, *lo = hi + nk;
        const float sum = *hi + *lo + wv;
        uint32_t u;
        memcpy(&u, &sum, 4);
        u &= 0xffffe000u;
        float t;
        memcpy(&t, &u, 4);
        *hi = t;
        *lo = sum - t;
      }
  }
  // the copy of every plane's chunk 0 behind its last chunk: a K step that starts at the last chunk
  // (rotated pairing) continues into it
  for (int g = 0; g < p.ngroups; ++g)
    for (int term = 0; term < 2; ++term) {
      float *base = out->data() + ((size_t)g * 2 + term) * nk;
      for (int c = 0; c < n_in; ++c)
        for (int n = 0; n < p.Npad; ++n)
          for (int j = 0; j < 4; ++j) {
            const size_t row = (size_t)4 * (n % 8) + (size_t)(8 * p.KtotB) * (n / 8) + j;
            base[row + 32 * ((size_t)c * (CP + 1) + CP)] = base[row + 32 * ((size_t)c * (CP + 1))];
          }
    }
}

int tc_launch_layer(dm_ctx *ctx, const TcPlan &p, const float *B, const float *bias, int n_in, int n_out, int kh, int kw,
                    int tanh_after, const float *in, float *out, int n_img, int h, int w, int pad_l, int pad_r, int pad_t,
                    int pad_b) {
  TcArgs a{};
  a.in = in;
  a.out = out;
  a.B = B;
  a.bias = bias;
  a.n_img = n_img;
  a.n_in = n_in;
  a.n_out = n_out;
  a.kh = kh;
  a.kw = kw;
  a.tanh_after = tanh_after;
  a.h = h;
  a.w = w;
  a.hout = h + pad_t + pad_b - kh + 1;
  a.wout = w + pad_l + pad_r - kw + 1;
  a.pad_t = pad_t;
  a.pad_l = pad_l;
  a.KP = p.KP;
  a.Ktot = p.Ktot;
  a.KtotB = p.KtotB;
  a.G = p.G;
  a.ngroups = p.ngroups;
  a.Npad = p.Npad;
  a.twv = kTcM - kw + 1;
  a.col_tiles = (a.wout + a.twv - 1) / a.twv;
  // row bands: enough units to fill the machine about three times, bands no shorter than 4 kernel heights
  const int base = a.ngroups * n_img * a.col_tiles * 4;   // x 4 residue classes of output rows
  int nb = (3 * ctx->num_sms + base - 1) / base;
  const int max_nb = std::max(1, a.hout / std::max(8 * kh, 64));   // a unit re-loads a whole window at its start
  nb = std::max(1, std::min(nb, max_nb));
  a.band = (a.hout + nb - 1) / nb;
  a.nbands = (a.hout + a.band - 1) / a.band;
  a.units = a.ngroups * n_img * a.col_tiles * a.nbands * 4;
  DM_CHECK(ensure_func_smem(ctx, (const void *)conv_tc_kernel, p.smem));
  const int grid = std::min(a.units, ctx->num_sms);
  a.prof = nullptr;
  if (ctx->opt.volume_debug == 9) {
    DM_CUDA(cudaMalloc(&a.prof, 8 * sizeof(long long)));
    DM_CUDA(cudaMemset(a.prof, 0, 8 * sizeof(long long)));
  }
  conv_tc_kernel<<<grid, kTcThreads, p.smem, ctx->stream>>>(a);
  DM_CUDA(cudaGetLastError());
  count_launch(ctx);
  if (a.prof) {
    long long h[8];
    DM_CUDA(cudaStreamSynchronize(ctx->stream));
    DM_CUDA(cudaMemcpy(h, a.prof, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(a.prof);
    const double st = (double)std::max(1LL, h[7]);
    fprintf(stderr, "[conv_tc %dx%dx%dx%d N=%d K=%d units=%d grid=%d] cycles/step of CTA 0: gather wait %.0f work %.0f | issuer wait-D %.0f "
            "wait-A %.0f issue %.0f | epilogue wait %.0f work %.0f (steps %.0f)\n", n_in, kh, kw, n_out, p.Npad, p.Ktot, a.units, grid,
            h[0] / st, h[1] / st, h[2] / st, h[3] / st, h[4] / st, h[5] / st, h[6] / st, st);
  }
  return DM_OK;
}

}  // namespace dm
