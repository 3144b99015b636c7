// match_fused.cu -- K1+K2: cost volume, softmax statistics, winner-take-all argmax,
// thresholded score extraction and soft/sub-pixel mean in ONE pass; the
// H1 x W1 x maxh x maxw volume never exists in HBM.
//
// Reference semantics being fused (all per output pixel, window entries k in scan
// order dy-major, 1-based):
//   v_k   = sum_c (in1[c,y,x] - in2[c,y+dy,x+dx])^2     nn.SpatialMatching  (opticalflow_model.lua:93)
//   p_k   = exp(min_v - v_k) / sum_j exp(min_v - v_j)   Minus + SoftMax     (:94-109)
//   index = first argmax_k p_k, middle-index tie rule   getOutputConfidences (:153-161)
//   thr   = extractOutput(p, 0.11)                      extract_output.cpp:63-155
//   y,x   = sum_k p_k*row_k, sum_k p_k*col_k            OutputExtractor.lua:21-35
//   flow  = (row,col) - ceil(max/2), pasted in a canvas processOutput        (:201-252)
//
// Numerics: SSD in fp32, channel-ascending.  Three forms (SsdMode, match_kernels.cuh): the
// difference form with FMA contraction, the same with separately rounded multiply and add
// (DM_FLAG_EXACT_SSD, bit-exact with the CPU path), and for large calls |a|^2+|b|^2-2a.b behind a
// device-side bound on the norms (half the FP32 work, absolute error a few ulp of the norms).
// The softmax runs online (flash-style running minimum) with ex2.approx, so probabilities agree
// with the two-pass CPU path to ~1e-6 relative, not bit-wise.  When no probability is asked for
// (index / flow / min_ssd only) the exponentials are skipped altogether (kEpiWta).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "match_kernels.cuh"

namespace dm {

constexpr int kBarBytes = 384;  // room for the 2*kNSlot mbarriers, keeps what follows 128-byte aligned
constexpr int kBar2Bytes = 512; // two-row sweep: 2*kNSlot mbarriers + kNSlot refill counters

struct ExtractParams {
  SweepGeom g;
  unsigned flags;
  int M;              // 8 if thr < 0.2 else 4 (extract_output.cpp:82-84)
  float p_none;       // a probability below this is below the threshold for sure (thr - 1e-4)
  float p_gt;         // a probability above this is above the threshold for sure (thr + 1e-4)
  float thr_lo;       // shortlist trigger: e_k > thr_lo * S (thr_lo slightly under thr)
  float p_clear;      // rescale factor under which everything seen before a new minimum is out
  int mid_dy, mid_blk[2], mid_r[2], middle;  // zero-flow entry: dy, then (block, column) for even / odd pixels; 1-based index
  int cy, cx;                          // ceil(maxh/2), ceil(maxw/2)
  int h_img, w_img, hoff, woff;
  long long *index;
  float *min_ssd, *pmax, *flow_full;
  long long *index_thr;
  float *score_thr, *soft_yx;
  unsigned long long *n_untouched;
  float *conf_marginal;  // kEpiSoft: 1 where max_dy sum_dx p(dy,dx) > prob_threshold
  float p_thr;           // prob_threshold as float
  // thresholded extraction (extractOutput on the probabilities), see ExtractEpi::tile_end
  int nwords;           // 32-bit words of the per-pixel (dy, dx-block) shortlist bitmap
  int gb;               // consecutive blocks that share one shortlist bit (1 unless the window is huge)
  int *todo;            // pixels that need the exact pass over their shortlist
  unsigned *todo_mask;  // [todo slot][nwords]
  unsigned *ntodo;
  float *vmin, *vinv;   // per-pixel min and 1/sum for the exact pass
  // kDot / kFma twin launch: both kernels read the largest |a|^2 and |b|^2 the norm pre-pass
  // found and only the one the bound selects runs (the other returns at once)
  const unsigned *stats;  // {max |a|^2, max |b|^2} as float bits; NULL: run unconditionally
  float dot_limit;        // kDot is taken when max |a|^2 + max |b|^2 <= dot_limit
  // kDot: window entries closer than tau = tau_rel * (max |a|^2 + max |b|^2) to the running minimum
  // cannot be ordered by the dot form (its absolute error is (C + 2) * 2^-24 of the norms, see
  // dot_error_rel).  Such pixels -- and, in every form, pixels whose zero-flow entry is within a few
  // ulp of the minimum -- are appended to `resc` and recomputed entry by entry in the difference
  // form with the reference's literal order of operations (generic_rescore).
  float tau_rel;
  int *resc;
  unsigned *nresc;
  int dbg;                // tuning switches of the two-row sweep (option volume_debug): 1 = no epilogue
  const float *nb;        // |b|^2 per frame-2 pixel as the norm pre-pass wrote it (two-row sweep: tile end)
  long long nb_sn, nb_sy;
  const float *in2;       // frame 2 as the kernels address it, for the exact re-score of the winner
  long long s2n, s2c, s2y;
};

// Per-thread state of the fused reduction: for each of the thread's kP pixels the running
// minimum (and its index), the softmax denominator and -- when SOFT -- the two first
// moments, all relative to the running minimum (flash-style rescaling when it moves).
// Shared memory per pixel: the shortlist bitmap [nwords] and the SSD of the zero-flow entry.
enum EpiKind { kEpiScores = 0, kEpiSoft = 1, kEpiWta = 2 };

// EPI: kEpiScores -- argmin + soft-max statistics + thresholded extraction; kEpiSoft -- the same
// plus the first moments; kEpiWta -- winner-take-all only (index / flow / min_ssd): the arg max
// of the soft-max is the arg min of the SSD, so when no probability is asked for the
// exponentials are skipped altogether (the zero-flow tie rule then compares exp(m - v_mid)
// with 1 instead of the normalised probabilities).
template <class Cfg, int EPI, bool DOT>
struct ExtractEpi {
  static constexpr bool SOFT = EPI == kEpiSoft;
  static constexpr bool WTA = EPI == kEpiWta;
  static constexpr int kCThreads = Cfg::kCThreads;
  const ExtractParams &P;
  float m[kP], S[kP], sx[kP], sy[kP];
  float rowsum[SOFT ? kP : 1], rowmax[SOFT ? kP : 1];  // kEpiSoft: marginal over dx of the current
                                                        // window row, and the largest one so far
  int idx[kP];
  // thresholded extraction: e2[p] = upper bound of the second largest exp(m - v) seen so far
  // (relative to the running minimum, rescaled with it); vfrom[p] = first shortlist bit still
  // in reach of the threshold
  float e2[kP];
  int vfrom[kP];
  unsigned amb;    // kDot: bit p = pixel p has a second entry within tau of its minimum
  float tau;       // kDot: see ExtractParams::tau_rel; 0 otherwise
  unsigned *mask;  // [nwords][kP][kCThreads] words, this thread's column
  float *vmid;     // [kP][kCThreads]

  __device__ ExtractEpi(const ExtractParams &p, unsigned *smem_extra)
      : P(p), mask(smem_extra + threadIdx.x) {
    vmid = reinterpret_cast<float *>(smem_extra + (size_t)p.nwords * kP * kCThreads) + threadIdx.x;
    tau = 0.0f;
    if (DOT && p.stats) tau = p.tau_rel * (__uint_as_float(p.stats[0]) + __uint_as_float(p.stats[1]));
  }

  __device__ __forceinline__ void tile_begin(int, int, int) {
#pragma unroll
    for (int p = 0; p < kP; ++p) {
      m[p] = __int_as_float(0x7f800000);
      S[p] = sx[p] = sy[p] = 0.0f;
      if (SOFT) rowsum[p] = rowmax[p] = 0.0f;
      idx[p] = 1;
      e2[p] = 0.0f;
      vfrom[p] = 0;
      vmid[p * kCThreads] = 0.0f;
    }
    amb = 0u;
    for (int w = 0; w < P.nwords * kP; ++w) mask[w * kCThreads] = 0u;
  }

  // a new running minimum v at 1-based index k for pixel p (strict <: the first occurrence
  // wins, like TH max): rescale what was accumulated relative to the old one
  __device__ __forceinline__ void new_min(int p, float v, int k, int bit) {
    if (WTA) {
      m[p] = v;
      idx[p] = k;
      return;
    }
    const float sc = ex2_approx((v - m[p]) * kLog2e);  // old m = +inf -> 0
    S[p] *= sc;
    if (SOFT) {
      sx[p] *= sc;
      sy[p] *= sc;
      rowsum[p] *= sc;
      rowmax[p] *= sc;
    }
    e2[p] = fmaxf(e2[p], 1.0f) * sc;  // the old minimum (e = 1) becomes a runner-up
    // old minimum (and everything before it) below the threshold for good: earlier bits are dead
    if (sc < P.p_clear) vfrom[p] = bit;
    m[p] = v;
    idx[p] = k;
  }

  // acc[p][r] is the SSD of pixel p at dx = 8*blk - (p & 1) + r (skewed blocks, see
  // BlockSchedule); R = 8, or 2 for a narrow tail
  template <int R>
  __device__ __forceinline__ void block(float (&acc)[kP][R], int dy, int blk) {
    const float inf = __int_as_float(0x7f800000);
    const int dx0 = blk * kR;  // dx of r = 0 for even pixels; odd pixels start one earlier
    // entries outside the window (warp-uniform branches): dx = -1 of the odd pixels in the
    // first block, dx >= maxw in the last one
    if (blk == 0) {
#pragma unroll
      for (int p = 1; p < kP; p += 2) acc[p][0] = inf;
    }
    if (dx0 + R > P.g.maxw) {
#pragma unroll
      for (int p = 0; p < kP; ++p)
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (dx0 - (p & 1) + r >= P.g.maxw) acc[p][r] = inf;
    }
    // zero-flow entry, for the tie rule (warp-uniform branch, once or twice per sweep)
    if (dy == P.mid_dy && (blk == P.mid_blk[0] || blk == P.mid_blk[1])) {
#pragma unroll
      for (int p = 0; p < kP; ++p)
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (blk == P.mid_blk[p & 1] && r == P.mid_r[p & 1]) vmid[p * kCThreads] = acc[p][r];
    }
    float bm[kP];
    bool any = false;
    const int blkid = dy * P.g.bs.per_row() + blk;
    const int bit = P.gb == 1 ? blkid : blkid / P.gb;
#pragma unroll
    for (int p = 0; p < kP; ++p) {
      bm[p] = acc[p][0];
#pragma unroll
      for (int r = 1; r < R; ++r) bm[p] = fminf(bm[p], acc[p][r]);
      any |= (DOT ? bm[p] - tau : bm[p]) < m[p];
    }
    const int kbase = dy * P.g.maxw + dx0 + 1;  // 1-based index of (even pixel, r = 0)
    if (any) {  // some pixel of this thread has a new running minimum in this block (kDot: or an
                // entry the dot form cannot tell from the minimum)
#pragma unroll
      for (int p = 0; p < kP; ++p)
        if (bm[p] < m[p]) {
          int rb = 0;
#pragma unroll
          for (int r = R - 1; r >= 0; --r)
            if (acc[p][r] == bm[p]) rb = r;
          if (DOT) {
            // everything seen before is >= the old minimum: ambiguous iff that one is within tau;
            // inside this block count the entries within tau of the new one
            int near = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) near += acc[p][r] - tau <= bm[p] ? 1 : 0;
            const bool a = m[p] - tau <= bm[p] || near > 1;
            amb = (amb & ~(1u << p)) | (a ? 1u << p : 0u);
          }
          new_min(p, bm[p], kbase - (p & 1) + rb, bit);
        } else if (DOT && bm[p] - tau <= m[p]) {
          amb |= 1u << p;
        }
    }
    if (WTA) return;
    const float rowf = (float)(dy + 1);
    bool cand = false;
    float eb[kP], ebm[kP];
    // soft-max terms two pixels at a time: acc[2pp][r] and acc[2pp+1][r] are the halves of one
    // packed accumulator, so the scale and the sums are FFMA2 / FADD2 (same roundings as scalar)
#pragma unroll
    for (int pp = 0; pp < kP / 2; ++pp) {
      const float2 mL2 = make_float2(m[2 * pp] * kLog2e, m[2 * pp + 1] * kLog2e);
      float2 eb2 = make_float2(0.0f, 0.0f), ex2 = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float2 t = __ffma2_rn(make_float2(acc[2 * pp][r], acc[2 * pp + 1][r]),
                                    make_float2(-kLog2e, -kLog2e), mL2);
        const float2 e = make_float2(ex2_approx(t.x), ex2_approx(t.y));
        eb2 = __fadd2_rn(eb2, e);
        if (SOFT) ex2 = __ffma2_rn(e, make_float2((float)(r + 1), (float)(r + 1)), ex2);
      }
      eb[2 * pp] = eb2.x;
      eb[2 * pp + 1] = eb2.y;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int p = 2 * pp + h;
        const float ex = h ? ex2.y : ex2.x;
        S[p] += eb[p];
        if (SOFT) {
          sx[p] += fmaf(eb[p], (float)(dx0 - (p & 1)), ex);  // column = dx + 1
          sy[p] = fmaf(eb[p], rowf, sy[p]);
          rowsum[p] += eb[p];
          if (blk == P.g.bs.per_row() - 1) {  // the window row is complete (warp-uniform)
            rowmax[p] = fmaxf(rowmax[p], rowsum[p]);
            rowsum[p] = 0.0f;
          }
        }
        // thresholded extraction: p_k(final) <= e_k / S(now), so an entry of this block can
        // end above the threshold only if the block's largest e exceeds thr * S
        ebm[p] = 0.0f;
        if (P.nwords) {
          ebm[p] = ex2_approx(fmaf(bm[p], -kLog2e, h ? mL2.y : mL2.x));
          cand |= ebm[p] > P.thr_lo * S[p];
        }
      }
    }
    if (cand) {
#pragma unroll
      for (int p = 0; p < kP; ++p) {
        const float lim = P.thr_lo * S[p];
        if (ebm[p] > lim) {
          mask[((bit >> 5) * kP + p) * kCThreads] |= 1u << (bit & 31);
          // runner-up bound: in the block that holds the minimum (first occurrence) the second
          // largest e is at most eb - ebm; in any other block it is the block's largest e
          // (the odd pixels' dx = -1 slot of block 0 aliases the previous row's last index)
          const int kb = kbase - (p & 1);
          const bool own = idx[p] >= kb + (((p & 1) && blk == 0) ? 1 : 0) && idx[p] < kb + R;
          e2[p] = fmaxf(e2[p], own ? eb[p] - ebm[p] : ebm[p]);
        }
      }
    }
  }

  // kDot: the winner is re-scored in the difference form (the dot form cancels: a perfect match
  // would read a few ulp of |a|^2 + |b|^2 instead of 0) and the soft-max sums, which were taken
  // relative to the dot-form minimum, are moved to the re-scored one: every term but the
  // winner's own (exactly 1) scales by c = exp(m_diff - m_dot).  a2 holds -2a.
  template <int CT>
  __device__ __forceinline__ void tile_rescore(const float2 (&a2)[CT][2], int n, int y, int x0) {
    if (!DOT || y >= P.g.H1) return;
    if (WTA && !P.min_ssd) return;
    if (!P.in2) return;   // statistics pass of the soft-max volume: min and sum stay in the dot form (minus |a|^2),
                          // which is what the volume sweep that follows computes too
#pragma unroll
    for (int p = 0; p < kP; ++p) {
      if (x0 + p >= P.g.W1) continue;
      const int dy = (idx[p] - 1) / P.g.maxw, dx = (idx[p] - 1) % P.g.maxw;
      const float *b = P.in2 + (long long)n * P.s2n + (long long)(y + dy) * P.s2y + (x0 + p + dx);
      float acc = 0.0f, na = 0.0f;
#pragma unroll
      for (int k = 0; k < CT; ++k) {
        const float2 av = a2[k][p >> 1];
        const float a = -0.5f * ((p & 1) ? av.y : av.x);
        na = fmaf(a, a, na);
        const float d = a - (k < P.g.Cin ? __ldg(b + (long long)k * P.s2c) : 0.0f);
        acc = fmaf(d, d, acc);
      }
      if (!WTA) {
        const float c = expf(acc - (m[p] + na));   // m is the dot-form minimum without |a|^2
        S[p] = fmaf(S[p] - 1.0f, c, 1.0f);
        e2[p] *= c;
        if (SOFT) {
          const float rw = (float)(dy + 1), cw = (float)(dx + 1);
          sy[p] = fmaf(sy[p] - rw, c, rw);
          sx[p] = fmaf(sx[p] - cw, c, cw);
        }
      }
      m[p] = acc;
    }
  }

  __device__ __forceinline__ float idx_min_ssd(int p) const { return m[p]; }

  __device__ __forceinline__ void tile_end(int n, int y, int x0) {
    const SweepGeom &g = P.g;
    if (y >= g.H1) return;
    unsigned untouched = 0;
#pragma unroll
    for (int p = 0; p < kP; ++p) {
      const int x = x0 + p;
      if (x >= g.W1) continue;
      const size_t o = ((size_t)n * g.H1 + y) * g.W1 + x;
      const float inv = WTA ? 1.0f : 1.0f / S[p];
      int win = idx[p];
      // Zero-flow tie rule, p[middle] == max p (opticalflow_model.lua:157-159), the same predicate in
      // every epilogue: equal SSDs are a tie; an SSD a few ulp above the minimum may or may not
      // round to the same probability, which only the reference's literal arithmetic decides --
      // those pixels, like the ones the dot form could not order, go to the rescore list and
      // every output of theirs is written by generic_rescore instead.
      bool rescue = DOT && ((amb >> p) & 1u);
      if (!DOT && (P.flags & DM_FLAG_TIE_MIDDLE) && win != P.middle) {
        const float gapm = vmid[p * kCThreads] - idx_min_ssd(p);
        if (gapm == 0.0f) win = P.middle;
        else if (gapm < 1.0e-6f) rescue = true;
      }
      if (rescue && P.resc) {
        P.resc[atomicAdd(P.nresc, 1u)] = (int)o;
        continue;
      }
      if (P.index) P.index[o] = win;
      if (P.min_ssd) P.min_ssd[o] = m[p];
      if (P.pmax && !WTA) P.pmax[o] = inv;
      if (SOFT && P.conf_marginal) P.conf_marginal[o] = rowmax[p] * inv > P.p_thr ? 1.0f : 0.0f;
      if (SOFT && P.soft_yx) {
        const size_t plane = (size_t)g.H1 * g.W1;
        const size_t so = (size_t)n * 2 * plane + (size_t)y * g.W1 + x;
        P.soft_yx[so] = sy[p] * inv;
        P.soft_yx[so + plane] = sx[p] * inv;
      }
      if (P.flow_full) {
        const int row = (win - 1) / g.maxw + 1, col = (win - 1) % g.maxw + 1;
        const size_t plane = (size_t)P.h_img * P.w_img;
        const size_t fo = (size_t)n * 2 * plane + (size_t)(y + P.hoff) * P.w_img + (x + P.woff);
        P.flow_full[fo] = (float)(row - P.cy);
        P.flow_full[fo + plane] = (float)(col - P.cx);
      }
      if (P.todo && !WTA) {
        // extractOutput(prob, thr) (extract_output.cpp:63-155).  pmax < thr: nothing qualifies,
        // the pixel stays untouched.  pmax > thr and the runner-up bound e2/S < thr: the list is
        // {pmax}, ret = its position, score = M * pmax (prefix sums of {pmax,0,..}).  Anything within 1e-4 of the threshold, and every pixel that may have
        // two or more entries above it, goes to the exact per-pixel pass, which re-scores only
        // the (dy, dx-block)s on the pixel's shortlist.
        long long ret = 0;
        float score = 0.0f;
        if (inv > P.p_gt && e2[p] * inv < P.p_none) {
          ret = idx[p];
          score = (float)((double)P.M * (double)inv);
        } else if (inv < P.p_none) {
          ++untouched;
        } else {
          const unsigned slot = atomicAdd(P.ntodo, 1u);
          P.todo[slot] = (int)o;
          for (int w = 0; w < P.nwords; ++w) {
            unsigned bits = mask[(w * kP + p) * kCThreads];
            const int lo = vfrom[p] - 32 * w;  // bits below vfrom are dead
            if (lo >= 32) bits = 0u;
            else if (lo > 0) bits &= ~0u << lo;
            P.todo_mask[(size_t)slot * P.nwords + w] = bits;
          }
          P.vmin[o] = m[p];
          P.vinv[o] = inv;
        }
        if (P.index_thr) P.index_thr[o] = ret;
        if (P.score_thr) P.score_thr[o] = score;
      }
    }
    if (P.n_untouched && untouched) atomicAdd(P.n_untouched + n, (unsigned long long)untouched);
  }
};

template <class Cfg, int CT, int MODE, int EPI>
__global__ void __launch_bounds__(Cfg::kThreads, Cfg::kWarps <= 6 ? 2 : 1)
match_extract_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_nb,
                     const ExtractParams P) {
  if (P.stats) {  // twin launch: the norm bound picks the dot or the difference form
    const bool dot_ok = __uint_as_float(P.stats[0]) + __uint_as_float(P.stats[1]) <= P.dot_limit;
    if (dot_ok != (MODE == kDot)) return;
  }
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *ring = reinterpret_cast<float *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)P.g.nslot * P.g.slab_floats);
  unsigned *extra = reinterpret_cast<unsigned *>(reinterpret_cast<unsigned char *>(full) + kBarBytes);
  ExtractEpi<Cfg, EPI, MODE == kDot> epi(P, extra);
  run_sweep<Cfg, CT, MODE>(&tmap, &tmap_nb, P.g, ring, full, epi);
}

}  // namespace dm
#include "match_sweep2.cuh"
namespace dm {

// ---------------------------------------------------------------- exact threshold pass
// sorting networks of the reference (extract_output.cpp:27-33, :35-61), same order; compare-
// exchanges with compile-time indices keep the eight (value, position) pairs in registers
#define DM_CX(a, b)                \
  if (val[b] > val[a]) {           \
    float t_ = val[a];             \
    val[a] = val[b];               \
    val[b] = t_;                   \
    t_ = pos[a];                   \
    pos[a] = pos[b];               \
    pos[b] = t_;                   \
  }

// Sort `M` (value,pos) pairs descending with the reference's network, return the
// reference's ret and score (prefix sums in fp32, total in double).
template <int M>
__device__ __forceinline__ void net_sort_score_m(float (&val)[8], float (&pos)[8], long long *ret, float *score) {
  if (M == 4) {
    DM_CX(0, 2) DM_CX(1, 3) DM_CX(0, 1) DM_CX(2, 3) DM_CX(1, 2)
  } else {
    DM_CX(0, 1) DM_CX(2, 3) DM_CX(4, 5) DM_CX(6, 7) DM_CX(0, 2) DM_CX(1, 3) DM_CX(4, 6) DM_CX(5, 7) DM_CX(1, 2)
    DM_CX(5, 6) DM_CX(0, 4) DM_CX(3, 7) DM_CX(1, 5) DM_CX(2, 6) DM_CX(1, 4) DM_CX(3, 6) DM_CX(2, 4) DM_CX(3, 5)
    DM_CX(3, 4)
  }
  *ret = (long long)pos[0];
#pragma unroll
  for (int k = 1; k < M; ++k) val[k] = __fadd_rn(val[k], val[k - 1]);
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < M; ++k) acc += (double)val[k];
  *score = (float)acc;
}
#undef DM_CX
__device__ __forceinline__ void net_sort_score(float (&val)[8], float (&pos)[8], int M, long long *ret,
                                               float *score) {
  if (M == 4)
    net_sort_score_m<4>(val, pos, ret, score);
  else
    net_sort_score_m<8>(val, pos, ret, score);
}

struct ThresholdPass {
  const float *in1, *in2;
  long long s1n, s1c, s1y, s2n, s2c, s2y;
  int N, C, H1, W1, maxh, maxw, nwords, gb, M, exact;
  int mark;  // diagnostics (option debug_todo): score_thr = -1 on the pixels this pass handled
  BlockSchedule bs;
  double thr;
  const int *todo;
  const unsigned *todo_mask, *ntodo;
  const float *vmin, *vinv;
  long long *index_thr;
  float *score_thr;
  unsigned long long *n_untouched;
};

// Exact extractOutput for the pixels the sweep could not decide: eight lanes per pixel (a
// block is at most 8 entries wide), four pixels per warp in flight.  A group walks its
// pixel's shortlist of (dy, dx-block)s in scan order, recomputes those SSDs with the sweep's
// arithmetic (all channel loads of a block issued before the first use, so a block costs one
// memory latency), and applies extract_output.cpp:63-155 literally.
template <int CT>
__global__ void __launch_bounds__(128) threshold_exact_kernel(const ThresholdPass T) {
  const int lane = threadIdx.x & 31, grp = lane >> 3, gl = lane & 7;
  const unsigned gmask = 0xffu << (8 * grp);
  const unsigned n = *T.ntodo;
  const int per_row = T.bs.per_row();
  const unsigned ngroups = gridDim.x * 16;
  for (unsigned e = (blockIdx.x * 4 + (threadIdx.x >> 5)) * 4 + grp; e < n; e += ngroups) {
    const int px = T.todo[e];  // the list holds 32-bit pixel indices: 32-bit divisions
    const int row = px / T.W1;
    const int x = px - row * T.W1, pn = row / T.H1, y = row - pn * T.H1;
    const float *a = T.in1 + pn * T.s1n + y * T.s1y + x;
    const float *b0 = T.in2 + pn * T.s2n + y * T.s2y + x;
    const float m = T.vmin[px], inv = T.vinv[px];
    float av[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) av[c] = c < T.C ? __ldg(a + c * T.s1c) : 0.0f;
    float val[8], pos[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) val[j] = pos[j] = 0.0f;
    int got = 0;
    for (int w = 0; w < T.nwords && got < T.M; ++w) {
      unsigned bits = T.todo_mask[(size_t)e * T.nwords + w];
      while (bits && got < T.M) {
        const int bitid = w * 32 + __ffs(bits) - 1;
        bits &= bits - 1;
       for (int id = bitid * T.gb; id < (bitid + 1) * T.gb && id < T.maxh * per_row && got < T.M; ++id) {
        const int dy = id / per_row, blk = id - dy * per_row;
        const int dxb = blk * kR - (x & 1);  // skewed blocks: odd pixels start one column earlier
        const int width = blk >= T.bs.n8 ? T.bs.tail_r : kR;
        const bool valid = gl < width && dxb + gl >= 0 && dxb + gl < T.maxw;
        float pk = 0.0f;
        if (valid) {
          const float *b = b0 + dy * T.s2y + dxb + gl;
          float bv[CT];
#pragma unroll
          for (int c = 0; c < CT; ++c) bv[c] = c < T.C ? __ldg(b + c * T.s2c) : 0.0f;
          float acc = 0.0f;
#pragma unroll
          for (int c = 0; c < CT; ++c)
            if (c < T.C) {
              const float d = av[c] - bv[c];
              acc = T.exact ? __fadd_rn(acc, __fmul_rn(d, d)) : fmaf(d, d, acc);
            }
          pk = expf(m - acc) * inv;
        }
        unsigned hit = (__ballot_sync(gmask, valid && (double)pk > T.thr) >> (8 * grp)) & 0xffu;
        while (hit && got < T.M) {
          const int src = __ffs(hit) - 1;
          hit &= hit - 1;
          const float pv = __shfl_sync(gmask, pk, 8 * grp + src);
          const float pp = (float)(dy * T.maxw + dxb + src + 1);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j == got) {
              val[j] = pv;
              pos[j] = pp;
            }
          ++got;
        }
       }
      }
    }
    if (gl != 0) continue;
    long long ret = 0;
    float score = 0.0f;
    if (got > 0)
      net_sort_score(val, pos, T.M, &ret, &score);
    else if (T.n_untouched)
      atomicAdd(T.n_untouched + pn, 1ull);
    if (T.index_thr) T.index_thr[px] = ret;
    if (T.score_thr) T.score_thr[px] = T.mark ? -1.0f : score;
  }
}

// ---------------------------------------------------------------- volume epilogue
struct VolumeParams {
  SweepGeom g;
  int mode;            // dm_volume_mode
  const float *vmin;   // [N][H1][W1], mode NEG_SOFTMAX: min_k v_k
  const float *vinv;   // [N][H1][W1] 1 / sum_k exp(min - v_k)
  float *out;          // [N][H1][W1][K]
  int debug;           // DM_VOLUME_DEBUG: 1 = skip the global stores, 2 = stores wrap into 32 MB
  // soft-max volume in the dot form (probabilities only need 1e-4): twin launch on the norm bound
  // as in the fused kernel; NULL = run unconditionally
  const unsigned *stats;
  float dot_limit;
};

// Store staging of the volume kernel.  The output tensor [px][K] gives every pixel a
// contiguous stream of K floats, K*4 bytes is in general not a multiple of the 32-byte DRAM
// sector, and a (pixel, block) produces only 8 stream entries: storing them as they come
// touches two partial sectors per run, and B200's L2 fills partially written sectors from
// DRAM (measured: +60 % reads, +47 % writes, 0.9 TB/s).  So each pixel's stream goes through a
// 16-entry circular buffer in shared memory (index = stream position mod 16) and leaves as
// whole, 32-byte-aligned sectors: after every block at most one sector per pixel completes, 8
// lanes write it.  Only the first and last sector of a pixel's stream (shared with the
// neighbouring pixels) are written partially.
constexpr int kStgPlane = 16 * 66;  // one parity plane: [16 positions][64 pixels + 2 pad]

struct VolumeEpi {
  const VolumeParams &P;
  float *stg;  // this warp's staging: plane 0 even pixels, plane 1 odd pixels
  float mL[kP], inv[kP];
  size_t obase;  // stream offset (floats) of pixel 0 of this warp's row in the tile
  int npx;       // valid pixels of this warp's row in the tile

  __device__ VolumeEpi(const VolumeParams &p, float *stg_all)
      : P(p), stg(stg_all + (threadIdx.x >> 5) * (2 * kStgPlane)) {}

  __device__ __forceinline__ void tile_begin(int n, int y, int x0) {
    const SweepGeom &g = P.g;
    const int lane = threadIdx.x & 31;
    const int xt = x0 - lane * kP;
    const bool rowok = y < g.H1;
    npx = rowok ? min(kTW, g.W1 - xt) : 0;
    obase = (((size_t)n * g.H1 + (rowok ? y : 0)) * g.W1 + xt) * (size_t)(g.maxh * g.maxw);
#pragma unroll
    for (int p = 0; p < kP; ++p) {
      mL[p] = 0.0f;
      inv[p] = 1.0f;
      if (P.mode == DM_VOLUME_NEG_SOFTMAX && rowok && x0 + p < g.W1) {
        const size_t o = ((size_t)n * g.H1 + y) * g.W1 + x0 + p;
        mL[p] = P.vmin[o] * kLog2e;
        inv[p] = P.vinv[o];
      }
    }
  }

  // acc[p][r] is window entry dx = 8*blk - (p & 1) + r of pixel p (skewed blocks)
  template <int R>
  __device__ __forceinline__ void block(float (&acc)[kP][R], int dy, int blk) {
    const int lane = threadIdx.x & 31;
    const int maxw = P.g.maxw, K = P.g.maxh * maxw;
    if (P.mode == DM_VOLUME_NEG_SOFTMAX) {
#pragma unroll
      for (int p = 0; p < kP; ++p)
#pragma unroll
        for (int r = 0; r < R; ++r)
          acc[p][r] = ex2_approx(fmaf(acc[p][r], -kLog2e, mL[p])) * inv[p];
    }
    // 1. stage: stream index q = dy*maxw + dx -> position q & 15, pixel pairs (0,2) and (1,3)
    const int q0 = dy * maxw + blk * kR;  // stream index of (even pixel, r = 0)
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int dxe = blk * kR + r, dxo = dxe - 1;
      if (dxe < maxw)
        *reinterpret_cast<float2 *>(stg + ((q0 + r) & 15) * 66 + 2 * lane) =
            make_float2(acc[0][r], acc[2][r]);
      if (dxo >= 0 && dxo < maxw)
        *reinterpret_cast<float2 *>(stg + kStgPlane + ((q0 + r - 1) & 15) * 66 + 2 * lane) =
            make_float2(acc[1][r], acc[3][r]);
    }
    __syncwarp();
    // 2. flush the sectors this block completed
    if (P.debug == 1) return;
    const long long s_even = (long long)dy * maxw + min(blk * kR + R, maxw);
    const bool ends = dy == P.g.maxh - 1 && blk * kR + R - 1 >= maxw;  // a pixel stream may end here
    if (s_even >= 24 && !ends)
      flush_fast<R>(dy, blk);
    else
      flush_edges<R>(dy, blk);
  }

  // Steady state: every completed sector lies fully inside its pixel's stream.  Two lanes per
  // sector (16 bytes each, one STG.128), 16 pixels per iteration.  Advancing 16 pixels moves the
  // stream start by 16*K floats, a multiple of the sector, so the lane's alignment phase, the
  // staging rows it reads and "did a sector complete" are fixed for the whole block.
  template <int R>
  __device__ __forceinline__ void flush_fast(int dy, int blk) {
    const int lane = threadIdx.x & 31;
    const int maxw = P.g.maxw, K = P.g.maxh * maxw;
    const int pxl = lane >> 1, h = lane & 1, par = pxl & 1;
    const int lo = max(blk * kR - par, 0), hi = min(blk * kR - par + R, maxw);
    const int nvalid = hi - lo;
    const int s_end = dy * maxw + hi;
    const int phase = (int)((obase + (size_t)pxl * K) & 7);
    const int u = (phase + s_end) & 7;       // floats of the still incomplete sector
    if (u >= nvalid) return;                 // no sector completed for this lane's pixels
    const int qs = s_end - u - 8 + 4 * h;    // stream index of this lane's 4 floats
    const float *src = stg + par * kStgPlane + (pxl >> 1);
    const int r0 = ((qs + 0) & 15) * 66, r1 = ((qs + 1) & 15) * 66, r2 = ((qs + 2) & 15) * 66,
              r3 = ((qs + 3) & 15) * 66;
    float *dst = P.out + obase + (size_t)pxl * K + qs;
    const size_t step = (size_t)16 * K;
#pragma unroll
    for (int it = 0; it < kTW / 16; ++it) {
      const float4 v = make_float4(src[r0 + 8 * it], src[r1 + 8 * it], src[r2 + 8 * it], src[r3 + 8 * it]);
      if (16 * it + pxl < npx) *reinterpret_cast<float4 *>(dst + it * step) = v;
    }
  }

  // First and last sectors of a pixel's stream are shared with the neighbouring pixels:
  // scalar stores, 8 lanes per sector, floats outside the stream masked.
  template <int R>
  __device__ __forceinline__ void flush_edges(int dy, int blk) {
    const int lane = threadIdx.x & 31;
    const int maxw = P.g.maxw, K = P.g.maxh * maxw;
    const int g4 = lane >> 3, i = lane & 7, par = g4 & 1;
    const int lo = max(blk * kR - par, 0), hi = min(blk * kR - par + R, maxw);
    const int nvalid = hi - lo;
    const long long s_end = (long long)dy * maxw + hi;  // stream index after this block
    const bool last = dy == P.g.maxh - 1 && hi == maxw;  // the pixel's stream ends here
    const float *src = stg + par * kStgPlane + (g4 >> 1);
    float *out = P.out;
#pragma unroll 4
    for (int it = 0; it < kTW / 4; ++it) {
      const int px = 4 * it + g4;
      const long long base = (long long)obase + (long long)px * K;  // stream start of the pixel
      const long long G = base + s_end - nvalid, Gend = base + s_end;
      const long long E = Gend & ~7LL;  // end of the last complete sector
      if (px < npx && nvalid > 0) {
        if (E > G) {  // sector [E-8, E) completed (floats before `base` are the previous pixel's)
          const long long f = E - 8 + i;
          if (f >= base) out[f] = src[(int)((f - base) & 15) * 66 + 2 * it];
        }
        if (last && i < (int)(Gend - E)) {
          const long long f = E + i;  // a stream shorter than a sector starts after E: mask as above
          if (f >= base) out[f] = src[(int)((f - base) & 15) * 66 + 2 * it];
        }
      }
    }
  }

  template <int CT>
  __device__ __forceinline__ void tile_rescore(const float2 (&)[CT][2], int, int, int) {}
  __device__ __forceinline__ void tile_end(int, int, int) {}
};

template <int CT, int MODE>
__global__ void __launch_bounds__(VolumeCfg::kThreads, 1)
match_volume_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_nb,
                    const VolumeParams P) {
  if (P.stats) {  // twin launch: the norm bound picks the dot or the difference form
    const bool dot_ok = __uint_as_float(P.stats[0]) + __uint_as_float(P.stats[1]) <= P.dot_limit;
    if (dot_ok != (MODE == kDot)) return;
  }
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *ring = reinterpret_cast<float *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)P.g.nslot * P.g.slab_floats);
  float *stg = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(full) + kBarBytes);
  VolumeEpi epi(P, stg);
  run_sweep<VolumeCfg, CT, MODE>(&tmap, &tmap_nb, P.g, ring, full, epi);
}

}  // namespace dm
#include "match_volume_px.cuh"
namespace dm {

// ---------------------------------------------------------------- host side
static size_t ring_bytes(const SweepGeom &g, int nslot) {
  return (size_t)nslot * g.slab_floats * sizeof(float);
}

struct Prepared {
  SweepGeom g;
  CUtensorMap tmap;
  int CT;
  const float *in1_dev;
  const float *in2_dev;
  long long s2n, s2c, s2y;  // strides of in2_dev (elements)
  int Cin;
};

// Stage inputs (host -> device, or repack device views TMA cannot describe) and
// build the tensor map of frame 2: dims {W2, H2, C, N}, box {WB, 1, CT, 1}.  The
// channel dim of the box is CT >= C: TMA zero-fills the missing channels and the
// kernel keeps a = 0 for them, so they add exactly 0 to every SSD.
static int prepare(Call &call, const dm_pair *in, int maxh, int maxw, int tile_rows, Prepared *out) {
  dm_ctx *ctx = call.ctx;
  DM_REQUIRE(in && in->in1 && in->in2, "input pointers are NULL");
  DM_REQUIRE(in->n_pairs >= 1 && in->channels >= 1, "n_pairs and channels must be >= 1");
  DM_REQUIRE(maxh >= 1 && maxw >= 1, "window must be at least 1x1 (got %dx%d)", maxh, maxw);
  DM_REQUIRE(in->h1 >= 1 && in->w1 >= 1, "empty frame-1 map (%dx%d)", in->h1, in->w1);
  DM_REQUIRE(in->h2 >= in->h1 + maxh - 1 && in->w2 >= in->w1 + maxw - 1,
             "frame 2 (%dx%d) smaller than frame 1 (%dx%d) + window (%dx%d) - 1", in->h2, in->w2,
             in->h1, in->w1, maxh, maxw);
  SweepGeom &g = out->g;
  g.N = in->n_pairs;
  g.C = in->channels;
  g.Cin = in->channels;
  g.H1 = in->h1;
  g.W1 = in->w1;
  g.H2 = in->h2;
  g.W2 = in->w2;
  g.maxh = maxh;
  g.maxw = maxw;
  g.bs = block_schedule(maxw);
  g.WB = slab_width(maxw);
  g.tiles_x = (g.W1 + kTW - 1) / kTW;
  g.tiles_y = (g.H1 + tile_rows - 1) / tile_rows;
  g.ntiles = g.tiles_x * g.tiles_y * g.N;
  out->CT = g.C <= 4 ? 4 : (g.C <= 10 ? 10 : 16);

  long long s1y = in->in1_stride_y ? in->in1_stride_y : g.W1;
  long long s1c = in->in1_stride_c ? in->in1_stride_c : (long long)g.H1 * s1y;
  long long s1n = in->in1_stride_n ? in->in1_stride_n : (long long)g.C * s1c;
  long long s2y = in->in2_stride_y ? in->in2_stride_y : g.W2;
  long long s2c = in->in2_stride_c ? in->in2_stride_c : (long long)g.H2 * s2y;
  long long s2n = in->in2_stride_n ? in->in2_stride_n : (long long)g.C * s2c;

  // frame 1: any strides work for the kernel; host data is copied as the smallest
  // enclosing span of the view.
  const float *d1 = in->in1;
  if (classify(in->in1) == PtrKind::Host) {
    const size_t span = (size_t)((g.N - 1) * s1n + (g.C - 1) * s1c + (g.H1 - 1) * s1y + g.W1);
    const size_t dense = (size_t)g.N * g.C * g.H1 * g.W1;
    if (span > dense + dense / 32 && s1y >= g.W1 && s1c % s1y == 0 && s1c / s1y >= g.H1 &&
        s1n == (long long)g.C * s1c) {
      // a cropped view (prepareInput's narrow): one 3-D copy moves only the rows of the crop
      // over PCIe and packs them
      void *buf = nullptr;
      DM_CHECK(call.alloc(&buf, dense * sizeof(float)));
      cudaMemcpy3DParms cp = {};
      cp.srcPtr = make_cudaPitchedPtr(const_cast<float *>(in->in1), (size_t)s1y * sizeof(float),
                                      (size_t)g.W1, (size_t)(s1c / s1y));
      cp.dstPtr = make_cudaPitchedPtr(buf, (size_t)g.W1 * sizeof(float), (size_t)g.W1, (size_t)g.H1);
      cp.extent = make_cudaExtent((size_t)g.W1 * sizeof(float), (size_t)g.H1, (size_t)g.N * g.C);
      cp.kind = cudaMemcpyHostToDevice;
      DM_CUDA(cudaMemcpy3DAsync(&cp, ctx->stream));
      ctx->call_has_host = true;
      d1 = static_cast<const float *>(buf);
      s1y = g.W1;
      s1c = (long long)g.H1 * g.W1;
      s1n = (long long)g.C * s1c;
    } else {
      const void *p = nullptr;
      DM_CHECK(call.in(in->in1, span * sizeof(float), &p));
      d1 = static_cast<const float *>(p);
    }
  }
  g.in1 = d1;
  g.s1n = s1n;
  g.s1c = s1c;
  g.s1y = s1y;

  // frame 2: TMA needs a 16-byte aligned base and 16-byte multiple strides.
  const float *d2 = in->in2;
  const bool host2 = classify(in->in2) == PtrKind::Host;
  const bool tma_ok = ((uintptr_t)in->in2 % 16 == 0) && (s2y % 4 == 0) && (s2c % 4 == 0) &&
                      (s2n % 4 == 0);
  const bool dense2 = s2y == g.W2 && s2c == (long long)g.H2 * g.W2 && s2n == (long long)g.C * s2c;
  if (host2 && dense2 && g.W2 % 4 == 0) {
    const void *p = nullptr;
    DM_CHECK(call.in(in->in2, (size_t)g.N * s2n * sizeof(float), &p));
    d2 = static_cast<const float *>(p);
  } else if (host2 || !tma_ok) {
    const long long py = (g.W2 + 3) & ~3LL;  // padded pitch
    void *buf = nullptr;
    DM_CHECK(call.alloc(&buf, (size_t)g.N * g.C * g.H2 * py * sizeof(float)));
    const cudaMemcpyKind kind = host2 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (s2c == (long long)g.H2 * s2y && s2n == (long long)g.C * s2c) {
      // planes follow each other row after row: one pitched copy for everything
      DM_CUDA(cudaMemcpy2DAsync(buf, py * sizeof(float), in->in2, s2y * sizeof(float),
                                g.W2 * sizeof(float), (size_t)g.N * g.C * g.H2, kind, ctx->stream));
    } else
    // one pitched copy per (n, c) plane keeps arbitrary strides simple
    for (int n = 0; n < g.N; ++n)
      for (int c = 0; c < g.C; ++c) {
        const float *src = in->in2 + n * s2n + c * s2c;
        float *dst = static_cast<float *>(buf) + ((size_t)n * g.C + c) * g.H2 * py;
        DM_CUDA(cudaMemcpy2DAsync(dst, py * sizeof(float), src, s2y * sizeof(float),
                                  g.W2 * sizeof(float), g.H2,
                                  host2 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                                  ctx->stream));
      }
    if (host2) ctx->call_has_host = true;
    d2 = static_cast<const float *>(buf);
    s2y = py;
    s2c = (long long)g.H2 * py;
    s2n = (long long)g.C * s2c;
  }
  out->in1_dev = d1;
  out->in2_dev = d2;
  out->s2n = s2n;
  out->s2c = s2c;
  out->s2y = s2y;
  out->Cin = g.C;
  const uint64_t dims[4] = {(uint64_t)g.W2, (uint64_t)g.H2, (uint64_t)g.C, (uint64_t)g.N};
  const uint64_t strides[3] = {(uint64_t)s2y * 4, (uint64_t)s2c * 4, (uint64_t)s2n * 4};
  const uint32_t box[4] = {(uint32_t)g.WB, 1u, (uint32_t)out->CT, 1u};
  DM_CHECK(tensor_map_4d(ctx, &out->tmap, d2, dims, strides, box));
  // the kernels address the ring with the box's channel count; slots start on 128-byte
  // boundaries (TMA destination alignment)
  g.C = out->CT;
  g.slab_floats = (g.C * g.WB + 31) & ~31;
  return DM_OK;
}

int generic_match_extract(Call &call, const dm_pair *in, int maxh, int maxw, unsigned flags,
                          double thr, int h_img, int w_img, const dm_extract_out *out);
int generic_match_volume(Call &call, const dm_pair *in, int maxh, int maxw, int mode, bool exact,
                         float *out);
// zero-flow column dx = cx - 1 in the skewed block layout: even pixels dx = 8*blk + r, odd
// pixels dx = 8*blk - 1 + r
static void set_middle(ExtractParams *P, int maxw) {
  const int dx = P->cx - 1;
  P->mid_blk[0] = dx / kR;
  P->mid_r[0] = dx % kR;
  P->mid_blk[1] = (dx + 1) / kR;
  P->mid_r[1] = (dx + 1) % kR;
  P->middle = P->mid_dy * maxw + P->cx;
}

// ring depth: as many slots as fit next to `extra` bytes, at most the configuration's, at
// least two more than the tile height (so that the producer can run ahead of the warps)
static int fit_ring(dm_ctx *ctx, SweepGeom *g, int tile_rows, int max_slots, size_t extra) {
  const size_t slab = (size_t)g->slab_floats * sizeof(float);
  long long n = ((long long)ctx->smem_optin - (long long)extra) / (long long)slab;
  if (n > max_slots) n = max_slots;
  if (n < tile_rows + 2) {
    set_error("window %dx%d with %d channels does not fit in shared memory (%zu-byte row slabs)",
              g->maxh, g->maxw, g->Cin, slab);
    return DM_ERR_UNSUPPORTED;
  }
  g->nslot = (int)n;
  return DM_OK;
}

// Can the tiled sweep take this shape?  The slab row must fit a TMA box (256 elements) and the
// ring needs tile_rows + 2 slots next to the epilogue's shared memory.  Otherwise the untiled
// kernel (match_generic.cu) does the job, slowly but for any window.
static bool tiled_fits(dm_ctx *ctx, int channels, int maxw, int tile_rows, size_t extra, bool with_norm_row) {
  const int CT = channels <= 4 ? 4 : (channels <= 10 ? 10 : 16);
  const int WB = slab_width(maxw);
  if (WB > 256) return false;
  size_t slab = (size_t)((CT * WB + 31) & ~31);
  if (with_norm_row) slab += (size_t)((WB + 31) & ~31);
  return (size_t)(tile_rows + 2) * slab * sizeof(float) + extra <= ctx->smem_optin;
}

static const void *pick_extract(bool small, int CT, int mode, int epi) {
#define DM_PICK3(cfg, ct, md)                                                                   \
  (epi == kEpiSoft ? (const void *)match_extract_kernel<cfg, ct, md, kEpiSoft>                  \
                   : (epi == kEpiWta ? (const void *)match_extract_kernel<cfg, ct, md, kEpiWta> \
                                     : (const void *)match_extract_kernel<cfg, ct, md, kEpiScores>))
#define DM_PICK(ct)                                                                                  \
  (small ? (mode == kExact ? DM_PICK3(ExtractCfgSmall, ct, kExact) : DM_PICK3(ExtractCfgSmall, ct, kFma)) \
         : (mode == kExact ? DM_PICK3(ExtractCfg, ct, kExact)                                        \
                           : (mode == kDot ? DM_PICK3(ExtractCfg, ct, kDot) : DM_PICK3(ExtractCfg, ct, kFma))))
  return CT == 4 ? DM_PICK(4) : (CT == 10 ? DM_PICK(10) : DM_PICK(16));
#undef DM_PICK
#undef DM_PICK3
}

// |x|^2 per pixel over the channels (one FMA chain, k ascending -- the sweep computes |a|^2 the
// same way) and the largest one, for the kDot bound.  out may be NULL (only the maximum).
__global__ void norm_kernel(const float *in, long long sn, long long sc, long long sy, int N, int C, int H,
                            int W, float *out, long long on, long long oy, unsigned *stat) {
  float mx = 0.0f;
  const long long total = (long long)N * H * W;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(t % W), y = (int)((t / W) % H), n = (int)(t / ((long long)W * H));
    const float *src = in + n * sn + y * sy + x;
    float s = 0.0f;
    for (int k = 0; k < C; ++k) {
      const float v = __ldg(src + k * sc);
      s = fmaf(v, v, s);
    }
    if (out) out[n * on + y * oy + x] = s;
    mx = fmaxf(mx, s);   // NaN inputs: fmaxf drops them here, the SSDs are NaN in either form
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.0f) atomicMax(stat, __float_as_uint(mx));
}

static int grid_for(dm_ctx *ctx, const void *kernel, int threads, size_t smem, int ntiles) {
  const int cap = ctx->num_sms * blocks_per_sm(ctx, kernel, threads, smem);
  return ntiles < cap ? ntiles : cap;
}

using S2Cfg = Sweep2Cfg<8, 6>;

// Launch of the two-row dot sweep.  DM_ERR_UNSUPPORTED (nothing launched, no error text) when the
// ring does not fit next to the shortlist bitmap: the caller falls back to the one-row kernel.
static int launch_sweep2(dm_ctx *ctx, const Prepared &pr, ExtractParams *Q, const CUtensorMap &nbmap) {
  SweepGeom &g = Q->g;
  g.tiles_y = (g.H1 + S2Cfg::kTH - 1) / S2Cfg::kTH;
  g.ntiles = g.tiles_x * g.tiles_y * g.N;
  if (Q->gb != 1) return DM_ERR_UNSUPPORTED;  // one shortlist bit per block only
  const size_t extra = kBar2Bytes + (size_t)Q->nwords * S2Cfg::kPx * S2Cfg::kCThreads * sizeof(unsigned);
  const size_t slab = (size_t)g.slab_floats * sizeof(float);
  long long nslot = ((long long)ctx->smem_optin - (long long)extra) / (long long)slab;
  if (nslot > S2Cfg::kNSlot) nslot = S2Cfg::kNSlot;
  if (nslot < S2Cfg::kTH + 2) return DM_ERR_UNSUPPORTED;
  g.nslot = (int)nslot;
  const size_t smem = ring_bytes(g, g.nslot) + extra;
  const bool wta = !Q->pmax && !Q->todo;
  // straight-line row body for the benchmark's window class (four full blocks + a 2-wide tail)
  const bool n8_4 = g.bs.n8 == 4 && g.bs.tail_r == 2 && ctx->opt.sweep != 3;
  const void *kfn;
#define DM_PICK2(ct, epi) (n8_4 ? (const void *)match_sweep2_kernel<S2Cfg, ct, epi, 4> : (const void *)match_sweep2_kernel<S2Cfg, ct, epi, 0>)
  if (pr.CT == 4)
    kfn = wta ? DM_PICK2(4, kEpiWta) : DM_PICK2(4, kEpiScores);
  else
    kfn = wta ? DM_PICK2(10, kEpiWta) : DM_PICK2(10, kEpiScores);
#undef DM_PICK2
  DM_CHECK(ensure_func_smem(ctx, kfn, smem));
  const int grid = g.ntiles < ctx->num_sms ? g.ntiles : ctx->num_sms;
  void *args[] = {(void *)&pr.tmap, (void *)&nbmap, (void *)Q};
  DM_CUDA(cudaLaunchKernel(kfn, dim3(grid), dim3(S2Cfg::kThreads), args, smem, ctx->stream));
  count_launch(ctx);
  return DM_OK;
}

}  // namespace dm

using namespace dm;

static int match_extract_impl(dm_ctx *ctx, const dm_pair *in, int maxh, int maxw, unsigned flags,
                              double prob_threshold, int h_img, int w_img, const dm_extract_out *out,
                              bool defer) {
  const bool want_thr = out->index_thr || out->score_thr || out->n_untouched;
  DM_REQUIRE(!want_thr || (prob_threshold > 0.1 && prob_threshold < 1.0),
             "dm_match_extract: fused thresholded extraction needs 0.1 < prob_threshold < 1 "
             "(got %g); use dm_match_volume + dm_extract_output for other thresholds",
             prob_threshold);
  if (out->flow_full)
    DM_REQUIRE(h_img >= in->h1 && w_img >= in->w1, "canvas %dx%d smaller than output %dx%d", h_img,
               w_img, in->h1, in->w1);
  Call call(ctx, defer);
  // worst-case epilogue footprint of the 15-row configuration (6 shortlist words)
  const size_t extra_max = kBarBytes + (size_t)7 * ExtractCfg::kCThreads * kP * sizeof(unsigned);
  if (in->channels > kMaxC || !tiled_fits(ctx, in->channels, maxw, ExtractCfg::kTH, extra_max, false)) {
    DM_REQUIRE(!out->conf_marginal, "dm_match_extract: conf_marginal is not available beyond %d channels or "
                                    "for windows the tiled kernel cannot hold", kMaxC);
    int rc = generic_match_extract(call, in, maxh, maxw, flags, prob_threshold, h_img, w_img, out);
    int rf = call.finish();
    return rc != DM_OK ? rc : rf;
  }
  // 15-row tiles leave most SMs idle on one small pair (fewer tiles than half the SMs): switch
  // to the 5-row configuration
  const long long big_tiles = (long long)((in->w1 + kTW - 1) / kTW) * ((in->h1 + ExtractCfg::kTH - 1) / ExtractCfg::kTH) *
                              in->n_pairs;
  const bool force_dot = ctx->opt.ssd_form == 2, force_diff = ctx->opt.ssd_form == 1;  // tuning and tests only
  // small calls (one 320x180 pair with a 17x17 window) are launch-bound: the norm pre-pass and the
  // twin launch of the dot form would cost more than it saves
  const bool big = (double)in->h1 * in->w1 * in->n_pairs * maxh * maxw * in->channels >= 5.0e8 ||
                   force_dot;
  const bool small = 2 * big_tiles < ctx->num_sms && !big && !ctx->opt.no_small_tiles;
  const int cfg_th = small ? ExtractCfgSmall::kTH : ExtractCfg::kTH;
  const int cfg_nslot = small ? ExtractCfgSmall::kNSlot : ExtractCfg::kNSlot;
  const int cfg_threads = small ? ExtractCfgSmall::kThreads : ExtractCfg::kThreads;
  const int cfg_cthreads = small ? ExtractCfgSmall::kCThreads : ExtractCfg::kCThreads;
  Prepared pr;
  DM_CHECK(prepare(call, in, maxh, maxw, cfg_th, &pr));
  const SweepGeom &g = pr.g;
  const size_t npx = (size_t)g.N * g.H1 * g.W1;
  DM_REQUIRE(npx < ((size_t)1 << 31), "dm_match_extract: %zu output pixels in one call; the pixel lists are "
                                      "32-bit, split the batch", npx);

  ExtractParams P;
  P.g = g;
  P.flags = flags;
  P.M = prob_threshold < 0.2 ? 8 : 4;
  // decide in the sweep only when it is safe by a margin; the rest goes to the exact pass
  P.p_none = (float)(prob_threshold - 1e-4);
  P.p_gt = (float)(prob_threshold + 1e-4);
  P.thr_lo = (float)(prob_threshold * 0.999);
  P.p_clear = (float)(prob_threshold * 0.99);
  P.cy = (maxh + 1) / 2;
  P.cx = (maxw + 1) / 2;
  P.mid_dy = P.cy - 1;
  set_middle(&P, maxw);
  P.h_img = h_img;
  P.w_img = w_img;
  P.hoff = (h_img - g.H1) / 2;
  P.woff = (w_img - g.W1) / 2;

  void *p;
#define DM_OUT(field, type, bytes)                                \
  P.field = nullptr;                                              \
  if (out->field) {                                               \
    DM_CHECK(call.out(out->field, (bytes), &p));                  \
    P.field = static_cast<type *>(p);                             \
  }
  DM_OUT(index, long long, npx * 8)
  DM_OUT(min_ssd, float, npx * 4)
  DM_OUT(pmax, float, npx * 4)
  DM_OUT(flow_full, float, (size_t)g.N * 2 * h_img * w_img * 4)
  DM_OUT(index_thr, long long, npx * 8)
  DM_OUT(score_thr, float, npx * 4)
  DM_OUT(soft_yx, float, npx * 2 * 4)
  DM_OUT(conf_marginal, float, npx * 4)
#undef DM_OUT
  P.p_thr = (float)prob_threshold;
  DM_REQUIRE(!P.conf_marginal || P.soft_yx, "dm_match_extract: conf_marginal belongs to the 'mean' extraction: "
                                           "ask for soft_yx too");
  P.n_untouched = nullptr;
  if (out->n_untouched) {
    DM_CHECK(call.out(out->n_untouched, (size_t)g.N * 8, &p));
    P.n_untouched = static_cast<unsigned long long *>(p);
    DM_CUDA(cudaMemsetAsync(p, 0, (size_t)g.N * 8, ctx->stream));
  }
  if (P.flow_full)
    DM_CUDA(cudaMemsetAsync(P.flow_full, 0, (size_t)g.N * 2 * h_img * w_img * 4, ctx->stream));

  P.nwords = 0;
  P.gb = 1;
  P.todo = nullptr;
  P.todo_mask = nullptr;
  P.ntodo = nullptr;
  P.vmin = P.vinv = nullptr;
  if (want_thr) {
    // one shortlist bit per (dy, dx-block); S >= 1 so p_k > thr needs v_k < min + ln(1/thr)
    const int nblocks = maxh * g.bs.per_row();
    P.gb = (nblocks + 191) / 192;  // at most 6 bitmap words per pixel
    P.nwords = ((nblocks + P.gb - 1) / P.gb + 31) / 32;
    void *scratch = nullptr;
    DM_CHECK(call.alloc(&scratch, 256 + npx * sizeof(int) * (1 + (size_t)P.nwords) + npx * 8));
    P.ntodo = static_cast<unsigned *>(scratch);
    P.todo = static_cast<int *>(scratch) + 64;
    P.todo_mask = reinterpret_cast<unsigned *>(P.todo + npx);
    P.vmin = reinterpret_cast<float *>(P.todo_mask + npx * P.nwords);
    P.vinv = P.vmin + npx;
    DM_CUDA(cudaMemsetAsync(P.ntodo, 0, sizeof(unsigned), ctx->stream));
  }
  const size_t extra = kBarBytes + (size_t)(P.nwords + 1) * cfg_cthreads * kP * sizeof(unsigned);
  const bool exact = flags & DM_FLAG_EXACT_SSD;
  P.stats = nullptr;
  P.dot_limit = 0.0f;
  P.tau_rel = 0.0f;
  P.dbg = 0;
  {
    // the rescore list (see ExtractParams::resc): worst case every pixel
    void *scratch = nullptr;
    DM_CHECK(call.alloc(&scratch, 256 + npx * sizeof(int)));
    P.nresc = static_cast<unsigned *>(scratch);
    P.resc = static_cast<int *>(scratch) + 64;
    DM_CUDA(cudaMemsetAsync(P.nresc, 0, sizeof(unsigned), ctx->stream));
  }
  P.in2 = pr.in2_dev;
  P.s2n = pr.s2n;
  P.s2c = pr.s2c;
  P.s2y = pr.s2y;
  DM_CHECK(fit_ring(ctx, &P.g, cfg_th, cfg_nslot, extra));
  auto launch = [&](const ExtractParams &Q, const CUtensorMap &nbmap, int mode) -> int {
    const size_t smem = ring_bytes(Q.g, Q.g.nslot) + extra;
    const bool wta = !Q.pmax && !Q.soft_yx && !Q.todo;  // index / flow / min_ssd only
    const void *kfn = pick_extract(small, pr.CT, mode, Q.soft_yx ? kEpiSoft : (wta ? kEpiWta : kEpiScores));
    DM_CHECK(ensure_func_smem(ctx, kfn, smem));
    const int grid = grid_for(ctx, kfn, cfg_threads, smem, Q.g.ntiles);
    void *args[] = {(void *)&pr.tmap, (void *)&nbmap, (void *)&Q};
    DM_CUDA(cudaLaunchKernel(kfn, dim3(grid), dim3(cfg_threads), args, smem, ctx->stream));
    count_launch(ctx);
    return DM_OK;
  };
  // The dot form (|a|^2 + |b|^2 - 2 a.b) halves the FP32 work of the sweep; its absolute error
  // is a few ulp of |a|^2 + |b|^2.  It is used when the largest norms keep that error under
  // ~1e-5 (the parity bar on the soft-max scores): a pre-pass writes |b|^2 per frame-2 pixel and
  // the maxima, then both kernels are launched and the device-side bound lets one of them run.
  const bool allow_dot = big && !small && !exact && !(flags & DM_FLAG_DIFF_SSD) && !force_diff;
  bool twin = false;
  ExtractParams Pd = P;
  CUtensorMap nbmap;
  if (allow_dot) {
    Pd.g.nb_off = (Pd.g.C * Pd.g.WB + 31) & ~31;
    Pd.g.slab_floats = Pd.g.nb_off + ((Pd.g.WB + 31) & ~31);
    twin = fit_ring(ctx, &Pd.g, cfg_th, cfg_nslot, extra) == DM_OK;
  }
  if (twin) {
    const long long w2p = (g.W2 + 3) & ~3LL;
    void *nbuf = nullptr, *stats = nullptr;
    DM_CHECK(call.alloc(&nbuf, (size_t)g.N * g.H2 * w2p * sizeof(float)));
    DM_CHECK(call.alloc(&stats, 256));
    DM_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(unsigned), ctx->stream));
    const int nthr = 256, nblk = ctx->num_sms * 8;
    prof_begin(ctx);
    norm_kernel<<<nblk, nthr, 0, ctx->stream>>>(g.in1, g.s1n, g.s1c, g.s1y, g.N, pr.Cin, g.H1, g.W1, nullptr, 0,
                                                0, static_cast<unsigned *>(stats));
    norm_kernel<<<nblk, nthr, 0, ctx->stream>>>(pr.in2_dev, pr.s2n, pr.s2c, pr.s2y, g.N, pr.Cin, g.H2, g.W2,
                                                static_cast<float *>(nbuf), (long long)g.H2 * w2p, w2p,
                                                static_cast<unsigned *>(stats) + 1);
    count_launch(ctx, 2);
    const uint64_t dims[4] = {(uint64_t)g.W2, (uint64_t)g.H2, 1u, (uint64_t)g.N};
    const uint64_t strides[3] = {(uint64_t)w2p * 4, (uint64_t)g.H2 * w2p * 4, (uint64_t)g.H2 * w2p * 4};
    const uint32_t box[4] = {(uint32_t)g.WB, 1u, 1u, 1u};
    DM_CHECK(tensor_map_4d(ctx, &nbmap, static_cast<const float *>(nbuf), dims, strides, box));
    P.stats = Pd.stats = static_cast<const unsigned *>(stats);
    // Error model of the dot form: na, nb and the C products are each rounded once at a magnitude
    // of at most |a|^2 + |b|^2, so |v_dot - v| <= e_rel * (|a|^2 + |b|^2) with e_rel = (C + 2) * 2^-24
    // (a 4M-sample fp32 simulation at C = 10 peaks at 0.85 of that bound).  Scores: the soft-max
    // sums move by the same relative amount, so the 1e-4 bar on scores allows norms up to
    // 1e-4 / e_rel (140 for 10 channels).  Indices: entries closer than tau = 4 * e_rel * norms
    // (twice the sum of two errors) to the minimum are not ordered here but in generic_rescore.
    const float e_rel = (float)(pr.Cin + 2) * 5.9604645e-8f;
    P.dot_limit = Pd.dot_limit = force_dot ? 3.0e38f : 1.0e-4f / e_rel;
    Pd.tau_rel = 4.0f * e_rel;
    Pd.dbg = ctx->opt.volume_debug;
    Pd.nb = static_cast<const float *>(nbuf);
    Pd.nb_sn = (long long)g.H2 * w2p;
    Pd.nb_sy = w2p;
    // the dot leg: the one-row sweep in its dot form.  The two-rows-per-warp variant
    // (match_sweep2.cuh, option sweep = 2) halves the shared-memory wavefronts per output but runs
    // two warps per scheduler in lock-step phases; measured slower on B200, so it
    // is opt-in.
    bool two_rows = false;
    if (!Pd.soft_yx && pr.CT <= 10 && (ctx->opt.sweep == 2 || ctx->opt.sweep == 3)) {
      ExtractParams P2 = Pd;
      int rc2 = launch_sweep2(ctx, pr, &P2, nbmap);
      if (rc2 == DM_OK) two_rows = true;
      else if (rc2 != DM_ERR_UNSUPPORTED) return rc2;
    }
    if (!two_rows) DM_CHECK(launch(Pd, nbmap, kDot));
    DM_CHECK(launch(P, pr.tmap, kFma));
    prof_end(ctx);
  } else {
    prof_begin(ctx);
    DM_CHECK(launch(P, pr.tmap, exact ? kExact : kFma));
    prof_end(ctx);
  }
  {
    GenericParams G;
    memset(&G, 0, sizeof(G));
    G.in1 = g.in1; G.s1n = g.s1n; G.s1c = g.s1c; G.s1y = g.s1y;
    G.in2 = pr.in2_dev; G.s2n = pr.s2n; G.s2c = pr.s2c; G.s2y = pr.s2y;
    G.N = g.N; G.C = pr.Cin; G.H1 = g.H1; G.W1 = g.W1; G.maxh = maxh; G.maxw = maxw;
    G.flags = flags;
    G.thr = prob_threshold;
    G.M = P.M; G.middle = P.middle; G.cy = P.cy; G.cx = P.cx;
    G.h_img = h_img; G.w_img = w_img; G.hoff = P.hoff; G.woff = P.woff;
    G.index = P.index; G.min_ssd = P.min_ssd; G.pmax = P.pmax; G.flow_full = P.flow_full;
    G.index_thr = P.index_thr; G.score_thr = P.score_thr; G.soft_yx = P.soft_yx;
    G.n_untouched = P.n_untouched; G.conf_marginal = P.conf_marginal;
    DM_CHECK(generic_rescore(ctx, G, P.resc, P.nresc));
  }
  if (want_thr) {
    ThresholdPass T;
    T.in1 = g.in1; T.s1n = g.s1n; T.s1c = g.s1c; T.s1y = g.s1y;
    T.in2 = pr.in2_dev; T.s2n = pr.s2n; T.s2c = pr.s2c; T.s2y = pr.s2y;
    T.N = g.N; T.C = pr.Cin; T.H1 = g.H1; T.W1 = g.W1; T.maxh = maxh; T.maxw = maxw;
    T.bs = g.bs; T.nwords = P.nwords; T.gb = P.gb; T.M = P.M; T.exact = exact ? 1 : 0;
    T.thr = prob_threshold;
    T.mark = ctx->opt.debug_todo ? 1 : 0;
    T.todo = P.todo; T.todo_mask = P.todo_mask; T.ntodo = P.ntodo;
    T.vmin = P.vmin; T.vinv = P.vinv;
    T.index_thr = P.index_thr; T.score_thr = P.score_thr; T.n_untouched = P.n_untouched;
    if (pr.CT == 4)
      threshold_exact_kernel<4><<<ctx->num_sms * 12, 128, 0, ctx->stream>>>(T);
    else if (pr.CT == 10)
      threshold_exact_kernel<10><<<ctx->num_sms * 12, 128, 0, ctx->stream>>>(T);
    else
      threshold_exact_kernel<16><<<ctx->num_sms * 12, 128, 0, ctx->stream>>>(T);
    DM_CUDA(cudaGetLastError());
    count_launch(ctx);
    if (ctx->opt.debug_todo) {  // diagnostics: how many pixels needed the exact pass
      unsigned n = 0;
      cudaMemcpyAsync(&n, P.ntodo, sizeof(n), cudaMemcpyDeviceToHost, ctx->stream);
      cudaStreamSynchronize(ctx->stream);
      std::vector<unsigned> hm((size_t)n * P.nwords);
      cudaMemcpy(hm.data(), P.todo_mask, hm.size() * sizeof(unsigned), cudaMemcpyDeviceToHost);
      unsigned long long bits = 0;
      for (unsigned v : hm) bits += (unsigned long long)__builtin_popcount(v);
      fprintf(stderr, "[depthmatch] exact pass: %u of %zu pixels, %.1f shortlisted blocks per pixel\n", n, npx,
              n ? (double)bits / n : 0.0);
    }
  }
  // the counters of this call, for dm_last_counts (the scratch they live in is recycled)
  DM_CUDA(cudaMemcpyAsync(ctx->counters, P.nresc, sizeof(unsigned), cudaMemcpyDeviceToDevice, ctx->stream));
  if (P.ntodo)
    DM_CUDA(cudaMemcpyAsync(ctx->counters + 1, P.ntodo, sizeof(unsigned), cudaMemcpyDeviceToDevice, ctx->stream));
  else
    DM_CUDA(cudaMemsetAsync(ctx->counters + 1, 0, sizeof(unsigned), ctx->stream));
  ctx->counters_valid = true;
  return call.finish();
}

// Host-buffer batches are cut in chunks that alternate between two private sub-contexts, so
// the H2D copy of one chunk overlaps the kernels of the previous one and the D2H of the one
// before (full-duplex PCIe + compute); device-buffer calls go straight through.
extern "C" int dm_match_extract(dm_ctx *ctx, const dm_pair *in, int maxh, int maxw, unsigned flags,
                                double prob_threshold, int h_img, int w_img,
                                const dm_extract_out *out) {
  DM_REQUIRE(ctx && in && out, "dm_match_extract: NULL argument");
  DM_CUDA(cudaSetDevice(ctx->device));
  const bool host_in = in->in1 && in->in2 && classify(in->in1) == PtrKind::Host &&
                       classify(in->in2) == PtrKind::Host;
  const int N = in->n_pairs;
  const bool async = (flags & DM_FLAG_ASYNC) != 0;  // the caller waits with dm_synchronize()
  if (!host_in || N < 4 || ctx->is_child || ctx->opt.no_pipeline)
    return match_extract_impl(ctx, in, maxh, maxw, flags, prob_threshold, h_img, w_img, out, async);
  for (int i = 0; i < 2; ++i)
    if (!ctx->pipe[i]) {
      DM_CHECK(dm_create(ctx->device, &ctx->pipe[i]));
      ctx->pipe[i]->is_child = true;
      ctx->pipe[i]->opt = ctx->opt;
    }
  const long long s1y = in->in1_stride_y ? in->in1_stride_y : in->w1;
  const long long s1c = in->in1_stride_c ? in->in1_stride_c : (long long)in->h1 * s1y;
  const long long s1n = in->in1_stride_n ? in->in1_stride_n : (long long)in->channels * s1c;
  const long long s2y = in->in2_stride_y ? in->in2_stride_y : in->w2;
  const long long s2c = in->in2_stride_c ? in->in2_stride_c : (long long)in->h2 * s2y;
  const long long s2n = in->in2_stride_n ? in->in2_stride_n : (long long)in->channels * s2c;
  const size_t npx1 = (size_t)in->h1 * in->w1, canvas = (size_t)2 * h_img * w_img;
  int chunk = N >= 16 ? 4 : (N >= 8 ? 2 : 1);  // 4 north pairs = 3 full waves of tiles on 148 SMs
  if (ctx->opt.pipe_chunk > 0) chunk = ctx->opt.pipe_chunk;
  int rc = DM_OK;
  for (int n0 = 0, c = 0; n0 < N && rc == DM_OK; n0 += chunk, ++c) {
    dm_pair sub = *in;
    sub.n_pairs = N - n0 < chunk ? N - n0 : chunk;
    sub.in1 = in->in1 + n0 * s1n;
    sub.in2 = in->in2 + n0 * s2n;
    sub.in1_stride_n = s1n; sub.in1_stride_c = s1c; sub.in1_stride_y = s1y;
    sub.in2_stride_n = s2n; sub.in2_stride_c = s2c; sub.in2_stride_y = s2y;
    dm_extract_out so = *out;
    if (so.index) so.index += n0 * npx1;
    if (so.min_ssd) so.min_ssd += n0 * npx1;
    if (so.pmax) so.pmax += n0 * npx1;
    if (so.flow_full) so.flow_full += n0 * canvas;
    if (so.index_thr) so.index_thr += n0 * npx1;
    if (so.score_thr) so.score_thr += n0 * npx1;
    if (so.soft_yx) so.soft_yx += n0 * 2 * npx1;
    if (so.conf_marginal) so.conf_marginal += n0 * npx1;
    if (so.n_untouched) so.n_untouched += n0;
    rc = match_extract_impl(ctx->pipe[c & 1], &sub, maxh, maxw, flags, prob_threshold, h_img, w_img, &so,
                            true);
  }
  for (int i = 0; i < 2 && !async; ++i) {
    cudaError_t e = cudaStreamSynchronize(ctx->pipe[i]->stream);
    if (e != cudaSuccess && rc == DM_OK) rc = cuda_fail(e, "pipeline synchronize", __FILE__, __LINE__);
  }
  return rc;
}

// ---------------------------------------------------------------- volume API
namespace dm {

// statistics pass for the soft-max volume: per-pixel min and 1/sum, nothing else
static int launch_stats(Call &call, const Prepared &pr, int ssd_mode, bool small, float *vmin, float *vinv,
                        const CUtensorMap *nbmap = nullptr, const unsigned *stats = nullptr, float dot_limit = 0.0f,
                        const float *in2 = nullptr) {
  dm_ctx *ctx = call.ctx;
  const SweepGeom &g = pr.g;
  const int cfg_th = small ? ExtractCfgSmall::kTH : ExtractCfg::kTH;
  const int cfg_nslot = small ? ExtractCfgSmall::kNSlot : ExtractCfg::kNSlot;
  const int cfg_threads = small ? ExtractCfgSmall::kThreads : ExtractCfg::kThreads;
  const int cfg_cthreads = small ? ExtractCfgSmall::kCThreads : ExtractCfg::kCThreads;
  ExtractParams P;
  memset(&P, 0, sizeof(P));
  P.g = g;
  P.flags = 0;
  P.M = 8;
  P.cy = (g.maxh + 1) / 2;
  P.cx = (g.maxw + 1) / 2;
  P.mid_dy = P.cy - 1;
  set_middle(&P, g.maxw);
  P.min_ssd = vmin;
  P.pmax = vinv;
  P.stats = stats;
  P.dot_limit = dot_limit;
  if (ssd_mode == kDot) {   // statistics in the dot form: no rescore list, the volume pass uses the same sums
    P.g.nb_off = (P.g.C * P.g.WB + 31) & ~31;
    P.g.slab_floats = P.g.nb_off + ((P.g.WB + 31) & ~31);
    (void)in2;            // P.in2 stays NULL: no re-score of the winner, see ExtractEpi::tile_rescore
  }
  const size_t extra = kBarBytes + (size_t)cfg_cthreads * kP * sizeof(unsigned);
  DM_CHECK(fit_ring(ctx, &P.g, cfg_th, cfg_nslot, extra));
  const size_t smem = ring_bytes(P.g, P.g.nslot) + extra;
  const void *kfn = pick_extract(small, pr.CT, ssd_mode, kEpiScores);
  DM_CHECK(ensure_func_smem(ctx, kfn, smem));
  const int grid = grid_for(ctx, kfn, cfg_threads, smem, g.ntiles);
  void *args[] = {(void *)&pr.tmap, (void *)(nbmap ? nbmap : &pr.tmap), (void *)&P};
  DM_CUDA(cudaLaunchKernel(kfn, dim3(grid), dim3(cfg_threads), args, smem, ctx->stream));
  count_launch(ctx);
  return DM_OK;
}

}  // namespace dm

namespace dm {
// the body of dm_match_volume on an open Call (dm_multiscale_extract runs several of these
// inside one call so that their outputs can live in the call's arena)
int match_volume_on(Call &call, const dm_pair *in, int maxh, int maxw, int mode, float *out) {
  dm_ctx *ctx = call.ctx;
  const bool exact = (mode & DM_VOLUME_EXACT) != 0;
  mode &= ~DM_VOLUME_EXACT;
  DM_REQUIRE(mode == DM_VOLUME_SSD || mode == DM_VOLUME_NEG_SOFTMAX, "dm_match_volume: bad mode %d",
             mode);
  if (in->channels > kMaxC ||
      !tiled_fits(ctx, in->channels, maxw, ExtractCfg::kTH,
                  kBarBytes + (size_t)VolumeCfg::kWarps * 2 * kStgPlane * sizeof(float), false))
    return generic_match_volume(call, in, maxh, maxw, mode, exact, out);
  Prepared pr;
  DM_CHECK(prepare(call, in, maxh, maxw, VolumeCfg::kTH, &pr));
  const SweepGeom &g = pr.g;
  const size_t npx = (size_t)g.N * g.H1 * g.W1;
  const size_t K = (size_t)maxh * maxw;
  void *p = nullptr;
  DM_CHECK(call.out(out, npx * K * sizeof(float), &p));
  VolumeParams P;
  P.g = g;
  P.mode = mode;
  P.vmin = P.vinv = nullptr;
  P.out = static_cast<float *>(p);
  P.debug = ctx->opt.volume_debug;
  P.stats = nullptr;
  P.dot_limit = 0.0f;
  // soft-max volume of a large call: both sweeps in the dot form (half the FP32 work; the outputs
  // are probabilities, whose 1e-4 bar the form meets under the same norm bound as the fused
  // kernel -- no index is produced here, so there is nothing to rescore)
  const bool big = (double)g.H1 * g.W1 * g.N * maxh * maxw * pr.Cin >= 5.0e8;
  bool dot = mode == DM_VOLUME_NEG_SOFTMAX && !exact && big && ctx->opt.ssd_form != 1;
  CUtensorMap nbmap = pr.tmap;
  const unsigned *stats = nullptr;
  float dot_limit = 0.0f;
  SweepGeom gdot = g;
  const float *nbuf_dev = nullptr;
  if (dot) {
    gdot.nb_off = (g.C * g.WB + 31) & ~31;
    gdot.slab_floats = gdot.nb_off + ((g.WB + 31) & ~31);
    const size_t extra_v = kBarBytes + (size_t)VolumeCfg::kWarps * 2 * kStgPlane * sizeof(float);
    dot = fit_ring(ctx, &gdot, VolumeCfg::kTH, VolumeCfg::kNSlot, extra_v) == DM_OK;
  }
  if (dot) {
    const long long w2p = (g.W2 + 3) & ~3LL;
    void *nbuf = nullptr, *st = nullptr;
    DM_CHECK(call.alloc(&nbuf, (size_t)g.N * g.H2 * w2p * sizeof(float)));
    DM_CHECK(call.alloc(&st, 256));
    DM_CUDA(cudaMemsetAsync(st, 0, 2 * sizeof(unsigned), ctx->stream));
    const int nthr = 256, nblk = ctx->num_sms * 8;
    norm_kernel<<<nblk, nthr, 0, ctx->stream>>>(g.in1, g.s1n, g.s1c, g.s1y, g.N, pr.Cin, g.H1, g.W1, nullptr, 0, 0,
                                                static_cast<unsigned *>(st));
    norm_kernel<<<nblk, nthr, 0, ctx->stream>>>(pr.in2_dev, pr.s2n, pr.s2c, pr.s2y, g.N, pr.Cin, g.H2, g.W2,
                                                static_cast<float *>(nbuf), (long long)g.H2 * w2p, w2p,
                                                static_cast<unsigned *>(st) + 1);
    count_launch(ctx, 2);
    const uint64_t dims[4] = {(uint64_t)g.W2, (uint64_t)g.H2, 1u, (uint64_t)g.N};
    const uint64_t strides[3] = {(uint64_t)w2p * 4, (uint64_t)g.H2 * w2p * 4, (uint64_t)g.H2 * w2p * 4};
    const uint32_t box[4] = {(uint32_t)g.WB, 1u, 1u, 1u};
    DM_CHECK(tensor_map_4d(ctx, &nbmap, static_cast<const float *>(nbuf), dims, strides, box));
    nbuf_dev = static_cast<const float *>(nbuf);
    stats = static_cast<const unsigned *>(st);
    dot_limit = ctx->opt.ssd_form == 2 ? 3.0e38f : 1.0e-4f / ((float)(pr.Cin + 2) * 5.9604645e-8f);
  }
  // The strip kernel (match_volume_px.cuh) writes whole pixel streams with bulk copies; it needs
  // 16-byte aligned runs of 4 pixels (W1 % 4 == 0) and two staging buffers of 16 streams next to a
  // ring of maxh + 3 rows.  What it cannot take stays with the tiled kernel above.
  PxGeom X = {};
  CUtensorMap pxmap = pr.tmap, pxnb = pr.tmap;
  bool strip = ctx->opt.volume_kernel != 1 && g.W1 % 4 == 0 && (reinterpret_cast<uintptr_t>(P.out) % 16) == 0;
  size_t px_smem = 0;
  if (strip) {
    const BlockSchedule bs = g.bs;
    const bool wide = bs.tail_r == kR;
    X.WBs = kPxW - kP + bs.n8 * kR + (wide ? kNB : kP);
    X.nb_off = (pr.CT * X.WBs + 31) & ~31;
    X.pitch = dot ? X.nb_off + ((X.WBs + 31) & ~31) : X.nb_off;
    X.nslot = maxh + kPxAhead;
    X.nwide = bs.n8 + (wide ? 1 : 0);
    const int rem = X.nwide % 4;
    X.nbp = 2 * (X.nwide / 4) + (rem + 1) / 2;
    // window rows in groups of four; one or two rows left over by a four-block-wide window (33 = 8 * 4 + 1)
    // are one item of their own instead of a mostly idle group
    X.nrow = (X.nwide == 4 && maxh >= 4 && (maxh % 4 == 1 || maxh % 4 == 2)) ? 1 : 0;
    X.ndg = X.nrow ? maxh / 4 : (maxh + 3) / 4;
    X.nfull = X.ndg * X.nbp;
    X.items = X.nfull + X.nrow + (wide ? 0 : (maxh + 7) / 8);
    X.strips = (g.W1 + kPxW - 1) / kPxW;
    X.ncw = kPxMaxWarps;
    {
      auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
      X.rot = X.items;
      while (gcd(X.rot, X.ncw) != 1) ++X.rot;
    }
    px_smem = ((size_t)X.nslot * X.pitch + 2 * (size_t)kPxW * K + 4 + (size_t)kPxARing * (pr.CT * kPxW + 2 * kPxW)) * sizeof(float) +
              (2 * kPxMaxSlot + 6 + kPxARing) * sizeof(uint64_t);
    strip = strip && X.WBs <= 256 && X.nslot <= kPxMaxSlot && X.nwide >= 1 && px_smem <= ctx->smem_optin;
    // a step must hold enough items to keep the warps busy between two hand-overs: small windows (the
    // multiscale 8x8: 3 items per step, measured 0.38 against 0.33 ms) stay with the tiled kernel
    if (X.items < 2 * kPxMaxWarps && ctx->opt.volume_kernel != 2) strip = false;
  }
  if (strip) {
    // rows per unit: the split of the strips into row bands that fills the CTAs' waves best (a unit
    // pays about three steps for its first maxh - 1 rows)
    const long long per_band = (long long)g.N * X.strips;
    long long best = -1;
    for (int nb = 1; nb <= g.H1 && nb <= 64; ++nb) {
      const int band = (g.H1 + nb - 1) / nb;
      if (band < 8 && nb > 1) break;
      const long long waves = (per_band * nb + ctx->num_sms - 1) / ctx->num_sms;
      const long long cost = waves * (band + 3);
      if (best < 0 || cost < best) {
        best = cost;
        X.band = band;
      }
    }
    X.nbands = (g.H1 + X.band - 1) / X.band;
    X.units = (int)(per_band * X.nbands);
    const uint64_t dims[4] = {(uint64_t)g.W2, (uint64_t)g.H2, (uint64_t)pr.Cin, (uint64_t)g.N};
    const uint64_t strides[3] = {(uint64_t)pr.s2y * 4, (uint64_t)pr.s2c * 4, (uint64_t)pr.s2n * 4};
    const uint32_t box[4] = {(uint32_t)X.WBs, 1u, (uint32_t)pr.CT, 1u};
    DM_CHECK(tensor_map_4d(ctx, &pxmap, pr.in2_dev, dims, strides, box));
    if (dot) {
      const long long w2p = (g.W2 + 3) & ~3LL;
      const uint64_t ndims[4] = {(uint64_t)g.W2, (uint64_t)g.H2, 1u, (uint64_t)g.N};
      const uint64_t nstr[3] = {(uint64_t)w2p * 4, (uint64_t)g.H2 * w2p * 4, (uint64_t)g.H2 * w2p * 4};
      const uint32_t nbox[4] = {(uint32_t)X.WBs, 1u, 1u, 1u};
      DM_CHECK(tensor_map_4d(ctx, &pxnb, nbuf_dev, ndims, nstr, nbox));
    }
  }
  // soft-max volume on the strip kernel: minimum and sum are taken in the staging buffer, no statistics sweep
  const bool fused_softmax = strip && mode == DM_VOLUME_NEG_SOFTMAX && !exact && K <= (size_t)kPxKRegs * 32 &&
                             !(ctx->opt.volume_debug & 32);
  if (mode == DM_VOLUME_NEG_SOFTMAX && !fused_softmax) {
    void *s = nullptr;
    DM_CHECK(call.alloc(&s, npx * 2 * sizeof(float)));
    float *vmin = static_cast<float *>(s), *vinv = vmin + npx;
    Prepared ps = pr;  // the statistics sweep runs with the extraction kernel's tile height
    const long long big_tiles = (long long)g.tiles_x * ((g.H1 + ExtractCfg::kTH - 1) / ExtractCfg::kTH) * g.N;
    const bool small = 2 * big_tiles < ctx->num_sms && !ctx->opt.no_small_tiles && !dot;  // e.g. the coarse scales
    const int th = small ? ExtractCfgSmall::kTH : ExtractCfg::kTH;
    ps.g.tiles_y = (g.H1 + th - 1) / th;
    ps.g.ntiles = ps.g.tiles_x * ps.g.tiles_y * g.N;
    if (dot) {
      // twin launch: the norm bound, read on the device, lets exactly one of the two run
      DM_CHECK(launch_stats(call, ps, kDot, small, vmin, vinv, &nbmap, stats, dot_limit, pr.in2_dev));
      DM_CHECK(launch_stats(call, ps, kFma, small, vmin, vinv, nullptr, stats, dot_limit));
    } else {
      DM_CHECK(launch_stats(call, ps, exact ? kExact : kFma, small, vmin, vinv));
    }
    P.vmin = vmin;
    P.vinv = vinv;
  }
  const size_t extra = kBarBytes + (size_t)VolumeCfg::kWarps * 2 * kStgPlane * sizeof(float);
  DM_CHECK(fit_ring(ctx, &P.g, VolumeCfg::kTH, VolumeCfg::kNSlot, extra));
  auto launch_vol = [&](const VolumeParams &Q, const CUtensorMap &nb, int ssd_mode) -> int {
    const size_t smem = ring_bytes(Q.g, Q.g.nslot) + extra;
#define DM_PICKV(ct)                                                                                      \
  (ssd_mode == kExact ? (const void *)match_volume_kernel<ct, kExact>                                     \
                      : (ssd_mode == kDot ? (const void *)match_volume_kernel<ct, kDot> : (const void *)match_volume_kernel<ct, kFma>))
    const void *kfn = pr.CT == 4 ? DM_PICKV(4) : (pr.CT == 10 ? DM_PICKV(10) : DM_PICKV(16));
#undef DM_PICKV
    DM_CHECK(ensure_func_smem(ctx, kfn, smem));
    const int grid = grid_for(ctx, kfn, VolumeCfg::kThreads, smem, Q.g.ntiles);
    void *args[] = {(void *)&pr.tmap, (void *)&nb, (void *)&Q};
    DM_CUDA(cudaLaunchKernel(kfn, dim3(grid), dim3(VolumeCfg::kThreads), args, smem, ctx->stream));
    count_launch(ctx);
    return DM_OK;
  };
  auto launch_px = [&](const VolumeParams &Q, int ssd_mode) -> int {
#define DM_PICKX(ct)                                                                                           \
  (ssd_mode == kExact ? (const void *)match_volume_px_kernel<ct, kExact, false>                                \
   : ssd_mode == kDot ? (fused_softmax ? (const void *)match_volume_px_kernel<ct, kDot, true>                  \
                                       : (const void *)match_volume_px_kernel<ct, kDot, false>)                \
                      : (fused_softmax ? (const void *)match_volume_px_kernel<ct, kFma, true>                  \
                                       : (const void *)match_volume_px_kernel<ct, kFma, false>))
    const void *kfn = pr.CT == 4 ? DM_PICKX(4) : (pr.CT == 10 ? DM_PICKX(10) : DM_PICKX(16));
#undef DM_PICKX
    DM_CHECK(ensure_func_smem(ctx, kfn, px_smem));
    const int grid = X.units < ctx->num_sms ? X.units : ctx->num_sms;
    void *args[] = {(void *)&pxmap, (void *)&pxnb, (void *)&Q, (void *)&X};
    DM_CUDA(cudaLaunchKernel(kfn, dim3(grid), dim3((X.ncw + 2) * 32), args, px_smem, ctx->stream));
    count_launch(ctx);
    return DM_OK;
  };
  prof_begin(ctx);
  if (dot) {
    VolumeParams Pd = P;
    Pd.g = gdot;
    Pd.stats = P.stats = stats;
    Pd.dot_limit = P.dot_limit = dot_limit;
    if (strip) {
      DM_CHECK(launch_px(Pd, kDot));
      DM_CHECK(launch_px(P, kFma));
    } else {
      DM_CHECK(launch_vol(Pd, nbmap, kDot));
      DM_CHECK(launch_vol(P, pr.tmap, kFma));
    }
  } else if (strip) {
    DM_CHECK(launch_px(P, exact ? kExact : kFma));
  } else {
    DM_CHECK(launch_vol(P, pr.tmap, exact ? kExact : kFma));
  }
  prof_end(ctx);
  return DM_OK;
}
}  // namespace dm

extern "C" int dm_match_volume(dm_ctx *ctx, const dm_pair *in, int maxh, int maxw, int mode,
                               float *out) {
  DM_REQUIRE(ctx && in && out, "dm_match_volume: NULL argument");
  DM_CUDA(cudaSetDevice(ctx->device));
  Call call(ctx);
  const int rc = match_volume_on(call, in, maxh, maxw, mode, out);
  const int rf = call.finish();
  return rc != DM_OK ? rc : rf;
}
