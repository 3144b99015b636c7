// match_sweep2.cuh -- the dot-form sweep with TWO output rows per warp and a branch-free epilogue.
// Included by match_fused.cu after ExtractParams.
//
// Why: with one output row per warp (match_kernels.cuh) a thread turns 10 slab floats per channel
// into 32 partial sums, and the dot form needs only one packed FFMA2 per two of them: the kernel
// is then bound by shared-memory wavefronts (ncu r01: 84 % LDS, FP32 pipe 60 %).  Output rows y and
// y+1 need the same frame-2 row at window rows dy and dy-1, so a warp that owns both rows reads
// every slab value once for 64 partial sums: half the wavefronts per output, and the FP32 pipe
// becomes the bound again.
//
// Work decomposition: CTA = NW warps, ALL of them consumers = a tile of 2*NW output rows x 128
// columns; warp w owns rows y0+2w and y0+2w+1, lane l four consecutive pixels of both (their C
// frame-1 values, pre-multiplied by -2, stay in registers: 2 x 4 x C).  The kernel needs ~230
// registers, i.e. two warps per scheduler: eight warps fill the four schedulers evenly, a ninth
// (a dedicated TMA producer) would cap every warp at 168 registers.  So the ring is refilled by
// the consumers themselves: the warp whose arrival completes a slot's `empty` phase (it sees the
// phase complete right after its own arrive; a compare-and-swap on a per-slot counter makes it
// exactly one) issues the TMA load of the row that takes the slot next.  At slab row j the warp
// computes window row d0 = j - 2w for its first output row and d0 - 1 for the second.
//
// The per-row code of the common case (both output rows inside their windows) is straight-line:
// the blocks of a window row are unrolled and the epilogue has no branch, so that the scheduler
// can issue one block's epilogue (ALU / XU pipes) under the next block's FFMA2 stream -- with two
// warps per scheduler there is nobody else to hide it.
//
// SSD form: v' = |b|^2 - 2 a.b  (the norm row rides in the ring as one more "channel" whose weight
// is the constant 1).  |a|^2 is constant per pixel, so minima, their order and the soft-max
// differences are those of v' itself; it is only added back for the reported min_ssd.  Error
// model and the tau band as in match_extract_impl.
//
// Epilogue, per (pixel, 8-block), no data-dependent branch:
//   b    = min of the block's 8 entries
//   m2   = min(m2, max(m, b))          second smallest BLOCK MINIMUM so far
//   wb   = b < m ? block id : wb       winning block (strict <: first occurrence)
//   m    = min(m, b)
//   scores:  sc = 2^((m_new - m_old) log2e);  S = S*sc + sum_r 2^((m_new - v_r) log2e)
// and, for the thresholded extraction, one shortlist bit per block whose largest term could end
// above the threshold (the only branch; taken for a few blocks per pixel).
// At tile end each pixel re-computes its winning block (same operation sequence, so bit-identical
// values) to find the first position of the minimum and the runner-up inside that block; a second
// entry within tau of the minimum sends the pixel to the rescore list (ExtractParams::resc).
#pragma once

namespace dm {

template <int NW, int PF>
struct Sweep2Cfg {
  static constexpr int kWarps = NW;
  static constexpr int kCThreads = NW * 32;
  static constexpr int kThreads = kCThreads;  // no producer warp
  static constexpr int kTH = 2 * NW;
  static constexpr int kNSlot = 2 * NW + PF;
  static constexpr int kPx = 2 * kP;  // pixels per thread
};

__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// packed block of both rows: acc2[row][pp][j] = sum_k a2[row][k][pp] * b[k][2pp + j] + nb[2pp + j]
template <int CT, int JW>
__device__ __forceinline__ void dot2_block(const float2 (&a2)[2][CT][2], const float *bsrc, const float *nbsrc,
                                           int WB, float2 (&acc2)[2][2][JW]) {
  constexpr int NBF = JW == kR ? kNB : kP;
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
    for (int row = 0; row < 2; ++row)
#pragma unroll
      for (int pp = 0; pp < 2; ++pp)
#pragma unroll
        for (int j = 0; j < JW; ++j) {
          const float2 bb = make_float2(b[2 * pp + j], b[2 * pp + j]);
          acc2[row][pp][j] = k == 0 ? __fmul2_rn(a2[row][0][pp], bb) : __ffma2_rn(a2[row][k][pp], bb, acc2[row][pp][j]);
        }
  }
  {
    float nb[NBF];
    const float4 *src = reinterpret_cast<const float4 *>(nbsrc);
#pragma unroll
    for (int j = 0; j < NBF / 4; ++j) {
      const float4 t = src[j];
      nb[4 * j + 0] = t.x;
      nb[4 * j + 1] = t.y;
      nb[4 * j + 2] = t.z;
      nb[4 * j + 3] = t.w;
    }
    const float2 one = make_float2(1.0f, 1.0f);
#pragma unroll
    for (int row = 0; row < 2; ++row)
#pragma unroll
      for (int pp = 0; pp < 2; ++pp)
#pragma unroll
        for (int j = 0; j < JW; ++j)
          acc2[row][pp][j] = __ffma2_rn(one, make_float2(nb[2 * pp + j], nb[2 * pp + j]), acc2[row][pp][j]);
  }
}

// EPI: kEpiWta or kEpiScores (kEpiSoft stays with the one-row kernel).
template <class Cfg, int CT, int EPI>
struct Epi2 {
  static constexpr bool WTA = EPI == kEpiWta;
  static constexpr int kCThreads = Cfg::kCThreads;
  static constexpr int NQ = 2 * kP;  // q = row * 4 + p
  const ExtractParams &P;
  float m[NQ], m2[NQ];
  int wb[NQ];
  float S[WTA ? 1 : NQ];
  int vfrom[WTA ? 1 : NQ];
  unsigned rowbits[WTA ? 1 : NQ];  // shortlist bits of the window row in progress (bit = block of the row)
  float tau;
  unsigned *mask;  // [nwords][NQ][kCThreads] words, this thread's column

  __device__ Epi2(const ExtractParams &p, unsigned *smem_extra) : P(p), mask(smem_extra + threadIdx.x) {
    tau = p.tau_rel * (__uint_as_float(p.stats[0]) + __uint_as_float(p.stats[1]));
  }

  __device__ __forceinline__ void tile_begin() {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      m[q] = m2[q] = __int_as_float(0x7f800000);
      wb[q] = 0;
      if (!WTA) {
        S[q] = 0.0f;
        vfrom[q] = 0;
        rowbits[q] = 0u;
      }
    }
    if (!WTA)
      for (int w = 0; w < P.nwords * NQ; ++w) mask[w * kCThreads] = 0u;
  }

  // one output row of the thread (ROW compile-time), one skewed block: acc2[pp][r].x is pixel 2pp at
  // dx = 8*blk + r, .y pixel 2pp+1 at dx = 8*blk - 1 + r
  // MAYBE_FIRST: blk may be 0 (the dx = -1 slot of the odd pixels is masked)
  template <int ROW, int R, bool MAYBE_FIRST>
  __device__ __forceinline__ void block(float2 (&acc2)[2][R], int dy, int blk) {
    const float inf = __int_as_float(0x7f800000);
    const int dx0 = blk * kR;
    if (P.dbg & 1) {  // tuning: no epilogue (results are wrong), keeps the sums alive
      m[ROW * kP] = fminf(m[ROW * kP], acc2[0][0].x + acc2[1][R - 1].y);
      return;
    }
    if (MAYBE_FIRST && blk == 0) {  // dx = -1 of the odd pixels
      acc2[0][0].y = inf;
      acc2[1][0].y = inf;
    }
    if ((R != kR || MAYBE_FIRST) && dx0 + R > P.g.maxw) {  // past the window (last block only)
#pragma unroll
      for (int pp = 0; pp < 2; ++pp)
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (dx0 + r >= P.g.maxw) acc2[pp][r].x = inf;
          if (dx0 - 1 + r >= P.g.maxw) acc2[pp][r].y = inf;
        }
    }
    const int bid = dy * P.g.bs.per_row() + blk;
    float mold[kP];
#pragma unroll
    for (int pp = 0; pp < 2; ++pp) {
      float bx = acc2[pp][0].x, by = acc2[pp][0].y;
#pragma unroll
      for (int r = 1; r < R; ++r) {
        bx = fminf(bx, acc2[pp][r].x);
        by = fminf(by, acc2[pp][r].y);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int q = ROW * kP + 2 * pp + h;
        const float b = h ? by : bx;
        m2[q] = fminf(m2[q], fmaxf(m[q], b));
        wb[q] = b < m[q] ? bid : wb[q];
        mold[2 * pp + h] = m[q];
        m[q] = fminf(m[q], b);
      }
    }
    if (WTA) return;
    // (the host only picks this kernel when one shortlist bit is one block: gb == 1)
#pragma unroll
    for (int pp = 0; pp < 2; ++pp) {
      const int q0 = ROW * kP + 2 * pp;
      const float2 mL2 = make_float2(m[q0] * kLog2e, m[q0 + 1] * kLog2e);
      // rescale of what was summed relative to the old minimum (old = +inf: 2^-inf = 0)
      const float sc0 = ex2_approx((m[q0] - mold[2 * pp]) * kLog2e);
      const float sc1 = ex2_approx((m[q0 + 1] - mold[2 * pp + 1]) * kLog2e);
      float2 eb2 = make_float2(0.0f, 0.0f);
      float ex = 0.0f, ey = 0.0f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float2 t = __ffma2_rn(acc2[pp][r], make_float2(-kLog2e, -kLog2e), mL2);
        const float2 e = make_float2(ex2_approx(t.x), ex2_approx(t.y));
        eb2 = __fadd2_rn(eb2, e);
        ex = fmaxf(ex, e.x);
        ey = fmaxf(ey, e.y);
      }
      S[q0] = fmaf(S[q0], sc0, eb2.x);
      S[q0 + 1] = fmaf(S[q0 + 1], sc1, eb2.y);
      if (P.nwords) {
        // everything before a new minimum that rescaled the old one below the threshold is dead
        vfrom[q0] = sc0 < P.p_clear ? bid : vfrom[q0];
        vfrom[q0 + 1] = sc1 < P.p_clear ? bid : vfrom[q0 + 1];
        // p_k(final) <= e_k / S(now): an entry of this block can end above the threshold only if
        // the block's largest term exceeds thr * S.  Collected per window row, flushed by row_end.
        rowbits[q0] |= ex > P.thr_lo * S[q0] ? 1u << blk : 0u;
        rowbits[q0 + 1] |= ey > P.thr_lo * S[q0 + 1] ? 1u << blk : 0u;
      }
    }
  }

  // after the last block of window row dy of output row ROW: move the row's shortlist bits into
  // the per-pixel bitmap in shared memory (the only data-dependent branch of the sweep)
  template <int ROW>
  __device__ __forceinline__ void row_end(int dy) {
    if (WTA) return;
    unsigned any = 0u;
#pragma unroll
    for (int p = 0; p < kP; ++p) any |= rowbits[ROW * kP + p];
    if (any) {
      const int bit0 = dy * P.g.bs.per_row();
      const int w0 = bit0 >> 5, sh = bit0 & 31;
#pragma unroll
      for (int p = 0; p < kP; ++p) {
        const int q = ROW * kP + p;
        const unsigned long long v = (unsigned long long)rowbits[q] << sh;
        if ((unsigned)v) mask[(w0 * NQ + q) * kCThreads] |= (unsigned)v;
        if ((unsigned)(v >> 32)) mask[((w0 + 1) * NQ + q) * kCThreads] |= (unsigned)(v >> 32);
        rowbits[q] = 0u;
      }
    }
  }
};

// Tile end of the two-row sweep, one pixel: resolve the winning block, decide, store.  Deliberately
// NOT inlined: the call sites are unrolled over the thread's 8 pixels (their state lives in
// registers under compile-time indices), the body exists once -- inlined eight times it made the
// kernel 230 KB of code and the instruction cache the bottleneck.  All loads of the block (8
// entries x (C + 1) rows) are issued before the first use: one L2 round trip per pixel.
struct Resolve2 {
  float m, m2, S, tau;
  int wb, vfrom, n, y, x, q;
};

template <int CT, bool WTA, int NQ, int CTHREADS>
__device__ __noinline__ unsigned sweep2_resolve(const ExtractParams &P, const Resolve2 R, const float (&av)[CT],
                                                const unsigned *mask) {
  const SweepGeom &g = P.g;
  const float inf = __int_as_float(0x7f800000);
  const int per_row = g.bs.per_row();
  const int y = R.y, x = R.x, n = R.n;
  const unsigned o = ((unsigned)n * g.H1 + y) * g.W1 + x;   // < 2^31 (checked on the host)
  // ---- the winning block again: same operation sequence as dot2_block, entry by entry
  const int dy = R.wb / per_row, blk = R.wb - dy * per_row;
  const int dxb = blk * kR - (x & 1);
  const int width = blk >= g.bs.n8 ? g.bs.tail_r : kR;
  // entries outside the window are loaded from a clamped column (always inside the frame row) and
  // masked afterwards, so that no load depends on a predicate
  const int rowoff2 = (int)((long long)n * P.s2n + (long long)(y + dy) * P.s2y) + x;   // fits: frames < 2^31 floats
  const int rowoffn = (int)((long long)n * P.nb_sn + (long long)(y + dy) * P.nb_sy) + x;
  int col[kR];
  float bv[CT][kR], nv[kR];
#pragma unroll
  for (int r = 0; r < kR; ++r) {
    const int dx = dxb + r;
    col[r] = dx < 0 ? 0 : (dx >= g.maxw ? g.maxw - 1 : dx);
  }
#pragma unroll
  for (int k = 0; k < CT; ++k) {
    const float *bp = P.in2 + (long long)(k < g.Cin ? k : 0) * P.s2c + rowoff2;
#pragma unroll
    for (int r = 0; r < kR; ++r) bv[k][r] = __ldg(bp + col[r]);
  }
#pragma unroll
  for (int r = 0; r < kR; ++r) nv[r] = __ldg(P.nb + rowoffn + col[r]);
  float v[kR];
#pragma unroll
  for (int r = 0; r < kR; ++r) {
    float acc = __fmul_rn(av[0], bv[0][r]);
#pragma unroll
    for (int k = 1; k < CT; ++k) acc = fmaf(av[k], k < g.Cin ? bv[k][r] : 0.0f, acc);
    acc = fmaf(1.0f, nv[r], acc);
    const int dx = dxb + r;
    v[r] = (r < width && dx >= 0 && dx < g.maxw) ? acc : inf;
  }
  int rb = -1;
#pragma unroll
  for (int r = kR - 1; r >= 0; --r)
    if (v[r] == R.m) rb = r;
  float second = R.m2;
#pragma unroll
  for (int r = 0; r < kR; ++r)
    if (r != rb) second = fminf(second, v[r]);
  // the dot form cannot order entries closer than tau: entry-by-entry rescore decides them (ties
  // included, so the zero-flow rule never has to be evaluated here).  rb < 0 cannot happen (the
  // block was computed with the same operations); it would be caught here too.
  if (second - R.tau <= R.m || rb < 0) {
    P.resc[atomicAdd(P.nresc, 1u)] = (int)o;
    return 0u;
  }
  const int dxw = dxb + rb;
  const int win = dy * g.maxw + dxw + 1;
  // ---- the winner in the difference form: min_ssd, and the reference point of the soft-max
  float ssd = 0.0f, na = 0.0f;
#pragma unroll
  for (int k = 0; k < CT; ++k) {
    const float a = -0.5f * av[k];
    na = fmaf(a, a, na);
    float bw = bv[k][0];
#pragma unroll
    for (int r = 1; r < kR; ++r) bw = r == rb ? bv[k][r] : bw;
    const float d = a - (k < g.Cin ? bw : 0.0f);
    ssd = fmaf(d, d, ssd);
  }
  if (P.index) P.index[o] = win;
  if (P.min_ssd) P.min_ssd[o] = ssd;
  if (P.flow_full) {
    const size_t plane = (size_t)P.h_img * P.w_img;
    const size_t fo = (size_t)n * 2 * plane + (size_t)(y + P.hoff) * P.w_img + (x + P.woff);
    P.flow_full[fo] = (float)(dy + 1 - P.cy);
    P.flow_full[fo + plane] = (float)(dxw + 1 - P.cx);
  }
  if (WTA) return 0u;
  // S was summed relative to the dot-form minimum m (+ |a|^2 = the dot-form SSD); move it to the
  // difference-form one: every term but the winner's own (exactly 1) scales by c
  const float c = expf(ssd - (R.m + na));
  const float Sx = fmaf(R.S - 1.0f, c, 1.0f);
  const float inv = 1.0f / Sx;
  if (P.pmax) P.pmax[o] = inv;
  unsigned untouched = 0u;
  if (P.todo) {
    // extractOutput(prob, thr) (extract_output.cpp:63-155), decided here when it is safe by a
    // margin: the runner-up's term is e2 = exp(m - second) (`second` is the true second smallest
    // entry).  pmax < thr: untouched.  pmax > thr and e2/S < thr: the list is {pmax}.  Otherwise
    // the exact per-pixel pass re-scores the shortlisted blocks.
    const float e2 = expf(R.m - second) * c;
    long long ret = 0;
    float score = 0.0f;
    if (inv > P.p_gt && e2 * inv < P.p_none) {
      ret = win;
      score = (float)((double)P.M * (double)inv);
    } else if (inv < P.p_none) {
      untouched = 1u;
    } else {
      const unsigned slot = atomicAdd(P.ntodo, 1u);
      P.todo[slot] = (int)o;
      for (int w = 0; w < P.nwords; ++w) {
        unsigned bits = mask[(w * NQ + R.q) * CTHREADS];
        const int lo = R.vfrom - 32 * w;  // bits below vfrom are dead
        if (lo >= 32) bits = 0u;
        else if (lo > 0) bits &= ~0u << lo;
        P.todo_mask[(size_t)slot * P.nwords + w] = bits;
      }
      P.vmin[o] = ssd;
      P.vinv[o] = inv;
    }
    if (P.index_thr) P.index_thr[o] = ret;
    if (P.score_thr) P.score_thr[o] = score;
  }
  return untouched;
}

template <class Cfg, int CT, int EPI>
__device__ __forceinline__ void sweep2_tile_end(Epi2<Cfg, CT, EPI> &E, const float2 (&a2)[2][CT][2], int n, int yrow0,
                                                int x0) {
  constexpr bool WTA = EPI == kEpiWta;
  constexpr int NQ = 2 * kP;
  const ExtractParams &P = E.P;
  unsigned untouched = 0;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int row = q / kP, p = q % kP;
    Resolve2 R;
    R.y = yrow0 + row;
    R.x = x0 + p;
    if (R.y >= P.g.H1 || R.x >= P.g.W1) continue;
    R.n = n;
    R.q = q;
    R.m = E.m[q];
    R.m2 = E.m2[q];
    R.wb = E.wb[q];
    R.tau = E.tau;
    R.S = WTA ? 0.0f : E.S[WTA ? 0 : q];
    R.vfrom = WTA ? 0 : E.vfrom[WTA ? 0 : q];
    float av[CT];
#pragma unroll
    for (int k = 0; k < CT; ++k) {
      const float2 t = a2[row][k][p >> 1];
      av[k] = (p & 1) ? t.y : t.x;  // -2a
    }
    untouched += sweep2_resolve<CT, WTA, NQ, Cfg::kCThreads>(P, R, av, E.mask);
  }
  if (P.n_untouched && untouched) atomicAdd(P.n_untouched + n, (unsigned long long)untouched);
}

// Refill of ring slot `slot` with row `seq` of this CTA's row sequence (tile seq / rows_total, slab
// row seq % rows_total), issued by one lane.
__device__ __forceinline__ void sweep2_issue(const CUtensorMap *tmap, const CUtensorMap *tmap_nb, const SweepGeom &g,
                                             float *ring, uint64_t *full, uint32_t seq, int slot, int rows_total, int th,
                                             uint32_t slab_bytes) {
  const int tl = (int)(seq / (uint32_t)rows_total), j = (int)(seq - (uint32_t)tl * rows_total);
  const int tile = blockIdx.x + tl * gridDim.x;
  const int tx = tile % g.tiles_x;
  const int ty = (tile / g.tiles_x) % g.tiles_y;
  const int n = tile / (g.tiles_x * g.tiles_y);
  const int y = ty * th + j, xt = tx * kTW;
  if (y < g.H2) {
    mbar_arrive_expect_tx(&full[slot], slab_bytes);
    tma_load_4d(ring + slot * g.slab_floats, tmap, &full[slot], xt, y, 0, n);
    tma_load_4d(ring + slot * g.slab_floats + g.nb_off, tmap_nb, &full[slot], xt, y, 0, n);
  } else {
    mbar_arrive(&full[slot]);  // row below the frame: only masked pixels read it
  }
}

// N8: number of full 8-wide blocks per window row when known at compile time (4 = a 33..34 wide
// window with a 2-wide tail) so that the row body is straight-line code; 0 = generic (rolled).
template <class Cfg, int CT, int EPI, int N8>
__global__ void __launch_bounds__(Cfg::kThreads, 1)
match_sweep2_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_nb,
                    const ExtractParams P) {
  {  // twin launch: this kernel is the dot form
    const bool dot_ok = __uint_as_float(P.stats[0]) + __uint_as_float(P.stats[1]) <= P.dot_limit;
    if (!dot_ok) return;
  }
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const SweepGeom &g = P.g;
  float *ring = reinterpret_cast<float *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)g.nslot * g.slab_floats);
  uint64_t *empty = full + Cfg::kNSlot;
  unsigned *issued = reinterpret_cast<unsigned *>(empty + Cfg::kNSlot);  // refills issued per slot
  unsigned *extra = reinterpret_cast<unsigned *>(reinterpret_cast<unsigned char *>(full) + kBar2Bytes);

  constexpr int kWarps = Cfg::kWarps, kTH = Cfg::kTH;
  const int kNSlot = g.nslot;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int rows_total = kTH + g.maxh - 1;
  const uint32_t slab_bytes = (uint32_t)((g.C + 1) * g.WB * sizeof(float));
  const int slab_floats = g.slab_floats;
  const int n8 = N8 ? N8 : g.bs.n8;
  const bool wide_tail = N8 ? false : g.bs.tail_r == kR;
  const int my_tiles = (g.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t seq_total = (uint32_t)my_tiles * (uint32_t)rows_total;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap);
    prefetch_tmap(&tmap_nb);
    for (int s = 0; s < kNSlot; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kWarps);
      issued[s] = 0u;
    }
    fence_mbar_init();
    for (uint32_t s = 0; s < (uint32_t)kNSlot && s < seq_total; ++s)
      sweep2_issue(&tmap, &tmap_nb, g, ring, full, s, (int)s, rows_total, kTH, slab_bytes);
  }
  __syncthreads();

  Epi2<Cfg, CT, EPI> epi(P, extra);
  uint32_t g0 = 0;
  for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
    const int tx = tile % g.tiles_x;
    const int ty = (tile / g.tiles_x) % g.tiles_y;
    const int n = tile / (g.tiles_x * g.tiles_y);
    const int y0 = ty * kTH, xt = tx * kTW;
    const int yr = y0 + 2 * warp, x0 = xt + lane * kP;

    float2 a2[2][CT][2];
#pragma unroll
    for (int row = 0; row < 2; ++row) {
      const bool rowok = yr + row < g.H1;
      const float *src = g.in1 + (long long)n * g.s1n + (long long)(rowok ? yr + row : 0) * g.s1y;
#pragma unroll
      for (int k = 0; k < CT; ++k) {
        float a[kP];
#pragma unroll
        for (int p = 0; p < kP; ++p)
          a[p] = (rowok && k < g.Cin && x0 + p < g.W1) ? -2.0f * __ldg(src + (long long)k * g.s1c + x0 + p) : 0.0f;
        a2[row][k][0] = make_float2(a[0], a[1]);
        a2[row][k][1] = make_float2(a[2], a[3]);
      }
    }
    epi.tile_begin();

#pragma unroll 1
    for (int j = 0; j < rows_total; ++j) {
      const uint32_t seq = g0 + (uint32_t)j;
      const int slot = (int)(seq % kNSlot);
      const uint32_t inst = seq / kNSlot;
      if (!(P.dbg & 2)) mbar_wait(&full[slot], inst & 1u);   // dbg 2 (tuning): no ring protocol at all
      const int d0 = j - 2 * warp;  // window row of the first output row; the second is at d0 - 1
      const float *brow = ring + slot * slab_floats + lane * kP;
      if (d0 >= 1 && d0 < g.maxh) {
        // ---- both output rows inside their windows: straight-line when N8 is known
        {
          float2 acc2[2][2][kR];
          dot2_block<CT, kR>(a2, brow, brow + g.nb_off, g.WB, acc2);
          epi.template block<0, kR, true>(acc2[0], d0, 0);
          epi.template block<1, kR, true>(acc2[1], d0 - 1, 0);
        }
        if (N8) {
#pragma unroll
          for (int blk = 1; blk < (N8 ? N8 : 1); ++blk) {
            float2 acc2[2][2][kR];
            dot2_block<CT, kR>(a2, brow + blk * kR, brow + g.nb_off + blk * kR, g.WB, acc2);
            epi.template block<0, kR, false>(acc2[0], d0, blk);
            epi.template block<1, kR, false>(acc2[1], d0 - 1, blk);
          }
        } else {
          const int nwide = n8 + (wide_tail ? 1 : 0);
#pragma unroll 1
          for (int blk = 1; blk < nwide; ++blk) {
            float2 acc2[2][2][kR];
            dot2_block<CT, kR>(a2, brow + blk * kR, brow + g.nb_off + blk * kR, g.WB, acc2);
            epi.template block<0, kR, false>(acc2[0], d0, blk);
            epi.template block<1, kR, false>(acc2[1], d0 - 1, blk);
          }
        }
        if (!wide_tail) {
          float2 acc2[2][2][2];
          dot2_block<CT, 2>(a2, brow + n8 * kR, brow + g.nb_off + n8 * kR, g.WB, acc2);
          epi.template block<0, 2, false>(acc2[0], d0, n8);
          epi.template block<1, 2, false>(acc2[1], d0 - 1, n8);
        }
        epi.template row_end<0>(d0);
        epi.template row_end<1>(d0 - 1);
      } else if (d0 == 0 || d0 == g.maxh) {
        // ---- first / last slab row of this warp's window span: one output row only (rolled code)
        const int nwide = n8 + (wide_tail ? 1 : 0);
#pragma unroll 1
        for (int blk = 0; blk < nwide; ++blk) {
          float2 acc2[2][2][kR];
          dot2_block<CT, kR>(a2, brow + blk * kR, brow + g.nb_off + blk * kR, g.WB, acc2);
          if (d0 == 0) epi.template block<0, kR, true>(acc2[0], d0, blk);
          else epi.template block<1, kR, true>(acc2[1], d0 - 1, blk);
        }
        if (!wide_tail) {
          float2 acc2[2][2][2];
          dot2_block<CT, 2>(a2, brow + n8 * kR, brow + g.nb_off + n8 * kR, g.WB, acc2);
          if (d0 == 0) epi.template block<0, 2, false>(acc2[0], d0, n8);
          else epi.template block<1, 2, false>(acc2[1], d0 - 1, n8);
        }
        if (d0 == 0) epi.template row_end<0>(d0);
        else epi.template row_end<1>(d0 - 1);
      }
      __syncwarp();
      if (lane == 0 && !(P.dbg & 2)) {
        // done with the row; whoever completes the slot's `empty` phase loads its next tenant
        mbar_arrive(&empty[slot]);
        const uint32_t next = seq + (uint32_t)kNSlot;
        if (next < seq_total && mbar_test(&empty[slot], inst & 1u) && atomicCAS(&issued[slot], inst, inst + 1u) == inst)
          sweep2_issue(&tmap, &tmap_nb, g, ring, full, next, slot, rows_total, kTH, slab_bytes);
      }
    }
    if (!(P.dbg & 4)) sweep2_tile_end<Cfg, CT, EPI>(epi, a2, n, yr, x0);   // dbg 4 (tuning): no resolve / stores
    else if (epi.m[0] == 123.456f) P.resc[0] = 1;
    g0 += (uint32_t)rows_total;
  }
}

}  // namespace dm
