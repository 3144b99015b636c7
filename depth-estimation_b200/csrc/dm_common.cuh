// dm_common.cuh -- shared host/device plumbing of libdepthmatch (sm_100a only).
//
// Context object, error reporting, host<->device staging, and the small PTX
// wrappers (mbarrier, TMA bulk-tensor copy) the kernels use.  No torch types, no
// CPU fallback: every entry point fails with DM_ERR_CUDA when no device is there.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/depthmatch.h"

namespace dm {

// ---------------------------------------------------------------- errors
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define DM_CUDA(expr)                                                       \
  do {                                                                      \
    cudaError_t _e = (expr);                                                \
    if (_e != cudaSuccess) return ::dm::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define DM_CHECK(expr)                \
  do {                                \
    int _s = (expr);                  \
    if (_s != DM_OK) return _s;       \
  } while (0)

#define DM_REQUIRE(cond, ...)         \
  do {                                \
    if (!(cond)) {                    \
      ::dm::set_error(__VA_ARGS__);   \
      return DM_ERR_INVALID;          \
    }                                 \
  } while (0)

// ---------------------------------------------------------------- context
struct Arena {
  char *base = nullptr;
  size_t cap = 0, used = 0;
  size_t demand = 0;  // bytes the current call asked for in total (arena + overflow blocks)
};

}  // namespace dm

namespace dm {
// Tuning / diagnostic switches.  Read from the environment ONCE, in dm_create (DM_SSD_FORM,
// DM_NO_SMALL_TILES, DM_NO_PIPELINE, DM_PIPE_CHUNK, DM_VOLUME_DEBUG, DM_DEBUG_TODO, DM_CONV_TILE);
// dm_set_option changes them on a live context.  None of them is needed for normal use.
struct Options {
  int ssd_form = 0;  // 0 auto, 1 difference form always, 2 dot form whatever the norms
  bool no_small_tiles = false, no_pipeline = false, debug_todo = false;
  int pipe_chunk = 0;
  // tuning only (results may be wrong): volume kernels 1 = no global stores; strip kernel 32 = soft-max volume
  // through the statistics sweep instead of the staging buffer; two-row sweep: its own bits (match_sweep2.cuh);
  // 9 = role cycle counts of the tensor-core filter kernel on stderr
  int volume_debug = 0;
  int conv_tile = 0, conv_target = 0;
  int sweep = 0;     // 0 auto; other values select a sweep variant (tuning)
  int volume_kernel = 0;  // 0 auto (strip kernel where it fits and pays), 1 = tiled kernel with sector stores, 2 = strip kernel wherever it fits
  int conv = 0;      // feature extractor: 2 = tensor-core (tcgen05) layers where they fit; else CUDA cores
};
}  // namespace dm

struct dm_ctx {
  int device = 0;
  dm::Options opt;
  // device counters of the most recent dm_match_extract on this context (dm_last_counts)
  // Launch plans: what a call computes once per (kernel, shape) instead of once per call.  Encoded
  // tensor maps live here (keyed by base pointer + geometry; a steady-state caller passes the same
  // buffers); the dynamic shared-memory grants and occupancy answers are per DEVICE (FuncPlan tables
  // in dm_context.cu), because function attributes are.
  struct FuncPlan {
    const void *fn;
    int threads;
    size_t smem;      // largest dynamic size granted so far
    int per_sm;       // occupancy at (threads, smem_occ); 0 = not asked yet
    size_t smem_occ;
  };
  struct MapPlan {
    const void *base;
    uint64_t dims[4], strides[3];
    uint32_t box[4];
    CUtensorMap map;
  };
  std::vector<MapPlan> map_plans;   // small, most recent first
  unsigned *counters = nullptr;  // [0] rescored, [1] exact pass; copied there on the stream after each call
  bool counters_valid = false;
  int num_sms = 0;
  size_t smem_optin = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  dm::Arena arena;       // device scratch + staging, bump-allocated per call
  int64_t launches = 0;  // kernels launched through this context
  bool profiling = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // bracket the sweep kernel when profiling
  // pending device->host copies of the current call
  struct Pending {
    void *host;
    const void *dev;
    size_t bytes;
  };
  std::vector<Pending> pending;
  bool call_has_host = false;
  // two private sub-contexts (own stream + arena) used to pipeline host-buffer batches:
  // H2D of chunk i+1 overlaps the kernels of chunk i and the D2H of chunk i-1
  dm_ctx *pipe[2] = {nullptr, nullptr};
  bool is_child = false;
  // side streams for independent sub-problems of one call (the scales of the multiscale model):
  // forked from and joined to `stream` with events, created on first use
  cudaStream_t aux[2] = {nullptr, nullptr};
  cudaEvent_t aux_fork = nullptr, aux_join[2] = {nullptr, nullptr};
};

namespace dm {

enum class PtrKind { Device, Host };
PtrKind classify(const void *p);

// Per-call staging helper.  begin() resets the arena; in()/out() return a device
// pointer for a user pointer (copying host data in, registering host outputs for
// copy-back); finish() copies outputs back and synchronises when any host pointer
// was involved.
struct Call {
  dm_ctx *ctx;
  bool defer = false;  // leave the final synchronise to the caller (pipelined chunks)
  explicit Call(dm_ctx *c, bool defer_sync = false);
  int alloc(void **dptr, size_t bytes);  // device scratch, 256-byte aligned
  int in(const void *user, size_t bytes, const void **dptr);
  // like in() but a pitched 2D copy: rows of row_bytes, user pitch -> dense device rows
  int out(void *user, size_t bytes, void **dptr, bool preload = false);
  int finish();
};

int ensure_arena(dm_ctx *ctx, size_t bytes);

inline void count_launch(dm_ctx *ctx, int n = 1) { ctx->launches += n; }
inline void prof_begin(dm_ctx *ctx) {
  if (ctx->profiling) cudaEventRecord(ctx->ev0, ctx->stream);
}
inline void prof_end(dm_ctx *ctx) {
  if (ctx->profiling) cudaEventRecord(ctx->ev1, ctx->stream);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda, so the
// library still loads on a box without a driver).
int encode_tensor_map_4d(CUtensorMap *map, const float *base, const uint64_t dims[4],
                         const uint64_t strides_bytes[3], const uint32_t box[4]);
// the same through the context's plan cache (dm_ctx::map_plans)
int tensor_map_4d(dm_ctx *ctx, CUtensorMap *map, const float *base, const uint64_t dims[4],
                  const uint64_t strides_bytes[3], const uint32_t box[4]);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) only when the kernel needs more than it was granted
int ensure_func_smem(dm_ctx *ctx, const void *fn, size_t smem);
// resident CTAs per SM of (fn, threads, smem), asked from the runtime once
int blocks_per_sm(dm_ctx *ctx, const void *fn, int threads, size_t smem);

constexpr float kLog2e = 1.4426950408889634f;

// Parameters of the untiled one-warp-per-pixel kernels (match_generic.cu); the tiled sweep
// fills one too for the pixels it hands over to generic_rescore.
struct GenericParams {
  const float *in1, *in2;
  long long s1n, s1c, s1y, s2n, s2c, s2y;
  int N, C, H1, W1, maxh, maxw;
  unsigned flags;
  double thr;
  int M, middle, cy, cx, h_img, w_img, hoff, woff;
  long long *index;
  float *min_ssd, *pmax, *flow_full;
  long long *index_thr;
  float *score_thr, *soft_yx;
  unsigned long long *n_untouched;
  float *radial_flow;  // argmin - 1 as float

  float *conf_marginal;  // getOutputConfidences2's confidence, see dm_extract_out
  // volume
  int mode;
  float *vol;
  // list mode (generic_rescore): only the pixels list[0 .. *nlist) are computed
  const int *list;
  const unsigned *nlist;
};
int generic_rescore(dm_ctx *ctx, GenericParams P, const int *list, const unsigned *nlist);

#ifdef __CUDACC__
// ---------------------------------------------------------------- device PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// same, for a thread that expects to wait long (the TMA producer): let the hardware suspend
// the thread until the phase completes (try_wait's suspend-time hint) instead of polling, so
// the wait costs the producer's scheduler almost no issue slots
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAITB_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONEB_%=;\n\t"
      "bra WAITB_%=;\n\t"
      "DONEB_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(1000u)
      : "memory");
}
// 4-D tiled TMA load global -> shared, completion on an mbarrier
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *map, uint64_t *bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#endif

}  // namespace dm
