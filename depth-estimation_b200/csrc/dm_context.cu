// dm_context.cu -- context, error reporting, pointer classification and staging.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <utility>

#include "dm_common.cuh"

namespace dm {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return DM_ERR_CUDA;
}

PtrKind classify(const void *p) {
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return PtrKind::Host;
  }
  if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) return PtrKind::Device;
  return PtrKind::Host;
}

int ensure_arena(dm_ctx *ctx, size_t bytes) {
  if (ctx->arena.cap >= bytes) return DM_OK;
  // growing invalidates earlier pointers of this call: callers size before use
  DM_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->arena.base) DM_CUDA(cudaFree(ctx->arena.base));
  ctx->arena.base = nullptr;
  ctx->arena.cap = 0;
  size_t want = bytes + (bytes >> 2) + (1u << 20);
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("out of device memory allocating %zu bytes of workspace", want);
    return DM_ERR_NOMEM;
  }
  ctx->arena.base = static_cast<char *>(p);
  ctx->arena.cap = want;
  return DM_OK;
}

Call::Call(dm_ctx *c, bool defer_sync) : ctx(c), defer(defer_sync) {
  // size the arena for what the previous call needed, so that steady-state calls never fall
  // back to per-call cudaMalloc blocks
  if (ctx->arena.demand > ctx->arena.cap) ensure_arena(ctx, ctx->arena.demand);
  ctx->arena.demand = 0;
  ctx->arena.used = 0;
  ctx->pending.clear();
  ctx->call_has_host = false;
}

// The arena is a bump allocator.  To keep pointers stable we never grow it in the
// middle of a call once something was handed out: instead a chain of extra
// cudaMalloc'd blocks is avoided by letting the first allocation of a call that
// does not fit trigger a grow only when nothing is live; otherwise we fall back to
// a dedicated allocation that is released at finish().
struct Extra {
  void *p;
};
static thread_local std::vector<void *> g_extra;

int Call::alloc(void **dptr, size_t bytes) {
  const size_t aligned = (bytes + 255) & ~size_t(255);
  ctx->arena.demand += aligned;
  if (ctx->arena.used + aligned > ctx->arena.cap) {
    if (ctx->arena.used == 0) {
      DM_CHECK(ensure_arena(ctx, aligned));
    } else {
      void *p = nullptr;
      cudaError_t e = cudaMalloc(&p, aligned);
      if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("out of device memory allocating %zu bytes", aligned);
        return DM_ERR_NOMEM;
      }
      g_extra.push_back(p);
      *dptr = p;
      return DM_OK;
    }
  }
  *dptr = ctx->arena.base + ctx->arena.used;
  ctx->arena.used += aligned;
  return DM_OK;
}

int Call::in(const void *user, size_t bytes, const void **dptr) {
  if (classify(user) == PtrKind::Device) {
    *dptr = user;
    return DM_OK;
  }
  void *d = nullptr;
  DM_CHECK(alloc(&d, bytes));
  DM_CUDA(cudaMemcpyAsync(d, user, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ctx->call_has_host = true;
  *dptr = d;
  return DM_OK;
}

int Call::out(void *user, size_t bytes, void **dptr, bool preload) {
  if (classify(user) == PtrKind::Device) {
    *dptr = user;
    return DM_OK;
  }
  void *d = nullptr;
  DM_CHECK(alloc(&d, bytes));
  if (preload) DM_CUDA(cudaMemcpyAsync(d, user, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ctx->pending.push_back({user, d, bytes});
  ctx->call_has_host = true;
  *dptr = d;
  return DM_OK;
}

int Call::finish() {
  int rc = DM_OK;
  for (auto &p : ctx->pending) {
    cudaError_t e = cudaMemcpyAsync(p.host, p.dev, p.bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e != cudaSuccess && rc == DM_OK) rc = cuda_fail(e, "copy-back", __FILE__, __LINE__);
  }
  ctx->pending.clear();
  if ((ctx->call_has_host && !defer) || !g_extra.empty()) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess && rc == DM_OK) rc = cuda_fail(e, "stream synchronize", __FILE__, __LINE__);
  }
  for (void *p : g_extra) cudaFree(p);
  g_extra.clear();
  if (rc == DM_OK) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(e, "kernel launch", __FILE__, __LINE__);
  }
  return rc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_tensor_map_4d(CUtensorMap *map, const float *base, const uint64_t dims[4],
                         const uint64_t strides_bytes[3], const uint32_t box[4]) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    DM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeTiled not available from the driver");
      return DM_ERR_CUDA;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  cuuint64_t gdims[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gstr[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t gbox[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(base), gdims, gstr,
                  gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)",
              (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)dims[2], (unsigned long long)dims[3], box[0], box[1], box[2],
              box[3]);
    return DM_ERR_CUDA;
  }
  return DM_OK;
}

int tensor_map_4d(dm_ctx *ctx, CUtensorMap *map, const float *base, const uint64_t dims[4],
                  const uint64_t strides_bytes[3], const uint32_t box[4]) {
  for (size_t i = 0; i < ctx->map_plans.size(); ++i) {
    const dm_ctx::MapPlan &p = ctx->map_plans[i];
    if (p.base == base && !memcmp(p.dims, dims, sizeof(p.dims)) && !memcmp(p.strides, strides_bytes, sizeof(p.strides)) &&
        !memcmp(p.box, box, sizeof(p.box))) {
      *map = p.map;
      if (i) std::swap(ctx->map_plans[i], ctx->map_plans[0]);
      return DM_OK;
    }
  }
  DM_CHECK(encode_tensor_map_4d(map, base, dims, strides_bytes, box));
  dm_ctx::MapPlan p;
  p.base = base;
  memcpy(p.dims, dims, sizeof(p.dims));
  memcpy(p.strides, strides_bytes, sizeof(p.strides));
  memcpy(p.box, box, sizeof(p.box));
  p.map = *map;
  ctx->map_plans.insert(ctx->map_plans.begin(), p);
  if (ctx->map_plans.size() > 16) ctx->map_plans.pop_back();
  return DM_OK;
}

// Function attributes belong to the device, not to a dm_ctx: two contexts that shared a kernel but
// kept separate books lowered each other's dynamic shared-memory grant.  One table per device,
// process-wide; a grant is only ever raised.
static std::mutex g_plan_mutex;
static std::vector<dm_ctx::FuncPlan> g_func_plans[64];

static dm_ctx::FuncPlan *func_plan(dm_ctx *ctx, const void *fn) {
  auto &plans = g_func_plans[ctx->device & 63];
  for (auto &p : plans)
    if (p.fn == fn) return &p;
  plans.push_back({fn, 0, 0, 0, 0});
  return &plans.back();
}

int ensure_func_smem(dm_ctx *ctx, const void *fn, size_t smem) {
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  dm_ctx::FuncPlan *p = func_plan(ctx, fn);
  if (smem <= p->smem) return DM_OK;
  DM_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  p->smem = smem;
  return DM_OK;
}

int blocks_per_sm(dm_ctx *ctx, const void *fn, int threads, size_t smem) {
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  dm_ctx::FuncPlan *p = func_plan(ctx, fn);
  if (p->per_sm > 0 && p->threads == threads && p->smem_occ == smem) return p->per_sm;
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem) != cudaSuccess) {
    cudaGetLastError();
    per_sm = 1;
  }
  if (per_sm < 1) per_sm = 1;
  p->per_sm = per_sm;
  p->threads = threads;
  p->smem_occ = smem;
  return per_sm;
}

}  // namespace dm

using namespace dm;

static int apply_option(Options *o, const char *name, const char *value) {
  const int iv = value ? atoi(value) : 0;
  if (!strcmp(name, "ssd_form")) {
    if (!value || !*value || !strcmp(value, "auto")) o->ssd_form = 0;
    else if (!strcmp(value, "diff")) o->ssd_form = 1;
    else if (!strcmp(value, "dot")) o->ssd_form = 2;
    else return DM_ERR_INVALID;
  } else if (!strcmp(name, "no_small_tiles")) o->no_small_tiles = iv != 0;
  else if (!strcmp(name, "no_pipeline")) o->no_pipeline = iv != 0;
  else if (!strcmp(name, "debug_todo")) o->debug_todo = iv != 0;
  else if (!strcmp(name, "pipe_chunk")) o->pipe_chunk = iv;
  else if (!strcmp(name, "volume_debug")) o->volume_debug = iv;
  else if (!strcmp(name, "sweep")) o->sweep = iv;
  else if (!strcmp(name, "conv")) o->conv = iv;
  else if (!strcmp(name, "volume_kernel")) o->volume_kernel = iv;
  else if (!strcmp(name, "conv_tile")) {
    o->conv_tile = o->conv_target = 0;
    if (value) sscanf(value, "%d,%d", &o->conv_tile, &o->conv_target);
  } else return DM_ERR_INVALID;
  return DM_OK;
}

static void options_from_env(Options *o) {
  static const char *const kEnv[][2] = {
      {"DM_SSD_FORM", "ssd_form"},       {"DM_NO_SMALL_TILES", "no_small_tiles"}, {"DM_NO_PIPELINE", "no_pipeline"},
      {"DM_DEBUG_TODO", "debug_todo"},   {"DM_PIPE_CHUNK", "pipe_chunk"},         {"DM_VOLUME_DEBUG", "volume_debug"},
      {"DM_CONV_TILE", "conv_tile"},     {"DM_SWEEP", "sweep"},
      {"DM_CONV", "conv"},               {"DM_VOLUME_KERNEL", "volume_kernel"}};
  for (const auto &e : kEnv)
    if (const char *v = getenv(e[0])) {
      // flags set to the empty string or anything non-numeric count as "on"
      const bool flag = !strcmp(e[1], "no_small_tiles") || !strcmp(e[1], "no_pipeline") || !strcmp(e[1], "debug_todo");
      apply_option(o, e[1], flag && atoi(v) == 0 && strcmp(v, "0") ? "1" : v);
    }
}

extern "C" {

int dm_version(void) { return DM_VERSION; }

int dm_set_option(dm_ctx *ctx, const char *name, const char *value) {
  DM_REQUIRE(ctx && name, "dm_set_option: NULL argument");
  const int rc = apply_option(&ctx->opt, name, value);
  if (rc != DM_OK) set_error("dm_set_option: unknown option or value '%s' = '%s'", name, value ? value : "");
  for (int i = 0; i < 2; ++i)
    if (ctx->pipe[i]) ctx->pipe[i]->opt = ctx->opt;
  return rc;
}

int dm_last_counts(dm_ctx *ctx, int64_t *rescored, int64_t *exact_pass) {
  DM_REQUIRE(ctx != nullptr, "dm_last_counts: ctx is NULL");
  DM_CUDA(cudaSetDevice(ctx->device));
  DM_CUDA(cudaStreamSynchronize(ctx->stream));
  unsigned v[2] = {0, 0};
  if (ctx->counters_valid) DM_CUDA(cudaMemcpy(v, ctx->counters, sizeof(v), cudaMemcpyDeviceToHost));
  if (rescored) *rescored = ctx->counters_valid ? (int64_t)v[0] : -1;
  if (exact_pass) *exact_pass = ctx->counters_valid ? (int64_t)v[1] : -1;
  return DM_OK;
}

const char *dm_last_error(void) { return dm::g_err; }

int dm_create(int device, dm_ctx **out) {
  DM_REQUIRE(out != nullptr, "dm_create: ctx pointer is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("dm_create: no CUDA device (%s); libdepthmatch has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return DM_ERR_CUDA;
  }
  DM_REQUIRE(device >= 0 && device < ndev, "dm_create: device %d out of range [0,%d)", device, ndev);
  DM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  DM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("dm_create: device %d is sm_%d%d; this library holds sm_100a code only", device,
              prop.major, prop.minor);
    return DM_ERR_UNSUPPORTED;
  }
  dm_ctx *ctx = new dm_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete ctx;
    return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__);
  }
  ctx->stream = ctx->own_stream;
  e = cudaMalloc(&ctx->counters, 4 * sizeof(unsigned));
  if (e != cudaSuccess) {
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
  }
  options_from_env(&ctx->opt);
  *out = ctx;
  return DM_OK;
}

int dm_destroy(dm_ctx *ctx) {
  if (!ctx) return DM_OK;
  cudaSetDevice(ctx->device);
  for (int i = 0; i < 2; ++i)
    if (ctx->pipe[i]) dm_destroy(ctx->pipe[i]);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->arena.base) cudaFree(ctx->arena.base);
  if (ctx->counters) cudaFree(ctx->counters);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  for (int i = 0; i < 2; ++i) {
    if (ctx->aux[i]) cudaStreamDestroy(ctx->aux[i]);
    if (ctx->aux_join[i]) cudaEventDestroy(ctx->aux_join[i]);
  }
  if (ctx->aux_fork) cudaEventDestroy(ctx->aux_fork);
  delete ctx;
  return DM_OK;
}

int dm_synchronize(dm_ctx *ctx) {
  DM_REQUIRE(ctx != nullptr, "dm_synchronize: ctx is NULL");
  DM_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < 2; ++i)  // chunks of a DM_FLAG_ASYNC host-buffer call still in flight
    if (ctx->pipe[i]) DM_CUDA(cudaStreamSynchronize(ctx->pipe[i]->stream));
  return DM_OK;
}

int dm_set_stream(dm_ctx *ctx, void *cuda_stream) {
  DM_REQUIRE(ctx != nullptr, "dm_set_stream: ctx is NULL");
  DM_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stream = static_cast<cudaStream_t>(cuda_stream);  // NULL is CUDA's legacy default stream
  return DM_OK;
}

int dm_reset_stream(dm_ctx *ctx) {
  DM_REQUIRE(ctx != nullptr, "dm_reset_stream: ctx is NULL");
  DM_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stream = ctx->own_stream;
  return DM_OK;
}

void *dm_get_stream(dm_ctx *ctx) { return ctx ? static_cast<void *>(ctx->stream) : nullptr; }

int dm_host_alloc(void **ptr, size_t bytes) {
  DM_REQUIRE(ptr != nullptr, "dm_host_alloc: ptr is NULL");
  DM_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
  return DM_OK;
}

int dm_host_free(void *ptr) {
  if (ptr) DM_CUDA(cudaFreeHost(ptr));
  return DM_OK;
}

int64_t dm_launch_count(dm_ctx *ctx) {
  if (!ctx) return 0;
  int64_t n = ctx->launches;
  for (int i = 0; i < 2; ++i)
    if (ctx->pipe[i]) n += ctx->pipe[i]->launches;
  return n;
}

int dm_set_profiling(dm_ctx *ctx, int on) {
  DM_REQUIRE(ctx != nullptr, "dm_set_profiling: ctx is NULL");
  DM_CUDA(cudaSetDevice(ctx->device));
  if (on && !ctx->ev0) {
    DM_CUDA(cudaEventCreate(&ctx->ev0));
    DM_CUDA(cudaEventCreate(&ctx->ev1));
  }
  ctx->profiling = on != 0;
  return DM_OK;
}

int dm_last_kernel_ms(dm_ctx *ctx, float *ms) {
  DM_REQUIRE(ctx && ms, "dm_last_kernel_ms: NULL argument");
  DM_REQUIRE(ctx->ev0 != nullptr, "dm_last_kernel_ms: profiling was never switched on");
  DM_CUDA(cudaEventSynchronize(ctx->ev1));
  DM_CUDA(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return DM_OK;
}

}  // extern "C"
