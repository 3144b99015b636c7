// tcgen05.mma kind::tf32 issue/execution rate on sm_100a: REP back-to-back MMAs (M = 128, K = 8) into one TMEM
// accumulator, A from TMEM (TS) or from shared memory (SS), B K-major no-swizzle from shared memory, for several N.
// One CTA, nothing else running.  Prints cycles per MMA (issue -> commit completion).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
constexpr int M = 128, K = 8, REP = 256;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3fff) | (uint64_t)((lbo >> 4) & 0x3fff) << 16 | (uint64_t)((sbo >> 4) & 0x3fff) << 32 | (uint64_t)1 << 46;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__global__ void rate(int N, int ts, int nsteps, int uniform, int nacc, long long *out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float *sB = reinterpret_cast<float *>(smem);          // N x (8 * nsteps) K-major
  float *sA = sB + 256 * 8 * 16;
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // the shuffle tells the compiler the value is warp-uniform
  for (int i = tid; i < 256 * 8 * 16 + 128 * 8 * 16; i += blockDim.x) sB[i] = 0.001f * (i % 97);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tptr, tmem0 = tptr;
  if (uniform && warp == 0) {
    // the whole warp walks the loop (operands are warp-uniform arithmetic), one elected lane issues
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t sbo = 32u * 8u * (uint32_t)nsteps;
    const uint64_t dbb = make_desc(smem_u32(sB), 128, sbo), dab = make_desc(smem_u32(sA), 128, sbo);
    const long long t0 = clock64();
    for (int r = 0; r < REP; r += 16) {
#pragma unroll
      for (int s = 0; s < 16; ++s) {
        const uint64_t db = dbb + (uint64_t)(s * 16), da = dab + (uint64_t)(s * 16);
        const uint32_t acc = (r | s) != 0;
        const uint32_t tmem = tmem0 + (uint32_t)((s % 4 < nacc ? s % 4 : 0) * N);
        // no C++ branch: every lane executes the block with warp-uniform operands, the MMA itself is predicated on elect.sync
        if (ts)
          asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem),
                       "r"(tmem0 + 384 + s * 8), "l"(db), "r"(idesc), "r"(acc)
                       : "memory");
        else
          asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                       "l"(da), "l"(db), "r"(idesc), "r"(acc)
                       : "memory");
      }
    }
    const long long t1 = clock64();
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    if (tid == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  if (!uniform && tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t sbo = 32u * 8u * (uint32_t)nsteps;
    const uint64_t dbb = make_desc(smem_u32(sB), 128, sbo), dab = make_desc(smem_u32(sA), 128, sbo);
    const long long t0 = clock64();
    for (int r = 0; r < REP; ++r) {
      const int s = r % nsteps;
      const uint64_t db = dbb + (uint64_t)(s * 16), da = dab + (uint64_t)(s * 16);
      const uint32_t acc = r > 0;
      if (ts)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem),
                     "r"(tmem + 256 + s * 8), "l"(db), "r"(idesc), "r"(acc)
                     : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                     "l"(da), "l"(db), "r"(idesc), "r"(acc)
                     : "memory");
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
int main() {
  long long *d, h[2];
  cudaMalloc(&d, 16);
  const size_t smem = (256 * 8 * 16 + 128 * 8 * 16) * 4;
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int uniform = 1; uniform >= 0; --uniform)
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {16, 80, 128, 256})
      for (int nacc : {1, 2, 4}) {
        const int nsteps = 16;
        if (!uniform && (N != 80 || nacc != 1)) continue;
        if (nacc * N > 384) continue;
        rate<<<1, 128, smem>>>(N, ts, nsteps, uniform, nacc, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%s %s N=%3d accumulators %d: issue %.1f cycles/MMA, issue+drain %.1f cycles/MMA  (floor 128*N/256 = %d)  %s\n", uniform ? "warp-uniform issue" : "one-thread issue  ", ts ? "TS" : "SS", N, nacc,
               (double)h[0] / REP, (double)h[1] / REP, 128 * N / 256, cudaGetErrorString(e));
      }
  return 0;
}
