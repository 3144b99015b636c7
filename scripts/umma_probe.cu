// Probe of tcgen05.mma kind::tf32 with no-swizzle (INTERLEAVE) shared-memory descriptors on sm_100a:
// D[128 x N] = A[128 x K] * B[N x K]^T with A MN-major (M contiguous, as the feature extractor
// stages image rows) and B K-major, K = 8 * KSTEPS.  Prints the largest deviation from a CPU product
// for the descriptor convention given on the command line.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu && ./umma_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int M = 128, N = 32, KSTEPS = 3, K = 8 * KSTEPS;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;                // layout_type = 0 (no swizzle), base_offset 0
}

// A (MN-major): element (m, k) at  (m%4) + 4*(k%8) + a_sbo_e*(m/4) + a_lbo_e*(k/8)      [floats]
// B (K-major) : element (n, k) at  (k%4) + 4*(n%8) + b_sbo_e*(n/8) + b_lbo_e*(k/4)      [floats]
__global__ void probe(const float *A, const float *B, float *D, int a_lbo, int a_sbo, int b_lbo, int b_sbo, int swap_a,
                      int swap_b, int mode) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float *sA = reinterpret_cast<float *>(smem);
  float *sB = sA + M * K;
  uint64_t *bar = reinterpret_cast<uint64_t *>(sB + 2 * N * K);
  uint32_t *tptr = reinterpret_cast<uint32_t *>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < M * K; i += blockDim.x) {
    const int m = i / K, k = i % K;
    if (mode == 1)  // K-major A, same form as B: (k%4) + 4*(m%8) + sbo*(m/8) + lbo*(k/4)
      sA[(k % 4) + 4 * (m % 8) + (a_sbo / 4) * (m / 8) + (a_lbo / 4) * (k / 4)] = A[m * K + k];
    else
      sA[(m % 4) + 4 * (k % 8) + (a_sbo / 4) * (m / 4) + (a_lbo / 4) * (k / 8)] = A[m * K + k];
  }
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    sB[(k % 4) + 4 * (n % 8) + (b_sbo / 4) * (n / 8) + (b_lbo / 4) * (k / 4)] = B[n * K + k];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tptr)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // generic-proxy writes of the operands must be visible to the async proxy (the tensor core)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tptr;
  if (mode == 2) {  // TMEM store / load round trip only
    const uint32_t taddr0 = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < N; ++c) {
      const uint32_t val = __float_as_uint((float)(tid * 100 + c));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr0 + c), "r"(val) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  if (mode >= 3) {
    // A straight from registers into TMEM: thread m = lane m writes its K values into columns 64.. (hi) and 128.. (lo)
    const uint32_t taddr0 = tmem + ((uint32_t)(warp * 32) << 16);
    for (int k = 0; k < K; ++k) {
      const float a = A[tid * K + k];
      const float hi = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr0 + 64 + k), "r"(__float_as_uint(a)) : "memory");
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr0 + 128 + k), "r"(__float_as_uint(a - hi)) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    // B lo right behind B hi in shared memory
    for (int i = tid; i < N * K; i += blockDim.x) {
      const int n = i / K, k = i % K;
      const float b = B[n * K + k];
      const float hi = __uint_as_float(__float_as_uint(b) & 0xffffe000u);
      sB[N * K + (k % 4) + 4 * (n % 8) + (b_sbo / 4) * (n / 8) + (b_lbo / 4) * (k / 4)] = b - hi;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) |
                             ((uint32_t)(M >> 4) << 24);
      int first = 1;
      for (int term = 0; term < (mode == 4 ? 3 : 1); ++term)
        for (int s = 0; s < KSTEPS; ++s) {
          const uint32_t ta = tmem + (term == 2 ? 128 : 64) + s * 8;
          const uint64_t db = make_desc(smem_u32(sB + (term == 1 ? N * K : 0)) + s * 2 * b_lbo, b_lbo, b_sbo);
          const uint32_t acc = first ? 0u : 1u;
          first = 0;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem),
              "r"(ta), "l"(db), "r"(idesc), "r"(acc)
              : "memory");
        }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
  }
  if (tid == 0 && mode < 2) {
    // instruction descriptor: c=f32, a=b=tf32, A MN-major, B K-major, N>>3, M>>4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((mode == 1 ? 0u : 1u) << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(M >> 4) << 24);
    for (int s = 0; s < KSTEPS; ++s) {
      // one K = 8 step: A advances one k-block (a_lbo), B two k-quads (2 * b_lbo)
      const uint32_t a_step = mode == 1 ? 2 * a_lbo : a_lbo;
      const uint64_t da = swap_a ? make_desc(smem_u32(sA) + s * a_step, a_sbo, a_lbo) : make_desc(smem_u32(sA) + s * a_step, a_lbo, a_sbo);
      const uint64_t db = swap_b ? make_desc(smem_u32(sB) + s * 2 * b_lbo, b_sbo, b_lbo) : make_desc(smem_u32(sB) + s * 2 * b_lbo, b_lbo, b_sbo);
      const uint32_t acc = s > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  // everybody waits for the MMAs
  if (mode != 2) {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok)
                   : "r"(smem_u32(bar))
                   : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[N];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
        "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
        "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < N; ++n) D[tid * N + n] = __uint_as_float(v[n]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

static float tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}

int main(int argc, char **argv) {
  float hA[M * K], hB[N * K], hD[M * N], ref[M * N];
  srand(1);
  for (int i = 0; i < M * K; ++i) hA[i] = tf32((rand() % 2001 - 1000) / 500.0f);
  for (int i = 0; i < N * K; ++i) hB[i] = tf32((rand() % 2001 - 1000) / 500.0f);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)hA[m * K + k] * hB[n * K + k];
      ref[m * N + n] = (float)s;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, sizeof(hA));
  cudaMalloc(&dB, sizeof(hB));
  cudaMalloc(&dD, sizeof(hD));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  // layouts: A core matrices (8 k x 4 m = 128 B) contiguous along m: SBO = 128 B, k-blocks LBO = 128*M/4 B
  //          B core matrices (8 n x 4 k = 128 B): k-quads LBO = 128 B apart, n-blocks SBO = 128 * K/4 B
  const int a_sbo = 128, a_lbo = 128 * (M / 4), b_lbo = 128, b_sbo = 128 * (K / 4);
  const size_t smem = (M * K + 2 * N * K) * 4 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int variant = 4; variant < 11; ++variant) {
    if (variant > 4 && variant < 8) continue;
    const int mode = variant < 4 ? 0 : (variant < 8 ? 1 : variant - 6);
    if (mode == 4) {  // full fp32 inputs, double reference
      for (int i = 0; i < M * K; ++i) hA[i] = (rand() % 200001 - 100000) / 37123.0f;
      for (int i = 0; i < N * K; ++i) hB[i] = (rand() % 200001 - 100000) / 41517.0f;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
          double s = 0;
          for (int k = 0; k < K; ++k) s += (double)hA[m * K + k] * hB[n * K + k];
          ref[m * N + n] = (float)s;
        }
      cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
      cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    }
    const int swap_a = variant & 1, swap_b = (variant >> 1) & 1;
    cudaMemset(dD, 0, sizeof(hD));
    if (mode == 1)
      probe<<<1, 128, smem>>>(dA, dB, dD, 128, 128 * (K / 4), b_lbo, b_sbo, swap_a, swap_b, mode);
    else
      probe<<<1, 128, smem>>>(dA, dB, dD, a_lbo, a_sbo, b_lbo, b_sbo, swap_a, swap_b, mode);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int i = 0; i < M * N; ++i) worst = fmax(worst, fabs((double)hD[i] - ref[i]));
    int nz = 0;
    for (int i = 0; i < M * N; ++i) nz += hD[i] != 0.0f;
    printf("mode %d nonzero %d  D[1][0..2] = %g %g %g  D[33][1] = %g | ", mode, nz, hD[N], hD[N + 1], hD[N + 2], hD[33 * N + 1]);
    printf("variant %d (A desc fields %s, B desc fields %s): %s  max |D - ref| = %g   D[0..3] = %g %g %g %g  ref = %g %g %g %g\n",
           variant, swap_a ? "swapped" : "lbo,sbo", swap_b ? "swapped" : "lbo,sbo", cudaGetErrorString(e), worst, hD[0], hD[1],
           hD[2], hD[3], ref[0], ref[1], ref[2], ref[3]);
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
