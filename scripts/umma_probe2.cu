// Probe 2: tcgen05.mma kind::tf32 with an MN-major B operand (N contiguous in shared memory, as image
// rows are) -- no-swizzle interleave and SWIZZLE_128B, including start addresses that are whole rows
// (128 B) into a swizzle atom (what a sliding window over image rows needs).
// A: K-major no-swizzle (known good, umma_probe.cu).  D[128 x N] = A[128 x K] * B[N x K]^T.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int M = 128, N = 64, KSTEPS = 2, K = 8 * KSTEPS, ROWS = 32;  // the B tile holds ROWS k-rows; the MMA uses K of them from row r0

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

// mode 0: B no-swizzle interleave: (n%4) + 4*(k%8) + 32*(n/4) [n-groups of 4: 128 B] + (32*N/4)*(k/8)
// mode 1: B SWIZZLE_128B: per n-group of 32: byte = k*128 + (((n%32)/4) ^ (k%8))*16 + (n%4)*4, groups ROWS*128 B apart
__global__ void probe(const float *A, const float *Bfull, float *D, int mode, int r0, int lbo, int sbo, int base_off, int swap) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float *sB = reinterpret_cast<float *>(smem);                 // 1024-aligned
  float *sA = sB + N * ROWS;
  uint64_t *bar = reinterpret_cast<uint64_t *>(sA + M * K);
  uint32_t *tptr = reinterpret_cast<uint32_t *>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < M * K; i += blockDim.x) {
    const int m = i / K, k = i % K;
    sA[(k % 4) + 4 * (m % 8) + (32 * K / 4) * (m / 8) + 32 * (k / 4)] = A[m * K + k];
  }
  for (int i = tid; i < N * ROWS; i += blockDim.x) {
    const int n = i / ROWS, k = i % ROWS;   // Bfull[n][row]
    const float v = Bfull[n * ROWS + k];
    if (mode == 0)
      sB[(n % 4) + 4 * (k % 8) + 32 * (n / 4) + (32 * N / 4) * (k / 8)] = v;
    else
      sB[(n / 32) * ROWS * 32 + k * 32 + ((((n % 32) / 4) ^ (k % 8)) * 4) + (n % 4)] = v;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tptr)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tptr;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(M >> 4) << 24);
    for (int s = 0; s < KSTEPS; ++s) {
      const uint64_t da = make_desc(smem_u32(sA) + s * 256, 128, 32 * K, 0, 0);
      uint32_t baddr;
      if (mode == 0) baddr = smem_u32(sB) + (uint32_t)(((r0 / 8) + s) * (128 * N / 4)) ;  // r0 must be a multiple of 8 here
      else baddr = smem_u32(sB) + (uint32_t)((r0 + 8 * s) * 128);
      const uint64_t db = swap ? make_desc(baddr, sbo, lbo, mode == 1 ? 2 : 0, base_off) : make_desc(baddr, lbo, sbo, mode == 1 ? 2 : 0, base_off);
      const uint32_t acc = s > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok)
                   : "r"(smem_u32(bar))
                   : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int n0 = 0; n0 < N; n0 += 16) {
    uint32_t v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr + n0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[tid * N + n0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

static float tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}

int main() {
  static float hA[M * K], hB[N * ROWS], hD[M * N], ref[M * N];
  srand(2);
  for (int i = 0; i < M * K; ++i) hA[i] = tf32((rand() % 2001 - 1000) / 500.0f);
  for (int i = 0; i < N * ROWS; ++i) hB[i] = tf32((rand() % 2001 - 1000) / 500.0f);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, sizeof(hA));
  cudaMalloc(&dB, sizeof(hB));
  cudaMalloc(&dD, sizeof(hD));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  const size_t smem = (N * ROWS + M * K) * 4 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  struct Case { int mode, r0, lbo, sbo, base_off, swap; const char *what; };
  const Case cases[] = {
      {0, 0, 128 * N / 4, 128, 0, 0, "interleave lbo=kgroup sbo=ngroup"},
      {0, 0, 128 * N / 4, 128, 0, 1, "interleave swapped"},
      {0, 8, 128 * N / 4, 128, 0, 0, "interleave r0=8"},
      {1, 0, ROWS * 128, 1024, 0, 0, "sw128 r0=0 lbo=ngroup(32 cols) sbo=1024"},
      {1, 0, ROWS * 128, 1024, 0, 1, "sw128 r0=0 swapped"},
      {1, 8, ROWS * 128, 1024, 0, 0, "sw128 r0=8"},
      {1, 3, ROWS * 128, 1024, 0, 0, "sw128 r0=3 base_off=0"},
      {1, 3, ROWS * 128, 1024, 3, 0, "sw128 r0=3 base_off=3"},
      {1, 5, ROWS * 128, 1024, 5, 0, "sw128 r0=5 base_off=5"},
      {1, 3, ROWS * 128, 1024, 0, 1, "sw128 r0=3 base_off=0 swapped"},
      {1, 3, ROWS * 128, 1024, 3, 1, "sw128 r0=3 base_off=3 swapped"},
  };
  for (const Case &c : cases) {
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += (double)hA[m * K + k] * hB[n * ROWS + c.r0 + k];
        ref[m * N + n] = (float)s;
      }
    cudaMemset(dD, 0, sizeof(hD));
    probe<<<1, 128, smem>>>(dA, dB, dD, c.mode, c.r0, c.lbo, c.sbo, c.base_off, c.swap);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    double worst = 0;
    int nz = 0, good = 0;
    for (int i = 0; i < M * N; ++i) {
      worst = fmax(worst, fabs((double)hD[i] - ref[i]));
      nz += hD[i] != 0.0f;
      good += fabs((double)hD[i] - ref[i]) < 1e-4;
    }
    printf("%-48s %s nonzero %d correct %d/%d max err %g  D[0][0..1]=%g %g ref %g %g  D[0][32]=%g ref %g\n", c.what,
           cudaGetErrorString(e), nz, good, M * N, worst, hD[0], hD[1], ref[0], ref[1], hD[32], ref[32]);
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
