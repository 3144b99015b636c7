// ubench_lds128.cu -- how does B200 split a 32-lane LDS.128 into wavefronts?  One warp, 16 warps of back-to-back
// LDS.128 with a given lane -> address pattern; reports cycles per instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o ubench_lds128 ubench_lds128.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int REP = 4096;
__global__ void k(const int *offs, long long *out, float *sink) {
  extern __shared__ uint4 sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = make_uint4(i, 1, 2, 3);
  __syncthreads();
  const int o = offs[threadIdx.x & 31];   // in 16-byte units
  uint4 acc = make_uint4(0, 0, 0, 0);
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm + o);
  const long long t0 = clock64();
#pragma unroll 8
  for (int r = 0; r < REP; ++r) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + (uint32_t)((r & 7) * 8192)));
    acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = clock64() - t0;
  sink[threadIdx.x] = __uint_as_float(acc.x + acc.y + acc.z + acc.w) + (float)(t1 - t0);
}
int main() {
  int *d; long long *o, h; float *s;
  cudaMalloc(&d, 128); cudaMalloc(&o, 8); cudaMalloc(&s, 4096);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
  struct { const char *name; int (*f)(int); } pats[] = {
      {"contiguous 512 B (lane * 16 B)", [](int l) { return l; }},
      {"4 rows 128 B apart in banks... rows at +1024 B: quarter q reads row q, 128 B contiguous", [](int l) { return (l >> 3) * 64 + (l & 7); }},
      {"same, rows at +1984 B (64 B bank shift per row)", [](int l) { return (l >> 3) * 124 + (l & 7); }},
      {"same, rows at +1056 B (32 B shift per row)", [](int l) { return (l >> 3) * 66 + (l & 7); }},
      {"strip kernel: lane (q, hb, ddy): row ddy at +1920 B, 16 B * q + 64 B * hb", [](int l) { return (l >> 3) * 120 + ((l >> 2) & 1) * 4 + (l & 3); }},
      {"half-warp pairs: lanes 0-7 row 0, lanes 8-15 row 0 + 128 B, lanes 16-23 row 1, ...", [](int l) { return (l >> 4) * 64 + (l & 15); }},
      {"all lanes one address (broadcast)", [](int l) { return 0; }},
      {"4 distinct addresses (a-ring read): lane & 3", [](int l) { return l & 3; }},
      {"8 rows x 4 quads, rows 1920 B apart (tail items)", [](int l) { return (l >> 2) * 120 + (l & 3); }},
      {"8 rows x 4 quads, rows 1920 + 64 B apart", [](int l) { return (l >> 2) * 124 + (l & 3); }},
      {"2-way conflict inside every quarter (lanes 0-3 and 4-7 same banks, different rows)", [](int l) { return (l >> 3) * 8 + ((l >> 2) & 1) * 64 + (l & 3); }},
  };
  for (auto &p : pats) {
    int hoffs[32];
    for (int l = 0; l < 32; ++l) hoffs[l] = p.f(l);
    cudaMemcpy(d, hoffs, 128, cudaMemcpyHostToDevice);
    k<<<1, 512, 8192 * 16>>>(d, o, s);
    cudaDeviceSynchronize();
    k<<<1, 512, 8192 * 16>>>(d, o, s);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, o, 8, cudaMemcpyDeviceToHost);
    printf("%-100s %.2f SM cycles per warp LDS.128 (16 warps)  %s\n", p.name, (double)h / REP / 16, cudaGetErrorString(e));
  }
  return 0;
}
