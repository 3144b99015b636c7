// Microbenchmark: throughput of scalar FFMA/FADD vs packed fma.rn.f32x2 / add.f32x2 on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float s) {
  float a[16], b = s + threadIdx.x * 1e-9f, c = 0.5f;
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = i * 0.01f + threadIdx.x * 1e-6f;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // scalar: 16 independent FFMA
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
    } else if (MODE == 1) {  // packed: 8 independent FFMA2 (same flops)
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        uint64_t x, y, z;
        asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(a[i + 1]));
        asm("mov.b64 %0, {%1, %1};" : "=l"(y) : "f"(b));
        asm("mov.b64 %0, {%1, %1};" : "=l"(z) : "f"(c));
        asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(y), "l"(z));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(x));
      }
    } else if (MODE == 2) {  // scalar sub+fma pairs like the SSD loop: 8 x (FSUB, FFMA)
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float d = a[i + 1] - b;
        a[i] = fmaf(d, d, a[i]);
      }
    } else {  // packed sub+fma: 4 x (FSUB2, FFMA2) = same flops as MODE 2
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        uint64_t acc, v, bb, d;
        asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(a[i]), "f"(a[i + 2]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a[i + 1]), "f"(a[i + 3]));
        asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
        asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(v), "l"(bb));
        asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc) : "l"(d));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 2]) : "l"(acc));
      }
    }
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
float run(float *out, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(out, iters, 0.999f);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(out, iters, 0.999f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  float *out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 20000;
  const double lanes = 148.0 * 8 * 256;
  float t0 = run<0>(out, iters), t1 = run<1>(out, iters), t2 = run<2>(out, iters), t3 = run<3>(out, iters);
  printf("scalar FFMA  : %.3f ms  %.2f T lane-FMA/s\n", t0, lanes * iters * 16 / t0 / 1e9);
  printf("packed FFMA2 : %.3f ms  %.2f T lane-FMA/s\n", t1, lanes * iters * 16 / t1 / 1e9);
  printf("scalar SUB+FMA (8 pairs): %.3f ms  %.2f T pairs/s\n", t2, lanes * iters * 8 / t2 / 1e9);
  printf("packed SUB2+FMA2 (4x2 pairs): %.3f ms  %.2f T pairs/s\n", t3, lanes * iters * 8 / t3 / 1e9);
  return 0;
}
