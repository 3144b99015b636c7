// Microbenchmark of the two-row dot block's FFMA2 stream (match_sweep2.cuh: dot2_block) on sm_100a:
// how many cycles per packed FFMA2 does one quadrant sustain with the kernel's real operand pattern
// (32 distinct accumulator pairs, 4 a-pairs and 10 broadcast scalars per channel), for several
// instruction orders, with operands from registers only and with the slab LDS included?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_dot2 ubench_dot2.cu && ./ubench_dot2
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// ORDER 0: row, pp, j (as dot2_block)   1: j, row, pp   2: pp, j, row (same b for two consecutive)
template <int ORDER, bool LDS, int CT>
__global__ void __launch_bounds__(256, 1) k(float *out, const float *in, int iters, long long *cycles) {
  extern __shared__ float slab[];
  for (int i = threadIdx.x; i < CT * 192 + 192; i += blockDim.x) slab[i] = in[i];
  __syncthreads();
  float2 a2[2][CT][2];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int kk = 0; kk < CT; ++kk)
#pragma unroll
      for (int pp = 0; pp < 2; ++pp) a2[r][kk][pp] = make_float2(in[(r * CT + kk) * 2 + pp + threadIdx.x], in[7 + threadIdx.x + kk]);
  float2 acc[2][2][8];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[r][pp][j] = make_float2(0.f, 0.f);
  const float *brow = slab + (threadIdx.x & 31) * 4;
  float breg[2][12];   // two alternating sets of "slab" values (a full [CT][12] would not fit)
  if (!LDS) {
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
#pragma unroll
      for (int j = 0; j < 12; ++j) breg[kk][j] = in[kk * 12 + j + (threadIdx.x & 3)];
  }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int kk = 0; kk < CT; ++kk) {
      float b[12];
      if (LDS) {
        const float4 *src = reinterpret_cast<const float4 *>(brow + kk * 192 + (it & 3) * 8);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          float4 t = src[j];
          b[4 * j] = t.x; b[4 * j + 1] = t.y; b[4 * j + 2] = t.z; b[4 * j + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 12; ++j) b[j] = breg[kk & 1][j];
      }
      if (ORDER == 0) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int pp = 0; pp < 2; ++pp)
#pragma unroll
            for (int j = 0; j < 8; ++j)
              acc[r][pp][j] = ffma2(a2[r][kk][pp], make_float2(b[2 * pp + j], b[2 * pp + j]), acc[r][pp][j]);
      } else if (ORDER == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int pp = 0; pp < 2; ++pp)
              acc[r][pp][j] = ffma2(a2[r][kk][pp], make_float2(b[2 * pp + j], b[2 * pp + j]), acc[r][pp][j]);
      } else if (ORDER == 2) {
#pragma unroll
        for (int pp = 0; pp < 2; ++pp)
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int r = 0; r < 2; ++r)
              acc[r][pp][j] = ffma2(a2[r][kk][pp], make_float2(b[2 * pp + j], b[2 * pp + j]), acc[r][pp][j]);
      } else if (ORDER == 3) {
        // snake: consecutive instructions share either the a pair (same row) or the b scalar (same j)
#pragma unroll
        for (int pp = 0; pp < 2; ++pp)
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const int r = (j & 1) ? 1 - rr : rr;
              acc[r][pp][j] = ffma2(a2[r][kk][pp], make_float2(b[2 * pp + j], b[2 * pp + j]), acc[r][pp][j]);
            }
      } else {
        // snake over all four a pairs: (r0,p0) (r1,p0) | b changes | (r1,p0)... with the j of pp=1 shifted by 2
        // so that b[2pp + j] is shared by four consecutive instructions
#pragma unroll
        for (int c = 0; c < 10; ++c)
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            const int q = (c & 1) ? 3 - s4 : s4;   // which a pair: snake through the four
            const int r = q >> 1, pp = q & 1, j = c - 2 * pp;
            if (j >= 0 && j < 8)
              acc[r][pp][j] = ffma2(a2[r][kk][pp], make_float2(b[c], b[c]), acc[r][pp][j]);
          }
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += acc[r][pp][j].x + acc[r][pp][j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// scalar reference: the same sums with FFMA (64 per channel)
template <int CT>
__global__ void __launch_bounds__(256, 1) ks(float *out, const float *in, int iters, long long *cycles) {
  float a[2][CT][4];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int kk = 0; kk < CT; ++kk)
#pragma unroll
      for (int p = 0; p < 4; ++p) a[r][kk][p] = in[(r * CT + kk) * 4 + p + threadIdx.x];
  float acc[2][4][8];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[r][p][j] = 0.f;
  float breg[2][12];
#pragma unroll
  for (int kk = 0; kk < 2; ++kk)
#pragma unroll
    for (int j = 0; j < 12; ++j) breg[kk][j] = in[kk * 12 + j + (threadIdx.x & 3)];
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int kk = 0; kk < CT; ++kk)
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[r][p][j] = fmaf(a[r][kk][p], breg[kk & 1][(p & 2) + j], acc[r][p][j]);
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += acc[r][p][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <class F>
void run(const char *name, F kern, int threads, int iters, double ffma2_per_iter, float *out, const float *in, long long *cyc) {
  size_t smem = (10 * 192 + 192) * 4;
  kern<<<148, threads, smem>>>(out, in, 10, cyc);
  kern<<<148, threads, smem>>>(out, in, iters, cyc);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const int warps_per_quadrant = threads / 128;
  printf("%-44s %d warp(s)/scheduler: %6.3f cycles per packed op per scheduler (%s)\n", name, warps_per_quadrant,
         (double)c / (ffma2_per_iter * iters * warps_per_quadrant), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float *out, *in;
  long long *cyc;
  cudaMalloc(&out, 148 * 256 * 4);
  cudaMalloc(&in, 1 << 16);
  cudaMemset(in, 0, 1 << 16);
  cudaMalloc(&cyc, 8);
  const int iters = 2000;
  for (int threads : {128, 256}) {
    run("FFMA2 order row,pp,j   regs only", k<0, false, 10>, threads, iters, 320, out, in, cyc);
    run("FFMA2 order j,row,pp   regs only", k<1, false, 10>, threads, iters, 320, out, in, cyc);
    run("FFMA2 order pp,j,row   regs only", k<2, false, 10>, threads, iters, 320, out, in, cyc);
    run("FFMA2 snake (j,row)     regs only", k<3, false, 10>, threads, iters, 320, out, in, cyc);
    run("FFMA2 snake b-major     regs only", k<4, false, 10>, threads, iters, 320, out, in, cyc);
    run("FFMA2 snake (j,row)     + slab LDS", k<3, true, 10>, threads, iters, 320, out, in, cyc);
    run("FFMA2 snake b-major     + slab LDS", k<4, true, 10>, threads, iters, 320, out, in, cyc);
    run("FFMA2 order row,pp,j   + slab LDS", k<0, true, 10>, threads, iters, 320, out, in, cyc);
    run("FFMA2 order pp,j,row   + slab LDS", k<2, true, 10>, threads, iters, 320, out, in, cyc);
    run("scalar FFMA (640 per channel-block)", ks<10>, threads, iters, 640, out, in, cyc);
  }
  return 0;
}
