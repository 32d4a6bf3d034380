// Microbenchmark: issue rate of FFMA, FFMA2 (fma.rn.f32x2), FADD2, FMNMX3, MUFU.EX2, F2FP, LEA for 1 / 2 warps per SMSP.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)), "l"(reinterpret_cast<uint64_t&>(c)));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)));
  return d;
}

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8], b = threadIdx.x * 1e-3f + 1.0f, c = 0.5f;
  float2 a2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = i + threadIdx.x; a2[i] = make_float2(i, threadIdx.x); }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
      if (MODE == 1) a2[i] = ffma2(a2[i], make_float2(b, b), make_float2(c, c));
      if (MODE == 2) a2[i] = fadd2(a2[i], make_float2(b, c));
      if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
      if (MODE == 5) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(b)); a[i] = __uint_as_float(r); }
      if (MODE == 6) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
      if (MODE == 7) { int r = __float_as_int(a[i]); asm volatile("mad.lo.s32 %0, %0, 8388608, %1;" : "+r"(r) : "r"(__float_as_int(b))); a[i] = __int_as_float(r); }
      if (MODE == 8) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2[i].x)); }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + a2[i].x + a2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<MODE><<<148, threads>>>(out, cyc, iters);
  k<MODE><<<148, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s warps/SMSP=%d: %.2f cycles per warp-instruction (per warp), %.2f issue-cycles per instr per SMSP\n", name, threads / 128,
         double(h) / (iters * 8.0), double(h) / (iters * 8.0) / (threads / 128));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {128, 256, 512}) {
    run<0>("FFMA", threads); run<1>("FFMA2 (f32x2)", threads); run<2>("FADD2 (f32x2)", threads); run<6>("FADD", threads);
    run<3>("MUFU.EX2", threads); run<4>("FMNMX3", threads); run<5>("F2FP.BF16x2", threads); run<7>("IMAD (shl+add)", threads);
    run<8>("FFMA + MUFU interleaved", threads);
  }
  return 0;
}
