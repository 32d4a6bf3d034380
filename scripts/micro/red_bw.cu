// Microbenchmark: throughput of fp32 reduce-add into L2-resident global memory from one CTA per SM.
//   mode 0: red.global.add.v4.f32 from registers (coalesced: lane -> consecutive 16 B)
//   mode 1: red.global.add.f32 scalar (coalesced)
//   mode 2: cp.reduce.async.bulk (smem -> global, 16 KB chunks, 2 in flight)
//   mode 3: st.global.v4 (plain stores, for reference)
// Each CTA adds `iters` tiles of 64 KB into its own 2 MB window (wrapping), like the dQ accumulation of the backward.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bw red_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void __launch_bounds__(128) k_red(float* base, int iters, int mode, int window_floats) {
  extern __shared__ __align__(128) float stage[];  // 2 x 4096 floats
  float* win = base + static_cast<size_t>(blockIdx.x) * window_floats;
  const int t = threadIdx.x;
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  if (mode == 2) {
    for (int i = t; i < 8192; i += 128) stage[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  for (int it = 0; it < iters; ++it) {
    float* tile = win + (static_cast<size_t>(it) * 16384) % window_floats;
    if (mode == 0) {
#pragma unroll 8
      for (int c = 0; c < 32; ++c) {
        float* p = tile + (c * 128 + t) * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      }
    } else if (mode == 1) {
#pragma unroll 8
      for (int c = 0; c < 128; ++c) {
        float* p = tile + c * 128 + t;
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v.x) : "memory");
      }
    } else if (mode == 3) {
#pragma unroll 8
      for (int c = 0; c < 32; ++c) {
        float* p = tile + (c * 128 + t) * 4;
        *reinterpret_cast<float4*>(p) = v;
      }
    } else {
      if (t == 0) {
        for (int c = 0; c < 4; ++c) {
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(tile + c * 4096),
                       "r"(smem_u32(stage + (c & 1) * 4096)), "r"(16384)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
      }
    }
  }
  if (mode == 2 && t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  int sms = 148;
  const int window_floats = 64 * 1024;  // 256 KB per CTA: 37 MB in total, L2 resident
  float* buf;
  cudaMalloc(&buf, static_cast<size_t>(sms) * window_floats * 4);
  cudaMemset(buf, 0, static_cast<size_t>(sms) * window_floats * 4);
  cudaFuncSetAttribute(k_red, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  const int iters = 400;
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t a, b;
      cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a);
      k_red<<<sms, 128, 32768>>>(buf, iters, mode, window_floats);
      cudaEventRecord(b);
      cudaError_t e = cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, a, b);
      double bytes = double(sms) * iters * 65536.0;
      if (rep == 1)
        printf("mode %d: %s  %.3f ms  %.1f GB/s chip  %.1f B/clk/SM @1.9GHz  (%.0f cycles per 64 KB tile)\n", mode,
               cudaGetErrorString(e), ms, bytes / ms * 1e-6, bytes / sms / (ms * 1e-3 * 1.9e9), ms * 1e-3 * 1.9e9 / iters);
    }
  }
  return 0;
}
