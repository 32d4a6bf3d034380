// Microbenchmark: how fast can an SM stream data it produced out to global memory (the dS spill of the dK/dV kernel)?
//   mode 0: st.global.v4        mode 1: st.global.cs.v4 (streaming)   mode 2: cp.async.bulk shared -> global, 32 KB chunks
// W warps per CTA (one CTA per SM), each warp instruction writes 512 contiguous bytes; per CTA 32 KB contiguous "tiles".
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../fa2_triton_b200/csrc -o store_bw store_bw.cu
#include <cstdio>
#include "ptx.cuh"
using namespace fa2;

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(uint8_t* out, long long tiles_per_cta, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warps = blockDim.x / 32, warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  uint8_t* base = out + (long long)blockIdx.x * tiles_per_cta * 32768;
  __syncthreads();
  const long long t0 = clock64();
  if (MODE == 2) {
    if (threadIdx.x == 0) {
      const uint64_t pol = l2_policy_evict_first();
      for (long long t = 0; t < tiles_per_cta; ++t) {
        bulk_store_1d_hint(base + t * 32768, smem + (t & 1) * 32768, 32768, pol);
        bulk_commit_group();
        bulk_wait_group_read<1>();
      }
      bulk_wait_group<0>();
    }
  } else {
    // a 32 KB tile = 64 warp-stores of 512 B, dealt round-robin to the warps
    for (long long t = 0; t < tiles_per_cta; ++t) {
      uint8_t* tile = base + t * 32768;
      for (int s = warp; s < 64; s += warps) {
        uint8_t* p = tile + s * 512 + lane * 16;
        if (MODE == 0) asm volatile("st.global.v4.b32 [%0], {%1, %1, %1, %1};" ::"l"(p), "r"(lane) : "memory");
        else asm volatile("st.global.cs.v4.b32 [%0], {%1, %1, %1, %1};" ::"l"(p), "r"(lane) : "memory");
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const long long tiles = 512;  // 16 MB per CTA, 2.4 GB in total
  uint8_t* out; long long* cyc;
  cudaMalloc(&out, (size_t)sms * tiles * 32768);
  cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int mode = 0; mode < 3; ++mode) {
    for (int warps : {4, 8, 16}) {
      if (mode == 2 && warps != 4) continue;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<sms, warps * 32, 0>>>(out, tiles, cyc);
        else if (mode == 1) k<1><<<sms, warps * 32, 0>>>(out, tiles, cyc);
        else k<2><<<sms, warps * 32, 65536>>>(out, tiles, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)sms * tiles * 32768;
      printf("mode %d (%s) %2d warps/SM: %.3f ms  %.2f TB/s  %.1f B/clk/SM (CTA 0: %lld cycles)  %s\n", mode,
             mode == 0 ? "st.global.v4" : mode == 1 ? "st.global.cs.v4" : "bulk smem->global", warps, ms, bytes / ms / 1e9,
             (double)tiles * 32768 / c, c, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
