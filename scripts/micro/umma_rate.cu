// Microbenchmark: tcgen05.mma issue/throughput for SS vs TS operands and N = 128 / 256, alone and with concurrent
// shared-memory traffic from other warps. Operand contents are garbage (uninitialised smem) - only timing matters.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../fa2_triton_b200/csrc -o umma_rate umma_rate.cu
#include <cstdio>
#include "ptx.cuh"
using namespace fa2;

// mode: 0 = SS N=128, 1 = TS N=128, 2 = SS N=256, 3 = TS N=256 ; traffic: 0 none, 1 LDS.128 stream, 2 STS.128 stream
__global__ void __launch_bounds__(288, 1) k(long long* out, int iters, int mode, int traffic) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc<512>(&slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  __shared__ volatile int done;
  if (threadIdx.x == 0) done = 0;
  __syncthreads();
  if (warp == 0) {
    if (elect_one()) {
      const int N = (mode >= 2) ? 256 : 128;
      const uint32_t idesc = make_idesc_f16(128, N, true, false, false);
      const SmemDesc a = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
      const SmemDesc b = make_smem_desc_sw128(smem_u32(smem) + 32768, 16, 1024);
      long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk / 4) * 16384 + (kk % 4) * 32;
          if (mode & 1) umma_ts(0, 256 + kk * 8, advance(b, off), idesc, 1u);
          else umma_ss(0, advance(a, off), advance(b, off), idesc, 1u);
        }
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      long long t1 = clock64();
      done = 1;
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
  } else if (traffic) {
    uint4* p = reinterpret_cast<uint4*>(smem + 98304) + (threadIdx.x - 32);
    uint4 acc = make_uint4(0, 0, 0, 0);
    long long n = 0;
    while (!done) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (traffic == 1) { uint4 v = p[j * 256]; acc.x ^= v.x; acc.y ^= v.w; }
        else p[j * 256] = acc;
      }
      n += 16;
    }
    if (acc.x == 0x12345678 && blockIdx.x == 0) out[1] = acc.y;
    if (threadIdx.x == 32 && blockIdx.x == 0) out[2] = n;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<512>(0u); }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  const char* mn[] = {"SS N=128", "TS N=128", "SS N=256", "TS N=256"};
  const char* tn[] = {"no other smem traffic", "8 warps LDS.128", "8 warps STS.128"};
  for (int mode = 0; mode < 4; ++mode)
    for (int tr = 0; tr < 3; ++tr) {
      cudaMemset(d, 0, 64);
      k<<<148, 288, 200 * 1024>>>(d, iters, mode, tr);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
      const double per = double(h[0]) / (iters * 8.0);
      const int N = mode >= 2 ? 256 : 128;
      printf("%-9s %-24s: %s  %.1f cycles per MMA (ideal %d)  other-warp smem ops: %.1f B/clk\n", mn[mode], tn[tr], cudaGetErrorString(e), per,
             N / 2, tr ? double(h[2]) * 8 * 32 * 16 / double(h[0]) : 0.0);
    }
  return 0;
}
