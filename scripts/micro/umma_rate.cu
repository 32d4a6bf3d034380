// Microbenchmark: tcgen05.mma throughput (cycles per instruction, one issuing thread per SM) for
//   SS vs TS operands, N = 128 / 256, one or two accumulator tiles, K-major vs MN-major B,
// optionally with concurrent shared-memory traffic from 8 other warps. Everything is a template parameter so that
// the issue loop is as tight as in the real kernels (descriptor add + UTCHMMA). Operand contents are garbage.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../fa2_triton_b200/csrc -o umma_rate umma_rate.cu
#include <cstdio>
#include "ptx.cuh"
using namespace fa2;

template <bool TS, int N, int NACC, bool BMN, bool SAME>
__global__ void __launch_bounds__(288, 1) k(long long* out, int iters, int traffic) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); done = 0; }
    __syncwarp();
    tmem_alloc<512>(&slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(128, N, true, false, BMN);
      const SmemDesc a = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
      const SmemDesc b = BMN ? make_smem_desc_sw128(smem_u32(smem) + 32768, 16384, 1024)
                             : make_smem_desc_sw128(smem_u32(smem) + 32768, 16, 1024);
      long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = SAME ? 0u : (kk / 4) * 16384 + (kk % 4) * 32;
          const uint32_t boff = SAME ? 0u : (BMN ? kk * 2048 : off);
          const uint32_t d = (kk % NACC) * 128;
          if (TS) umma_ts(d, 448 + kk * 8, advance(b, boff), idesc, 1u);
          else umma_ss(d, advance(a, off), advance(b, boff), idesc, 1u);
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
    uint4 acc = make_uint4(threadIdx.x, 0, 0, 0);
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

template <bool TS, int N, int NACC, bool BMN, bool SAME>
void run(const char* name, long long* d) {
  auto kern = k<TS, N, NACC, BMN, SAME>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  for (int tr = 0; tr < 3; ++tr) {
    cudaMemset(d, 0, 64);
    kern<<<148, 288, 200 * 1024>>>(d, iters, tr);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("%-44s %-16s: %s  %6.1f cycles per MMA (math alone: %d)\n", name,
           tr == 0 ? "alone" : (tr == 1 ? "+8 warps LDS.128" : "+8 warps STS.128"), cudaGetErrorString(e),
           double(h[0]) / (iters * 8.0), N / 2);
  }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  run<false, 128, 1, false, false>("SS N=128 K-major B, 1 accumulator", d);
  run<false, 128, 2, false, false>("SS N=128 K-major B, 2 accumulators", d);
  run<false, 128, 1, false, true>("SS N=128 K-major B, same operands", d);
  run<false, 128, 1, true, false>("SS N=128 MN-major B, 1 accumulator", d);
  run<true, 128, 1, false, false>("TS N=128 K-major B, 1 accumulator", d);
  run<true, 128, 1, true, false>("TS N=128 MN-major B, 1 accumulator", d);
  run<true, 128, 2, true, false>("TS N=128 MN-major B, 2 accumulators", d);
  run<true, 128, 1, true, true>("TS N=128 MN-major B, same operands", d);
  run<false, 64, 1, false, false>("SS N=64 K-major B, 1 accumulator", d);
  run<false, 64, 2, false, false>("SS N=64 K-major B, 2 accumulators", d);
  run<false, 64, 1, false, true>("SS N=64 K-major B, same operands", d);
  run<true, 64, 1, false, false>("TS N=64 K-major B, 1 accumulator", d);
  run<true, 64, 1, true, false>("TS N=64 MN-major B, 1 accumulator", d);
  run<false, 32, 1, false, false>("SS N=32 K-major B, 1 accumulator", d);
  run<false, 96, 1, false, false>("SS N=96 K-major B, 1 accumulator", d);
  run<false, 192, 1, false, false>("SS N=192 K-major B, 1 accumulator", d);
  run<false, 256, 1, false, false>("SS N=256 K-major B, 1 accumulator", d);
  run<true, 256, 1, false, false>("TS N=256 K-major B, 1 accumulator", d);
  return 0;
}
