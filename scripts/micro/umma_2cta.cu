// Microbenchmark / probe for round 2: tcgen05.mma.cta_group::2 (a pair of CTAs on the two SMs of a TPC computes one
// 256 x N x 16 product; each CTA holds its own 128 rows of A and HALF of B, so the B-operand shared-memory reads per SM
// halve). Two parts:
//   1. correctness: D[256, N] = A[256, K] * B[K, N] (K = 64) with thread-written SWIZZLE_128B operands against a host
//      reference - pins the descriptor conventions (A: own 128 rows, K-major; B: own N/2 rows of B^T, K-major; idesc M = 256)
//      and the accumulator placement (CTA r gets rows [128 r, 128 r + 128) in its own TMEM);
//   2. rate and power: a long issue loop, 1-CTA 128 x 128 x 16 SS products on every SM vs 2-CTA 256 x 128 x 16 on every
//      pair - cycles per product, and seconds to sample board power for (nvidia-smi in the calling script).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../fa2_triton_b200/csrc -o umma_2cta umma_2cta.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "ptx.cuh"
using namespace fa2;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void umma_ss_2cta(uint32_t d_tmem, SmemDesc a, SmemDesc b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n"
      :
      : "r"(d_tmem), "r"(a.lo), "r"(a.hi), "r"(b.lo), "r"(b.hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// part 1: correctness. A: [256, 64] row-major, Bt: [N, 64] row-major (= B^T), D: [256, N] fp32. One cluster of two CTAs.
// ---------------------------------------------------------------------------------------------------------------------
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const __nv_bfloat16* A, const __nv_bfloat16* Bt, float* D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;           // [128 x 64] bf16, SWIZZLE_128B
  uint8_t* sB = smem + 16384;   // [N/2 x 64] bf16, SWIZZLE_128B
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (warp == 0) tmem_alloc_2cta(&slot);
  // operands: row r of the tile = 128 bytes = 8 chunks of 16 B
  for (int idx = tid; idx < 128 * 8; idx += 128) {
    const int r = idx / 8, c = idx % 8;
    *reinterpret_cast<uint4*>(sA + sw128_offset(r, c)) = *reinterpret_cast<const uint4*>(A + (rank * 128 + r) * 64 + c * 8);
  }
  for (int idx = tid; idx < (N / 2) * 8; idx += 128) {
    const int r = idx / 8, c = idx % 8;
    *reinterpret_cast<uint4*>(sB + sw128_offset(r, c)) = *reinterpret_cast<const uint4*>(Bt + (rank * (N / 2) + r) * 64 + c * 8);
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = slot;
  if (rank == 0 && warp == 0 && elect_one()) {
    const uint32_t idesc = make_idesc_f16(256, N, true, false, false);
    const SmemDesc a = make_smem_desc_sw128(smem_u32(sA), 16, 1024);
    const SmemDesc b = make_smem_desc_sw128(smem_u32(sB), 16, 1024);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_ss_2cta(tmem_base, advance(a, k * 32), advance(b, k * 32), idesc, k > 0);
    umma_commit_2cta(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t regs[32];
    tmem_ld_32x32b_x32(tmem_base + lane_sel + c0, regs);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) D[(rank * 128 + tid) * N + c0 + j] = __uint_as_float(regs[j]);
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc_2cta(tmem_base); }
}

// ---------------------------------------------------------------------------------------------------------------------
// part 2: issue-rate loops (operand contents are whatever shared memory holds)
// ---------------------------------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128, 1) rate_1cta(long long* out, int iters, int fill_zero) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  // operands: pseudo-random bf16 in [-1, 1) (tensor-core power depends on the data: all-zero operands draw far less)
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) {
    uint32_t h = (i + 1u) * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const float a = (int(h & 0xffff) - 32768) / 32768.0f, b = (int(h >> 16) - 32768) / 32768.0f;
    reinterpret_cast<uint32_t*>(smem)[i] = fill_zero ? 0u : pack2<true>(a, b);
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (warp == 0) tmem_alloc<512>(&slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 0 && elect_one()) {
    constexpr uint32_t idesc = make_idesc_f16(128, N, true, false, false);
    const SmemDesc a = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const SmemDesc b = make_smem_desc_sw128(smem_u32(smem) + 32768, 16, 1024);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t off = (k / 4) * 16384 + (k % 4) * 32;
        umma_ss(0, advance(a, off), advance(b, off), idesc, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    if (blockIdx.x == 0) out[0] = clock64() - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<512>(0u); }
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate_2cta(long long* out, int iters, int fill_zero) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5;
  // operands: pseudo-random bf16 in [-1, 1) (tensor-core power depends on the data: all-zero operands draw far less)
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) {
    uint32_t h = (i + 1u) * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const float a = (int(h & 0xffff) - 32768) / 32768.0f, b = (int(h >> 16) - 32768) / 32768.0f;
    reinterpret_cast<uint32_t*>(smem)[i] = fill_zero ? 0u : pack2<true>(a, b);
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (warp == 0) tmem_alloc_2cta(&slot);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  if (rank == 0 && warp == 0 && elect_one()) {
    constexpr uint32_t idesc = make_idesc_f16(256, N, true, false, false);
    const SmemDesc a = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const SmemDesc b = make_smem_desc_sw128(smem_u32(smem) + 32768, 16, 1024);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t off = (k / 4) * 16384 + (k % 4) * 32;
        umma_ss_2cta(slot, advance(a, off), advance(b, off), idesc, 1u);
      }
    }
    umma_commit_2cta(&bar);
    mbar_wait(&bar, 0);
    if (blockIdx.x == 0) out[0] = clock64() - t0;
  } else {
    mbar_wait(&bar, 0);  // the peer's barrier receives the multicast arrive as well
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc_2cta(slot); }
}

template <int N>
static int run_probe() {
  const int K = 64;
  std::vector<__nv_bfloat16> hA(256 * K), hB(N * K);
  std::vector<float> fA(256 * K), fB(N * K);
  srand(1);
  for (int i = 0; i < 256 * K; ++i) { float v = (rand() % 17 - 8) / 8.0f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { float v = (rand() % 13 - 6) / 4.0f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 256 * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, 256 * N * 4);
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  probe<N><<<2, 128, 65536>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> hD(256 * N);
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += double(fA[m * K + k]) * fB[n * K + k];
      worst = fmax(worst, fabs(ref - hD[m * N + n]));
    }
  printf("2-CTA probe N=%d: %s, max |err| = %.3g %s\n", N, cudaGetErrorString(e), worst, worst < 1e-3 ? "OK" : "MISMATCH");
  return worst < 1e-3 ? 0 : 1;
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;     // 0: probe + rates, 1: long 1-CTA loop, 2 / 3: long 2-CTA loop N = 128 / 256 (for power); argv[3] = 1: zero operands
  const double seconds = argc > 2 ? atof(argv[2]) : 3.0;
  const int zero = argc > 3 ? atoi(argv[3]) : 0;
  long long* d; cudaMalloc(&d, 64);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(rate_1cta<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
  cudaFuncSetAttribute(rate_2cta<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
  cudaFuncSetAttribute(rate_2cta<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
  if (mode == 0) {
    int bad = run_probe<128>() + run_probe<256>();
    const int iters = 4000;
    long long h;
    rate_1cta<128><<<sms, 128, 98304>>>(d, iters, 0); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("1-CTA 128x128x16 SS: %.1f cycles per product (%s)\n", double(h) / (iters * 8.0), cudaGetErrorString(cudaGetLastError()));
    rate_2cta<128><<<sms / 2 * 2, 128, 98304>>>(d, iters, 0); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("2-CTA 256x128x16 SS: %.1f cycles per product (%s)\n", double(h) / (iters * 8.0), cudaGetErrorString(cudaGetLastError()));
    rate_2cta<256><<<sms / 2 * 2, 128, 98304>>>(d, iters, 0); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("2-CTA 256x256x16 SS: %.1f cycles per product (%s)\n", double(h) / (iters * 8.0), cudaGetErrorString(cudaGetLastError()));
    return bad;
  }
  // power runs: repeat a ~50 ms launch for `seconds`
  const int iters = 150000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  int launches = 0;
  float ms = 0;
  do {
    if (mode == 1) rate_1cta<128><<<sms, 128, 98304>>>(d, iters, zero);
    else if (mode == 2) rate_2cta<128><<<sms / 2 * 2, 128, 98304>>>(d, iters, zero);
    else rate_2cta<256><<<sms / 2 * 2, 128, 98304>>>(d, iters / 2, zero);
    ++launches;
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  } while (ms < seconds * 1e3);
  const double flop = double(launches) * iters * 8.0 * 2.0 * 128 * 128 * 16 * sms;  // per SM 128 x 128 x 16 either way
  printf("mode %d: %d launches, %.1f ms, %.0f TFLOP/s (%s)\n", mode, launches, ms, flop / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
