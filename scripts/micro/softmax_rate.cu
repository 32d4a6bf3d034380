// Microbenchmark: cycles per 64-key half-block of the forward kernel's softmax loop body (fwd2_sm100.cu), without the
// tensor pipe: thread = row, 64 scores loaded from tensor memory, block maximum, 64 exponentials, row sum, 16-bit pack,
// 32 columns stored back. 1 or 2 warps per scheduler (the kernel has 2: one per Q tile). Variants:
//   0 load + max + exp + store in that order (the kernel before speculation)   1 exponentials merged with the maximum
//   2 load + exp + store (no maximum at all)   3 load + max + store (no exponentials)
//   4 variant 1 with exponentials software-pipelined by hand (group of 8 issued before the previous group is consumed)
// kEmul of every 4 pairs are evaluated on the FMA pipes (ex2_emul2).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../fa2_triton_b200/csrc -o softmax_rate softmax_rate.cu
#include <cstdio>
#include "ptx.cuh"
using namespace fa2;

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

template <int V, int kEmul>
__global__ void __launch_bounds__(256, 1) k(long long* out, float* sink, int iters, float c, int nwarps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t tS = lane_sel + (warp >> 2) * 128;
  {  // plausible scores in both buffers
    uint32_t init[32];
    for (int b = 0; b < 4; ++b) {
      for (int e = 0; e < 32; ++e) init[e] = __float_as_uint(-3.f + 0.05f * ((threadIdx.x * 7 + e * 13 + b * 5) % 97));
      tmem_st_32x32b_x32(tS + b * 32, init);
    }
    tmem_wait_st();
  }
  __syncthreads();
  float m_ref = 2.f, l = 0.f;
  const float2 c2 = make_float2(c, c);
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      const uint32_t tSb = tS + (it & 1) * 64;
      float s[64];
      {
        uint32_t su[2][32];
        tmem_ld_32x32b_x32(tSb, su[0]);
        tmem_ld_32x32b_x32(tSb + 32, su[1]);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 64; ++e) s[e] = __uint_as_float(su[e / 32][e % 32]);
      }
      uint32_t pk[32];
      float2 ls[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      const float neg_m = -m_ref;
      const float2 neg_m2 = make_float2(neg_m, neg_m);
      auto exp_pair = [&](int e) {
        const float2 x2 = ffma2(make_float2(s[e], s[e + 1]), c2, neg_m2);
        float2 p2;
        if (((e / 2) % 4) < kEmul) {
          p2 = ex2_emul2(x2);
        } else {
          p2.x = ex2_approx(x2.x);
          p2.y = ex2_approx(x2.y);
        }
        ls[(e / 2) % 2] = fadd2(ls[(e / 2) % 2], p2);
        pk[e / 2] = pack2<true>(p2.x, p2.y);
      };
      float mb = 0.f;
      if (V == 0 || V == 3) {
        float mx[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) mx[e] = fmax3(s[e], s[e + 4], s[e + 8]);
#pragma unroll
        for (int e = 12; e < 60; e += 8) {
#pragma unroll
          for (int q = 0; q < 4; ++q) mx[q] = fmax3(mx[q], s[e + q], s[e + q + 4]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) mx[q] = fmaxf(mx[q], s[60 + q]);
        mb = fmax3(fmaxf(mx[0], mx[1]), mx[2], mx[3]) * c;
        if (mb > m_ref + 8.f) m_ref = mb;  // never true with these scores; keeps the dependency of the kernel
      }
      if (V == 0 || V == 2) {
        const float nm = -m_ref;
        const float2 nm2 = make_float2(nm, nm);
#pragma unroll
        for (int e = 0; e < 64; e += 2) {
          const float2 x2 = ffma2(make_float2(s[e], s[e + 1]), c2, nm2);
          float2 p2;
          if (((e / 2) % 4) < kEmul) {
            p2 = ex2_emul2(x2);
          } else {
            p2.x = ex2_approx(x2.x);
            p2.y = ex2_approx(x2.y);
          }
          ls[(e / 2) % 2] = fadd2(ls[(e / 2) % 2], p2);
          pk[e / 2] = pack2<true>(p2.x, p2.y);
        }
      }
      if (V == 1) {
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int e = 0; e < 64; e += 8) {
#pragma unroll
          for (int q = 0; q < 4; ++q) mx[q] = fmax3(mx[q], s[e + q], s[e + q + 4]);
#pragma unroll
          for (int u = 0; u < 8; u += 2) exp_pair(e + u);
        }
        mb = fmax3(fmaxf(mx[0], mx[1]), mx[2], mx[3]) * c;
        if (__any_sync(0xffffffffu, mb > m_ref + 8.f)) m_ref = mb;
      }
      if (V == 4) {
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        float x[64];
#pragma unroll
        for (int e = 0; e < 64; e += 2) {
          const float2 x2 = ffma2(make_float2(s[e], s[e + 1]), c2, neg_m2);
          x[e] = x2.x;
          x[e + 1] = x2.y;
        }
#pragma unroll
        for (int g = 0; g <= 8; ++g) {
          if (g < 8) {
#pragma unroll
            for (int e = g * 8; e < g * 8 + 8; ++e) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x[e]) : "f"(x[e]));
#pragma unroll
            for (int q = 0; q < 4; ++q) mx[q] = fmax3(mx[q], s[g * 8 + q], s[g * 8 + q + 4]);
          }
          if (g > 0) {
#pragma unroll
            for (int e = (g - 1) * 8; e < g * 8; e += 2) {
              float2 p2 = make_float2(x[e], x[e + 1]);
              asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(reinterpret_cast<uint64_t&>(ls[(e / 2) % 2])) : "l"(reinterpret_cast<uint64_t&>(p2)));
              pk[e / 2] = pack2<true>(x[e], x[e + 1]);
            }
          }
        }
        mb = fmax3(fmaxf(mx[0], mx[1]), mx[2], mx[3]) * c;
        if (__any_sync(0xffffffffu, mb > m_ref + 8.f)) m_ref = mb;
      }
      if (V == 3) {
#pragma unroll
        for (int e = 0; e < 32; ++e) pk[e] = __float_as_uint(s[2 * e] + mb);
      }
      l += (ls[0].x + ls[1].x) + (ls[0].y + ls[1].y);
      tmem_st_32x32b_x32(tSb, pk);
      tmem_wait_st();
      // put scores back so that the next visit of this buffer sees floats again (not timed separately; 32 columns)
      m_ref += l * 1e-30f;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (l == 12345.f) sink[threadIdx.x] = l + m_ref;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<512>(0u); }
}

template <int V, int kEmul>
void run(const char* name, long long* d, float* sink) {
  const int iters = 2000;
  for (int nw : {4, 8}) {
    cudaMemset(d, 0, 64);
    k<V, kEmul><<<148, 256>>>(d, sink, iters, 0.1275f, nw);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-52s emul %d/4, %d warp(s) per scheduler: %s %7.1f cycles per half-block\n", name, kEmul, nw / 4,
           cudaGetErrorString(e), double(h) / iters);
  }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  float* sink; cudaMalloc(&sink, 4096);
  run<0, 0>("load, max, exp, store (kernel before)", d, sink);
  run<1, 0>("load, max merged into exp, store", d, sink);
  run<4, 0>("load, max merged, exp pipelined by hand, store", d, sink);
  run<2, 0>("load, exp, store (no max)", d, sink);
  run<3, 0>("load, max, store (no exp)", d, sink);
  run<0, 1>("load, max, exp, store (kernel before)", d, sink);
  run<1, 1>("load, max merged into exp, store", d, sink);
  run<0, 2>("load, max, exp, store (kernel before)", d, sink);
  run<1, 2>("load, max merged into exp, store", d, sink);
  run<2, 2>("load, exp, store (no max)", d, sink);
  run<1, 3>("load, max merged into exp, store", d, sink);
  run<1, 4>("load, max merged into exp, store", d, sink);
  return 0;
}
