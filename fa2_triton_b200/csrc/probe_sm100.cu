// UMMA / TMA probe: D[128, N] = A[128, K] * B[K, N] through each operand path used by the attention kernels
// (K-major smem via TMA, MN-major smem via TMA or thread-written SWIZZLE_128B, A from tensor memory).
// Test infrastructure (tests/test_umma_probe.py): pins the descriptor encodings in ptx.cuh on real hardware.
#include "fa2_kernels.cuh"
#include "ptx.cuh"

namespace fa2 {

template <bool kBf16>
__global__ void __launch_bounds__(128, 1)
fa2_probe_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                // up to 32 KB
  uint8_t* sB = smem + 32768;        // up to 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536);
  uint64_t* load_bar = bars;
  uint64_t* mma_bar = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int N = p.N, K = p.K;
  const int ksub = K / 64;  // 64-wide K chunks (K-major operands)
  const int nsub = N / 64;  // 64-wide N chunks (MN-major B)

  if (warp == 0) {
    if (tid == 0) {
      mbar_init(load_bar, 1);
      mbar_init(mma_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
  const uint16_t* ga = reinterpret_cast<const uint16_t*>(p.a);

  // ---- operand A ----
  if (p.a_mode == 1) {
    // global [K, 128] row-major -> smem MN-major SWIZZLE_128B: two [K x 64] sub-tiles of K*128 bytes
    for (int idx = tid; idx < K * 16; idx += 128) {
      const int k = idx / 16, c = idx % 16;  // c: 16-byte chunk along M (8 elements)
      const uint4 v = *reinterpret_cast<const uint4*>(ga + k * 128 + c * 8);
      *reinterpret_cast<uint4*>(sA + (c / 8) * (K * 128) + sw128_offset(k, c % 8)) = v;
    }
    fence_proxy_async_smem();
  } else if (p.a_mode == 3 || p.a_mode == 4) {
    // global [K, 128] row-major -> smem MN-major, no swizzle: 16 slabs (8 M-elements each) of [K rows x 16 bytes]
    for (int idx = tid; idx < K * 16; idx += 128) {
      const int k = idx / 16, c = idx % 16;
      const uint4 v = *reinterpret_cast<const uint4*>(ga + k * 128 + c * 8);
      *reinterpret_cast<uint4*>(sA + c * (K * 16) + k * 16) = v;
    }
    fence_proxy_async_smem();
  } else if (p.a_mode == 2) {
    // global [128, K] row-major -> TMEM columns [256, 256 + K/2), lane == row, two 16-bit values per column
    for (int c0 = 0; c0 < K / 2; c0 += 32) {
      uint32_t regs[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) regs[j] = *reinterpret_cast<const uint32_t*>(ga + tid * K + (c0 + j) * 2);
      tmem_st_32x32b_x32(tmem_base + lane_sel + 256 + c0, regs);
    }
    tmem_wait_st();
    tcgen05_fence_before();
  }
  __syncthreads();

  if (tid == 0) {
    tcgen05_fence_after();
    uint32_t bytes = 0;
    if (p.a_mode == 0) bytes += 128 * K * 2;
    bytes += N * K * 2;
    mbar_arrive_expect_tx(load_bar, bytes);
    if (p.a_mode == 0)
      for (int j = 0; j < ksub; ++j) tma_load_4d(sA + j * 16384, &tmap_a, load_bar, j * 64, 0, 0, 0);
    if (p.b_mode == 0) {
      for (int j = 0; j < ksub; ++j) tma_load_4d(sB + j * (N * 128), &tmap_b, load_bar, j * 64, 0, 0, 0);
    } else {
      for (int j = 0; j < nsub; ++j) tma_load_4d(sB + j * (K * 128), &tmap_b, load_bar, j * 64, 0, 0, 0);
    }
    mbar_wait(load_bar, 0);
    tcgen05_fence_after();

    const uint32_t idesc = make_idesc_f16(128, N, kBf16, p.a_mode == 1 || p.a_mode >= 3, p.b_mode == 1);
    const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
    for (int k = 0; k < K / 16; ++k) {
      SmemDesc bdesc;
      if (p.b_mode == 0)
        bdesc = make_smem_desc_sw128(b_addr + (k / 4) * (N * 128) + (k % 4) * 32, 16, 1024);
      else
        bdesc = make_smem_desc_sw128(b_addr + k * 2048, K * 128, 1024);
      if (p.a_mode == 2) {
        umma_ts(tmem_base, tmem_base + 256 + k * 8, bdesc, idesc, k > 0);
      } else {
        SmemDesc adesc;
        if (p.a_mode == 0)
          adesc = make_smem_desc_sw128(a_addr + (k / 4) * 16384 + (k % 4) * 32, 16, 1024);
        else if (p.a_mode == 3)  // LBO: 8-row groups along K (128 B apart), SBO: 8-element groups along M (slabs)
          adesc = make_smem_desc_noswizzle(a_addr + k * 256, 128, K * 16);
        else if (p.a_mode == 4)  // the other reading of LBO / SBO (expected to be wrong; kept for the probe script)
          adesc = make_smem_desc_noswizzle(a_addr + k * 256, K * 16, 128);
        else
          adesc = make_smem_desc_sw128(a_addr + k * 2048, K * 128, 1024);
        umma_ss(tmem_base, adesc, bdesc, idesc, k > 0);
      }
    }
    umma_commit(mma_bar);
  }
  __syncwarp();
  mbar_wait(mma_bar, 0);
  tcgen05_fence_after();

  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t regs[32];
    tmem_ld_32x32b_x32(tmem_base + lane_sel + c0, regs);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) p.d[tid * N + c0 + j] = __uint_as_float(regs[j]);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

cudaError_t launch_probe(bool bf16, const CUtensorMap& ta, const CUtensorMap& tb, const ProbeParams& p,
                         cudaStream_t stream) {
  constexpr int smem = 65536 + 64 + 1024;
  auto kernel = bf16 ? fa2_probe_kernel<true> : fa2_probe_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  kernel<<<1, 128, smem, stream>>>(ta, tb, p);
  return cudaGetLastError();
}

}  // namespace fa2

// ------------------------------------------------------------------------------------------------
// C entry point of the TEST-ONLY library libfa2_probe.so (not part of libfa2_sm100.so, not in include/fa2_sm100.h):
//   int fa2_probe_umma(a, b, d, N, K, a_mode, b_mode, dtype /* 16 = fp16, 17 = bf16 */, stream)  -> 0 or an error code
// ------------------------------------------------------------------------------------------------
namespace {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// [rows, cols] row-major 16-bit matrix -> 2-D-in-4-D map (cols, rows, 1, 1), box (64, box_rows), 128-byte swizzle
int probe_tmap(CUtensorMap* map, const void* ptr, bool bf16, int rows, int cols, int box_rows) {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return -7;
  const cuuint64_t gdim[4] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), 1, 1};
  const cuuint64_t row_bytes = static_cast<cuuint64_t>(cols) * 2;
  const cuuint64_t gstride[3] = {row_bytes, row_bytes * rows, row_bytes * rows};
  const cuuint32_t box[4] = {64u, static_cast<cuuint32_t>(box_rows), 1u, 1u};
  const cuuint32_t estride[4] = {1u, 1u, 1u, 1u};
  CUresult r = reinterpret_cast<EncodeTiledFn>(sym)(
      map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), gdim, gstride,
      box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

}  // namespace

extern "C" int fa2_probe_umma(const void* a, const void* b, float* d, int N, int K, int a_mode, int b_mode, int dtype,
                              void* stream) {
  if (dtype != 16 && dtype != 17) return -2;
  if ((N != 64 && N != 128) || (K != 64 && K != 128)) return -3;
  if (a_mode < 0 || a_mode > 4 || b_mode < 0 || b_mode > 1) return -3;
  if (!a || !b || !d) return -5;
  const bool bf16 = dtype == 17;
  CUtensorMap ta, tb;
  if (int e = probe_tmap(&ta, a, bf16, 128, K, 128)) return e;            // A as [128, K]
  if (b_mode == 0) {
    if (int e = probe_tmap(&tb, b, bf16, N, K, N)) return e;              // B^T as [N, K]
  } else {
    if (int e = probe_tmap(&tb, b, bf16, K, N, K)) return e;              // B as [K, N]
  }
  fa2::ProbeParams p{};
  p.a = a; p.b = b; p.d = d; p.N = N; p.K = K; p.a_mode = a_mode; p.b_mode = b_mode;
  return static_cast<int>(fa2::launch_probe(bf16, ta, tb, p, static_cast<cudaStream_t>(stream)));
}

// Host-side probe of the item order of the causal grids (fa2_kernels.cuh: tail_balanced_item): item `lin` of a grid with
// nx blocks per head, ny heads and nbh (batch, head) pairs -> out = {x, head, batch}. tests/test_item_order.py checks
// that the order is a permutation of the grid for every tail length.
extern "C" void fa2_probe_tail_item(int lin, int nx, int ny, int nbh, int tail_heads, int* out) {
  const fa2::ItemCoord c = fa2::tail_balanced_item(lin, nx, ny, nbh, tail_heads);
  out[0] = c.x;
  out[1] = c.head;
  out[2] = c.batch;
}
extern "C" void fa2_probe_tail_block(int x, int y, int z, int nx, int ny, int nz, int tail_log2, int* out) {
  const fa2::ItemCoord c = fa2::tail_balanced_block(x, y, z, nx, ny, nz, tail_log2);
  out[0] = c.x;
  out[1] = c.head;
  out[2] = c.batch;
}
extern "C" int fa2_probe_tail_log2_for(int tail_heads, int heads) { return fa2::tail_log2_for(tail_heads, heads); }
extern "C" int fa2_probe_tail_heads_for(long long streamed_bytes_per_head, int causal) {
  return fa2::tail_heads_for(static_cast<size_t>(streamed_bytes_per_head), causal);
}
