// dK / dV kernel of the deterministic backward for sm_100a (B200), optionally spilling dS for the dQ kernel.
//
// Replaces the "pid < NUM_BLOCKS_KV" half of the reference's `_bwd_kernel` (src/backward/kernel.py:65-153):
// `_compute_column_blocks_dkdv` / `_compute_single_block_dkdv` (src/backward/compute_dkdv.py:116-296, 8-112).
// One CTA owns a 128-row K/V block and walks the query blocks that see it.
//
// Arithmetic: S^T = K Q^T, P^T = exp2(S^T c - LSE2), dV += P^T dO with P rounded to the input
// dtype, dP^T = V dO^T, dS^T = P^T o (dP^T - delta) rounded to the input dtype, dK += dS^T Q (scaled by softmax_scale
// in the epilogue).
//
// Shared memory bandwidth is the scarce resource of this kernel, not the tensor pipe: a 128 x 128 x 16 product whose
// two operands come from shared memory reads 8 KB in its 64 cycles - all of the SM's 128 B/clk - so every operand that
// can live in tensor memory should (measured: the version of this kernel that staged P^T and dS^T through shared
// memory, 416 KB of shared-memory traffic per block pair, ran at 1.12-1.28 ms; the one that kept P^T in tensor memory
// and only dS^T in shared memory, 320 KB, at 0.94-0.97 ms). Here
//   * P^T (16-bit) overwrites S^T and dS^T (16-bit) overwrites dP^T in tensor memory, each thread in its own columns,
//     and dV / dK take them as the A operand from there: 256 KB of shared-memory traffic per block pair
//     (S^T, dP^T: 64 KB each; dV, dK: 32 KB each; the Q and dO tiles arriving by TMA: 64 KB);
//   * dO is double buffered like Q (the shared memory a dS^T tile would need pays for it);
//   * kSpill: every thread also writes its slice of dS^T to global memory straight from registers, in a layout made
//     for that: [16 groups of 8 queries][128 keys][16 B], so that a warp's store instruction covers 512 contiguous
//     bytes. That is the no-swizzle MN-major operand layout of tcgen05.mma, so bwd_dq_spill_sm100.cu bulk-copies the
//     32 KB tile into shared memory as it is and forms dQ with ONE product per block pair in a fixed order - the
//     deterministic backward then costs 5 products per pair like the atomic one, not 7.
//
// CTA = kCmpWgs compute warpgroups + one issue warpgroup:
//   compute : thread = key row (TMEM lane), warpgroup g owns query columns [g * 128 / kCmpWgs, ...)
//   issue   : warp 0 MMA issuer, warp 1 TMA producer, warps 2-3 idle
// TMEM (512 columns): S^T / P^T [0,128) | dP^T / dS^T [128,256) | dV [256,256+D) | dK [384,384+D)
// Tensor-pipe order per query block i:  dV(i) | S^T(i+1) | dK(i) | dP^T(i+1)
// Head dim 64 (dV and dK are 64 columns each): S^T is DOUBLE-BUFFERED - S^T / P^T [0,128) and [128,256) | dP^T / dS^T
// [256,384) | dV [384,448) | dK [448,512), order  dV(i) | S^T(i+2) | dK(i) | dP^T(i+1)  - so that the compute warps find the
// scores of block i + 1 waiting when they are done with block i. At this head dim the four products of a block pair take
// 1024 tensor cycles against ~1300 issue cycles of exponentials and dS arithmetic on every scheduler: the compute warps
// are the bottleneck, and with one S^T buffer they also sat out the round trip P^T(i) -> dV(i), S^T(i+1) in every
// iteration (2350 cycles per block pair at S = 16384).
#include "fa2_kernels.cuh"
#include "ptx.cuh"

namespace fa2 {

#ifdef FA2_TRACE
__device__ long long g_dkdv_trace[4 * 64 * 8];
#define FA2_STAMP(role, it, ev)                                                                          \
  do {                                                                                                   \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (it) < 64)                               \
      g_dkdv_trace[((role) * 64 + (it)) * 8 + (ev)] = clock64();                                         \
  } while (0)
cudaError_t fetch_dkdv_trace(long long* host, size_t n) {
  return cudaMemcpyFromSymbol(host, g_dkdv_trace, n * sizeof(long long));
}
#else
#define FA2_STAMP(role, it, ev) do { } while (0)
#endif

}  // namespace fa2

#include "bwd_dkdv_body.cuh"

namespace fa2 {

using namespace dkdv;

// kExtra: 0 = the plain kernel; 1 = attention bias; 2 = variable-length batches; 3 = split head groups (bwd_dkdv_body.cuh)
template <int D, bool kBf16, bool kSpill, int kExtra>
__global__ void __launch_bounds__(kThreads, 1)
fa2_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                    const __grid_constant__ CUtensorMap tmap_dk, const __grid_constant__ CUtensorMap tmap_dv,
                    const __grid_constant__ CUtensorMap tmap_bias, const DkdvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // causal: low key blocks are seen by the most query blocks - longest first
  dkdv_cta<D, kBf16, kSpill, kExtra, 0>(smem, tmap_q, tmap_k, tmap_v, tmap_do, tmap_dk, tmap_dv, tmap_bias, p,
                                            static_cast<int>(blockIdx.x), static_cast<int>(blockIdx.y),
                                            static_cast<int>(blockIdx.z), [] {});
}

template <int D, bool kBf16, bool kSpill, int kExtra>
static cudaError_t launch_bwd_dkdv_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                                     const CUtensorMap& tdo, const CUtensorMap& tdk, const CUtensorMap& tdv,
                                     const DkdvParams& p, cudaStream_t stream) {
  auto kernel = fa2_bwd_dkdv_kernel<D, kBf16, kSpill, kExtra>;
  constexpr int smem = DkdvSmem<D, q_stages(D, false, kExtra), kExtra == 1, false>::kBytes;
  const CUtensorMap& tbias = p.bias_tma ? *p.bias_map : tq;  // placeholder when the bias is read through plain loads
  if (cudaError_t e = ensure_dynamic_smem<fa2_bwd_dkdv_kernel<D, kBf16, kSpill, kExtra>>(smem); e != cudaSuccess) return e;
  dim3 grid((p.Sk + kBlk - 1) / kBlk, p.Hkv * (kExtra == 3 ? p.kv_split : 1), p.B);
  return launch_pdl(kernel, grid, dim3(kThreads), smem, stream, tq, tk, tv, tdo, tdk, tdv, tbias, p);
}

cudaError_t launch_bwd_dkdv(int D, bool bf16, bool spill, const CUtensorMap& tq, const CUtensorMap& tk,
                            const CUtensorMap& tv, const CUtensorMap& tdo, const CUtensorMap& tdk,
                            const CUtensorMap& tdv, const DkdvParams& p, cudaStream_t stream) {
  const int extra = p.seqlens != nullptr ? 2 : (p.bias.ptr != nullptr ? 1 : (p.kv_split > 1 ? 3 : 0));
  if (p.seqlens != nullptr && p.bias.ptr != nullptr) return cudaErrorInvalidValue;
  if (p.kv_split > 1 && extra != 3) return cudaErrorInvalidValue;
#define FA2_DKDV_CASE(DD, BF, SP)                                                                                     \
  if (D == DD && bf16 == BF && spill == SP) {                                                                         \
    if (extra == 0) return launch_bwd_dkdv_t<DD, BF, SP, 0>(tq, tk, tv, tdo, tdk, tdv, p, stream);                    \
    if (extra == 1) return launch_bwd_dkdv_t<DD, BF, SP, 1>(tq, tk, tv, tdo, tdk, tdv, p, stream);                    \
    if (extra == 2) return launch_bwd_dkdv_t<DD, BF, SP, 2>(tq, tk, tv, tdo, tdk, tdv, p, stream);                    \
    return launch_bwd_dkdv_t<DD, BF, SP, 3>(tq, tk, tv, tdo, tdk, tdv, p, stream);                                    \
  }
  FA2_DKDV_CASE(128, true, true) FA2_DKDV_CASE(128, true, false) FA2_DKDV_CASE(128, false, true) FA2_DKDV_CASE(128, false, false)
  FA2_DKDV_CASE(64, true, true) FA2_DKDV_CASE(64, true, false) FA2_DKDV_CASE(64, false, true) FA2_DKDV_CASE(64, false, false)
#undef FA2_DKDV_CASE
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------
// sum of the fp32 dK / dV partials of the split head groups (kExtra == 3): fixed order, one rounding
// ------------------------------------------------------------------------------------------------
template <bool kBf16>
__global__ void __launch_bounds__(256) fa2_dkdv_reduce_kernel(const DkdvReduceParams p) {
  const int64_t n4 = static_cast<int64_t>(p.B) * p.Sk * p.Hkv * (p.D / 4);  // float4 groups per output tensor
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < 2 * n4; i += static_cast<int64_t>(gridDim.x) * 256) {
    const bool is_dv = i >= n4;
    const int64_t e = is_dv ? i - n4 : i;
    const int d4 = static_cast<int>(e % (p.D / 4));
    const int h = static_cast<int>((e / (p.D / 4)) % p.Hkv);
    const int64_t bs = e / (p.D / 4) / p.Hkv;  // b * Sk + s
    const float* src = (is_dv ? p.part_dv : p.part_dk) + ((bs * p.Hkv + h) * p.kv_split) * p.D + d4 * 4;
    float4 acc = *reinterpret_cast<const float4*>(src);
    for (int j = 1; j < p.kv_split; ++j) {
      const float4 x = *reinterpret_cast<const float4*>(src + static_cast<int64_t>(j) * p.D);
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    const int64_t* st = is_dv ? p.dv_strides : p.dk_strides;
    const int b = static_cast<int>(bs / p.Sk), srow = static_cast<int>(bs % p.Sk);
    uint16_t* dst = static_cast<uint16_t*>(is_dv ? p.dv : p.dk) + b * st[0] + h * st[1] + srow * st[2] + d4 * 4;
    *reinterpret_cast<uint2*>(dst) = make_uint2(pack2<kBf16>(acc.x, acc.y), pack2<kBf16>(acc.z, acc.w));
  }
}

cudaError_t launch_dkdv_reduce(bool bf16, const DkdvReduceParams& p, cudaStream_t stream) {
  const int64_t n = 2ll * p.B * p.Sk * p.Hkv * (p.D / 4);
  const int blocks = static_cast<int>(n / 256 + 1 < 148 * 8 ? n / 256 + 1 : 148 * 8);
  if (bf16) fa2_dkdv_reduce_kernel<true><<<blocks, 256, 0, stream>>>(p);
  else fa2_dkdv_reduce_kernel<false><<<blocks, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace fa2
