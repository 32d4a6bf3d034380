// dQ kernel of the deterministic backward for sm_100a (B200), fed by the dS tiles the dK/dV kernel spilled.
//
// Replaces the "pid >= NUM_BLOCKS_KV" half of the reference's `_bwd_kernel` (src/backward/kernel.py:154-182):
// `_compute_row_blocks_dq` (src/backward/compute_dq.py:82-261). The reference recomputes S, P, dP and dS for every
// block pair there (3 products); here dS^T(m, n) - already rounded to the input dtype - comes back from the workspace
// written by bwd_dkdv_sm100.cu (in the no-swizzle MN-major operand layout, so a linear 32 KB bulk copy loads it) and only
// dQ(m) = softmax_scale * sum_n dS(m, n) K(n)  is left: one product per pair, summed in ascending key-block order in a
// tensor-memory accumulator owned by one CTA - deterministic by construction, no atomics, no cross-CTA traffic.
//
// HBM-bound: every 32 KB dS tile is read exactly once (config 3: 2.2 GB); K comes from L2 (items are dealt head-major,
// longest first, by an atomicAdd scheduler, so the CTAs running at any moment share a handful of heads).
//
// Persistent, one CTA per SM, 6 warps:
//   warps 0-3 : epilogue (thread = query row = TMEM lane): accumulator -> registers -> caller's dQ
//   warp 4    : MMA issuer        warp 5 : TMA producer + scheduler
// TMEM: two accumulators [0, D) and [256, 256 + D): the epilogue of item t overlaps the products of item t + 1.
// The CTA's body lives in bwd_dq_spill_body.cuh (the pipelined backward, bwd_pipe_sm100.cu, runs it too).
#include "bwd_dq_spill_body.cuh"

namespace fa2 {

using namespace dqs;

template <int D, bool kBf16>
__global__ void __launch_bounds__(Roles<false>::kThreads, 1)
fa2_bwd_dq_spill_kernel(const __grid_constant__ CUtensorMap tmap_k, const DqSpillParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  dq_spill_cta<D, kBf16, 0>(smem, tmap_k, p, PipeParams{}, 0, [] {});
}

template <int D, bool kBf16>
static cudaError_t launch_bwd_dq_spill_t(const CUtensorMap& tk, const DqSpillParams& p, cudaStream_t stream) {
  auto kernel = fa2_bwd_dq_spill_kernel<D, kBf16>;
  constexpr int smem = DqSpillSmem<D>::kBytes;
  if (cudaError_t e = ensure_dynamic_smem<fa2_bwd_dq_spill_kernel<D, kBf16>>(smem); e != cudaSuccess) return e;
  const int n_items = ((p.Sq + kBlk - 1) / kBlk) * p.Hq * p.B;
  const int sms = current_sm_count();
  dim3 grid(n_items < sms ? n_items : sms);  // persistent: one CTA per SM
  return launch_pdl(kernel, grid, dim3(Roles<false>::kThreads), smem, stream, tk, p);
}

cudaError_t launch_bwd_dq_spill(int D, bool bf16, const CUtensorMap& tk, const DqSpillParams& p, cudaStream_t stream) {
  if (D == 128) return bf16 ? launch_bwd_dq_spill_t<128, true>(tk, p, stream) : launch_bwd_dq_spill_t<128, false>(tk, p, stream);
  if (D == 64) return bf16 ? launch_bwd_dq_spill_t<64, true>(tk, p, stream) : launch_bwd_dq_spill_t<64, false>(tk, p, stream);
  return cudaErrorInvalidValue;
}

}  // namespace fa2
