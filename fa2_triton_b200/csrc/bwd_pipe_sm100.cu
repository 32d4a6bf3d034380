// Pipelined deterministic backward for sm_100a (B200): dK, dV and dQ in ONE persistent kernel.
//
// The deterministic backward forms dK / dV per key block (bwd_dkdv_sm100.cu, 4 products per block pair) and spills each
// 16-bit dS tile; dQ is then one product per tile, summed per query block in a fixed order (bwd_dq_spill_sm100.cu). Run as
// two kernels one after the other, every tile travels to HBM and back (config 3: 2.4 GB written, 2.35 GB read - 5x the
// backward's algorithmic traffic, and the dQ pass is HBM-bound). Running the two kernels NEXT TO each other on a static
// split of the SMs was measured twice and lost both times (profiles/r02_experiments.txt, items 1 and 8): CTAs that wait
// for tiles hold SMs the tensor-bound kernel needs.
//
// Here every SM runs one persistent CTA that switches roles item by item:
//   * dQ items (query block, head, batch) are taken FIRST, but only when READY - all dS tiles of the query block have
//     reached global memory (one counter per query block, incremented by the producing CTA after the tile's bulk store
//     has completed). The tiles are then still in L2; once a tile sits in shared memory its lines are discarded from L2
//     (discard.global.L2), so most tiles never travel to HBM in either direction.
//   * otherwise the CTA takes the next dK/dV item (key block, kv head, batch), in the stand-alone kernel's order;
//   * when no dK/dV item is left it waits for the remaining dQ items to become ready.
// No CTA ever waits for data while other work exists, and a dQ item only ever waits for dK/dV items that are already
// running or done - no deadlock whatever the hardware's CTA placement.
//
// The role bodies are the stand-alone kernels' bodies (bwd_dkdv_body.cuh, bwd_dq_spill_body.cuh): same arithmetic, same
// summation order, bit-identical results (tests/test_backward_gpu.py). The CTA has the dK/dV kernel's shape (4 compute
// warpgroups + 1 issue warpgroup, register split by setmaxnreg, all of tensor memory); a chain of dQ items uses warps
// 0-3 (epilogue) and the issue warpgroup (MMA issuer, TMA producer + claims, two L2-discard warps).
#include <type_traits>

#include "bwd_dkdv_body.cuh"
#include "bwd_dq_spill_body.cuh"

namespace fa2 {

namespace {

template <int D>
struct PipeSmem {
  using Lk = dkdv::DkdvSmem<D, dkdv::q_stages(D, true, 0), false, true>;
  using Lq = dqs::DqSpillSmem<D>;
  static constexpr int kItemOff = (Lk::kBytes > Lq::kBytes ? Lk::kBytes : Lq::kBytes);  // 8 ints (see the kernel)
  static constexpr int kTmemSlotOff = kItemOff + 32;
  static constexpr int kBytes = kTmemSlotOff + 16;
  static_assert(kBytes <= 232448, "pipelined backward smem budget exceeded");
};

}  // namespace

template <int D, bool kBf16>
__global__ void __launch_bounds__(dkdv::kThreads, 1)
fa2_bwd_pipe_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                    const __grid_constant__ CUtensorMap tmap_dk, const __grid_constant__ CUtensorMap tmap_dv,
                    const DkdvParams pk, const DqSpillParams pq, const PipeParams pp) {
  static_assert(dkdv::kThreads == dqs::Roles<true>::kThreads && dkdv::kMmaWarp == dqs::Roles<true>::kMmaWarp &&
                    dkdv::kTmaWarp == dqs::Roles<true>::kTmaWarp,
                "the dQ role is laid out over the dK/dV kernel's warps");
  using L = PipeSmem<D>;
  extern __shared__ __align__(1024) uint8_t smem[];
  volatile int* item = reinterpret_cast<volatile int*>(smem + L::kItemOff);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlotOff);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == dkdv::kMmaWarp) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (*tmem_slot != 0u) __trap();  // whole-TMEM allocation => base 0 (keeps tcgen05 addresses warp-uniform)
  // PDL: everything above overlapped the tail of the previous kernel of the stream (delta)
  pdl_wait();
  if (threadIdx.x == 0) pdl_launch_dependents();

  // The warps split FIRST and each half runs its own copy of the item loop, so that ptxas compiles the role bodies' halves
  // for their own register budgets (a setmaxnreg hoisted above a common loop made it compile everything for the smaller
  // one). The halves meet at the bodies' __syncthreads (barrier 0 counts arrivals, whichever instruction they come from).
  // item: [0] kind and [1] index of the current item, [2] / [3] of the next one (chosen by the producer thread of the
  // current item once it has issued its last load), [4] "no dK/dV item left".
  auto decide = [&] { dqs::pipe_decide(pp, pq.Sq, pq.Sk, pq.causal, item); };
  auto run_item = [&](auto role) {
    constexpr int kRole = decltype(role)::value;
    const int kind = item[0], idx = item[1];
    if (kind == dqs::kItemNone) return false;
    if (kind == dqs::kItemDkdv) {
      // same order as the stand-alone kernel's grid: key block fastest (0 = most query blocks first), then kv head, batch
      const int nblk = idx % pp.n_k_blocks;
      const int rest = idx / pp.n_k_blocks;
      dkdv::dkdv_cta<D, kBf16, true, 0, kRole>(smem, tmap_q, tmap_k, tmap_v, tmap_do, tmap_dk, tmap_dv, tmap_q, pk, nblk,
                                                rest % pk.Hkv, rest / pk.Hkv, decide);
    } else {
      dqs::dq_spill_cta<D, kBf16, kRole>(smem, tmap_k, pq, pp, idx, decide);
    }
    return true;  // (both bodies end with a __syncthreads: `item` may be rewritten)
  };

  if (warp >= dkdv::kMmaWarp) {
    reg_dealloc<dkdv::kRegsIssue>();
    if (warp == dkdv::kTmaWarp && lane == 0) {
      item[4] = 0;
      item[2] = dqs::kItemUndecided;
    }
    for (;;) {
      if (warp == dkdv::kTmaWarp && lane == 0) {
        if (item[2] == dqs::kItemUndecided) decide();  // the first item; an item whose producer thread had nothing to do
        item[0] = item[2];
        item[1] = item[3];
        item[2] = dqs::kItemUndecided;
      }
      __syncthreads();
      if (!run_item(std::integral_constant<int, 1>{})) break;
    }
  } else {
    reg_alloc<dkdv::kRegsCmp>();
    for (;;) {
      __syncthreads();
      if (!run_item(std::integral_constant<int, 2>{})) break;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == dkdv::kMmaWarp) {
    tcgen05_fence_after();
    tmem_dealloc<512>(0u);
  }
}

template <int D, bool kBf16>
static cudaError_t launch_bwd_pipe_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                                     const CUtensorMap& tdo, const CUtensorMap& tdk, const CUtensorMap& tdv,
                                     const DkdvParams& k, const DqSpillParams& d, const PipeParams& pp,
                                     cudaStream_t stream) {
  auto kernel = fa2_bwd_pipe_kernel<D, kBf16>;
  constexpr int smem = PipeSmem<D>::kBytes;
  if (cudaError_t e = ensure_dynamic_smem<fa2_bwd_pipe_kernel<D, kBf16>>(smem); e != cudaSuccess) return e;
  const int n_items = pp.n_dkdv + pp.n_heads * pp.n_q_blocks;
  const int sms = current_sm_count();
  dim3 grid(n_items < sms ? n_items : sms);  // persistent: one CTA per SM
  return launch_pdl(kernel, grid, dim3(dkdv::kThreads), smem, stream, tq, tk, tv, tdo, tdk, tdv, k, d, pp);
}

cudaError_t launch_bwd_pipe(int D, bool bf16, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                            const CUtensorMap& tdo, const CUtensorMap& tdk, const CUtensorMap& tdv, const DkdvParams& k,
                            const DqSpillParams& d, const PipeParams& pp, cudaStream_t stream) {
  if (k.seqlens != nullptr || k.bias.ptr != nullptr || k.kv_split > 1 || k.done == nullptr || k.ds_spill == nullptr)
    return cudaErrorInvalidValue;
  if (D == 128) return bf16 ? launch_bwd_pipe_t<128, true>(tq, tk, tv, tdo, tdk, tdv, k, d, pp, stream)
                            : launch_bwd_pipe_t<128, false>(tq, tk, tv, tdo, tdk, tdv, k, d, pp, stream);
  return cudaErrorInvalidValue;
}

}  // namespace fa2
