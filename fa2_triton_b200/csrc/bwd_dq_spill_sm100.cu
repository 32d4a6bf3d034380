// dQ kernel of the deterministic backward for sm_100a (B200), fed by the dS tiles the dK/dV kernel spilled.
//
// Replaces the "pid >= NUM_BLOCKS_KV" half of the reference's `_bwd_kernel` (src/backward/kernel.py:154-182):
// `_compute_row_blocks_dq` (src/backward/compute_dq.py:82-261). The reference recomputes S, P, dP and dS for every
// block pair there (3 products); here dS^T(m, n) - already rounded to the input dtype - comes back from the workspace
// written by bwd_dkdv_sm100.cu (in the no-swizzle MN-major operand layout, so a linear 32 KB bulk copy loads it) and only  dQ(m) = softmax_scale * sum_n dS(m, n) K(n)  is left: one product per pair,
// summed in ascending key-block order in a tensor-memory accumulator owned by one CTA - deterministic by construction,
// no atomics, no cross-CTA traffic.
//
// HBM-bound: every 32 KB dS tile is read exactly once (config 3: 2.2 GB); K comes from L2 (items are dealt head-major,
// longest first, by an atomicAdd scheduler, so the CTAs running at any moment share a handful of heads).
//
// Persistent, one CTA per SM, 6 warps:
//   warps 0-3 : epilogue (thread = query row = TMEM lane): accumulator -> registers -> caller's dQ
//   warp 4    : MMA issuer        warp 5 : TMA producer + scheduler
// TMEM: two accumulators [0, D) and [256, 256 + D): the epilogue of item t overlaps the products of item t + 1.
#include "fa2_kernels.cuh"
#include "ptx.cuh"

namespace fa2 {

namespace {

constexpr int kBlk = 128;
constexpr int kThreads = 192;
constexpr int kMmaWarp = 4, kTmaWarp = 5;
constexpr int kSchedSlots = 4;

template <int D>
struct DqSpillSmem {
  static constexpr int kStages = (D == 128) ? 3 : 4;
  static constexpr int kKTileBytes = kBlk * D * 2;
  static constexpr int kStageBytes = kDsTileBytes + kKTileBytes;
  static constexpr int kBarOff = kStages * kStageBytes;
  // full[S], empty[S], acc_full[2], acc_empty[2], item_full[4], item_empty[4]
  static constexpr int kNumBars = 2 * kStages + 4 + 2 * kSchedSlots;
  static constexpr int kSchedOff = kBarOff + kNumBars * 8;
  static constexpr int kTmemSlotOff = kSchedOff + kSchedSlots * 4;
  static constexpr int kBytes = kTmemSlotOff + 16;
  static_assert(kBytes <= 232448, "dQ spill kernel smem budget exceeded");
};

struct Item {
  int m, head, batch;
};
// work item w -> (query block, head, batch): query blocks of one (batch, head) are consecutive, longest first; the last
// heads are dealt block-major (fa2_kernels.cuh: tail_balanced_item)
__device__ __forceinline__ Item item_of(int w, int n_q_blocks, int Hq, int B, int tail_heads) {
  const ItemCoord c = tail_balanced_item(w, n_q_blocks, Hq, Hq * B, tail_heads);
  Item it;
  it.m = n_q_blocks - 1 - c.x;
  it.head = c.head;
  it.batch = c.batch;
  return it;
}

}  // namespace

template <int D, bool kBf16>
__global__ void __launch_bounds__(kThreads, 1)
fa2_bwd_dq_spill_kernel(const __grid_constant__ CUtensorMap tmap_k, const DqSpillParams p) {
  using L = DqSpillSmem<D>;
  constexpr int kStages = L::kStages;
  constexpr int kSub = D / 64;
  constexpr int kSubBytes = kBlk * 128;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* full = bars;
  uint64_t* empty = full + kStages;
  uint64_t* acc_full = empty + kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* item_full = acc_empty + 2;
  uint64_t* item_empty = item_full + kSchedSlots;
  volatile int* sched = reinterpret_cast<volatile int*>(smem + L::kSchedOff);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlotOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_q_blocks = (p.Sq + kBlk - 1) / kBlk;
  const int n_items = n_q_blocks * p.Hq * p.B;
  const int group = p.Hq / p.Hkv;

  if ((smem_u32(smem) & 1023u) != 0) __trap();

  if (warp == kTmaWarp && lane == 0) tma_prefetch_desc(&tmap_k);
  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&acc_full[b], 1);
        mbar_init(&acc_empty[b], 128);
      }
      for (int i = 0; i < kSchedSlots; ++i) {
        mbar_init(&item_full[i], 1);
        mbar_init(&item_empty[i], 1 + 128);  // MMA thread + every epilogue thread
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (*tmem_slot != 0u) __trap();  // whole-TMEM allocation => base 0 (keeps tcgen05 addresses warp-uniform)
  // PDL: the prologue above overlapped the tail of the dK/dV kernel; its dS tiles are complete only after this wait
  pdl_wait();
  if (threadIdx.x == 0) pdl_launch_dependents();

  if (warp == kTmaWarp) {
    // ===================================== TMA producer + scheduler ================================
    if (elect_one()) {
      const uint64_t policy = l2_policy_evict_first();  // dS is read once
      const uint8_t* base = static_cast<const uint8_t*>(p.ds_spill);
      uint32_t c = 0;  // stages filled so far
      for (uint32_t t = 0;; ++t) {
        const uint32_t slot = t % kSchedSlots;
        mbar_wait(&item_empty[slot], ((t / kSchedSlots) & 1) ^ 1);
        const int w = atomicAdd(p.sched, 1);
        // every CTA draws exactly one id past the end; the very last draw resets the counter for the next launch
        if (w == n_items + static_cast<int>(gridDim.x) - 1) atomicExch(p.sched, 0);
        sched[slot] = (w < n_items) ? w : -1;
        mbar_arrive(&item_full[slot]);
        if (w >= n_items) break;
        const Item it = item_of(w, n_q_blocks, p.Hq, p.B, p.tail_heads);
        const int head_kv = it.head / group;
        const int len = p.seqlens != nullptr ? min(max(p.seqlens[it.batch], 0), p.Sq) : 0;  // variable-length mode: this sequence's length
        const int nblk = p.seqlens != nullptr ? visible_kv_blocks_of(it.m, len, len, p.causal)
                                              : visible_kv_blocks_of(it.m, p.Sq, p.Sk, p.causal);
        const long long tile0 = (static_cast<long long>(it.batch) * p.Hq + it.head) * p.tiles_per_head +
                                ds_tile_prefix(it.m, p.Sq, p.Sk, p.causal);
        // variable-length mode: the K map of this batch element (bounded by its length; batch coordinate 0)
        const CUtensorMap* mk = p.seqlens != nullptr ? p.vmaps + static_cast<size_t>(it.batch) * kVarlenMapsPerBatch + 1 : &tmap_k;
        const int bc = p.seqlens != nullptr ? 0 : it.batch;
        if (p.seqlens != nullptr) tmap_acquire(mk);
        for (int i = 0; i < nblk; ++i, ++c) {
          const uint32_t s = c % kStages;
          uint8_t* stage = smem + s * L::kStageBytes;
          mbar_wait(&empty[s], ((c / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&full[s], L::kStageBytes);
          bulk_load_1d_hint(stage, base + (tile0 + i) * kDsTileBytes, kDsTileBytes, &full[s], policy);
          for (int j = 0; j < kSub; ++j)
            tma_load_4d(stage + kDsTileBytes + j * kSubBytes, mk, &full[s], j * 64, i * kBlk, head_kv, bc);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ========================================== MMA issuer =========================================
    if (elect_one()) {
      // dQ[q, d] (+)= sum_kv dS^T[kv, q] K[kv, d]: both operands MN-major (the rows of either tile are the reduction index)
      const uint32_t idesc = make_idesc_f16(kBlk, (p.d_actual + 15) / 16 * 16, kBf16, true, true);  // N = the tensors' head dim (x16)
      uint32_t c = 0;
      for (uint32_t t = 0;; ++t) {
        const uint32_t slot = t % kSchedSlots;
        mbar_wait(&item_full[slot], (t / kSchedSlots) & 1);
        const int w = sched[slot];
        mbar_arrive(&item_empty[slot]);
        if (w < 0) break;
        const Item it = item_of(w, n_q_blocks, p.Hq, p.B, p.tail_heads);
        const int len = p.seqlens != nullptr ? min(max(p.seqlens[it.batch], 0), p.Sq) : 0;  // variable-length mode: this sequence's length
        const int nblk = p.seqlens != nullptr ? visible_kv_blocks_of(it.m, len, len, p.causal)
                                              : visible_kv_blocks_of(it.m, p.Sq, p.Sk, p.causal);
        const uint32_t buf = t & 1;
        const uint32_t acc = buf * 256;
        mbar_wait(&acc_empty[buf], ((t >> 1) & 1) ^ 1);  // the epilogue of item t - 2 has read this accumulator
        tcgen05_fence_after();
        for (int i = 0; i < nblk; ++i, ++c) {
          const uint32_t s = c % kStages;
          mbar_wait(&full[s], (c / kStages) & 1);
          tcgen05_fence_after();
          const uint32_t stage = smem_u32(smem + s * L::kStageBytes);
          // dS^T tile as spilled: [16 pieces of 8 queries][128 keys][16 B] = MN-major without swizzle; 8-key groups are
          // 128 B apart (LBO), 8-query groups 2 KB (SBO); one k-step = 16 keys = 256 B
          const SmemDesc ds_mn = make_smem_desc_noswizzle(stage, 128, kBlk * 16);
          const SmemDesc k_mn = make_smem_desc_sw128(stage + kDsTileBytes, kSubBytes, 1024);
#pragma unroll
          for (int k = 0; k < kBlk / 16; ++k)
            umma_ss(acc, advance(ds_mn, k * 256), advance(k_mn, k * 2048), idesc, (i > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ========================================== epilogue ===========================================
    const int r = threadIdx.x & 127;  // query row inside the block == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float mul = p.softmax_scale;
    for (uint32_t t = 0;; ++t) {
      const uint32_t slot = t % kSchedSlots;
      mbar_wait(&item_full[slot], (t / kSchedSlots) & 1);
      const int w = sched[slot];
      mbar_arrive(&item_empty[slot]);
      if (w < 0) break;
      const Item it = item_of(w, n_q_blocks, p.Hq, p.B, p.tail_heads);
      const int row = it.m * kBlk + r;
      const int len = p.seqlens != nullptr ? min(max(p.seqlens[it.batch], 0), p.Sq) : 0;  // variable-length mode: this sequence's length
        const int nblk = p.seqlens != nullptr ? visible_kv_blocks_of(it.m, len, len, p.causal)
                                              : visible_kv_blocks_of(it.m, p.Sq, p.Sk, p.causal);
      const uint32_t buf = t & 1;
      mbar_wait(&acc_full[buf], (t >> 1) & 1);
      tcgen05_fence_after();
      const int64_t off0 = it.batch * p.dq_strides[0] + it.head * p.dq_strides[1] + static_cast<int64_t>(row) * p.dq_strides[2];
#pragma unroll
      for (int chunk = 0; chunk < D / 32; ++chunk) {
        uint32_t o[32];
        if (nblk > 0) {
          tmem_ld_32x32b_x32(lane_sel + buf * 256 + chunk * 32, o);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = 0u;  // fully masked query rows: dQ = 0 (src/backward/compute_dq.py:248-249)
        }
        if (chunk == D / 32 - 1) {
          tcgen05_fence_before();
          mbar_arrive(&acc_empty[buf]);
        }
        if (row < (p.seqlens != nullptr ? len : p.Sq)) {
          const int64_t off = off0 + chunk * 32;
          if (p.dq_is_16bit) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.dq) + off);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              if (chunk * 32 + q4 * 8 >= p.d_actual) break;  // zero-padded head dim: nothing to store there
              uint4 v;
              v.x = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 0]) * mul, __uint_as_float(o[q4 * 8 + 1]) * mul);
              v.y = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 2]) * mul, __uint_as_float(o[q4 * 8 + 3]) * mul);
              v.z = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 4]) * mul, __uint_as_float(o[q4 * 8 + 5]) * mul);
              v.w = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 6]) * mul, __uint_as_float(o[q4 * 8 + 7]) * mul);
              dst[q4] = v;
            }
          } else {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.dq) + off);
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
              if (chunk * 32 + q4 * 4 >= p.d_actual) break;
              dst[q4] = make_float4(__uint_as_float(o[q4 * 4 + 0]) * mul, __uint_as_float(o[q4 * 4 + 1]) * mul,
                                    __uint_as_float(o[q4 * 4 + 2]) * mul, __uint_as_float(o[q4 * 4 + 3]) * mul);
            }
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tcgen05_fence_after();
    tmem_dealloc<512>(0u);
  }
}

template <int D, bool kBf16>
static cudaError_t launch_bwd_dq_spill_t(const CUtensorMap& tk, const DqSpillParams& p, cudaStream_t stream) {
  auto kernel = fa2_bwd_dq_spill_kernel<D, kBf16>;
  constexpr int smem = DqSpillSmem<D>::kBytes;
  if (cudaError_t e = ensure_dynamic_smem<fa2_bwd_dq_spill_kernel<D, kBf16>>(smem); e != cudaSuccess) return e;
  const int n_items = ((p.Sq + kBlk - 1) / kBlk) * p.Hq * p.B;
  const int sms = current_sm_count();
  dim3 grid(n_items < sms ? n_items : sms);  // persistent: one CTA per SM
  DqSpillParams pp = p;
  pp.tail_heads = tail_heads_for(static_cast<size_t>(p.Sk) * D * 2, p.causal);  // a head's items stream its K
  return launch_pdl(kernel, grid, dim3(kThreads), smem, stream, tk, pp);
}

cudaError_t launch_bwd_dq_spill(int D, bool bf16, const CUtensorMap& tk, const DqSpillParams& p, cudaStream_t stream) {
  if (D == 128) return bf16 ? launch_bwd_dq_spill_t<128, true>(tk, p, stream) : launch_bwd_dq_spill_t<128, false>(tk, p, stream);
  if (D == 64) return bf16 ? launch_bwd_dq_spill_t<64, true>(tk, p, stream) : launch_bwd_dq_spill_t<64, false>(tk, p, stream);
  return cudaErrorInvalidValue;
}

}  // namespace fa2
