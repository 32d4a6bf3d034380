// Body of the dQ-from-spilled-dS kernel (see bwd_dq_spill_sm100.cu for the design notes): shared by the stand-alone
// persistent kernel (bwd_dq_spill_sm100.cu) and the pipelined backward (bwd_pipe_sm100.cu), whose persistent CTAs switch
// between dK/dV items and chains of dQ items.
#pragma once
#include "fa2_kernels.cuh"
#include "ptx.cuh"

namespace fa2 {
namespace dqs {

constexpr int kBlk = 128;
constexpr int kSchedSlots = 4;
constexpr int kDiscardWarps = 2;

// Warp roles. Stand-alone kernel: 6 warps (epilogue 0-3, MMA 4, TMA 5). Pipelined backward: the CTA has the dK/dV
// kernel's 20 warps and register split (bwd_dkdv_body.cuh): the epilogue stays with warps 0-3 (compute registers), the
// MMA issuer, the TMA producer and two L2-discard warps are the four warps of the issue warpgroup; warps 4-15 sit out.
template <bool kPipe>
struct Roles {
  static constexpr int kMmaWarp = kPipe ? 16 : 4;
  static constexpr int kTmaWarp = kPipe ? 17 : 5;
  static constexpr int kDiscardWarp0 = kPipe ? 18 : 6;  // (no discard warps in the stand-alone kernel)
  static constexpr int kThreads = kPipe ? 640 : 192;
};

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
// work item w -> (query block, head, batch): query blocks of one (batch, head) are consecutive; the stand-alone kernel
// deals them longest first, the pipelined backward in the order the dK/dV items complete them (ascending)
__device__ __forceinline__ Item item_of(int w, int n_q_blocks, int Hq, bool ascending) {
  const int bh = w / n_q_blocks;
  Item it;
  it.m = ascending ? (w - bh * n_q_blocks) : n_q_blocks - 1 - (w - bh * n_q_blocks);
  it.head = bh % Hq;
  it.batch = bh / Hq;
  return it;
}

__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// acquire side of the hand-over of a claimed item's dS tiles: the counters are polled with relaxed loads (an acquire load
// per polled word cost ~5 us each next to the TMA traffic of a busy SM - measured: 40 us per scan of 8 heads), ONE fence
// follows a successful claim, and the bulk copies of the tiles (async proxy) are ordered behind it
__device__ __forceinline__ void pipe_acquire_claimed() {
#ifndef FA2_PIPE_NO_ACQ_FENCE
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
  asm volatile("fence.proxy.async;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// Work distribution of the pipelined backward (one thread per CTA calls these).
//
// dQ items: query block m of (batch, head) may be formed once all of its dS tiles have reached global memory -
// done[(batch * Hq + head) * n_q_blocks + m] counts them. The blocks of a head become ready in ascending order (the
// CTA of key block 0 is the last to pass each of them), so every head has ONE cursor (next_row[bh]) and blocks are
// claimed by advancing it with a compare-and-swap - only when they are ready: nobody ever holds an SM waiting for data
// while dK/dV items are left. Several heads are in flight at once; the scan looks at a window of kWindow heads starting
// at the oldest unfinished one (lo_hint), oldest first, so that tiles are consumed in the order they were written.
// pipe_try_claim_dq returns the item index (bh * n_q_blocks + m), -1 if nothing is ready now, -2 if every item has been claimed.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kWindow = 8;
// query blocks are claimed in groups of pp.rows_per_claim consecutive blocks of one head (one scan and one role switch per
// group instead of per block); a group is ready when its LAST block is - the blocks of a head become ready in ascending
// order. Returns the index of the group's first block.
__device__ __forceinline__ int pipe_group_end(const PipeParams& pp, int w) {  // one past the last item of w's group
  const int bh = w / pp.n_q_blocks, m = w - bh * pp.n_q_blocks;
  const int e = m - m % pp.rows_per_claim + pp.rows_per_claim;
  return bh * pp.n_q_blocks + (e < pp.n_q_blocks ? e : pp.n_q_blocks);
}
__device__ __forceinline__ int pipe_try_claim_dq(const PipeParams& pp, int Sq, int Sk, int causal) {
  const int BH = pp.n_heads, n_rows = pp.n_q_blocks, G = pp.rows_per_claim;
  const int n_groups = (n_rows + G - 1) / G;
  int lo = ld_relaxed_gpu(pp.lo_hint);
  while (lo < BH) {
    // ONE pass over the window: all cursors, then all counters, then at most one compare-and-swap per head. A lost race
    // is not retried (the winner took that group; re-scanning after every lost race made ~2/3 of all scans futile when
    // every CTA that finishes an item goes for the same newly completed block) - the caller has other work or comes back.
    int grp[kWindow], seen[kWindow], last[kWindow];
#pragma unroll
    for (int i = 0; i < kWindow; ++i) grp[i] = (lo + i < BH) ? ld_relaxed_gpu(pp.next_row + lo + i) : n_groups;
#pragma unroll
    for (int i = 0; i < kWindow; ++i) {
      last[i] = grp[i] * G + G - 1 < n_rows ? grp[i] * G + G - 1 : n_rows - 1;
      seen[i] = (grp[i] < n_groups) ? ld_relaxed_gpu(pp.done + static_cast<long long>(lo + i) * n_rows + last[i]) : 0;
    }
    int first_open = -1;
#pragma unroll
    for (int i = 0; i < kWindow; ++i) {
      if (grp[i] >= n_groups) continue;
      if (first_open < 0) first_open = i;
      if (seen[i] >= visible_kv_blocks_of(last[i], Sq, Sk, causal) &&
          atomicCAS(pp.next_row + lo + i, grp[i], grp[i] + 1) == grp[i]) {
        if (first_open > 0) atomicMax(pp.lo_hint, lo + first_open);
        pipe_acquire_claimed();
        return (lo + i) * n_rows + grp[i] * G;
      }
    }
    if (first_open >= 0) {
      if (first_open > 0) atomicMax(pp.lo_hint, lo + first_open);
      return -1;
    }
    lo = lo + kWindow < BH ? lo + kWindow : BH;  // every head of the window is finished
    atomicMax(pp.lo_hint, lo);
  }
  return -2;
}

// The next item of a CTA of the pipelined backward: a READY dQ item if there is one, else the next dK/dV item, else (no
// dK/dV item left) wait for a dQ item to become ready - its tiles are being produced by dK/dV items that are running.
// Called by ONE thread per CTA, from inside the role bodies once the producer thread of the current item has issued its
// last load - the decision (a few dependent L2 round trips: 2.5-6 us, measured) is then ready when the item ends instead
// of sitting between two items with the whole SM idle. st: shared-memory words [2] next kind, [3] next index, [4]
// "no dK/dV item left".
constexpr int kItemNone = 0, kItemDkdv = 1, kItemDq = 2, kItemUndecided = -1;
__device__ __forceinline__ void pipe_decide(const PipeParams& pp, int Sq, int Sk, int causal, volatile int* st) {
  int kind = kItemNone, idx = 0;
  long long spins = 0;
  for (;;) {
    const int w = pipe_try_claim_dq(pp, Sq, Sk, causal);
    if (w >= 0) {
      kind = kItemDq;
      idx = w;
      break;
    }
    if (st[4] == 0) {
      const int d = atomicAdd(pp.dkdv_next, 1);
      if (d < pp.n_dkdv) {
        kind = kItemDkdv;
        idx = d;
        break;
      }
      st[4] = 1;
      continue;
    }
    if (w == -2) break;  // every item of both kinds has been claimed
    __nanosleep(500);
    if (++spins > (1ll << 24)) __trap();  // watchdog (seconds): a producer died
  }
  st[3] = idx;
  st[2] = kind;
}

// kRole 0    : the stand-alone persistent kernel's CTA (items from the atomicAdd counter p.sched).
// kRole 1 / 2: a chain of dQ items inside the pipelined backward: `first_item` was claimed by the caller, the producer
//   thread claims further READY items while there are any and ends the chain otherwise. Tensor memory, the register split
//   and the grid dependency belong to the caller, which has split the warps already: 1 compiles the issue warpgroup's half
//   (MMA issuer, TMA producer, discard warps), 2 the compute warps' half (epilogue in warps 0-3; see bwd_dkdv_body.cuh for
//   why). Every thread of the CTA enters and leaves through a __syncthreads.
// at_producer_end(): called by the producer thread after its last load has been issued (pipelined backward: pipe_decide).
template <int D, bool kBf16, int kRole, typename AtProducerEnd>
__device__ __forceinline__ void dq_spill_cta(uint8_t* smem, const CUtensorMap& tmap_k, const DqSpillParams& p,
                                             const PipeParams& pp, const int first_item, AtProducerEnd at_producer_end) {
  constexpr bool kPipe = kRole != 0;
  using L = DqSpillSmem<D>;
  using R = Roles<kPipe>;
  constexpr int kStages = L::kStages;
  constexpr int kSub = D / 64;
  constexpr int kSubBytes = kBlk * 128;
  constexpr int kMmaWarp = R::kMmaWarp, kTmaWarp = R::kTmaWarp, kDiscardWarp0 = R::kDiscardWarp0;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* full = bars;
  uint64_t* empty = full + kStages;
  uint64_t* acc_full = empty + kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* item_full = acc_empty + 2;
  uint64_t* item_empty = item_full + kSchedSlots;
  volatile int* sched = reinterpret_cast<volatile int*>(smem + L::kSchedOff);
  [[maybe_unused]] uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlotOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_q_blocks = (p.Sq + kBlk - 1) / kBlk;
  const int n_items = n_q_blocks * p.Hq * p.B;
  const int group = p.Hq / p.Hkv;

  if ((smem_u32(smem) & 1023u) != 0) __trap();

  if (kRole != 2 && warp == kTmaWarp && lane == 0) tma_prefetch_desc(&tmap_k);
  if (kRole != 2 && warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], kPipe ? 1 + kDiscardWarps : 1);  // the products' commit (+ one arrival per discard warp)
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&acc_full[b], 1);
        mbar_init(&acc_empty[b], 128);
      }
      for (int i = 0; i < kSchedSlots; ++i) {
        mbar_init(&item_full[i], 1);
        mbar_init(&item_empty[i], 1 + 128 + (kPipe ? kDiscardWarps : 0));  // MMA thread + every epilogue thread (+ discard warps)
      }
      fence_barrier_init();
    }
    __syncwarp();
    if constexpr (!kPipe) tmem_alloc<512>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if constexpr (!kPipe) {
    if (*tmem_slot != 0u) __trap();  // whole-TMEM allocation => base 0 (keeps tcgen05 addresses warp-uniform)
    // PDL: the prologue above overlapped the tail of the dK/dV kernel; its dS tiles are complete only after this wait
    pdl_wait();
    if (threadIdx.x == 0) pdl_launch_dependents();
  }

  if (kRole != 2 && warp == kTmaWarp) {
    // ===================================== TMA producer + scheduler ================================
    if (elect_one()) {
      // stand-alone: dS is read once, much later than it was written; pipelined: it sits in L2 and is discarded there
      const uint64_t policy = l2_policy_evict_first();
      const uint8_t* base = static_cast<const uint8_t*>(p.ds_spill);
      uint32_t c = 0;  // stages filled so far
      [[maybe_unused]] int next_w = -1, group_end = 0;
      for (uint32_t t = 0;; ++t) {
        const uint32_t slot = t % kSchedSlots;
        mbar_wait(&item_empty[slot], ((t / kSchedSlots) & 1) ^ 1);
        int w;
        if constexpr (kPipe) {
          // the blocks of the claimed group one after the other, then another READY group if there is one (a successful
          // claim has seen every tile of the group delivered and ordered this thread's copies behind that)
          if (t == 0) {
            next_w = first_item;
            group_end = pipe_group_end(pp, first_item);
          } else if (next_w >= group_end) {
            next_w = pipe_try_claim_dq(pp, p.Sq, p.Sk, p.causal);
            group_end = next_w >= 0 ? pipe_group_end(pp, next_w) : 0;
          }
          w = next_w >= 0 ? next_w++ : -1;
        } else {
          w = atomicAdd(p.sched, 1);
          // every CTA draws exactly one id past the end; the very last draw resets the counter for the next launch
          if (w == n_items + static_cast<int>(gridDim.x) - 1) atomicExch(p.sched, 0);
          if (w >= n_items) w = -1;
        }
        sched[slot] = w;
        mbar_arrive(&item_full[slot]);
        if (w < 0) break;
        const Item it = item_of(w, n_q_blocks, p.Hq, kPipe);
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
      at_producer_end();
    }
  } else if (kRole != 2 && warp == kMmaWarp) {
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
        const Item it = item_of(w, n_q_blocks, p.Hq, kPipe);
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
  } else if (kRole == 1 && warp >= kDiscardWarp0 && warp < kDiscardWarp0 + kDiscardWarps) {
    // ======================================= L2 discard warps ======================================
    // Follow the tile stream of the producer: once a tile has landed in shared memory (full[s]) its 256 lines in global
    // memory are dead - discard them from L2 so that the dirty lines are never written back to HBM. Their arrival on
    // empty[s] keeps them in step with the ring.
    const uint8_t* base = static_cast<const uint8_t*>(p.ds_spill);
    const int lane_in_pair = (warp - kDiscardWarp0) * 32 + lane;  // 0 .. 63: four lines each
    uint32_t c = 0;
    for (uint32_t t = 0;; ++t) {
      const uint32_t slot = t % kSchedSlots;
      mbar_wait(&item_full[slot], (t / kSchedSlots) & 1);
      const int w = sched[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(&item_empty[slot]);
      if (w < 0) break;
      const Item it = item_of(w, n_q_blocks, p.Hq, kPipe);
      const int nblk = visible_kv_blocks_of(it.m, p.Sq, p.Sk, p.causal);
      const long long tile0 = (static_cast<long long>(it.batch) * p.Hq + it.head) * p.tiles_per_head +
                              ds_tile_prefix(it.m, p.Sq, p.Sk, p.causal);
      for (int i = 0; i < nblk; ++i, ++c) {
        const uint32_t s = c % kStages;
        mbar_wait(&full[s], (c / kStages) & 1);
        if (pp.discard) {
          const uint8_t* tile = base + (tile0 + i) * kDsTileBytes;
#pragma unroll
          for (int q = 0; q < kDsTileBytes / 128 / (kDiscardWarps * 32); ++q)
            asm volatile("discard.global.L2 [%0], 128;" ::"l"(tile + (q * kDiscardWarps * 32 + lane_in_pair) * 128) : "memory");
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
    }
  } else if (kRole != 1 && warp < 4) {
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
      const Item it = item_of(w, n_q_blocks, p.Hq, kPipe);
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
  if constexpr (!kPipe) {
    if (warp == kMmaWarp) {
      tcgen05_fence_after();
      tmem_dealloc<512>(0u);
    }
  } else if constexpr (kRole == 1) {
    // the next item initialises its barriers afresh, possibly in other places (the dK/dV role's layout): retire these
    // first (ordered before that by the caller's __syncthreads ahead of the next item)
    if (warp == kMmaWarp && lane == 0) {
      for (int i = 0; i < L::kNumBars; ++i) mbar_inval(&bars[i]);
    }
  }
}

}  // namespace dqs
}  // namespace fa2
