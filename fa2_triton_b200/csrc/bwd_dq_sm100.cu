// dQ kernel of the deterministic backward for sm_100a (B200).
//
// Replaces the "pid >= NUM_BLOCKS_KV" half of the reference's `_bwd_kernel` (src/backward/kernel.py:154-182):
// `_compute_row_blocks_dq` / `_compute_single_block_dq` (src/backward/compute_dq.py:82-261, 8-78). Like the
// reference, one CTA owns one 128-row query block and walks the key blocks it sees in ascending order, so dQ is
// a fixed-order sum with no atomics and no cross-CTA traffic - deterministic by construction.
//
// Arithmetic: S = Q K^T, P = exp2(S c - LSE2), dP = dO V^T, dS = P o (dP - delta) rounded to
// the input dtype, dQ = scale * (dS K) in fp32, stored in the caller's dtype.
//
// Persistent: one CTA per SM walks a longest-first list of (query block, head, batch) items; the TMA producer and the
// tensor pipe run ahead into the next item while the compute warps store the current dQ, so the ~10 k cycles of
// per-item prologue (operand loads) and epilogue are off the critical path (they were 35 % of the non-persistent
// version of this kernel).
//
// CTA = 20 warps:
//   warps 0-15  : four compute warpgroups; thread = query row (TMEM lane), warpgroup g owns key columns [32g, 32g+32).
//                 Four warps per scheduler instead of two: one warp of this code sustains ~0.3 instructions per clock
//                 (dependent FFMA -> MUFU -> FMUL -> F2FP chains), so the block's softmax/dS latency - which sits on the
//                 critical path between the tensor-pipe products - roughly halves (profiles/, ncu stall reasons).
//   warp 16 : MMA issuer       warp 17 : TMA producer       warps 18-19 : idle
// TMEM (512 columns): S ring [0,128) [128,256), dS(i) (16-bit) goes over S(i) | dP [256,384) | dQ [384,384+D)
// Tensor-pipe order:  S(0) dP(0) | S(i+1) dP(i+1) dQ(i) ...   - S is double buffered, so the exponentials of block i+1
// start as soon as dS(i) is handed over, and dP(i+1) - the next thing they need - is issued before dQ(i).
#include "fa2_kernels.cuh"
#include "ptx.cuh"

namespace fa2 {

#ifdef FA2_TRACE
__device__ long long g_dq_trace[4 * 64 * 8];
#define FA2_STAMP(role, it, ev)                                                                          \
  do {                                                                                                   \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (it) < 64)                              \
      g_dq_trace[((role) * 64 + (it)) * 8 + (ev)] = clock64();                                           \
  } while (0)
cudaError_t fetch_dq_trace(long long* host, size_t n) {
  return cudaMemcpyFromSymbol(host, g_dq_trace, n * sizeof(long long));
}
#else
#define FA2_STAMP(role, it, ev) do { } while (0)
#endif

namespace {

constexpr int kBlk = 128;
#ifndef FA2_DQ_CMP_WGS
#define FA2_DQ_CMP_WGS 4
#endif
constexpr int kCmpWgs = FA2_DQ_CMP_WGS;   // compute warpgroups (2 or 4)
constexpr int kCmpCols = kBlk / kCmpWgs;  // key columns per compute thread
constexpr int kDqThreads = (kCmpWgs + 1) * 128;
constexpr int kMmaWarp = kCmpWgs * 4, kTmaWarp = kMmaWarp + 1;
// register split (setmaxnreg): the launch bound gives floor(65536 / threads / 8) * 8 per thread
constexpr int kRegsCmp = kCmpWgs == 4 ? 104 : 184;
constexpr int kRegsIssue = kCmpWgs == 4 ? 64 : 88;
constexpr int kKStages = 3;
constexpr int kVStages = 2;
constexpr int kSchedSlots = 4;
// how many of every 4 pairs of exponentials are evaluated on the FMA pipes (ex2_emul2) instead of the MUFU: at D = 64 the
// three products of a block pair take 768 tensor-pipe cycles, its 16384 exponentials 1024 MUFU cycles
#ifndef FA2_DQ_EMUL
#define FA2_DQ_EMUL 1  // measured, dQ kernel alone: D = 64, S = 16384: 0/4 -> 3.45 ms, 1/4 -> 3.30, 2/4 -> 3.39; D = 128, S = 4096: 0/4 -> 0.786, 2/4 -> 0.748
#endif
__host__ __device__ constexpr bool dq_emulate_pair(int q) {
  return FA2_DQ_EMUL == 4 ? true : FA2_DQ_EMUL == 3 ? (q % 4 != 0) : FA2_DQ_EMUL == 2 ? (q % 2 == 1)
       : FA2_DQ_EMUL == 1 ? (q % 4 == 3) : false;
}

template <int D>
struct DqSmem {
  static constexpr int kTileBytes = kBlk * D * 2;
  static constexpr int kQOff = 0;
  static constexpr int kDOOff = kQOff + kTileBytes;
  static constexpr int kKOff = kDOOff + kTileBytes;
  static constexpr int kVOff = kKOff + kKStages * kTileBytes;
  static constexpr int kBarOff = kVOff + kVStages * kTileBytes;
  // q_full, qdo_empty, k_full[3], k_empty[3], v_full[2], v_empty[2], s_full[2], dp_full, ds_full, dq_done, dq_empty,
  // item_full[4], item_empty[4]
  static constexpr int kNumBars = 2 + 2 * kKStages + 2 * kVStages + 2 + 4 + 2 * kSchedSlots;
  static constexpr int kSchedOff = kBarOff + kNumBars * 8;  // kSchedSlots work-item ids
  static constexpr int kTmemSlotOff = kSchedOff + kSchedSlots * 4;
  static constexpr int kBytes = kTmemSlotOff + 16;
  static_assert(kBytes <= 232448, "dQ kernel smem budget exceeded");
};

__device__ __forceinline__ int visible_kv_blocks(int row0, int Sq, int Sk, int causal) {
  if (row0 >= Sq) return 0;
  const int last_row = min(row0 + kBlk, Sq) - 1;
  const int cols = causal ? min(Sk, last_row + Sk - Sq + 1) : Sk;
  return cols <= 0 ? 0 : (cols + kBlk - 1) / kBlk;
}

// Work item w -> (query block, head, batch): the query blocks of one (batch, head) are consecutive items, longest
// first, so the CTAs that are running at any moment share the K / V of a handful of heads in L2 (with a query-block-
// major order every CTA streams its own head from HBM: measured 0.78 ms instead of 0.68). Items are handed out
// dynamically (one atomicAdd per item by the producer thread, published to the other roles through a 4-slot ring).
struct DqItem {
  int m, head, batch;
};
// The last heads are dealt block-major (fa2_kernels.cuh: tail_balanced_item).
__device__ __forceinline__ DqItem dq_item(int w, int n_q_blocks, int Hq, int B, int tail_heads) {
  const ItemCoord c = tail_balanced_item(w, n_q_blocks, Hq, Hq * B, tail_heads);
  DqItem it;
  it.m = n_q_blocks - 1 - c.x;
  it.head = c.head;
  it.batch = c.batch;
  return it;
}

}  // namespace

// kBias: the variant that adds the attention bias to the recomputed scores (extras.cuh)
template <int D, bool kBf16, bool kBias>
__global__ void __launch_bounds__(kDqThreads, 1)
fa2_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                  const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                  const __grid_constant__ CUtensorMap tmap_dq, const DqParams p) {
  using L = DqSmem<D>;
  constexpr int kSub = D / 64;
  constexpr int kSubBytes = kBlk * 128;
  constexpr uint32_t kTmemS = 0, kTmemDP = 256, kTmemDQ = 384;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem + L::kQOff;
  uint8_t* sDO = smem + L::kDOOff;
  uint8_t* sK = smem + L::kKOff;
  uint8_t* sV = smem + L::kVOff;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* q_full = bars;
  uint64_t* qdo_empty = bars + 1;
  uint64_t* k_full = bars + 2;
  uint64_t* k_empty = k_full + kKStages;
  uint64_t* v_full = k_empty + kKStages;
  uint64_t* v_empty = v_full + kVStages;
  uint64_t* s_full = v_empty + kVStages;
  uint64_t* dp_full = s_full + 2;
  uint64_t* ds_full = dp_full + 1;
  uint64_t* dq_done = ds_full + 1;
  uint64_t* dq_empty = dq_done + 1;
  uint64_t* item_full = dq_empty + 1;
  uint64_t* item_empty = item_full + kSchedSlots;
  volatile int* sched = reinterpret_cast<volatile int*>(smem + L::kSchedOff);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlotOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int n_q_blocks = (p.Sq + kBlk - 1) / kBlk;
  const int n_items = n_q_blocks * p.Hq * p.B;
  const int group = p.Hq / p.Hkv;
  const int shift = p.Sk - p.Sq;

  if ((smem_u32(smem) & 1023u) != 0) __trap();

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_do);
  }
  if (warp == kMmaWarp) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      mbar_init(qdo_empty, 1);
      for (int s = 0; s < kKStages; ++s) {
        mbar_init(&k_full[s], 1);
        mbar_init(&k_empty[s], 1);
      }
      for (int s = 0; s < kVStages; ++s) {
        mbar_init(&v_full[s], 1);
        mbar_init(&v_empty[s], 1);
      }
      mbar_init(&s_full[0], 1);
      mbar_init(&s_full[1], 1);
      mbar_init(dp_full, 1);
      mbar_init(ds_full, kCmpWgs * 128);
      mbar_init(dq_done, 1);
      mbar_init(dq_empty, kCmpWgs * 128);
      for (int i = 0; i < kSchedSlots; ++i) {
        mbar_init(&item_full[i], 1);
        mbar_init(&item_empty[i], 1 + kCmpWgs * 128);  // MMA thread + every compute thread
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

  // register pool: 640 threads x 96 (launch bound) = 61440 = 4 x 128 x 104 + 128 x 64   (4 compute warpgroups)
  //                384 threads x 168               = 64512 >= 2 x 128 x 184 + 128 x 88  (2 compute warpgroups)
  if (warp >= kMmaWarp) {
    reg_dealloc<kRegsIssue>();
  if (warp == kTmaWarp) {
    // ========================================= TMA producer ========================================
    // Runs ahead of the consumers across work items: K / V of the next item stream into the rings as slots free up,
    // its Q / dO as soon as the last S and dP of the current item have been issued.
    if (elect_one()) {
      uint32_t kc = 0, vc = 0;  // K and V blocks loaded so far
      for (uint32_t t = 0;; ++t) {
        // fetch the next item and publish it to the MMA and compute roles
        const uint32_t slot = t % kSchedSlots;
        mbar_wait(&item_empty[slot], ((t / kSchedSlots) & 1) ^ 1);
        const int w = atomicAdd(p.sched, 1);
        // every CTA draws exactly one id past the end; the very last draw resets the counter for the next launch
        if (w == n_items + static_cast<int>(gridDim.x) - 1) atomicExch(p.sched, 0);
        sched[slot] = (w < n_items) ? w : -1;
        mbar_arrive(&item_full[slot]);
        if (w >= n_items) break;
        const DqItem it = dq_item(w, n_q_blocks, p.Hq, p.B, p.tail_heads);
        const int row0 = it.m * kBlk;
        const int head_kv = it.head / group;
        const int nblk = visible_kv_blocks(row0, p.Sq, p.Sk, p.causal);
        mbar_wait(qdo_empty, (t & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, 2 * L::kTileBytes);
        for (int j = 0; j < kSub; ++j) {
          tma_load_4d(sQ + j * kSubBytes, &tmap_q, q_full, j * 64, row0, it.head, it.batch);
          tma_load_4d(sDO + j * kSubBytes, &tmap_do, q_full, j * 64, row0, it.head, it.batch);
        }
        for (int i = 0; i < nblk; ++i, ++kc, ++vc) {
          const uint32_t ks = kc % kKStages, vs = vc % kVStages;
          mbar_wait(&k_empty[ks], ((kc / kKStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&k_full[ks], L::kTileBytes);
          for (int j = 0; j < kSub; ++j)
            tma_load_4d(sK + ks * L::kTileBytes + j * kSubBytes, &tmap_k, &k_full[ks], j * 64, i * kBlk, head_kv, it.batch);
          mbar_wait(&v_empty[vs], ((vc / kVStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[vs], L::kTileBytes);
          for (int j = 0; j < kSub; ++j)
            tma_load_4d(sV + vs * L::kTileBytes + j * kSubBytes, &tmap_v, &v_full[vs], j * 64, i * kBlk, head_kv, it.batch);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ========================================== MMA issuer =========================================
    if (elect_one()) {
      constexpr uint32_t idesc_kk = make_idesc_f16(kBlk, kBlk, kBf16, false, false);  // S, dP: [q x D] [kv x D]^T
      constexpr uint32_t idesc_dq = make_idesc_f16(kBlk, D, kBf16, false, true);      // dQ: dS (TMEM) x K (MN-major)
      const SmemDesc q_kmaj = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const SmemDesc do_kmaj = make_smem_desc_sw128(smem_u32(sDO), 16, 1024);
      const SmemDesc k_kmaj = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const SmemDesc v_kmaj = make_smem_desc_sw128(smem_u32(sV), 16, 1024);
      const SmemDesc k_mn = make_smem_desc_sw128(smem_u32(sK), kSubBytes, 1024);

      auto issue_kmajor = [&](uint32_t d_tmem, SmemDesc a, SmemDesc b) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = (k / 4) * kSubBytes + (k % 4) * 32;
          umma_ss(d_tmem, advance(a, off), advance(b, off), idesc_kk, k > 0 ? 1u : 0u);
        }
      };
      // g = index of this item's first key block in the CTA-wide block sequence: K(g+i) lives in ring stage
      // (g+i) % kKStages, S(g+i) in TMEM buffer (g+i) & 1, and the dp_full / ds_full phases follow g+i as well
      auto issue_s = [&](uint32_t gi) {
        const uint32_t ks = gi % kKStages;
        mbar_wait(&k_full[ks], (gi / kKStages) & 1);
        tcgen05_fence_after();
        issue_kmajor(kTmemS + (gi & 1) * 128, q_kmaj, advance(k_kmaj, ks * L::kTileBytes));
        umma_commit(&s_full[gi & 1]);
      };
      auto issue_dp = [&](uint32_t gi) {
        const uint32_t vs = gi % kVStages;
        mbar_wait(&v_full[vs], (gi / kVStages) & 1);
        tcgen05_fence_after();
        issue_kmajor(kTmemDP, do_kmaj, advance(v_kmaj, vs * L::kTileBytes));
        umma_commit(dp_full);
        umma_commit(&v_empty[vs]);
      };

      uint32_t g = 0;
      for (uint32_t t = 0;; ++t) {
        const uint32_t slot = t % kSchedSlots;
        mbar_wait(&item_full[slot], (t / kSchedSlots) & 1);
        const int w = sched[slot];
        mbar_arrive(&item_empty[slot]);
        if (w < 0) break;
        const DqItem it = dq_item(w, n_q_blocks, p.Hq, p.B, p.tail_heads);
        const int nblk = visible_kv_blocks(it.m * kBlk, p.Sq, p.Sk, p.causal);
        mbar_wait(q_full, t & 1);
        tcgen05_fence_after();
        if (nblk > 0) {
          issue_s(g);
          issue_dp(g);
          if (nblk == 1) umma_commit(qdo_empty);  // no further S / dP will read this item's Q and dO
        } else {
          umma_commit(qdo_empty);
        }
        for (int i = 0; i < nblk; ++i) {
          const uint32_t gi = g + i;
          const uint32_t ks = gi % kKStages;
          // S(i+1) = Q K(i+1)^T into the other S buffer. Issued here, one block ahead, rather than two blocks ahead
          // right after dQ(i-1): K(i+1) was requested when dQ(i-2) retired, so this wait is (almost) never a stall,
          // while the thread would otherwise sit idle waiting for dS(i) anyway.
          FA2_STAMP(2, i, 4);
          if (i + 1 < nblk) issue_s(gi + 1);
          FA2_STAMP(2, i, 5);
          mbar_wait(ds_full, gi & 1);  // dS(i) is in tensor memory, and dP(i) / S(i) have been read
          tcgen05_fence_after();
          FA2_STAMP(2, i, 0);
          // dP(i+1) = dO V(i+1)^T first: it is what the compute warps wait for next
          if (i + 1 < nblk) issue_dp(gi + 1);
          FA2_STAMP(2, i, 1);
          // dQ (+)= dS(i) K(i)   (A = dS from TMEM: compute warpgroup g packed its 32 columns, k-steps 2g and 2g+1,
          // into the first 16 columns of its slice of the S buffer block i was computed in)
          if (i == 0) {
            mbar_wait(dq_empty, (t & 1) ^ 1);  // the previous item's dQ has left tensor memory
            tcgen05_fence_after();
          }
#pragma unroll
          for (int k = 0; k < kBlk / 16; ++k) {
            constexpr int kStepsPerWg = 8 / kCmpWgs;  // k-steps (16 key columns each) per compute warpgroup
            const uint32_t a_col = (k / kStepsPerWg) * kCmpCols + (k % kStepsPerWg) * 8;
            umma_ts(kTmemDQ, kTmemS + (gi & 1) * 128 + a_col, advance(k_mn, ks * L::kTileBytes + k * 2048), idesc_dq,
                    (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&k_empty[ks]);  // K(i) has fed S(i) and dQ(i)
          FA2_STAMP(2, i, 2);
          // Q and dO are dead once the last S and the last dP (both issued at i = nblk-2) are in the pipe
          if (i == nblk - 2) umma_commit(qdo_empty);
          FA2_STAMP(2, i, 3);
        }
        umma_commit(dq_done);
        g += nblk;
      }
    }
  }
  } else {
    // ====================================== compute warpgroups =====================================
    reg_alloc<kRegsCmp>();
    const int wg = warp >> 2;
    const int r = threadIdx.x & 127;  // query row inside the block == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tDP = lane_sel + kTmemDP + wg * kCmpCols;
    const float c = p.scale_log2;

    uint32_t g = 0;
    for (uint32_t t = 0;; ++t) {
      const uint32_t slot = t % kSchedSlots;
      mbar_wait(&item_full[slot], (t / kSchedSlots) & 1);
      const int w = sched[slot];
      mbar_arrive(&item_empty[slot]);
      if (w < 0) break;
      const DqItem it = dq_item(w, n_q_blocks, p.Hq, p.B, p.tail_heads);
      const int row0 = it.m * kBlk;
      const int row = row0 + r;
      const int nblk = visible_kv_blocks(row0, p.Sq, p.Sk, p.causal);
      // padded rows carry LSE = delta = 0 (forward / delta kernels); their dQ rows are never stored
      const int64_t stat_off = (static_cast<int64_t>(it.batch) * p.Hq + it.head) * p.lse_stride + row;
      const float nlse = -p.lse[stat_off];  // -inf -> +inf for fully masked rows: handled by the mask path
      const float ndelta = -p.delta[stat_off];

      for (int i = 0; i < nblk; ++i) {
        const uint32_t gi = g + i;
        const uint32_t tS = lane_sel + kTmemS + (gi & 1) * 128 + wg * kCmpCols;
        const int col0 = i * kBlk + wg * kCmpCols;
        const bool need_mask = (i * kBlk + kBlk > p.Sk) || (p.causal && (i * kBlk + kBlk - 1 > row0 + shift));

        // ---- P = exp2(S c - LSE) ----
        if (r == 0 && (wg == 0 || wg == kCmpWgs - 1)) FA2_STAMP(wg ? 1 : 0, i, 0);
        mbar_wait(&s_full[gi & 1], (gi >> 1) & 1);
        tcgen05_fence_after();
        if (r == 0 && (wg == 0 || wg == kCmpWgs - 1)) FA2_STAMP(wg ? 1 : 0, i, 1);
        // packed f32x2 arithmetic throughout: the four schedulers' issue slots - shared with the MMA-issuing thread on
        // scheduler 0 - are the scarce resource here, not the FMA or MUFU pipes
        constexpr int kPairs = kCmpCols / 2;
        uint32_t su[kCmpCols];
#pragma unroll
        for (int h = 0; h < kCmpCols / 32; ++h) tmem_ld_32x32b_x32(tS + h * 32, *reinterpret_cast<uint32_t(*)[32]>(&su[h * 32]));
        tmem_wait_ld();
        float2 pf[kPairs];
        const float2 c2 = make_float2(c, c), nl2 = make_float2(nlse, nlse);
#pragma unroll
        for (int j = 0; j < kPairs; ++j) {
          const float2 x = ffma2(make_float2(__uint_as_float(su[2 * j]), __uint_as_float(su[2 * j + 1])), c2, nl2);
          if constexpr (kBias) {
            // P = exp2(S c + bias log2(e) - LSE) (src/backward/compute_dq.py:40-51); on the MUFU so that a bias of
            // -inf gives exactly 0; a row hidden entirely that way (LSE = -inf) gets P = 0
            float xx[2] = {x.x, x.y};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = col0 + 2 * j + e;
              float b = 0.f;
              if (row < p.Sq && col < p.Sk)
                b = bias_load<kBf16>(p.bias, it.batch * p.bias.sb + it.head * p.bias.sh + static_cast<int64_t>(row) * p.bias.sm + col);
              xx[e] = (nlse == INFINITY) ? -INFINITY : fmaf(b, 1.4426950408889634f, xx[e]);
            }
            pf[j] = make_float2(ex2_approx(xx[0]), ex2_approx(xx[1]));
          } else {
          if (dq_emulate_pair(j)) pf[j] = ex2_emul2(x);
          else pf[j] = make_float2(ex2_approx(x.x), ex2_approx(x.y));
          }
        }
        if (need_mask) {
          const int limit = p.causal ? min(p.Sk - 1, row + shift) : (p.Sk - 1);
#pragma unroll
          for (int j = 0; j < kPairs; ++j) {
            pf[j].x = (col0 + 2 * j <= limit) ? pf[j].x : 0.f;
            pf[j].y = (col0 + 2 * j + 1 <= limit) ? pf[j].y : 0.f;
          }
        }

        // ---- dS = P o (dP - delta), 16-bit, written over the head of this warpgroup's (now dead) S columns, so
        //      that the dP columns are free for dP(i+1) at once ----
        if (r == 0 && (wg == 0 || wg == kCmpWgs - 1)) FA2_STAMP(wg ? 1 : 0, i, 2);
        mbar_wait(dp_full, gi & 1);
        tcgen05_fence_after();
        if (r == 0 && (wg == 0 || wg == kCmpWgs - 1)) FA2_STAMP(wg ? 1 : 0, i, 3);
        uint32_t dpu[kCmpCols];
#pragma unroll
        for (int h = 0; h < kCmpCols / 32; ++h) tmem_ld_32x32b_x32(tDP + h * 32, *reinterpret_cast<uint32_t(*)[32]>(&dpu[h * 32]));
        tmem_wait_ld();
        uint32_t pk[kPairs];
        const float2 nd2 = make_float2(ndelta, ndelta);
#pragma unroll
        for (int j = 0; j < kPairs; ++j) {
          const float2 d = fmul2(pf[j], fadd2(make_float2(__uint_as_float(dpu[2 * j]), __uint_as_float(dpu[2 * j + 1])), nd2));
          pk[j] = pack2<kBf16>(d.x, d.y);
        }
        if constexpr (kPairs == 16) tmem_st_32x32b_x16(tS, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
        else tmem_st_32x32b_x32(tS, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
        tmem_wait_st();
        tcgen05_fence_before();
        mbar_arrive(ds_full);
        if (r == 0 && (wg == 0 || wg == kCmpWgs - 1)) FA2_STAMP(wg ? 1 : 0, i, 4);
      }
      g += nblk;

      // ---- epilogue: dQ * scale -> caller's tensor (the tensor pipe is already working on the next item) ----
      mbar_wait(dq_done, t & 1);
      tcgen05_fence_after();
      const float mul = p.softmax_scale;
      // 32 head-dim columns per chunk; warpgroup wg takes chunks wg, wg + kCmpWgs, ...
      constexpr int kChunksPerWg = (D / 32 + kCmpWgs - 1) / kCmpWgs;
      uint32_t oo[kChunksPerWg][32];
#pragma unroll
      for (int cc = 0; cc < kChunksPerWg; ++cc) {
        const int chunk = wg + cc * kCmpWgs;
        if (chunk < D / 32 && nblk > 0) {
          tmem_ld_32x32b_x32(lane_sel + kTmemDQ + chunk * 32, oo[cc]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) oo[cc][j] = 0u;
        }
      }
      tmem_wait_ld();
      tcgen05_fence_before();
      mbar_arrive(dq_empty);  // the next item's first dQ product may overwrite the accumulator
#pragma unroll
      for (int cc = 0; cc < kChunksPerWg; ++cc) {
      const int chunk = wg + cc * kCmpWgs;
      const bool active = chunk < D / 32;
      const uint32_t (&o)[32] = oo[cc];
      const int col = chunk * 32;  // first head-dim column of this thread's chunk
      // Straight from registers: 64 contiguous bytes (16-bit) or 128 (fp32) per thread and row. With every shared-memory
      // byte spent on operand rings there is no staging tile for a TMA store, and none is needed: it is 32 KB per item.
      if (active && row < p.Sq) {
        const int64_t off = it.batch * p.dq_strides[0] + it.head * p.dq_strides[1] + static_cast<int64_t>(row) * p.dq_strides[2] + col;
        if (p.dq_is_16bit) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.dq) + off);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            if (col + q4 * 8 >= p.d_actual) break;  // zero-padded head dim: nothing to store there
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
            if (col + q4 * 4 >= p.d_actual) break;
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

template <int D, bool kBf16, bool kBias>
static cudaError_t launch_bwd_dq_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                                   const CUtensorMap& tdo, const CUtensorMap& tdq, const DqParams& p,
                                   cudaStream_t stream) {
  auto kernel = fa2_bwd_dq_kernel<D, kBf16, kBias>;
  constexpr int smem = DqSmem<D>::kBytes;
  if (cudaError_t e = ensure_dynamic_smem<fa2_bwd_dq_kernel<D, kBf16, kBias>>(smem); e != cudaSuccess) return e;
  const int n_items = ((p.Sq + kBlk - 1) / kBlk) * p.Hq * p.B;
  const int sms = current_sm_count();
  dim3 grid(n_items < sms ? n_items : sms);  // persistent: one CTA per SM, items dealt dynamically
  DqParams pp = p;
  pp.tail_heads = tail_heads_for(static_cast<size_t>(p.Sk) * D * 2 * 2, p.causal);  // a head's items stream its K and V
  kernel<<<grid, kDqThreads, smem, stream>>>(tq, tk, tv, tdo, tdq, pp);
  return cudaGetLastError();
}

cudaError_t launch_bwd_dq(int D, bool bf16, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                          const CUtensorMap& tdo, const CUtensorMap& tdq, const DqParams& p, cudaStream_t stream) {
  const bool has_bias = p.bias.ptr != nullptr;
#define FA2_DQ_CASE(DD, BF)                                                                     \
  if (D == DD && bf16 == BF)                                                                    \
    return has_bias ? launch_bwd_dq_t<DD, BF, true>(tq, tk, tv, tdo, tdq, p, stream)            \
                    : launch_bwd_dq_t<DD, BF, false>(tq, tk, tv, tdo, tdq, p, stream);
  FA2_DQ_CASE(128, true) FA2_DQ_CASE(128, false) FA2_DQ_CASE(64, true) FA2_DQ_CASE(64, false)
#undef FA2_DQ_CASE
  return cudaErrorInvalidValue;
}

}  // namespace fa2
