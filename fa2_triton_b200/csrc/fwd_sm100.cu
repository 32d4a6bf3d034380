// FlashAttention-2 forward for sm_100a (B200), the variants with an additive bias and / or forward dropout
// (SURVEY.md section 8f rank 4). The plain forward lives in fwd2_sm100.cu; this translation unit keeps the first-generation
// pipeline - TMA -> smem ring -> tcgen05.mma (S = Q K^T, O += P V) with accumulators in tensor memory, two 128-row Q
// tiles per CTA whose softmax warpgroups ping-pong against the single MMA-issuing warp - because the bias staging
// (one [128 x 128] block per tile through shared memory by TMA) and the dropout draw are built around 128-key blocks.
//
// Replaces the reference's `_fwd_kernel` (src/forward/kernel.py:61-291) + `compute_row_block`
// (src/forward/compute_row_blocks.py:8-103) for `BIAS_ON` / `USE_DROPOUT`. Same arithmetic contract:
//   S = Q K^T (fp32 acc) + bias / scale; masked where col >= Sk or (causal) col > row + Sk - Sq;
//   log2-domain online softmax with c = softmax_scale * log2(e); P rounded to the input dtype before PV;
//   O = (P V) / l in the input dtype;  LSE2 = m + log2(l) in fp32 (src/forward/kernel.py:119,259-272).
//
// CTA = 2 softmax warpgroups (one per Q tile; thread = query row = TMEM lane) + one issue warpgroup
//   (warp 0 MMA issuer, warp 1 TMA producer, 2 idle warps).
// TMEM (512 columns): S0 [0,128) | S1 [128,256) | O0 [256,256+D) | O1 [256+D,256+2D).
//   P_t (16-bit) overwrites the head of the S_t columns of the thread that wrote it, once it has read its row.
#include <type_traits>

#include "fa2_kernels.cuh"
#include "ptx.cuh"

namespace fa2 {

namespace {

constexpr int kBM = 128;  // query rows per tile (= TMEM lanes)
constexpr int kBN = 128;  // key rows per block
constexpr int kCols = kBN;                       // key columns per softmax thread
constexpr int kSmxThreads = 2 * 128;             // softmax threads per CTA
constexpr int kTileSmxThreads = 128;             // ... per Q tile
constexpr int kFwdThreads = kSmxThreads + 128;
constexpr int kMmaWarp = kSmxThreads / 32, kTmaWarp = kMmaWarp + 1;
// register pools (setmaxnreg): 384 threads x 168 = 64512 = 2 x 128 x 208 + 128 x 88
constexpr int kRegsSmx = 208;
constexpr int kRegsIssue = 88;
constexpr float kRescaleThreshold = 8.0f;  // log2 units; P stays <= 2^8 between rescales

template <int D, bool kExt = false, bool kBiasTma = false>
struct FwdSmem {
  static constexpr int kStages = (D == 128) ? 2 : 4;  // K ring
  // The TMA-staged bias needs a whole [128 x 128] block per Q tile in flight (64 KB): at head dim 128 it takes the second
  // V stage (V(i + 1) is requested when P V(i) of both tiles has completed and needed about half a block period later)
  static constexpr int kVStages = (kBiasTma && D == 128) ? 1 : kStages;
  static constexpr int kTileBytes = kBM * D * 2;  // Q, K and V tiles all have 128 rows
  static constexpr int kQOff = 0;
  static constexpr int kKOff = kQOff + 2 * kTileBytes;
  static constexpr int kVOff = kKOff + kStages * kTileBytes;
  // kExt (bias variants): a region right behind the V ring (1024-byte aligned). A 16-bit bias that TMA can address lands
  // there as two [128 rows x 64 keys] 128-byte-swizzled sub-tiles per Q tile and block (2 x 32 KB); any other bias is
  // staged by the softmax warps themselves - per warp 32 rows x 64 bytes, rows 80 bytes apart so that both the coalesced
  // writes (8 rows x 64 B per instruction) and the row-per-lane 16-byte reads are conflict-free (8 x 2560 B in 32 KB).
  static constexpr int kBiasOff = kVOff + kVStages * kTileBytes;
  static constexpr int kBiasSubBytes = kBM * 128;
  static constexpr int kBiasBufBytes = (kBiasTma ? 2 : 1) * kBiasSubBytes;
  static constexpr int kStageOff = kBiasOff;
  static constexpr int kStageRow = 80;
  static constexpr int kStageWarpBytes = 32 * kStageRow;
  static constexpr int kBarOff = kBiasOff + (kExt ? 2 * kBiasBufBytes : 0);
  // barriers: q_full, k_full[S], k_empty[S], v_full[S], v_empty[S], s_full[2], p_full[2], o_full[2], bias_full[2], bias_empty[2]
  static constexpr int kNumBars = 1 + 4 * kStages + 6 + 4;
  static constexpr int kTmemSlotOff = kBarOff + kNumBars * 8;
  static constexpr int kBytes = kTmemSlotOff + 16;
  static_assert(8 * kStageWarpBytes <= 2 * kBiasBufBytes, "the two bias layouts share one region");
  static constexpr int kDynBytes = kBytes + 1024;  // slack for manual 1024-byte alignment
  static_assert(kDynBytes <= 227 * 1024, "forward shared memory exceeds the 227 KB a CTA can have");
};

// number of key blocks query tile [row0, row0+128) has to visit
__device__ __forceinline__ int num_kv_blocks(int row0, int Sq, int Sk, int causal) {
  if (row0 >= Sq) return 0;
  int last_row = min(row0 + kBM, Sq) - 1;
  int cols = causal ? min(Sk, last_row + Sk - Sq + 1) : Sk;
  return cols <= 0 ? 0 : (cols + kBN - 1) / kBN;
}

}  // namespace

// kExt: variants with an additive bias (bit 0; with bit 2 it is staged through shared memory by TMA) and / or forward
// dropout (bit 1), see extras.cuh; 0 = the plain kernel
template <int D, bool kBf16, int kExt>
__global__ void __launch_bounds__(kFwdThreads, 1)
fa2_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
               const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
               const __grid_constant__ CUtensorMap tmap_bias, const FwdParams p) {
  constexpr bool kBiasTma = (kExt & 4) != 0;
  using L = FwdSmem<D, (kExt & 1) != 0, kBiasTma>;
  constexpr int kVStages = L::kVStages;  // the bias travels through shared memory by TMA (with bit 0)
  constexpr int kStages = L::kStages;
  constexpr int kSub = D / 64;           // 64-column (128-byte) sub-tiles per row
  constexpr int kSubBytes = kBM * 128;   // one [128 x 64] 16-bit sub-tile

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + L::kQOff;
  uint8_t* sK = smem + L::kKOff;
  uint8_t* sV = smem + L::kVOff;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + kStages;
  uint64_t* v_full = k_empty + kStages;
  uint64_t* v_empty = v_full + kStages;
  uint64_t* s_full = v_empty + kStages;
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 2;
  [[maybe_unused]] uint64_t* bias_full = o_full + 2;
  [[maybe_unused]] uint64_t* bias_empty = bias_full + 2;
  [[maybe_unused]] uint8_t* sBias = smem + L::kBiasOff;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlotOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // heaviest (largest row index) tile pairs first: better tail behaviour under causal masking
  const ItemCoord cta = tail_balanced_cta(p.tail_log2);  // the last heads block-major (fa2_kernels.cuh)
  const int pair = p.n_pairs - 1 - cta.x;
  const int head = cta.head;
  const int batch = cta.batch;
  const int head_kv = head / (p.Hq / p.Hkv);
  const int row0 = pair * 2 * kBM;
  // variable-length mode: this batch element's own length and tensor maps (bounded by it; batch coordinate 0)
  int Sq = p.Sq, Sk = p.Sk, bc = batch;
  const CUtensorMap *mq = &tmap_q, *mk = &tmap_k, *mv = &tmap_v, *mo = &tmap_o;
  if (p.seqlens != nullptr) {
    Sq = Sk = min(max(p.seqlens[batch], 0), p.Sq);  // clamped like the tensor maps (varlen_sm100.cu)
    const CUtensorMap* m = p.vmaps + static_cast<size_t>(batch) * kVarlenMapsPerBatch;
    mq = m; mk = m + 1; mv = m + 2; mo = m + 3; bc = 0;
    if (row0 >= Sq) return;  // tile pair in the padding: its rows of O and LSE were zero-filled by the caller
  }

  const int nblk0 = num_kv_blocks(row0, Sq, Sk, p.causal);
  const int nblk1 = num_kv_blocks(row0 + kBM, Sq, Sk, p.causal);
  const int nblk_max = max(nblk0, nblk1);

  if (warp == kTmaWarp && lane == 0) {
    if (p.seqlens != nullptr) {  // maps published by fa2_varlen_maps_kernel
      tmap_acquire(mq);
      tmap_acquire(mk);
      tmap_acquire(mv);
    }
    tma_prefetch_desc(mq);
    tma_prefetch_desc(mk);
    tma_prefetch_desc(mv);
    if (p.seqlens == nullptr) tma_prefetch_desc(mo);
  }
  if (warp == kMmaWarp) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int s = 0; s < kStages; ++s) {
        mbar_init(&k_full[s], 1);
        mbar_init(&k_empty[s], 1);
        mbar_init(&v_full[s], 1);
        mbar_init(&v_empty[s], 1);
      }
      for (int t = 0; t < 2; ++t) {
        mbar_init(&s_full[t], 1);
        mbar_init(&p_full[t], kTileSmxThreads);
        mbar_init(&o_full[t], 1);
        mbar_init(&bias_full[t], 1);
        mbar_init(&bias_empty[t], kTileSmxThreads);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  // This CTA allocated all 512 columns, so the allocation starts at lane 0 / column 0. Using the literal 0 (after
  // checking it) keeps every tcgen05 address provably warp-uniform: ptxas then feeds UTCHMMA / LDTM / STTM from
  // uniform registers directly instead of wrapping each one in an ELECT + R2UR.BROADCAST "waterfall" (which cost
  // ~100 issue cycles per MMA and made the single issuing thread the bottleneck - DESIGN.md section 4;
  // profiles/r01d_fa2_fwd_kernel_opcodes.txt shows what is left of R2UR).
  if (*tmem_slot != 0u) __trap();
  constexpr uint32_t tmem_base = 0u;

  // setmaxnreg at the top of each role branch (ptxas widens the budget only for dominated code)
  if (warp >= kMmaWarp) {
    reg_dealloc<kRegsIssue>();
  if (warp == kTmaWarp) {
    // ======================================= TMA producer =======================================
    if (nblk_max > 0 && elect_one()) {
      const int n_q_tiles = (row0 + kBM < Sq) ? 2 : 1;
      mbar_arrive_expect_tx(q_full, n_q_tiles * L::kTileBytes);
      for (int t = 0; t < n_q_tiles; ++t)
        for (int j = 0; j < kSub; ++j)
          tma_load_4d(sQ + t * L::kTileBytes + j * kSubBytes, mq, q_full, j * 64, row0 + t * kBM, head, bc);
      for (int i = 0; i < nblk_max; ++i) {
        const int s = i % kStages;
        const uint32_t ph = (i / kStages) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], L::kTileBytes);
        for (int j = 0; j < kSub; ++j)
          tma_load_4d(sK + s * L::kTileBytes + j * kSubBytes, mk, &k_full[s], j * 64, i * kBN, head_kv, bc);
        const int sv = i % kVStages;
        const uint32_t phv = (i / kVStages) & 1;
        mbar_wait(&v_empty[sv], phv ^ 1);
        mbar_arrive_expect_tx(&v_full[sv], L::kTileBytes);
        for (int j = 0; j < kSub; ++j)
          tma_load_4d(sV + sv * L::kTileBytes + j * kSubBytes, mv, &v_full[sv], j * 64, i * kBN, head_kv, bc);
      }
    }
  } else if (warp == kMmaWarp) {
    // ======================================== MMA issuer ========================================
    if (nblk_max > 0 && elect_one()) {
      constexpr uint32_t idesc_qk = make_idesc_f16(kBM, kBN, kBf16, false, false);
      const uint32_t idesc_pv = make_idesc_f16(kBM, p.d_cols, kBf16, false, true);  // N = the tensors' head dim (x16)
      const int d_steps = p.d_steps;
      const int nblk[2] = {nblk0, nblk1};

      const SmemDesc q_desc = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);       // K-major
      const SmemDesc k_desc = make_smem_desc_sw128(smem_u32(sK), 16, 1024);       // K-major
      const SmemDesc v_desc = make_smem_desc_sw128(smem_u32(sV), kSubBytes, 1024);  // MN-major
      auto issue_qk = [&](int t, int stage) {
        const SmemDesc qa = advance(q_desc, t * L::kTileBytes);
        const SmemDesc ka = advance(k_desc, stage * L::kTileBytes);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = (k / 4) * kSubBytes + (k % 4) * 32;
          if (k < d_steps) umma_ss(tmem_base + t * kBN, advance(qa, off), advance(ka, off), idesc_qk, k > 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int t, int stage, bool accumulate) {
        const SmemDesc va = advance(v_desc, stage * L::kTileBytes);
#pragma unroll
        for (int k = 0; k < kBN / 16; ++k) {
          // softmax thread h of a row packed its kCols probabilities (k-steps h * kCols / 16 ...) into the first
          // kCols / 2 columns of its own slice of S
          constexpr int kStepsPerHalf = kCols / 16;
          const uint32_t a_col = (k / kStepsPerHalf) * kCols + (k % kStepsPerHalf) * 8;
          umma_ts(tmem_base + 256 + t * D, tmem_base + t * kBN + a_col, advance(va, k * 2048), idesc_pv,
                  (accumulate || k > 0) ? 1u : 0u);
        }
      };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tcgen05_fence_after();
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (nblk[t] > 0) {
          issue_qk(t, 0);
          umma_commit(&s_full[t]);
        }
      }
      umma_commit(&k_empty[0]);

      for (int i = 0; i < nblk_max; ++i) {
        const int sn = (i + 1) % kStages;
        const uint32_t phn = ((i + 1) / kStages) & 1;
        bool next_k_ready = false;
        const int sv = i % kVStages;
        mbar_wait(&v_full[sv], (i / kVStages) & 1);
        if (i + 1 < nblk_max) {  // K(i+1) is usually there long before P(i): take that wait off the P -> S chain
          mbar_wait(&k_full[sn], phn);
          next_k_ready = true;
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (i < nblk[t]) {
            mbar_wait(&p_full[t], i & 1);
            tcgen05_fence_after();
            issue_pv(t, sv, i > 0);
            if (i + 1 < nblk[t]) {
              if (!next_k_ready) {
                mbar_wait(&k_full[sn], phn);
                tcgen05_fence_after();
                next_k_ready = true;
              }
              issue_qk(t, sn);
              umma_commit(&s_full[t]);
            } else {
              umma_commit(&o_full[t]);
            }
          }
        }
        umma_commit(&v_empty[sv]);
        if (i + 1 < nblk_max) umma_commit(&k_empty[sn]);
      }
    }
  }
  } else {
    // ===================================== softmax warpgroups ====================================
    reg_alloc<kRegsSmx>();
    const int t = warp >> 2;                  // Q tile
    const int r = threadIdx.x & 127;          // row inside the tile == TMEM lane
    const int tile_row0 = row0 + t * kBM;
    const int row = tile_row0 + r;
    const int nblk = t == 0 ? nblk0 : nblk1;
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_sel + t * kBN;  // this thread's scores; P goes over their head
    constexpr int kOCols = D;                            // columns of O this thread rescales / stores
    const uint32_t tO = tmem_base + lane_sel + 256 + t * D;
    const float c = p.scale_log2;
    const int shift = Sk - Sq;

    float m_ref = -INFINITY;
    float l = 0.f;
    // linear index of this row's first score in a [B, Hq, Sq, Sk] tensor, modulo 2^32: the dropout counter
    // (src/forward/kernel.py:147-149); drop_offset is what the caller adds for a sequence of a packed batch
    const uint32_t drop_row_off = p.drop_offset +
        ((static_cast<uint32_t>(batch) * p.Hq + head) * static_cast<uint32_t>(p.Sq) + static_cast<uint32_t>(row)) *
        static_cast<uint32_t>(p.Sk);
    (void)drop_row_off;

    // ---- kExt bits 0 + 2, a 16-bit bias TMA can address: the tile's bias passes through shared memory one
    // [128 rows x 128 keys] block at a time (two 128-byte-swizzled sub-tiles of 64 keys, one 32 KB buffer per Q tile).
    // Block i + 1 is requested by thread 0 of the tile as soon as all 128 rows of block i sit in registers (bias_empty),
    // i.e. before the scores of block i have even arrived, and travels during that block's softmax. (With 16 KB per tile
    // - half a block in flight - every block exposed one TMA round trip: 1.17 ms instead of the plain forward's 0.57.)
    // A thread's row is 2 x 128 contiguous bytes whose 16-byte pieces sit at (piece ^ (row & 7)): the 8 rows of a
    // quarter warp cover all 32 banks, no conflicts.
    [[maybe_unused]] uint8_t* my_bias = sBias + t * L::kBiasBufBytes;
    [[maybe_unused]] const int bias_h = head * p.bias_head_mul, bias_b = batch * p.bias_batch_mul;
    [[maybe_unused]] auto bias_issue = [&](int blk) {
      mbar_arrive_expect_tx(&bias_full[t], 2 * L::kBiasSubBytes);
      for (int j = 0; j < 2; ++j)
        tma_load_4d(my_bias + j * L::kBiasSubBytes, &tmap_bias, &bias_full[t], blk * kBN + j * 64, tile_row0, bias_h, bias_b);
    };
    if constexpr ((kExt & 1) != 0) {
      if (kBiasTma && nblk > 0 && r == 0) {
        tma_prefetch_desc(&tmap_bias);
        bias_issue(0);
      }
    }

    for (int i = 0; i < nblk; ++i) {
      // ---- kExt: additive bias (src/forward/compute_row_blocks.py:48-54, 68-69) ----
      // It joins the raw scores as bias / softmax_scale, so that the common factor c = softmax_scale * log2(e) applies
      // to both; hidden entries are set to -inf afterwards. A thread owns a score row, so reading its 128 bias values
      // straight from global memory makes every warp load touch 32 different lines (measured: forward 7.5x slower),
      // and loading piecewise after the scores have arrived exposes the L2 latency 16 times per block (5x slower).
      // So: (1) BEFORE waiting for S, each warp fetches the block's bias for its 32 rows with coalesced 16-byte loads
      // (8 rows x 64 B per instruction) into registers - the score registers are not live yet and the latency hides
      // behind the Q K^T product; (2) once S is there, 64-byte pieces of the rows go through a warp-private staging
      // tile in shared memory, every lane reads its own row back, adds it to its scores IN TENSOR MEMORY and the
      // usual path below takes over. (Everything unrolled at once made the kernel 17 000 instructions long and
      // instruction-cache bound; the loops below are sized to stay at ~5 000.)
      uint4 breg[16];
      bool bias_fast = false;
      [[maybe_unused]] const int bias_col0 = i * kBN;
      [[maybe_unused]] const int wrow0 = tile_row0 + (warp & 3) * 32;  // first query row of this warp
      [[maybe_unused]] auto bias_fetch = [&](auto elem_tag, int group) {
        constexpr int kElem = decltype(elem_tag)::value;  // bytes per bias element: 2 or 4
        constexpr int kChunkCols = 64 / kElem;            // columns per 64-byte piece of a row
        constexpr int kPer16 = 16 / kElem;                // elements per 16-byte load
        const uint8_t* base = static_cast<const uint8_t*>(p.bias.ptr);
        const int64_t head_off = batch * p.bias.sb + head * p.bias.sh;
        const int gc0 = bias_col0 + group * 4 * kChunkCols;
        // whole pieces in bounds and 16-byte aligned for every row of the warp (warp-uniform)
        bias_fast = (wrow0 + 32 <= Sq) && (bias_col0 + kBN <= Sk) && ((p.bias.sm * kElem) % 16 == 0) &&
                    ((reinterpret_cast<uintptr_t>(base) + (head_off + bias_col0) * kElem) % 16 == 0);
#pragma unroll
        for (int li = 0; li < 16; ++li) {
          const int ch = li / 4, rr = (li % 4) * 8 + (lane >> 2), piece = lane & 3;
          const int grow = wrow0 + rr;
          const int gcol = gc0 + ch * kChunkCols + piece * kPer16;
          const int64_t idx = head_off + static_cast<int64_t>(grow) * p.bias.sm + gcol;
          if (bias_fast) {
            breg[li] = __ldg(reinterpret_cast<const uint4*>(base + idx * kElem));
          } else {
            uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int e = 0; e < kPer16; ++e) {
              const bool ok = grow < Sq && gcol + e < Sk;
              if constexpr (kElem == 4) w[e] = ok ? __ldg(reinterpret_cast<const uint32_t*>(base) + idx + e) : 0u;
              else w[e / 2] |= (ok ? static_cast<uint32_t>(__ldg(reinterpret_cast<const uint16_t*>(base) + idx + e)) : 0u) << (16 * (e % 2));
            }
            breg[li] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      };
      [[maybe_unused]] auto bias_apply = [&](auto elem_tag, int group) {
        constexpr int kElem = decltype(elem_tag)::value;
        constexpr int kChunkCols = 64 / kElem;
        uint8_t* stage = smem + L::kStageOff + warp * L::kStageWarpBytes;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
          for (int li = 0; li < 4; ++li)
            *reinterpret_cast<uint4*>(stage + (li * 8 + (lane >> 2)) * L::kStageRow + (lane & 3) * 16) = breg[ch * 4 + li];
          __syncwarp();
          uint4 mine[4];
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) mine[k4] = *reinterpret_cast<const uint4*>(stage + lane * L::kStageRow + k4 * 16);
          __syncwarp();
          const uint32_t tcol = tS + (group * 4 + ch) * kChunkCols;
          uint32_t sv[kChunkCols];
          if constexpr (kChunkCols == 32) tmem_ld_32x32b_x32(tcol, sv); else tmem_ld_32x32b_x16(tcol, sv);
          tmem_wait_ld();
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint32_t w[4] = {mine[k4].x, mine[k4].y, mine[k4].z, mine[k4].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if constexpr (kElem == 4) {
                const int j = k4 * 4 + e;
                sv[j] = __float_as_uint(fmaf(__uint_as_float(w[e]), p.bias_mul, __uint_as_float(sv[j])));
              } else {
                const int j = k4 * 8 + 2 * e;
                sv[j] = __float_as_uint(fmaf(bias_to_float<kBf16>(static_cast<uint16_t>(w[e] & 0xffffu)), p.bias_mul, __uint_as_float(sv[j])));
                sv[j + 1] = __float_as_uint(fmaf(bias_to_float<kBf16>(static_cast<uint16_t>(w[e] >> 16)), p.bias_mul, __uint_as_float(sv[j + 1])));
              }
            }
          }
          if constexpr (kChunkCols == 32) tmem_st_32x32b_x32(tcol, sv); else tmem_st_32x32b_x16(tcol, sv);
        }
      };
      // TMA-staged bias: copy this thread's 128 values of block i out of the tile's buffer and hand it back
      uint4 bhalf[8], bhalf1[8];
      [[maybe_unused]] auto bias_take = [&]() {
        mbar_wait(&bias_full[t], i & 1);
        const uint32_t rowp = smem_u32(my_bias) + r * 128;
#pragma unroll
        for (int c8 = 0; c8 < 16; ++c8) {
          uint4& d = c8 < 8 ? bhalf[c8] : bhalf1[c8 - 8];
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w)
                       : "r"(rowp + (c8 / 8) * L::kBiasSubBytes + (((c8 % 8) ^ (r & 7)) << 4)) : "memory");
        }
        // The buffer is overwritten by TMA once all 128 rows have arrived, so this thread's loads must have RETURNED by
        // then - mbarrier.arrive alone does not wait for them (measured: without this a row now and then saw the next
        // tile, 1 run in 3 differed in a few rows at B=3, H=4, S=640). The arrive's address is made to depend on the
        // loaded registers (a term that is always 0, which the compiler cannot know); fence.proxy.async before the
        // arrive works as well.
        uint32_t dep = 0;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) dep ^= bhalf[c8].x ^ bhalf1[c8].x;
        dep &= static_cast<uint32_t>(p.bias_head_mul) >> 1;  // bias_head_mul is 0 or 1
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bias_empty[t]) + dep) : "memory");
        if (r == 0 && i + 1 < nblk) {
          mbar_wait(&bias_empty[t], i & 1);
          bias_issue(i + 1);
        }
      };
      if constexpr ((kExt & 1) != 0) {
        if (p.bias.ptr != nullptr) {
          if constexpr (kBiasTma) bias_take();
          else if (p.bias.is_fp32) bias_fetch(std::integral_constant<int, 4>{}, 0);
          else bias_fetch(std::integral_constant<int, 2>{}, 0);
        }
      }
      // ---- kExt bit 1: dropout. One keep bit per score of this thread's row and block, drawn here, before the scores
      // arrive: the registers are free and the Philox rounds fill the wait for the Q K^T product. They stay in a
      // rolled loop (see the note on code size above).
      uint32_t keep_bits[kCols / 32];
      if constexpr ((kExt & 2) != 0) {
#pragma unroll
        for (int wd = 0; wd < kCols / 32; ++wd) {
          uint32_t bits = 0xffffffffu;
          if (p.dropout_p > 0.f) {
            bits = 0u;
            const uint32_t off = drop_row_off + static_cast<uint32_t>(i * kBN + wd * 32);
#pragma unroll 2
            for (int b = 0; b < 32; ++b)
              bits |= static_cast<uint32_t>(dropout_keep(p.seed_lo, p.seed_hi, off + b, p.dropout_p)) << b;
          }
          keep_bits[wd] = bits;
        }
      }
      (void)keep_bits;
      mbar_wait(&s_full[t], i & 1);
      tcgen05_fence_after();
      if constexpr ((kExt & 1) != 0) {
        if (p.bias.ptr != nullptr) {
          if constexpr (kBiasTma) {
            // added to the score registers below, without the round trip through tensor memory
          } else if (p.bias.is_fp32) {  // 64 columns per pass: the second fetch is exposed (fp32 biases are the rare case)
            bias_apply(std::integral_constant<int, 4>{}, 0);
            bias_fetch(std::integral_constant<int, 4>{}, 1);
            bias_apply(std::integral_constant<int, 4>{}, 1);
          } else {
            bias_apply(std::integral_constant<int, 2>{}, 0);
          }
          tmem_wait_st();
        }
      }

      float s[kCols];  // this thread's part of the score row
      auto load_scores = [&]() {
        uint32_t su[kCols / 32][32];
#pragma unroll
        for (int cidx = 0; cidx < kCols / 32; ++cidx) tmem_ld_32x32b_x32(tS + cidx * 32, su[cidx]);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < kCols; ++j) s[j] = __uint_as_float(su[j / 32][j % 32]);
      };
      bool scores_loaded = false;
      if constexpr ((kExt & 1) != 0) {
        if constexpr (kBiasTma) {  // 64 scores at a time, so that at most 32 bias registers are live next to all 128 scores
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t su[2][32];
            tmem_ld_32x32b_x32(tS + h * 64, su[0]);
            tmem_ld_32x32b_x32(tS + h * 64 + 32, su[1]);
            tmem_wait_ld();
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
              const uint4 q4 = h == 0 ? bhalf[c8] : bhalf1[c8];
              const uint32_t w[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int jj = c8 * 8 + 2 * e;
                s[h * 64 + jj] = fmaf(bias_to_float<kBf16>(static_cast<uint16_t>(w[e] & 0xffffu)), p.bias_mul, __uint_as_float(su[jj / 32][jj % 32]));
                s[h * 64 + jj + 1] = fmaf(bias_to_float<kBf16>(static_cast<uint16_t>(w[e] >> 16)), p.bias_mul, __uint_as_float(su[(jj + 1) / 32][(jj + 1) % 32]));
              }
            }
          }
          scores_loaded = true;
        }
      }
      if (!scores_loaded) load_scores();

      const int col0 = i * kBN;
      const int block_last = i * kBN + kBN - 1;
      const bool need_mask = (block_last >= Sk) || (p.causal && block_last > tile_row0 + shift);
      if (need_mask) {
        const int limit = p.causal ? min(Sk - 1, row + shift) : (Sk - 1);
#pragma unroll
        for (int j = 0; j < kCols; ++j) s[j] = (col0 + j <= limit) ? s[j] : -INFINITY;
      }

      uint32_t pk[kCols / 2];
      float mx[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) mx[j] = fmaxf(s[j], s[j + 8]);
#pragma unroll
      for (int j = 16; j < kCols; j += 16) {
#pragma unroll
        for (int e = 0; e < 8; ++e) mx[e] = fmaxf(mx[e], fmaxf(s[j + e], s[j + e + 8]));
      }
      const float mb = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7]))) * c;

      // lazy rescale: only move the reference max when the block max exceeds it by > 2^8
      const bool need = mb > m_ref + kRescaleThreshold;
      float alpha = 1.f;
      if (need) {
        alpha = ex2_approx(m_ref - mb);  // m_ref == -inf -> 0
        m_ref = mb;
      }
      if (i > 0 && __any_sync(0xffffffffu, need)) {
        // rare (the reference maximum moved by more than 2^8): 16 columns at a time keeps the live scores in registers
#pragma unroll 1
        for (int cidx = 0; cidx < kOCols / 16; ++cidx) {
          uint32_t o[16];
          tmem_ld_32x32b_x16(tO + cidx * 16, o);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
          tmem_st_32x32b_x16(tO + cidx * 16, o);
        }
      }
      l *= alpha;

      // every exponential on the MUFU: a bias may hold -inf (used as a mask), which must give exactly 0
      const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < kCols; j += 2) {
        float pr[2];
        pr[0] = ex2_approx(fmaf(s[j], c, neg_m));
        pr[1] = ex2_approx(fmaf(s[j + 1], c, neg_m));
        ls[j % 4] += pr[0];
        ls[j % 4 + 1] += pr[1];
        if constexpr ((kExt & 2) != 0) {
          // dropout hits P after its row sum was taken (src/forward/compute_row_blocks.py:73-79)
          pr[0] = ((keep_bits[j / 32] >> (j % 32)) & 1u) ? pr[0] : 0.f;
          pr[1] = ((keep_bits[(j + 1) / 32] >> ((j + 1) % 32)) & 1u) ? pr[1] : 0.f;
        }
        pk[j / 2] = pack2<kBf16>(pr[0], pr[1]);
      }
      l += (ls[0] + ls[2]) + (ls[1] + ls[3]);
#pragma unroll
      for (int cidx = 0; cidx < kCols / 64; ++cidx) tmem_st_32x32b_x32(tS + cidx * 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[cidx * 32]));
      tmem_wait_st();
      tcgen05_fence_before();
      mbar_arrive(&p_full[t]);
    }

    // ------------------------------ epilogue: O / l -> smem -> TMA store ---------------------------
    if (tile_row0 < Sq) {
      float inv_l = (l > 0.f) ? 1.0f / l : 0.f;
      if constexpr ((kExt & 2) != 0) {
        if (p.dropout_p > 0.f) inv_l *= p.inv_keep;  // src/forward/kernel.py:256-257
      }
      if (nblk > 0) {
        mbar_wait(&o_full[t], 0);
        tcgen05_fence_after();
      } else if (nblk_max > 0) {
        // fully masked tile next to a live one: its Q tile is still being loaded under the shared q_full
        // barrier; the zero rows staged below must not be overwritten by that load
        mbar_wait(q_full, 0);
      }
      uint8_t* sO = sQ + t * L::kTileBytes;  // the Q tile is dead once o_full fired
#pragma unroll
      for (int cidx = 0; cidx < kOCols / 32; ++cidx) {
        uint32_t o[32];
        if (nblk > 0) {
          tmem_ld_32x32b_x32(tO + cidx * 32, o);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = 0u;
        }
        const int col = cidx * 32;  // first head-dim column of this chunk
        uint8_t* sub = sO + (col / 64) * kSubBytes;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 v;
          v.x = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 0]) * inv_l, __uint_as_float(o[q4 * 8 + 1]) * inv_l);
          v.y = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 2]) * inv_l, __uint_as_float(o[q4 * 8 + 3]) * inv_l);
          v.z = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 4]) * inv_l, __uint_as_float(o[q4 * 8 + 5]) * inv_l);
          v.w = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 6]) * inv_l, __uint_as_float(o[q4 * 8 + 7]) * inv_l);
          const uint32_t chunk16 = ((col % 64) / 8) + q4;
          *reinterpret_cast<uint4*>(sub + sw128_offset(r, chunk16)) = v;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + t, kTileSmxThreads);
      if (r == 0) {
        if (p.seqlens != nullptr) tmap_acquire(mo);
        for (int j = 0; j < kSub; ++j) tma_store_4d(mo, sO + j * kSubBytes, j * 64, tile_row0, head, bc);
        bulk_commit_group();
        bulk_wait_group_read<0>();
      }
      // LSE (log2 domain), padded rows get 0
      float* lse_row = p.lse + (static_cast<int64_t>(batch) * p.Hq + head) * p.lse_stride;
      if (row < p.lse_stride) {
        float lse = 0.f;
        if (row < Sq) lse = (l > 0.f) ? m_ref + lg2_approx(l) : -INFINITY;
        lse_row[row] = lse;
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

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
template <int D, bool kBf16, int kExt>
static cudaError_t launch_fwd_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                                const CUtensorMap& to, const FwdParams& p, cudaStream_t stream) {
  constexpr int smem = FwdSmem<D, (kExt & 1) != 0, (kExt & 4) != 0>::kDynBytes;
  if (cudaError_t e = ensure_dynamic_smem<fa2_fwd_kernel<D, kBf16, kExt>>(smem); e != cudaSuccess) return e;
  dim3 grid(p.n_pairs, p.Hq, p.B);
  const CUtensorMap& tbias = (kExt & 4) ? *p.bias_map : tq;  // placeholder when no bias goes through TMA
  FwdParams pp = p;
  // a head's CTAs stream its K and V - and, with a per-head bias, its [Sq x Sk] bias: no block-major tail then
  pp.tail_log2 = (p.bias.ptr != nullptr && p.bias.sh != 0) ? 0 : tail_log2_for(tail_heads_for(static_cast<size_t>(p.Sk) * D * 2 * 2, p.causal), p.Hq);
  fa2_fwd_kernel<D, kBf16, kExt><<<grid, kFwdThreads, smem, stream>>>(tq, tk, tv, to, tbias, pp);
  return cudaGetLastError();
}

// D: the kernel head dim (64, 128 or 256). The plain forward and every head dim above 128 go to fwd2_sm100.cu.
cudaError_t launch_fwd(int D, bool bf16, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                       const CUtensorMap& to, const FwdParams& p, cudaStream_t stream) {
  const int ext = (p.bias.ptr != nullptr ? (p.bias_tma && p.bias_map ? 5 : 1) : 0) | (p.dropout_p > 0.f ? 2 : 0);
  if (ext == 0 || D == 256) return launch_fwd2(D, bf16, tq, tk, tv, to, p, stream);
#define FA2_FWD_CASE(d, b)                                                                       \
  if (D == d && bf16 == b) {                                                                     \
    switch (ext) {                                                                               \
      case 1: return launch_fwd_t<d, b, 1>(tq, tk, tv, to, p, stream);                           \
      case 2: return launch_fwd_t<d, b, 2>(tq, tk, tv, to, p, stream);                           \
      case 3: return launch_fwd_t<d, b, 3>(tq, tk, tv, to, p, stream);                           \
      case 5: return launch_fwd_t<d, b, 5>(tq, tk, tv, to, p, stream);                           \
      default: return launch_fwd_t<d, b, 7>(tq, tk, tv, to, p, stream);                          \
    }                                                                                            \
  }
  FA2_FWD_CASE(128, true) FA2_FWD_CASE(128, false) FA2_FWD_CASE(64, true) FA2_FWD_CASE(64, false)
#undef FA2_FWD_CASE
  return cudaErrorInvalidValue;
}

}  // namespace fa2
