// FlashAttention-2 forward for sm_100a (B200): TMA -> smem ring -> tcgen05.mma (S = Q K^T, O += P V) with
// accumulators in tensor memory, two 128-row Q tiles per CTA whose softmax warpgroups ping-pong against
// the single MMA-issuing warp.
//
// Replaces the reference's `_fwd_kernel` (src/forward/kernel.py:61-291) + `compute_row_block`
// (src/forward/compute_row_blocks.py:8-103). Same arithmetic contract:
//   S = Q K^T (fp32 acc); masked where col >= Sk or (causal) col > row + Sk - Sq;
//   log2-domain online softmax with c = softmax_scale * log2(e); P rounded to the input dtype before PV;
//   O = (P V) / l in the input dtype;  LSE2 = m + log2(l) in fp32 (src/forward/kernel.py:119,259-272).
//
// CTA = 2 * kHalves softmax warpgroups + one issue warpgroup (kHalves = FA2_FWD_HALVES, 1 or 2):
//   softmax warpgroup (t, h): Q tile t, key columns [h * 128 / kHalves, ...) of every block; thread = query row
//                             (TMEM lane == row). With kHalves = 2 the two threads of a row exchange their block
//                             maxima through shared memory (one named barrier per block) and keep partial row sums.
//                             Four softmax warps per scheduler instead of two: the row's 128 exponentials are the
//                             dependent chain softmax -> PV -> QK(next) of a tile is waiting for, and a lone warp
//                             only keeps the MUFU ~2/3 busy (measured T_softmax 2100 cycles -> see DESIGN.md).
//   issue warpgroup: warp 0 MMA issuer (one elected lane), warp 1 TMA producer (one elected lane), 2 idle warps
// TMEM (512 columns): S0 [0,128) | S1 [128,256) | O0 [256,256+D) | O1 [256+D,256+2D).
//   P_t (16-bit) overwrites the head of the S_t columns of the thread that wrote it, once it has read its row.
#include <type_traits>

#include "fa2_kernels.cuh"
#include "ptx.cuh"

namespace fa2 {

#ifdef FA2_TRACE
// Debug build only (make TRACE=1): clock64() stamps of one CTA's pipeline events, [role][iteration][event].
__device__ long long g_fwd_trace[3 * 64 * 8];
#define FA2_STAMP(role, it, ev)                                                                          \
  do {                                                                                                   \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (it) < 64)                              \
      g_fwd_trace[((role) * 64 + (it)) * 8 + (ev)] = clock64();                                          \
  } while (0)
#else
#define FA2_STAMP(role, it, ev) do { } while (0)
#endif

namespace {

constexpr int kBM = 128;  // query rows per tile (= TMEM lanes)
constexpr int kBN = 128;  // key rows per block
// Measured on B200 (B=4, H=32, S=4096, D=128 causal; same box, same call): 1 thread per row 0.566 ms, + speculative
// exponentials 0.567, 2 threads per row 0.574, 2 threads + speculative 0.643, + a quarter of the exponentials on the FMA
// pipes 0.637. Shortening the softmax chain does not move the kernel: with ~1 kW drawn it runs at the power cap (SM clock
// 1.55-1.75 GHz instead of 1.965), so what counts is work per step, not bubbles - the leanest variant is the default.
#ifndef FA2_FWD_HALVES
#define FA2_FWD_HALVES 1
#endif
constexpr int kHalves = FA2_FWD_HALVES;          // softmax threads per query row
constexpr int kCols = kBN / kHalves;             // key columns per softmax thread
constexpr int kSmxThreads = 2 * kHalves * 128;   // softmax threads per CTA
constexpr int kTileSmxThreads = kHalves * 128;   // ... per Q tile
constexpr int kFwdThreads = kSmxThreads + 128;
constexpr int kMmaWarp = kSmxThreads / 32, kTmaWarp = kMmaWarp + 1;
// register pools (setmaxnreg): 384 threads x 168 = 64512 = 2 x 128 x 208 + 128 x 88
//                              640 threads x  96 = 61440 = 4 x 128 x 104 + 128 x 64
constexpr int kRegsSmx = kHalves == 2 ? 104 : 208;
constexpr int kRegsIssue = kHalves == 2 ? 64 : 88;
constexpr float kRescaleThreshold = 8.0f;  // log2 units; P stays <= 2^8 between rescales
// Of every FA2_FWD_EMUL_DEN probabilities, FA2_FWD_EMUL_NUM are exponentiated on the FMA pipes (ex2_emul) and the rest
// on the MUFU. Measured issue costs per warp instruction and scheduler (scripts/micro/pipe_rates.cu): MUFU.EX2 8 cycles,
// FFMA / FADD / IMAD 1, FMNMX / F2FP 2 - an emulated exponential is 8 instructions, so about a quarter of them
// would balance the pipes on paper. Measured on B200 (B=4,H=32,S=4096,D=128 causal): 0/4 -> 0.567 ms, 1/4 -> 0.591,
// 1/3 -> 0.616, 1/2 -> 0.651: with one softmax warp per scheduler the phase is latency-bound, not MUFU-bound, and the
// extra dependent instructions cost more than the MUFU slots they free. Default: off.
#ifndef FA2_FWD_EMUL_NUM
#define FA2_FWD_EMUL_NUM (D == 64 ? 1 : 0)  // evaluated inside the kernel template
#endif
#ifndef FA2_FWD_EMUL_DEN
#define FA2_FWD_EMUL_DEN 4
#endif
// Packed flavour (whole pairs through ex2_emul2, f32x2 arithmetic). Measured, same box and call: D = 64 (S = 16384,
// non-causal, where the 16384 exponentials of a tile take twice as long on the MUFU as its two products on the tensor
// pipe): 1 of 4 pairs 3.15 -> 2.92 ms, 2 of 4 3.21; D = 128 (S = 4096 causal): 0.560 -> 0.566 / 0.561. Hence the
// default: a quarter of the pairs at D = 64, none at D = 128.
// Energy / time ablations for DESIGN.md (results are WRONG on purpose, never shipped): 1 = exponentials replaced by a
// move, 2 = additionally no maximum pass / rescale, 3 = no softmax arithmetic at all (P = rounded raw scores, no row sum)
#ifndef FA2_FWD_ABLATE
#define FA2_FWD_ABLATE 0
#endif
// Experiment: both exponentials of a pair in one MUFU.EX2.F16x2 (the arguments and results pass through fp16)
#ifndef FA2_FWD_EX2_F16
#define FA2_FWD_EX2_F16 0
#endif
#ifndef FA2_FWD_EMUL_PACKED
#define FA2_FWD_EMUL_PACKED 1
#endif
// Speculative exponentials against the previous reference maximum (see the softmax loop). Off: no gain (see above).
#ifndef FA2_FWD_SPEC
#define FA2_FWD_SPEC 0
#endif

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
  static constexpr int kBarOff = kBiasOff + ((kExt && kHalves == 1) ? 2 * kBiasBufBytes : 0);
  // barriers: q_full, k_full[S], k_empty[S], v_full[S], v_empty[S], s_full[2], p_full[2], o_full[2], bias_full[2], bias_empty[2]
  static constexpr int kNumBars = 1 + 4 * kStages + 6 + 4;
  static constexpr int kTmemSlotOff = kBarOff + kNumBars * 8;
  // row statistics exchanged between the kHalves threads of a row: [parity 2 x kind 2][tile 2][half 2][128 rows] fp32
  static constexpr int kXchgOff = kTmemSlotOff + 16;
  static constexpr int kBytes = kXchgOff + (kHalves == 2 ? 4 * 2 * 2 * kBM * 4 : 0);
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
  float* xchg = reinterpret_cast<float*>(smem + L::kXchgOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // heaviest (largest row index) tile pairs first: better tail behaviour under causal masking
  const int pair = p.n_pairs - 1 - static_cast<int>(blockIdx.x);
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
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
            FA2_STAMP(2, i, t * 4 + 0);
            issue_pv(t, sv, i > 0);
            FA2_STAMP(2, i, t * 4 + 1);
            if (i + 1 < nblk[t]) {
              if (!next_k_ready) {
                mbar_wait(&k_full[sn], phn);
                tcgen05_fence_after();
                next_k_ready = true;
              }
              FA2_STAMP(2, i, t * 4 + 2);
              issue_qk(t, sn);
              umma_commit(&s_full[t]);
              FA2_STAMP(2, i, t * 4 + 3);
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
    const int t = warp / (4 * kHalves);       // Q tile
    const int half = (warp >> 2) % kHalves;   // which kCols key columns of every block
    const int r = threadIdx.x & 127;          // row inside the tile == TMEM lane
    const int tile_row0 = row0 + t * kBM;
    const int row = tile_row0 + r;
    const int nblk = t == 0 ? nblk0 : nblk1;
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_sel + t * kBN + half * kCols;  // this thread's scores; P goes over their head
    constexpr int kOCols = D / kHalves;                                 // columns of O this thread rescales / stores
    const uint32_t tO = tmem_base + lane_sel + 256 + t * D + half * kOCols;
    const float c = p.scale_log2;
    const int shift = Sk - Sq;
    auto xchg_slot = [&](int par, int h, int kind = 0) { return xchg + (((par * 2 + kind) * 2 + t) * 2 + h) * kBM + r; };

    float m_ref = -INFINITY;
    float l = 0.f;  // kHalves == 2: the sum over this thread's columns only
    // linear index of this row's first score in a [B, Hq, Sq, Sk] tensor, modulo 2^32: the dropout counter
    // (src/forward/kernel.py:147-149)
    const uint32_t drop_row_off =
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
    if constexpr ((kExt & 1) && kHalves == 1) {
      if (kBiasTma && nblk > 0 && r == 0) {
        tma_prefetch_desc(&tmap_bias);
        bias_issue(0);
      }
    }

    for (int i = 0; i < nblk; ++i) {
      if (r == 0 && half == 0) FA2_STAMP(t, i, 0);
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
      if constexpr ((kExt & 1) && kHalves == 1) {
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
            const uint32_t off = drop_row_off + static_cast<uint32_t>(i * kBN + half * kCols + wd * 32);
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
      if (r == 0 && half == 0) FA2_STAMP(t, i, 1);
      if constexpr ((kExt & 1) != 0) {
        if (p.bias.ptr != nullptr) {
          if constexpr (kHalves == 1) {
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
          } else {
            __trap();  // the staged bias path assumes one softmax thread per row (FA2_FWD_HALVES = 1, the default)
          }
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
      if constexpr ((kExt & 1) && kHalves == 1) {
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
      if (r == 0 && half == 0) FA2_STAMP(t, i, 2);

      const int col0 = i * kBN + half * kCols;
      const int block_last = i * kBN + kBN - 1;
      const bool need_mask = (block_last >= Sk) || (p.causal && block_last > tile_row0 + shift);
      if (need_mask) {
        const int limit = p.causal ? min(Sk - 1, row + shift) : (Sk - 1);
#pragma unroll
        for (int j = 0; j < kCols; ++j) s[j] = (col0 + j <= limit) ? s[j] : -INFINITY;
      }

      uint32_t pk[kCols / 2];
      // Speculative fast path (every block but the first and the masked ones): exponentiate against the current
      // reference maximum right away and find the block maximum on the side - the FMNMX run on the ALU pipe in the
      // shadow of the MUFU. Only if some row of this warp turns out to exceed the reference by more than the rescale
      // threshold is the block redone the slow way (scores reloaded from tensor memory - P has not been stored yet -
      // maximum first, rescale O, exponentiate). Results are identical to always taking the slow path; the maximum
      // pass leaves the softmax -> PV -> QK chain.
      bool done = false;
      if (FA2_FWD_SPEC && !(kExt & 1) && i > 0 && !need_mask) {
        const float neg_m = -m_ref;  // m_ref == -inf (row without any visible key so far): x = +inf forces the slow path
        float ls[4] = {0.f, 0.f, 0.f, 0.f};
        float xm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < kCols; j += 2) {
          const float x0 = fmaf(s[j], c, neg_m), x1 = fmaf(s[j + 1], c, neg_m);
          xm[(j / 2) % 4] = fmaxf(xm[(j / 2) % 4], fmaxf(x0, x1));
          float p0, p1;
          if (FA2_FWD_EMUL_NUM > 0 && ((j / 2) % FA2_FWD_EMUL_DEN) < FA2_FWD_EMUL_NUM) {
            p0 = ex2_emul(x0);
            p1 = ex2_emul(x1);
          } else {
            p0 = ex2_approx(x0);
            p1 = ex2_approx(x1);
          }
          ls[j % 4] += p0;
          ls[j % 4 + 1] += p1;
          pk[j / 2] = pack2<kBf16>(p0, p1);
        }
        float xmax = fmaxf(fmaxf(xm[0], xm[1]), fmaxf(xm[2], xm[3]));
        if constexpr (kHalves == 2) {
          *xchg_slot(i & 1, half, 1) = xmax;
          named_bar_sync(1 + t, kTileSmxThreads);
          xmax = fmaxf(xmax, *xchg_slot(i & 1, half ^ 1, 1));
        }
        const bool redo = !(xmax <= kRescaleThreshold);
        if (!__any_sync(0xffffffffu, redo)) {
          l += (ls[0] + ls[2]) + (ls[1] + ls[3]);
          done = true;
        } else {
          load_scores();
        }
      }
      if (!done) {
#if FA2_FWD_ABLATE >= 3
      {
#pragma unroll
        for (int j = 0; j < kCols; j += 2) pk[j / 2] = pack2<kBf16>(s[j], s[j + 1]);
        l = 1.f; m_ref = 0.f; done = true;
      }
      if (!done) {
#endif
      float mx[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) mx[j] = fmaxf(s[j], s[j + 8]);
#pragma unroll
      for (int j = 16; j < kCols; j += 16) {
#pragma unroll
        for (int e = 0; e < 8; ++e) mx[e] = fmaxf(mx[e], fmaxf(s[j + e], s[j + e + 8]));
      }
      float mb = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7]))) * c;
#if FA2_FWD_ABLATE == 2
      mb = s[0] * c;
#endif
      if constexpr (kHalves == 2) {
        // the two threads of a row must use the same reference maximum: swap block maxima (double buffered by parity:
        // a slot is rewritten two blocks later, after the barrier of the block in between)
        *xchg_slot(i & 1, half) = mb;
        named_bar_sync(1 + t, kTileSmxThreads);
        mb = fmaxf(mb, *xchg_slot(i & 1, half ^ 1));
      }

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

      if (r == 0 && half == 0) FA2_STAMP(t, i, 3);
      const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      auto exp_block = [&](auto emulate) {
#pragma unroll
        for (int j = 0; j < kCols; j += 2) {
          float pr[2];
#if FA2_FWD_EMUL_PACKED
          // whole pairs on the FMA pipes with the f32x2 flavour (half the issue slots of the scalar one)
          if (decltype(emulate)::value && ((j / 2) % FA2_FWD_EMUL_DEN) < FA2_FWD_EMUL_NUM) {
            const float2 x2 = ffma2(make_float2(s[j], s[j + 1]), make_float2(c, c), make_float2(neg_m, neg_m));
            const float2 p2 = ex2_emul2(x2);
            pr[0] = p2.x;
            pr[1] = p2.y;
          } else {
#if FA2_FWD_ABLATE >= 1
            pr[0] = fmaf(s[j], c, neg_m);
            pr[1] = fmaf(s[j + 1], c, neg_m);
#elif FA2_FWD_EX2_F16
            {
              uint32_t xh, ph;
              asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(xh) : "f"(fmaf(s[j + 1], c, neg_m)), "f"(fmaf(s[j], c, neg_m)));
              asm("ex2.approx.f16x2 %0, %1;" : "=r"(ph) : "r"(xh));
              const float2 pf2 = __half22float2(*reinterpret_cast<const __half2*>(&ph));
              pr[0] = pf2.x;
              pr[1] = pf2.y;
            }
#else
            pr[0] = ex2_approx(fmaf(s[j], c, neg_m));
            pr[1] = ex2_approx(fmaf(s[j + 1], c, neg_m));
#endif
          }
          ls[j % 4] += pr[0];
          ls[j % 4 + 1] += pr[1];
          if constexpr ((kExt & 2) != 0) {
            // dropout hits P after its row sum was taken (src/forward/compute_row_blocks.py:73-79)
            pr[0] = ((keep_bits[j / 32] >> (j % 32)) & 1u) ? pr[0] : 0.f;
            pr[1] = ((keep_bits[(j + 1) / 32] >> ((j + 1) % 32)) & 1u) ? pr[1] : 0.f;
          }
#else
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float x = fmaf(s[j + e], c, neg_m);
            if (decltype(emulate)::value && ((j + e) % FA2_FWD_EMUL_DEN) < FA2_FWD_EMUL_NUM) {
              pr[e] = ex2_emul(x);
            } else {
              pr[e] = ex2_approx(x);
            }
            ls[(j + e) % 4] += pr[e];
          }
#endif
          pk[j / 2] = pack2<kBf16>(pr[0], pr[1]);
        }
      };
      // masked blocks keep every exponential on the MUFU (exact 0 for -inf scores); everywhere else a fixed share
      // is evaluated on the FMA pipes
      if (need_mask) exp_block(std::false_type{}); else exp_block(std::true_type{});
      l += (ls[0] + ls[2]) + (ls[1] + ls[3]);
      }
#if FA2_FWD_ABLATE >= 3
      }
#endif
      if (r == 0 && half == 0) FA2_STAMP(t, i, 4);
#pragma unroll
      for (int cidx = 0; cidx < kCols / 64; ++cidx) tmem_st_32x32b_x32(tS + cidx * 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[cidx * 32]));
      tmem_wait_st();
      tcgen05_fence_before();
      mbar_arrive(&p_full[t]);
      if (r == 0 && half == 0) FA2_STAMP(t, i, 5);
    }

    // ------------------------------ epilogue: O / l -> smem -> TMA store ---------------------------
    if (tile_row0 < Sq) {
      if constexpr (kHalves == 2) {
        // full row sum = the two partial sums (both relative to the same m_ref); slots of parity nblk & 1 are free:
        // the last block used parity (nblk - 1) & 1 and every thread has passed that block's barrier
        *xchg_slot(nblk & 1, half) = l;
        named_bar_sync(1 + t, kTileSmxThreads);
        l += *xchg_slot(nblk & 1, half ^ 1);
      }
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
        const int col = half * kOCols + cidx * 32;  // first head-dim column of this chunk
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
      if (r == 0 && half == 0) {
        if (p.seqlens != nullptr) tmap_acquire(mo);
        for (int j = 0; j < kSub; ++j) tma_store_4d(mo, sO + j * kSubBytes, j * 64, tile_row0, head, bc);
        bulk_commit_group();
        bulk_wait_group_read<0>();
      }
      // LSE (log2 domain), padded rows get 0
      float* lse_row = p.lse + (static_cast<int64_t>(batch) * p.Hq + head) * p.lse_stride;
      if (half == 0 && row < p.lse_stride) {
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
  auto kernel = fa2_fwd_kernel<D, kBf16, kExt>;
  constexpr int smem = FwdSmem<D, (kExt & 1) != 0, (kExt & 4) != 0>::kDynBytes;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  dim3 grid(p.n_pairs, p.Hq, p.B);
  const CUtensorMap& tbias = (kExt & 4) ? *p.bias_map : tq;  // placeholder when no bias goes through TMA
  kernel<<<grid, kFwdThreads, smem, stream>>>(tq, tk, tv, to, tbias, p);
  return cudaGetLastError();
}

#ifdef FA2_TRACE
cudaError_t fetch_fwd_trace(long long* host, size_t n) {
  return cudaMemcpyFromSymbol(host, g_fwd_trace, n * sizeof(long long));
}
#endif

cudaError_t launch_fwd(int D, bool bf16, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                       const CUtensorMap& to, const FwdParams& p, cudaStream_t stream) {
  const int ext = (p.bias.ptr != nullptr ? (p.bias_tma && p.bias_map && kHalves == 1 ? 5 : 1) : 0) | (p.dropout_p > 0.f ? 2 : 0);
#define FA2_FWD_CASE(d, b)                                                                       \
  if (D == d && bf16 == b) {                                                                     \
    switch (ext) {                                                                               \
      case 0: return launch_fwd_t<d, b, 0>(tq, tk, tv, to, p, stream);                           \
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
