// FlashAttention-2 forward for sm_100a (B200), second generation: 64-key half-blocks with the scores of a Q tile
// DOUBLE-BUFFERED inside its 128 tensor-memory columns, so that the chain  softmax -> P V -> Q K^T -> softmax  of a tile
// no longer runs in series.
//
// Replaces the reference's `_fwd_kernel` (src/forward/kernel.py:61-291) + `compute_row_block`
// (src/forward/compute_row_blocks.py:8-103). Same arithmetic contract:
//   S = Q K^T (fp32 acc); masked where col >= Sk or (causal) col > row + Sk - Sq;
//   log2-domain online softmax with c = softmax_scale * log2(e); P rounded to the input dtype before PV;
//   O = (P V) / l in the input dtype;  LSE2 = m + log2(l) in fp32 (src/forward/kernel.py:119,259-272).
//
// Why (round 1 measurements): with one 128-column S buffer per tile that P overwrites in place, the
// tensor pipe sat idle while a tile's softmax ran, because Q K^T of the next block could not start before P V of this one
// had read P: period = T_softmax + 1024 + hand-over latencies (~3500 cycles per 128 keys and tile pair against 2048 of
// tensor work; tensor pipe 49 % active). Here a tile's S columns hold TWO half-blocks of 64 keys:
//   buffer a = columns [0, 64), buffer b = columns [64, 128); P(j) (16-bit, 32 columns) overwrites the head of buffer j & 1.
// While the softmax warpgroup works on half-block j in one buffer, the tensor pipe runs  P V(j-1)  and  Q K^T(j+1)  in the
// other: the only serial dependency left is on the softmax warps themselves. Price: Q K^T is issued as N = 64 products,
// which re-read the Q operand from shared memory for every 64 keys (6 KB per 128 x 64 x 16 product in 32 tensor cycles =
// 192 B/clk against the 128 B/clk an SM's shared memory delivers): Q K^T costs 1.5x its tensor time. Per 128 keys and
// tile pair: 2 x (768 + 512) = 2560 cycles, shared-memory bound, instead of ~3500 chain bound.
//
// CTA = kTiles softmax warpgroups (one per 128-row Q tile; thread = query row = TMEM lane) + one issue warpgroup
//       (warp 0 MMA issuer of tile 0, warp 1 TMA producer, warp 2 MMA issuer of tile 1, 1 idle warp).
//   head dim <= 128: kTiles = 2, K / V tiles of 128 keys (two half-blocks), TMEM S0 S1 O0 O1.
//   head dim 129..256 (instantiation D = 256): kTiles = 1, K / V tiles of 64 keys, TMEM S [0,128) O [256,512).
// MMA order per tile t and half-block j:  P V_t(j)  Q K^T_t(j+2); the two tiles' streams come from two issuing threads.
#include <type_traits>
#include <utility>

#include "fa2_kernels.cuh"
#include "ptx.cuh"

namespace fa2 {

#ifdef FA2_TRACE
// Debug build only (make TRACE=1): clock64() stamps of one CTA's pipeline events, [role][half-block][event];
// roles 0 / 1 = softmax thread 0 of tile 0 / 1, roles 2 / 3 = MMA issuer of tile 0 / 1 (scripts/fwd2_trace.py)
__device__ long long g_fwd_trace[4 * 64 * 8];
#define FA2_STAMP(role, it, ev)                                                                          \
  do {                                                                                                   \
    if (blockIdx.x == FA2_TRACE_CTA && blockIdx.y == 0 && blockIdx.z == 0 && (it) < 64)                  \
      g_fwd_trace[((role) * 64 + (it)) * 8 + (ev)] = clock64();                                          \
  } while (0)
#ifndef FA2_TRACE_CTA
#define FA2_TRACE_CTA 0
#endif
cudaError_t fetch_fwd_trace(long long* host, size_t n) {
  return cudaMemcpyFromSymbol(host, g_fwd_trace, n * sizeof(long long));
}
#else
#define FA2_STAMP(role, it, ev) do { } while (0)
#endif

namespace {

constexpr int kBM = 128;  // query rows per tile (= TMEM lanes)
constexpr int kHB = 64;   // keys per half-block
constexpr float kRescaleThreshold = 8.0f;  // log2 units; P stays <= 2^8 between rescales

template <int D>
struct Fwd2Cfg {
  static_assert(D == 64 || D == 128 || D == 256, "instantiated head dims");
  static constexpr int kTiles = D <= 128 ? 2 : 1;
  static constexpr int kKvRows = D <= 128 ? 128 : 64;  // keys per K / V tile (TMA box rows)
  static constexpr int kHpt = kKvRows / kHB;           // half-blocks per K / V tile
  static constexpr int kStages = D == 64 ? 4 : 2;      // K ring and V ring
  static constexpr int kSub = D / 64;                  // 64-column (128-byte) sub-tiles per row
  static constexpr int kQSubBytes = kBM * 128;
  static constexpr int kKvSubBytes = kKvRows * 128;
  static constexpr int kQTileBytes = kBM * D * 2;
  static constexpr int kKvTileBytes = kKvRows * D * 2;
  static constexpr int kQOff = 0;
  static constexpr int kKOff = kQOff + kTiles * kQTileBytes;
  static constexpr int kVOff = kKOff + kStages * kKvTileBytes;
  static constexpr int kBarOff = kVOff + kStages * kKvTileBytes;
  // q_full, k_full[S], k_empty[S], v_full[S], v_empty[S], s_full[2][2], p_full[2][2], pv_done[2], o_full[2]
  static constexpr int kNumBars = 1 + 4 * kStages + 4 + 4 + 2 + 2;
  static constexpr int kTmemSlotOff = kBarOff + kNumBars * 8;
  static constexpr int kBytes = kTmemSlotOff + 16;
  static constexpr int kDynBytes = kBytes + 1024;  // slack for manual 1024-byte alignment
  static_assert(kDynBytes <= 227 * 1024, "forward shared memory exceeds the 227 KB a CTA can have");
  static constexpr int kThreads = kTiles * 128 + 128;
  static constexpr int kMmaWarp = kTiles * 4, kTmaWarp = kMmaWarp + 1, kMmaWarp2 = kMmaWarp + 2;  // issuer of tile 0, producer, issuer of tile 1
  // half-blocks after which ring stage, K / V tile half and S buffer repeat: the MMA issuers unroll that window
  static constexpr int kWin = (kStages * kHpt) % 2 == 0 ? kStages * kHpt : 2 * kStages * kHpt;
  // register pools (setmaxnreg), two tiles: 384 threads x 168 = 64512 = 2 x 128 x 208 + 128 x 88
  static constexpr int kRegsSmx = 208, kRegsIssue = 88;
  // of every 4 pairs of probabilities this many are exponentiated on the FMA pipes (ex2_emul2) instead of the MUFU:
  // at head dim 64 the exponentials of a half-block take twice the MUFU time its products take on the tensor pipe
#ifdef FA2_FWD2_EMUL
  static constexpr int kEmul = FA2_FWD2_EMUL;
#else
  static constexpr int kEmul = D == 64 ? 1 : 0;
#endif
  // exponentials formed with the current reference maximum while the block maximum is still being taken (softmax loop)
  static constexpr bool kSpeculate = D == 64;
};

// number of 64-key half-blocks query tile [row0, row0 + 128) has to visit
__device__ __forceinline__ int num_half_blocks(int row0, int Sq, int Sk, int causal) {
  if (row0 >= Sq) return 0;
  const int last_row = min(row0 + kBM, Sq) - 1;
  const int cols = causal ? min(Sk, last_row + Sk - Sq + 1) : Sk;
  return cols <= 0 ? 0 : (cols + kHB - 1) / kHB;
}

// f(integral_constant<0>), f(integral_constant<1>), ...: an unrolled loop whose index is a compile-time constant
template <typename F, int... U>
__device__ __forceinline__ void static_for(F&& f, std::integer_sequence<int, U...>) {
  (f(std::integral_constant<int, U>{}), ...);
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

}  // namespace

// kBias: additive bias read straight from global memory (the head-dim 129..256 instantiation only; head dims up to 128
// keep the TMA-staged bias variants of fwd_sm100.cu)
template <int D, bool kBf16, bool kBias>
__global__ void __launch_bounds__(Fwd2Cfg<D>::kThreads, 1)
fa2_fwd2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
                const FwdParams p) {
  using L = Fwd2Cfg<D>;
  constexpr int kTiles = L::kTiles, kStages = L::kStages, kSub = L::kSub, kHpt = L::kHpt;
  constexpr uint32_t kTmemO = 256;

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
  uint64_t* s_full = v_empty + kStages;  // [tile][buffer]
  uint64_t* p_full = s_full + 4;         // [tile][buffer]
  uint64_t* pv_done = p_full + 4;        // [tile]: completion n = P V(n) of that tile has retired
  uint64_t* o_full = pv_done + 2;        // [tile]: the last P V of that tile has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlotOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // heaviest (largest row index) CTAs first, the last heads block-major: better tail behaviour under causal masking
  // (fa2_kernels.cuh: tail_balanced_cta)
  const int cta_rows = kTiles * kBM;
  const ItemCoord cta = tail_balanced_cta(p.tail_log2);
  const int slot = p.n_pairs - 1 - cta.x;
  const int head = cta.head;
  const int batch = cta.batch;
  const int head_kv = head / (p.Hq / p.Hkv);
  const int row0 = slot * cta_rows;
  // variable-length mode: this batch element's own length and tensor maps (bounded by it; batch coordinate 0)
  int Sq = p.Sq, Sk = p.Sk, bc = batch;
  const CUtensorMap *mq = &tmap_q, *mk = &tmap_k, *mv = &tmap_v, *mo = &tmap_o;
  if (p.seqlens != nullptr) {
    Sq = Sk = min(max(p.seqlens[batch], 0), p.Sq);  // clamped like the tensor maps (varlen_sm100.cu)
    const CUtensorMap* m = p.vmaps + static_cast<size_t>(batch) * kVarlenMapsPerBatch;
    mq = m; mk = m + 1; mv = m + 2; mo = m + 3; bc = 0;
    if (row0 >= Sq) return;  // CTA in the padding: its rows of O and LSE were zero-filled by the caller
  }

  int nh[2] = {num_half_blocks(row0, Sq, Sk, p.causal), 0};
  if constexpr (kTiles == 2) nh[1] = num_half_blocks(row0 + kBM, Sq, Sk, p.causal);
  const int nh_max = max(nh[0], nh[1]);
  const int n_kv_tiles = (nh_max + kHpt - 1) / kHpt;

  if (warp == L::kTmaWarp && lane == 0) {
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
  if (warp == L::kMmaWarp) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int s = 0; s < kStages; ++s) {
        mbar_init(&k_full[s], 1);
        mbar_init(&k_empty[s], kTiles);  // one commit per MMA issuer
        mbar_init(&v_full[s], 1);
        mbar_init(&v_empty[s], kTiles);
      }
      for (int i = 0; i < 4; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&p_full[i], kBM);
      }
      for (int t = 0; t < 2; ++t) {
        mbar_init(&pv_done[t], 1);
        mbar_init(&o_full[t], 1);
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
  // checking it) keeps every tcgen05 address provably warp-uniform: ptxas then feeds UTCHMMA / LDTM / STTM from uniform
  // registers directly instead of wrapping each one in an ELECT + R2UR.BROADCAST "waterfall".
  if (*tmem_slot != 0u) __trap();
  // PDL: everything above overlapped the tail of the previous kernel of the stream; nothing below may run before it has
  // completed. The next kernel may be scheduled as soon as every CTA of this grid got here.
  pdl_wait();
  if (threadIdx.x == 0) pdl_launch_dependents();

  if (warp >= L::kMmaWarp) {
    if constexpr (kTiles == 2) reg_dealloc<L::kRegsIssue>();
    if (warp == L::kTmaWarp) {
      // ======================================= TMA producer =======================================
      if (nh_max > 0 && elect_one()) {
        const int n_q_tiles = (kTiles == 2 && row0 + kBM < Sq) ? 2 : 1;
        mbar_arrive_expect_tx(q_full, n_q_tiles * L::kQTileBytes);
        for (int t = 0; t < n_q_tiles; ++t)
          for (int j = 0; j < kSub; ++j)
            tma_load_4d(sQ + t * L::kQTileBytes + j * L::kQSubBytes, mq, q_full, j * 64, row0 + t * kBM, head, bc);
        for (int i = 0; i < n_kv_tiles; ++i) {
          const int s = i % kStages;
          const uint32_t ph = (i / kStages) & 1;
          mbar_wait(&k_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[s], L::kKvTileBytes);
          for (int j = 0; j < kSub; ++j)
            tma_load_4d(sK + s * L::kKvTileBytes + j * L::kKvSubBytes, mk, &k_full[s], j * 64, i * L::kKvRows, head_kv, bc);
          mbar_wait(&v_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[s], L::kKvTileBytes);
          for (int j = 0; j < kSub; ++j)
            tma_load_4d(sV + s * L::kKvTileBytes + j * L::kKvSubBytes, mv, &v_full[s], j * 64, i * L::kKvRows, head_kv, bc);
        }
      }
    } else if (warp == L::kMmaWarp || (kTiles == 2 && warp == L::kMmaWarp2)) {
      // ======================================== MMA issuers ========================================
      // One issuing thread PER Q TILE (warps kMmaWarp and kMmaWarp2). ncu of the single-issuer version of this kernel
      // (profiles/r02a_fwd2_*): the issuing warp was busy 90 % of the time - ~330 instructions per half-block step,
      // 5.5 cycles each on the uniform datapath, head-of-line blocked on tile 0's P while tile 1's was ready - and the
      // softmax warps waited for S 26 % of theirs. Two independent issuers halve that stream; the window of kWin
      // half-blocks is unrolled so that every shared-memory descriptor offset and tensor-memory address of a product is
      // a compile-time constant (ring stage, half of the K / V tile, S buffer), and the full-head-dim case (no
      // `k < d_steps` predicate per product) has its own copy of the loop.
      if (nh_max > 0 && elect_one()) {
        const int t = (warp == L::kMmaWarp) ? 0 : 1;
        constexpr uint32_t idesc_qk = make_idesc_f16(kBM, kHB, kBf16, false, false);
        const uint32_t idesc_pv = make_idesc_f16(kBM, p.d_cols, kBf16, false, true);  // N = the tensors' head dim (x16)
        const int d_steps = p.d_steps;
        const int nht = nh[t];

        const SmemDesc q_desc = make_smem_desc_sw128(smem_u32(sQ) + t * L::kQTileBytes, 16, 1024);  // K-major
        const SmemDesc k_desc = make_smem_desc_sw128(smem_u32(sK), 16, 1024);                        // K-major
        const SmemDesc v_desc = make_smem_desc_sw128(smem_u32(sV), L::kKvSubBytes, 1024);            // MN-major
        const uint32_t tmem_s = t * 128, tmem_o = kTmemO + t * D;
        // S_t[buffer ujj & 1] = Q_t K(half-block jj)^T, ujj = jj modulo the unrolled window (compile-time)
        auto issue_qk = [&](auto ujj_tag, auto full_tag) {
          constexpr int ujj = decltype(ujj_tag)::value;
          constexpr int stage = (ujj / kHpt) % kStages, h = ujj % kHpt;
          const SmemDesc ka = advance(k_desc, stage * L::kKvTileBytes + h * (kHB * 128));
          const uint32_t d_tmem = tmem_s + (ujj & 1) * kHB;
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            if (decltype(full_tag)::value || k < d_steps)
              umma_ss(d_tmem, advance(q_desc, (k / 4) * L::kQSubBytes + (k % 4) * 32),
                      advance(ka, (k / 4) * L::kKvSubBytes + (k % 4) * 32), idesc_qk, k > 0 ? 1u : 0u);
          }
        };
        // O_t (+)= P_t(j) V(half-block j); the thread of a row packed its 64 probabilities into the first 32 columns
        // of buffer j & 1
        auto issue_pv = [&](auto u_tag, bool first) {
          constexpr int u = decltype(u_tag)::value;
          constexpr int stage = (u / kHpt) % kStages, h = u % kHpt;
          const SmemDesc va = advance(v_desc, stage * L::kKvTileBytes + h * (kHB / 16) * 2048);
          const uint32_t a_tmem = tmem_s + (u & 1) * kHB;
#pragma unroll
          for (int k = 0; k < kHB / 16; ++k)
            umma_ts(tmem_o, a_tmem + k * 8, advance(va, k * 2048), idesc_pv, (!first || k > 0) ? 1u : 0u);
        };
        int k_ready = -1, v_ready = -1;  // highest K / V tile whose arrival has been waited for
        auto need_k = [&](int i) {
          if (i > k_ready) {
            mbar_wait(&k_full[i % kStages], (i / kStages) & 1);
            tcgen05_fence_after();
            k_ready = i;
          }
        };
        auto need_v = [&](int i) {
          if (i > v_ready) {
            mbar_wait(&v_full[i % kStages], (i / kStages) & 1);
            tcgen05_fence_after();
            v_ready = i;
          }
        };
        auto last_use_of_tile = [&](int jj) { return (jj % kHpt == kHpt - 1) || jj == nh_max - 1; };
        // Both issuers release every ring stage (the `empty` barriers count kTiles arrivals), whether or not their own
        // tile used it: with tcgen05.commit after the products that read the stage, else with a plain arrive - but only
        // after having seen the stage FILLED, so that an issuer with nothing to do (a tile with fewer half-blocks, or one
        // beyond Sq) can never run a phase ahead of the other one and complete a barrier phase with two arrivals of its own.
        auto release_k = [&](int jlast) {
          const int i = jlast / kHpt;
          if (i * kHpt < nht) { umma_commit(&k_empty[i % kStages]); } else { need_k(i); mbar_arrive(&k_empty[i % kStages]); }
        };
        auto release_v = [&](int jlast) {
          const int i = jlast / kHpt;
          if (i * kHpt < nht) { umma_commit(&v_empty[i % kStages]); } else { need_v(i); mbar_arrive(&v_empty[i % kStages]); }
        };
        auto run = [&](auto full_tag) {
          mbar_wait(q_full, 0);
          tcgen05_fence_after();
          auto prologue_step = [&](auto ujj_tag) {
            constexpr int jj = decltype(ujj_tag)::value;
            if (jj < nh_max) {
              if (jj < nht) {
                need_k(jj / kHpt);
                issue_qk(ujj_tag, full_tag);
                umma_commit(&s_full[t * 2 + (jj & 1)]);
              }
              if (last_use_of_tile(jj)) release_k(jj);
            }
          };
          prologue_step(std::integral_constant<int, 0>{});
          prologue_step(std::integral_constant<int, 1>{});
          for (int j0 = 0; j0 < nh_max; j0 += L::kWin) {
            auto step = [&](auto u_tag) {
              constexpr int u = decltype(u_tag)::value;
              const int j = j0 + u, jj = j + 2;
              if (j >= nh_max) return;
              if (j < nht) {
                // K(jj) and V(j) are usually there long before P(j): take those waits off the P -> PV chain
                need_v(j / kHpt);
                if (jj < nht) need_k(jj / kHpt);
                FA2_STAMP(2 + t, j, 0);
                mbar_wait(&p_full[t * 2 + (u & 1)], (j >> 1) & 1);
                tcgen05_fence_after();
                FA2_STAMP(2 + t, j, 1);
                issue_pv(u_tag, j == 0);
                umma_commit(&pv_done[t]);
                if (j == nht - 1) umma_commit(&o_full[t]);
                if (jj < nht) {
                  issue_qk(std::integral_constant<int, (u + 2) % L::kWin>{}, full_tag);
                  umma_commit(&s_full[t * 2 + (u & 1)]);
                }
                FA2_STAMP(2 + t, j, 2);
              }
              if (last_use_of_tile(j)) release_v(j);
              if (jj < nh_max && last_use_of_tile(jj)) release_k(jj);
            };
            static_for(step, std::make_integer_sequence<int, L::kWin>{});
          }
        };
        if (d_steps == D / 16) run(std::true_type{}); else run(std::false_type{});
      }
    }
  } else {
    // ===================================== softmax warpgroups ====================================
    if constexpr (kTiles == 2) reg_alloc<L::kRegsSmx>();
    const int t = warp >> 2;          // Q tile
    const int r = threadIdx.x & 127;  // row inside the tile == TMEM lane
    const int tile_row0 = row0 + t * kBM;
    const int row = tile_row0 + r;
    const int nhb = nh[t];
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = lane_sel + t * 128;
    const uint32_t tO = lane_sel + kTmemO + t * D;
    const float c = p.scale_log2;
    const float2 c2 = make_float2(c, c);
    const int shift = Sk - Sq;

    float m_ref = -INFINITY;
    float l = 0.f;

    for (int j = 0; j < nhb; ++j) {
      const int buf = j & 1;
      const uint32_t tSb = tS + buf * kHB;
      const int col0 = j * kHB;
      // kBias: this row's 64 bias values of the half-block, requested before the scores are waited for
      [[maybe_unused]] float bias_v[kBias ? kHB : 1];
      if constexpr (kBias) {
        const bool row_ok = row < Sq;
        const int64_t off = batch * p.bias.sb + head * p.bias.sh + static_cast<int64_t>(row) * p.bias.sm + col0;
        if (p.bias.is_fp32) {
          const float* bp = static_cast<const float*>(p.bias.ptr) + off;
          const bool fast = row_ok && col0 + kHB <= Sk && (reinterpret_cast<uintptr_t>(bp) & 15) == 0;
          if (fast) {
#pragma unroll
            for (int q4 = 0; q4 < kHB / 4; ++q4) {
              const float4 v4 = __ldg(reinterpret_cast<const float4*>(bp) + q4);
              bias_v[q4 * 4 + 0] = v4.x; bias_v[q4 * 4 + 1] = v4.y; bias_v[q4 * 4 + 2] = v4.z; bias_v[q4 * 4 + 3] = v4.w;
            }
          } else {
#pragma unroll
            for (int e = 0; e < kHB; ++e) bias_v[e] = (row_ok && col0 + e < Sk) ? __ldg(bp + e) : 0.f;
          }
        } else {
          const uint16_t* bp = static_cast<const uint16_t*>(p.bias.ptr) + off;
          const bool fast = row_ok && col0 + kHB <= Sk && (reinterpret_cast<uintptr_t>(bp) & 15) == 0;
          if (fast) {
#pragma unroll
            for (int q8 = 0; q8 < kHB / 8; ++q8) {
              const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(bp) + q8);
              const uint32_t w[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                bias_v[q8 * 8 + 2 * e] = bias_to_float<kBf16>(static_cast<uint16_t>(w[e] & 0xffffu));
                bias_v[q8 * 8 + 2 * e + 1] = bias_to_float<kBf16>(static_cast<uint16_t>(w[e] >> 16));
              }
            }
          } else {
#pragma unroll
            for (int e = 0; e < kHB; ++e)
              bias_v[e] = (row_ok && col0 + e < Sk) ? bias_to_float<kBf16>(__ldg(bp + e)) : 0.f;
          }
        }
      }

      if (r == 0) FA2_STAMP(t, j, 0);
      mbar_wait(&s_full[t * 2 + buf], (j >> 1) & 1);
      tcgen05_fence_after();
      if (r == 0) FA2_STAMP(t, j, 1);
      float s[kHB];
      {
        uint32_t su[2][32];
        tmem_ld_32x32b_x32(tSb, su[0]);
        tmem_ld_32x32b_x32(tSb + 32, su[1]);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < kHB; ++e) s[e] = __uint_as_float(su[e / 32][e % 32]);
      }
      if constexpr (kBias) {
        // the bias joins the raw scores as bias / softmax_scale, so that the common factor c applies to both
        // (src/forward/compute_row_blocks.py:48-54, 68-69)
#pragma unroll
        for (int e = 0; e < kHB; ++e) s[e] = fmaf(bias_v[e], p.bias_mul, s[e]);
      }

      if (r == 0) FA2_STAMP(t, j, 2);
      const int col_last = col0 + kHB - 1;
      const bool need_mask = (col_last >= Sk) || (p.causal && col_last > tile_row0 + shift);
      if (need_mask) {
        const int limit = p.causal ? min(Sk - 1, row + shift) : (Sk - 1);
#pragma unroll
        for (int e = 0; e < kHB; ++e) s[e] = (col0 + e <= limit) ? s[e] : -INFINITY;
      }

      uint32_t pk[kHB / 2];
      float2 ls[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      // one pair of probabilities: p = 2^(s c - m), added to the row sum in fp32, rounded to 16 bits for P V
      auto exp_pair = [&](int e, const float2 neg_m2, float2 (&acc)[2], auto emulate) {
        const float2 x2 = ffma2(make_float2(s[e], s[e + 1]), c2, neg_m2);
        float2 p2;
        if (decltype(emulate)::value && ((e / 2) % 4) < L::kEmul) {
          p2 = ex2_emul2(x2);
        } else {
          p2.x = ex2_approx(x2.x);
          p2.y = ex2_approx(x2.y);
        }
        acc[(e / 2) % 2] = fadd2(acc[(e / 2) % 2], p2);
        pk[e / 2] = pack2<kBf16>(p2.x, p2.y);
      };

      // Fast path at head dim 64 (every half-block but the first and the masked ones): the exponentials do not wait for
      // the block maximum. The reference maximum only moves when a block exceeds it by 2^8 (lazy rescale), which after
      // the first blocks of a row is rare, so the probabilities are formed with the CURRENT reference while the maximum
      // is taken in the same instruction stream - the maxima fill the issue slots the MUFU leaves - and are kept unless
      // some row of the warp does need a new reference (then the block is redone below, from the untouched scores).
      // Same arithmetic either way. Measured (profiles/r02_experiments.txt): S = 16384, D = 64 non-causal 2.98-3.06 ->
      // 2.83-2.86 ms; at head dim 128, where the tensor pipe (1280 of the ~1600 cycles per half-block) and not the MUFU
      // is the tighter resource, 0.500 -> 0.520 ms, so it is off there.
      bool kept = false;
      if constexpr (!kBias && L::kSpeculate) {
        if (!need_mask && j > 0) {
          const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
          const float2 neg_m2 = make_float2(neg_m, neg_m);
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int e = 0; e < kHB; e += 8) {
#pragma unroll
            for (int q = 0; q < 4; ++q) mx[q] = fmax3(mx[q], s[e + q], s[e + q + 4]);
#pragma unroll
            for (int u = 0; u < 8; u += 2) exp_pair(e + u, neg_m2, ls, std::true_type{});
          }
          const float mb = fmax3(fmaxf(mx[0], mx[1]), mx[2], mx[3]) * c;
          kept = !__any_sync(0xffffffffu, mb > m_ref + kRescaleThreshold);
        }
      }
      if (!kept) {
        // block maximum: 3-input maxima (FMNMX3), four independent chains
        float mx[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) mx[e] = fmax3(s[e], s[e + 4], s[e + 8]);
#pragma unroll
        for (int e = 12; e < kHB - 4; e += 8) {
#pragma unroll
          for (int q = 0; q < 4; ++q) mx[q] = fmax3(mx[q], s[e + q], s[e + q + 4]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) mx[q] = fmaxf(mx[q], s[kHB - 4 + q]);
        const float mb = fmax3(fmaxf(mx[0], mx[1]), mx[2], mx[3]) * c;

        // lazy rescale: only move the reference max when the block max exceeds it by > 2^8
        const bool need = mb > m_ref + kRescaleThreshold;
        float alpha = 1.f;
        if (need) {
          alpha = ex2_approx(m_ref - mb);  // m_ref == -inf -> 0
          m_ref = mb;
        }
        if (j > 0 && __any_sync(0xffffffffu, need)) {
          // rare. O must be at rest: P V(j-1) - issued after this thread handed over P(j-1) - has to have retired
          mbar_wait(&pv_done[t], (j - 1) & 1);
          tcgen05_fence_after();
#pragma unroll 1
          for (int cidx = 0; cidx < D / 16; ++cidx) {
            uint32_t o[16];
            tmem_ld_32x32b_x16(tO + cidx * 16, o);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            tmem_st_32x32b_x16(tO + cidx * 16, o);
          }
          tmem_wait_st();
        }
        l *= alpha;

        if (r == 0) FA2_STAMP(t, j, 3);
        const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
        const float2 neg_m2 = make_float2(neg_m, neg_m);
        ls[0] = ls[1] = make_float2(0.f, 0.f);
        // masked half-blocks (and every half-block with a bias, which may hold -inf) keep all exponentials on the MUFU:
        // exact 0 for -inf scores
        if (need_mask || kBias) {
#pragma unroll
          for (int e = 0; e < kHB; e += 2) exp_pair(e, neg_m2, ls, std::false_type{});
        } else {
#pragma unroll
          for (int e = 0; e < kHB; e += 2) exp_pair(e, neg_m2, ls, std::true_type{});
        }
      }
      l += (ls[0].x + ls[1].x) + (ls[0].y + ls[1].y);
      if (r == 0) FA2_STAMP(t, j, 4);

      tmem_st_32x32b_x32(tSb, pk);
      tmem_wait_st();
      tcgen05_fence_before();
      mbar_arrive(&p_full[t * 2 + buf]);
      if (r == 0) FA2_STAMP(t, j, 5);
    }

    // ------------------------------ epilogue: O / l -> smem -> TMA store ---------------------------
    if (tile_row0 < Sq) {
      const float inv_l = (l > 0.f) ? 1.0f / l : 0.f;
      if (nhb > 0) {
        mbar_wait(&o_full[t], 0);
        tcgen05_fence_after();
      } else if (nh_max > 0) {
        // fully masked tile next to a live one: its Q tile is still being loaded under the shared q_full barrier; the
        // zero rows staged below must not be overwritten by that load
        mbar_wait(q_full, 0);
      }
      uint8_t* sO = sQ + t * L::kQTileBytes;  // the Q tile is dead once the last product of the tile retired
#pragma unroll
      for (int cidx = 0; cidx < D / 32; ++cidx) {
        uint32_t o[32];
        if (nhb > 0) {
          tmem_ld_32x32b_x32(tO + cidx * 32, o);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) o[e] = 0u;
        }
        const int col = cidx * 32;  // first head-dim column of this chunk
        uint8_t* sub = sO + (col / 64) * L::kQSubBytes;
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
      named_bar_sync(1 + t, kBM);
      if (r == 0) {
        if (p.seqlens != nullptr) tmap_acquire(mo);
        // sub-tiles that lie entirely beyond the tensors' head dim hold zeros only: nothing to store
        for (int j = 0; j < kSub; ++j)
          if (j * 64 < p.d_cols) tma_store_4d(mo, sO + j * L::kQSubBytes, j * 64, tile_row0, head, bc);
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
  if (warp == L::kMmaWarp) {
    tcgen05_fence_after();
    tmem_dealloc<512>(0u);
  }
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
template <int D, bool kBf16, bool kBias>
static cudaError_t launch_fwd2_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                                 const CUtensorMap& to, const FwdParams& p, cudaStream_t stream) {
  using L = Fwd2Cfg<D>;
  if (cudaError_t e = ensure_dynamic_smem<fa2_fwd2_kernel<D, kBf16, kBias>>(L::kDynBytes); e != cudaSuccess) return e;
  FwdParams pp = p;
  pp.n_pairs = (p.Sq + L::kTiles * kBM - 1) / (L::kTiles * kBM);
  // a head's CTAs stream its K and V (GQA: shared by the group, counted per query head all the same - an upper bound)
  pp.tail_log2 = tail_log2_for(tail_heads_for(static_cast<size_t>(p.Sk) * D * 2 * 2, p.causal), p.Hq);
  dim3 grid(pp.n_pairs, p.Hq, p.B);
  return launch_pdl(fa2_fwd2_kernel<D, kBf16, kBias>, grid, dim3(L::kThreads), L::kDynBytes, stream, tq, tk, tv, to, pp);
}

// plain forward (no bias, no dropout) for head-dim instantiations 64 / 128 / 256; with a bias only at 256
cudaError_t launch_fwd2(int D, bool bf16, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                        const CUtensorMap& to, const FwdParams& p, cudaStream_t stream) {
  const bool bias = p.bias.ptr != nullptr;
  if (p.dropout_p > 0.f) return cudaErrorInvalidValue;
  if (D == 256) {
    if (bias) return bf16 ? launch_fwd2_t<256, true, true>(tq, tk, tv, to, p, stream)
                          : launch_fwd2_t<256, false, true>(tq, tk, tv, to, p, stream);
    return bf16 ? launch_fwd2_t<256, true, false>(tq, tk, tv, to, p, stream)
                : launch_fwd2_t<256, false, false>(tq, tk, tv, to, p, stream);
  }
  if (bias) return cudaErrorInvalidValue;
  if (D == 128) return bf16 ? launch_fwd2_t<128, true, false>(tq, tk, tv, to, p, stream)
                            : launch_fwd2_t<128, false, false>(tq, tk, tv, to, p, stream);
  if (D == 64) return bf16 ? launch_fwd2_t<64, true, false>(tq, tk, tv, to, p, stream)
                           : launch_fwd2_t<64, false, false>(tq, tk, tv, to, p, stream);
  return cudaErrorInvalidValue;
}

}  // namespace fa2
