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

namespace {

constexpr int kBlk = 128;
#ifndef FA2_DKDV_WGS
#define FA2_DKDV_WGS 4
#endif
constexpr int kCmpWgs = FA2_DKDV_WGS;     // compute warpgroups (2 or 4)
constexpr int kCmpCols = kBlk / kCmpWgs;  // query columns per compute thread
constexpr int kCmpThreads = kCmpWgs * 128;
constexpr int kThreads = kCmpThreads + 128;
constexpr int kMmaWarp = kCmpWgs * 4, kTmaWarp = kMmaWarp + 1;
// register pool: 640 threads x 96 (launch bound) = 61440 = 4 x 128 x 104 + 128 x 64
//                384 threads x 168               = 64512 = 2 x 128 x 208 + 128 x 88
constexpr int kRegsCmp = kCmpWgs == 4 ? 104 : 208;
constexpr int kRegsIssue = kCmpWgs == 4 ? 64 : 88;
constexpr int kQStages = 3;   // Q (+LSE) ring: Q(i) lives from S^T(i) to dK(i), and Q(i+1) is wanted right after dV(i)
constexpr int kQStagesDbl = 4;  // head dim 64, double-buffered S^T: Q(i+2) is wanted right after dV(i)
constexpr int kDoStages = 2;  // dO (+delta) ring
// the bias variant at head dim 128 gives its third Q stage to the bias tile
__host__ __device__ constexpr int q_stages(int D, int extra) {
  return D == 64 ? kQStagesDbl : (extra == 1 && D == 128) ? 2 : kQStages;
}
// How many of every 4 pairs of exponentials are evaluated on the FMA pipes (ex2_emul2_neg) instead of the MUFU. Measured
// (scripts/build_variants.sh rebuilds with another value for A/B runs): head dim 128, S = 4096 causal, sustained: 0/4
// 1.096 ms, 1/4 1.063, 2/4 1.02-1.03, 3/4 1.08, 4/4 1.14; head dim 64 with the double-buffered S^T, S = 16384: 0/4
// 3.90-3.96 ms, 1/4 4.02, 2/4 4.10-4.20, 3/4 4.28-4.42 - there the issue slots, not the MUFU, are what the compute warps
// run out of, and an emulated pair costs ~12 instructions against 2.
#ifndef FA2_DKDV_EMUL
#define FA2_DKDV_EMUL 2
#endif
#ifndef FA2_DKDV_EMUL_D64
#define FA2_DKDV_EMUL_D64 0
#endif
template <int D>
__host__ __device__ constexpr bool emulate_pair(int q) {
  constexpr int n = D == 64 ? FA2_DKDV_EMUL_D64 : FA2_DKDV_EMUL;
  return n == 4 ? true : n == 3 ? (q % 4 != 0) : n == 2 ? (q % 2 == 1) : n == 1 ? (q % 4 == 3) : false;
}

// kQS: stages of the Q ring; kBiasTile: room for one [128 queries x 128 keys] 16-bit tile of the attention bias (the
// bias variant; at head dim 128 it takes the place of the third Q stage)
template <int D, int kQS = kQStages, bool kBiasTile = false>
struct DkdvSmem {
  static constexpr int kTileBytes = kBlk * D * 2;  // K, V, Q, dO tiles
  static constexpr int kBiasBytes = kBlk * kBlk * 2;
  static constexpr int kKOff = 0;
  static constexpr int kVOff = kKOff + kTileBytes;
  static constexpr int kQOff = kVOff + kTileBytes;
  static constexpr int kBiasOff = kQOff + kQS * kTileBytes;  // multiples of 16 KB: 1024-byte aligned for the 128-byte swizzle
  static constexpr int kDOOff = kBiasOff + (kBiasTile ? kBiasBytes : 0);
  // The LSE row of query block i rides with Q(i) (stage released by dK(i), i.e. after dS(i)). Its delta row rides with
  // dO(i), but is read by dS(i) - possibly after dV(i), which releases the dO stage, has retired: the slot has its own
  // "empty" barrier on which every compute thread arrives once it has its delta values. (Without it the load of block
  // i + 2 could overwrite the row under a late reader: run-to-run differences at S = 16384.)
  static constexpr int kLseOff = kDOOff + kDoStages * kTileBytes;   // kQS x 512 B
  static constexpr int kDeltaOff = kLseOff + kQS * kBlk * 4;        // kDoStages x 512 B
  static constexpr int kBarOff = kDeltaOff + kDoStages * kBlk * 4;
  static constexpr int kNumBars = 24;
  static constexpr int kTmemSlotOff = kBarOff + kNumBars * 8;
  static constexpr int kBytes = kTmemSlotOff + 16;
  static_assert(kBytes <= 232448, "dK/dV kernel smem budget exceeded");
};

// mbarrier slots (kQS = stages of the Q ring)
template <int kQS>
struct DkdvBars {
  static constexpr int kKvFull = 0;
  static constexpr int kQFull = 1;                // [kQS]
  static constexpr int kQEmpty = kQFull + kQS;    // [kQS]
  static constexpr int kDoFull = kQEmpty + kQS;   // [2]
  static constexpr int kDoEmpty = kDoFull + 2;    // [2]
  static constexpr int kSFull = kDoEmpty + 2;     // [2] S^T(i) is in tensor memory (buffer i & 1 when double-buffered, else slot 0)
  static constexpr int kPFull = kSFull + 2;       // [2] P^T(i) is in tensor memory (over S^T)
  static constexpr int kDpFull = kPFull + 2;
  static constexpr int kDsFull = kDpFull + 1;     // dS^T(i) is in tensor memory (over dP^T)
  static constexpr int kDkvFull = kDsFull + 1;
  static constexpr int kDeltaEmpty = kDkvFull + 1;  // [2] every compute thread has read the delta row in this slot
  static constexpr int kBiasFull = kDeltaEmpty + 2;   // bias variant: the bias tile of this iteration has landed
  static constexpr int kBiasEmpty = kBiasFull + 1;    // ... and every compute thread has read its values
  static constexpr int kCount = kBiasEmpty + 1;
  static_assert(kCount <= 24, "more mbarriers than DkdvSmem reserves");
};

// streaming 16-byte store (written once, read once by the dQ kernel: evict first)
__device__ __forceinline__ void st_global_cs_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

}  // namespace

// kExtra: 0 = the plain kernel; 1 = adds the attention bias to the recomputed scores (extras.cuh); 2 = variable-length
// batches (per-sequence lengths and tensor maps, varlen_sm100.cu). Separate instantiations: the compute warps of the plain
// kernel have no register to spare (measured: the runtime-selected form spilled and cost 12 %).
template <int D, bool kBf16, bool kSpill, int kExtra>
__global__ void __launch_bounds__(kThreads, 1)
fa2_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                    const __grid_constant__ CUtensorMap tmap_dk, const __grid_constant__ CUtensorMap tmap_dv,
                    const __grid_constant__ CUtensorMap tmap_bias, const DkdvParams p) {
  constexpr bool kDbl = D == 64;  // double-buffered S^T (see the top of the file)
  constexpr int kQS = q_stages(D, kExtra);
  using L = DkdvSmem<D, kQS, kExtra == 1>;
  using Bar = DkdvBars<kQS>;
  constexpr int kKvFull = Bar::kKvFull, kQFull = Bar::kQFull, kQEmpty = Bar::kQEmpty, kDoFull = Bar::kDoFull,
                kDoEmpty = Bar::kDoEmpty, kSFull = Bar::kSFull, kPFull = Bar::kPFull, kDpFull = Bar::kDpFull,
                kDsFull = Bar::kDsFull, kDkvFull = Bar::kDkvFull, kDeltaEmpty = Bar::kDeltaEmpty;
  [[maybe_unused]] constexpr int kBiasFull = Bar::kBiasFull, kBiasEmpty = Bar::kBiasEmpty;
  constexpr int kSub = D / 64;
  constexpr int kSubBytes = kBlk * 128;  // [128 rows x 64 x 16-bit]
  constexpr uint32_t kTmemS = 0, kTmemDP = kDbl ? 256 : 128, kTmemDV = kDbl ? 384 : 256, kTmemDK = kDbl ? 448 : 384;
  constexpr int kAhead = kDbl ? 2 : 1;  // S^T(i + kAhead) is issued right after dV(i)
  // S^T buffer of query block i, the barrier slot that goes with it and the parity of its i-th use
  auto sbuf = [](int i) { return kDbl ? (i & 1) : 0; };
  auto spar = [](int i) { return static_cast<uint32_t>(kDbl ? (i >> 1) & 1 : i & 1); };
  constexpr int kPairs = kCmpCols / 2;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem + L::kKOff;
  uint8_t* sV = smem + L::kVOff;
  uint8_t* sQ = smem + L::kQOff;
  uint8_t* sDO = smem + L::kDOOff;
  [[maybe_unused]] uint8_t* sBias = smem + L::kBiasOff;
  float* sLse = reinterpret_cast<float*>(smem + L::kLseOff);
  float* sDelta = reinterpret_cast<float*>(smem + L::kDeltaOff);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlotOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // Causal: low key blocks are seen by the most query blocks - longest first. (NOT the block-major tail of the other
  // kernels, fa2_kernels.cuh: the remapped coordinates cost this kernel's compute warps registers they do not have -
  // 4 -> 48 bytes of spills and +16 % cycles, against the 1.8 % the better tail is worth here.)
  const int nblk = static_cast<int>(blockIdx.x);
  constexpr bool kBias = kExtra == 1, kVarlen = kExtra == 2, kSplit = kExtra == 3;
  const int head_kv = blockIdx.y;  // kSplit: the (kv head, part) index; the query heads below follow from it unchanged
  const int head_kv_src = kSplit ? head_kv / p.kv_split : head_kv;  // the kv head whose K / V rows are loaded
  const int batch = blockIdx.z;
  const int kv0 = nblk * kBlk;
  // variable-length mode: this batch element's own length and tensor maps (bounded by it; batch coordinate 0)
  const int len = kVarlen ? min(max(p.seqlens[batch], 0), p.Sq) : 0;  // clamped like the tensor maps
  const int Sq = kVarlen ? len : p.Sq, Sk = kVarlen ? len : p.Sk, bc = kVarlen ? 0 : batch;
  const CUtensorMap* vm = kVarlen ? p.vmaps + static_cast<size_t>(batch) * kVarlenMapsPerBatch : nullptr;
  const CUtensorMap* mq = kVarlen ? vm : &tmap_q;
  const CUtensorMap* mk = kVarlen ? vm + 1 : &tmap_k;
  const CUtensorMap* mv = kVarlen ? vm + 2 : &tmap_v;
  const CUtensorMap* mdo = kVarlen ? vm + 3 : &tmap_do;
  const CUtensorMap* mdk = kVarlen ? vm + 4 : &tmap_dk;
  const CUtensorMap* mdv = kVarlen ? vm + 5 : &tmap_dv;
  if (kVarlen && kv0 >= Sk) return;  // key block in the padding: its dK / dV rows were zero-filled by the caller
  const int n_q_blocks = (Sq + kBlk - 1) / kBlk;
  const int group = kSplit ? p.Hq / (p.Hkv * p.kv_split) : p.Hq / p.Hkv;  // query heads this CTA loops over
  const int shift = Sk - Sq;

  // first query block that sees this key block (bottom-right aligned causal mask)
  int m_start = 0;
  if (p.causal) m_start = max(0, (kv0 - shift) / kBlk);
  const int m_count = max(0, n_q_blocks - m_start);
  const int T = m_count * group;  // iterations: (head in group) x (query block)

  if ((smem_u32(smem) & 1023u) != 0) __trap();

  if (warp == kTmaWarp && lane == 0) {
    if constexpr (kVarlen) {  // maps published by fa2_varlen_maps_kernel
      tmap_acquire(mq);
      tmap_acquire(mk);
      tmap_acquire(mv);
      tmap_acquire(mdo);
    }
    tma_prefetch_desc(mq);
    tma_prefetch_desc(mk);
    tma_prefetch_desc(mv);
    tma_prefetch_desc(mdo);
    if constexpr (!kVarlen) {
      tma_prefetch_desc(mdk);
      tma_prefetch_desc(mdv);
    }
  }
  if (warp == kMmaWarp) {
    if (lane == 0) {
      mbar_init(&bars[kKvFull], 1);
      for (int s = 0; s < kQS; ++s) {
        mbar_init(&bars[kQFull + s], 1);
        mbar_init(&bars[kQEmpty + s], 1);
      }
      for (int s = 0; s < kDoStages; ++s) {
        mbar_init(&bars[kDoFull + s], 1);
        mbar_init(&bars[kDoEmpty + s], 1);
        mbar_init(&bars[kDeltaEmpty + s], kCmpThreads);
      }
      mbar_init(&bars[kBiasFull], 1);
      mbar_init(&bars[kBiasEmpty], kCmpThreads);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&bars[kSFull + b], 1);
        mbar_init(&bars[kPFull + b], kCmpThreads);
      }
      mbar_init(&bars[kDpFull], 1);
      mbar_init(&bars[kDsFull], kCmpThreads);
      mbar_init(&bars[kDkvFull], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  // whole-TMEM allocation => base 0; the literal keeps tcgen05 addresses warp-uniform for ptxas (see fwd_sm100.cu)
  if (*tmem_slot != 0u) __trap();
  // PDL: the prologue above overlapped the tail of the previous kernel of the stream (the delta kernel); nothing below may
  // run before it has completed
  pdl_wait();
  if (threadIdx.x == 0) {
    pdl_launch_dependents();
    // the dQ kernel that follows on the stream draws its work items from this counter
    if (p.zero_word != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) *p.zero_word = 0;
  }

  if (warp >= kMmaWarp) {
    reg_dealloc<kRegsIssue>();
    if (warp == kTmaWarp) {
      // ========================================= TMA producer ========================================
      if (T > 0 && elect_one()) {
        mbar_arrive_expect_tx(&bars[kKvFull], 2 * L::kTileBytes);
        for (int j = 0; j < kSub; ++j) {
          tma_load_4d(sK + j * kSubBytes, mk, &bars[kKvFull], j * 64, kv0, head_kv_src, bc);
          tma_load_4d(sV + j * kSubBytes, mv, &bars[kKvFull], j * 64, kv0, head_kv_src, bc);
        }
        // Q(i) + its LSE row: released by dK(i - kQS)
        auto load_q = [&](int i) {
          const int qs = i % kQS;
          const uint32_t qph = (i / kQS) & 1;
          const int head = head_kv * group + i / m_count;
          const int m = m_start + i % m_count;
          const int64_t row_off = (static_cast<int64_t>(batch) * p.Hq + head) * p.lse_stride + m * kBlk;
          mbar_wait(&bars[kQEmpty + qs], qph ^ 1);
          mbar_arrive_expect_tx(&bars[kQFull + qs], L::kTileBytes + kBlk * 4);
          for (int j = 0; j < kSub; ++j)
            tma_load_4d(sQ + qs * L::kTileBytes + j * kSubBytes, mq, &bars[kQFull + qs], j * 64, m * kBlk, head, bc);
          bulk_load_1d(sLse + qs * kBlk, p.lse + row_off, kBlk * 4, &bars[kQFull + qs]);
        };
        // the Q ring runs kAhead - 1 blocks ahead of the dO ring (double-buffered S^T: Q(i + 2) is wanted right after dV(i))
        for (int i = 0; i < kAhead - 1 && i < T; ++i) load_q(i);
        for (int it = 0; it < T; ++it) {
          const int ds = it % kDoStages;
          const uint32_t dph = (it / kDoStages) & 1;
          const int head = head_kv * group + it / m_count;
          const int m = m_start + it % m_count;
          const int64_t row_off = (static_cast<int64_t>(batch) * p.Hq + head) * p.lse_stride + m * kBlk;
          if constexpr (kBias) {
            if (p.bias_tma) {  // bias tile (it): two [128 queries x 64 keys] sub-tiles; released by the readers of tile it - 1
              mbar_wait(&bars[kBiasEmpty], (it & 1) ^ 1);
              mbar_arrive_expect_tx(&bars[kBiasFull], L::kBiasBytes);
              for (int j = 0; j < 2; ++j)
                tma_load_4d(sBias + j * (kBlk * 128), &tmap_bias, &bars[kBiasFull], kv0 + j * 64, m * kBlk,
                            p.bias_head_mul * head, p.bias_batch_mul * batch);
            }
          }
          if (it + kAhead - 1 < T) load_q(it + kAhead - 1);
          // dO(it) + its delta row: released by dV(it - 2) and by the readers of delta(it - 2)
          mbar_wait(&bars[kDoEmpty + ds], dph ^ 1);
          mbar_wait(&bars[kDeltaEmpty + ds], dph ^ 1);
          mbar_arrive_expect_tx(&bars[kDoFull + ds], L::kTileBytes + kBlk * 4);
          for (int j = 0; j < kSub; ++j)
            tma_load_4d(sDO + ds * L::kTileBytes + j * kSubBytes, mdo, &bars[kDoFull + ds], j * 64, m * kBlk, head, bc);
          bulk_load_1d(sDelta + ds * kBlk, p.delta + row_off, kBlk * 4, &bars[kDoFull + ds]);
        }
      }
    } else if (warp == kMmaWarp) {
      // ========================================== MMA issuer =========================================
      if (T > 0 && elect_one()) {
        constexpr uint32_t idesc_kk = make_idesc_f16(kBlk, kBlk, kBf16, false, false);  // S^T, dP^T
        const uint32_t idesc_tn = make_idesc_f16(kBlk, p.d_cols, kBf16, false, true);  // dV, dK: N = the tensors' head dim (x16)
        const int d_steps = p.d_steps;
        // K-major views (rows of the tile are the M/N index, the reduction runs along the 128-byte rows)
        const SmemDesc k_kmaj = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
        const SmemDesc v_kmaj = make_smem_desc_sw128(smem_u32(sV), 16, 1024);
        const SmemDesc q_kmaj = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
        const SmemDesc do_kmaj = make_smem_desc_sw128(smem_u32(sDO), 16, 1024);
        // MN-major views (rows of the tile are the reduction index)
        const SmemDesc q_mn = make_smem_desc_sw128(smem_u32(sQ), kSubBytes, 1024);
        const SmemDesc do_mn = make_smem_desc_sw128(smem_u32(sDO), kSubBytes, 1024);

        // [128 x D] x [128 x D]^T, both K-major, reduction over D
        auto issue_kmajor = [&](uint32_t d_tmem, SmemDesc a, SmemDesc b) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const uint32_t off = (k / 4) * kSubBytes + (k % 4) * 32;
            if (k < d_steps) umma_ss(d_tmem, advance(a, off), advance(b, off), idesc_kk, k > 0 ? 1u : 0u);
          }
        };
        // [128 kv x 128 q] (P^T or dS^T, 16-bit, in tensor memory) x [128 q x D] (dO or Q, MN-major), reduction over q.
        // Compute warpgroup g packed its kCmpCols query columns (k-steps g * kCmpCols / 16 ...) into the first
        // kCmpCols / 2 columns of its slice.
        auto issue_acc = [&](uint32_t d_tmem, uint32_t a_tmem, SmemDesc b_mn, bool accumulate) {
#pragma unroll
          for (int k = 0; k < kBlk / 16; ++k) {
            constexpr int kStepsPerWg = kCmpCols / 16;
            const uint32_t a_col = (k / kStepsPerWg) * kCmpCols + (k % kStepsPerWg) * 8;
            umma_ts(d_tmem, a_tmem + a_col, advance(b_mn, k * 2048), idesc_tn, (accumulate || k > 0) ? 1u : 0u);
          }
        };

        mbar_wait(&bars[kKvFull], 0);
        mbar_wait(&bars[kQFull + 0], 0);
        tcgen05_fence_after();
        issue_kmajor(kTmemS, k_kmaj, q_kmaj);  // S^T(0) = K Q(0)^T
        umma_commit(&bars[kSFull]);
        if (kDbl && T > 1) {  // S^T(1) into the second buffer
          mbar_wait(&bars[kQFull + 1], 0);
          tcgen05_fence_after();
          issue_kmajor(kTmemS + 128, k_kmaj, advance(q_kmaj, L::kTileBytes));
          umma_commit(&bars[kSFull + 1]);
        }
        mbar_wait(&bars[kDoFull + 0], 0);
        tcgen05_fence_after();
        issue_kmajor(kTmemDP, v_kmaj, do_kmaj);  // dP^T(0) = V dO(0)^T
        umma_commit(&bars[kDpFull]);

        for (int it = 0; it < T; ++it) {
          const uint32_t par = it & 1;
          const int qs = it % kQS, qsn = (it + kAhead) % kQS;
          const int ds = it % kDoStages, dsn = (it + 1) % kDoStages;
          const uint32_t qphn = ((it + kAhead) / kQS) & 1, dphn = ((it + 1) / kDoStages) & 1;
          const uint32_t tmem_s = kTmemS + sbuf(it) * 128;
          // dV += P^T(it) dO(it)
          mbar_wait(&bars[kPFull + sbuf(it)], spar(it));
          tcgen05_fence_after();
          FA2_STAMP(2, it, 0);
          issue_acc(kTmemDV, tmem_s, advance(do_mn, ds * L::kTileBytes), it > 0);
          umma_commit(&bars[kDoEmpty + ds]);  // dO(it) has fed dP(it) and dV(it)
          FA2_STAMP(2, it, 1);
          // S^T(it + kAhead) = K Q(it + kAhead)^T into the buffer P^T(it) was just read from (the tensor pipe executes in
          // issue order)
          if (it + kAhead < T) {
            mbar_wait(&bars[kQFull + qsn], qphn);
            tcgen05_fence_after();
            issue_kmajor(tmem_s, k_kmaj, advance(q_kmaj, qsn * L::kTileBytes));
            umma_commit(&bars[kSFull + sbuf(it)]);
          }
          FA2_STAMP(2, it, 2);
          // dK += dS^T(it) Q(it)
          mbar_wait(&bars[kDsFull], par);
          tcgen05_fence_after();
          FA2_STAMP(2, it, 3);
          issue_acc(kTmemDK, kTmemDP, advance(q_mn, qs * L::kTileBytes), it > 0);
          umma_commit(&bars[kQEmpty + qs]);  // Q(it) has fed S(it) and dK(it)
          FA2_STAMP(2, it, 4);
          // dP^T(it+1) = V dO(it+1)^T (overwrites dS^T(it))
          if (it + 1 < T) {
            mbar_wait(&bars[kDoFull + dsn], dphn);
            tcgen05_fence_after();
            issue_kmajor(kTmemDP, v_kmaj, advance(do_kmaj, dsn * L::kTileBytes));
            umma_commit(&bars[kDpFull]);
          }
          FA2_STAMP(2, it, 5);
        }
        umma_commit(&bars[kDkvFull]);
      }
    }
  } else {
    // ====================================== compute warpgroups =====================================
    reg_alloc<kRegsCmp>();
    const int wg = warp >> 2;
    const int r = threadIdx.x & 127;  // key row inside the block == TMEM lane
    const int kv = kv0 + r;
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS0 = lane_sel + kTmemS + wg * kCmpCols;  // P^T (16-bit) goes over the head of the same columns
    const uint32_t tDP = lane_sel + kTmemDP + wg * kCmpCols;  // ... and dS^T over the head of these
    const float2 neg_c2 = make_float2(-p.scale_log2, -p.scale_log2);
    // spill: this thread's 16-byte pieces of dS^T(m, n) go to [piece = 8 queries][key row][16 B] inside the 32 KB tile,
    // so that one warp store covers 512 contiguous bytes. (A dedicated spill warpgroup reading dS^T back from tensor
    // memory was slower: 1.09 ms against 1.02 - four warps do not keep the 32 B/clk/SM store path busy. Replacing the
    // closed-form tile address below - a division and a 64-bit product per thread and iteration, ~15 % of the kernel's
    // instructions - by incremental pointers was slower as well: 1.08 ms against 1.04, the stores then pile up
    // earlier. With the chip at its power cap and the store path ~45 % busy, fewer instructions are not faster.)
    uint8_t* spill_row = nullptr;
    long long prefix = 0;
    // The store path takes ~32 B/clk/SM: the 32 KB of a block pair occupy it for >= 1000 cycles, and a warp that issues
    // its four stores back to back sits in them for most of that. So every thread writes half of its pieces right
    // after handing dS^T over and carries the other half to the same point after P^T of the next block - each time
    // just before a barrier it would have to wait at anyway.
    constexpr int kPieces = kCmpCols / 8;
    uint32_t carry[kPieces / 2 * 4];
    uint8_t* carry_dst = nullptr;
    if (kSpill) spill_row = static_cast<uint8_t*>(p.ds_spill) + (static_cast<size_t>(wg) * (kCmpCols / 8) * kBlk + r) * 16;

    for (int it = 0; it < T; ++it) {
      const uint32_t par = it & 1;
      const int qs = it % kQS, ds = it % kDoStages;
      const uint32_t qph = (it / kQS) & 1, dph = (it / kDoStages) & 1;
      const int j_in_head = it % m_count;
      const int m = m_start + j_in_head;
      const int q0 = m * kBlk + wg * kCmpCols;  // first query column of this thread's slice
      const float4* lse_s = reinterpret_cast<const float4*>(sLse + qs * kBlk + wg * kCmpCols);
      const float4* delta_s = reinterpret_cast<const float4*>(sDelta + ds * kBlk + wg * kCmpCols);
      const bool need_mask = (kv0 + kBlk > Sk) || (m * kBlk + kBlk > Sq) || (p.causal && (kv0 + kBlk - 1 > m * kBlk + shift));
      // kBias: this thread's 32 bias values bias[q0 + j][kv] (one coalesced 64-byte warp load per query row), fetched
      // before the waits below so that their latency hides behind the S^T product; scaled to the log2 domain.
      // 16-bit values travel as packed pairs, fp32 ones unpacked.
      uint32_t braw[kCmpCols];
      if constexpr (kBias) {
        if (!p.bias_tma) {
        const int64_t off0 = batch * p.bias.sb + (head_kv * group + it / m_count) * p.bias.sh +
                             static_cast<int64_t>(q0) * p.bias.sm + kv;
        const bool col_ok = kv < Sk;
        if (p.bias.is_fp32) {
          const float* bp = static_cast<const float*>(p.bias.ptr) + off0;
#pragma unroll
          for (int j = 0; j < kCmpCols; ++j)
            braw[j] = (col_ok && q0 + j < Sq) ? __float_as_uint(__ldg(bp + j * p.bias.sm)) : 0u;
        } else {
          const uint16_t* bp = static_cast<const uint16_t*>(p.bias.ptr) + off0;
#pragma unroll
          for (int j = 0; j < kCmpCols; j += 2) {
            const uint32_t lo = (col_ok && q0 + j < Sq) ? __ldg(bp + j * p.bias.sm) : 0u;
            const uint32_t hi = (col_ok && q0 + j + 1 < Sq) ? __ldg(bp + (j + 1) * p.bias.sm) : 0u;
            braw[j / 2] = lo | (hi << 16);
          }
        }
        }
      }

      // ---- P^T = exp2(S^T * c - LSE) = exp2(-(LSE - S^T * c)) ----
      if (r == 0 && (wg == 0 || wg == kCmpWgs - 1)) FA2_STAMP(wg ? 1 : 0, it, 0);
      const uint32_t tS = tS0 + sbuf(it) * 128;
      mbar_wait(&bars[kQFull + qs], qph);  // the LSE row travels with Q
      mbar_wait(&bars[kSFull + sbuf(it)], spar(it));
      tcgen05_fence_after();
      if (r == 0 && (wg == 0 || wg == kCmpWgs - 1)) FA2_STAMP(wg ? 1 : 0, it, 1);
      uint32_t su[kCmpCols];
#pragma unroll
      for (int h = 0; h < kCmpCols / 32; ++h) tmem_ld_32x32b_x32(tS + h * 32, *reinterpret_cast<uint32_t(*)[32]>(&su[h * 32]));
      tmem_wait_ld();
      if constexpr (kBias) {
        if (p.bias_tma) {
          // this thread's column (its key) of the staged bias tile, 32 query rows: 128-byte swizzled [query][64 keys]
          // sub-tiles; the 32 lanes of a warp read 64 contiguous (permuted) bytes of one row - conflict free
          mbar_wait(&bars[kBiasFull], it & 1);
          const uint8_t* sub = sBias + (r >> 6) * (kBlk * 128) + (r & 7) * 2;
          const uint32_t chunk = (r & 63) >> 3;
#pragma unroll
          for (int j = 0; j < kCmpCols; j += 2) {
            const uint32_t q_lo = wg * kCmpCols + j, q_hi = q_lo + 1;
            const uint32_t lo = *reinterpret_cast<const uint16_t*>(sub + sw128_offset(q_lo, chunk));
            const uint32_t hi = *reinterpret_cast<const uint16_t*>(sub + sw128_offset(q_hi, chunk));
            braw[j / 2] = lo | (hi << 16);
          }
          // TMA overwrites the tile once every reader has arrived, so the loads above must have RETURNED before this
          // thread arrives (mbarrier.arrive does not wait for them; the forward kernel's bias buffer showed the race,
          // fwd_sm100.cu): the arrive's address depends on the loaded registers through a term that is always 0
          uint32_t dep = 0;
#pragma unroll
          for (int j = 0; j < kCmpCols / 2; ++j) dep ^= braw[j];
          dep &= static_cast<uint32_t>(p.bias_head_mul) >> 1;  // bias_head_mul is 0 or 1
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[kBiasEmpty]) + dep) : "memory");
        }
      }
      float2 pf[kPairs];  // probabilities, fp32, pairs of neighbouring query columns
#pragma unroll
      for (int j4 = 0; j4 < kCmpCols / 4; ++j4) {
        const float4 l4 = lse_s[j4];
        const int j = j4 * 4;
        const float2 t0 = ffma2(make_float2(__uint_as_float(su[j]), __uint_as_float(su[j + 1])), neg_c2, make_float2(l4.x, l4.y));
        const float2 t1 = ffma2(make_float2(__uint_as_float(su[j + 2]), __uint_as_float(su[j + 3])), neg_c2, make_float2(l4.z, l4.w));
        // a fully masked query row carries LSE = -inf: exp2(+inf) = inf here, replaced by 0 in the masked path below
        // (such rows only occur in blocks that need masking)
        if constexpr (kBias) {
          // P = exp2(S c + bias log2(e) - LSE) (src/backward/compute_dkdv.py:60-69). All on the MUFU: a bias of -inf
          // (used as a mask) must give exactly 0, and a row hidden entirely that way carries LSE = -inf
          float tt[4] = {t0.x, t0.y, t1.x, t1.y};
          const float ll[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float b;
            if (p.bias.is_fp32) b = __uint_as_float(braw[j + e]);
            else b = bias_to_float<kBf16>(static_cast<uint16_t>((braw[(j + e) / 2] >> (16 * (e % 2))) & 0xffffu));
            tt[e] = fmaf(b, -1.4426950408889634f, tt[e]);
            tt[e] = (ll[e] == -INFINITY) ? INFINITY : tt[e];
          }
          pf[j4 * 2] = make_float2(ex2_approx(-tt[0]), ex2_approx(-tt[1]));
          pf[j4 * 2 + 1] = make_float2(ex2_approx(-tt[2]), ex2_approx(-tt[3]));
        } else {
        if (emulate_pair<D>(j4 * 2)) pf[j4 * 2] = ex2_emul2_neg(t0);
        else pf[j4 * 2] = make_float2(ex2_approx(-t0.x), ex2_approx(-t0.y));
        if (emulate_pair<D>(j4 * 2 + 1)) pf[j4 * 2 + 1] = ex2_emul2_neg(t1);
        else pf[j4 * 2 + 1] = make_float2(ex2_approx(-t1.x), ex2_approx(-t1.y));
        }
      }
      if (need_mask) {
#pragma unroll
        for (int j = 0; j < kPairs; ++j) {
          const int qi = q0 + 2 * j;
          const bool row_ok = kv < Sk;
          const bool ok0 = row_ok && (qi < Sq) && (!p.causal || kv <= qi + shift);
          const bool ok1 = row_ok && (qi + 1 < Sq) && (!p.causal || kv <= qi + 1 + shift);
          pf[j].x = ok0 ? pf[j].x : 0.f;
          pf[j].y = ok1 ? pf[j].y : 0.f;
        }
      }
      {
        uint32_t pk[kPairs];
#pragma unroll
        for (int j = 0; j < kPairs; ++j) pk[j] = pack2<kBf16>(pf[j].x, pf[j].y);
        if constexpr (kPairs == 16) tmem_st_32x32b_x16(tS, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
        else tmem_st_32x32b_x32(tS, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
        tmem_wait_st();
      }
      tcgen05_fence_before();
      mbar_arrive(&bars[kPFull + sbuf(it)]);
      if (r == 0 && (wg == 0 || wg == kCmpWgs - 1)) FA2_STAMP(wg ? 1 : 0, it, 2);
      if (kSpill && it > 0) {
#pragma unroll
        for (int c = 0; c < kPieces / 2; ++c)
          st_global_cs_v4(carry_dst + (kPieces / 2 + c) * (kBlk * 16), carry[c * 4], carry[c * 4 + 1], carry[c * 4 + 2], carry[c * 4 + 3]);
      }

      // ---- dS^T = P^T o (dP^T - delta) ----
      mbar_wait(&bars[kDoFull + ds], dph);  // the delta row travels with dO
      mbar_wait(&bars[kDpFull], par);
      tcgen05_fence_after();
      if (r == 0 && (wg == 0 || wg == kCmpWgs - 1)) FA2_STAMP(wg ? 1 : 0, it, 3);
      uint32_t dpu[kCmpCols];
#pragma unroll
      for (int h = 0; h < kCmpCols / 32; ++h) tmem_ld_32x32b_x32(tDP + h * 32, *reinterpret_cast<uint32_t(*)[32]>(&dpu[h * 32]));
      tmem_wait_ld();
      uint32_t w[kPairs];
#pragma unroll
      for (int j4 = 0; j4 < kCmpCols / 4; ++j4) {
        const float4 d4 = delta_s[j4];
        const int j = j4 * 4;
        const float2 x0 = fsub2(make_float2(__uint_as_float(dpu[j]), __uint_as_float(dpu[j + 1])), make_float2(d4.x, d4.y));
        const float2 x1 = fsub2(make_float2(__uint_as_float(dpu[j + 2]), __uint_as_float(dpu[j + 3])), make_float2(d4.z, d4.w));
        const float2 s0 = fmul2(pf[j4 * 2], x0);
        const float2 s1 = fmul2(pf[j4 * 2 + 1], x1);
        w[j4 * 2] = pack2<kBf16>(s0.x, s0.y);
        w[j4 * 2 + 1] = pack2<kBf16>(s1.x, s1.y);
      }
      mbar_arrive(&bars[kDeltaEmpty + ds]);  // this thread is done with the delta row
      if constexpr (kPairs == 16) tmem_st_32x32b_x16(tDP, *reinterpret_cast<uint32_t(*)[16]>(&w[0]));
      else tmem_st_32x32b_x32(tDP, *reinterpret_cast<uint32_t(*)[32]>(&w[0]));
      tmem_wait_st();
      tcgen05_fence_before();
      mbar_arrive(&bars[kDsFull]);
      if (r == 0 && (wg == 0 || wg == kCmpWgs - 1)) FA2_STAMP(wg ? 1 : 0, it, 4);

      if (kSpill) {
        const int head = head_kv * group + it / m_count;
        if (j_in_head == 0) prefix = ds_tile_prefix(m_start, p.Sq, p.Sk, p.causal);
        const long long tile = (static_cast<long long>(batch) * p.Hq + head) * p.tiles_per_head + prefix + nblk;
        prefix += visible_kv_blocks_of(m, p.Sq, p.Sk, p.causal);
        uint8_t* dst = spill_row + tile * kDsTileBytes;
#pragma unroll
        for (int c = 0; c < kPieces / 2; ++c)
          st_global_cs_v4(dst + c * (kBlk * 16), w[c * 4], w[c * 4 + 1], w[c * 4 + 2], w[c * 4 + 3]);
#pragma unroll
        for (int j = 0; j < kPieces / 2 * 4; ++j) carry[j] = w[kPieces / 2 * 4 + j];
        carry_dst = dst;
      }
    }
    if (kSpill && T > 0) {
#pragma unroll
      for (int c = 0; c < kPieces / 2; ++c)
        st_global_cs_v4(carry_dst + (kPieces / 2 + c) * (kBlk * 16), carry[c * 4], carry[c * 4 + 1], carry[c * 4 + 2], carry[c * 4 + 3]);
    }

    // ---- epilogue: dK * scale, dV -> smem (the K / V tiles are dead once every MMA retired) -> TMA store ----
    if (T > 0) {
      mbar_wait(&bars[kDkvFull], 0);
      tcgen05_fence_after();
    }
    constexpr int kChunksPerOut = D / 32;            // 32-column chunks of dK (and of dV)
    constexpr int kChunks = 2 * kChunksPerOut;       // warpgroup wg takes chunks wg, wg + 4, ...
#pragma unroll
    for (int cc = 0; cc < (kChunks + kCmpWgs - 1) / kCmpWgs; ++cc) {
      const int chunk = wg + cc * kCmpWgs;
      if (chunk < kChunks) {
        const int which = chunk / kChunksPerOut;  // 0: dK, 1: dV
        const int col = (chunk % kChunksPerOut) * 32;
        uint8_t* sOut = (which == 0) ? sK : sV;
        const float mul = (which == 0) ? p.softmax_scale : 1.0f;
        uint32_t o[32];
        if (T > 0) {
          tmem_ld_32x32b_x32(lane_sel + (which == 0 ? kTmemDK : kTmemDV) + col, o);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = 0u;
        }
        if constexpr (kSplit) {
          // fp32 partial of this part, straight from registers: 128 contiguous bytes per thread (key row) and chunk
          if (kv < Sk) {
            float* dst = (which == 0 ? p.part_dk : p.part_dv) +
                         ((static_cast<int64_t>(batch) * Sk + kv) * (p.Hkv * p.kv_split) + head_kv) * D + col;
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4)
              *reinterpret_cast<float4*>(dst + q4 * 4) =
                  make_float4(__uint_as_float(o[q4 * 4 + 0]) * mul, __uint_as_float(o[q4 * 4 + 1]) * mul,
                              __uint_as_float(o[q4 * 4 + 2]) * mul, __uint_as_float(o[q4 * 4 + 3]) * mul);
          }
          continue;
        }
        uint8_t* sub = sOut + (col / 64) * kSubBytes;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 v;
          v.x = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 0]) * mul, __uint_as_float(o[q4 * 8 + 1]) * mul);
          v.y = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 2]) * mul, __uint_as_float(o[q4 * 8 + 3]) * mul);
          v.z = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 4]) * mul, __uint_as_float(o[q4 * 8 + 5]) * mul);
          v.w = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 6]) * mul, __uint_as_float(o[q4 * 8 + 7]) * mul);
          *reinterpret_cast<uint4*>(sub + sw128_offset(r, ((col % 64) / 8) + q4)) = v;
        }
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(1, kCmpThreads);
    if (!kSplit && threadIdx.x == 0) {
      if constexpr (kVarlen) {
        tmap_acquire(mdk);
        tmap_acquire(mdv);
      }
      for (int j = 0; j < kSub; ++j) {
        tma_store_4d(mdk, sK + j * kSubBytes, j * 64, kv0, head_kv, bc);
        tma_store_4d(mdv, sV + j * kSubBytes, j * 64, kv0, head_kv, bc);
      }
      bulk_commit_group();
      bulk_wait_group_read<0>();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tcgen05_fence_after();
    tmem_dealloc<512>(0u);
  }
}

template <int D, bool kBf16, bool kSpill, int kExtra>
static cudaError_t launch_bwd_dkdv_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                                     const CUtensorMap& tdo, const CUtensorMap& tdk, const CUtensorMap& tdv,
                                     const DkdvParams& p, cudaStream_t stream) {
  auto kernel = fa2_bwd_dkdv_kernel<D, kBf16, kSpill, kExtra>;
  constexpr int smem = DkdvSmem<D, q_stages(D, kExtra), kExtra == 1>::kBytes;
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
