// Deterministic backward for head dims 129..256 on sm_100a (B200) - the 207 and 256 of the reference's grid
// (tests/test_fwd_bwd.py:16; the reference serves them through BLOCK_HEADDIM = next_power_of_2, src/forward/caller.py:77).
//
// At these head dims dK and dV alone (2 x 256 fp32 columns) fill the 512 columns of tensor memory, so the 5-product
// schedule of bwd_dkdv_sm100.cu cannot hold S^T and dP^T next to them. The work is cut differently:
//
//   fa2_bwd_dq256_kernel   one CTA per 128-row query block, thread = query row (like the forward): for every 64-key
//                          half-block  S = Q K^T,  dP = dO V^T  (both double-buffered in tensor memory),
//                          P = exp2(S c - LSE2),  dS = P o (dP - delta);  dS (16-bit) goes back over dP and feeds
//                          dQ += dS K  as the A operand from tensor memory (256 fp32 columns, one CTA, ascending key
//                          order: deterministic like `_compute_row_blocks_dq`, src/backward/compute_dq.py:82-261).
//                          P and dS are also SPILLED, 16-bit, as [128 x 128] tiles [16 groups of 8 keys][128 queries][16 B]:
//                          a warp's store covers 512 contiguous bytes, and the tile is the no-swizzle MN-major operand
//                          layout of tcgen05.mma with M = keys (pinned by tests/test_umma_probe.py, a_mode 3).
//   fa2_bwd_acc256_kernel  dK(n) = softmax_scale * sum_m dS(m, n)^T Q(m)  and  dV(n) = sum_m P(m, n)^T dO(m)  from those
//                          tiles: persistent, one product per block pair and output, 256-column accumulators, the
//                          query heads of a kv head summed in tensor memory in a fixed order
//                          (`_compute_column_blocks_dkdv`, src/backward/compute_dkdv.py:116-296 + the GQA sum of
//                          src/backward/caller.py:163-165).
//
// 5 products per block pair like the spilled-dS path at head dims <= 128; the price is 2 x 32 KB of spill per pair
// (dS is read once, P once). Not a BASELINE configuration: built for coverage of the reference's grid, measured in
// profiles/r02_widened_rows.jsonl.
#include "fa2_kernels.cuh"
#include "ptx.cuh"

namespace fa2 {

namespace {

constexpr int kBlk = 128;
constexpr int kHB = 64;
constexpr int kD = 256;                       // instantiated head dim
constexpr int kSub = kD / 64;                 // 64-column sub-tiles per row
constexpr int kQSubBytes = kBlk * 128;        // [128 rows x 64] 16-bit
constexpr int kKvSubBytes = kHB * 128;        // [64 keys x 64]
constexpr int kQTileBytes = kBlk * kD * 2;    // 64 KB
constexpr int kKvTileBytes = kHB * kD * 2;    // 32 KB

// ------------------------------------------------------------------------------------------------
// dQ + spill kernel
// ------------------------------------------------------------------------------------------------
struct Dq256Smem {
  static constexpr int kKStages = 2;  // K(j) lives from S(j) to dQ(j)
  static constexpr int kQOff = 0;
  static constexpr int kDOOff = kQOff + kQTileBytes;
  static constexpr int kKOff = kDOOff + kQTileBytes;
  static constexpr int kVOff = kKOff + kKStages * kKvTileBytes;
  static constexpr int kBarOff = kVOff + kKvTileBytes;
  // qdo_full, k_full[2], k_empty[2], v_full, v_empty, s_full[2], s_free[2], dp_full[2], ds_full[2], dq_full
  static constexpr int kNumBars = 1 + 4 + 2 + 8 + 1;
  static constexpr int kTmemSlotOff = kBarOff + kNumBars * 8;
  static constexpr int kBytes = kTmemSlotOff + 16;
  static_assert(kBytes <= 232448, "dQ (head dim 256) kernel smem budget exceeded");
};
constexpr int kDqThreads = 256;  // warps 0-3 compute (thread = query row), warp 4 MMA issuer, warp 5 TMA producer

__device__ __forceinline__ int half_blocks_of(int row0, int Sq, int Sk, int causal) {
  if (row0 >= Sq) return 0;
  const int last_row = min(row0 + kBlk, Sq) - 1;
  const int cols = causal ? min(Sk, last_row + Sk - Sq + 1) : Sk;
  return cols <= 0 ? 0 : (cols + kHB - 1) / kHB;
}

__device__ __forceinline__ void st_global_cs_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

}  // namespace

template <bool kBf16, bool kBias>
__global__ void __launch_bounds__(kDqThreads, 1)
fa2_bwd_dq256_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                     const Bwd256Params p) {
  using L = Dq256Smem;
  constexpr uint32_t kTmemS = 0, kTmemDP = 128, kTmemDQ = 256;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem + L::kQOff;
  uint8_t* sDO = smem + L::kDOOff;
  uint8_t* sK = smem + L::kKOff;
  uint8_t* sV = smem + L::kVOff;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* qdo_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + 2;
  uint64_t* v_full = k_empty + 2;
  uint64_t* v_empty = v_full + 1;
  uint64_t* s_full = v_empty + 1;   // [2] S(j) is in tensor memory
  uint64_t* s_free = s_full + 2;    // [2] every compute thread has its row of S(j) in registers
  uint64_t* dp_full = s_free + 2;   // [2]
  uint64_t* ds_full = dp_full + 2;  // [2] dS(j) (16-bit) is in tensor memory, over dP(j)
  uint64_t* dq_full = ds_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlotOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m = static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x);  // causal: long query blocks first
  const int head = blockIdx.y, batch = blockIdx.z;
  const int head_kv = head / (p.Hq / p.Hkv);
  const int row0 = m * kBlk;
  const int Sq = p.Sq, Sk = p.Sk;
  const int nh = half_blocks_of(row0, Sq, Sk, p.causal);

  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (warp == 5 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_do);
  }
  if (warp == 4) {
    if (lane == 0) {
      mbar_init(qdo_full, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&k_full[s], 1);
        mbar_init(&k_empty[s], 1);
        mbar_init(&s_full[s], 1);
        mbar_init(&s_free[s], kBlk);
        mbar_init(&dp_full[s], 1);
        mbar_init(&ds_full[s], kBlk);
      }
      mbar_init(v_full, 1);
      mbar_init(v_empty, 1);
      mbar_init(dq_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (*tmem_slot != 0u) __trap();  // whole-TMEM allocation => base 0 (keeps tcgen05 addresses warp-uniform)

  if (warp == 5) {
    // ========================================= TMA producer ========================================
    if (nh > 0 && elect_one()) {
      mbar_arrive_expect_tx(qdo_full, 2 * kQTileBytes);
      for (int j = 0; j < kSub; ++j) {
        tma_load_4d(sQ + j * kQSubBytes, &tmap_q, qdo_full, j * 64, row0, head, batch);
        tma_load_4d(sDO + j * kQSubBytes, &tmap_do, qdo_full, j * 64, row0, head, batch);
      }
      for (int j = 0; j < nh; ++j) {
        const int s = j & 1;
        mbar_wait(&k_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[s], kKvTileBytes);
        for (int c = 0; c < kSub; ++c)
          tma_load_4d(sK + s * kKvTileBytes + c * kKvSubBytes, &tmap_k, &k_full[s], c * 64, j * kHB, head_kv, batch);
        mbar_wait(v_empty, (j & 1) ^ 1);
        mbar_arrive_expect_tx(v_full, kKvTileBytes);
        for (int c = 0; c < kSub; ++c)
          tma_load_4d(sV + c * kKvSubBytes, &tmap_v, v_full, c * 64, j * kHB, head_kv, batch);
      }
    }
  } else if (warp == 4) {
    // ========================================== MMA issuer =========================================
    if (nh > 0 && elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(kBlk, kHB, kBf16, false, false);          // S, dP: [128 x 64], K-major x K-major
      const uint32_t idesc_dq = make_idesc_f16(kBlk, p.d_cols, kBf16, false, true);          // dQ: A from TMEM, B = K MN-major
      const int d_steps = p.d_steps;
      const SmemDesc q_desc = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const SmemDesc do_desc = make_smem_desc_sw128(smem_u32(sDO), 16, 1024);
      const SmemDesc k_desc = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const SmemDesc v_desc = make_smem_desc_sw128(smem_u32(sV), 16, 1024);
      const SmemDesc k_mn = make_smem_desc_sw128(smem_u32(sK), kKvSubBytes, 1024);
      auto issue_kk = [&](uint32_t d_tmem, SmemDesc a, SmemDesc b) {  // [128 x D] x [64 x D]^T, reduction over D
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) {
          if (k < d_steps)
            umma_ss(d_tmem, advance(a, (k / 4) * kQSubBytes + (k % 4) * 32), advance(b, (k / 4) * kKvSubBytes + (k % 4) * 32),
                    idesc_s, k > 0 ? 1u : 0u);
        }
      };
      auto issue_dq = [&](int j) {  // dQ += dS(j) [128 x 64, tensor memory] K(j) [64 x D, MN-major]
        const SmemDesc kb = advance(k_mn, (j & 1) * kKvTileBytes);
#pragma unroll
        for (int k = 0; k < kHB / 16; ++k)
          umma_ts(kTmemDQ, kTmemDP + (j & 1) * kHB + k * 8, advance(kb, k * 2048), idesc_dq, (j > 0 || k > 0) ? 1u : 0u);
      };
      mbar_wait(qdo_full, 0);
      tcgen05_fence_after();
      for (int j = 0; j < nh; ++j) {
        const int b = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        // S(j): needs K(j) and the S buffer read out (S(j - 2))
        mbar_wait(&k_full[b], ph);
        mbar_wait(&s_free[b], ph ^ 1);
        tcgen05_fence_after();
        issue_kk(kTmemS + b * kHB, q_desc, advance(k_desc, b * kKvTileBytes));
        umma_commit(&s_full[b]);
        // dP(j): its buffer held dS(j - 2), read by dQ(j - 2) which was issued before (the tensor pipe runs in order)
        mbar_wait(v_full, j & 1);
        tcgen05_fence_after();
        issue_kk(kTmemDP + b * kHB, do_desc, v_desc);
        umma_commit(&dp_full[b]);
        umma_commit(v_empty);
        // dQ(j - 1)
        if (j > 0) {
          mbar_wait(&ds_full[b ^ 1], ((j - 1) >> 1) & 1);
          tcgen05_fence_after();
          issue_dq(j - 1);
          umma_commit(&k_empty[b ^ 1]);
        }
      }
      mbar_wait(&ds_full[(nh - 1) & 1], ((nh - 1) >> 1) & 1);
      tcgen05_fence_after();
      issue_dq(nh - 1);
      umma_commit(dq_full);
    }
  } else if (warp < 4) {
    // ========================================== compute ============================================
    const int r = threadIdx.x;  // query row inside the block == TMEM lane
    const int row = row0 + r;
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    const float c = p.scale_log2;
    const int shift = Sk - Sq;
    const int64_t stat_off = (static_cast<int64_t>(batch) * p.Hq + head) * p.lse_stride + row;
    const bool row_ok = row < Sq;
    const float lse_r = row_ok ? p.lse[stat_off] : -INFINITY;
    const float delta_r = row_ok ? p.delta[stat_off] : 0.f;
    const bool row_live = row_ok && lse_r != -INFINITY;  // a fully hidden row has LSE = -inf: P = 0
    const float neg_lse = row_live ? -lse_r : 0.f;
    const int limit = !row_live ? -1 : (p.causal ? min(Sk - 1, row + shift) : Sk - 1);
    // spilled tiles of this (batch, head, query block): consecutive, key block n at tile0 + n
    const long long tile0 = (static_cast<long long>(batch) * p.Hq + head) * p.tiles_per_head + ds_tile_prefix(m, Sq, Sk, p.causal);
    uint8_t* ds_base = static_cast<uint8_t*>(p.ds_spill) + tile0 * kDsTileBytes + r * 16;
    uint8_t* p_base = static_cast<uint8_t*>(p.p_spill) + tile0 * kDsTileBytes + r * 16;

    for (int j = 0; j < nh; ++j) {
      const int b = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int col0 = j * kHB;
      [[maybe_unused]] float bias_v[kBias ? kHB : 1];
      if constexpr (kBias) {
        const int64_t off = batch * p.bias.sb + head * p.bias.sh + static_cast<int64_t>(row) * p.bias.sm + col0;
        if (p.bias.is_fp32) {
          const float* bp = static_cast<const float*>(p.bias.ptr) + off;
#pragma unroll
          for (int e = 0; e < kHB; ++e) bias_v[e] = (row_ok && col0 + e < Sk) ? __ldg(bp + e) : 0.f;
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

      // ---- P = exp2(S c (+ bias log2 e) - LSE2) ----
      mbar_wait(&s_full[b], ph);
      tcgen05_fence_after();
      float pf[kHB];
      {
        uint32_t su[2][32];
        tmem_ld_32x32b_x32(lane_sel + kTmemS + b * kHB, su[0]);
        tmem_ld_32x32b_x32(lane_sel + kTmemS + b * kHB + 32, su[1]);
        tmem_wait_ld();
        tcgen05_fence_before();
        mbar_arrive(&s_free[b]);
#pragma unroll
        for (int e = 0; e < kHB; ++e) {
          float x = fmaf(__uint_as_float(su[e / 32][e % 32]), c, neg_lse);
          if constexpr (kBias) x = fmaf(bias_v[e], 1.4426950408889634f, x);
          pf[e] = (col0 + e <= limit) ? ex2_approx(x) : 0.f;
        }
      }
      uint32_t pk[kHB / 2];
#pragma unroll
      for (int e = 0; e < kHB; e += 2) pk[e / 2] = pack2<kBf16>(pf[e], pf[e + 1]);
      {
        uint8_t* dst = p_base + static_cast<long long>(j >> 1) * kDsTileBytes + (j & 1) * (8 * kBlk * 16);
#pragma unroll
        for (int g = 0; g < 8; ++g) st_global_cs_v4(dst + g * (kBlk * 16), pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
      }

      // ---- dS = P o (dP - delta), 16-bit, back over dP for dQ += dS K ----
      mbar_wait(&dp_full[b], ph);
      tcgen05_fence_after();
      {
        uint32_t du[2][32];
        tmem_ld_32x32b_x32(lane_sel + kTmemDP + b * kHB, du[0]);
        tmem_ld_32x32b_x32(lane_sel + kTmemDP + b * kHB + 32, du[1]);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < kHB; e += 2)
          pk[e / 2] = pack2<kBf16>(pf[e] * (__uint_as_float(du[e / 32][e % 32]) - delta_r),
                                   pf[e + 1] * (__uint_as_float(du[(e + 1) / 32][(e + 1) % 32]) - delta_r));
      }
      tmem_st_32x32b_x32(lane_sel + kTmemDP + b * kHB, pk);
      tmem_wait_st();
      tcgen05_fence_before();
      mbar_arrive(&ds_full[b]);
      {
        uint8_t* dst = ds_base + static_cast<long long>(j >> 1) * kDsTileBytes + (j & 1) * (8 * kBlk * 16);
#pragma unroll
        for (int g = 0; g < 8; ++g) st_global_cs_v4(dst + g * (kBlk * 16), pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
      }
    }
    if (nh & 1) {  // the key block of the last half-block is spilled as a whole tile: its unvisited half holds zeros
      const long long off = static_cast<long long>(nh >> 1) * kDsTileBytes + 8 * kBlk * 16;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        st_global_cs_v4(p_base + off + g * (kBlk * 16), 0u, 0u, 0u, 0u);
        st_global_cs_v4(ds_base + off + g * (kBlk * 16), 0u, 0u, 0u, 0u);
      }
    }

    // ---- epilogue: dQ * softmax_scale -> caller's tensor, straight from registers ----
    if (nh > 0) {
      mbar_wait(dq_full, 0);
      tcgen05_fence_after();
    }
    const float mul = p.softmax_scale;
    const int64_t off0 = batch * p.dq_strides[0] + head * p.dq_strides[1] + static_cast<int64_t>(row) * p.dq_strides[2];
#pragma unroll 1
    for (int chunk = 0; chunk < kD / 32; ++chunk) {
      if (chunk * 32 >= p.d_actual) break;
      uint32_t o[32];
      if (nh > 0) {
        tmem_ld_32x32b_x32(lane_sel + kTmemDQ + chunk * 32, o);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) o[e] = 0u;  // fully hidden query rows: dQ = 0 (src/backward/compute_dq.py:248-249)
      }
      if (row_ok && p.dq != nullptr) {  // dq == nullptr: dQ not wanted (dK / dV only), the spills were the point
        const int64_t off = off0 + chunk * 32;
        if (p.dq_is_16bit) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.dq) + off);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            if (chunk * 32 + q4 * 8 >= p.d_actual) break;
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

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) {
    tcgen05_fence_after();
    tmem_dealloc<512>(0u);
  }
}

// ------------------------------------------------------------------------------------------------
// dK / dV from the spilled tiles
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int kAccThreads = 192;  // warps 0-3 epilogue (thread = key row), warp 4 MMA issuer, warp 5 TMA producer + scheduler
constexpr int kSchedSlots = 4;

struct Acc256Smem {
  static constexpr int kStages = 2;
  static constexpr int kStageBytes = kDsTileBytes + kQTileBytes;  // one 32 KB tile + one [128 x 256] Q / dO tile
  static constexpr int kBarOff = kStages * kStageBytes;
  // full[S], empty[S], acc_full[2], acc_empty[2], item_full[4], item_empty[4]
  static constexpr int kNumBars = 2 * kStages + 4 + 2 * kSchedSlots;
  static constexpr int kSchedOff = kBarOff + kNumBars * 8;
  static constexpr int kTmemSlotOff = kSchedOff + kSchedSlots * 4;
  static constexpr int kBytes = kTmemSlotOff + 16;
  static_assert(kBytes <= 232448, "dK/dV (head dim 256) kernel smem budget exceeded");
};

// work item w -> (which output, key block, kv head, batch): both outputs of a key block are consecutive, key block 0
// (seen by the most query blocks under the causal mask) first
struct AccItem {
  int which, n, head_kv, batch;  // which: 0 = dK (dS tiles x Q), 1 = dV (P tiles x dO)
};
__device__ __forceinline__ AccItem acc_item_of(int w, int n_kv_blocks, int Hkv) {
  AccItem it;
  it.which = w & 1;
  w >>= 1;
  const int bh = w / n_kv_blocks;
  it.n = w - bh * n_kv_blocks;
  it.head_kv = bh % Hkv;
  it.batch = bh / Hkv;
  return it;
}

}  // namespace

template <bool kBf16>
__global__ void __launch_bounds__(kAccThreads, 1)
fa2_bwd_acc256_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                      const Bwd256Params p) {
  using L = Acc256Smem;
  constexpr int kStages = L::kStages;

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
  const int n_kv_blocks = (p.Sk + kBlk - 1) / kBlk;
  const int n_items = 2 * n_kv_blocks * p.Hkv * p.B;
  const int group = p.Hq / p.Hkv;
  const int shift = p.Sk - p.Sq;
  // first query block that sees key block n (bottom-right aligned causal mask), as in bwd_dkdv_sm100.cu
  auto m_start_of = [&](int n) { return p.causal ? max(0, (n * kBlk - shift) / kBlk) : 0; };

  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (warp == 5 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_do);
  }
  if (warp == 4) {
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
  if (*tmem_slot != 0u) __trap();

  if (warp == 5) {
    // ===================================== TMA producer + scheduler ================================
    if (elect_one()) {
      const uint64_t policy = l2_policy_evict_first();  // every spilled tile is read once per output
      uint32_t c = 0;
      for (uint32_t t = 0;; ++t) {
        const uint32_t slot = t % kSchedSlots;
        mbar_wait(&item_empty[slot], ((t / kSchedSlots) & 1) ^ 1);
        const int w = atomicAdd(p.sched, 1);
        // every CTA draws exactly one id past the end; the very last draw resets the counter for the next launch
        if (w == n_items + static_cast<int>(gridDim.x) - 1) atomicExch(p.sched, 0);
        sched[slot] = (w < n_items) ? w : -1;
        mbar_arrive(&item_full[slot]);
        if (w >= n_items) break;
        const AccItem it = acc_item_of(w, n_kv_blocks, p.Hkv);
        const uint8_t* base = static_cast<const uint8_t*>(it.which == 0 ? p.ds_spill : p.p_spill);
        const CUtensorMap* mb = it.which == 0 ? &tmap_q : &tmap_do;
        const int m0 = m_start_of(it.n);
        for (int g = 0; g < group; ++g) {
          const int head = it.head_kv * group + g;
          long long tile = (static_cast<long long>(it.batch) * p.Hq + head) * p.tiles_per_head +
                           ds_tile_prefix(m0, p.Sq, p.Sk, p.causal) + it.n;
          for (int mm = m0; mm < n_q_blocks; ++mm, ++c) {
            const uint32_t s = c % kStages;
            uint8_t* stage = smem + s * L::kStageBytes;
            mbar_wait(&empty[s], ((c / kStages) & 1) ^ 1);
            mbar_arrive_expect_tx(&full[s], L::kStageBytes);
            bulk_load_1d_hint(stage, base + tile * kDsTileBytes, kDsTileBytes, &full[s], policy);
            for (int j = 0; j < kSub; ++j)
              tma_load_4d(stage + kDsTileBytes + j * kQSubBytes, mb, &full[s], j * 64, mm * kBlk, head, it.batch);
            tile += visible_kv_blocks_of(mm, p.Sq, p.Sk, p.causal);
          }
        }
      }
    }
  } else if (warp == 4) {
    // ========================================== MMA issuer =========================================
    if (elect_one()) {
      // out[kv, d] (+)= sum_q T[kv, q] X[q, d]: both operands MN-major (the rows of the Q / dO tile and the 16-byte
      // pieces of the spilled tile are the reduction index)
      const uint32_t idesc = make_idesc_f16(kBlk, p.d_cols, kBf16, true, true);
      uint32_t c = 0;
      for (uint32_t t = 0;; ++t) {
        const uint32_t slot = t % kSchedSlots;
        mbar_wait(&item_full[slot], (t / kSchedSlots) & 1);
        const int w = sched[slot];
        mbar_arrive(&item_empty[slot]);
        if (w < 0) break;
        const AccItem it = acc_item_of(w, n_kv_blocks, p.Hkv);
        const int iters = group * max(0, n_q_blocks - m_start_of(it.n));
        const uint32_t buf = t & 1;
        const uint32_t acc = buf * 256;
        mbar_wait(&acc_empty[buf], ((t >> 1) & 1) ^ 1);  // the epilogue of item t - 2 has read this accumulator
        tcgen05_fence_after();
        for (int i = 0; i < iters; ++i, ++c) {
          const uint32_t s = c % kStages;
          mbar_wait(&full[s], (c / kStages) & 1);
          tcgen05_fence_after();
          const uint32_t stage = smem_u32(smem + s * L::kStageBytes);
          // spilled tile [16 groups of 8 keys][128 queries][16 B] = MN-major without swizzle, M = keys: 8-query groups
          // (the reduction index) are 128 B apart (LBO), 8-key groups 2 KB (SBO); one k-step = 16 queries = 256 B
          const SmemDesc t_mn = make_smem_desc_noswizzle(stage, 128, kBlk * 16);
          const SmemDesc x_mn = make_smem_desc_sw128(stage + kDsTileBytes, kQSubBytes, 1024);
#pragma unroll
          for (int k = 0; k < kBlk / 16; ++k)
            umma_ss(acc, advance(t_mn, k * 256), advance(x_mn, k * 2048), idesc, (i > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ========================================== epilogue ===========================================
    const int r = threadIdx.x & 127;  // key row inside the block == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    for (uint32_t t = 0;; ++t) {
      const uint32_t slot = t % kSchedSlots;
      mbar_wait(&item_full[slot], (t / kSchedSlots) & 1);
      const int w = sched[slot];
      mbar_arrive(&item_empty[slot]);
      if (w < 0) break;
      const AccItem it = acc_item_of(w, n_kv_blocks, p.Hkv);
      const int iters = group * max(0, n_q_blocks - m_start_of(it.n));
      const int kv = it.n * kBlk + r;
      const uint32_t buf = t & 1;
      mbar_wait(&acc_full[buf], (t >> 1) & 1);
      tcgen05_fence_after();
      const float mul = it.which == 0 ? p.softmax_scale : 1.0f;
      const int64_t* st = it.which == 0 ? p.dk_strides : p.dv_strides;
      uint16_t* out = reinterpret_cast<uint16_t*>(it.which == 0 ? p.dk : p.dv) + it.batch * st[0] + it.head_kv * st[1] +
                      static_cast<int64_t>(kv) * st[2];
#pragma unroll 1
      for (int chunk = 0; chunk < kD / 32; ++chunk) {
        uint32_t o[32];
        if (iters > 0 && chunk * 32 < p.d_actual) {
          tmem_ld_32x32b_x32(lane_sel + buf * 256 + chunk * 32, o);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) o[e] = 0u;  // a key block no query sees: dK = dV = 0
        }
        if (chunk == kD / 32 - 1) {
          tcgen05_fence_before();
          mbar_arrive(&acc_empty[buf]);
        }
        if (kv < p.Sk) {
          uint4* dst = reinterpret_cast<uint4*>(out + chunk * 32);
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
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) {
    tcgen05_fence_after();
    tmem_dealloc<512>(0u);
  }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
template <bool kBf16, bool kBias>
static cudaError_t launch_dq256_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                                  const Bwd256Params& p, cudaStream_t stream) {
  if (cudaError_t e = ensure_dynamic_smem<fa2_bwd_dq256_kernel<kBf16, kBias>>(Dq256Smem::kBytes); e != cudaSuccess) return e;
  dim3 grid((p.Sq + kBlk - 1) / kBlk, p.Hq, p.B);
  fa2_bwd_dq256_kernel<kBf16, kBias><<<grid, kDqThreads, Dq256Smem::kBytes, stream>>>(tq, tk, tv, tdo, p);
  return cudaGetLastError();
}

cudaError_t launch_bwd_dq256(bool bf16, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                             const CUtensorMap& tdo, const Bwd256Params& p, cudaStream_t stream) {
  const bool bias = p.bias.ptr != nullptr;
  if (bf16) return bias ? launch_dq256_t<true, true>(tq, tk, tv, tdo, p, stream) : launch_dq256_t<true, false>(tq, tk, tv, tdo, p, stream);
  return bias ? launch_dq256_t<false, true>(tq, tk, tv, tdo, p, stream) : launch_dq256_t<false, false>(tq, tk, tv, tdo, p, stream);
}

template <bool kBf16>
static cudaError_t launch_acc256_t(const CUtensorMap& tq, const CUtensorMap& tdo, const Bwd256Params& p, cudaStream_t stream) {
  if (cudaError_t e = ensure_dynamic_smem<fa2_bwd_acc256_kernel<kBf16>>(Acc256Smem::kBytes); e != cudaSuccess) return e;
  const int sms = current_sm_count();
  const int n_items = 2 * ((p.Sk + kBlk - 1) / kBlk) * p.Hkv * p.B;
  dim3 grid(n_items < sms ? n_items : sms);  // persistent: one CTA per SM
  fa2_bwd_acc256_kernel<kBf16><<<grid, kAccThreads, Acc256Smem::kBytes, stream>>>(tq, tdo, p);
  return cudaGetLastError();
}

cudaError_t launch_bwd_acc256(bool bf16, const CUtensorMap& tq, const CUtensorMap& tdo, const Bwd256Params& p,
                              cudaStream_t stream) {
  return bf16 ? launch_acc256_t<true>(tq, tdo, p, stream) : launch_acc256_t<false>(tq, tdo, p, stream);
}

}  // namespace fa2
