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
// CTA = 20 warps:
//   warps 0-15  : four compute warpgroups; thread = query row (TMEM lane), warpgroup g owns key columns [32g, 32g+32).
//                 Four warps per scheduler instead of two: one warp of this code sustains ~0.3 instructions per clock
//                 (dependent FFMA -> MUFU -> FMUL -> F2FP chains), so the block's softmax/dS latency - which sits on the
//                 critical path between the tensor-pipe products - roughly halves (profiles/, ncu stall reasons).
//   warp 16 : MMA issuer       warp 17 : TMA producer       warps 18-19 : idle
// TMEM (512 columns): S ring [0,128) [128,256), dS(i) (16-bit) goes over S(i) | dP [256,384) | dQ [384,384+D)
// Tensor-pipe order:  S(0) dP(0) S(1) | dP(i+1) dQ(i) S(i+2) ...   - S is double buffered, so the exponentials of
// block i+1 start as soon as dS(i) is handed over, and dP(i+1) - the next thing they need - is issued first.
#include "fa2_kernels.cuh"
#include "ptx.cuh"

namespace fa2 {

namespace {

constexpr int kBlk = 128;
constexpr int kDqThreads = 640;
constexpr int kCmpWgs = 4;             // compute warpgroups
constexpr int kCmpCols = kBlk / kCmpWgs;  // key columns per compute thread
constexpr int kKStages = 3;
constexpr int kVStages = 2;

template <int D>
struct DqSmem {
  static constexpr int kTileBytes = kBlk * D * 2;
  static constexpr int kQOff = 0;
  static constexpr int kDOOff = kQOff + kTileBytes;
  static constexpr int kKOff = kDOOff + kTileBytes;
  static constexpr int kVOff = kKOff + kKStages * kTileBytes;
  static constexpr int kBarOff = kVOff + kVStages * kTileBytes;
  // q_full, k_full[3], k_empty[3], v_full[2], v_empty[2], s_full[2], dp_full, ds_full, dq_done
  static constexpr int kNumBars = 1 + 2 * kKStages + 2 * kVStages + 2 + 3;
  static constexpr int kTmemSlotOff = kBarOff + kNumBars * 8;
  static constexpr int kBytes = kTmemSlotOff + 16;
  static_assert(kBytes <= 232448, "dQ kernel smem budget exceeded");
};

__device__ __forceinline__ int visible_kv_blocks(int row0, int Sq, int Sk, int causal) {
  if (row0 >= Sq) return 0;
  const int last_row = min(row0 + kBlk, Sq) - 1;
  const int cols = causal ? min(Sk, last_row + Sk - Sq + 1) : Sk;
  return cols <= 0 ? 0 : (cols + kBlk - 1) / kBlk;
}

}  // namespace

template <int D, bool kBf16>
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
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + kKStages;
  uint64_t* v_full = k_empty + kKStages;
  uint64_t* v_empty = v_full + kVStages;
  uint64_t* s_full = v_empty + kVStages;
  uint64_t* dp_full = s_full + 2;
  uint64_t* ds_full = dp_full + 1;
  uint64_t* dq_done = ds_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlotOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m = static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x);  // long (late) query blocks first
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int head_kv = head / (p.Hq / p.Hkv);
  const int row0 = m * kBlk;
  const int shift = p.Sk - p.Sq;
  const int nblk = visible_kv_blocks(row0, p.Sq, p.Sk, p.causal);

  if ((smem_u32(smem) & 1023u) != 0) __trap();

  if (warp == 17 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_do);
    if (p.dq_is_16bit) tma_prefetch_desc(&tmap_dq);
  }
  if (warp == 16) {
    if (lane == 0) {
      mbar_init(q_full, 1);
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
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (*tmem_slot != 0u) __trap();  // whole-TMEM allocation => base 0 (keeps tcgen05 addresses warp-uniform)

  // register pool: 640 threads x 96 (launch bound) = 61440 >= 4 x 128 x 104 + 128 x 64
  if (warp >= 16) {
    reg_dealloc<64>();
  if (warp == 17) {
    // ========================================= TMA producer ========================================
    if (nblk > 0 && elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * L::kTileBytes);
      for (int j = 0; j < kSub; ++j) {
        tma_load_4d(sQ + j * kSubBytes, &tmap_q, q_full, j * 64, row0, head, batch);
        tma_load_4d(sDO + j * kSubBytes, &tmap_do, q_full, j * 64, row0, head, batch);
      }
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      for (int i = 0; i < nblk; ++i) {
        mbar_wait(&k_empty[ks], kph ^ 1);
        mbar_arrive_expect_tx(&k_full[ks], L::kTileBytes);
        for (int j = 0; j < kSub; ++j)
          tma_load_4d(sK + ks * L::kTileBytes + j * kSubBytes, &tmap_k, &k_full[ks], j * 64, i * kBlk, head_kv, batch);
        mbar_wait(&v_empty[vs], vph ^ 1);
        mbar_arrive_expect_tx(&v_full[vs], L::kTileBytes);
        for (int j = 0; j < kSub; ++j)
          tma_load_4d(sV + vs * L::kTileBytes + j * kSubBytes, &tmap_v, &v_full[vs], j * 64, i * kBlk, head_kv, batch);
        if (++ks == kKStages) { ks = 0; kph ^= 1; }
        if (++vs == kVStages) { vs = 0; vph ^= 1; }
      }
    }
  } else if (warp == 16) {
    // ========================================== MMA issuer =========================================
    if (nblk > 0 && elect_one()) {
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
      // stage / phase of K(i) and V(i) as running counters (no divisions in the loop)
      auto k_stage = [](int i) { return i % kKStages; };
      auto k_phase = [](int i) { return static_cast<uint32_t>((i / kKStages) & 1); };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tcgen05_fence_after();
      issue_kmajor(kTmemS, q_kmaj, k_kmaj);  // S(0)
      umma_commit(&s_full[0]);
      mbar_wait(&v_full[0], 0);
      tcgen05_fence_after();
      issue_kmajor(kTmemDP, do_kmaj, v_kmaj);  // dP(0)
      umma_commit(dp_full);
      umma_commit(&v_empty[0]);
      if (nblk > 1) {
        mbar_wait(&k_full[1], 0);
        tcgen05_fence_after();
        issue_kmajor(kTmemS + 128, q_kmaj, advance(k_kmaj, L::kTileBytes));  // S(1)
        umma_commit(&s_full[1]);
      }
      for (int i = 0; i < nblk; ++i) {
        const int ks = k_stage(i);
        mbar_wait(ds_full, i & 1);  // dS(i) is in tensor memory, and dP(i) / S(i) have been read
        tcgen05_fence_after();
        // dP(i+1) = dO V(i+1)^T first: it is what the compute warps wait for next
        if (i + 1 < nblk) {
          const int vs = (i + 1) % kVStages;
          mbar_wait(&v_full[vs], static_cast<uint32_t>(((i + 1) / kVStages) & 1));
          tcgen05_fence_after();
          issue_kmajor(kTmemDP, do_kmaj, advance(v_kmaj, vs * L::kTileBytes));
          umma_commit(dp_full);
          umma_commit(&v_empty[vs]);
        }
        // dQ += dS(i) K(i)   (A = dS from TMEM: compute warpgroup g packed its 32 columns, k-steps 2g and 2g+1, into
        // the first 16 columns of its slice of the S buffer block i was computed in)
#pragma unroll
        for (int k = 0; k < kBlk / 16; ++k) {
          const uint32_t a_col = (k / 2) * kCmpCols + (k % 2) * 8;
          umma_ts(kTmemDQ, kTmemS + (i & 1) * 128 + a_col, advance(k_mn, ks * L::kTileBytes + k * 2048), idesc_dq,
                  (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&k_empty[ks]);  // K(i) has fed S(i) and dQ(i)
        // S(i+2) = Q K(i+2)^T into the same buffer (ordered after dQ(i) by the tensor pipe)
        if (i + 2 < nblk) {
          const int ks2 = k_stage(i + 2);
          mbar_wait(&k_full[ks2], k_phase(i + 2));
          tcgen05_fence_after();
          issue_kmajor(kTmemS + (i & 1) * 128, q_kmaj, advance(k_kmaj, ks2 * L::kTileBytes));
          umma_commit(&s_full[i & 1]);
        }
      }
      umma_commit(dq_done);
    }
  }
  } else {
    // ====================================== compute warpgroups =====================================
    reg_alloc<104>();
    const int wg = warp >> 2;
    const int r = threadIdx.x & 127;  // query row inside the block == TMEM lane
    const int row = row0 + r;
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tDP = lane_sel + kTmemDP + wg * kCmpCols;
    const float c = p.scale_log2;
    // padded rows carry LSE = delta = 0 (forward / delta kernels); their dQ rows are never stored
    const int64_t stat_off = (static_cast<int64_t>(batch) * p.Hq + head) * p.lse_stride + row;
    const float nlse = -p.lse[stat_off];  // -inf -> +inf for fully masked rows: handled by the mask path
    const float ndelta = -p.delta[stat_off];

    for (int i = 0; i < nblk; ++i) {
      const uint32_t tS = lane_sel + kTmemS + (i & 1) * 128 + wg * kCmpCols;
      const int col0 = i * kBlk + wg * kCmpCols;
      const bool need_mask = (i * kBlk + kBlk > p.Sk) || (p.causal && (i * kBlk + kBlk - 1 > row0 + shift));

      // ---- P = exp2(S c - LSE) ----
      mbar_wait(&s_full[i & 1], (i >> 1) & 1);
      tcgen05_fence_after();
      uint32_t su[32];
      tmem_ld_32x32b_x32(tS, su);
      tmem_wait_ld();
      float pf[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float t = fmaf(__uint_as_float(su[j]), c, nlse);
        // every fourth exponential on the FMA pipes, the others on the MUFU
        pf[j] = ((j & 3) == 3) ? ex2_emul(t) : ex2_approx(t);
      }
      if (need_mask) {
        const int limit = p.causal ? min(p.Sk - 1, row + shift) : (p.Sk - 1);
#pragma unroll
        for (int j = 0; j < 32; ++j) pf[j] = (col0 + j <= limit) ? pf[j] : 0.f;
      }

      // ---- dS = P o (dP - delta), 16-bit, written over the head of this warpgroup's (now dead) S columns, so that
      //      the dP columns are free for dP(i+1) at once ----
      mbar_wait(dp_full, i & 1);
      tcgen05_fence_after();
      uint32_t dpu[32];
      tmem_ld_32x32b_x32(tDP, dpu);
      tmem_wait_ld();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float d0 = pf[2 * j] * (__uint_as_float(dpu[2 * j]) + ndelta);
        const float d1 = pf[2 * j + 1] * (__uint_as_float(dpu[2 * j + 1]) + ndelta);
        pk[j] = pack2<kBf16>(d0, d1);
      }
      tmem_st_32x32b_x16(tS, pk);
      tmem_wait_st();
      tcgen05_fence_before();
      mbar_arrive(ds_full);
    }

    // ---- epilogue: dQ * scale -> caller's tensor ----
    if (row0 < p.Sq) {
      if (nblk > 0) {
        mbar_wait(dq_done, 0);
        tcgen05_fence_after();
      }
      // 32 head-dim columns per warpgroup: D / 32 of the four warpgroups take part
      const uint32_t tDQ = lane_sel + kTmemDQ + wg * 32;
      const float mul = p.softmax_scale;
      if (wg < D / 32) {
        constexpr int cidx = 0;
        uint32_t o[32];
        if (nblk > 0) {
          tmem_ld_32x32b_x32(tDQ + cidx * 32, o);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = 0u;
        }
        const int col = wg * 32;  // first head-dim column of this chunk
        if (p.dq_is_16bit) {
          uint8_t* sub = sQ + (col / 64) * kSubBytes;  // the Q tile is dead once dq_done fired
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint4 v;
            v.x = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 0]) * mul, __uint_as_float(o[q4 * 8 + 1]) * mul);
            v.y = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 2]) * mul, __uint_as_float(o[q4 * 8 + 3]) * mul);
            v.z = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 4]) * mul, __uint_as_float(o[q4 * 8 + 5]) * mul);
            v.w = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 6]) * mul, __uint_as_float(o[q4 * 8 + 7]) * mul);
            *reinterpret_cast<uint4*>(sub + sw128_offset(r, ((col % 64) / 8) + q4)) = v;
          }
        } else if (row < p.Sq) {
          float* dst = reinterpret_cast<float*>(p.dq) + batch * p.dq_strides[0] + head * p.dq_strides[1] +
                       static_cast<int64_t>(row) * p.dq_strides[2] + col;
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            *reinterpret_cast<float4*>(dst + q4 * 4) =
                make_float4(__uint_as_float(o[q4 * 4 + 0]) * mul, __uint_as_float(o[q4 * 4 + 1]) * mul,
                            __uint_as_float(o[q4 * 4 + 2]) * mul, __uint_as_float(o[q4 * 4 + 3]) * mul);
          }
        }
      }
      if (p.dq_is_16bit) {
        // (when nblk == 0 nothing was loaded into the Q tile - q_full is only armed for nblk > 0 - so no load can
        // race the zeros staged above)
        fence_proxy_async_smem();
        named_bar_sync(1, kCmpWgs * 128);
        if (threadIdx.x == 0) {
          for (int j = 0; j < kSub; ++j) tma_store_4d(&tmap_dq, sQ + j * kSubBytes, j * 64, row0, head, batch);
          bulk_commit_group();
          bulk_wait_group_read<0>();
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 16) {
    tcgen05_fence_after();
    tmem_dealloc<512>(0u);
  }
}

template <int D, bool kBf16>
static cudaError_t launch_bwd_dq_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                                   const CUtensorMap& tdo, const CUtensorMap& tdq, const DqParams& p,
                                   cudaStream_t stream) {
  auto kernel = fa2_bwd_dq_kernel<D, kBf16>;
  constexpr int smem = DqSmem<D>::kBytes;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  dim3 grid((p.Sq + kBlk - 1) / kBlk, p.Hq, p.B);
  kernel<<<grid, kDqThreads, smem, stream>>>(tq, tk, tv, tdo, tdq, p);
  return cudaGetLastError();
}

cudaError_t launch_bwd_dq(int D, bool bf16, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                          const CUtensorMap& tdo, const CUtensorMap& tdq, const DqParams& p, cudaStream_t stream) {
  if (D == 128) return bf16 ? launch_bwd_dq_t<128, true>(tq, tk, tv, tdo, tdq, p, stream)
                            : launch_bwd_dq_t<128, false>(tq, tk, tv, tdo, tdq, p, stream);
  if (D == 64) return bf16 ? launch_bwd_dq_t<64, true>(tq, tk, tv, tdo, tdq, p, stream)
                           : launch_bwd_dq_t<64, false>(tq, tk, tv, tdo, tdq, p, stream);
  return cudaErrorInvalidValue;
}

}  // namespace fa2
