// FlashAttention-2 backward for sm_100a (B200).
//
//   fa2_delta_kernel      delta = rowsum(O * dO)                 <- `_compute_delta`  src/backward/compute_delta.py:18-73
//   fa2_bwd_kernel        dK, dV and the dQ partial sums         <- `_bwd_kernel`     src/backward/kernel.py:65-182
//                                                                   (compute_dkdv.py:8-112, compute_dq.py:8-78)
//   fa2_dq_convert_kernel fp32 dQ accumulator -> caller's dQ     <- the fp32 `dq` of  src/backward/caller.py:86
//
// Arithmetic contract: P = exp2(S*c - LSE2) recomputed from the saved log2-domain LSE,
// dV = P^T dO with P rounded to the input dtype, dP = dO V^T, dS = P o (dP - delta) rounded to the input dtype,
// dK = dS^T Q * scale, dQ = dS K * scale. (The reference folds `scale` into dS before rounding; we round dS and
// apply `scale` to the fp32 sums - the same value up to one rounding of a 16-bit number.)
//
// Unlike the reference (one program per K/V block for dK/dV plus one program per Q block that recomputes
// everything for dQ, 7 matmuls) this kernel does the minimal 5 matmuls: one CTA owns a 128-row K/V block, walks
// the query blocks that see it and adds its dQ contribution into an fp32 accumulator with TMA bulk reductions.
// Deterministic mode serialises those additions per query tile in a fixed key-block order with a ticket.
//
// CTA = 16 warps:
//   warps 0-3 / 4-7 : compute warpgroups; thread = key row (TMEM lane), warpgroup w owns query columns [64w, 64w+64)
//   warps 8-11      : dQ drain: TMEM -> registers -> smem staging -> cp.reduce.async.bulk (fp32 add) into global
//   warp 12         : MMA issuer          warp 13 : TMA producer          warps 14-15 : idle
// TMEM (512 columns): S^T / P^T [0,128) | dV [128,128+D) | dK [256,256+D) | dP^T, then dQ [384,512)
// MMA issue order per query block i:  dV(i) | dP(i) | S(i+1) | dK(i) | dQ(i)
#include "fa2_kernels.cuh"
#include "ptx.cuh"

namespace fa2 {

#ifdef FA2_TRACE
// Debug build only (make TRACE=1): clock64() stamps of one CTA's pipeline events, [role][iteration][event].
__device__ long long g_bwd_trace[4 * 64 * 8];
#define FA2_STAMP(role, it, ev)                                                                          \
  do {                                                                                                   \
    if (blockIdx.x == gridDim.x - 1 && blockIdx.y == 0 && blockIdx.z == 0 && (it) < 64)                  \
      g_bwd_trace[((role) * 64 + (it)) * 8 + (ev)] = clock64();                                          \
  } while (0)
cudaError_t fetch_bwd_trace(long long* host, size_t n) {
  return cudaMemcpyFromSymbol(host, g_bwd_trace, n * sizeof(long long));
}
#else
#define FA2_STAMP(role, it, ev) do { } while (0)
#endif

namespace {

constexpr int kBlk = 128;  // rows per Q block and per K/V block
constexpr int kBwdThreads = 512;
// register re-partition (setmaxnreg): 3 math warpgroups x 128 x 152 + 1 issue warpgroup x 128 x 56 = 65536
constexpr int kRegsMath = 152;
constexpr int kRegsIssue = 56;

template <int D>
struct BwdSmem {
  static constexpr int kQStages = 2;
  static constexpr int kTileBytes = kBlk * D * 2;     // K, V, Q, dO tiles
  static constexpr int kDsBytes = kBlk * kBlk * 2;    // dS^T tile: 2 sub-tiles [128 kv x 64 q]
  static constexpr int kChunkBytes = kBlk * 32 * 4;   // dQ staging chunk: [128 q x 32 d] fp32 = 16 KB
  static constexpr int kKOff = 0;
  static constexpr int kVOff = kKOff + kTileBytes;
  static constexpr int kQOff = kVOff + kTileBytes;                // kQStages stages
  static constexpr int kDOOff = kQOff + kQStages * kTileBytes;    // 1 stage
  static constexpr int kDsOff = kDOOff + kTileBytes;
  static constexpr int kStageOff = kDsOff + kDsBytes;             // 2 chunks
  static constexpr int kLseOff = kStageOff + 2 * kChunkBytes;     // kQStages x 512 B (rides with Q)
  static constexpr int kDeltaOff = kLseOff + kQStages * kBlk * 4; // 2 x 512 B (rides with dO, own barriers)
  static constexpr int kBarOff = kDeltaOff + 2 * kBlk * 4;
  static constexpr int kNumBars = 22;
  static constexpr int kTmemSlotOff = kBarOff + kNumBars * 8;
  static constexpr int kBytes = kTmemSlotOff + 16;
  static_assert(kBytes <= 232448, "backward smem budget exceeded");
};

enum BwdBar {
  kKvFull = 0,
  kQFull = 1,   // [2]
  kQEmpty = 3,  // [2]
  kStageFull = 5,   // [2] dQ staging chunk written by the drain warps
  kDoFull = 7,
  kDoEmpty = 8,
  kSFull = 9,
  kPFull = 10,
  kDpFull = 11,
  kDsFull = 12,
  kDqFull = 13,
  kDqEmpty = 14,
  kDkvFull = 15,
  kDeltaFull = 16,   // [2]
  kStageEmpty = 18,  // [2] the bulk reduction has finished reading the chunk
};

__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

}  // namespace

// =================================================================================================
// delta = rowsum(O * dO)
// =================================================================================================
template <int D, bool kBf16>
__global__ void __launch_bounds__(256) fa2_delta_kernel(const DeltaParams p) {
  constexpr int kThreadsPerRow = D / 8;
  constexpr int kRowsPerPass = 256 / kThreadsPerRow;
  constexpr int kPasses = 4;  // rows per thread: all 2 x kPasses 16-byte loads are issued before the first use
  constexpr int kRowsPerBlock = kRowsPerPass * kPasses;
  pdl_wait();  // O comes from the forward kernel that may still be running (programmatic dependent launch)
  if (threadIdx.x == 0) pdl_launch_dependents();
  const int head = blockIdx.y, batch = blockIdx.z;
  const int sq_valid = p.seqlens != nullptr ? min(max(p.seqlens[batch], 0), p.Sq) : p.Sq;  // variable-length mode: this sequence's rows
  const int row0 = blockIdx.x * kRowsPerBlock + threadIdx.x / kThreadsPerRow;
  const int part = threadIdx.x % kThreadsPerRow;
  const uint16_t* obase = reinterpret_cast<const uint16_t*>(p.o) + batch * p.o_strides[0] + head * p.o_strides[1] + part * 8;
  const uint16_t* gbase = reinterpret_cast<const uint16_t*>(p.dout) + batch * p.do_strides[0] + head * p.do_strides[1] + part * 8;
  uint4 ov[kPasses], gv[kPasses];
#pragma unroll
  for (int i = 0; i < kPasses; ++i) {
    const int row = row0 + i * kRowsPerPass;
    ov[i] = gv[i] = make_uint4(0u, 0u, 0u, 0u);
    if (row < sq_valid && part * 8 < p.D) {  // p.D < D: the tensors' head dim is zero-padded up to the kernel's
      ov[i] = *reinterpret_cast<const uint4*>(obase + static_cast<int64_t>(row) * p.o_strides[2]);
      gv[i] = *reinterpret_cast<const uint4*>(gbase + static_cast<int64_t>(row) * p.do_strides[2]);
    }
  }
#pragma unroll
  for (int i = 0; i < kPasses; ++i) {
    const int row = row0 + i * kRowsPerPass;
    const uint32_t ow[4] = {ov[i].x, ov[i].y, ov[i].z, ov[i].w};
    const uint32_t gw[4] = {gv[i].x, gv[i].y, gv[i].z, gv[i].w};
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 a, b;
      if constexpr (kBf16) {
        a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ow[j]));
        b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[j]));
      } else {
        a = __half22float2(*reinterpret_cast<const __half2*>(&ow[j]));
        b = __half22float2(*reinterpret_cast<const __half2*>(&gw[j]));
      }
      acc = fmaf(a.x, b.x, acc);
      acc = fmaf(a.y, b.y, acc);
    }
#pragma unroll
    for (int off = kThreadsPerRow / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (part == 0 && row < p.lse_stride)
      p.delta[(static_cast<int64_t>(batch) * p.Hq + head) * p.lse_stride + row] = acc;  // padded rows: 0
  }
  // Clear this CTA's share of the backward workspace (fp32 dQ accumulator + ordered-reduction tickets), so the
  // fused kernel that follows on the stream needs no separate memset: kRowsPerBlock * D floats.
  if (p.dq_accum != nullptr) {
    const int64_t bh = static_cast<int64_t>(batch) * p.Hq + head;
    const int rows_here = min(kRowsPerBlock, p.lse_stride - static_cast<int>(blockIdx.x) * kRowsPerBlock);
    float4* dst = reinterpret_cast<float4*>(p.dq_accum + (bh * p.lse_stride + static_cast<int64_t>(blockIdx.x) * kRowsPerBlock) * D);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = threadIdx.x; i < rows_here * D / 4; i += 256) dst[i] = z;
    if (blockIdx.x == 0 && p.dq_sem != nullptr) {
      const int tiles = p.lse_stride / 128;
      for (int i = threadIdx.x; i < tiles; i += 256) p.dq_sem[bh * tiles + i] = 0;
      if (bh == 0 && threadIdx.x == 0) p.dq_sem[static_cast<int64_t>(p.B) * p.Hq * tiles] = 0;  // the CTA-order counter
    }
  }
}

cudaError_t launch_delta(bool bf16, const DeltaParams& p, cudaStream_t stream) {
  const int dk = p.D <= 64 ? 64 : p.D <= 128 ? 128 : 256;  // instantiation; head dims in between are handled by the column guard
  const int rows_per_block = 4 * (256 / (dk / 8));
  dim3 grid((p.lse_stride + rows_per_block - 1) / rows_per_block, p.Hq, p.B);
  if (p.D < 8 || p.D > 256 || p.D % 8) return cudaErrorInvalidValue;
  if (p.dq_accum != nullptr && (p.D != dk || dk == 256)) return cudaErrorInvalidValue;  // the accumulator layout needs D in {64, 128}
  const dim3 block(256);
  if (dk == 256) return bf16 ? launch_pdl(fa2_delta_kernel<256, true>, grid, block, 0, stream, p) : launch_pdl(fa2_delta_kernel<256, false>, grid, block, 0, stream, p);
  if (dk == 128) return bf16 ? launch_pdl(fa2_delta_kernel<128, true>, grid, block, 0, stream, p) : launch_pdl(fa2_delta_kernel<128, false>, grid, block, 0, stream, p);
  return bf16 ? launch_pdl(fa2_delta_kernel<64, true>, grid, block, 0, stream, p) : launch_pdl(fa2_delta_kernel<64, false>, grid, block, 0, stream, p);
}

// =================================================================================================
// fused dK / dV / dQ
// =================================================================================================
// dQ accumulator layout (opaque workspace, undone by fa2_dq_convert_kernel):
//   [B][Hq][Sq_pad/128 tiles][D/32 chunks][128 rows][32 fp32], and inside a 128-byte row the 16-byte piece index
//   is XOR-ed with (row & 7) so that the staging writes (thread = row) are bank-conflict free. One chunk is one
//   16 KB cp.reduce.async.bulk (fp32 add performed by the L2 atomic units).
// The reduction is asynchronous and paced by L2 (measured: ~22 B/clk/SM chip-wide, scripts/micro/red_bw.cu), so it
// gets its own thread (warp 14) and a 2-chunk staging ring: the drain warps only ever wait for a free chunk, the
// compute and MMA warps never wait for the reduction at all. (Adding from registers with REDG.128 instead keeps the
// LSU busy for the whole L2 round trip and stalls the compute warps' shared-memory traffic behind it.)
//
// Tensor-pipe order per query block i (all five products are 128 x 128 x 128 at D = 128):
//     dV(i) | S(i+1) | dQ(i) | dK(i) | dP(i+1)
// dP(i+1) does not depend on P(i+1), so it is issued one block ahead: when the compute warps finish P(i+1) its dP
// is already in tensor memory and dS follows immediately. dQ(i) goes before dK(i) so that the drain warps empty
// the shared dP/dQ columns while dK(i) runs.
template <int D, bool kBf16, bool kFusedDq>
__global__ void __launch_bounds__(kBwdThreads, 1)
fa2_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
               const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
               const __grid_constant__ CUtensorMap tmap_dk, const __grid_constant__ CUtensorMap tmap_dv,
               const BwdParams p) {
  using L = BwdSmem<D>;
  constexpr int kSub = D / 64;
  constexpr int kSubBytes = kBlk * 128;  // [128 rows x 64 x 16-bit]
  constexpr int kQStages = L::kQStages;
  constexpr uint32_t kTmemS = 0, kTmemDV = 128, kTmemDK = 256, kTmemDP = 384;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem + L::kKOff;
  uint8_t* sV = smem + L::kVOff;
  uint8_t* sQ = smem + L::kQOff;
  uint8_t* sDO = smem + L::kDOOff;
  uint8_t* sDS = smem + L::kDsOff;
  float* sLse = reinterpret_cast<float*>(smem + L::kLseOff);
  float* sDelta = reinterpret_cast<float*>(smem + L::kDeltaOff);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlotOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int n_kv_blocks = (p.Sk + kBlk - 1) / kBlk;
  const int n_q_blocks = (p.Sq + kBlk - 1) / kBlk;
  // Ticket-ordered mode: the CTA's place in the grid's linear order is drawn from a counter when it starts (see
  // BwdParams::order_ctr) instead of being read off blockIdx - a CTA then only ever waits for CTAs that have started.
  int lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  if (p.order_ctr != nullptr) {
    if (threadIdx.x == 0) tmem_slot[1] = static_cast<uint32_t>(atomicAdd(p.order_ctr, 1));
    __syncthreads();
    lin = static_cast<int>(tmem_slot[1]);
  }
  const int cta_x = lin % static_cast<int>(gridDim.x), cta_yz = lin / static_cast<int>(gridDim.x);
  const int nblk = n_kv_blocks - 1 - cta_x;  // high key blocks are started first
  const int head_kv = cta_yz % static_cast<int>(gridDim.y);
  const int batch = cta_yz / static_cast<int>(gridDim.y);
  const int group = p.Hq / p.Hkv;
  const int shift = p.Sk - p.Sq;
  const int kv0 = nblk * kBlk;

  // first query block that sees this key block (bottom-right aligned causal mask)
  int m_start = 0;
  if (p.causal) m_start = max(0, (kv0 - shift) / kBlk);
  const int m_count = max(0, n_q_blocks - m_start);
  const int T = m_count * group;  // iterations: (head in group) x (query block)
  // Non-causal: key block n starts at query tile n and wraps, so concurrently running CTAs touch different dQ
  // tiles (and, in deterministic mode, reach each tile in ticket order without a start-up stagger).
  const bool rotate = !p.causal && (!p.deterministic || (n_kv_blocks <= n_q_blocks && n_kv_blocks <= 128));
  auto q_block_of = [&](int it) {
    const int j = it % m_count;
    return rotate ? (nblk + j) % n_q_blocks : m_start + j;
  };

  if ((smem_u32(smem) & 1023u) != 0) __trap();

  if (warp == 13 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_dk);
    tma_prefetch_desc(&tmap_dv);
  }
  if (warp == 12) {
    if (lane == 0) {
      mbar_init(&bars[kKvFull], 1);
      for (int s = 0; s < kQStages; ++s) {
        mbar_init(&bars[kQFull + s], 1);
        mbar_init(&bars[kQEmpty + s], 1);
      }
      mbar_init(&bars[kDoFull], 1);
      mbar_init(&bars[kDoEmpty], 1);
      mbar_init(&bars[kSFull], 1);
      mbar_init(&bars[kPFull], 256);
      mbar_init(&bars[kDpFull], 1);
      mbar_init(&bars[kDsFull], 256);
      mbar_init(&bars[kDqFull], 1);
      mbar_init(&bars[kDqEmpty], 128);
      mbar_init(&bars[kDkvFull], 1);
      mbar_init(&bars[kDeltaFull + 0], 1);
      mbar_init(&bars[kDeltaFull + 1], 1);
      for (int c = 0; c < 2; ++c) {
        mbar_init(&bars[kStageFull + c], 128);
        mbar_init(&bars[kStageEmpty + c], 1);
      }
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
  constexpr uint32_t tmem_base = 0u;

  // setmaxnreg sits at the top of each role's branch: ptxas only widens the register budget for code the
  // instruction dominates.
  if (warp >= 12) {
    reg_dealloc<kRegsIssue>();
  if (warp == 13) {
    // ========================================= TMA producer ========================================
    if (T > 0 && elect_one()) {
      mbar_arrive_expect_tx(&bars[kKvFull], 2 * L::kTileBytes);
      for (int j = 0; j < kSub; ++j) {
        tma_load_4d(sK + j * kSubBytes, &tmap_k, &bars[kKvFull], j * 64, kv0, head_kv, batch);
        tma_load_4d(sV + j * kSubBytes, &tmap_v, &bars[kKvFull], j * 64, kv0, head_kv, batch);
      }
      int qs = 0;
      uint32_t qph = 0;
      for (int it = 0; it < T; ++it) {
        const int head = head_kv * group + it / m_count;
        const int m = q_block_of(it);
        const int64_t row_off = (static_cast<int64_t>(batch) * p.Hq + head) * p.lse_stride + m * kBlk;
        // Q(it) + its LSE and delta rows: ring of kQStages (released by dK(it - kQStages))
        mbar_wait(&bars[kQEmpty + qs], qph ^ 1);
        mbar_arrive_expect_tx(&bars[kQFull + qs], L::kTileBytes + kBlk * 4);
        for (int j = 0; j < kSub; ++j)
          tma_load_4d(sQ + qs * L::kTileBytes + j * kSubBytes, &tmap_q, &bars[kQFull + qs], j * 64, m * kBlk, head, batch);
        bulk_load_1d(sLse + qs * kBlk, p.lse + row_off, kBlk * 4, &bars[kQFull + qs]);
        // dO(it): single buffer, free again once dV(it - 1) has been computed. Its delta row goes to slot it & 1
        // (still read by the compute warps after dO itself is dead), which dV(it - 1) has released as well.
        mbar_wait(&bars[kDoEmpty], (it & 1) ^ 1);
        mbar_arrive_expect_tx(&bars[kDeltaFull + (it & 1)], kBlk * 4);
        bulk_load_1d(sDelta + (it & 1) * kBlk, p.delta + row_off, kBlk * 4, &bars[kDeltaFull + (it & 1)]);
        mbar_arrive_expect_tx(&bars[kDoFull], L::kTileBytes);
        for (int j = 0; j < kSub; ++j)
          tma_load_4d(sDO + j * kSubBytes, &tmap_do, &bars[kDoFull], j * 64, m * kBlk, head, batch);
        if (++qs == kQStages) { qs = 0; qph ^= 1; }
      }
    }
  } else if (warp == 12) {
    // ========================================== MMA issuer =========================================
    if (T > 0 && elect_one()) {
      constexpr uint32_t idesc_kk = make_idesc_f16(kBlk, kBlk, kBf16, false, false);  // S^T, dP^T
      constexpr uint32_t idesc_tn = make_idesc_f16(kBlk, D, kBf16, false, true);      // dV (TS), dK
      constexpr uint32_t idesc_nn = make_idesc_f16(kBlk, D, kBf16, true, true);       // dQ
      // K-major views (rows of the tile are the M/N index, reduction runs along the 128-byte rows)
      const SmemDesc k_kmaj = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const SmemDesc v_kmaj = make_smem_desc_sw128(smem_u32(sV), 16, 1024);
      const SmemDesc q_kmaj = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const SmemDesc do_kmaj = make_smem_desc_sw128(smem_u32(sDO), 16, 1024);
      const SmemDesc ds_kmaj = make_smem_desc_sw128(smem_u32(sDS), 16, 1024);
      // MN-major views (rows of the tile are the reduction index)
      const SmemDesc k_mn = make_smem_desc_sw128(smem_u32(sK), kSubBytes, 1024);
      const SmemDesc q_mn = make_smem_desc_sw128(smem_u32(sQ), kSubBytes, 1024);
      const SmemDesc do_mn = make_smem_desc_sw128(smem_u32(sDO), kSubBytes, 1024);
      const SmemDesc ds_mn = make_smem_desc_sw128(smem_u32(sDS), kSubBytes, 1024);

      // [128 x D] x [128 x D]^T, both K-major, reduction over D
      auto issue_kmajor = [&](uint32_t d_tmem, SmemDesc a, SmemDesc b) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = (k / 4) * kSubBytes + (k % 4) * 32;
          umma_ss(d_tmem, advance(a, off), advance(b, off), idesc_kk, k > 0 ? 1u : 0u);
        }
      };

      mbar_wait(&bars[kKvFull], 0);
      mbar_wait(&bars[kQFull + 0], 0);
      tcgen05_fence_after();
      issue_kmajor(tmem_base + kTmemS, k_kmaj, q_kmaj);  // S^T(0) = K Q(0)^T
      umma_commit(&bars[kSFull]);
      mbar_wait(&bars[kDoFull], 0);
      tcgen05_fence_after();
      issue_kmajor(tmem_base + kTmemDP, v_kmaj, do_kmaj);  // dP^T(0) = V dO(0)^T
      umma_commit(&bars[kDpFull]);

      int qs = 0, qs_next = 1 % kQStages;
      uint32_t qph_next = (kQStages == 1) ? 1u : 0u;
      for (int it = 0; it < T; ++it) {
        const uint32_t par = it & 1;
        const uint32_t q_off = qs * L::kTileBytes;

        // dV += P^T dO   (A = P^T from TMEM: k-steps 0-3 in columns [0,32), 4-7 in [64,96))
        mbar_wait(&bars[kPFull], par);
        tcgen05_fence_after();
        FA2_STAMP(2, it, 0);
#pragma unroll
        for (int k = 0; k < kBlk / 16; ++k) {
          const uint32_t a_col = (k < 4) ? k * 8 : 64 + (k - 4) * 8;
          umma_ts(tmem_base + kTmemDV, tmem_base + kTmemS + a_col, advance(do_mn, k * 2048), idesc_tn,
                  (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&bars[kDoEmpty]);  // dO(it) has now fed both dP(it) and dV(it)
        FA2_STAMP(2, it, 1);
        // S^T(next) = K Q(next)^T
        if (it + 1 < T) {
          mbar_wait(&bars[kQFull + qs_next], qph_next);
          tcgen05_fence_after();
          issue_kmajor(tmem_base + kTmemS, k_kmaj, advance(q_kmaj, qs_next * L::kTileBytes));
          umma_commit(&bars[kSFull]);
        }
        FA2_STAMP(2, it, 2);
        // dQ = dS K, then dK += dS^T Q
        mbar_wait(&bars[kDsFull], par);
        tcgen05_fence_after();
        FA2_STAMP(2, it, 3);
        if constexpr (kFusedDq) {
#pragma unroll
          for (int k = 0; k < kBlk / 16; ++k) {
            umma_ss(tmem_base + kTmemDP, advance(ds_mn, k * 2048), advance(k_mn, k * 2048), idesc_nn, k > 0 ? 1u : 0u);
          }
          umma_commit(&bars[kDqFull]);
        }
#pragma unroll
        for (int k = 0; k < kBlk / 16; ++k) {
          const uint32_t off = (k / 4) * kSubBytes + (k % 4) * 32;
          umma_ss(tmem_base + kTmemDK, advance(ds_kmaj, off), advance(q_mn, q_off + k * 2048), idesc_tn,
                  (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&bars[kQEmpty + qs]);
        FA2_STAMP(2, it, 4);
        // dP^T(next) = V dO(next)^T  (needs dO(next) and the dP/dQ columns back from the drain warps)
        if (it + 1 < T) {
          mbar_wait(&bars[kDoFull], (it + 1) & 1);
          if constexpr (kFusedDq) mbar_wait(&bars[kDqEmpty], par);
          tcgen05_fence_after();
          issue_kmajor(tmem_base + kTmemDP, v_kmaj, do_kmaj);
          umma_commit(&bars[kDpFull]);
        }
        FA2_STAMP(2, it, 5);
        qs = qs_next;
        if (++qs_next == kQStages) { qs_next = 0; qph_next ^= 1; }
      }
      umma_commit(&bars[kDkvFull]);
    }
  } else if (warp == 14) {
    // ========================================= dQ reducer ==========================================
    // One thread turns staged chunks into cp.reduce.async.bulk fp32 adds. Deterministic mode: the key blocks that
    // see one query tile add in a fixed order, serialised by a ticket per tile.
    if (kFusedDq && T > 0 && elect_one()) {
      constexpr int kChunks = D / 32;
      uint32_t n = 0;         // chunks issued so far (slot = n & 1, phase = (n >> 1) & 1)
      uint32_t released = 0;  // chunks whose staging slot has been handed back to the drain warps
      for (int it = 0; it < T; ++it) {
        const int head = head_kv * group + it / m_count;
        const int m = q_block_of(it);
        float* gtile = p.dq_accum + ((static_cast<int64_t>(batch) * p.Hq + head) * (p.sq_pad / kBlk) + m) *
                                        static_cast<int64_t>(kBlk) * D;
        int* ticket = nullptr;
        int my_turn = 0;
        if (p.deterministic) {
          ticket = p.dq_sem + (static_cast<int64_t>(batch) * p.Hq + head) * (p.sq_pad / kBlk) + m;
          if (p.causal) {
            // key blocks visible to query tile m: 0 .. n_last; they add in descending order
            const int last_row = min(m * kBlk + kBlk, p.Sq) - 1;
            const int n_last = min(n_kv_blocks - 1, (last_row + shift) / kBlk);
            my_turn = n_last - nblk;
          } else if (rotate) {
            // rotated schedule: key block n reaches query tile m at step (m - n) mod n_q_blocks
            const int top = min(m, n_kv_blocks - 1);
            my_turn = (nblk <= m) ? (top - nblk) : (top + 1 + (n_kv_blocks - 1 - nblk));
          } else {
            my_turn = n_kv_blocks - 1 - nblk;
          }
        }
        for (int c = 0; c < kChunks; ++c) {
          const uint32_t slot = n & 1;
          mbar_wait(&bars[kStageFull + slot], (n >> 1) & 1);
          if (p.deterministic && c == 0) {
            const long long start = clock64();
            while (ld_acquire(ticket) != my_turn) {
              if (clock64() - start > FA2_WATCHDOG_CYCLES) __trap();
            }
          }
          bulk_reduce_add_f32(gtile + c * (kBlk * 32), smem + L::kStageOff + slot * L::kChunkBytes, L::kChunkBytes);
          bulk_commit_group();
          ++n;
          bulk_wait_group_read<1>();  // every chunk but the one just issued has been read out of its slot
          for (; released + 1 < n; ++released) mbar_arrive(&bars[kStageEmpty + (released & 1)]);
        }
        if (p.deterministic) {
          bulk_wait_group<0>();  // every add of this tile is performed ...
          for (; released < n; ++released) mbar_arrive(&bars[kStageEmpty + (released & 1)]);
          __threadfence();
          st_release(ticket, my_turn + 1);  // ... before the next key block may start its own
        }
        FA2_STAMP(3, it, 3);
      }
      bulk_wait_group<0>();
    }
  }
  } else if (warp < 8) {
    // ====================================== compute warpgroups =====================================
    reg_alloc<kRegsMath>();
    const int wg = warp >> 2;
    const int r = threadIdx.x & 127;  // key row inside the block == TMEM lane
    const int kv = kv0 + r;
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_sel + kTmemS + wg * 64;
    const uint32_t tP = tmem_base + lane_sel + kTmemS + wg * 64;  // P^T (16-bit) over this warpgroup's own S^T columns
    const uint32_t tDP = tmem_base + lane_sel + kTmemDP + wg * 64;
    const float2 neg_c2 = make_float2(-p.scale_log2, -p.scale_log2);
    uint8_t* ds_sub = sDS + wg * kSubBytes;

    int qs = 0;
    uint32_t qph = 0;
    for (int it = 0; it < T; ++it) {
      const uint32_t par = it & 1;
      const int m = q_block_of(it);
      const int q0 = m * kBlk + wg * 64;  // first query column of this thread's slice
      const float4* lse_s = reinterpret_cast<const float4*>(sLse + qs * kBlk + wg * 64);
      const float4* delta_s = reinterpret_cast<const float4*>(sDelta + (it & 1) * kBlk + wg * 64);
      const bool need_mask = (kv0 + kBlk > p.Sk) || (m * kBlk + kBlk > p.Sq) || (p.causal && (kv0 + kBlk - 1 > m * kBlk + shift));

      // ---- P^T = exp2(S^T * c - LSE) = exp2(-(LSE - S^T * c)) ----
      if (r == 0) FA2_STAMP(wg, it, 0);
      mbar_wait(&bars[kQFull + qs], qph);  // the LSE row travels with Q
      mbar_wait(&bars[kSFull], par);
      tcgen05_fence_after();
      if (r == 0) FA2_STAMP(wg, it, 1);
      uint32_t su[2][32];
      tmem_ld_32x32b_x32(tS, su[0]);
      tmem_ld_32x32b_x32(tS + 32, su[1]);
      tmem_wait_ld();
      float2 pf[32];  // probabilities, fp32, pairs of neighbouring query columns
#pragma unroll
      for (int j4 = 0; j4 < 16; ++j4) {
        const float4 l4 = lse_s[j4];
        const int j = j4 * 4;
        const float2 t0 = ffma2(make_float2(__uint_as_float(su[j / 32][j % 32]), __uint_as_float(su[j / 32][j % 32 + 1])),
                                neg_c2, make_float2(l4.x, l4.y));
        const float2 t1 = ffma2(make_float2(__uint_as_float(su[j / 32][j % 32 + 2]), __uint_as_float(su[j / 32][j % 32 + 3])),
                                neg_c2, make_float2(l4.z, l4.w));
        // a fully masked query row carries LSE = -inf: exp2(+inf) = inf here, replaced by 0 in the masked path below
        // (such rows only occur in blocks that need masking)
        // one pair on the MUFU, one on the FMA pipes (ex2_emul2_neg): halves the MUFU time of this phase
        pf[j4 * 2] = make_float2(ex2_approx(-t0.x), ex2_approx(-t0.y));
        pf[j4 * 2 + 1] = ex2_emul2_neg(t1);
      }
      if (need_mask) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int qi = q0 + 2 * j;
          const bool row_ok = kv < p.Sk;
          const bool ok0 = row_ok && (qi < p.Sq) && (!p.causal || kv <= qi + shift);
          const bool ok1 = row_ok && (qi + 1 < p.Sq) && (!p.causal || kv <= qi + 1 + shift);
          pf[j].x = ok0 ? pf[j].x : 0.f;
          pf[j].y = ok1 ? pf[j].y : 0.f;
        }
      }
      {
        uint32_t pk[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) pk[j] = pack2<kBf16>(pf[j].x, pf[j].y);
        tmem_st_32x32b_x32(tP, pk);
        tmem_wait_st();
      }
      tcgen05_fence_before();
      mbar_arrive(&bars[kPFull]);
      if (r == 0) FA2_STAMP(wg, it, 2);

      // ---- dS^T = P^T o (dP^T - delta) ----
      mbar_wait(&bars[kDeltaFull + (it & 1)], (it >> 1) & 1);
      mbar_wait(&bars[kDpFull], par);
      tcgen05_fence_after();
      if (r == 0) FA2_STAMP(wg, it, 3);
      uint32_t dpu[2][32];
      tmem_ld_32x32b_x32(tDP, dpu[0]);
      tmem_wait_ld();
      tmem_ld_32x32b_x32(tDP + 32, dpu[1]);  // in flight while the first half is processed
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (half == 1) tmem_wait_ld();
#pragma unroll
        for (int c16 = 0; c16 < 4; ++c16) {
          uint32_t w[4];
#pragma unroll
          for (int e4 = 0; e4 < 2; ++e4) {
            const float4 d4 = delta_s[half * 8 + c16 * 2 + e4];
            const int j = c16 * 8 + e4 * 4;
            const float2 x0 = fsub2(make_float2(__uint_as_float(dpu[half][j]), __uint_as_float(dpu[half][j + 1])),
                                    make_float2(d4.x, d4.y));
            const float2 x1 = fsub2(make_float2(__uint_as_float(dpu[half][j + 2]), __uint_as_float(dpu[half][j + 3])),
                                    make_float2(d4.z, d4.w));
            const float2 s0 = fmul2(pf[half * 16 + c16 * 4 + e4 * 2], x0);
            const float2 s1 = fmul2(pf[half * 16 + c16 * 4 + e4 * 2 + 1], x1);
            w[e4 * 2] = pack2<kBf16>(s0.x, s0.y);
            w[e4 * 2 + 1] = pack2<kBf16>(s1.x, s1.y);
          }
          *reinterpret_cast<uint4*>(ds_sub + sw128_offset(r, half * 4 + c16)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      mbar_arrive(&bars[kDsFull]);
      if (r == 0) FA2_STAMP(wg, it, 4);
      if (++qs == kQStages) { qs = 0; qph ^= 1; }
    }

    // ---- epilogue: dK * scale, dV -> smem (the K / V tiles are dead once every MMA retired) -> TMA store ----
    if (T > 0) {
      mbar_wait(&bars[kDkvFull], 0);
      tcgen05_fence_after();
    }
    constexpr int kColsPerWg = D / 2;
#pragma unroll
    for (int which = 0; which < 2; ++which) {  // 0: dK, 1: dV
      uint8_t* sOut = (which == 0) ? sK : sV;
      const uint32_t tsrc = tmem_base + lane_sel + (which == 0 ? kTmemDK : kTmemDV) + wg * kColsPerWg;
      const float mul = (which == 0) ? p.softmax_scale : 1.0f;
#pragma unroll
      for (int cidx = 0; cidx < kColsPerWg / 32; ++cidx) {
        uint32_t o[32];
        if (T > 0) {
          tmem_ld_32x32b_x32(tsrc + cidx * 32, o);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = 0u;
        }
        const int col = wg * kColsPerWg + cidx * 32;  // first head-dim column of this chunk
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
    named_bar_sync(1, 256);
    if (threadIdx.x == 0) {
      for (int j = 0; j < kSub; ++j) {
        tma_store_4d(&tmap_dk, sK + j * kSubBytes, j * 64, kv0, head_kv, batch);
        tma_store_4d(&tmap_dv, sV + j * kSubBytes, j * 64, kv0, head_kv, batch);
      }
      bulk_commit_group();
      bulk_wait_group_read<0>();
    }
  } else {
    // ========================================== dQ drain ===========================================
    // TMEM -> registers (frees the dP/dQ columns at once) -> 16 KB staging chunks -> the reducer thread.
    reg_alloc<kRegsMath>();
    const int r = threadIdx.x & 127;  // query row inside the block == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tDQ = tmem_base + lane_sel + kTmemDP;
    constexpr int kChunks = D / 32;
    uint32_t n = 0;  // running chunk counter: slot = n & 1

    for (int it = 0; kFusedDq && it < T; ++it) {
      const uint32_t par = it & 1;
      mbar_wait(&bars[kDqFull], par);
      tcgen05_fence_after();
      if (r == 0) FA2_STAMP(3, it, 0);
      uint32_t acc[kChunks][32];
#pragma unroll
      for (int c = 0; c < kChunks; ++c) tmem_ld_32x32b_x32(tDQ + c * 32, acc[c]);
      tmem_wait_ld();
      tcgen05_fence_before();
      mbar_arrive(&bars[kDqEmpty]);  // dP^T(next) may overwrite these columns now
      if (r == 0) FA2_STAMP(3, it, 1);
#pragma unroll
      for (int c = 0; c < kChunks; ++c, ++n) {
        const uint32_t slot = n & 1;
        mbar_wait(&bars[kStageEmpty + slot], ((n >> 1) & 1) ^ 1);
        uint8_t* row = smem + L::kStageOff + slot * L::kChunkBytes + r * 128;
#pragma unroll
        for (int pc = 0; pc < 8; ++pc) {
          *reinterpret_cast<uint4*>(row + ((pc ^ (r & 7)) << 4)) =
              make_uint4(acc[c][pc * 4], acc[c][pc * 4 + 1], acc[c][pc * 4 + 2], acc[c][pc * 4 + 3]);
        }
        fence_proxy_async_smem();
        mbar_arrive(&bars[kStageFull + slot]);
      }
      if (r == 0) FA2_STAMP(3, it, 2);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 12) {
    tcgen05_fence_after();
    tmem_dealloc<512>(0u);
  }
}

template <int D, bool kBf16, bool kFusedDq>
static cudaError_t launch_bwd_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                                const CUtensorMap& tdo, const CUtensorMap& tdk, const CUtensorMap& tdv,
                                const BwdParams& p, cudaStream_t stream) {
  auto kernel = fa2_bwd_kernel<D, kBf16, kFusedDq>;
  constexpr int smem = BwdSmem<D>::kBytes;
  if (cudaError_t e = ensure_dynamic_smem<fa2_bwd_kernel<D, kBf16, kFusedDq>>(smem); e != cudaSuccess) return e;
  dim3 grid((p.Sk + kBlk - 1) / kBlk, p.Hkv, p.B);
  kernel<<<grid, kBwdThreads, smem, stream>>>(tq, tk, tv, tdo, tdk, tdv, p);
  return cudaGetLastError();
}

cudaError_t launch_bwd(int D, bool bf16, bool fused_dq, const CUtensorMap& tq, const CUtensorMap& tk,
                       const CUtensorMap& tv, const CUtensorMap& tdo, const CUtensorMap& tdk, const CUtensorMap& tdv,
                       const BwdParams& p, cudaStream_t stream) {
#define FA2_BWD_CASE(DD, BF, FU) \
  if (D == DD && bf16 == BF && fused_dq == FU) return launch_bwd_t<DD, BF, FU>(tq, tk, tv, tdo, tdk, tdv, p, stream);
  // only the fused instantiation is built: dK / dV alone come from bwd_dkdv_sm100.cu
  FA2_BWD_CASE(128, true, true) FA2_BWD_CASE(128, false, true) FA2_BWD_CASE(64, true, true) FA2_BWD_CASE(64, false, true)
#undef FA2_BWD_CASE
  return cudaErrorInvalidValue;
}

// =================================================================================================
// dQ accumulator -> dQ
// =================================================================================================
// One CTA per 64 accumulator rows: reads the chunked / swizzled layout with coalesced float4 loads (128 contiguous
// bytes per row and chunk), un-swizzles through shared memory and writes rows of the caller's [B, Sq, Hq, D]
// tensor in 16-byte pieces.
template <int D>
__global__ void __launch_bounds__(256) fa2_dq_convert_kernel(const DqConvertParams p) {
  constexpr int kRows = 64;
  constexpr int kPitch = D + 4;  // floats; keeps both the transposing writes and the row reads conflict-free
  __shared__ __align__(16) float tile[kRows * kPitch];
  const int half = blockIdx.x & 1, tile_idx = blockIdx.x >> 1;
  const int head = blockIdx.y, batch = blockIdx.z;
  const int row0 = tile_idx * 128 + half * kRows;
  if (row0 >= p.Sq) return;
  const float* src = p.dq_accum + ((static_cast<int64_t>(batch) * p.Hq + head) * (p.sq_pad / 128) + tile_idx) * 128LL * D;
  // accumulator: [chunk (32 columns)][row][8 pieces of 16 B], piece p of row r holds logical piece p ^ (r & 7)
  for (int idx = threadIdx.x; idx < (D / 32) * kRows * 8; idx += 256) {
    const int pc = idx & 7, r = (idx >> 3) & (kRows - 1), c = idx / (8 * kRows);
    const int row_in_tile = half * kRows + r;
    const float4 v = *reinterpret_cast<const float4*>(src + ((c * 128 + row_in_tile) * 8 + pc) * 4);
    const int col = c * 32 + ((pc ^ (row_in_tile & 7)) << 2);
    *reinterpret_cast<float4*>(&tile[r * kPitch + col]) = make_float4(v.x * p.scale, v.y * p.scale, v.z * p.scale, v.w * p.scale);
  }
  __syncthreads();
  const int esize = (p.out_dtype == 32) ? 4 : 2;
  const int per16 = 16 / esize;                 // elements per 16-byte piece
  const int pieces_per_row = D / per16;
  uint8_t* out = reinterpret_cast<uint8_t*>(p.dq);
  for (int idx = threadIdx.x; idx < kRows * pieces_per_row; idx += 256) {
    const int rr = idx / pieces_per_row, pc = idx % pieces_per_row;
    const int row = row0 + rr;
    if (row >= p.Sq) continue;
    const float* t = &tile[rr * kPitch + pc * per16];
    const int64_t dst = (batch * p.dq_strides[0] + head * p.dq_strides[1] + static_cast<int64_t>(row) * p.dq_strides[2] + pc * per16) * esize;
    if (p.out_dtype == 32) {
      *reinterpret_cast<float4*>(out + dst) = *reinterpret_cast<const float4*>(t);
    } else {
      const float4 a = *reinterpret_cast<const float4*>(t);
      const float4 b = *reinterpret_cast<const float4*>(t + 4);
      uint4 o;
      if (p.out_dtype == 17) {
        o = make_uint4(pack2<true>(a.x, a.y), pack2<true>(a.z, a.w), pack2<true>(b.x, b.y), pack2<true>(b.z, b.w));
      } else {
        o = make_uint4(pack2<false>(a.x, a.y), pack2<false>(a.z, a.w), pack2<false>(b.x, b.y), pack2<false>(b.z, b.w));
      }
      *reinterpret_cast<uint4*>(out + dst) = o;
    }
  }
}

cudaError_t launch_dq_convert(const DqConvertParams& p, cudaStream_t stream) {
  dim3 grid(2 * (p.sq_pad / 128), p.Hq, p.B);
  if (p.D == 128) fa2_dq_convert_kernel<128><<<grid, 256, 0, stream>>>(p);
  else if (p.D == 64) fa2_dq_convert_kernel<64><<<grid, 256, 0, stream>>>(p);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace fa2
