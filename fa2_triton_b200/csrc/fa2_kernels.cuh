// Internal interface between the C-ABI translation unit (api.cu) and the kernel translation units.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <utility>
#include <vector>

#include "extras.cuh"

namespace fa2 {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device); safe to call from several host threads
template <auto Kernel>
inline cudaError_t ensure_dynamic_smem(int bytes) {
  static std::mutex mu;
  static std::vector<char> done;  // indexed by device ordinal
  int dev = 0;
  if (cudaError_t e = cudaGetDevice(&dev); e != cudaSuccess) return e;
  std::lock_guard<std::mutex> guard(mu);
  if (dev < static_cast<int>(done.size()) && done[dev]) return cudaSuccess;
  if (cudaError_t e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); e != cudaSuccess) return e;
  if (static_cast<int>(done.size()) <= dev) done.resize(dev + 1, 0);
  done[dev] = 1;
  return cudaSuccess;
}

// Launch with programmatic stream serialization (PDL): the kernel may be scheduled before the previous kernel of the
// stream has finished, and orders itself behind it with griddepcontrol.wait (ptx.cuh: pdl_wait / pdl_launch_dependents).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// multiprocessor count of the current device (persistent kernels launch one CTA per SM)
inline int current_sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
    return 148;
  return sms;
}

// ---------------------------------------------------------------------------------------------
// Order in which a causal grid is worked through
// ---------------------------------------------------------------------------------------------
// The grids are (x = block of the sequence, heaviest first; y = head; z = batch) and the hardware hands out CTAs in
// linear order: head by head, which keeps a head's K / V (forward, dQ) or Q / dO (dK/dV) in L2 while its blocks run.
// Under a causal mask the blocks of a head differ in length by up to 32x, and with ~28 uneven CTAs per SM the last
// wave leaves SMs idle: list-scheduling the real item lengths on 148 SMs gives 92 % (forward), 96 % (dK/dV, dQ) of the
// ideal makespan at config 3. So the LAST heads are dealt block-major instead - the heaviest blocks of all of them
// first, their lightest last (98-99 %) - as many heads as keep their streamed operands within ~48 MB of L2. Which item
// a CTA works on changes, not what it computes: results are bit-identical.
//
// Two forms. The forward (a grid kernel): the tail is the last 2^k heads of the LAST batch element, so that the new
// coordinates follow from blockIdx with shifts and masks alone, in the uniform datapath. The dQ kernels (persistent; items
// decoded by a few threads): any number of (batch, head) pairs. NOT the dK/dV kernel: remapped coordinates - with a
// division per CTA or without - cost its compute warps registers they do not have (4 -> 48-56 bytes of spills, +16 %
// cycles under ncu; profiles/r02_experiments.txt, item 10), far more than the 1.8 % the better tail is worth there.
#ifndef FA2_TAIL_HEADS_MAX
#define FA2_TAIL_HEADS_MAX 32
#endif
#ifndef FA2_TAIL_L2_MB
#define FA2_TAIL_L2_MB 48  // L2 budget for the streamed operands of the tail's heads
#endif
inline int tail_heads_for(size_t streamed_bytes_per_head, int causal) {
  if (!causal || streamed_bytes_per_head == 0) return 0;
  const size_t k = (static_cast<size_t>(FA2_TAIL_L2_MB) << 20) / streamed_bytes_per_head;
  const int t = static_cast<int>(k < FA2_TAIL_HEADS_MAX ? k : FA2_TAIL_HEADS_MAX);
  return t < 2 ? 0 : t;
}
// log2 of the largest power of two <= min(tail_heads, heads); 0 = no block-major tail
inline int tail_log2_for(int tail_heads, int heads) {
  const int t = tail_heads < heads ? tail_heads : heads;
  int l = 0;
  while ((2 << l) <= t) ++l;
  return l;
}
struct ItemCoord {
  int x, head, batch;
};
// grid kernels: block (x, y, z) of an (nx, ny, nz) grid -> the item it works on
__host__ __device__ inline ItemCoord tail_balanced_block(int x, int y, int z, int nx, int ny, int nz, int tail_log2) {
  // (branch-free, on purpose: with `if (a && b && c)` ptxas evaluated the condition per thread - predicates packed into
  // general registers, 86 more LOP3 - instead of once in the uniform datapath)
  const int h0 = ny - (1 << tail_log2);  // first head of the tail
  const int in_tail = -static_cast<int>((tail_log2 > 0) & (z == nz - 1) & (y >= h0));  // all ones or zero
  const int r = (y - h0) * nx + x;       // place inside the tail in launch order
  const int xt = r >> tail_log2, ht = h0 + (r & ((1 << tail_log2) - 1));
  return ItemCoord{x ^ ((x ^ xt) & in_tail), y ^ ((y ^ ht) & in_tail), z};
}
// persistent kernels: linear item index -> (x, head, batch) for nx blocks per head, ny heads, nbh = ny * batches pairs
__host__ __device__ inline ItemCoord tail_balanced_item(int lin, int nx, int ny, int nbh, int tail_heads) {
  const int K = tail_heads < nbh ? tail_heads : nbh;
  const int body = (nbh - K) * nx;
  int x, bh;
  if (K < 2 || lin < body) {
    bh = lin / nx;
    x = lin - bh * nx;
  } else {
    const int r = lin - body;
    x = r / K;
    bh = nbh - K + (r - x * K);
  }
  return ItemCoord{x, bh % ny, bh / ny};
}
#ifdef __CUDACC__
__device__ __forceinline__ ItemCoord tail_balanced_cta(int tail_log2) {
  return tail_balanced_block(blockIdx.x, blockIdx.y, blockIdx.z, gridDim.x, gridDim.y, gridDim.z, tail_log2);
}
#endif

// ---------------------------------------------------------------------------------------------
// forward (fwd_sm100.cu)
// ---------------------------------------------------------------------------------------------
// Variable-length batches (right-padded [B, S, ...] tensors, one length per batch element; query and key lengths
// coincide like in the reference's attention_mask mode, src/forward/caller.py:25-28): `seqlens` (device, [B]) gives the
// lengths and `vmaps` (device) per-batch tensor maps bounded by them - [B][kVarlenMapsPerBatch], order q, k, v, o for the
// forward and q, k, v, dO, dK, dV for the backward - so that TMA zero-fills / clips at each sequence's own end and
// nothing stored in the padding can reach a result. Both nullptr outside that mode.
constexpr int kVarlenMapsPerBatch = 8;

// setup kernel (varlen_sm100.cu): templates[i] is the launch's full map of tensor i, out[b][i] its per-sequence copy
struct VarlenSetupParams {
  const int* seqlens;            // device, [B]
  CUtensorMap* out;              // device, [B][kVarlenMapsPerBatch], 128-byte aligned
  uint64_t base[6];              // device address of tensor i
  uint64_t batch_stride_bytes[6];
  int n_maps;                    // <= 6
  int S;                         // padded sequence length: lengths are clamped to [0, S]
};
cudaError_t launch_varlen_maps(const CUtensorMap* templates, const VarlenSetupParams& p, int B, cudaStream_t stream);

struct FwdParams {
  float* lse;        // [B, Hq, lse_stride] fp32, log2 domain
  int B, Hq, Hkv, Sq, Sk;
  int lse_stride;    // ceil(Sq / 128) * 128
  float scale_log2;  // softmax_scale * log2(e)
  int causal;
  int n_pairs;       // ceil(Sq / 256): one CTA per pair of 128-row query tiles
  int tail_log2;     // log2 of the heads at the end of the grid that are dealt block-major (tail_balanced_cta); set by the launcher
  // extras (kernel variant kExt; all zero otherwise)
  BiasView bias;
  float bias_mul;       // 1 / softmax_scale: the bias joins the raw scores, which are multiplied by scale_log2 later
  // a 16-bit bias TMA can address: staged through shared memory with the map in bias_map ([B', Sq, H', Sk]; a broadcast
  // batch / head dimension has extent 1 there, so its coordinate is multiplied by 0)
  int bias_tma, bias_head_mul, bias_batch_mul;
  const CUtensorMap* bias_map;  // host pointer, read by the launcher only
  float dropout_p;      // 0: no dropout
  float inv_keep;       // 1 / (1 - dropout_p)
  uint32_t seed_lo, seed_hi;
  uint32_t drop_offset;  // added to every dropout counter (modulo 2^32); 0 for a whole [B, Hq, Sq, Sk] call
  const int* seqlens;         // variable-length mode (see above), else nullptr
  const CUtensorMap* vmaps;
  // head dims below the kernel's: the zero columns TMA fills in need no products. d_steps = ceil(d / 16) k-steps of
  // Q K^T, d_cols = d rounded up to 16 = the N of P V. Equal to D / 16 and D for a full head dim.
  int d_steps, d_cols;
};

cudaError_t launch_fwd(int D, bool bf16, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                       const CUtensorMap& to, const FwdParams& p, cudaStream_t stream);
// fwd2_sm100.cu: the plain forward (no bias, no dropout) with double-buffered 64-key half-blocks, head-dim instantiations
// 64 / 128 / 256 (256 also with a bias). The K / V tensor maps must have box rows 128 (D <= 128) or 64 (D = 256).
cudaError_t launch_fwd2(int D, bool bf16, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                        const CUtensorMap& to, const FwdParams& p, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// backward (bwd_sm100.cu)
// ---------------------------------------------------------------------------------------------
struct DeltaParams {
  const void* o;
  const void* dout;
  float* delta;  // [B, Hq, lse_stride]
  int B, Hq, Sq, D;       // D: the tensors' head dim (a multiple of 8, <= 128); the kernel is instantiated for 64 or 128
  int lse_stride;
  int64_t o_strides[3];   // batch, head, seq (elements)
  int64_t do_strides[3];
  float* dq_accum;  // optional: backward workspace to clear ([B, Hq, lse_stride, D] fp32), or nullptr
  int* dq_sem;      // optional: [B, Hq, lse_stride / 128] tickets (+ one CTA-order counter behind them) to clear, or nullptr
  const int* seqlens;  // variable-length mode: rows at or beyond a sequence's length get delta = 0 unread
};
cudaError_t launch_delta(bool bf16, const DeltaParams& p, cudaStream_t stream);

struct BwdParams {
  const float* lse;    // [B, Hq, lse_stride] log2 domain
  const float* delta;  // [B, Hq, lse_stride]
  float* dq_accum;     // [B, Hq, Sq_pad, D] fp32, tile-swizzled (see bwd_sm100.cu)
  int* dq_sem;         // [B, Hq, Sq_pad/128] ordered-reduction tickets (deterministic mode) or nullptr
  // deterministic mode: the word behind the tickets. Every CTA draws its place in the grid's linear order from it as it
  // STARTS, so a CTA that waits for a ticket only ever waits for CTAs that are already running or done - whatever order
  // the hardware starts thread blocks in. Zero on entry (cleared with the tickets by the delta kernel).
  int* order_ctr;
  int B, Hq, Hkv, Sq, Sk;
  int lse_stride;
  int sq_pad;  // ceil(Sq / 128) * 128
  float softmax_scale;
  float scale_log2;
  int causal;
  int deterministic;
};
// fused_dq = true : dK, dV and dQ partial sums (atomic mode; 5 products per block pair)
// fused_dq = false: dK, dV only (deterministic mode; dQ comes from launch_bwd_dq)
cudaError_t launch_bwd(int D, bool bf16, bool fused_dq, const CUtensorMap& tq, const CUtensorMap& tk,
                       const CUtensorMap& tv, const CUtensorMap& tdo, const CUtensorMap& tdk, const CUtensorMap& tdv,
                       const BwdParams& p, cudaStream_t stream);

// dQ kernel of the deterministic backward (bwd_dq_sm100.cu): one CTA per 128-row query block, fixed key-block order
struct DqParams {
  const float* lse;    // [B, Hq, lse_stride] log2 domain
  const float* delta;  // [B, Hq, lse_stride]
  int* sched;          // work-item counter of the persistent scheduler: one int, 0 on entry (the kernel re-zeroes it)
  void* dq;            // [B, Sq, Hq, D]
  int64_t dq_strides[3];  // batch, head, seq (elements)
  int dq_is_16bit;     // 1: dq has the input dtype (TMA store), 0: dq is fp32 (vector stores)
  int B, Hq, Hkv, Sq, Sk;
  int lse_stride;
  float softmax_scale;
  float scale_log2;
  int causal;
  BiasView bias;       // kernel variant kBias; ptr == nullptr otherwise
  int d_actual;        // head dim of the tensors (multiple of 8, <= the kernel's D): columns beyond it are not stored
  int tail_heads;      // (batch, head) pairs at the end of the item list that are dealt block-major; set by the launcher
};
cudaError_t launch_bwd_dq(int D, bool bf16, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                          const CUtensorMap& tdo, const CUtensorMap& tdq, const DqParams& p, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// deterministic backward with spilled dS (bwd_dkdv_sm100.cu, bwd_dq_spill_sm100.cu)
// ---------------------------------------------------------------------------------------------
// The dK/dV kernel writes every dS^T tile it forms (128 keys x 128 queries, 16-bit = 32 KB, laid out as
// [16 groups of 8 queries][128 keys][16 B]: coalesced to write from registers and directly usable as a no-swizzle
// MN-major tcgen05.mma operand) to a workspace; the dQ kernel then only has to form dQ = sum_n dS(m, n) K(n) - one
// product per block pair instead of three, in a fixed order.
// Workspace layout: [B][Hq][visible (m, n) pairs, m-major][32 KB]; the tiles of one query block are contiguous.
constexpr int kDsTileBytes = 128 * 128 * 2;

// key blocks (128 rows) visible to query block m under the bottom-right aligned causal mask
__host__ __device__ inline int visible_kv_blocks_of(int m, int Sq, int Sk, int causal) {
  const int row0 = m * 128;
  if (row0 >= Sq) return 0;
  const int last_row = (row0 + 128 < Sq ? row0 + 128 : Sq) - 1;
  int cols = Sk;
  if (causal && last_row + Sk - Sq + 1 < cols) cols = last_row + Sk - Sq + 1;
  return cols <= 0 ? 0 : (cols + 127) / 128;
}
// number of spilled tiles that precede query block m inside one (batch, head)
__host__ __device__ inline long long ds_tile_prefix(int m, int Sq, int Sk, int causal) {
  if (!causal) return static_cast<long long>(m) * ((Sk + 127) / 128);
  long long s = 0;
  for (int i = 0; i < m; ++i) s += visible_kv_blocks_of(i, Sq, Sk, causal);
  return s;
}

// dK / dV kernel of the deterministic backward; `ds_spill` != nullptr: also write the dS tiles
struct DkdvParams {
  const float* lse;    // [B, Hq, lse_stride] log2 domain
  const float* delta;  // [B, Hq, lse_stride]
  void* ds_spill;      // workspace described above, or nullptr
  long long tiles_per_head;
  int B, Hq, Hkv, Sq, Sk;
  int lse_stride;
  float softmax_scale;
  float scale_log2;
  int causal;
  int* zero_word;         // work-item counter of the dQ kernel that follows on the stream: cleared by CTA (0, 0, 0), or nullptr
  BiasView bias;          // kernel variant kBias (one CTA per key block only); ptr == nullptr otherwise
  // kBias: a 16-bit bias whose rows TMA can address (16-byte aligned base and strides) is staged tile by tile through
  // shared memory with `bias_map` ((Sk, Sq, 1 | Hq, 1 | B), box 64 x 128); otherwise every thread loads its values itself
  int bias_tma;
  const CUtensorMap* bias_map;          // host pointer, read by the launcher only
  int bias_head_mul, bias_batch_mul;    // 0 for a broadcast dimension, else 1
  const int* seqlens;     // variable-length mode (one CTA per key block only), else nullptr
  const CUtensorMap* vmaps;
  int d_steps, d_cols;    // as in FwdParams: k-steps of the products that reduce over the head dim, N of those that span it
  // kernel variant kExtra == 3 (large query-head groups): the grid's "kv head" index counts Hkv * kv_split parts, part
  // i of kv head h covering the query heads [h * group + i * group / kv_split, ...); every part writes fp32 partial
  // dK / dV to `part_dk` / `part_dv` ([B, Sk, Hkv * kv_split, D] fp32, contiguous) and fa2_dkdv_reduce_kernel sums them.
  int kv_split;
  float* part_dk;
  float* part_dv;
};

// dk[b, s, h, :] = sum over the kv_split partials of kv head h, in a fixed order, rounded once to the output dtype
struct DkdvReduceParams {
  const float* part_dk;
  const float* part_dv;
  void* dk;
  void* dv;
  int64_t dk_strides[3], dv_strides[3];  // batch, head, seq (elements)
  int B, Sk, Hkv, D, kv_split;
};
cudaError_t launch_dkdv_reduce(bool bf16, const DkdvReduceParams& p, cudaStream_t stream);
cudaError_t launch_bwd_dkdv(int D, bool bf16, bool spill, const CUtensorMap& tq, const CUtensorMap& tk,
                            const CUtensorMap& tv, const CUtensorMap& tdo, const CUtensorMap& tdk,
                            const CUtensorMap& tdv, const DkdvParams& p, cudaStream_t stream);

// dQ from the spilled dS tiles: persistent, one product per block pair, HBM-bound (reads every tile once)
struct DqSpillParams {
  const void* ds_spill;
  long long tiles_per_head;
  int* sched;          // work-item counter, 0 on entry (the kernel re-zeroes it)
  void* dq;            // [B, Sq, Hq, D]
  int64_t dq_strides[3];
  int dq_is_16bit;
  int B, Hq, Hkv, Sq, Sk;
  float softmax_scale;
  int causal;
  int d_actual;        // head dim of the tensors (multiple of 8, <= the kernel's D): columns beyond it are not stored
  const int* seqlens;  // variable-length mode, else nullptr
  const CUtensorMap* vmaps;
  int tail_heads;      // (batch, head) pairs at the end of the item list that are dealt block-major; set by the launcher
};
cudaError_t launch_bwd_dq_spill(int D, bool bf16, const CUtensorMap& tk, const DqSpillParams& p, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// deterministic backward for head dims 129..256 (bwd256_sm100.cu)
// ---------------------------------------------------------------------------------------------
// Workspace: [scheduler counter, 1 KB][dS tiles][P tiles], each tile region [B][Hq][visible (m, n) pairs, m-major][32 KB]
// like the spilled-dS path above, a tile being [16 groups of 8 keys][128 queries][16 B].
struct Bwd256Params {
  const float* lse;    // [B, Hq, lse_stride] log2 domain
  const float* delta;  // [B, Hq, lse_stride]
  void* ds_spill;
  void* p_spill;
  long long tiles_per_head;
  int* sched;          // work-item counter of the dK / dV kernel, 0 on entry (the kernel re-zeroes it)
  void* dq;            // [B, Sq, Hq, D] or nullptr (dQ not wanted)
  int64_t dq_strides[3];  // batch, head, seq (elements)
  int dq_is_16bit;
  void* dk;            // [B, Sk, Hkv, D], input dtype
  void* dv;
  int64_t dk_strides[3];
  int64_t dv_strides[3];
  int B, Hq, Hkv, Sq, Sk;
  int lse_stride;
  float softmax_scale;
  float scale_log2;
  int causal;
  BiasView bias;       // kernel variant kBias; ptr == nullptr otherwise
  int d_actual;        // head dim of the tensors (multiple of 8, 129..256)
  int d_steps, d_cols; // ceil(d / 16) k-steps of the products that reduce over the head dim; d rounded up to 16 = N of the others
};
// K / V maps with box rows 64, Q / dO maps with box rows 128
cudaError_t launch_bwd_dq256(bool bf16, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                             const CUtensorMap& tdo, const Bwd256Params& p, cudaStream_t stream);
cudaError_t launch_bwd_acc256(bool bf16, const CUtensorMap& tq, const CUtensorMap& tdo, const Bwd256Params& p,
                              cudaStream_t stream);

struct DqConvertParams {
  const float* dq_accum;  // as written by the backward kernel
  void* dq;               // [B, Sq, Hq, D] with strides below
  int B, Hq, Sq, D;
  int sq_pad;
  int64_t dq_strides[3];  // batch, head, seq (elements)
  float scale;            // softmax_scale (dS is accumulated unscaled)
  int out_dtype;          // 16 = fp16, 17 = bf16, 32 = fp32
};
cudaError_t launch_dq_convert(const DqConvertParams& p, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// UMMA / TMA probe (probe_sm100.cu): one 128 x N x K tile product through every operand path
// the attention kernels use. Test infrastructure for tests/test_umma_probe.py.
// ---------------------------------------------------------------------------------------------
struct ProbeParams {
  const void* a;  // 16-bit, layout depends on a_mode
  const void* b;  // 16-bit, layout depends on b_mode
  float* d;       // [128, N] fp32 row-major
  int N;          // 64 or 128
  int K;          // 64 or 128
  int a_mode;     // 0: smem K-major via TMA ([128,K] row-major)   1: smem MN-major, written by threads ([K,128] row-major)
                  // 2: TMEM (threads pack 16-bit pairs, tcgen05.st)  ([128,K] row-major)
  int b_mode;     // 0: smem K-major via TMA ([N,K] row-major)     1: smem MN-major via TMA ([K,N] row-major)
};
cudaError_t launch_probe(bool bf16, const CUtensorMap& ta, const CUtensorMap& tb, const ProbeParams& p,
                         cudaStream_t stream);

#ifdef FA2_TRACE
cudaError_t fetch_fwd_trace(long long* host, size_t n);
cudaError_t fetch_bwd_trace(long long* host, size_t n);
cudaError_t fetch_dq_trace(long long* host, size_t n);
cudaError_t fetch_dkdv_trace(long long* host, size_t n);
#endif

}  // namespace fa2
