/*
 * fa2_sm100.h - C ABI of libfa2_sm100.so: FlashAttention-2 forward / backward for NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for the hot path of remi-or/fa2_triton. Each entry point replaces one
 * Triton launch site of the reference (paths relative to the reference repository):
 *
 *   fa2_sm100_fwd            <- `_fwd_kernel[grid](...)`       src/forward/caller.py:83-116
 *   fa2_sm100_bwd_delta      <- `_compute_delta[grid](...)`    src/backward/caller.py:94-114
 *   fa2_sm100_bwd            <- `_bwd_kernel[grid](...)`       src/backward/caller.py:118-160
 *   fa2_sm100_fwd_ext / fa2_sm100_bwd_ext: the same two launch sites with the reference's optional arguments
 *                               `bias` (src/forward/caller.py:66, 89, 97) and forward dropout (`dropout_p`,
 *                               `dropout_seed`, src/forward/caller.py:67, 91-92)
 *   fa2_sm100_*_varlen       <- the same launch sites in the reference's VARLEN mode (cum_seqlens_q argument,
 *                               src/forward/caller.py:46-59, src/backward/caller.py:56-79), without its pack / unpack copies
 *   fa2_sm100_dq_convert     <- the fp32 `dq` the reference allocates at src/backward/caller.py:86
 *                               (in the atomic modes our dQ accumulator is an opaque tile-swizzled workspace; this
 *                               call turns it into the caller-visible [B, Sq, Hq, D] tensor)
 *
 * Conventions shared by every call
 *   - tensors are device pointers; q/o/do are [B, Sq, Hq, D], k/v/dk/dv are [B, Sk, Hkv, D] with a contiguous
 *     last dimension; `*_strides[3]` are the (batch, head, seq) strides in ELEMENTS - the same triple the
 *     reference passes as `x.stride(0), x.stride(2), x.stride(1)` (src/forward/caller.py:93-96)
 *   - `dtype` uses the reference's autotune-key encoding (src/utils.py:102-109): 16 = float16, 17 = bfloat16
 *   - lse / delta are fp32 [B, Hq, lse_row_stride] with lse_row_stride = ceil(Sq / 128) * 128
 *     (src/forward/caller.py:73-74); LSE is in the log2 domain: m + log2(l) with scores pre-multiplied by
 *     softmax_scale * log2(e) (src/forward/kernel.py:119)
 *   - causal masking is bottom-right aligned: key j is visible to query i iff j <= i + Sk - Sq
 *     (src/forward/compute_row_blocks.py:64-66)
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work (no allocation, no sync)
 *   - return value: 0 on success; > 0 a cudaError_t / CUresult from the driver; < 0 an FA2_E_* argument error.
 *     fa2_sm100_last_error() returns a thread-local description of the last failure.
 *   - thread safety: calls may be made concurrently from different host threads.
 *
 * Supported in this build: head dims D = 8, 16, ..., 256 (the kernels exist for 64, 128 and 256; a smaller multiple of 8
 * runs in the next one, TMA zero-filling the missing columns - the reference pads to a power of two inside its kernels the
 * same way, src/forward/caller.py:77-78). Above 128: forward without dropout; backward through one deterministic path
 * (`deterministic` 1, 3, 4 or 5; no variable-length mode). The fused backward modes 0 and 2 need D = 64 or 128.
 * Hq % Hkv == 0; any Sq, Sk >= 1; base pointers 16-byte aligned and every stride a multiple of 8 elements (TMA
 * requirement). A stride of 0 is accepted only for a dimension of extent 1: an expanded (broadcast) tensor must be
 * materialised by the caller (FA2_E_BAD_SHAPE otherwise) - the Python host shim does so.
 */
#ifndef FA2_SM100_H_
#define FA2_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FA2_SM100_ABI_VERSION 13

#define FA2_DTYPE_FP16 16
#define FA2_DTYPE_BF16 17
#define FA2_DTYPE_FP32 32

/* how dQ is summed over key blocks (the `deterministic` argument) */
#define FA2_DQ_ATOMIC 0         /* fused kernel, fp32 adds into an accumulator in arrival order (fast mode) */
#define FA2_DQ_DETERMINISTIC 1  /* separate dQ kernel, one CTA per query block, fixed key-block order, no atomics */
#define FA2_DQ_FUSED_ORDERED 2  /* fused kernel, adds serialised per query tile by a ticket (bit-repeatable, slowest) */
#define FA2_DQ_NONE 3           /* dK and dV only (q does not require a gradient); dq may be NULL */
#define FA2_DQ_ONLY 4           /* dQ only, deterministic kernel (k, v do not require gradients); dk, dv may be NULL */
#define FA2_DQ_DET_SPILL 5      /* deterministic, 5 products per block pair: the dK/dV kernel also writes its 16-bit dS tiles
                                   to the workspace and the dQ kernel sums dS K over them in a fixed key-block order
                                   (no recomputation, no atomics; workspace = 32 KB per visible 128 x 128 block pair) */
#define FA2_DQ_SPILL_DKDV 6     /* first half of mode 5 alone: dK, dV and the dS tiles (dq may be NULL) */
#define FA2_DQ_SPILL_DQ 7       /* second half of mode 5 alone: dQ from the tiles a mode 5 / 6 call left in `workspace`
                                   with the same shapes (dk, dv may be NULL) */

#define FA2_E_UNSUPPORTED_HEAD_DIM (-1)
#define FA2_E_UNSUPPORTED_DTYPE (-2)
#define FA2_E_BAD_SHAPE (-3)
#define FA2_E_BAD_ALIGNMENT (-4)
#define FA2_E_NULL_POINTER (-5)
#define FA2_E_WORKSPACE_TOO_SMALL (-6)
#define FA2_E_NO_DRIVER (-7)

int fa2_sm100_version(void);
const char* fa2_sm100_last_error(void);

/* O = softmax(Q K^T * softmax_scale [+ causal mask]) V and the log2-domain LSE. */
int fa2_sm100_fwd(const void* q, const void* k, const void* v, void* o, float* lse,
                  int B, int Hq, int Hkv, int Sq, int Sk, int D,
                  const int64_t q_strides[3], const int64_t k_strides[3], const int64_t v_strides[3],
                  const int64_t o_strides[3], int64_t lse_row_stride,
                  float softmax_scale, int causal, int dtype, void* stream);

/* fa2_sm100_fwd with the reference's optional score modifiers (either may be off):
 *   bias          additive attention bias, logically [B, Hq, Sq, Sk] with a contiguous last dimension, added to the scaled
 *                 scores before the softmax (src/forward/compute_row_blocks.py:48-54, 68-69). `bias_strides` = (batch,
 *                 head, row) strides in elements; 0 for a dimension of size 1 that broadcasts. `bias_dtype` = `dtype` or
 *                 32 (fp32). NULL: no bias. -inf entries act as a mask. Fast path (both passes): a 16-bit bias whose base is
 *                 16-byte aligned and whose strides are multiples of 8 elements is staged through shared memory by TMA; any
 *                 other bias is read with plain loads (same results bit for bit, about 1.8x the forward time).
 *   dropout_p     in [0, 1): each probability is dropped iff tl.rand(dropout_seed, linear index in [B, Hq, Sq, Sk]) <= p
 *                 - Triton's Philox-4x32-10 stream, bit-identical to the reference's mask (src/forward/
 *                 compute_row_blocks.py:76-79, tests/utils.py:169-207) - after the row sum was taken, and O is scaled by
 *                 1 / (1 - p) (src/forward/kernel.py:256-257). The LSE is that of the undropped row. 0: no dropout.
 *   dropout_offset  added to every counter, modulo 2^32 (ABI v12). 0 for a whole batch. With `attention_mask` the reference
 *                 numbers the counters of a sequence of length L that starts at row c of the PACKED batch as
 *                 L * (c + L * head) + row * L + col (src/forward/kernel.py:105-112, 147-149): a call on that sequence's
 *                 [1, L, H, D] views with dropout_offset = L * c draws exactly that mask.
 * The reference has no dropout in its backward (src/utils.py:87-88: NotImplementedError); neither has this library. */
int fa2_sm100_fwd_ext(const void* q, const void* k, const void* v, void* o, float* lse,
                      int B, int Hq, int Hkv, int Sq, int Sk, int D,
                      const int64_t q_strides[3], const int64_t k_strides[3], const int64_t v_strides[3],
                      const int64_t o_strides[3], int64_t lse_row_stride,
                      float softmax_scale, int causal, int dtype,
                      const void* bias, int bias_dtype, const int64_t bias_strides[3],
                      float dropout_p, uint64_t dropout_seed, uint32_t dropout_offset, void* stream);

/* delta[b, h, i] = sum_d O[b, i, h, d] * dO[b, i, h, d] in fp32; rows in [Sq, lse_row_stride) are set to 0.
 * If `workspace` is not NULL the same launch also zero-fills it (the reference zero-allocates dq at
 * src/backward/caller.py:86); that is the state fa2_sm100_bwd expects to find its workspace in. */
int fa2_sm100_bwd_delta(const void* o, const void* dout, float* delta,
                        int B, int Hq, int Sq, int D,
                        const int64_t o_strides[3], const int64_t do_strides[3], int64_t lse_row_stride,
                        int dtype, void* workspace, size_t workspace_bytes, void* stream);

/* Bytes of scratch fa2_sm100_bwd needs. FA2_DQ_ATOMIC / FA2_DQ_FUSED_ORDERED: the fp32 dQ accumulator plus one int per
 * (batch, head, 128-row query tile) and one more (mode 2 numbers its thread blocks in the order they START with it, so a
 * block that waits for its turn only ever waits for blocks that are running: no assumption about the order in which the
 * hardware starts thread blocks). FA2_DQ_DETERMINISTIC / FA2_DQ_ONLY / FA2_DQ_NONE: 256 bytes (work-item counters of
 * the persistent kernels; cleared by the call itself). FA2_DQ_DET_SPILL (and 6, 7): 1 KB + B * Hq * (visible block pairs) * 32 KB
 * (`causal` decides which pairs are visible; B=4, H=32, S=4096 causal: 2.06 GiB).
 * The workspace must be 16-byte aligned. */
size_t fa2_sm100_bwd_workspace_bytes(int B, int Hq, int Sq, int Sk, int D, int causal, int deterministic);

/* The same with room for splitting large groups of query heads: the dK / dV kernel runs one CTA per (key block, kv head,
 * batch element) that loops over all Hq / Hkv query heads of the kv head - with few kv heads (MQA: B = 4, S = 4096 gives
 * 128 CTAs of 32 heads each for 148 SMs) the grid neither fills nor balances the chip. Given a workspace of THIS size,
 * fa2_sm100_bwd cuts a group of 8 or more query heads into 2, 4 or 8 parts of at least 4 heads, lets each part write fp32
 * partial dK / dV to the tail of the workspace and sums them in a fixed order with one rounding (deterministic; the
 * reference sums ALL query heads in 16-bit torch ops, src/backward/caller.py:163-165). With the smaller size above the
 * call runs unsplit. The choice depends on the group size only: a (batch, kv-head) shard of a call stays bit-identical
 * to the whole call. */
size_t fa2_sm100_bwd_workspace_bytes_gqa(int B, int Hq, int Hkv, int Sq, int Sk, int D, int causal, int deterministic);

/* dQ ([B, Sq, Hq, D], `dq_dtype` = the input dtype or 32 for fp32 as the reference allocates it at
 * src/backward/caller.py:86), dK and dV (input dtype, [B, Sk, Hkv, D], GQA head groups already summed).
 * `deterministic` selects the dQ path (FA2_DQ_*). Modes 0 and 2 need `workspace` all zero on entry: pass it to
 * fa2_sm100_bwd_delta (same stream) or clear it yourself. Modes 1, 4 and 5 clear what they need themselves; one call
 * per workspace may be in flight at a time. Results of modes 1, 2, 4 and 5 are bit-identical run to run. */
int fa2_sm100_bwd(const void* dout, const void* q, const void* k, const void* v,
                  const float* lse, const float* delta,
                  void* dq, void* dk, void* dv, void* workspace, size_t workspace_bytes,
                  int B, int Hq, int Hkv, int Sq, int Sk, int D,
                  const int64_t do_strides[3], const int64_t q_strides[3], const int64_t k_strides[3],
                  const int64_t v_strides[3], const int64_t dq_strides[3], const int64_t dk_strides[3],
                  const int64_t dv_strides[3],
                  int64_t lse_row_stride, float softmax_scale, int causal, int deterministic, int dtype, int dq_dtype,
                  void* stream);

/* fa2_sm100_bwd with the additive attention bias of the forward (`bias`, `bias_dtype`, `bias_strides` as in
 * fa2_sm100_fwd_ext; the bias itself gets no gradient, like the reference: src/wrapper.py:86). Every deterministic mode
 * (1, 3, 4, 5, 6, 7) takes a bias; the fused modes 0 and 2 return FA2_E_BAD_SHAPE when one is passed. */
int fa2_sm100_bwd_ext(const void* dout, const void* q, const void* k, const void* v,
                      const float* lse, const float* delta,
                      void* dq, void* dk, void* dv, void* workspace, size_t workspace_bytes,
                      int B, int Hq, int Hkv, int Sq, int Sk, int D,
                      const int64_t do_strides[3], const int64_t q_strides[3], const int64_t k_strides[3],
                      const int64_t v_strides[3], const int64_t dq_strides[3], const int64_t dk_strides[3],
                      const int64_t dv_strides[3],
                      int64_t lse_row_stride, float softmax_scale, int causal, int deterministic, int dtype, int dq_dtype,
                      const void* bias, int bias_dtype, const int64_t bias_strides[3], void* stream);

/* ---- variable-length batches (the reference's attention_mask mode, src/forward/caller.py:25-59) ----
 * Tensors stay right-padded [B, S, H, D]; `seqlens` is a DEVICE array of B int32 lengths (query and key lengths coincide,
 * src/forward/caller.py:28; values are clamped to [0, S]). Nothing is packed or unpacked and the host never reads a
 * length: a small setup kernel derives, for every batch element, tensor maps bounded by its length (tensormap.replace),
 * into `tmap_workspace` (device, fa2_sm100_varlen_workspace_bytes(B) bytes, 128-byte aligned, must stay untouched until
 * the call's kernels have run), and every attention kernel runs ONCE over the whole batch. TMA zero-fills rows at or
 * beyond a sequence's end on load and clips them on store, so the content of the padding (NaNs included) cannot reach a
 * result. Rows of o / lse / dq / dk / dv that belong to the padding are NOT written: pass them zero-filled, as the
 * reference allocates them (src/utils.py:26, src/forward/caller.py:70-74, src/backward/caller.py:86-88).
 * fa2_sm100_bwd_varlen takes `deterministic` = 5 (FA2_DQ_DET_SPILL), 3 (FA2_DQ_NONE) or 6; the workspace size is that of
 * the padded shape (fa2_sm100_bwd_workspace_bytes with Sq = Sk = S). */
size_t fa2_sm100_varlen_workspace_bytes(int B);

int fa2_sm100_fwd_varlen(const void* q, const void* k, const void* v, void* o, float* lse,
                         int B, int Hq, int Hkv, int S, int D,
                         const int64_t q_strides[3], const int64_t k_strides[3], const int64_t v_strides[3],
                         const int64_t o_strides[3], int64_t lse_row_stride,
                         float softmax_scale, int causal, int dtype,
                         const int32_t* seqlens, void* tmap_workspace, size_t tmap_workspace_bytes, void* stream);

int fa2_sm100_bwd_delta_varlen(const void* o, const void* dout, float* delta,
                               int B, int Hq, int S, int D,
                               const int64_t o_strides[3], const int64_t do_strides[3], int64_t lse_row_stride,
                               int dtype, const int32_t* seqlens, void* stream);

int fa2_sm100_bwd_varlen(const void* dout, const void* q, const void* k, const void* v,
                         const float* lse, const float* delta,
                         void* dq, void* dk, void* dv, void* workspace, size_t workspace_bytes,
                         int B, int Hq, int Hkv, int S, int D,
                         const int64_t do_strides[3], const int64_t q_strides[3], const int64_t k_strides[3],
                         const int64_t v_strides[3], const int64_t dq_strides[3], const int64_t dk_strides[3],
                         const int64_t dv_strides[3],
                         int64_t lse_row_stride, float softmax_scale, int causal, int deterministic, int dtype, int dq_dtype,
                         const int32_t* seqlens, void* tmap_workspace, size_t tmap_workspace_bytes, void* stream);

/* dq[b, i, h, :] = softmax_scale * accumulator[b, h, i, :], written as `out_dtype` (16, 17 or 32). Called by
 * fa2_sm100_bwd itself in modes 0 and 2; exported for callers that keep several accumulators in flight. */
int fa2_sm100_dq_convert(const void* workspace, void* dq,
                         int B, int Hq, int Sq, int D,
                         const int64_t dq_strides[3], float softmax_scale, int out_dtype, void* stream);

/* Pitched asynchronous copy between (pinned) host memory and the device on `stream`: `rows` rows of `width_bytes`,
 * `to_device` != 0: host -> device, else device -> host. This is how a host-buffer caller moves one (batch, head range)
 * slice of a [B, S, H, D] tensor without packing it first (fa2_triton_b200/host_api.py; bench.py's `e2e`). */
int fa2_sm100_copy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes,
                           size_t rows, int to_device, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* FA2_SM100_H_ */
