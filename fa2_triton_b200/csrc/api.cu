// C ABI of libfa2_sm100.so (see include/fa2_sm100.h): argument validation, TMA tensor-map encoding and
// kernel launches. Host-side restatement of what the reference's callers do around their Triton launches
// (src/forward/caller.py:33-116, src/backward/caller.py:46-160), minus allocation - the caller owns memory.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/fa2_sm100.h"
#include "fa2_kernels.cuh"

namespace {

thread_local char g_last_error[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(static_cast<int>(e), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

bool is_16bit(int dtype) { return dtype == FA2_DTYPE_FP16 || dtype == FA2_DTYPE_BF16; }

// [B, S, H, D] tensor with (batch, head, seq) element strides -> 4-D map (D, S, H, B), box (64, box_rows, 1, 1),
// 128-byte swizzle, zero fill for out-of-bounds rows.
int make_tmap_bshd(CUtensorMap* map, const void* ptr, int dtype, int B, int S, int H, int D, const int64_t strides[3],
                   int box_rows, const char* name) {
  if (!ptr) return fail(FA2_E_NULL_POINTER, "%s: null pointer", name);
  if (reinterpret_cast<uintptr_t>(ptr) % 16 != 0) return fail(FA2_E_BAD_ALIGNMENT, "%s: base not 16-byte aligned", name);
  const int extent[3] = {B, H, S};  // strides[] = (batch, head, seq)
  for (int i = 0; i < 3; ++i) {
    if (strides[i] < 0 || (strides[i] * 2) % 16 != 0)
      return fail(FA2_E_BAD_ALIGNMENT, "%s: stride[%d]=%lld elements is not a non-negative multiple of 8", name, i,
                  static_cast<long long>(strides[i]));
    // A zero stride is only meaningful for a dimension of extent 1 (torch reports arbitrary strides there). An expanded
    // (broadcast) tensor must be materialised by the caller: TMA cannot address it and the rows would alias.
    if (strides[i] == 0 && extent[i] > 1)
      return fail(FA2_E_BAD_SHAPE, "%s: stride[%d] is 0 for a dimension of extent %d (expanded tensor): pass a "
                                   "materialised (contiguous) tensor", name, i, extent[i]);
  }
  // A tensor map is a pure function of (address, dtype, extents, strides, box): a small per-thread cache saves the driver
  // call for the tensors a training step passes again and again (11 maps per forward + backward).
  struct Key {
    const void* ptr; int dtype, B, S, H, D, box_rows; int64_t st[3];
    bool operator==(const Key& o) const {
      return ptr == o.ptr && dtype == o.dtype && B == o.B && S == o.S && H == o.H && D == o.D && box_rows == o.box_rows &&
             st[0] == o.st[0] && st[1] == o.st[1] && st[2] == o.st[2];
    }
  };
  struct Slot { Key key; CUtensorMap map; bool valid; };
  constexpr int kSlots = 64;
  thread_local Slot cache[kSlots] = {};
  const Key key{ptr, dtype, B, S, H, D, box_rows, {strides[0], strides[1], strides[2]}};
  uint64_t h = reinterpret_cast<uintptr_t>(ptr) >> 4;
  h = (h ^ (h >> 17)) * 0x9E3779B97F4A7C15ull + static_cast<uint64_t>(S) * 31 + static_cast<uint64_t>(box_rows) * 7 + static_cast<uint64_t>(strides[2]);
  Slot& slot = cache[(h >> 20) % kSlots];
  if (slot.valid && slot.key == key) {
    *map = slot.map;
    return 0;
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(FA2_E_NO_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t gdim[4] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(S), static_cast<cuuint64_t>(H),
                              static_cast<cuuint64_t>(B)};
  // a size-1 dimension may carry a zero stride in torch; TMA wants something positive and 16-byte granular
  auto fix = [&](int64_t s) { return static_cast<cuuint64_t>((s > 0 ? s : static_cast<int64_t>(D)) * 2); };
  const cuuint64_t gstride[3] = {fix(strides[2]), fix(strides[1]), fix(strides[0])};
  const cuuint32_t box[4] = {64u, static_cast<cuuint32_t>(box_rows), 1u, 1u};
  const cuuint32_t estride[4] = {1u, 1u, 1u, 1u};
  const CUtensorMapDataType dt = dtype == FA2_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(map, dt, 4, const_cast<void*>(ptr), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(static_cast<int>(r), "%s: cuTensorMapEncodeTiled failed with CUresult %d", name, static_cast<int>(r));
  slot.key = key;
  slot.map = *map;
  slot.valid = true;
  return 0;
}

int check_common(int B, int Hq, int Hkv, int Sq, int Sk, int D, int dtype) {
  if (!is_16bit(dtype)) return fail(FA2_E_UNSUPPORTED_DTYPE, "dtype code %d: only 16 (fp16) and 17 (bf16) are supported", dtype);
  if (D < 8 || D > 256 || D % 8 != 0)
    return fail(FA2_E_UNSUPPORTED_HEAD_DIM, "head_dim %d: this build supports multiples of 8 up to 256", D);
  if (B < 1 || Hq < 1 || Hkv < 1 || Sq < 1 || Sk < 1) return fail(FA2_E_BAD_SHAPE, "non-positive dimension");
  if (Hq % Hkv != 0) return fail(FA2_E_BAD_SHAPE, "nheads_q = %d is not divisible by nheads_kv = %d", Hq, Hkv);
  if (Hq > 65535 || B > 65535) return fail(FA2_E_BAD_SHAPE, "batch and head counts must fit a grid dimension");
  return 0;
}

constexpr float kLog2e = 1.4426950408889634f;

// Into how many parts the dK / dV kernel cuts the group of query heads that share a kv head (see
// fa2_sm100_bwd_workspace_bytes_gqa in the header). Looks at the group size only, so that a (batch, kv-head) shard of a
// call makes the same choice as the whole call and stays bit-identical to it.
int kv_split_of(int Hq, int Hkv, int D, int deterministic) {
  if (Hkv < 1 || Hq % Hkv) return 1;
  if (deterministic == FA2_DQ_ATOMIC || deterministic == FA2_DQ_FUSED_ORDERED || deterministic == FA2_DQ_ONLY ||
      deterministic == FA2_DQ_SPILL_DQ)
    return 1;  // the fused kernel has no split variant; modes 4 and 7 do not run the dK / dV kernel
  if (D != 64 && D != 128) return 1;
  const int group = Hq / Hkv;
  for (int s = 8; s >= 2; s /= 2)
    if (group % s == 0 && group / s >= 4) return s;
  return 1;
}

size_t kv_split_bytes(int B, int Hkv, int Sk, int D, int split) {  // the two fp32 partial tensors
  return split > 1 ? 2 * static_cast<size_t>(B) * Sk * Hkv * split * D * sizeof(float) : 0;
}

// variable-length mode: per-sequence copies of the launch's tensor maps, built on the device (varlen_sm100.cu)
struct VarlenTensor { const CUtensorMap* map; const void* ptr; int64_t batch_stride_elems; };

int setup_varlen_maps(const VarlenTensor* tensors, int n, const int32_t* seqlens_dev, void* tmap_ws, size_t tmap_ws_bytes,
                      int B, int S, cudaStream_t st) {
  if (!seqlens_dev || !tmap_ws) return fail(FA2_E_NULL_POINTER, "seqlens / tensor-map workspace: null pointer");
  if (tmap_ws_bytes < fa2_sm100_varlen_workspace_bytes(B))
    return fail(FA2_E_WORKSPACE_TOO_SMALL, "tensor-map workspace has %zu bytes, %zu needed", tmap_ws_bytes,
                fa2_sm100_varlen_workspace_bytes(B));
  if (reinterpret_cast<uintptr_t>(tmap_ws) % 128) return fail(FA2_E_BAD_ALIGNMENT, "tensor-map workspace not 128-byte aligned");
  CUtensorMap templates[6];
  fa2::VarlenSetupParams sp{};
  sp.seqlens = seqlens_dev;
  sp.out = static_cast<CUtensorMap*>(tmap_ws);
  sp.n_maps = n;
  sp.S = S;
  for (int i = 0; i < n; ++i) {
    templates[i] = *tensors[i].map;
    sp.base[i] = reinterpret_cast<uint64_t>(tensors[i].ptr);
    sp.batch_stride_bytes[i] = static_cast<uint64_t>(tensors[i].batch_stride_elems) * 2u;
  }
  cudaError_t ce = fa2::launch_varlen_maps(templates, sp, B, st);
  if (ce != cudaSuccess) return cuda_fail(ce, "fa2_varlen_maps_kernel launch");
  return 0;
}

// bias: [1 | B, 1 | Hq, Sq, Sk] with (batch, head, row) element strides; a broadcast dimension carries stride 0
int make_bias_view(fa2::BiasView* view, const void* bias, int bias_dtype, const int64_t bias_strides[3], int dtype) {
  *view = fa2::BiasView{};
  if (!bias) return 0;
  if (bias_dtype != dtype && bias_dtype != FA2_DTYPE_FP32)
    return fail(FA2_E_UNSUPPORTED_DTYPE, "bias dtype code %d: must be the input dtype (%d) or 32 (fp32)", bias_dtype, dtype);
  if (!bias_strides) return fail(FA2_E_NULL_POINTER, "bias_strides: null pointer");
  const int elem = bias_dtype == FA2_DTYPE_FP32 ? 4 : 2;
  if (reinterpret_cast<uintptr_t>(bias) % elem) return fail(FA2_E_BAD_ALIGNMENT, "bias: base not element aligned");
  for (int i = 0; i < 3; ++i)
    if (bias_strides[i] < 0) return fail(FA2_E_BAD_SHAPE, "bias: negative stride");
  view->ptr = bias;
  view->is_fp32 = bias_dtype == FA2_DTYPE_FP32 ? 1 : 0;
  view->sb = bias_strides[0];
  view->sh = bias_strides[1];
  view->sm = bias_strides[2];
  return 0;
}

int round_up(int x, int m) { return (x + m - 1) / m * m; }

// The kernels are instantiated for head dims 64, 128 and 256. Any multiple of 8 below runs in the next instantiation: the
// tensor maps carry the real head dim, so TMA zero-fills the missing columns on load and clips them on store.
int kernel_head_dim(int D) { return D <= 64 ? 64 : D <= 128 ? 128 : 256; }

int fwd_impl(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Hq, int Hkv, int Sq,
             int Sk, int D, const int64_t q_strides[3], const int64_t k_strides[3], const int64_t v_strides[3],
             const int64_t o_strides[3], int64_t lse_row_stride, float softmax_scale, int causal, int dtype,
             const void* bias, int bias_dtype, const int64_t bias_strides[3], float dropout_p,
             uint64_t dropout_seed, uint32_t dropout_offset, const int32_t* seqlens, void* tmap_ws, size_t tmap_ws_bytes,
             void* stream);
int bwd_impl(const void* dout, const void* q, const void* k, const void* v, const float* lse,
             const float* delta, void* dq, void* dk, void* dv, void* workspace, size_t workspace_bytes, int B,
             int Hq, int Hkv, int Sq, int Sk, int D, const int64_t do_strides[3], const int64_t q_strides[3],
             const int64_t k_strides[3], const int64_t v_strides[3], const int64_t dq_strides[3],
             const int64_t dk_strides[3], const int64_t dv_strides[3], int64_t lse_row_stride,
             float softmax_scale, int causal, int deterministic, int dtype, int dq_dtype, const void* bias,
             int bias_dtype, const int64_t bias_strides[3], const int32_t* seqlens, void* tmap_ws, size_t tmap_ws_bytes,
             void* stream);
int delta_impl(const void* o, const void* dout, float* delta, int B, int Hq, int Sq, int D,
               const int64_t o_strides[3], const int64_t do_strides[3], int64_t lse_row_stride, int dtype,
               void* workspace, size_t workspace_bytes, const int32_t* seqlens, void* stream);

}  // namespace

extern "C" {

size_t fa2_sm100_varlen_workspace_bytes(int B) {
  return B < 1 ? 0 : static_cast<size_t>(B) * fa2::kVarlenMapsPerBatch * sizeof(CUtensorMap);
}

int fa2_sm100_version(void) { return FA2_SM100_ABI_VERSION; }

const char* fa2_sm100_last_error(void) { return g_last_error; }

int fa2_sm100_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Hq, int Hkv, int Sq,
                  int Sk, int D, const int64_t q_strides[3], const int64_t k_strides[3], const int64_t v_strides[3],
                  const int64_t o_strides[3], int64_t lse_row_stride, float softmax_scale, int causal, int dtype,
                  void* stream) {
  return fa2_sm100_fwd_ext(q, k, v, o, lse, B, Hq, Hkv, Sq, Sk, D, q_strides, k_strides, v_strides, o_strides,
                           lse_row_stride, softmax_scale, causal, dtype, nullptr, 0, nullptr, 0.f, 0ull, 0u, stream);
}

int fa2_sm100_fwd_ext(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Hq, int Hkv, int Sq,
                      int Sk, int D, const int64_t q_strides[3], const int64_t k_strides[3], const int64_t v_strides[3],
                      const int64_t o_strides[3], int64_t lse_row_stride, float softmax_scale, int causal, int dtype,
                      const void* bias, int bias_dtype, const int64_t bias_strides[3], float dropout_p,
                      uint64_t dropout_seed, uint32_t dropout_offset, void* stream) {
  return fwd_impl(q, k, v, o, lse, B, Hq, Hkv, Sq, Sk, D, q_strides, k_strides, v_strides, o_strides, lse_row_stride,
                  softmax_scale, causal, dtype, bias, bias_dtype, bias_strides, dropout_p, dropout_seed, dropout_offset,
                  nullptr, nullptr, 0, stream);
}

int fa2_sm100_fwd_varlen(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Hq, int Hkv, int S,
                         int D, const int64_t q_strides[3], const int64_t k_strides[3], const int64_t v_strides[3],
                         const int64_t o_strides[3], int64_t lse_row_stride, float softmax_scale, int causal, int dtype,
                         const int32_t* seqlens, void* tmap_workspace, size_t tmap_workspace_bytes, void* stream) {
  if (!seqlens) return fail(FA2_E_NULL_POINTER, "seqlens: null pointer");
  return fwd_impl(q, k, v, o, lse, B, Hq, Hkv, S, S, D, q_strides, k_strides, v_strides, o_strides, lse_row_stride,
                  softmax_scale, causal, dtype, nullptr, 0, nullptr, 0.f, 0ull, 0u, seqlens, tmap_workspace,
                  tmap_workspace_bytes, stream);
}

}  // extern "C"

namespace {

int fwd_impl(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Hq, int Hkv, int Sq,
             int Sk, int D, const int64_t q_strides[3], const int64_t k_strides[3], const int64_t v_strides[3],
             const int64_t o_strides[3], int64_t lse_row_stride, float softmax_scale, int causal, int dtype,
             const void* bias, int bias_dtype, const int64_t bias_strides[3], float dropout_p,
             uint64_t dropout_seed, uint32_t dropout_offset, const int32_t* seqlens, void* tmap_ws, size_t tmap_ws_bytes,
             void* stream) {
  if (int e = check_common(B, Hq, Hkv, Sq, Sk, D, dtype)) return e;
  if (!(dropout_p >= 0.f && dropout_p < 1.f)) return fail(FA2_E_BAD_SHAPE, "dropout_p must be in [0, 1)");
  if (!lse) return fail(FA2_E_NULL_POINTER, "lse: null pointer");
  if (lse_row_stride != round_up(Sq, 128)) return fail(FA2_E_BAD_SHAPE, "lse_row_stride must be ceil(Sq/128)*128");
  if (!(softmax_scale > 0.f)) return fail(FA2_E_BAD_SHAPE, "softmax_scale must be positive");
  CUtensorMap tq, tk, tv, to;
  if (int e = make_tmap_bshd(&tq, q, dtype, B, Sq, Hq, D, q_strides, 128, "q")) return e;
  // head dims above 128: the forward stages K / V in tiles of 64 keys (fwd2_sm100.cu)
  const int kv_box_rows = kernel_head_dim(D) == 256 ? 64 : 128;
  if (int e = make_tmap_bshd(&tk, k, dtype, B, Sk, Hkv, D, k_strides, kv_box_rows, "k")) return e;
  if (int e = make_tmap_bshd(&tv, v, dtype, B, Sk, Hkv, D, v_strides, kv_box_rows, "v")) return e;
  if (int e = make_tmap_bshd(&to, o, dtype, B, Sq, Hq, D, o_strides, 128, "o")) return e;
  if (D > 128 && dropout_p > 0.f)
    return fail(FA2_E_UNSUPPORTED_HEAD_DIM, "head_dim %d: forward dropout is built for head dims up to 128", D);
  if (D > 128 && seqlens && bias) return fail(FA2_E_BAD_SHAPE, "bias together with variable-length mode");
  fa2::FwdParams p{};
  p.lse = lse;
  p.B = B; p.Hq = Hq; p.Hkv = Hkv; p.Sq = Sq; p.Sk = Sk;
  p.lse_stride = static_cast<int>(lse_row_stride);
  p.scale_log2 = softmax_scale * kLog2e;
  p.causal = causal ? 1 : 0;
  p.n_pairs = (Sq + 255) / 256;
  // head dims below the kernel's: the products of the zero-filled columns are not issued
  p.d_steps = (D + 15) / 16;
  p.d_cols = round_up(D, 16);
  if (int e = make_bias_view(&p.bias, bias, bias_dtype, bias_strides, dtype)) return e;
  p.bias_mul = 1.0f / softmax_scale;
  CUtensorMap tbias;
  if (bias && bias_dtype == dtype && D <= 128) {
    // same rule as the dK / dV kernel's bias map below: [1 | B, 1 | Hq, Sq, Sk] seen as a [B', Sq, H', Sk] tensor, only if
    // TMA can address it (16-byte aligned base, strides multiples of 8 elements); plain loads serve every other bias
    const int Bb = bias_strides[0] ? B : 1, Hb = bias_strides[1] ? Hq : 1;
    const int64_t bs[3] = {bias_strides[0], bias_strides[1], bias_strides[2]};
    const bool ok = reinterpret_cast<uintptr_t>(bias) % 16 == 0 && bs[0] % 8 == 0 && bs[1] % 8 == 0 && bs[2] % 8 == 0 &&
                    bs[2] > 0;
    if (ok && make_tmap_bshd(&tbias, bias, dtype, Bb, Sq, Hb, Sk, bs, 128, "bias") == 0) {
      p.bias_tma = 1;
      p.bias_map = &tbias;
      p.bias_head_mul = Hb > 1 ? 1 : 0;
      p.bias_batch_mul = Bb > 1 ? 1 : 0;
    }
  }
  p.dropout_p = dropout_p;
  p.inv_keep = 1.0f / (1.0f - dropout_p);
  p.seed_lo = static_cast<uint32_t>(dropout_seed & 0xffffffffull);
  p.seed_hi = static_cast<uint32_t>(dropout_seed >> 32);
  p.drop_offset = dropout_offset;
  if (seqlens) {
    const VarlenTensor vt[4] = {{&tq, q, q_strides[0]}, {&tk, k, k_strides[0]}, {&tv, v, v_strides[0]}, {&to, o, o_strides[0]}};
    if (int e = setup_varlen_maps(vt, 4, seqlens, tmap_ws, tmap_ws_bytes, B, Sq, static_cast<cudaStream_t>(stream))) return e;
    p.seqlens = seqlens;
    p.vmaps = static_cast<const CUtensorMap*>(tmap_ws);
  }
  cudaError_t ce = fa2::launch_fwd(kernel_head_dim(D), dtype == FA2_DTYPE_BF16, tq, tk, tv, to, p, static_cast<cudaStream_t>(stream));
  if (ce != cudaSuccess) return cuda_fail(ce, "fa2_sm100_fwd launch");
  return 0;
}

}  // namespace

extern "C" {

int fa2_sm100_bwd_delta(const void* o, const void* dout, float* delta, int B, int Hq, int Sq, int D,
                        const int64_t o_strides[3], const int64_t do_strides[3], int64_t lse_row_stride, int dtype,
                        void* workspace, size_t workspace_bytes, void* stream) {
  return delta_impl(o, dout, delta, B, Hq, Sq, D, o_strides, do_strides, lse_row_stride, dtype, workspace, workspace_bytes,
                    nullptr, stream);
}

int fa2_sm100_bwd_delta_varlen(const void* o, const void* dout, float* delta, int B, int Hq, int S, int D,
                               const int64_t o_strides[3], const int64_t do_strides[3], int64_t lse_row_stride, int dtype,
                               const int32_t* seqlens, void* stream) {
  if (!seqlens) return fail(FA2_E_NULL_POINTER, "seqlens: null pointer");
  return delta_impl(o, dout, delta, B, Hq, S, D, o_strides, do_strides, lse_row_stride, dtype, nullptr, 0, seqlens, stream);
}

}  // extern "C"

namespace {

int delta_impl(const void* o, const void* dout, float* delta, int B, int Hq, int Sq, int D,
               const int64_t o_strides[3], const int64_t do_strides[3], int64_t lse_row_stride, int dtype,
               void* workspace, size_t workspace_bytes, const int32_t* seqlens, void* stream) {
  if (int e = check_common(B, Hq, Hq, Sq, 1, D, dtype)) return e;
  if (!o || !dout || !delta) return fail(FA2_E_NULL_POINTER, "o / dout / delta: null pointer");
  if (lse_row_stride != round_up(Sq, 128)) return fail(FA2_E_BAD_SHAPE, "lse_row_stride must be ceil(Sq/128)*128");
  if (reinterpret_cast<uintptr_t>(o) % 16 || reinterpret_cast<uintptr_t>(dout) % 16)
    return fail(FA2_E_BAD_ALIGNMENT, "o / dout: base not 16-byte aligned");
  for (int i = 0; i < 3; ++i)
    if (o_strides[i] % 8 || do_strides[i] % 8) return fail(FA2_E_BAD_ALIGNMENT, "o / dout: strides must be multiples of 8 elements");
  fa2::DeltaParams p{};
  p.o = o; p.dout = dout; p.delta = delta;
  p.B = B; p.Hq = Hq; p.Sq = Sq; p.D = D;
  p.lse_stride = static_cast<int>(lse_row_stride);
  for (int i = 0; i < 3; ++i) { p.o_strides[i] = o_strides[i]; p.do_strides[i] = do_strides[i]; }
  if (workspace) {
    if (D != kernel_head_dim(D))
      return fail(FA2_E_UNSUPPORTED_HEAD_DIM, "head_dim %d: the dQ accumulator of the fused backward needs 64 or 128", D);
    const size_t need = fa2_sm100_bwd_workspace_bytes(B, Hq, Sq, 1, D, 0, FA2_DQ_ATOMIC);
    if (workspace_bytes < need)
      return fail(FA2_E_WORKSPACE_TOO_SMALL, "workspace has %zu bytes, %zu needed", workspace_bytes, need);
    if (reinterpret_cast<uintptr_t>(workspace) % 16) return fail(FA2_E_BAD_ALIGNMENT, "workspace not 16-byte aligned");
    p.dq_accum = static_cast<float*>(workspace);
    p.dq_sem = reinterpret_cast<int*>(static_cast<char*>(workspace) +
                                      static_cast<size_t>(B) * Hq * lse_row_stride * D * sizeof(float));
  }
  p.seqlens = seqlens;
  cudaError_t ce = fa2::launch_delta(dtype == FA2_DTYPE_BF16, p, static_cast<cudaStream_t>(stream));
  if (ce != cudaSuccess) return cuda_fail(ce, "fa2_sm100_bwd_delta launch");
  return 0;
}

}  // namespace

extern "C" {

int fa2_sm100_dq_convert(const void* workspace, void* dq, int B, int Hq, int Sq, int D, const int64_t dq_strides[3],
                         float softmax_scale, int out_dtype, void* stream);

size_t fa2_sm100_bwd_workspace_bytes(int B, int Hq, int Sq, int Sk, int D, int causal, int deterministic) {
  if (B < 1 || Hq < 1 || Sq < 1 || D < 1) return 0;
  if (D > 128) {
    // head dims 129..256 (bwd256_sm100.cu), every supported mode: counter + the dS tiles + the P tiles
    if (Sk < 1) return 0;
    const long long tiles = fa2::ds_tile_prefix((Sq + 127) / 128, Sq, Sk, causal ? 1 : 0);
    return 1024 + 2 * static_cast<size_t>(B) * Hq * static_cast<size_t>(tiles) * fa2::kDsTileBytes;
  }
  if (deterministic == FA2_DQ_NONE || deterministic == FA2_DQ_DETERMINISTIC || deterministic == FA2_DQ_ONLY)
    return 256;  // no accumulator, only the work-item counters of the persistent kernels' schedulers
  if (deterministic == FA2_DQ_DET_SPILL || deterministic == FA2_DQ_SPILL_DKDV || deterministic == FA2_DQ_SPILL_DQ) {
    // scheduler counter (padded to 1 KB) + one 32 KB dS tile per visible (query block, key block) pair
    if (Sk < 1) return 0;
    const long long tiles = fa2::ds_tile_prefix((Sq + 127) / 128, Sq, Sk, causal ? 1 : 0);
    return 1024 + static_cast<size_t>(B) * Hq * static_cast<size_t>(tiles) * fa2::kDsTileBytes;
  }
  const size_t sq_pad = static_cast<size_t>(round_up(Sq, 128));
  const size_t accum = static_cast<size_t>(B) * Hq * sq_pad * D * sizeof(float);
  // one ticket per (batch, head, query tile) + the counter that numbers the CTAs of the ticket-ordered kernel as they start
  const size_t sem = (static_cast<size_t>(B) * Hq * (sq_pad / 128) + 1) * sizeof(int);
  return accum + ((sem + 15) / 16) * 16;
}

size_t fa2_sm100_bwd_workspace_bytes_gqa(int B, int Hq, int Hkv, int Sq, int Sk, int D, int causal, int deterministic) {
  const size_t base = fa2_sm100_bwd_workspace_bytes(B, Hq, Sq, Sk, D, causal, deterministic);
  if (base == 0) return 0;
  return (base + 15) / 16 * 16 + kv_split_bytes(B, Hkv, Sk, D, kv_split_of(Hq, Hkv, D, deterministic));
}

int fa2_sm100_bwd(const void* dout, const void* q, const void* k, const void* v, const float* lse, const float* delta,
                  void* dq, void* dk, void* dv, void* workspace, size_t workspace_bytes, int B, int Hq, int Hkv,
                  int Sq, int Sk, int D, const int64_t do_strides[3], const int64_t q_strides[3],
                  const int64_t k_strides[3], const int64_t v_strides[3], const int64_t dq_strides[3],
                  const int64_t dk_strides[3], const int64_t dv_strides[3], int64_t lse_row_stride, float softmax_scale,
                  int causal, int deterministic, int dtype, int dq_dtype, void* stream) {
  return fa2_sm100_bwd_ext(dout, q, k, v, lse, delta, dq, dk, dv, workspace, workspace_bytes, B, Hq, Hkv, Sq, Sk, D,
                           do_strides, q_strides, k_strides, v_strides, dq_strides, dk_strides, dv_strides,
                           lse_row_stride, softmax_scale, causal, deterministic, dtype, dq_dtype, nullptr, 0, nullptr,
                           stream);
}

int fa2_sm100_bwd_ext(const void* dout, const void* q, const void* k, const void* v, const float* lse,
                      const float* delta, void* dq, void* dk, void* dv, void* workspace, size_t workspace_bytes, int B,
                      int Hq, int Hkv, int Sq, int Sk, int D, const int64_t do_strides[3], const int64_t q_strides[3],
                      const int64_t k_strides[3], const int64_t v_strides[3], const int64_t dq_strides[3],
                      const int64_t dk_strides[3], const int64_t dv_strides[3], int64_t lse_row_stride,
                      float softmax_scale, int causal, int deterministic, int dtype, int dq_dtype, const void* bias,
                      int bias_dtype, const int64_t bias_strides[3], void* stream) {
  return bwd_impl(dout, q, k, v, lse, delta, dq, dk, dv, workspace, workspace_bytes, B, Hq, Hkv, Sq, Sk, D, do_strides,
                  q_strides, k_strides, v_strides, dq_strides, dk_strides, dv_strides, lse_row_stride, softmax_scale, causal,
                  deterministic, dtype, dq_dtype, bias, bias_dtype, bias_strides, nullptr, nullptr, 0, stream);
}

int fa2_sm100_bwd_varlen(const void* dout, const void* q, const void* k, const void* v, const float* lse,
                         const float* delta, void* dq, void* dk, void* dv, void* workspace, size_t workspace_bytes, int B,
                         int Hq, int Hkv, int S, int D, const int64_t do_strides[3], const int64_t q_strides[3],
                         const int64_t k_strides[3], const int64_t v_strides[3], const int64_t dq_strides[3],
                         const int64_t dk_strides[3], const int64_t dv_strides[3], int64_t lse_row_stride,
                         float softmax_scale, int causal, int deterministic, int dtype, int dq_dtype,
                         const int32_t* seqlens, void* tmap_workspace, size_t tmap_workspace_bytes, void* stream) {
  if (!seqlens) return fail(FA2_E_NULL_POINTER, "seqlens: null pointer");
  if (deterministic != FA2_DQ_DET_SPILL && deterministic != FA2_DQ_NONE && deterministic != FA2_DQ_SPILL_DKDV)
    return fail(FA2_E_BAD_SHAPE, "variable-length mode: `deterministic` must be 5 (FA2_DQ_DET_SPILL), 3 (FA2_DQ_NONE) or 6");
  if (!dk || !dv) return fail(FA2_E_NULL_POINTER, "dk / dv: null pointer");
  return bwd_impl(dout, q, k, v, lse, delta, dq, dk, dv, workspace, workspace_bytes, B, Hq, Hkv, S, S, D, do_strides,
                  q_strides, k_strides, v_strides, dq_strides, dk_strides, dv_strides, lse_row_stride, softmax_scale, causal,
                  deterministic, dtype, dq_dtype, nullptr, 0, nullptr, seqlens, tmap_workspace, tmap_workspace_bytes, stream);
}

}  // extern "C"

namespace {

int bwd_impl(const void* dout, const void* q, const void* k, const void* v, const float* lse,
             const float* delta, void* dq, void* dk, void* dv, void* workspace, size_t workspace_bytes, int B,
             int Hq, int Hkv, int Sq, int Sk, int D, const int64_t do_strides[3], const int64_t q_strides[3],
             const int64_t k_strides[3], const int64_t v_strides[3], const int64_t dq_strides[3],
             const int64_t dk_strides[3], const int64_t dv_strides[3], int64_t lse_row_stride,
             float softmax_scale, int causal, int deterministic, int dtype, int dq_dtype, const void* bias,
             int bias_dtype, const int64_t bias_strides[3], const int32_t* seqlens, void* tmap_ws, size_t tmap_ws_bytes,
             void* stream) {
  if (int e = check_common(B, Hq, Hkv, Sq, Sk, D, dtype)) return e;
  fa2::BiasView bias_view{};
  if (int e = make_bias_view(&bias_view, bias, bias_dtype, bias_strides, dtype)) return e;
  const int Dk = kernel_head_dim(D);
  if (D > 128) {
    // Head dims 129..256 (bwd256_sm100.cu): one deterministic path - dQ with spilled P / dS tiles, then dK / dV from the
    // tiles. `deterministic` 1 (recompute) and 5 (spill) both select it; 3 / 4 skip dQ / dK + dV.
    if (deterministic == FA2_DQ_ATOMIC || deterministic == FA2_DQ_FUSED_ORDERED || deterministic == FA2_DQ_SPILL_DKDV ||
        deterministic == FA2_DQ_SPILL_DQ)
      return fail(FA2_E_UNSUPPORTED_HEAD_DIM, "head_dim %d: the backward above head dim 128 has the deterministic modes 1, 3, 4, 5 only", D);
    if (seqlens) return fail(FA2_E_UNSUPPORTED_HEAD_DIM, "head_dim %d: variable-length mode stops at head dim 128 in the backward", D);
    const bool want_dq = deterministic != FA2_DQ_NONE, want_dkdv = deterministic != FA2_DQ_ONLY;
    if (!lse || !delta) return fail(FA2_E_NULL_POINTER, "lse / delta: null pointer");
    if (want_dq && !dq) return fail(FA2_E_NULL_POINTER, "dq: null pointer");
    if (want_dkdv && (!dk || !dv)) return fail(FA2_E_NULL_POINTER, "dk / dv: null pointer");
    if (lse_row_stride != round_up(Sq, 128)) return fail(FA2_E_BAD_SHAPE, "lse_row_stride must be ceil(Sq/128)*128");
    if (!(softmax_scale > 0.f)) return fail(FA2_E_BAD_SHAPE, "softmax_scale must be positive");
    if (want_dq) {
      if (dq_dtype != dtype && dq_dtype != FA2_DTYPE_FP32)
        return fail(FA2_E_UNSUPPORTED_DTYPE, "dq_dtype %d: must be the input dtype (%d) or 32 (fp32)", dq_dtype, dtype);
      if (reinterpret_cast<uintptr_t>(dq) % 16) return fail(FA2_E_BAD_ALIGNMENT, "dq: base not 16-byte aligned");
      for (int i = 0; i < 3; ++i)
        if (dq_strides[i] % 8) return fail(FA2_E_BAD_ALIGNMENT, "dq: strides must be multiples of 8 elements");
    }
    if (want_dkdv) {
      if (reinterpret_cast<uintptr_t>(dk) % 16 || reinterpret_cast<uintptr_t>(dv) % 16)
        return fail(FA2_E_BAD_ALIGNMENT, "dk / dv: base not 16-byte aligned");
      for (int i = 0; i < 3; ++i)
        if (dk_strides[i] % 8 || dv_strides[i] % 8) return fail(FA2_E_BAD_ALIGNMENT, "dk / dv: strides must be multiples of 8 elements");
    }
    const size_t need256 = fa2_sm100_bwd_workspace_bytes(B, Hq, Sq, Sk, D, causal, FA2_DQ_DET_SPILL);
    if (!workspace) return fail(FA2_E_NULL_POINTER, "workspace: null pointer");
    if (workspace_bytes < need256) return fail(FA2_E_WORKSPACE_TOO_SMALL, "workspace has %zu bytes, %zu needed", workspace_bytes, need256);
    if (reinterpret_cast<uintptr_t>(workspace) % 16) return fail(FA2_E_BAD_ALIGNMENT, "workspace not 16-byte aligned");
    CUtensorMap tq, tk, tv, tdo;
    if (int e = make_tmap_bshd(&tq, q, dtype, B, Sq, Hq, D, q_strides, 128, "q")) return e;
    if (int e = make_tmap_bshd(&tk, k, dtype, B, Sk, Hkv, D, k_strides, 64, "k")) return e;
    if (int e = make_tmap_bshd(&tv, v, dtype, B, Sk, Hkv, D, v_strides, 64, "v")) return e;
    if (int e = make_tmap_bshd(&tdo, dout, dtype, B, Sq, Hq, D, do_strides, 128, "dout")) return e;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t ce = cudaMemsetAsync(workspace, 0, 128, st);  // scheduler counter (the kernel re-zeroes it; first use of the buffer)
    if (ce != cudaSuccess) return cuda_fail(ce, "fa2_sm100_bwd scheduler counter memset");
    const long long tiles = fa2::ds_tile_prefix((Sq + 127) / 128, Sq, Sk, causal ? 1 : 0);
    const size_t region = static_cast<size_t>(B) * Hq * static_cast<size_t>(tiles) * fa2::kDsTileBytes;
    fa2::Bwd256Params b{};
    b.lse = lse; b.delta = delta;
    b.ds_spill = static_cast<char*>(workspace) + 1024;
    b.p_spill = static_cast<char*>(workspace) + 1024 + region;
    b.tiles_per_head = tiles;
    b.sched = static_cast<int*>(workspace);
    b.dq = want_dq ? dq : nullptr;
    b.dk = dk; b.dv = dv;
    for (int i = 0; i < 3; ++i) {
      b.dq_strides[i] = want_dq ? dq_strides[i] : 0;
      b.dk_strides[i] = want_dkdv ? dk_strides[i] : 0;
      b.dv_strides[i] = want_dkdv ? dv_strides[i] : 0;
    }
    b.dq_is_16bit = dq_dtype == dtype ? 1 : 0;
    b.B = B; b.Hq = Hq; b.Hkv = Hkv; b.Sq = Sq; b.Sk = Sk;
    b.lse_stride = static_cast<int>(lse_row_stride);
    b.softmax_scale = softmax_scale;
    b.scale_log2 = softmax_scale * kLog2e;
    b.causal = causal ? 1 : 0;
    b.bias = bias_view;
    b.d_actual = D;
    b.d_steps = (D + 15) / 16;
    b.d_cols = round_up(D, 16);
    ce = fa2::launch_bwd_dq256(dtype == FA2_DTYPE_BF16, tq, tk, tv, tdo, b, st);
    if (ce != cudaSuccess) return cuda_fail(ce, "fa2_sm100_bwd dQ (head dim 129..256) launch");
    if (want_dkdv) {
      ce = fa2::launch_bwd_acc256(dtype == FA2_DTYPE_BF16, tq, tdo, b, st);
      if (ce != cudaSuccess) return cuda_fail(ce, "fa2_sm100_bwd dK / dV (head dim 129..256) launch");
    }
    return 0;
  }
  if (D != Dk && (deterministic == FA2_DQ_ATOMIC || deterministic == FA2_DQ_FUSED_ORDERED))
    return fail(FA2_E_UNSUPPORTED_HEAD_DIM, "head_dim %d: the fused (atomic / ticket-ordered) backward needs 64 or 128; "
                                            "the deterministic modes take any multiple of 8 up to 128", D);
  if (bias && (deterministic == FA2_DQ_ATOMIC || deterministic == FA2_DQ_FUSED_ORDERED))
    return fail(FA2_E_BAD_SHAPE, "bias: the fused (atomic / ticket-ordered) backward has no bias variant; use a "
                                 "deterministic mode (1, 3, 4, 5, 6)");
  const bool want_dq = deterministic != FA2_DQ_NONE && deterministic != FA2_DQ_SPILL_DKDV;
  const bool want_dkdv = deterministic != FA2_DQ_ONLY && deterministic != FA2_DQ_SPILL_DQ;
  if (!lse || !delta) return fail(FA2_E_NULL_POINTER, "lse / delta: null pointer");
  if (want_dq && !dq) return fail(FA2_E_NULL_POINTER, "dq: null pointer");
  if (lse_row_stride != round_up(Sq, 128)) return fail(FA2_E_BAD_SHAPE, "lse_row_stride must be ceil(Sq/128)*128");
  if (!(softmax_scale > 0.f)) return fail(FA2_E_BAD_SHAPE, "softmax_scale must be positive");
  if (deterministic < 0 || deterministic > 7) return fail(FA2_E_BAD_SHAPE, "deterministic must be one of FA2_DQ_* (0..7)");
  if (want_dq) {
    if (dq_dtype != dtype && dq_dtype != FA2_DTYPE_FP32)
      return fail(FA2_E_UNSUPPORTED_DTYPE, "dq_dtype %d: must be the input dtype (%d) or 32 (fp32)", dq_dtype, dtype);
    if (reinterpret_cast<uintptr_t>(dq) % 16) return fail(FA2_E_BAD_ALIGNMENT, "dq: base not 16-byte aligned");
    for (int i = 0; i < 3; ++i)
      if (dq_strides[i] % 8) return fail(FA2_E_BAD_ALIGNMENT, "dq: strides must be multiples of 8 elements");
  }
  const size_t need = fa2_sm100_bwd_workspace_bytes(B, Hq, Sq, Sk, D, causal, deterministic);
  if (need > 0) {
    if (!workspace) return fail(FA2_E_NULL_POINTER, "workspace: null pointer");
    if (workspace_bytes < need)
      return fail(FA2_E_WORKSPACE_TOO_SMALL, "workspace has %zu bytes, %zu needed", workspace_bytes, need);
    if (reinterpret_cast<uintptr_t>(workspace) % 16) return fail(FA2_E_BAD_ALIGNMENT, "workspace not 16-byte aligned");
  }
  CUtensorMap tq, tk, tv, tdo, tdk, tdv;
  if (int e = make_tmap_bshd(&tq, q, dtype, B, Sq, Hq, D, q_strides, 128, "q")) return e;
  if (int e = make_tmap_bshd(&tk, k, dtype, B, Sk, Hkv, D, k_strides, 128, "k")) return e;
  if (int e = make_tmap_bshd(&tv, v, dtype, B, Sk, Hkv, D, v_strides, 128, "v")) return e;
  if (int e = make_tmap_bshd(&tdo, dout, dtype, B, Sq, Hq, D, do_strides, 128, "dout")) return e;
  if (want_dkdv) {
    if (int e = make_tmap_bshd(&tdk, dk, dtype, B, Sk, Hkv, D, dk_strides, 128, "dk")) return e;
    if (int e = make_tmap_bshd(&tdv, dv, dtype, B, Sk, Hkv, D, dv_strides, 128, "dv")) return e;
  }

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (seqlens) {
    const VarlenTensor vt[6] = {{&tq, q, q_strides[0]}, {&tk, k, k_strides[0]}, {&tv, v, v_strides[0]},
                                {&tdo, dout, do_strides[0]}, {&tdk, dk, dk_strides[0]}, {&tdv, dv, dv_strides[0]}};
    if (int e = setup_varlen_maps(vt, 6, seqlens, tmap_ws, tmap_ws_bytes, B, Sq, st)) return e;
  }
  const CUtensorMap* vmaps = seqlens ? static_cast<const CUtensorMap*>(tmap_ws) : nullptr;
  const int sq_pad = round_up(Sq, 128);
  const bool fused = deterministic == FA2_DQ_ATOMIC || deterministic == FA2_DQ_FUSED_ORDERED;
  fa2::BwdParams p{};
  p.lse = lse; p.delta = delta;
  p.dq_accum = static_cast<float*>(workspace);
  p.dq_sem = fused ? reinterpret_cast<int*>(static_cast<char*>(workspace) + static_cast<size_t>(B) * Hq * sq_pad * D * sizeof(float))
                   : nullptr;
  p.B = B; p.Hq = Hq; p.Hkv = Hkv; p.Sq = Sq; p.Sk = Sk;
  p.lse_stride = static_cast<int>(lse_row_stride);
  p.sq_pad = sq_pad;
  p.softmax_scale = softmax_scale;
  p.scale_log2 = softmax_scale * kLog2e;
  p.causal = causal ? 1 : 0;
  p.deterministic = deterministic == FA2_DQ_FUSED_ORDERED ? 1 : 0;
  p.order_ctr = p.deterministic ? p.dq_sem + static_cast<size_t>(B) * Hq * (sq_pad / 128) : nullptr;
  cudaError_t ce = cudaSuccess;
  const bool spill = deterministic == FA2_DQ_DET_SPILL || deterministic == FA2_DQ_SPILL_DKDV || deterministic == FA2_DQ_SPILL_DQ;
  const long long tiles_per_head = fa2::ds_tile_prefix((Sq + 127) / 128, Sq, Sk, causal ? 1 : 0);
  void* spill_base = spill ? static_cast<char*>(workspace) + 1024 : nullptr;
  // Work-item counter of the persistent dQ kernel (byte 0 of the workspace; the kernel re-zeroes it when it is done, this
  // covers the first use of a buffer). When the dK/dV kernel runs first, its CTA (0, 0, 0) clears the word - no memset
  // node between the kernels, which would also break their programmatic dependent launch chain.
  const bool dkdv_clears_counter = !fused && want_dkdv && !seqlens;
  if (!fused && !dkdv_clears_counter) {
    ce = cudaMemsetAsync(workspace, 0, 128, st);
    if (ce != cudaSuccess) return cuda_fail(ce, "fa2_sm100_bwd scheduler counter memset");
  }
  int kv_split_used = 1;
  float* part_dk_used = nullptr;
  if (want_dkdv) {
    if (fused) {
      ce = fa2::launch_bwd(Dk, dtype == FA2_DTYPE_BF16, true, tq, tk, tv, tdo, tdk, tdv, p, st);
    } else {
      fa2::DkdvParams k{};
      k.lse = lse; k.delta = delta;
      k.ds_spill = spill_base;
      k.tiles_per_head = tiles_per_head;
      k.B = B; k.Hq = Hq; k.Hkv = Hkv; k.Sq = Sq; k.Sk = Sk;
      k.lse_stride = static_cast<int>(lse_row_stride);
      k.softmax_scale = softmax_scale;
      k.scale_log2 = softmax_scale * kLog2e;
      k.causal = causal ? 1 : 0;
      k.zero_word = dkdv_clears_counter ? static_cast<int*>(workspace) : nullptr;
      k.bias = bias_view;
      CUtensorMap tbias;
      if (bias && bias_dtype == dtype) {
        // [1 | B, 1 | Hq, Sq, Sk] seen as a [B', Sq, H', Sk] tensor: same map builder as q / k / v. Only if TMA can address
        // it (16-byte aligned base, strides multiples of 8 elements); the kernel's plain loads serve every other bias.
        const int Bb = bias_strides[0] ? B : 1, Hb = bias_strides[1] ? Hq : 1;
        const int64_t bs[3] = {bias_strides[0], bias_strides[1], bias_strides[2]};
        const bool ok = reinterpret_cast<uintptr_t>(bias) % 16 == 0 && bs[0] % 8 == 0 && bs[1] % 8 == 0 && bs[2] % 8 == 0 &&
                        bs[2] > 0;
        if (ok && make_tmap_bshd(&tbias, bias, dtype, Bb, Sq, Hb, Sk, bs, 128, "bias") == 0) {
          k.bias_tma = 1;
          k.bias_map = &tbias;
          k.bias_head_mul = Hb > 1 ? 1 : 0;
          k.bias_batch_mul = Bb > 1 ? 1 : 0;
        }
      }
      k.seqlens = seqlens;
      k.vmaps = vmaps;
      // large query-head groups: split over several CTAs when the caller's workspace has room for the fp32 partials
      // (sized with fa2_sm100_bwd_workspace_bytes_gqa); otherwise one CTA per (key block, kv head, batch) as before
      int kv_split = (bias || seqlens) ? 1 : kv_split_of(Hq, Hkv, D, deterministic);
      const size_t part_off = (need + 15) / 16 * 16;
      if (kv_split > 1 && workspace_bytes < part_off + kv_split_bytes(B, Hkv, Sk, D, kv_split)) kv_split = 1;
      k.kv_split = kv_split;
      if (kv_split > 1) {
        k.part_dk = reinterpret_cast<float*>(static_cast<char*>(workspace) + part_off);
        k.part_dv = k.part_dk + static_cast<size_t>(B) * Sk * Hkv * kv_split * D;
      }
      kv_split_used = kv_split;
      part_dk_used = k.part_dk;
      k.d_steps = (D + 15) / 16;
      k.d_cols = round_up(D, 16);
      ce = fa2::launch_bwd_dkdv(Dk, dtype == FA2_DTYPE_BF16, spill, tq, tk, tv, tdo, tdk, tdv, k, st);
    }
    if (ce != cudaSuccess) return cuda_fail(ce, "fa2_sm100_bwd launch");
    if (kv_split_used > 1) {
      fa2::DkdvReduceParams r{};
      r.part_dk = part_dk_used;
      r.part_dv = part_dk_used + static_cast<size_t>(B) * Sk * Hkv * kv_split_used * D;
      r.dk = dk; r.dv = dv;
      for (int i = 0; i < 3; ++i) { r.dk_strides[i] = dk_strides[i]; r.dv_strides[i] = dv_strides[i]; }
      r.B = B; r.Sk = Sk; r.Hkv = Hkv; r.D = D; r.kv_split = kv_split_used;
      ce = fa2::launch_dkdv_reduce(dtype == FA2_DTYPE_BF16, r, st);
      if (ce != cudaSuccess) return cuda_fail(ce, "fa2_dkdv_reduce_kernel launch");
    }
  }
  if (!want_dq) return 0;

  if (fused) {
    // dQ accumulator -> caller's tensor
    return fa2_sm100_dq_convert(workspace, dq, B, Hq, Sq, D, dq_strides, softmax_scale, dq_dtype, stream);
  }
  if (spill) {
    fa2::DqSpillParams d{};
    d.ds_spill = spill_base;
    d.tiles_per_head = tiles_per_head;
    d.sched = static_cast<int*>(workspace);
    d.dq = dq;
    for (int i = 0; i < 3; ++i) d.dq_strides[i] = dq_strides[i];
    d.dq_is_16bit = dq_dtype == dtype ? 1 : 0;
    d.B = B; d.Hq = Hq; d.Hkv = Hkv; d.Sq = Sq; d.Sk = Sk;
    d.softmax_scale = softmax_scale;
    d.causal = causal ? 1 : 0;
    d.d_actual = D;
    d.seqlens = seqlens;
    d.vmaps = vmaps;
    ce = fa2::launch_bwd_dq_spill(Dk, dtype == FA2_DTYPE_BF16, tk, d, st);
    if (ce != cudaSuccess) return cuda_fail(ce, "fa2_sm100_bwd dQ (spill) launch");
    return 0;
  }
  CUtensorMap tdq = tq;  // placeholder when dq is fp32 (the kernel then stores through plain pointers)
  if (dq_dtype == dtype)
    if (int e = make_tmap_bshd(&tdq, dq, dtype, B, Sq, Hq, D, dq_strides, 128, "dq")) return e;
  fa2::DqParams d{};
  d.lse = lse; d.delta = delta; d.dq = dq;
  d.sched = static_cast<int*>(workspace);
  for (int i = 0; i < 3; ++i) d.dq_strides[i] = dq_strides[i];
  d.dq_is_16bit = dq_dtype == dtype ? 1 : 0;
  d.B = B; d.Hq = Hq; d.Hkv = Hkv; d.Sq = Sq; d.Sk = Sk;
  d.lse_stride = static_cast<int>(lse_row_stride);
  d.softmax_scale = softmax_scale;
  d.scale_log2 = softmax_scale * kLog2e;
  d.causal = causal ? 1 : 0;
  d.bias = bias_view;
  d.d_actual = D;
  ce = fa2::launch_bwd_dq(Dk, dtype == FA2_DTYPE_BF16, tq, tk, tv, tdo, tdq, d, st);
  if (ce != cudaSuccess) return cuda_fail(ce, "fa2_sm100_bwd dQ launch");
  return 0;
}

}  // namespace

extern "C" {

int fa2_sm100_dq_convert(const void* workspace, void* dq, int B, int Hq, int Sq, int D, const int64_t dq_strides[3],
                         float softmax_scale, int out_dtype, void* stream) {
  if (!workspace || !dq) return fail(FA2_E_NULL_POINTER, "workspace / dq: null pointer");
  if (D != 64 && D != 128) return fail(FA2_E_UNSUPPORTED_HEAD_DIM, "head_dim %d: this build supports 64 and 128", D);
  if (B < 1 || Hq < 1 || Sq < 1) return fail(FA2_E_BAD_SHAPE, "non-positive dimension");
  if (out_dtype != FA2_DTYPE_FP16 && out_dtype != FA2_DTYPE_BF16 && out_dtype != FA2_DTYPE_FP32)
    return fail(FA2_E_UNSUPPORTED_DTYPE, "out_dtype code %d: expected 16, 17 or 32", out_dtype);
  if (reinterpret_cast<uintptr_t>(dq) % 16) return fail(FA2_E_BAD_ALIGNMENT, "dq: base not 16-byte aligned");
  for (int i = 0; i < 3; ++i)
    if (dq_strides[i] % 8) return fail(FA2_E_BAD_ALIGNMENT, "dq: strides must be multiples of 8 elements");
  fa2::DqConvertParams p{};
  p.dq_accum = static_cast<const float*>(workspace);
  p.dq = dq;
  p.B = B; p.Hq = Hq; p.Sq = Sq; p.D = D;
  p.sq_pad = round_up(Sq, 128);
  for (int i = 0; i < 3; ++i) p.dq_strides[i] = dq_strides[i];
  p.scale = softmax_scale;
  p.out_dtype = out_dtype;
  cudaError_t ce = fa2::launch_dq_convert(p, static_cast<cudaStream_t>(stream));
  if (ce != cudaSuccess) return cuda_fail(ce, "fa2_sm100_dq_convert launch");
  return 0;
}

int fa2_sm100_copy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes,
                           size_t rows, int to_device, void* stream) {
  if (!dst || !src) return fail(FA2_E_NULL_POINTER, "copy2d: null pointer");
  if (width_bytes > dst_pitch || width_bytes > src_pitch) return fail(FA2_E_BAD_SHAPE, "copy2d: row wider than its pitch");
  cudaError_t ce = cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, rows,
                                     to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                                     static_cast<cudaStream_t>(stream));
  if (ce != cudaSuccess) return cuda_fail(ce, "fa2_sm100_copy2d_async");
  return 0;
}

#ifdef FA2_TRACE
int fa2_sm100_debug_fetch_trace(int which, long long* host, size_t n) {
  switch (which) {
    case 0: return static_cast<int>(fa2::fetch_fwd_trace(host, n));
    case 1: return static_cast<int>(fa2::fetch_bwd_trace(host, n));
    case 2: return static_cast<int>(fa2::fetch_dq_trace(host, n));
    default: return static_cast<int>(fa2::fetch_dkdv_trace(host, n));
  }
}
#endif

}  // extern "C"
