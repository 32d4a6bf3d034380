// Device helpers of the "extras" kernel variants: additive attention bias and forward dropout
// (reference src/forward/compute_row_blocks.py:48-54, 68-69, 76-79). Kept out of the default instantiations: the
// BASELINE configurations use neither, and their kernels must not pay a register or an instruction for them.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace fa2 {

// Additive bias, broadcastable to [B, Hq, Sq, Sk]: element (b, h, i, j) lives at
// base + b * sb + h * sh + i * sm + j (elements; sb / sh are 0 for broadcast dimensions, the last one is contiguous).
struct BiasView {
  const void* ptr;   // nullptr: no bias
  int is_fp32;       // 0: same 16-bit type as q, 1: fp32
  int64_t sb, sh, sm;
};

template <bool kBf16>
__device__ __forceinline__ float bias_to_float(uint16_t raw) {
  if constexpr (kBf16) return __uint_as_float(static_cast<uint32_t>(raw) << 16);
  else return __half2float(__ushort_as_half(raw));
}

// one element, `idx` counted in elements from the view's base
template <bool kBf16>
__device__ __forceinline__ float bias_load(const BiasView& b, int64_t idx) {
  if (b.is_fp32) return __ldg(static_cast<const float*>(b.ptr) + idx);
  return bias_to_float<kBf16>(__ldg(static_cast<const uint16_t*>(b.ptr) + idx));
}

// The reference's dropout draw (src/forward/compute_row_blocks.py:76-79): keep iff tl.rand(seed, offset) > p, where
// tl.rand is Triton's Philox-4x32-10 on counter (offset, 0, 0, 0) and key (seed low word, seed high word), first output
// word x reinterpreted as int32 and mapped to [0, 1) as (x >= 0 ? x : -x - 1) * 4.6566127342e-10 in fp32
// (triton/language/random.py; restated and pinned bit for bit in oracle/philox.py + tests/golden/triton_dropout_*.npz).
__device__ __forceinline__ bool dropout_keep(uint32_t seed_lo, uint32_t seed_hi, uint32_t offset, float p) {
  uint32_t c0 = offset, c1 = 0u, c2 = 0u, c3 = 0u, k0 = seed_lo, k1 = seed_hi;
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint32_t a_hi = __umulhi(0xD2511F53u, c0), a_lo = 0xD2511F53u * c0;
    const uint32_t b_hi = __umulhi(0xCD9E8D57u, c2), b_lo = 0xCD9E8D57u * c2;
    c0 = b_hi ^ c1 ^ k0;
    c2 = a_hi ^ c3 ^ k1;
    c1 = b_lo;
    c3 = a_lo;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  int x = static_cast<int>(c0);
  x = x < 0 ? -x - 1 : x;
  return __fmul_rn(__int2float_rn(x), 4.6566127342e-10f) > p;
}

}  // namespace fa2
