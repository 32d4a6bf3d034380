// Per-sequence tensor maps for variable-length batches, built ON THE DEVICE from the lengths in device memory - no host
// synchronisation, unlike the reference's pack / unpack path, which reads every length with .item()
// (src/utils.py:8-31, src/forward/caller.py:46-59, 119-120).
//
// For every batch element b and every tensor of the launch, the kernel below copies the launch's full tensor map
// (dimensions (D, S, H, B)) into shared memory, moves its base address to element b and shrinks its sequence dimension
// to that sequence's length with `tensormap.replace`, and publishes it to global memory with
// `tensormap.cp_fenceproxy` (release). The attention kernels pick the map of their batch element, acquire it with
// `fence.proxy.tensormap::generic.acquire` and address it with batch coordinate 0: TMA then zero-fills every row at or
// beyond the sequence's own end on load and clips it on store, exactly as it does for the ragged tail of a dense
// batch - whatever the padding rows hold (NaNs included) cannot reach a result.
#include "fa2_kernels.cuh"
#include "ptx.cuh"

namespace fa2 {

namespace {

__global__ void __launch_bounds__(32 * kVarlenMapsPerBatch)
fa2_varlen_maps_kernel(const __grid_constant__ CUtensorMap t0, const __grid_constant__ CUtensorMap t1,
                       const __grid_constant__ CUtensorMap t2, const __grid_constant__ CUtensorMap t3,
                       const __grid_constant__ CUtensorMap t4, const __grid_constant__ CUtensorMap t5,
                       const VarlenSetupParams p) {
  __shared__ __align__(128) uint8_t staged[kVarlenMapsPerBatch][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  if (warp >= p.n_maps) return;
  const CUtensorMap* src = warp == 0 ? &t0 : warp == 1 ? &t1 : warp == 2 ? &t2 : warp == 3 ? &t3 : warp == 4 ? &t4 : &t5;
  if (lane < 8) reinterpret_cast<uint4*>(staged[warp])[lane] = reinterpret_cast<const uint4*>(src)[lane];
  __syncwarp();
  if (lane == 0) {
    int n = p.seqlens[b];
    n = n < 0 ? 0 : (n > p.S ? p.S : n);
    const uint32_t rows = n > 0 ? static_cast<uint32_t>(n) : 1u;  // a dimension of 0 is not a valid tensor map; the
                                                                   // kernels never touch the map of an empty sequence
    const uint64_t addr = p.base[warp] + static_cast<uint64_t>(b) * p.batch_stride_bytes[warp];
    const uint32_t smem_map = smem_u32(staged[warp]);
    asm volatile("tensormap.replace.tile.global_address.shared::cta.b1024.b64 [%0], %1;" ::"r"(smem_map), "l"(addr) : "memory");
    asm volatile("tensormap.replace.tile.global_dim.shared::cta.b1024.b32 [%0], 1, %1;" ::"r"(smem_map), "r"(rows) : "memory");
    asm volatile("tensormap.replace.tile.global_dim.shared::cta.b1024.b32 [%0], 3, %1;" ::"r"(smem_map), "r"(1u) : "memory");
  }
  __syncwarp();
  CUtensorMap* dst = p.out + static_cast<size_t>(b) * kVarlenMapsPerBatch + warp;
  asm volatile("tensormap.cp_fenceproxy.global.shared::cta.tensormap::generic.release.gpu.sync.aligned [%0], [%1], 128;"
               ::"l"(dst), "r"(smem_u32(staged[warp])) : "memory");
}

}  // namespace

cudaError_t launch_varlen_maps(const CUtensorMap* templates, const VarlenSetupParams& p, int B, cudaStream_t stream) {
  if (p.n_maps < 1 || p.n_maps > 6) return cudaErrorInvalidValue;
  const CUtensorMap& z = templates[0];
  auto t = [&](int i) -> const CUtensorMap& { return i < p.n_maps ? templates[i] : z; };
  fa2_varlen_maps_kernel<<<B, 32 * kVarlenMapsPerBatch, 0, stream>>>(t(0), t(1), t(2), t(3), t(4), t(5), p);
  return cudaGetLastError();
}

}  // namespace fa2
