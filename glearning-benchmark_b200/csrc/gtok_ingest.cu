// gtok_ingest.cu - device side of the compact host->device wire format.
//
// The reference keeps every graph as int64 PyG tensors on the host (edge_index int64[2,E], x int64[N,1],
// edge_attr int64[E]; graph_data_loader/zinc_dataset_autograph.py:51-73) and the engine's resident layout is
// int32 (include/gtok.h GtokGraphBatch).  Over PCIe that layout is the end-to-end bottleneck (0.68 GB per
// 1M molecules, 13 ms, vs 0.7 ms of tokenization), although local node ids, atom and bond types of molecule-
// sized graphs fit one byte and per-graph counts fit 8/16 bits.  The host side therefore ships the NARROWEST
// lossless unsigned type per array (gtok/batch.py PackedGraphBatch) and these kernels restore the int32 /
// int64-offset layout in HBM:
//   gtok_widen_i32          u8 / u16 / i32 -> int32, 16 bytes in, 16-byte stores out
//   gtok_counts_to_offsets  per-graph counts (u8 / u16 / i32) -> int64[G+1] exclusive offsets (widen + K0 scan)
// Pure streaming: (w + 4) bytes per element.
#include "gtok_common.cuh"

namespace gtok {

constexpr int kWidenThreads = 256;

template <typename InT, int PER>   // PER elements per thread = 16 bytes of input
__global__ void __launch_bounds__(kWidenThreads)
widen_kernel(const InT* __restrict__ in, int64_t n, int32_t* __restrict__ out) {
  const int64_t nvec = n / PER;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 raw = __ldcs(reinterpret_cast<const uint4*>(in) + i);
    const unsigned wds[4] = {raw.x, raw.y, raw.z, raw.w};
    int4* o = reinterpret_cast<int4*>(out + i * PER);
    if (sizeof(InT) == 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        __stcs(o + k, make_int4((int)(wds[k] & 0xffu), (int)((wds[k] >> 8) & 0xffu), (int)((wds[k] >> 16) & 0xffu),
                                (int)(wds[k] >> 24)));
    } else {
#pragma unroll
      for (int k = 0; k < 2; ++k)
        __stcs(o + k, make_int4((int)(wds[2 * k] & 0xffffu), (int)(wds[2 * k] >> 16), (int)(wds[2 * k + 1] & 0xffffu),
                                (int)(wds[2 * k + 1] >> 16)));
    }
  }
  // tail (fewer than PER elements)
  const int64_t t0 = nvec * PER;
  const int64_t j = t0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[j] = (int32_t)in[j];
}

static int widen_launch(const void* in, int elem_bytes, int64_t n, int32_t* out, cudaStream_t st) {
  if (n == 0) return GTOK_OK;
  if (elem_bytes == 4) {
    GTOK_CUDA_TRY(cudaMemcpyAsync(out, in, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    return GTOK_OK;
  }
  const int per = 16 / elem_bytes;
  int64_t blocks = (n / per + kWidenThreads - 1) / kWidenThreads;
  const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (elem_bytes == 1)
    widen_kernel<uint8_t, 16><<<(unsigned)blocks, kWidenThreads, 0, st>>>(static_cast<const uint8_t*>(in), n, out);
  else
    widen_kernel<uint16_t, 8><<<(unsigned)blocks, kWidenThreads, 0, st>>>(static_cast<const uint16_t*>(in), n, out);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

}  // namespace gtok

extern "C" int gtok_widen_i32(const void* in, int32_t elem_bytes, int64_t n, int32_t* out, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(n >= 0, GTOK_EINVAL, "widen_i32: negative length");
  GTOK_REQUIRE(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4, GTOK_EINVAL, "widen_i32: element size must be 1, 2 or 4");
  if (n == 0) return GTOK_OK;
  GTOK_REQUIRE(in && out, GTOK_EINVAL, "widen_i32: null pointer");
  GTOK_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, GTOK_EINVAL,
               "widen_i32: buffers must be 16-byte aligned");
  return widen_launch(in, elem_bytes, n, out, (cudaStream_t)stream);
}

extern "C" int gtok_counts_to_offsets(const void* counts, int32_t elem_bytes, int64_t num_items, int64_t* off_out,
                                      int32_t* tmp_i32, void* ws, size_t ws_bytes, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(num_items >= 0, GTOK_EINVAL, "counts_to_offsets: negative length");
  GTOK_REQUIRE(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4, GTOK_EINVAL, "counts_to_offsets: element size must be 1, 2 or 4");
  GTOK_REQUIRE(off_out != nullptr, GTOK_EINVAL, "counts_to_offsets: null off_out");
  if (num_items > 0) {
    GTOK_REQUIRE(counts && tmp_i32, GTOK_EINVAL, "counts_to_offsets: null pointer");
    GTOK_REQUIRE((reinterpret_cast<uintptr_t>(counts) & 15) == 0 && (reinterpret_cast<uintptr_t>(tmp_i32) & 15) == 0, GTOK_EINVAL,
                 "counts_to_offsets: buffers must be 16-byte aligned");
    const int rc = widen_launch(counts, elem_bytes, num_items, tmp_i32, (cudaStream_t)stream);
    if (rc != GTOK_OK) return rc;
  }
  return gtok_scan_offsets(tmp_i32, num_items, off_out, ws, ws_bytes, stream);
}
