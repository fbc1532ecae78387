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

// ---- the reference's OWN storage -> resident layouts ----------------------------------------------------------------
// PyG keeps a dataset as int64 tensors (InMemoryDataset: edge_index int64[2, E_total] with LOCAL node ids, x
// int64[N_total, 1], edge_attr int64[E_total], slices int64[G+1]; zinc_dataset_autograph.py:51-73).  The slices ARE
// node_off / edge_off; these kernels narrow the rest once the bytes are in HBM, so an end-to-end step that starts
// from the reference's tensors needs no host-side conversion at all:
//   gtok_pyg_pack_mol   -> one 16-bit word per column (src | dst << 6 | edge_attr << 12) + one byte per atom
//                          (GtokPackedMolBatch), 24 B in / 2 B out per column, 8 B in / 1 B out per atom
//   gtok_pyg_narrow     -> the int32 arrays of GtokGraphBatch (any graph size)
// Values that do not fit are reported through `status` (OR of GTOK_PYG_* bits), never silently truncated.
namespace gtok {

constexpr int kPygThreads = 256;

__device__ __forceinline__ void ld8_i64(const int64_t* __restrict__ p, long long (&v)[8]) {
  const longlong2* q = reinterpret_cast<const longlong2*>(p);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const longlong2 t = __ldcs(q + k);
    v[2 * k] = t.x;
    v[2 * k + 1] = t.y;
  }
}

template <bool HAS_ATTR>
__global__ void __launch_bounds__(kPygThreads)
pyg_pack_cols_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, const int64_t* __restrict__ attr,
                     int64_t n, uint16_t* __restrict__ col16, int32_t* __restrict__ status) {
  const int64_t nvec = n >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long bad_id = 0ull, bad_attr = 0ull;          // OR of every value's bits above the field width
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    long long a[8], c[8], t[8];
    ld8_i64(src + 8 * i, a);
    ld8_i64(dst + 8 * i, c);
    if (HAS_ATTR) ld8_i64(attr + 8 * i, t);
    unsigned w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      bad_id |= (unsigned long long)(a[k] | c[k]);
      unsigned v = ((unsigned)a[k] & 63u) | (((unsigned)c[k] & 63u) << 6);
      if (HAS_ATTR) {
        bad_attr |= (unsigned long long)t[k];
        v |= ((unsigned)t[k] & 15u) << 12;
      }
      w[k] = v;
    }
    __stcs(reinterpret_cast<uint4*>(col16) + i,
           make_uint4(w[0] | (w[1] << 16), w[2] | (w[3] << 16), w[4] | (w[5] << 16), w[6] | (w[7] << 16)));
  }
  const int64_t j = (nvec << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // tail: fewer than 8 columns
  if (j < n) {
    const long long a = src[j], c = dst[j], t = HAS_ATTR ? attr[j] : 0;
    bad_id |= (unsigned long long)(a | c);
    bad_attr |= (unsigned long long)t;
    col16[j] = (uint16_t)(((unsigned)a & 63u) | (((unsigned)c & 63u) << 6) | (((unsigned)t & 15u) << 12));
  }
  int flags = ((bad_id & ~63ull) ? GTOK_PYG_BAD_NODE_ID : 0) | ((bad_attr & ~15ull) ? GTOK_PYG_BAD_EDGE_ATTR : 0);
  flags = __reduce_or_sync(kFull, flags);
  if (flags && lane_id() == 0) atomicOr(status, flags);
}

__global__ void __launch_bounds__(kPygThreads)
pyg_pack_atoms_kernel(const int64_t* __restrict__ x, int64_t n, uint8_t* __restrict__ x8, int32_t* __restrict__ status) {
  const int64_t nvec = n >> 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long bad = 0ull;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    long long a[8], b[8];
    ld8_i64(x + 16 * i, a);
    ld8_i64(x + 16 * i + 8, b);
    unsigned w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      bad |= (unsigned long long)(a[k] | b[k]);
      w[k >> 2] |= ((unsigned)a[k] & 255u) << (8 * (k & 3));
      w[2 + (k >> 2)] |= ((unsigned)b[k] & 255u) << (8 * (k & 3));
    }
    __stcs(reinterpret_cast<uint4*>(x8) + i, make_uint4(w[0], w[1], w[2], w[3]));
  }
  const int64_t j = (nvec << 4) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) {
    const long long a = x[j];
    bad |= (unsigned long long)a;
    x8[j] = (uint8_t)a;
  }
  int flags = (bad & ~255ull) ? GTOK_PYG_BAD_NODE_TYPE : 0;
  flags = __reduce_or_sync(kFull, flags);
  if (flags && lane_id() == 0) atomicOr(status, flags);
}

// int64 -> int32, 4 per thread (32 B in, 16 B out)
__global__ void __launch_bounds__(kPygThreads)
pyg_narrow_kernel(const int64_t* __restrict__ in, int64_t n, int32_t* __restrict__ out, int32_t* __restrict__ status) {
  const int64_t nvec = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  auto fits = [](long long v) { return v == (long long)(int)v; };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const longlong2 p = __ldcs(reinterpret_cast<const longlong2*>(in) + 2 * i);
    const longlong2 q = __ldcs(reinterpret_cast<const longlong2*>(in) + 2 * i + 1);
    bad |= !(fits(p.x) && fits(p.y) && fits(q.x) && fits(q.y));
    __stcs(reinterpret_cast<int4*>(out) + i, make_int4((int)p.x, (int)p.y, (int)q.x, (int)q.y));
  }
  const int64_t j = (nvec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) {
    const long long v = in[j];
    bad |= !fits(v);
    out[j] = (int)v;
  }
  const int flags = __reduce_or_sync(kFull, bad ? GTOK_PYG_BAD_RANGE : 0);
  if (flags && lane_id() == 0) atomicOr(status, flags);
}

static unsigned pyg_grid(int64_t nvec) {
  int64_t blocks = (nvec + kPygThreads - 1) / kPygThreads;
  const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 8;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace gtok

extern "C" int gtok_pyg_pack_mol(const int64_t* src64, const int64_t* dst64, const int64_t* attr64, int64_t num_entries,
                                 const int64_t* x64, int64_t num_nodes, uint16_t* col16, uint8_t* x8, int32_t* status,
                                 void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(num_entries >= 0 && num_nodes >= 0, GTOK_EINVAL, "pyg_pack_mol: negative length");
  GTOK_REQUIRE(status != nullptr, GTOK_EINVAL, "pyg_pack_mol: null status");
  cudaStream_t st = (cudaStream_t)stream;
  if (num_entries > 0) {
    GTOK_REQUIRE(src64 && dst64 && col16, GTOK_EINVAL, "pyg_pack_mol: null edge_index / col16");
    GTOK_REQUIRE(al16(src64) && al16(dst64) && al16(attr64) && al16(col16), GTOK_EINVAL,
                 "pyg_pack_mol: buffers must be 16-byte aligned");
    if (attr64)
      pyg_pack_cols_kernel<true><<<pyg_grid(num_entries >> 3), kPygThreads, 0, st>>>(src64, dst64, attr64, num_entries, col16, status);
    else
      pyg_pack_cols_kernel<false><<<pyg_grid(num_entries >> 3), kPygThreads, 0, st>>>(src64, dst64, nullptr, num_entries, col16, status);
    GTOK_CUDA_TRY(cudaGetLastError());
  }
  if (num_nodes > 0) {
    GTOK_REQUIRE(x64 && x8, GTOK_EINVAL, "pyg_pack_mol: null x / x8");
    GTOK_REQUIRE(al16(x64) && al16(x8), GTOK_EINVAL, "pyg_pack_mol: buffers must be 16-byte aligned");
    pyg_pack_atoms_kernel<<<pyg_grid(num_nodes >> 4), kPygThreads, 0, st>>>(x64, num_nodes, x8, status);
    GTOK_CUDA_TRY(cudaGetLastError());
  }
  return GTOK_OK;
}

extern "C" int gtok_pyg_narrow(const int64_t* in64, int64_t n, int32_t* out, int32_t* status, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(n >= 0, GTOK_EINVAL, "pyg_narrow: negative length");
  GTOK_REQUIRE(status != nullptr, GTOK_EINVAL, "pyg_narrow: null status");
  if (n == 0) return GTOK_OK;
  GTOK_REQUIRE(in64 && out, GTOK_EINVAL, "pyg_narrow: null pointer");
  GTOK_REQUIRE(al16(in64) && al16(out), GTOK_EINVAL, "pyg_narrow: buffers must be 16-byte aligned");
  pyg_narrow_kernel<<<pyg_grid(n >> 2), kPygThreads, 0, (cudaStream_t)stream>>>(in64, n, out, status);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}
