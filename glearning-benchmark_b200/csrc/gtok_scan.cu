// gtok_scan.cu - K0: single-pass exclusive scan of per-graph token counts into
// ragged offsets (int32 lengths -> int64 offsets, G+1 entries).
//
// Replaces: the implicit offsets of the reference's Python list building
// (self.seqs.append(...), data_loader.py:483).  Hand-written chained scan with
// decoupled look-back: one launch, each tile publishes {flag|value} in ONE 64-bit
// word so no fence ordering between flag and value is needed; tile ids come from
// an atomic ticket so a tile's predecessors are always already resident.
// Traffic: 4 B read + 8 B written per graph - noise next to the token stream.
#include "gtok_common.cuh"

namespace gtok {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <typename OutT>
__global__ void __launch_bounds__(kScanThreads)
scan_offsets_kernel(const int32_t* __restrict__ len, int64_t n, OutT* __restrict__ off,
                    unsigned int* __restrict__ ticket, unsigned long long* __restrict__ status) {
  __shared__ unsigned int s_tile;
  __shared__ long long s_warp[kScanThreads / 32];
  __shared__ long long s_excl;
  const int t = threadIdx.x;
  if (t == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const int64_t tile = s_tile;
  const int64_t base = tile * kScanTile + (int64_t)t * kScanItems;

  int v[kScanItems];
  if (base + kScanItems <= n) {
    const int4 a = __ldg(reinterpret_cast<const int4*>(len + base));
    const int4 b = __ldg(reinterpret_cast<const int4*>(len + base + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) v[i] = (base + i < n) ? __ldg(len + base + i) : 0;
  }
  long long tsum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) tsum += v[i];

  // block exclusive scan of the per-thread sums
  long long incl = tsum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    long long o = __shfl_up_sync(kFull, incl, d);
    if ((t & 31) >= d) incl += o;
  }
  if ((t & 31) == 31) s_warp[t >> 5] = incl;
  __syncthreads();
  long long wbase = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    const long long s = s_warp[w];
    if (w < (t >> 5)) wbase += s;
    total += s;
  }
  const long long texcl = wbase + incl - tsum;

  // publish the tile aggregate, then look back (warp 0)
  if (t < 32) {
    const long long excl = lookback_exclusive(status, tile, total);
    if (t == 0) s_excl = excl;
  }
  __syncthreads();
  long long run = s_excl + texcl;
  if (base + kScanItems <= n) {
    OutT o[kScanItems];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) { o[i] = (OutT)run; run += v[i]; }
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) off[base + i] = o[i];
    if (base + kScanItems == n) off[n] = (OutT)run;
  } else {
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      if (base + i < n) { off[base + i] = (OutT)run; run += v[i]; }
      if (base + i + 1 == n) off[n] = (OutT)run;
    }
  }
}

static size_t scan_ws_bytes(int64_t n) {
  const int64_t tiles = (n + kScanTile - 1) / kScanTile;
  return 16 + (size_t)(tiles > 0 ? tiles : 1) * 8;
}

template <typename OutT>
static int scan_launch(const int32_t* len, int64_t n, OutT* off, void* ws, size_t ws_bytes, cudaStream_t st) {
  GTOK_REQUIRE(n >= 0, GTOK_EINVAL, "scan: negative item count");
  GTOK_REQUIRE(off != nullptr, GTOK_EINVAL, "scan: null output");
  if (n == 0) {
    GTOK_CUDA_TRY(cudaMemsetAsync(off, 0, sizeof(OutT), st));
    return GTOK_OK;
  }
  GTOK_REQUIRE(len != nullptr && ws != nullptr, GTOK_EINVAL, "scan: null input/workspace");
  const size_t need = scan_ws_bytes(n);
  GTOK_REQUIRE(ws_bytes >= need, GTOK_EWORKSPACE, "scan: workspace %zu < %zu bytes", ws_bytes, need);
  GTOK_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0 && (reinterpret_cast<uintptr_t>(len) & 15) == 0,
               GTOK_EINVAL, "scan: len and workspace must be 16-byte aligned");
  GTOK_CUDA_TRY(cudaMemsetAsync(ws, 0, need, st));
  const int64_t tiles = (n + kScanTile - 1) / kScanTile;
  scan_offsets_kernel<OutT><<<(unsigned)tiles, kScanThreads, 0, st>>>(
      len, n, off, reinterpret_cast<unsigned int*>(ws),
      reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ws) + 16));
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

}  // namespace gtok

extern "C" size_t gtok_scan_workspace_bytes(int64_t num_items) { return gtok::scan_ws_bytes(num_items < 0 ? 0 : num_items); }

extern "C" int gtok_scan_offsets(const int32_t* len, int64_t num_items, int64_t* off_out, void* ws,
                                 size_t ws_bytes, void* stream) {
  return gtok::scan_launch<int64_t>(len, num_items, off_out, ws, ws_bytes, (cudaStream_t)stream);
}
