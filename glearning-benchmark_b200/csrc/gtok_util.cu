// gtok_util.cu - small library kernels for the plumbing around the tokenizers, so that no torch op (boolean-mask
// indexing, repeat_interleave, argsort) computes anything on the product path:
//   gtok_compact_rows     strided rows [G, stride] + lengths -> ragged stream (the AGTT shim serves rows from it;
//                         Graph2TrailTokenizer.__call__ per item, trainer/train_agtt.py:250)
//   gtok_segment_counts / gtok_gather_segments
//                         gather the graphs idx[0], idx[1], ... of a batch into a new batch (TokenDataset over a
//                         shuffled / subset view of a PyG dataset: ds._indices, data_loader.py:465-486)
//   gtok_balance_order    processing order for the thread-per-graph SENT kernels: graphs sorted by (nodes, CSR entries)
//                         inside windows of consecutive graphs, so the 32 lanes of a warp walk graphs of similar work
#include "gtok_common.cuh"

namespace gtok {

constexpr int kUtilThreads = 256;

__global__ void __launch_bounds__(kUtilThreads)
compact_rows_kernel(const int64_t* __restrict__ tok, int64_t stride, const int32_t* __restrict__ row_len,
                    const int64_t* __restrict__ off, int64_t rows, int64_t* __restrict__ out) {
  const int lane = lane_id();
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = w0; r < rows; r += nw) {
    const int len = __ldg(row_len + r);
    const int64_t* __restrict__ src = tok + r * stride;
    int64_t* __restrict__ dst = out + __ldg(off + r);
    for (int p = lane; p < len; p += kWarp) __stcs(dst + p, __ldcs(src + p));
  }
}

__global__ void __launch_bounds__(kUtilThreads)
segment_counts_kernel(const int64_t* __restrict__ off, const int64_t* __restrict__ idx, int64_t n, int64_t num_graphs,
                      int32_t* __restrict__ cnt, int32_t* __restrict__ status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t g = __ldg(idx + i);
  if (g < 0 || g >= num_graphs) {
    cnt[i] = 0;
    atomicOr(status, 1);
    return;
  }
  cnt[i] = (int32_t)(__ldg(off + g + 1) - __ldg(off + g));
}

__global__ void __launch_bounds__(kUtilThreads)
gather_segments_kernel(const int32_t* __restrict__ src, const int64_t* __restrict__ src_off, const int64_t* __restrict__ idx,
                       const int64_t* __restrict__ dst_off, int64_t n, int64_t num_graphs, int32_t* __restrict__ dst) {
  const int lane = lane_id();
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = w0; i < n; i += nw) {
    const int64_t g = __ldg(idx + i);
    if (g < 0 || g >= num_graphs) continue;
    const int64_t s0 = __ldg(src_off + g), d0 = __ldg(dst_off + i);
    const int len = (int)(__ldg(dst_off + i + 1) - d0);
    for (int p = lane; p < len; p += kWarp) dst[d0 + p] = __ldg(src + s0 + p);
  }
}

// rows + per-row tails -> new ragged stream: row r = (tok[off[r] .. off[r]+len[r]) ++ tail[r, :tail_len[r]])[:max_len].
// new_len is computed by append_lengths_kernel, scanned by the caller into new_off.
__global__ void __launch_bounds__(kUtilThreads)
append_lengths_kernel(const int32_t* __restrict__ len, const int32_t* __restrict__ tail_len, int64_t rows, int max_len,
                      int32_t* __restrict__ new_len) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) {
    const int l = __ldg(len + r) + __ldg(tail_len + r);
    new_len[r] = l < max_len ? l : max_len;
  }
}
__global__ void __launch_bounds__(kUtilThreads)
append_rows_kernel(const int64_t* __restrict__ tok, const int64_t* __restrict__ off, const int32_t* __restrict__ len,
                   const int64_t* __restrict__ tail, int tail_stride, const int64_t* __restrict__ new_off, int64_t rows,
                   int64_t* __restrict__ out) {
  const int lane = lane_id();
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = w0; r < rows; r += nw) {
    const int l = __ldg(len + r);
    const int64_t d0 = __ldg(new_off + r);
    const int nl = (int)(__ldg(new_off + r + 1) - d0);
    const int64_t* __restrict__ src = tok + __ldg(off + r);
    for (int p = lane; p < nl; p += kWarp) out[d0 + p] = p < l ? __ldg(src + p) : __ldg(tail + r * tail_stride + (p - l));
  }
}

// one CTA per window of W <= 4096 consecutive graphs: bitonic sort of (key << 32 | local index) in shared memory
constexpr int kBalanceThreads = 512;
__global__ void __launch_bounds__(kBalanceThreads)
balance_order_kernel(const int64_t* __restrict__ node_off, const int32_t* __restrict__ rowptr, int64_t G, int window,
                     int pow2, int32_t* __restrict__ order) {
  extern __shared__ unsigned long long s_key[];
  const int64_t g0 = (int64_t)blockIdx.x * window;
  const int cnt = (int)((G - g0 < window) ? (G - g0) : window);
  for (int i = threadIdx.x; i < pow2; i += blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < cnt) {
      const int64_t n0 = __ldg(node_off + g0 + i), n1 = __ldg(node_off + g0 + i + 1);
      const unsigned n = (unsigned)(n1 - n0);
      unsigned nnz = (unsigned)(__ldg(rowptr + n1) - __ldg(rowptr + n0));
      if (nnz > 0xFFFFFu) nnz = 0xFFFFFu;
      k = ((unsigned long long)((n << 20) | nnz) << 32) | (unsigned)i;     // index in the low word: a stable order
    }
    s_key[i] = k;
  }
  __syncthreads();
  for (int size = 2; size <= pow2; size <<= 1) {
    for (int st = size >> 1; st > 0; st >>= 1) {
      for (int i = threadIdx.x; i < (pow2 >> 1); i += blockDim.x) {
        const int lo = ((i / st) * st * 2) + (i % st), hi = lo + st;
        const bool up = ((lo & size) == 0);
        const unsigned long long a = s_key[lo], b = s_key[hi];
        if ((a > b) == up) { s_key[lo] = b; s_key[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) order[g0 + i] = (int32_t)(g0 + (unsigned)(s_key[i] & 0xffffffffu));
}

static unsigned util_grid(int64_t warps) {
  int64_t blocks = (warps * 32 + kUtilThreads - 1) / kUtilThreads;
  const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 8;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

}  // namespace gtok

extern "C" int gtok_compact_rows(const int64_t* tok, int64_t stride, const int32_t* row_len, const int64_t* off,
                                 int64_t num_rows, int64_t* out, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(num_rows >= 0 && stride >= 0, GTOK_EINVAL, "compact_rows: negative size");
  if (num_rows == 0) return GTOK_OK;
  GTOK_REQUIRE(tok && row_len && off && out, GTOK_EINVAL, "compact_rows: null pointer");
  compact_rows_kernel<<<util_grid(num_rows), kUtilThreads, 0, (cudaStream_t)stream>>>(tok, stride, row_len, off, num_rows, out);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

extern "C" int gtok_segment_counts(const int64_t* off, int64_t num_graphs, const int64_t* idx, int64_t num_idx,
                                   int32_t* cnt_out, int32_t* status, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(num_idx >= 0 && num_graphs >= 0, GTOK_EINVAL, "segment_counts: negative size");
  GTOK_REQUIRE(status != nullptr, GTOK_EINVAL, "segment_counts: null status");
  if (num_idx == 0) return GTOK_OK;
  GTOK_REQUIRE(off && idx && cnt_out, GTOK_EINVAL, "segment_counts: null pointer");
  segment_counts_kernel<<<(unsigned)((num_idx + kUtilThreads - 1) / kUtilThreads), kUtilThreads, 0, (cudaStream_t)stream>>>(
      off, idx, num_idx, num_graphs, cnt_out, status);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

extern "C" int gtok_gather_segments(const int32_t* src, const int64_t* src_off, int64_t num_graphs, const int64_t* idx,
                                    const int64_t* dst_off, int64_t num_idx, int32_t* dst, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(num_idx >= 0 && num_graphs >= 0, GTOK_EINVAL, "gather_segments: negative size");
  if (num_idx == 0) return GTOK_OK;
  GTOK_REQUIRE(src_off && idx && dst_off, GTOK_EINVAL, "gather_segments: null pointer");
  if (!src || !dst) return GTOK_OK;                       // (nothing to move: every selected segment is empty)
  gather_segments_kernel<<<util_grid(num_idx), kUtilThreads, 0, (cudaStream_t)stream>>>(src, src_off, idx, dst_off, num_idx,
                                                                                         num_graphs, dst);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

extern "C" int gtok_balance_order(const int64_t* node_off, const int32_t* rowptr, int64_t num_graphs, int32_t window,
                                  int32_t* order_out, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(num_graphs >= 0, GTOK_EINVAL, "balance_order: negative graph count");
  GTOK_REQUIRE(window >= 2 && window <= 4096, GTOK_EUNSUPPORTED, "balance_order: window must be in [2, 4096] (got %d)", window);
  if (num_graphs == 0) return GTOK_OK;
  GTOK_REQUIRE(node_off && rowptr && order_out, GTOK_EINVAL, "balance_order: null pointer");
  int pow2 = 2;
  while (pow2 < window) pow2 <<= 1;
  const int64_t blocks = (num_graphs + window - 1) / window;
  balance_order_kernel<<<(unsigned)blocks, kBalanceThreads, (size_t)pow2 * 8, (cudaStream_t)stream>>>(node_off, rowptr, num_graphs,
                                                                                                     window, pow2, order_out);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

extern "C" int gtok_ragged_append_lengths(const int32_t* row_len, const int32_t* tail_len, int64_t num_rows, int32_t max_len,
                                          int32_t* new_len, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(num_rows >= 0 && max_len >= 0, GTOK_EINVAL, "ragged_append_lengths: negative size");
  if (num_rows == 0) return GTOK_OK;
  GTOK_REQUIRE(row_len && tail_len && new_len, GTOK_EINVAL, "ragged_append_lengths: null pointer");
  append_lengths_kernel<<<(unsigned)((num_rows + kUtilThreads - 1) / kUtilThreads), kUtilThreads, 0, (cudaStream_t)stream>>>(
      row_len, tail_len, num_rows, max_len, new_len);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

extern "C" int gtok_ragged_append(const int64_t* tok, const int64_t* off, const int32_t* row_len, const int64_t* tail,
                                  int32_t tail_stride, const int64_t* new_off, int64_t num_rows, int64_t* out, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(num_rows >= 0 && tail_stride >= 0, GTOK_EINVAL, "ragged_append: negative size");
  if (num_rows == 0) return GTOK_OK;
  GTOK_REQUIRE(off && row_len && new_off && out, GTOK_EINVAL, "ragged_append: null pointer");
  append_rows_kernel<<<util_grid(num_rows), kUtilThreads, 0, (cudaStream_t)stream>>>(tok, off, row_len, tail, tail_stride, new_off,
                                                                                     num_rows, out);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}
