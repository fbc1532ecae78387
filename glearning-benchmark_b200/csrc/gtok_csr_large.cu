// gtok_csr_large.cu - K7 for graphs of ANY size (up to the 32767 nodes the SENT kernels accept): canonical undirected
// CSR from the caller's edge_index columns with library kernels only (gtok_csr.cu covers <= 256 nodes per graph through
// adjacency bit rows; beyond that a bit row per node would be N^2 / 8 bytes per graph).
//
// Same canonical form (oracle/sent_oracle.py canonical_csr; AutoGraph's to_undirected + coalesce, call site
// trainer/train_agtt.py:250): no self loops, every undirected edge once per direction, rows sorted by neighbour, label
// of {u, v} = edge_attr of its FIRST column in caller order, either direction.
//   1. count     per column: degree upper bounds (repeats counted) by atomicAdd            -> caller scans (K0)
//   2. scatter   per column: key = neighbour << 32 | column index into both endpoint rows (atomic cursor per row)
//   3. sort      per row: ascending keys, then keep the first key of every neighbour (smallest column index = first
//                column of the pair): rows of <= 32 entries in registers (warp bitonic network), <= 4096 in shared
//                memory (CTA bitonic), longer rows through a per-CTA table indexed by neighbour (atomicMin of the
//                column index, then an ordered sweep)                                       -> unique degrees, caller scans
//   4. emit      per row: col / eattr / int32 row pointers from the compacted keys
#include "gtok_common.cuh"

namespace gtok {
namespace {

constexpr int kLgThreads = 256;
constexpr int kLgSmemRow = 4096;                  // entries sorted in shared memory by one CTA
constexpr unsigned long long kInf = ~0ull;

__global__ void __launch_bounds__(kLgThreads)
lg_count_kernel(GtokGraphBatch b, int32_t* __restrict__ deg) {
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = warp0; g < b.num_graphs; g += nwarps) {
    const int64_t n0 = __ldg(b.node_off + g), e0 = __ldg(b.edge_off + g);
    const int N = (int)(__ldg(b.node_off + g + 1) - n0);
    const int64_t E = __ldg(b.edge_off + g + 1) - e0;
    for (int64_t e = lane; e < E; e += kWarp) {
      const int u = __ldcs(b.src + e0 + e), v = __ldcs(b.dst + e0 + e);
      if (u == v || (unsigned)u >= (unsigned)N || (unsigned)v >= (unsigned)N) continue;
      atomicAdd(deg + n0 + u, 1);
      atomicAdd(deg + n0 + v, 1);
    }
  }
}

__global__ void __launch_bounds__(kLgThreads)
lg_scatter_kernel(GtokGraphBatch b, const int64_t* __restrict__ rowptr_ub, int32_t* __restrict__ cursor,
                  unsigned long long* __restrict__ entries) {
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = warp0; g < b.num_graphs; g += nwarps) {
    const int64_t n0 = __ldg(b.node_off + g), e0 = __ldg(b.edge_off + g);
    const int N = (int)(__ldg(b.node_off + g + 1) - n0);
    const int64_t E = __ldg(b.edge_off + g + 1) - e0;
    for (int64_t e = lane; e < E; e += kWarp) {
      const int u = __ldcs(b.src + e0 + e), v = __ldcs(b.dst + e0 + e);
      if (u == v || (unsigned)u >= (unsigned)N || (unsigned)v >= (unsigned)N) continue;
      const int su = atomicAdd(cursor + n0 + u, 1), sv = atomicAdd(cursor + n0 + v, 1);
      entries[__ldg(rowptr_ub + n0 + u) + su] = ((unsigned long long)(unsigned)v << 32) | (unsigned long long)(unsigned)e;
      entries[__ldg(rowptr_ub + n0 + v) + sv] = ((unsigned long long)(unsigned)u << 32) | (unsigned long long)(unsigned)e;
    }
  }
}

// rows of <= 32 entries: one warp, key per lane, bitonic network over shuffles, first key of every neighbour kept
__global__ void __launch_bounds__(kLgThreads)
lg_sort_small_kernel(int64_t n_total, const int64_t* __restrict__ rowptr_ub, unsigned long long* __restrict__ entries,
                     int32_t* __restrict__ deg_out) {
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n_total; r += nwarps) {
    const int64_t p0 = __ldg(rowptr_ub + r);
    const int len = (int)(__ldg(rowptr_ub + r + 1) - p0);
    if (len > 32) continue;                               // (lg_sort_big_kernel)
    if (len == 0) {
      if (lane == 0) deg_out[r] = 0;
      continue;
    }
    unsigned long long key = lane < len ? entries[p0 + lane] : kInf;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        const unsigned long long other = __shfl_xor_sync(kFull, key, j);
        const bool up = (lane & k) == 0, lower = (lane & j) == 0;
        const bool take_min = up == lower;
        key = take_min ? (key < other ? key : other) : (key > other ? key : other);
      }
    }
    const unsigned long long prev = __shfl_up_sync(kFull, key, 1);
    const bool uniq = lane < len && (lane == 0 || (key >> 32) != (prev >> 32));
    const unsigned m = __ballot_sync(kFull, uniq);
    if (uniq) entries[p0 + __popc(m & ((1u << lane) - 1u))] = key;
    if (lane == 0) deg_out[r] = __popc(m);
  }
}

// longer rows: one CTA per row
__global__ void __launch_bounds__(kLgThreads)
lg_sort_big_kernel(int64_t n_total, const int64_t* __restrict__ node_row_graph_n /*unused*/, const int64_t* __restrict__ rowptr_ub,
                   unsigned long long* __restrict__ entries, int32_t* __restrict__ deg_out, int max_nodes,
                   unsigned int* __restrict__ table /*[gridDim.x * max_nodes]*/) {
  __shared__ unsigned long long s_key[kLgSmemRow];
  __shared__ int s_scan[kLgThreads];
  const int tid = threadIdx.x;
  for (int64_t r = blockIdx.x; r < n_total; r += gridDim.x) {
    const int64_t p0 = __ldg(rowptr_ub + r);
    const int64_t len64 = __ldg(rowptr_ub + r + 1) - p0;
    if (len64 <= 32) continue;
    unsigned long long* __restrict__ row = entries + p0;
    if (len64 <= kLgSmemRow) {
      const int len = (int)len64;
      int pow2 = 64;
      while (pow2 < len) pow2 <<= 1;
      for (int i = tid; i < pow2; i += kLgThreads) s_key[i] = i < len ? row[i] : kInf;
      __syncthreads();
      for (int k = 2; k <= pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = tid; i < (pow2 >> 1); i += kLgThreads) {
            const int lo = ((i / j) * j * 2) + (i % j), hi = lo + j;
            const bool up = (lo & k) == 0;
            const unsigned long long a = s_key[lo], c = s_key[hi];
            if ((a > c) == up) { s_key[lo] = c; s_key[hi] = a; }
          }
          __syncthreads();
        }
      }
      // keep the first key of every neighbour: block scan of the flags, chunk by chunk
      int base = 0;
      for (int c0 = 0; c0 < len; c0 += kLgThreads) {
        const int i = c0 + tid;
        const bool uniq = i < len && (i == 0 || (s_key[i] >> 32) != (s_key[i - 1] >> 32));
        s_scan[tid] = uniq ? 1 : 0;
        __syncthreads();
        for (int d = 1; d < kLgThreads; d <<= 1) {
          const int v = tid >= d ? s_scan[tid - d] : 0;
          __syncthreads();
          s_scan[tid] += v;
          __syncthreads();
        }
        if (uniq) row[base + s_scan[tid] - 1] = s_key[i];
        base += s_scan[kLgThreads - 1];
        __syncthreads();
      }
      if (tid == 0) deg_out[r] = base;
    } else {
      // table indexed by neighbour: smallest column index per neighbour, then an ordered sweep
      unsigned int* __restrict__ tb = table + (size_t)blockIdx.x * max_nodes;
      for (int i = tid; i < max_nodes; i += kLgThreads) tb[i] = 0xffffffffu;
      __syncthreads();
      for (int64_t i = tid; i < len64; i += kLgThreads) {
        const unsigned long long k = row[i];
        atomicMin(tb + (unsigned)(k >> 32), (unsigned)(k & 0xffffffffu));
      }
      __syncthreads();
      int base = 0;
      for (int c0 = 0; c0 < max_nodes; c0 += kLgThreads) {
        const int i = c0 + tid;
        const unsigned e = i < max_nodes ? tb[i] : 0xffffffffu;
        const bool has = e != 0xffffffffu;
        s_scan[tid] = has ? 1 : 0;
        __syncthreads();
        for (int d = 1; d < kLgThreads; d <<= 1) {
          const int v = tid >= d ? s_scan[tid - d] : 0;
          __syncthreads();
          s_scan[tid] += v;
          __syncthreads();
        }
        if (has) row[base + s_scan[tid] - 1] = ((unsigned long long)(unsigned)i << 32) | e;
        base += s_scan[kLgThreads - 1];
        __syncthreads();
      }
      if (tid == 0) deg_out[r] = base;
    }
    __syncthreads();
  }
  (void)node_row_graph_n;
}

__global__ void __launch_bounds__(kLgThreads)
lg_emit_kernel(GtokGraphBatch b, int64_t n_total, const int64_t* __restrict__ rowptr_ub,
               const unsigned long long* __restrict__ entries, const int64_t* __restrict__ rowptr64,
               int32_t* __restrict__ rowptr_out, int32_t* __restrict__ col, int32_t* __restrict__ eattr) {
  // one warp per graph, lanes over its rows
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  if (warp0 == 0 && lane == 0) rowptr_out[n_total] = (int32_t)__ldg(rowptr64 + n_total);
  for (int64_t g = warp0; g < b.num_graphs; g += nwarps) {
    const int64_t n0 = __ldg(b.node_off + g), e0 = __ldg(b.edge_off + g);
    const int N = (int)(__ldg(b.node_off + g + 1) - n0);
    for (int u = lane; u < N; u += kWarp) {
      const int64_t q0 = __ldg(rowptr64 + n0 + u), q1 = __ldg(rowptr64 + n0 + u + 1);
      const unsigned long long* __restrict__ src = entries + __ldg(rowptr_ub + n0 + u);
      rowptr_out[n0 + u] = (int32_t)q0;
      for (int64_t q = q0; q < q1; ++q) {
        const unsigned long long k = src[q - q0];
        col[q] = (int32_t)(k >> 32);
        if (eattr) eattr[q] = __ldg(b.edge_attr + e0 + (int64_t)(k & 0xffffffffu));
      }
    }
  }
}

unsigned lg_grid(int64_t threads) {
  int64_t blocks = (threads + kLgThreads - 1) / kLgThreads;
  const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 8;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

int lg_check(const GtokGraphBatch* b, const char* who) {
  GTOK_REQUIRE(b != nullptr, GTOK_EINVAL, "%s: null batch", who);
  GTOK_REQUIRE(b->num_graphs >= 0, GTOK_EINVAL, "%s: negative graph count", who);
  GTOK_REQUIRE(b->num_graphs == 0 || (b->node_off && b->edge_off), GTOK_EINVAL, "%s: null offsets", who);
  return GTOK_OK;
}

}  // namespace
}  // namespace gtok

extern "C" int gtok_csrl_count(const GtokGraphBatch* batch, int64_t total_nodes, int32_t* degree_ub, void* stream) {
  using namespace gtok;
  if (int rc = lg_check(batch, "csrl_count")) return rc;
  GTOK_REQUIRE(total_nodes >= 0, GTOK_EINVAL, "csrl_count: negative node count");
  if (total_nodes == 0) return GTOK_OK;
  GTOK_REQUIRE(degree_ub != nullptr, GTOK_EINVAL, "csrl_count: null output");
  cudaStream_t st = (cudaStream_t)stream;
  GTOK_CUDA_TRY(cudaMemsetAsync(degree_ub, 0, (size_t)total_nodes * 4, st));
  if (batch->num_graphs > 0 && batch->src && batch->dst) {
    lg_count_kernel<<<lg_grid(batch->num_graphs * 32), kLgThreads, 0, st>>>(*batch, degree_ub);
    GTOK_CUDA_TRY(cudaGetLastError());
  }
  return GTOK_OK;
}

extern "C" size_t gtok_csrl_table_bytes(int32_t max_nodes) {
  const int64_t ctas = (int64_t)(gtok::sm_count() > 0 ? gtok::sm_count() : 148) * 2;
  return (size_t)ctas * (size_t)(max_nodes > 0 ? max_nodes : 1) * 4;
}

extern "C" int gtok_csrl_sort(const GtokGraphBatch* batch, int64_t total_nodes, const int64_t* rowptr_ub, int32_t* cursor_ws,
                              uint64_t* entries, int32_t* degree_out, void* table_ws, size_t table_bytes, void* stream) {
  using namespace gtok;
  if (int rc = lg_check(batch, "csrl_sort")) return rc;
  GTOK_REQUIRE(total_nodes >= 0, GTOK_EINVAL, "csrl_sort: negative node count");
  if (total_nodes == 0) return GTOK_OK;
  GTOK_REQUIRE(rowptr_ub && cursor_ws && degree_out, GTOK_EINVAL, "csrl_sort: null pointer");
  GTOK_REQUIRE(batch->max_nodes > 0, GTOK_EINVAL, "csrl_sort: batch.max_nodes bound required");
  GTOK_REQUIRE(table_bytes >= gtok_csrl_table_bytes(batch->max_nodes) && table_ws, GTOK_EWORKSPACE,
               "csrl_sort: table workspace %zu < %zu bytes", table_bytes, gtok_csrl_table_bytes(batch->max_nodes));
  cudaStream_t st = (cudaStream_t)stream;
  GTOK_CUDA_TRY(cudaMemsetAsync(cursor_ws, 0, (size_t)total_nodes * 4, st));
  auto* ent = reinterpret_cast<unsigned long long*>(entries);
  if (batch->num_graphs > 0 && batch->src && batch->dst && entries) {
    lg_scatter_kernel<<<lg_grid(batch->num_graphs * 32), kLgThreads, 0, st>>>(*batch, rowptr_ub, cursor_ws, ent);
    GTOK_CUDA_TRY(cudaGetLastError());
  }
  lg_sort_small_kernel<<<lg_grid(total_nodes * 32), kLgThreads, 0, st>>>(total_nodes, rowptr_ub, ent, degree_out);
  GTOK_CUDA_TRY(cudaGetLastError());
  const int ctas = (sm_count() > 0 ? sm_count() : 148) * 2;
  lg_sort_big_kernel<<<ctas, kLgThreads, 0, st>>>(total_nodes, nullptr, rowptr_ub, ent, degree_out, batch->max_nodes,
                                                  static_cast<unsigned int*>(table_ws));
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

extern "C" int gtok_csrl_emit(const GtokGraphBatch* batch, int64_t total_nodes, const int64_t* rowptr_ub, const uint64_t* entries,
                              const int64_t* rowptr64, int64_t nnz, int32_t* rowptr_out, int32_t* col_out, int32_t* eattr_out,
                              void* stream) {
  using namespace gtok;
  if (int rc = lg_check(batch, "csrl_emit")) return rc;
  GTOK_REQUIRE(total_nodes >= 0 && nnz >= 0 && nnz < (1ll << 31), GTOK_EINVAL, "csrl_emit: sizes out of range (nnz must stay below 2^31)");
  GTOK_REQUIRE(rowptr64 && rowptr_out, GTOK_EINVAL, "csrl_emit: null row pointers");
  GTOK_REQUIRE(nnz == 0 || (col_out && entries && rowptr_ub), GTOK_EINVAL, "csrl_emit: null arrays");
  GTOK_REQUIRE(!eattr_out || batch->edge_attr || nnz == 0, GTOK_EINVAL, "csrl_emit: eattr_out without batch.edge_attr");
  cudaStream_t st = (cudaStream_t)stream;
  if (batch->num_graphs == 0 || total_nodes == 0) {
    GTOK_CUDA_TRY(cudaMemsetAsync(rowptr_out, 0, 4, st));
    return GTOK_OK;
  }
  lg_emit_kernel<<<lg_grid(batch->num_graphs * 32), kLgThreads, 0, st>>>(*batch, total_nodes, rowptr_ub,
                                                                         reinterpret_cast<const unsigned long long*>(entries), rowptr64,
                                                                         rowptr_out, col_out, eattr_out);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}
