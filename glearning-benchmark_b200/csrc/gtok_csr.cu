// gtok_csr.cu - K7: canonical undirected CSR + adjacency bit rows from the caller's edge_index columns, for graphs of
// at most 256 nodes.
//
// What AutoGraph's tokenizer does per call on the host (to_undirected + coalesce of edge_index, reference call site
// trainer/train_agtt.py:250) happens here ONCE per dataset on the device; the SENT kernels then walk the result every
// epoch.  Canonical form (oracle/sent_oracle.py canonical_csr): no self loops, every undirected edge once per
// direction, rows sorted by neighbour id, and the label of {u, v} is the edge_attr of its FIRST column in caller
// order, whichever direction that column has.
//   1. adjacency_kernel  warp per graph, lane <-> column: atomicOr of both directions into 64-bit adjacency rows
//                        (adj_words words per node) - set semantics remove repeats for free;
//   2. degree_kernel     thread per node: popcount of its row -> degree (the caller scans it into row pointers, K0);
//   3. fill_kernel       thread per node: enumerate the set bits ascending -> neighbour list, int32 row pointers;
//   4. first_col_kernel  warp per graph, lane <-> column: slot of (u, v) = rowptr[u] + popcount of row u below v;
//                        atomicMin of the column index into the slots of both directions;
//   5. label_kernel      thread per CSR entry: eattr = edge_attr[first column].
// Bytes: 8 (+4 labels) per column read twice, 8 * adj_words + 8 per node, 4 (+8 labels) per CSR entry written.
#include "gtok_common.cuh"

namespace gtok {
namespace {

constexpr int kCsrThreads = 256;

__global__ void __launch_bounds__(kCsrThreads)
adjacency_kernel(GtokGraphBatch b, int W, unsigned long long* __restrict__ adj) {
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = warp0; g < b.num_graphs; g += nwarps) {
    const int64_t n0 = __ldg(b.node_off + g), e0 = __ldg(b.edge_off + g);
    const int N = (int)(__ldg(b.node_off + g + 1) - n0), E = (int)(__ldg(b.edge_off + g + 1) - e0);
    for (int e = lane; e < E; e += kWarp) {
      const int u = __ldcs(b.src + e0 + e), v = __ldcs(b.dst + e0 + e);
      if (u == v || (unsigned)u >= (unsigned)N || (unsigned)v >= (unsigned)N) continue;      // self loop / not a node
      atomicOr(adj + (n0 + u) * W + (v >> 6), 1ull << (v & 63));
      atomicOr(adj + (n0 + v) * W + (u >> 6), 1ull << (u & 63));
    }
  }
}

__global__ void __launch_bounds__(kCsrThreads)
degree_kernel(const unsigned long long* __restrict__ adj, int W, int64_t n_total, int32_t* __restrict__ degree) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += stride) {
    int d = 0;
    for (int k = 0; k < W; ++k) d += __popcll(__ldg(adj + i * W + k));
    degree[i] = d;
  }
}

__global__ void __launch_bounds__(kCsrThreads)
fill_kernel(const unsigned long long* __restrict__ adj, int W, int64_t n_total, const int64_t* __restrict__ rowptr64,
            int32_t* __restrict__ rowptr, int32_t* __restrict__ col) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_total; i += stride) {
    const int64_t p0 = __ldg(rowptr64 + i);
    rowptr[i] = (int32_t)p0;
    if (i == n_total) continue;
    int64_t p = p0;
    for (int k = 0; k < W; ++k) {
      unsigned long long w = __ldg(adj + i * W + k);
      while (w) {
        col[p++] = 64 * k + __ffsll((long long)w) - 1;
        w &= w - 1;
      }
    }
  }
}

// CSR slot of neighbour v in the row of global node `node`
__device__ __forceinline__ int64_t slot_of(const unsigned long long* __restrict__ adj, int W, const int64_t* __restrict__ rowptr64,
                                           int64_t node, int v) {
  int rank = 0;
  const int wv = v >> 6;
  for (int k = 0; k < wv; ++k) rank += __popcll(__ldg(adj + node * W + k));
  rank += __popcll(__ldg(adj + node * W + wv) & ((1ull << (v & 63)) - 1ull));
  return __ldg(rowptr64 + node) + rank;
}

__global__ void __launch_bounds__(kCsrThreads)
first_col_kernel(GtokGraphBatch b, int W, const unsigned long long* __restrict__ adj, const int64_t* __restrict__ rowptr64,
                 unsigned long long* __restrict__ first_col) {
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = warp0; g < b.num_graphs; g += nwarps) {
    const int64_t n0 = __ldg(b.node_off + g), e0 = __ldg(b.edge_off + g);
    const int N = (int)(__ldg(b.node_off + g + 1) - n0), E = (int)(__ldg(b.edge_off + g + 1) - e0);
    for (int e = lane; e < E; e += kWarp) {
      const int u = __ldcs(b.src + e0 + e), v = __ldcs(b.dst + e0 + e);
      if (u == v || (unsigned)u >= (unsigned)N || (unsigned)v >= (unsigned)N) continue;
      const unsigned long long ce = (unsigned long long)(e0 + e);
      atomicMin(first_col + slot_of(adj, W, rowptr64, n0 + u, v), ce);
      atomicMin(first_col + slot_of(adj, W, rowptr64, n0 + v, u), ce);
    }
  }
}

__global__ void __launch_bounds__(kCsrThreads)
label_kernel(const unsigned long long* __restrict__ first_col, const int32_t* __restrict__ edge_attr, int64_t nnz,
             int32_t* __restrict__ eattr) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += stride)
    eattr[p] = __ldg(edge_attr + __ldg(first_col + p));
}

int grid_for(int64_t items) {
  const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 8;
  int64_t blocks = (items + kCsrThreads - 1) / kCsrThreads;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

int csr_check(const GtokGraphBatch* b, int32_t W, const char* who) {
  GTOK_REQUIRE(b != nullptr, GTOK_EINVAL, "%s: null batch", who);
  GTOK_REQUIRE(b->num_graphs >= 0, GTOK_EINVAL, "%s: negative graph count", who);
  GTOK_REQUIRE(W == 1 || W == 2 || W == 4, GTOK_EINVAL, "%s: adj_words must be 1, 2 or 4 (got %d)", who, W);
  if (b->num_graphs > 0) {
    GTOK_REQUIRE(b->node_off && b->edge_off, GTOK_EINVAL, "%s: null node_off / edge_off", who);   // src / dst may be null without columns
    GTOK_REQUIRE(b->max_nodes > 0 && b->max_nodes <= 64 * W, GTOK_EINVAL,
                 "%s: max_nodes bound required and must fit %d adjacency words (got %d)", who, W, b->max_nodes);
  }
  return GTOK_OK;
}

}  // namespace
}  // namespace gtok

extern "C" int gtok_csr_adjacency(const GtokGraphBatch* batch, int32_t adj_words, int64_t total_nodes, uint64_t* adj64,
                                  int32_t* degree_out, void* stream) {
  using namespace gtok;
  if (int rc = csr_check(batch, adj_words, "csr_adjacency")) return rc;
  GTOK_REQUIRE(total_nodes >= 0, GTOK_EINVAL, "csr_adjacency: negative node count");
  if (total_nodes == 0) return GTOK_OK;
  GTOK_REQUIRE(adj64 && degree_out, GTOK_EINVAL, "csr_adjacency: null output");
  cudaStream_t st = (cudaStream_t)stream;
  GTOK_CUDA_TRY(cudaMemsetAsync(adj64, 0, (size_t)total_nodes * adj_words * 8, st));
  if (batch->num_graphs > 0)
    adjacency_kernel<<<grid_for(batch->num_graphs * 32), kCsrThreads, 0, st>>>(*batch, adj_words,
                                                                               reinterpret_cast<unsigned long long*>(adj64));
  degree_kernel<<<grid_for(total_nodes), kCsrThreads, 0, st>>>(reinterpret_cast<const unsigned long long*>(adj64), adj_words,
                                                               total_nodes, degree_out);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

extern "C" int gtok_csr_fill(const GtokGraphBatch* batch, int32_t adj_words, int64_t total_nodes, const uint64_t* adj64,
                             const int64_t* rowptr64, int64_t nnz, int32_t* rowptr_out, int32_t* col_out, int32_t* eattr_out,
                             uint64_t* first_col_ws, void* stream) {
  using namespace gtok;
  if (int rc = csr_check(batch, adj_words, "csr_fill")) return rc;
  GTOK_REQUIRE(total_nodes >= 0 && nnz >= 0 && nnz < (1ll << 31), GTOK_EINVAL, "csr_fill: sizes out of range (nnz must stay below 2^31)");
  GTOK_REQUIRE(rowptr64 && rowptr_out, GTOK_EINVAL, "csr_fill: null row pointers");
  GTOK_REQUIRE(total_nodes == 0 || adj64, GTOK_EINVAL, "csr_fill: null adjacency rows");
  GTOK_REQUIRE(nnz == 0 || col_out, GTOK_EINVAL, "csr_fill: null col_out");
  cudaStream_t st = (cudaStream_t)stream;
  const auto* adj = reinterpret_cast<const unsigned long long*>(adj64);
  fill_kernel<<<grid_for(total_nodes + 1), kCsrThreads, 0, st>>>(adj, adj_words, total_nodes, rowptr64, rowptr_out, col_out);
  if (eattr_out && nnz > 0) {
    GTOK_REQUIRE(batch->edge_attr && first_col_ws, GTOK_EINVAL, "csr_fill: labels need batch.edge_attr and the first-column workspace");
    auto* fc = reinterpret_cast<unsigned long long*>(first_col_ws);
    GTOK_CUDA_TRY(cudaMemsetAsync(fc, 0xff, (size_t)nnz * 8, st));
    first_col_kernel<<<grid_for(batch->num_graphs * 32), kCsrThreads, 0, st>>>(*batch, adj_words, adj, rowptr64, fc);
    label_kernel<<<grid_for(nnz), kCsrThreads, 0, st>>>(fc, batch->edge_attr, nnz, eattr_out);
  }
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}
