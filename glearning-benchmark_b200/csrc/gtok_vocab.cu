// gtok_vocab.cu - K5: token statistics for build_vocab_from_texts, straight from the graphs.
//
// Replaces the O(total tokens) host pass of graph_data_loader/data_loader.py:451-463
//   cnt = Counter(); for text in texts: cnt.update(text.split()); ... cnt.most_common()
// for graph-token texts  <bos> (u v <e>)*E <n> 0..N-1 <q> TASK [u v] <p> LABEL <eos>
// (docs/synthetic_data.md:46-68).  Everything except the node-index strings occurs a closed-form number
// of times (<e>: sum E, TASK: G, ...), so the device only has to produce, for every node index k:
//   counts[k]  how often the string str(k) occurs over all texts (edge endpoints + node list + query), and
//   first[k]   the (text, position) of its FIRST occurrence as one 64-bit key (g << 32 | position) -
//              Counter.most_common() is a stable sort over first-insertion order, so ties between equally
//              frequent tokens are broken by exactly this key.
// The host finishes the job on <= max_nodes + a handful of entries (gtok/engine.py vocab_from_graphs).
// One warp per graph, statistics privatised per CTA in shared memory, merged with one atomic per bin.
#include "gtok_common.cuh"

namespace gtok {

constexpr int kVocThreads = 256;
constexpr int kVocBinsMax = 4096;      // bins kept in shared memory (48 KB); larger tables update HBM directly

__global__ void __launch_bounds__(kVocThreads)
vocab_hist_kernel(GtokGraphBatch b, int task, int bins, int smem_bins, unsigned long long* __restrict__ counts,
                  unsigned long long* __restrict__ first) {
  extern __shared__ __align__(16) unsigned char v_smem[];
  unsigned* s_cnt = reinterpret_cast<unsigned*>(v_smem);
  unsigned long long* s_first = reinterpret_cast<unsigned long long*>(v_smem + (((size_t)smem_bins * 4 + 7) & ~(size_t)7));
  for (int i = threadIdx.x; i < smem_bins; i += kVocThreads) { s_cnt[i] = 0u; s_first[i] = ~0ull; }
  __syncthreads();
  auto hit = [&](int k, unsigned long long key) {
    if ((unsigned)k >= (unsigned)bins) return;
    if (k < smem_bins) {
      atomicAdd(&s_cnt[k], 1u);
      if (key < s_first[k]) atomicMin(&s_first[k], key);     // settles after the first few graphs
    } else {
      atomicAdd(&counts[k], 1ull);
      if (key < first[k]) atomicMin(&first[k], key);
    }
  };
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = warp0; g < b.num_graphs; g += nwarps) {
    const int64_t n0 = __ldg(b.node_off + g), e0 = __ldg(b.edge_off + g);
    const int N = (int)(__ldg(b.node_off + g + 1) - n0), E = (int)(__ldg(b.edge_off + g + 1) - e0);
    const unsigned long long gk = (unsigned long long)g << 32;
    for (int e = lane; e < E; e += kWarp) {
      hit(__ldcs(b.src + e0 + e), gk | (unsigned)(1 + 3 * e));
      hit(__ldcs(b.dst + e0 + e), gk | (unsigned)(2 + 3 * e));
    }
    for (int k = lane; k < N; k += kWarp) hit(k, gk | (unsigned)(3 * E + 2 + k));
    if (task == GTOK_TASK_SHORTEST_PATH && lane < 2)
      hit(lane == 0 ? __ldg(b.query_u + g) : __ldg(b.query_v + g), gk | (unsigned)(3 * E + N + 4 + lane));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < smem_bins; i += kVocThreads) {
    if (s_cnt[i]) {
      atomicAdd(&counts[i], (unsigned long long)s_cnt[i]);
      atomicMin(&first[i], s_first[i]);
    }
  }
}

}  // namespace gtok

extern "C" int gtok_vocab_histogram(const GtokGraphBatch* batch, int32_t task, int32_t num_bins, uint64_t* counts,
                                    uint64_t* first, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(batch != nullptr, GTOK_EINVAL, "vocab_histogram: null batch");
  GTOK_REQUIRE(batch->num_graphs >= 0 && num_bins >= 0, GTOK_EINVAL, "vocab_histogram: negative size");
  GTOK_REQUIRE(task == GTOK_TASK_CYCLE_CHECK || task == GTOK_TASK_SHORTEST_PATH, GTOK_EINVAL, "vocab_histogram: unknown task %d", task);
  if (num_bins == 0) return GTOK_OK;
  GTOK_REQUIRE(counts && first, GTOK_EINVAL, "vocab_histogram: null output");
  cudaStream_t st = (cudaStream_t)stream;
  GTOK_CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)num_bins * 8, st));
  GTOK_CUDA_TRY(cudaMemsetAsync(first, 0xff, (size_t)num_bins * 8, st));
  if (batch->num_graphs == 0) return GTOK_OK;
  GTOK_REQUIRE(batch->node_off && batch->edge_off, GTOK_EINVAL, "vocab_histogram: null graph arrays");   // src/dst may be null without edges
  GTOK_REQUIRE(task != GTOK_TASK_SHORTEST_PATH || (batch->query_u && batch->query_v), GTOK_EINVAL,
               "vocab_histogram: shortest_path needs query_u/query_v");
  const int smem_bins = num_bins < kVocBinsMax ? num_bins : kVocBinsMax;
  const size_t smem = (((size_t)smem_bins * 4 + 7) & ~(size_t)7) + (size_t)smem_bins * 8;
  GTOK_CUDA_TRY(cudaFuncSetAttribute(vocab_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t work = (batch->num_graphs + (kVocThreads / 32) - 1) / (kVocThreads / 32);
  const int grid = persistent_grid(vocab_hist_kernel, kVocThreads, smem, work);
  vocab_hist_kernel<<<grid, kVocThreads, smem, st>>>(*batch, task, num_bins, smem_bins,
                                                      reinterpret_cast<unsigned long long*>(counts),
                                                      reinterpret_cast<unsigned long long*>(first));
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}
