// gtok_zinc.cu - K2: index-based serialization of ZINC-shaped molecular graphs.
//
// Replaces (reference, pure Python, per molecule):
//   ZINCTokenizationDataset.tokenize_molecule / decode_atom_features /
//   decode_bond_features / __getitem__ truncation
//       (graph_data_loader/zinc_dataset_indexbase.py:86-227)
//   + TokenDataset.__init__ id lookup / strip after <p> / [:max_len]
//       (graph_data_loader/data_loader.py:465-486)
//   with the fixed vocabulary of graph_data_loader/zinc_vocab.py:35-78.
// Row layout, E = first-seen undirected pairs:
//   [<bos>] [<atom> id(atom x_i)]*N [<bond> id(type) id(str(u)) id(str(v))]*E [<q> regression <p>]
// "first-seen" = edge_index column i is emitted iff no earlier column holds the
// same unordered pair (set() dedupe, zinc_dataset_indexbase.py:176-184), in column
// order and in THAT column's (u, v) orientation.
//
// One warp per molecule.  Dedupe without sorting: each column inserts its packed
// unordered pair into a per-warp shared-memory open-addressing table (atomicCAS
// claims the slot for the key, atomicMin keeps the smallest column index); a
// column survives iff it is its slot's minimum.  Survivors are compacted in
// column order with ballot/popc, then the row is written position-centric with
// 16-byte streaming stores (gtok_common.cuh: warp_write_row).
// Algorithmic bytes / molecule: 8*Ed (edge_index) + 4*Ed (edge_attr) + 4N (x) +
// 16 (offsets) + 8L.  This v1 reads edge_index twice (lengths pass + write pass).
#include "gtok_common.cuh"

namespace gtok {

constexpr int kZincThreads = 256;
constexpr int kZincWarps = kZincThreads / 32;
constexpr unsigned kEmpty = 0xffffffffu;

__host__ __device__ __forceinline__ int zinc_row_length(int N, int Eu, int item_max, int ds_max, int* eos_pos) {
  // oracle/ibtt_oracle.py: zinc_length.  Full text has T = 2N+4E+6 tokens, <p> at T-3.
  const long long T = 2ll * N + 4ll * Eu + 6;
  long long L;
  *eos_pos = -1;
  if (T <= item_max) L = T - 2;
  else if (T == (long long)item_max + 1) L = item_max - 1;
  else { L = item_max; *eos_pos = item_max - 1; }
  if (L > ds_max) L = ds_max;
  if (L < 0) L = 0;
  return (int)L;
}

template <int H, int MAXE, bool WRITE>
__global__ void __launch_bounds__(kZincThreads)
zinc_kernel(const GtokGraphBatch b, const __grid_constant__ GtokZincVocab v, const int64_t* __restrict__ off,
            int32_t* __restrict__ len_out, int64_t* __restrict__ out) {
  __shared__ unsigned s_key[kZincWarps][H];
  __shared__ unsigned s_min[kZincWarps][H];
  __shared__ unsigned short s_kept[WRITE ? kZincWarps : 1][WRITE ? MAXE : 1];

  const int lane = lane_id();
  const int w = threadIdx.x >> 5;
  unsigned* key_tbl = s_key[w];
  unsigned* min_tbl = s_min[w];
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int kShift = 32 - __builtin_ctz(H);

  for (int64_t g = warp0; g < b.num_graphs; g += nwarps) {
    int64_t s = 0;
    if (lane < 2) s = __ldg(b.node_off + g + lane);
    else if (lane < 4) s = __ldg(b.edge_off + g + lane - 2);
    else if (WRITE && lane < 6) s = __ldg(off + g + lane - 4);
    const int64_t n0 = __shfl_sync(kFull, s, 0), n1 = __shfl_sync(kFull, s, 1);
    const int64_t e0 = __shfl_sync(kFull, s, 2), e1 = __shfl_sync(kFull, s, 3);
    const int64_t o0 = __shfl_sync(kFull, s, 4), o1 = __shfl_sync(kFull, s, 5);
    const int N = (int)(n1 - n0);
    const int Ed = (int)(e1 - e0);
    const int32_t* __restrict__ src = b.src + e0;
    const int32_t* __restrict__ dst = b.dst + e0;

    for (int i = lane; i < H; i += kWarp) { key_tbl[i] = kEmpty; min_tbl[i] = kEmpty; }
    __syncwarp();
    // insert every column; remember nothing, re-probe below (probe chains are ~1 long at load <= 0.5)
    for (int i = lane; i < Ed; i += kWarp) {
      const unsigned a = (unsigned)__ldg(src + i), c = (unsigned)__ldg(dst + i);
      const unsigned key = (min(a, c) << 16) | max(a, c);
      unsigned h = (key * 0x9E3779B1u) >> kShift;
      while (true) {
        const unsigned old = atomicCAS(&key_tbl[h], kEmpty, key);
        if (old == kEmpty || old == key) break;
        h = (h + 1) & (H - 1);
      }
      atomicMin(&min_tbl[h], (unsigned)i);
    }
    __syncwarp();
    int Eu = 0;
    for (int i0 = 0; i0 < Ed; i0 += kWarp) {
      const int i = i0 + lane;
      bool keep = false;
      if (i < Ed) {
        const unsigned a = (unsigned)__ldg(src + i), c = (unsigned)__ldg(dst + i);
        const unsigned key = (min(a, c) << 16) | max(a, c);
        unsigned h = (key * 0x9E3779B1u) >> kShift;
        while (key_tbl[h] != key) h = (h + 1) & (H - 1);
        keep = (min_tbl[h] == (unsigned)i);
      }
      const unsigned m = __ballot_sync(kFull, keep);
      if (WRITE && keep) s_kept[w][Eu + __popc(m & ((1u << lane) - 1))] = (unsigned short)i;
      Eu += __popc(m);
    }
    int eos_pos;
    const int Lfull = zinc_row_length(N, Eu, v.item_max_len, v.ds_max_len, &eos_pos);
    if (!WRITE) {
      if (lane == 0) len_out[g] = Lfull;
      __syncwarp();
      continue;
    }
    __syncwarp();
    const int L = (int)(o1 - o0);  // == Lfull by construction
    const int32_t* __restrict__ x = b.x + n0;
    const int32_t* __restrict__ attr = b.edge_attr ? b.edge_attr + e0 : nullptr;
    const int32_t* __restrict__ node_tok = v.node_tok;
    const int tbl = v.node_tok_len;
    const int A2 = 2 * N, B4 = 4 * Eu;
    const unsigned short* kept = s_kept[w];
    auto node = [&](int k) -> int { return ((unsigned)k < (unsigned)tbl) ? __ldg(node_tok + k) : v.oov; };
    auto tok = [&](int p) -> int {
      if (p == eos_pos) return v.tok_eos;
      if (p == 0) return v.tok_bos;
      int q = p - 1;
      if (q < A2) {
        if ((q & 1) == 0) return v.tok_atom;
        const int xi = __ldg(x + (q >> 1));
        return ((unsigned)xi < 9u) ? v.atom_tok[xi] : v.atom_oov;
      }
      q -= A2;
      if (q < B4) {
        const int i = kept[q >> 2];
        const int r = q & 3;
        if (r == 0) return v.tok_bond;
        if (r == 1) {
          const int bt = attr ? __ldg(attr + i) : 0;
          return ((unsigned)(bt - 1) < 4u) ? v.bond_tok[bt - 1] : v.bond_oov;
        }
        return node(r == 2 ? __ldg(src + i) : __ldg(dst + i));
      }
      q -= B4;
      return q == 0 ? v.tok_q : (q == 1 ? v.tok_regression : v.tok_p);
    };
    warp_write_row(out + o0, L, tok);
    __syncwarp();  // kept[] / tables are reused by the next molecule
  }
}

static int zinc_check(const GtokGraphBatch* b, const GtokZincVocab* v, const char* who) {
  GTOK_REQUIRE(b != nullptr && v != nullptr, GTOK_EINVAL, "%s: null batch/vocab", who);
  GTOK_REQUIRE(b->num_graphs >= 0, GTOK_EINVAL, "%s: negative graph count", who);
  if (b->num_graphs == 0) return GTOK_OK;
  GTOK_REQUIRE(b->node_off && b->edge_off && b->src && b->dst, GTOK_EINVAL, "%s: null graph arrays", who);
  GTOK_REQUIRE(v->item_max_len >= 2 && v->ds_max_len >= 0, GTOK_EINVAL, "%s: max_len must be >= 2", who);
  GTOK_REQUIRE(v->node_tok_len == 0 || v->node_tok != nullptr, GTOK_EINVAL, "%s: null node_tok table", who);
  GTOK_REQUIRE(b->max_entries > 0 && b->max_nodes > 0, GTOK_EINVAL,
               "%s: batch.max_nodes / batch.max_entries bounds are required", who);
  GTOK_REQUIRE(b->max_nodes <= 65535, GTOK_EUNSUPPORTED, "%s: molecules with more than 65535 atoms are not supported", who);
  GTOK_REQUIRE(b->max_entries <= 256, GTOK_EUNSUPPORTED,
               "%s: more than 256 edge_index columns per molecule (got bound %d) is not supported", who, b->max_entries);
  return GTOK_OK;
}

template <int H, int MAXE, bool WRITE>
static int zinc_launch(const GtokGraphBatch* b, const GtokZincVocab* v, const int64_t* off, int32_t* len_out,
                       int64_t* out, cudaStream_t st) {
  auto k = zinc_kernel<H, MAXE, WRITE>;
  const int64_t work = (b->num_graphs + kZincWarps - 1) / kZincWarps;
  const int grid = persistent_grid(k, kZincThreads, 0, work);
  k<<<grid, kZincThreads, 0, st>>>(*b, *v, off, len_out, out);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

template <bool WRITE>
static int zinc_dispatch(const GtokGraphBatch* b, const GtokZincVocab* v, const int64_t* off, int32_t* len_out,
                         int64_t* out, cudaStream_t st) {
  // static shared memory budget is 48 KB per CTA: 8 warps * (8 B * H + 2 B * MAXE)
  if (b->max_entries <= 128) return zinc_launch<256, 128, WRITE>(b, v, off, len_out, out, st);
  return zinc_launch<512, 256, WRITE>(b, v, off, len_out, out, st);
}

}  // namespace gtok

extern "C" int gtok_zinc_lengths(const GtokGraphBatch* batch, const GtokZincVocab* vocab, int32_t* len_out,
                                 void* stream) {
  using namespace gtok;
  int rc = zinc_check(batch, vocab, "zinc_lengths");
  if (rc) return rc;
  if (batch->num_graphs == 0) return GTOK_OK;
  GTOK_REQUIRE(len_out != nullptr, GTOK_EINVAL, "zinc_lengths: null output");
  return zinc_dispatch<false>(batch, vocab, nullptr, len_out, nullptr, (cudaStream_t)stream);
}

extern "C" int gtok_zinc_serialize(const GtokGraphBatch* batch, const GtokZincVocab* vocab, const int64_t* off,
                                   int64_t* tok_out, void* stream) {
  using namespace gtok;
  int rc = zinc_check(batch, vocab, "zinc_serialize");
  if (rc) return rc;
  if (batch->num_graphs == 0) return GTOK_OK;
  GTOK_REQUIRE(off != nullptr && tok_out != nullptr, GTOK_EINVAL, "zinc_serialize: null off/tok_out");
  GTOK_REQUIRE(batch->x != nullptr, GTOK_EINVAL, "zinc_serialize: batch.x (atom types) is required");
  GTOK_REQUIRE((reinterpret_cast<uintptr_t>(tok_out) & 15) == 0, GTOK_EINVAL, "zinc_serialize: tok_out must be 16-byte aligned");
  return zinc_dispatch<true>(batch, vocab, off, nullptr, tok_out, (cudaStream_t)stream);
}
