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
// One warp per molecule, two code paths chosen by batch.max_nodes:
//  * N <= 64 (every real molecule set): the "seen" set is an N x 64-bit adjacency
//    bitmap in the warp's shared memory.  Columns are taken 32 at a time in order;
//    __match_any_sync groups equal pairs inside the chunk (lowest lane = first
//    occurrence), the bitmap answers "seen in an earlier chunk", survivors are
//    ranked with ballot/popc and write their four tokens straight into a staged
//    int32 row in shared memory; atoms are staged lane-per-atom; the row then
//    leaves with aligned 16-byte streaming stores.  ~200 warp instructions per
//    molecule (v1's generic position-centric path needed ~670 and was issue bound,
//    profiles/r1a_v1_full.txt).  The lengths pass only needs the NUMBER of distinct
//    pairs: one atomicOr per column, then a popcount of the bitmap.
//  * N > 64: open-addressing hash in shared memory (atomicCAS claims the slot for
//    the key, atomicMin keeps the smallest column index), position-centric write.
// Algorithmic bytes / molecule: 8*Ed (edge_index) + 4*Ed (edge_attr) + 4N (x) +
// 16 (offsets) + 8L.  The lengths pass re-reads edge_index (8*Ed).
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
            int32_t* __restrict__ len_out, int64_t* __restrict__ out, const int32_t* __restrict__ row_len) {
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
    const int L = row_len ? __ldg(row_len + g) : (int)(o1 - o0);  // == Lfull by construction (row_len: rows placed by the caller)
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


// ---------------------------------------------------------------------------
// N <= 64 path
// ---------------------------------------------------------------------------
constexpr int kSmallN = 64;

__device__ __forceinline__ void load_graph_header(const GtokGraphBatch& b, const int64_t* __restrict__ off, int64_t g,
                                                  int lane, bool with_off, int64_t& n0, int& N, int64_t& e0, int& Ed,
                                                  int64_t& o0, int& L) {
  int64_t s = 0;
  if (lane < 2) s = __ldg(b.node_off + g + lane);
  else if (lane < 4) s = __ldg(b.edge_off + g + lane - 2);
  else if (with_off && lane < 6) s = __ldg(off + g + lane - 4);
  n0 = __shfl_sync(kFull, s, 0);
  N = (int)(__shfl_sync(kFull, s, 1) - n0);
  e0 = __shfl_sync(kFull, s, 2);
  Ed = (int)(__shfl_sync(kFull, s, 3) - e0);
  o0 = __shfl_sync(kFull, s, 4);
  L = (int)(__shfl_sync(kFull, s, 5) - o0);
}

__global__ void __launch_bounds__(kZincThreads)
zinc_small_lengths_kernel(const GtokGraphBatch b, int item_max, int ds_max, int32_t* __restrict__ len_out) {
  __shared__ unsigned s_rows[kZincWarps][2 * kSmallN];
  const int lane = lane_id();
  unsigned* rows = s_rows[threadIdx.x >> 5];
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = warp0; g < b.num_graphs; g += nwarps) {
    int64_t n0, e0, o0; int N, Ed, L;
    load_graph_header(b, nullptr, g, lane, false, n0, N, e0, Ed, o0, L);
    const int32_t* __restrict__ src = b.src + e0;
    const int32_t* __restrict__ dst = b.dst + e0;
    for (int i = lane; i < 2 * N; i += kWarp) rows[i] = 0u;
    __syncwarp();
    for (int i = lane; i < Ed; i += kWarp) {
      const unsigned a = (unsigned)__ldg(src + i) & 63u, c = (unsigned)__ldg(dst + i) & 63u;
      const unsigned mn = min(a, c), mx = max(a, c);
      atomicOr(&rows[2 * mn + (mx >> 5)], 1u << (mx & 31));
    }
    __syncwarp();
    int cnt = 0;
    for (int i = lane; i < 2 * N; i += kWarp) cnt += __popc(rows[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(kFull, cnt, d);
    int eos_pos;
    if (lane == 0) len_out[g] = zinc_row_length(N, cnt, item_max, ds_max, &eos_pos);
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kZincThreads, 8)
zinc_small_serialize_kernel(const GtokGraphBatch b, const __grid_constant__ GtokZincVocab v,
                            const int64_t* __restrict__ off, int64_t* __restrict__ out, int cap,
                            const int32_t* __restrict__ row_len) {
  extern __shared__ __align__(16) unsigned char z_smem[];
  const int lane = lane_id();
  const int w = threadIdx.x >> 5;
  // per warp: 2*64 bitmap words, then the staged row (cap int32)
  unsigned* rows = reinterpret_cast<unsigned*>(z_smem) + (size_t)w * (2 * kSmallN + cap);
  int* stage = reinterpret_cast<int*>(rows + 2 * kSmallN);
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int32_t* __restrict__ node_tok = v.node_tok;
  const int tbl = v.node_tok_len;
  const unsigned lt = (1u << lane) - 1u;

  for (int64_t g = warp0; g < b.num_graphs; g += nwarps) {
    int64_t n0, e0, o0; int N, Ed, L;
    load_graph_header(b, off, g, lane, true, n0, N, e0, Ed, o0, L);
    if (row_len) L = __ldg(row_len + g);               // rows placed by the caller (padded batches): off[g] = row start
    const int32_t* __restrict__ src = b.src + e0;
    const int32_t* __restrict__ dst = b.dst + e0;
    const int32_t* __restrict__ attr = b.edge_attr ? b.edge_attr + e0 : nullptr;
    const int32_t* __restrict__ x = b.x + n0;

    for (int i = lane; i < 2 * N; i += kWarp) rows[i] = 0u;
    if (lane == 0) stage[0] = v.tok_bos;
    // atoms: lane per atom, positions 1+2i, 2+2i
    for (int i = lane; i < N; i += kWarp) {
      const int xi = __ldg(x + i);
      const int p = 1 + 2 * i;
      if (p < cap) stage[p] = v.tok_atom;
      if (p + 1 < cap) stage[p + 1] = ((unsigned)xi < 9u) ? v.atom_tok[xi] : v.atom_oov;
    }
    __syncwarp();
    // bonds: 32 columns at a time, in column order
    const int bbase = 1 + 2 * N;
    int Eu = 0;
    for (int i0 = 0; i0 < Ed; i0 += kWarp) {
      const int i = i0 + lane;
      const bool valid = i < Ed;
      int a = 0, c = 0, t = 0;
      if (valid) {
        a = __ldg(src + i);
        c = __ldg(dst + i);
        t = attr ? __ldg(attr + i) : 0;
      }
      const unsigned mn = (unsigned)min(a, c) & 63u, mx = (unsigned)max(a, c) & 63u;
      const unsigned key = valid ? (mn << 6 | mx) : (0x10000u | (unsigned)lane);
      const unsigned grp = __match_any_sync(kFull, key);
      const unsigned word = 2 * mn + (mx >> 5), bit = 1u << (mx & 31);
      const bool keep = valid && ((grp & lt) == 0u) && ((rows[word] & bit) == 0u);
      __syncwarp();
      if (keep) atomicOr(&rows[word], bit);
      const unsigned m = __ballot_sync(kFull, keep);
      if (keep) {
        const int p = bbase + 4 * (Eu + __popc(m & lt));
        if (p + 3 < cap) {
          stage[p] = v.tok_bond;
          stage[p + 1] = ((unsigned)(t - 1) < 4u) ? v.bond_tok[t - 1] : v.bond_oov;
          stage[p + 2] = ((unsigned)a < (unsigned)tbl) ? __ldg(node_tok + a) : v.oov;
          stage[p + 3] = ((unsigned)c < (unsigned)tbl) ? __ldg(node_tok + c) : v.oov;
        } else {
          if (p < cap) stage[p] = v.tok_bond;
          if (p + 1 < cap) stage[p + 1] = ((unsigned)(t - 1) < 4u) ? v.bond_tok[t - 1] : v.bond_oov;
          if (p + 2 < cap) stage[p + 2] = ((unsigned)a < (unsigned)tbl) ? __ldg(node_tok + a) : v.oov;
        }
      }
      Eu += __popc(m);
      __syncwarp();
    }
    int eos_pos;
    zinc_row_length(N, Eu, v.item_max_len, v.ds_max_len, &eos_pos);
    const int tbase = bbase + 4 * Eu;
    if (lane < 3 && tbase + lane < cap) stage[tbase + lane] = lane == 0 ? v.tok_q : (lane == 1 ? v.tok_regression : v.tok_p);
    __syncwarp();
    if (lane == 0 && eos_pos >= 0 && eos_pos < cap) stage[eos_pos] = v.tok_eos;
    __syncwarp();
    // aligned 16-byte streaming write-out of the staged row
    int64_t* __restrict__ row = out + o0;
    const int head = (int)(o0 & 1);
    if (head && lane == 0 && L > 0) st_stream_i64(row, (int64_t)stage[0]);
    const int npair = (L - head) >> 1;
    for (int j = lane; j < npair; j += kWarp) {
      const int p = head + 2 * j;
      st_stream_i64x2(row + p, (int64_t)stage[p], (int64_t)stage[p + 1]);
    }
    const int tail = head + 2 * npair;
    if (tail < L && lane == 31) st_stream_i64(row + tail, (int64_t)stage[tail]);
    __syncwarp();
  }
}

static int zinc_check(const GtokGraphBatch* b, const GtokZincVocab* v, const char* who) {
  GTOK_REQUIRE(b != nullptr && v != nullptr, GTOK_EINVAL, "%s: null batch/vocab", who);
  GTOK_REQUIRE(b->num_graphs >= 0, GTOK_EINVAL, "%s: negative graph count", who);
  if (b->num_graphs == 0) return GTOK_OK;
  GTOK_REQUIRE(b->node_off && b->edge_off, GTOK_EINVAL, "%s: null graph arrays", who);   // src/dst may be null without edges
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
                       int64_t* out, const int32_t* row_len, cudaStream_t st) {
  auto k = zinc_kernel<H, MAXE, WRITE>;
  const int64_t work = (b->num_graphs + kZincWarps - 1) / kZincWarps;
  const int grid = persistent_grid(k, kZincThreads, 0, work);
  k<<<grid, kZincThreads, 0, st>>>(*b, *v, off, len_out, out, row_len);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

static int zinc_small_launch(bool write, const GtokGraphBatch* b, const GtokZincVocab* v, const int64_t* off,
                             int32_t* len_out, int64_t* out, const int32_t* row_len, cudaStream_t st) {
  const int64_t work = (b->num_graphs + kZincWarps - 1) / kZincWarps;
  if (!write) {
    const int grid = persistent_grid(zinc_small_lengths_kernel, kZincThreads, 0, work);
    zinc_small_lengths_kernel<<<grid, kZincThreads, 0, st>>>(*b, v->item_max_len, v->ds_max_len, len_out);
    GTOK_CUDA_TRY(cudaGetLastError());
    return GTOK_OK;
  }
  long long bound = 2ll * b->max_nodes + 4ll * b->max_entries + 4;
  const long long lim = (long long)(v->item_max_len > v->ds_max_len ? v->item_max_len : v->ds_max_len) + 1;
  if (bound > lim) bound = lim;
  const int cap = (int)((bound + 4 + 3) & ~3ll);
  const size_t smem = (size_t)kZincWarps * (2 * kSmallN + cap) * 4;
  GTOK_CUDA_TRY(cudaFuncSetAttribute(zinc_small_serialize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = persistent_grid(zinc_small_serialize_kernel, kZincThreads, smem, work);
  zinc_small_serialize_kernel<<<grid, kZincThreads, smem, st>>>(*b, *v, off, out, cap, row_len);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

template <bool WRITE>
static int zinc_dispatch(const GtokGraphBatch* b, const GtokZincVocab* v, const int64_t* off, int32_t* len_out,
                         int64_t* out, cudaStream_t st, const int32_t* row_len = nullptr) {
  if (b->max_nodes <= kSmallN && b->max_entries <= 256) return zinc_small_launch(WRITE, b, v, off, len_out, out, row_len, st);
  // static shared memory budget is 48 KB per CTA: 8 warps * (8 B * H + 2 B * MAXE)
  if (b->max_entries <= 128) return zinc_launch<256, 128, WRITE>(b, v, off, len_out, out, row_len, st);
  return zinc_launch<512, 256, WRITE>(b, v, off, len_out, out, row_len, st);
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

extern "C" int gtok_zinc_serialize_rows(const GtokGraphBatch* batch, const GtokZincVocab* vocab, const int64_t* row_off,
                                        const int32_t* row_len, int64_t* tok_out, void* stream) {
  using namespace gtok;
  int rc = zinc_check(batch, vocab, "zinc_serialize_rows");
  if (rc) return rc;
  if (batch->num_graphs == 0) return GTOK_OK;
  GTOK_REQUIRE(row_off && row_len && tok_out, GTOK_EINVAL, "zinc_serialize_rows: null row_off/row_len/tok_out");
  GTOK_REQUIRE(batch->x != nullptr, GTOK_EINVAL, "zinc_serialize_rows: batch.x (atom types) is required");
  return zinc_dispatch<true>(batch, vocab, row_off, nullptr, tok_out, (cudaStream_t)stream, row_len);
}
