// gtok_ibtt.cu - K1: index-based serialization of graph-token graphs.
//
// Replaces (reference, pure Python, per graph):
//   * the external graph-token serializer's text
//       <bos> (u v <e>)*E <n> 0..N-1 <q> TASK.. <p> LABEL <eos>
//     (docs/synthetic_data.md:46-68, docs/models.md:22), and
//   * TokenDataset.__init__ (graph_data_loader/data_loader.py:465-486): split,
//     cut after the first <p>, vocab.get(tok, <pad>), [:max_len], torch.long.
// Closed form: row g has L = min(max_len, 3E+N+Q) tokens,
// Q = 5 (cycle_check) / 7 (shortest_path), and every token is a pure function of
// (g, position) - so the kernel is a gather with NO inter-thread dependency:
//     p = 0            <bos>
//     1 <= p <= 3E     edge (p-1)/3: node_tok[src], node_tok[dst], <e>
//     p = 3E+1         <n>
//     next N           node_tok[p-3E-2]
//     tail             <q> TASK [node_tok[u] node_tok[v]] <p>
// node ids go through the host-built table node_tok[k] = vocab.get(str(k), <pad>)
// (vocab ids are frequency ordered, data_loader.py:451-463; never base+k).
//
// Mapping: one warp per graph, persistent grid-stride; each lane produces two
// consecutive tokens and issues one 16-byte streaming store, so a warp store
// instruction covers 512 contiguous bytes of the ragged int64 stream.
// Algorithmic bytes / graph: 8E (src,dst) + 16 (offsets) [+8 query] + 8L.
#include "gtok_common.cuh"

namespace gtok {

constexpr int kIbttThreads = 256;

__global__ void __launch_bounds__(kIbttThreads)
ibtt_lengths_kernel(GtokGraphBatch b, int max_len, int qlen, int32_t* __restrict__ len_out) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= b.num_graphs) return;
  const int64_t n = b.node_off[g + 1] - b.node_off[g];
  const int64_t e = b.edge_off[g + 1] - b.edge_off[g];
  const int64_t full = 3 * e + n + qlen;
  len_out[g] = (int32_t)(full < max_len ? full : max_len);
}

template <int TASK>
__global__ void __launch_bounds__(kIbttThreads)
ibtt_serialize_kernel(GtokGraphBatch b, GtokIbttRules r, const int64_t* __restrict__ off,
                      int64_t* __restrict__ out) {
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int32_t* __restrict__ node_tok = r.node_tok;
  const int tbl = r.node_tok_len;
  const int oov = r.oov;

  for (int64_t g = warp0; g < b.num_graphs; g += nwarps) {
    // 6 warp-uniform scalars; lanes 0..5 fetch one each, then broadcast.
    int64_t s = 0;
    if (lane < 2) s = __ldg(b.node_off + g + lane);
    else if (lane < 4) s = __ldg(b.edge_off + g + lane - 2);
    else if (lane < 6) s = __ldg(off + g + lane - 4);
    const int64_t n0 = __shfl_sync(kFull, s, 0), n1 = __shfl_sync(kFull, s, 1);
    const int64_t e0 = __shfl_sync(kFull, s, 2), e1 = __shfl_sync(kFull, s, 3);
    const int64_t o0 = __shfl_sync(kFull, s, 4), o1 = __shfl_sync(kFull, s, 5);
    const int N = (int)(n1 - n0);
    const int E3 = 3 * (int)(e1 - e0);
    const int L = (int)(o1 - o0);
    const int32_t* __restrict__ src = b.src + e0;
    const int32_t* __restrict__ dst = b.dst + e0;
    int qu = 0, qv = 0;
    if (TASK == GTOK_TASK_SHORTEST_PATH) {
      qu = __ldg(b.query_u + g);
      qv = __ldg(b.query_v + g);
    }
    auto node = [&](int k) -> int { return ((unsigned)k < (unsigned)tbl) ? __ldg(node_tok + k) : oov; };
    auto tok = [&](int p) -> int {
      if (p == 0) return r.tok_bos;
      if (p <= E3) {
        const int e = (int)(__umulhi((unsigned)(p - 1), 0xAAAAAAABu) >> 1);  // (p-1)/3, exact for u32
        const int rem = (p - 1) - 3 * e;
        if (rem == 2) return r.tok_e;
        return node(rem == 0 ? __ldg(src + e) : __ldg(dst + e));
      }
      const int t = p - E3 - 1;
      if (t == 0) return r.tok_n;
      if (t <= N) return node(t - 1);
      const int u = t - N - 1;  // 0: <q>, 1: TASK, ...
      if (u == 0) return r.tok_q;
      if (u == 1) return r.tok_task;
      if (TASK == GTOK_TASK_SHORTEST_PATH) {
        if (u == 2) return node(qu);
        if (u == 3) return node(qv);
      }
      return r.tok_p;
    };
    warp_write_row(out + o0, L, tok);
  }
}

// K1+K4 fused: graph-token graphs -> padded batches, the ragged stream never exists.  One CTA per batch
// (grid-stride), one warp per row; batch extents come from the closed-form lengths (gtok_ibtt_lengths ->
// gtok_batch_extents -> gtok_scan_offsets, 16 bytes per graph), so there is no inter-CTA dependency.
// Rows that fit the warp's shared-memory row buffer are built segment-wise (lane <-> edge: 3 tokens, lane <->
// node: 1 token) and streamed out with pad fill + mask; longer rows use the closed form per position.
constexpr int kIcThreads = 256;
constexpr int kIcWarps = kIcThreads / 32;
constexpr int kIcSpan = 4096;           // cells per work item when rows can be long

template <int TASK, typename StageT>
__global__ void __launch_bounds__(kIcThreads)
ibtt_collate_kernel(GtokGraphBatch b, GtokIbttRules r, const int64_t* __restrict__ batch_idx, int64_t G, int B,
                    int64_t nb, const int32_t* __restrict__ lmax, const int64_t* __restrict__ xoff, int64_t pad,
                    int row_cap, int tbl_cap, int spans, int64_t* __restrict__ X, uint8_t* __restrict__ attn) {
  extern __shared__ __align__(16) unsigned char ic_smem[];
  StageT* s_tbl = reinterpret_cast<StageT*>(ic_smem);                     // node_tok[] (first tbl_cap entries)
  const int lane = lane_id(), w = threadIdx.x >> 5;
  StageT* rb = reinterpret_cast<StageT*>(ic_smem + (((size_t)tbl_cap * sizeof(StageT) + 15) & ~(size_t)15)) +
               (size_t)w * row_cap;
  const int32_t* __restrict__ node_tok = r.node_tok;
  const int tbl = r.node_tok_len;
  const int oov = r.oov;
  const int qlen = TASK == GTOK_TASK_CYCLE_CHECK ? 5 : 7;
  for (int i = threadIdx.x; i < tbl_cap; i += kIcThreads) s_tbl[i] = (StageT)(i < tbl ? __ldg(node_tok + i) : oov);
  __syncthreads();
  auto node = [&](int k) -> int {
    if ((unsigned)k < (unsigned)tbl_cap) return (int)s_tbl[k];
    return ((unsigned)k < (unsigned)tbl) ? __ldg(node_tok + k) : oov;
  };

  // rows are dealt to all warps of the grid (consecutive rows -> consecutive warps), so a few huge batches
  // (large-graph sweep) still fill the machine; each warp looks up its batch's extents
  // Rows longer than kIcSpan cells are cut into `spans` independent work items (the closed form needs no
  // neighbour), so a handful of 50k-token rows still occupies every warp.
  const int64_t nwarps = (int64_t)gridDim.x * kIcWarps;
  const int64_t items = G * spans;
  if (spans == 1) {
    // ---- the common case (rows fit the row buffer): a 2-deep software pipeline per warp -------------------
    // row i's tokens are built and streamed out while row i+1's edge columns and row i+2's header are already
    // in flight: with the loads issued right before their use, 63 % of all stall samples sat on them (r1q).
    struct Hdr { int64_t g, e0; int N, E; bool valid; };
    auto load_hdr = [&](int64_t row, Hdr& h) {
      h.valid = row < G;
      h.g = 0; h.e0 = 0; h.N = 0; h.E = 0;
      if (h.valid) {
        h.g = batch_idx ? __ldg(batch_idx + row) : row;
        int64_t s = 0;
        if (lane < 2) s = __ldg(b.node_off + h.g + lane);
        else if (lane < 4) s = __ldg(b.edge_off + h.g + lane - 2);
        const int64_t n0 = __shfl_sync(kFull, s, 0), n1 = __shfl_sync(kFull, s, 1);
        const int64_t e0 = __shfl_sync(kFull, s, 2), e1 = __shfl_sync(kFull, s, 3);
        h.N = (int)(n1 - n0); h.E = (int)(e1 - e0); h.e0 = e0;
      }
    };
    auto load_edges = [&](const Hdr& h, int (&es)[2], int (&ed)[2]) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int e = c * kWarp + lane;
        es[c] = 0; ed[c] = 0;
        if (h.valid && e < h.E) { es[c] = __ldcs(b.src + h.e0 + e); ed[c] = __ldcs(b.dst + h.e0 + e); }
      }
    };
    const int64_t row0 = (int64_t)blockIdx.x * kIcWarps + w;
    Hdr h0, h1, h2;
    int cs[2], cd[2], ns[2], nd[2];
    load_hdr(row0, h0);
    load_hdr(row0 + nwarps, h1);
    load_edges(h0, cs, cd);
    for (int64_t row = row0; row < G; row += nwarps) {
      load_edges(h1, ns, nd);                              // next row's columns: in flight during this row's build + flush
      load_hdr(row + 2 * nwarps, h2);
      const int64_t k = row / B;
      const int rr = (int)(row - k * B);
      const int Lm = __ldg(lmax + k);
      const int64_t o = __ldg(xoff + k) + (int64_t)rr * Lm;
      const int N = h0.N, E = h0.E, E3 = 3 * E;
      const long long full = 3ll * E + N + qlen;
      const int L = (int)(full < r.max_len ? full : r.max_len);
      const int32_t* __restrict__ src = b.src + h0.e0;
      const int32_t* __restrict__ dst = b.dst + h0.e0;
      int qu = 0, qv = 0;
      if (TASK == GTOK_TASK_SHORTEST_PATH) {
        qu = __ldg(b.query_u + h0.g);
        qv = __ldg(b.query_v + h0.g);
      }
      if (full <= row_cap) {
        if (lane == 0) rb[0] = (StageT)r.tok_bos;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int e = c * kWarp + lane;
          if (e < E) {
            rb[1 + 3 * e] = (StageT)node(cs[c]);
            rb[2 + 3 * e] = (StageT)node(cd[c]);
            rb[3 + 3 * e] = (StageT)r.tok_e;
          }
        }
        for (int e = 2 * kWarp + lane; e < E; e += kWarp) {   // more than 64 edges: the rest on demand
          rb[1 + 3 * e] = (StageT)node(__ldcs(src + e));
          rb[2 + 3 * e] = (StageT)node(__ldcs(dst + e));
          rb[3 + 3 * e] = (StageT)r.tok_e;
        }
        if (lane == 0) rb[E3 + 1] = (StageT)r.tok_n;
        for (int i = lane; i < N; i += kWarp) rb[E3 + 2 + i] = (StageT)node(i);
        if (lane == 0) {
          int t = E3 + 2 + N;
          rb[t++] = (StageT)r.tok_q;
          rb[t++] = (StageT)r.tok_task;
          if (TASK == GTOK_TASK_SHORTEST_PATH) { rb[t++] = (StageT)node(qu); rb[t++] = (StageT)node(qv); }
          rb[t] = (StageT)r.tok_p;
        }
        __syncwarp();
        const StageT* __restrict__ rbc = rb;
        warp_pad_row(L, Lm, pad, X + o, attn + o, [&](int p) -> int { return (int)rbc[p]; });
        __syncwarp();                                     // row buffer drained before the next row overwrites it
      } else {
        auto tok = [&](int p) -> int {                    // closed form per position (same as ibtt_serialize_kernel)
          if (p == 0) return r.tok_bos;
          if (p <= E3) {
            const int e = (int)(__umulhi((unsigned)(p - 1), 0xAAAAAAABu) >> 1);
            const int rem = (p - 1) - 3 * e;
            if (rem == 2) return r.tok_e;
            return node(rem == 0 ? __ldg(src + e) : __ldg(dst + e));
          }
          const int t = p - E3 - 1;
          if (t == 0) return r.tok_n;
          if (t <= N) return node(t - 1);
          const int u = t - N - 1;
          if (u == 0) return r.tok_q;
          if (u == 1) return r.tok_task;
          if (TASK == GTOK_TASK_SHORTEST_PATH) {
            if (u == 2) return node(qu);
            if (u == 3) return node(qv);
          }
          return r.tok_p;
        };
        warp_pad_row(L, Lm, pad, X + o, attn + o, tok);
      }
      h0 = h1; h1 = h2;
      cs[0] = ns[0]; cs[1] = ns[1]; cd[0] = nd[0]; cd[1] = nd[1];
    }
    return;
  }
  for (int64_t item = (int64_t)blockIdx.x * kIcWarps + w; item < items; item += nwarps) {
    const int64_t row = spans == 1 ? item : item / spans;
    const int span = spans == 1 ? 0 : (int)(item - row * spans);
    const int64_t k = row / B;
    const int rr = (int)(row - k * B);
    const int Lm = __ldg(lmax + k);
    if ((int64_t)span * kIcSpan >= Lm && span > 0) continue;
    const int64_t base = __ldg(xoff + k);
    {
      const int64_t g = batch_idx ? __ldg(batch_idx + row) : row;
      int64_t s = 0;
      if (lane < 2) s = __ldg(b.node_off + g + lane);
      else if (lane < 4) s = __ldg(b.edge_off + g + lane - 2);
      const int64_t n0 = __shfl_sync(kFull, s, 0), n1 = __shfl_sync(kFull, s, 1);
      const int64_t e0 = __shfl_sync(kFull, s, 2), e1 = __shfl_sync(kFull, s, 3);
      const int N = (int)(n1 - n0), E = (int)(e1 - e0), E3 = 3 * E;
      const long long full = 3ll * E + N + qlen;
      const int L = (int)(full < r.max_len ? full : r.max_len);
      const int32_t* __restrict__ src = b.src + e0;
      const int32_t* __restrict__ dst = b.dst + e0;
      int qu = 0, qv = 0;
      if (TASK == GTOK_TASK_SHORTEST_PATH) {
        qu = __ldg(b.query_u + g);
        qv = __ldg(b.query_v + g);
      }
      const int64_t o = base + (int64_t)rr * Lm;
      if (spans == 1 && full <= row_cap) {
        // ---- build the untruncated row in shared memory, segment by segment -------------------------
        if (lane == 0) rb[0] = (StageT)r.tok_bos;
        for (int e = lane; e < E; e += kWarp) {
          rb[1 + 3 * e] = (StageT)node(__ldcs(src + e));
          rb[2 + 3 * e] = (StageT)node(__ldcs(dst + e));
          rb[3 + 3 * e] = (StageT)r.tok_e;
        }
        if (lane == 0) rb[E3 + 1] = (StageT)r.tok_n;
        for (int i = lane; i < N; i += kWarp) rb[E3 + 2 + i] = (StageT)node(i);
        if (lane == 0) {
          int t = E3 + 2 + N;
          rb[t++] = (StageT)r.tok_q;
          rb[t++] = (StageT)r.tok_task;
          if (TASK == GTOK_TASK_SHORTEST_PATH) { rb[t++] = (StageT)node(qu); rb[t++] = (StageT)node(qv); }
          rb[t] = (StageT)r.tok_p;
        }
        __syncwarp();
        const StageT* __restrict__ rbc = rb;
        warp_pad_row(L, Lm, pad, X + o, attn + o, [&](int p) -> int { return (int)rbc[p]; });
        __syncwarp();                                     // row buffer drained before the next row overwrites it
      } else {
        auto tok = [&](int p) -> int {                    // closed form per position (same as ibtt_serialize_kernel)
          if (p == 0) return r.tok_bos;
          if (p <= E3) {
            const int e = (int)(__umulhi((unsigned)(p - 1), 0xAAAAAAABu) >> 1);
            const int rem = (p - 1) - 3 * e;
            if (rem == 2) return r.tok_e;
            return node(rem == 0 ? __ldg(src + e) : __ldg(dst + e));
          }
          const int t = p - E3 - 1;
          if (t == 0) return r.tok_n;
          if (t <= N) return node(t - 1);
          const int u = t - N - 1;
          if (u == 0) return r.tok_q;
          if (u == 1) return r.tok_task;
          if (TASK == GTOK_TASK_SHORTEST_PATH) {
            if (u == 2) return node(qu);
            if (u == 3) return node(qv);
          }
          return r.tok_p;
        };
        if (spans == 1) {
          warp_pad_row(L, Lm, pad, X + o, attn + o, tok);
        } else {
          const int p0 = span * kIcSpan;
          const int plen = (span == spans - 1 || Lm - p0 < kIcSpan) ? Lm - p0 : kIcSpan;   // the last item takes the rest
          const int vlen = L - p0 < 0 ? 0 : (L - p0 < plen ? L - p0 : plen);
          warp_pad_row(vlen, plen, pad, X + o + p0, attn + o + p0, [&](int p) -> int { return tok(p0 + p); });
        }
      }
    }
  }
}

static int check_batch(const GtokGraphBatch* b, const char* who) {
  GTOK_REQUIRE(b != nullptr, GTOK_EINVAL, "%s: null batch", who);
  GTOK_REQUIRE(b->num_graphs >= 0, GTOK_EINVAL, "%s: negative graph count", who);
  if (b->num_graphs == 0) return GTOK_OK;
  GTOK_REQUIRE(b->node_off && b->edge_off, GTOK_EINVAL, "%s: null node_off/edge_off", who);
  return GTOK_OK;
}

static int check_rules(const GtokIbttRules* r, const GtokGraphBatch* b) {
  GTOK_REQUIRE(r != nullptr, GTOK_EINVAL, "ibtt: null rules");
  GTOK_REQUIRE(r->task == GTOK_TASK_CYCLE_CHECK || r->task == GTOK_TASK_SHORTEST_PATH, GTOK_EINVAL,
               "ibtt: unknown task %d", r->task);
  GTOK_REQUIRE(r->max_len >= 0, GTOK_EINVAL, "ibtt: negative max_len");
  GTOK_REQUIRE(r->node_tok_len == 0 || r->node_tok != nullptr, GTOK_EINVAL, "ibtt: null node_tok table");
  if (r->task == GTOK_TASK_SHORTEST_PATH && b->num_graphs > 0)
    GTOK_REQUIRE(b->query_u && b->query_v, GTOK_EINVAL, "ibtt: shortest_path needs query_u/query_v");
  return GTOK_OK;
}

}  // namespace gtok

extern "C" int gtok_ibtt_lengths(const GtokGraphBatch* batch, const GtokIbttRules* rules, int32_t* len_out,
                                 void* stream) {
  using namespace gtok;
  int rc = check_batch(batch, "ibtt_lengths");
  if (rc) return rc;
  rc = check_rules(rules, batch);
  if (rc) return rc;
  if (batch->num_graphs == 0) return GTOK_OK;
  GTOK_REQUIRE(len_out != nullptr, GTOK_EINVAL, "ibtt_lengths: null output");
  const int qlen = rules->task == GTOK_TASK_CYCLE_CHECK ? 5 : 7;
  const int64_t blocks = (batch->num_graphs + kIbttThreads - 1) / kIbttThreads;
  ibtt_lengths_kernel<<<(unsigned)blocks, kIbttThreads, 0, (cudaStream_t)stream>>>(*batch, rules->max_len, qlen, len_out);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

extern "C" int gtok_ibtt_serialize(const GtokGraphBatch* batch, const GtokIbttRules* rules, const int64_t* off,
                                   int64_t* tok_out, void* stream) {
  using namespace gtok;
  int rc = check_batch(batch, "ibtt_serialize");
  if (rc) return rc;
  rc = check_rules(rules, batch);
  if (rc) return rc;
  if (batch->num_graphs == 0) return GTOK_OK;
  GTOK_REQUIRE(off != nullptr && tok_out != nullptr, GTOK_EINVAL, "ibtt_serialize: null off/tok_out");
  // (src / dst may be null when the batch has no edge at all: the kernels never dereference them then)
  GTOK_REQUIRE((reinterpret_cast<uintptr_t>(tok_out) & 15) == 0, GTOK_EINVAL, "ibtt_serialize: tok_out must be 16-byte aligned");
  const int64_t work = (batch->num_graphs + (kIbttThreads / 32) - 1) / (kIbttThreads / 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (rules->task == GTOK_TASK_CYCLE_CHECK) {
    const int grid = persistent_grid(ibtt_serialize_kernel<GTOK_TASK_CYCLE_CHECK>, kIbttThreads, 0, work);
    ibtt_serialize_kernel<GTOK_TASK_CYCLE_CHECK><<<grid, kIbttThreads, 0, st>>>(*batch, *rules, off, tok_out);
  } else {
    const int grid = persistent_grid(ibtt_serialize_kernel<GTOK_TASK_SHORTEST_PATH>, kIbttThreads, 0, work);
    ibtt_serialize_kernel<GTOK_TASK_SHORTEST_PATH><<<grid, kIbttThreads, 0, st>>>(*batch, *rules, off, tok_out);
  }
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

extern "C" int gtok_ibtt_collate(const GtokGraphBatch* batch, const GtokIbttRules* rules, const int64_t* batch_idx,
                                 int64_t num_rows, int64_t batch_size, const int32_t* lmax, const int64_t* xoff,
                                 int64_t pad_id, int64_t* X, uint8_t* attn, void* stream) {
  using namespace gtok;
  int rc = check_batch(batch, "ibtt_collate");
  if (rc) return rc;
  rc = check_rules(rules, batch);
  if (rc) return rc;
  GTOK_REQUIRE(num_rows >= 0 && batch_size > 0 && batch_size < (1ll << 31), GTOK_EINVAL, "ibtt_collate: need num_rows >= 0, 0 < batch_size < 2^31");
  if (num_rows == 0) return GTOK_OK;
  GTOK_REQUIRE(batch->num_graphs > 0, GTOK_EINVAL, "ibtt_collate: rows requested from an empty batch");
  GTOK_REQUIRE(lmax && xoff && X && attn, GTOK_EINVAL, "ibtt_collate: null pointer");
  GTOK_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0, GTOK_EINVAL, "ibtt_collate: X must be 16-byte aligned");
  const int64_t nb = (num_rows + batch_size - 1) / batch_size;
  // shared memory: node table (up to 4096 entries) + one row buffer per warp; rows longer than the buffer use the closed form
  int maxid = rules->tok_bos;
  const int ids[6] = {rules->tok_e, rules->tok_n, rules->tok_q, rules->tok_p, rules->tok_task, rules->oov};
  for (int i = 0; i < 6; ++i) maxid = ids[i] > maxid ? ids[i] : maxid;
  const bool narrow = maxid < 65536 && rules->max_token_id > 0 && rules->max_token_id < 65536;
  const int esz = narrow ? 2 : 4;
  const int tbl_cap = rules->node_tok_len < 4096 ? rules->node_tok_len : 4096;
  // rows up to 4096 tokens (16-bit staging: 8 KB per warp, still 3 CTAs per SM) go through the row buffer when EVERY row fits;
  // workloads with longer rows keep the 2048-token buffer (more resident warps for the closed-form path they mostly take)
  int row_cap = rules->max_len + 8;
  if (row_cap > (narrow ? 4096 : 2048)) row_cap = 2048;
  row_cap = (row_cap + 7) & ~7;
  const size_t smem = (((size_t)tbl_cap * esz + 15) & ~(size_t)15) + (size_t)kIcWarps * row_cap * esz;
  cudaStream_t st = (cudaStream_t)stream;
  void (*k)(GtokGraphBatch, GtokIbttRules, const int64_t*, int64_t, int, int64_t, const int32_t*, const int64_t*, int64_t,
            int, int, int, int64_t*, uint8_t*);
  if (rules->task == GTOK_TASK_CYCLE_CHECK)
    k = narrow ? ibtt_collate_kernel<GTOK_TASK_CYCLE_CHECK, uint16_t> : ibtt_collate_kernel<GTOK_TASK_CYCLE_CHECK, int32_t>;
  else
    k = narrow ? ibtt_collate_kernel<GTOK_TASK_SHORTEST_PATH, uint16_t> : ibtt_collate_kernel<GTOK_TASK_SHORTEST_PATH, int32_t>;
  GTOK_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int spans = rules->max_len > 2 * kIcSpan ? (rules->max_len + kIcSpan - 1) / kIcSpan : 1;
  if (spans > 64) spans = 64;
  const int grid = persistent_grid(k, kIcThreads, smem, (num_rows * spans + kIcWarps - 1) / kIcWarps);
  k<<<grid, kIcThreads, smem, st>>>(*batch, *rules, batch_idx, num_rows, (int)batch_size, nb, lmax, xoff, pad_id, row_cap,
                                    tbl_cap, spans, X, attn);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}
