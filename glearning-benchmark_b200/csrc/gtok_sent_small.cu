// gtok_sent_small.cu - K3 for small graphs (<= 64 nodes): ONE THREAD PER GRAPH.
//
// Same function as gtok_sent.cu (Graph2TrailTokenizer.__call__ + remap_zinc_tokens + query
// append, trainer/train_agtt.py:171-273; spec oracle/sent_oracle.py), different mapping.
// The warp-per-graph kernel spends ~11k warp instructions per graph because every walk step
// is scalar work replicated over 32 lanes (profiles/: 2.5 % of the HBM roofline, issue bound).
// For graphs of at most 64 nodes the whole walk state is a handful of 64-bit masks, so a
// THREAD can own a graph:
//   adjacency   64-bit row mask per node (GtokCsrBatch.adj64, built once per dataset), read
//               through L1 straight from HBM: 8 B per node instead of rowptr+col, no staging,
//               so shared memory only holds the visit-index bytes and occupancy stays high;
//   visited     one 64-bit register; candidates = adj[cur] & ~visited; k-th set bit = the pick;
//   bracket     members = adj[cur] & visited & ~prev; their visit indices are OR-ed into a
//               64-bit mask in visit-index space, whose set bits enumerate ascending - the
//               sorted neighbourhood set without a sort;
//   restart     visited nodes with an unvisited neighbour = (OR of adj over unvisited nodes) & visited;
//   output      tokens leave in pairs as 16-byte stores straight to the graph's row.
// A warp owns 32 consecutive graphs, lane j walks graph j; lanes of a warp touch neighbouring
// rows of adj64 / x / rowptr, so the L1 lines they pull are shared.
// Random words and every decision are exactly those of the oracle (Philox-4x32-10 keyed by the
// seed, counter = (global graph id, decision >> 2), or the explicit choice stream).
#include "gtok_sent_common.cuh"

namespace gtok {

namespace {

struct SmallLayout {
  int o_vis, o_rem, o_ord, o_xs, o_ea, o_ring;   // byte offsets inside a warp slice
  int warp_bytes;
};

constexpr int kRemapTable = 1024;  // tokens below this go through a shared-memory remap table

constexpr int kRing = 64;      // ring slots per lane (tokens leave 16 at a time = four full 32-byte sectors)

}  // namespace

static SmallLayout small_layout(int nmax, int zmax, bool labeled, bool stage_labels, int ring_elem) {
  SmallLayout L;
  int o = 0;
  L.o_vis = o; o += 32 * nmax;                 // u8  [v][lane]  node -> visit index
  L.o_rem = o; o += 32 * nmax;                 // u8  [v][lane]  unvisited neighbours left (visited nodes only)
  L.o_ord = o; o += labeled ? 32 * nmax : 0;   // u8  [k][lane]  visit index -> node
  L.o_xs = o; o += (labeled && stage_labels) ? 32 * nmax : 0;   // u8 [v][lane]  node label
  L.o_ea = o; o += (labeled && stage_labels) ? 32 * zmax : 0;   // u8 [e][lane]  edge label per CSR entry
  o = (o + 15) & ~15;
  L.o_ring = o; o += kRing * 32 * ring_elem;   // [slot][lane] pending tokens (16-bit when every token fits)
  L.warp_bytes = (o + 15) & ~15;
  return L;
}

// One lockstep loop per warp (lane j <-> graph j); every iteration is one EVENT (arrival at a node) and
// takes exactly one pick, so the Philox block refresh falls on the same iteration for all lanes.  All 32
// lanes stay in the loop until the slowest graph is done, which gives two warp-synchronous service points
// per iteration.  History (profiles/r1g..r1n): the first thread-per-graph version ran 1900 warp instructions
// per graph at 7 of 32 lanes active; what changed:
//   * emit() is a shared-memory store into a per-lane ring ([slot][lane], conflict free) at each of its ten
//     call sites instead of a bounds check + remap + parity branch + global store;
//   * flush point: when ANY lane has 32 tokens pending, EVERY lane with 16 or more stores 16 (eight 16-byte
//     stores = four full sectors) - the flush code runs a handful of times per 32 graphs with most lanes
//     active (lane-by-lane flushing was 31 % of all instructions);
//   * restart in O(1): a stuck lane needs the visited nodes that still have an unvisited neighbour.  Every
//     visited node keeps a counter of unvisited neighbours (u8 in shared memory) and a bit in `open`; a node's
//     first visit already enumerates its visited neighbours (trail predecessor + bracket members), so each
//     edge costs one decrement, and the restart set is just `open` - no adjacency scan (the scan, per lane or
//     served by the whole warp, was 15-20 % of instructions and the top stall);
//   * one select64 site (the pick among unvisited neighbours and the restart pick share it); a bracket with
//     one member (the common case) skips the visit-index ordering.
//   * labeled graphs: node and edge labels of the lane's graph are copied into shared memory as bytes before the
//     walk (independent loads, issued back to back); the walk itself then never waits for HBM on a label (the
//     dependent eattr / x loads were 30 % of all stall samples, profiles/r1p).
template <bool LABELED, bool STAGE_LABELS, typename RingT>
__global__ void __launch_bounds__(kSmallThreads)
sent_small_kernel(const GtokCsrBatch c, const GtokSentCfg cfg, const GtokChoice ch, const SmallLayout lay,
                  int32_t* __restrict__ len_out, int64_t* __restrict__ out, int64_t stride) {
  extern __shared__ __align__(16) unsigned char s_smem[];
  const int lane = lane_id();
  const int w = threadIdx.x >> 5;
  unsigned char* base = s_smem + (size_t)w * lay.warp_bytes;
  unsigned char* vis = base + lay.o_vis + lane;
  unsigned char* rem = base + lay.o_rem + lane;
  unsigned char* ord = base + lay.o_ord + lane;
  unsigned char* xs = base + lay.o_xs + lane;
  unsigned char* eas = base + lay.o_ea + lane;
  RingT* ring = reinterpret_cast<RingT*>(base + lay.o_ring) + lane;
  const unsigned long long* __restrict__ adj64 = reinterpret_cast<const unsigned long long*>(c.adj64);

  const int64_t G = c.num_graphs;
  const int64_t wid0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int Lcap = cfg.max_length;
  const unsigned k0 = (unsigned)(ch.seed & 0xffffffffu), k1 = (unsigned)(ch.seed >> 32);
  const bool remap = cfg.remap_zinc != 0;
  const bool restart_unvisited = cfg.restart_policy == GTOK_SENT_RESTART_UNVISITED;
  const bool bracket_desc = cfg.bracket_order == GTOK_SENT_BRACKET_VISIT_DESC;
  const int idx_off = cfg.idx_offset, node_off = cfg.node_idx_offset, edge_off = cfg.edge_idx_offset;
  __shared__ unsigned short s_remap[kRemapTable];
  if (remap) {
    for (int t = threadIdx.x; t < kRemapTable; t += blockDim.x)
      s_remap[t] = (unsigned short)remap_zinc_tok(t, idx_off, node_off, edge_off);
    __syncthreads();
  }

  for (int64_t g0 = wid0 * 32; g0 < G; g0 += nw * 32) {
    const int Tc = (int)((G - g0 < 32) ? (G - g0) : 32);
    const bool valid = lane < Tc;
    const int64_t g = valid ? (c.order ? (int64_t)__ldg(c.order + g0 + lane) : g0 + lane) : g0;
    const int64_t my_n0 = __ldg(c.node_off + g);
    const int N = valid ? (int)(__ldg(c.node_off + g + 1) - my_n0) : 0;
    const unsigned long long* __restrict__ arow = adj64 + my_n0;
    int64_t* __restrict__ row = out + g * stride;                // rows are 16-byte aligned (even stride, aligned base)

    int pos = 0, flushed = 0;                                    // tokens emitted / tokens already in HBM (multiple of 16)
    auto emit = [&](int t) {
      if (pos < Lcap) ring[(pos & (kRing - 1)) * 32] = (RingT)t;  // tokens past max_length are never stored
      ++pos;
    };
    auto tok = [&](int p) -> long long {                         // pending token p, post-processed
      const int t = (int)ring[(p & (kRing - 1)) * 32];
      if (!remap) return t;
      return (unsigned)t < (unsigned)kRemapTable ? (int)s_remap[t] : remap_zinc_tok(t, idx_off, node_off, edge_off);
    };
    auto flush16 = [&]() {                                       // 16 tokens = four full 32-byte sectors of the row
#pragma unroll
      for (int i = 0; i < 16; i += 2) st_stream_i64x2(row + flushed + i, tok(flushed + i), tok(flushed + i + 1));
      flushed += 16;
    };
    auto flush_guard = [&]() {                                   // inside a long bracket: never let the ring wrap (rare)
      if (pos - flushed >= kRing - 8 && flushed + 16 <= Lcap) flush16();
    };
    // labels into shared memory (bytes): x[0..N), eattr[rp0 .. rp0+nnz)
    int rp0 = 0;
    if (LABELED && STAGE_LABELS) {
      rp0 = __ldg(c.rowptr + my_n0);
      const int nnz = valid ? __ldg(c.rowptr + my_n0 + N) - rp0 : 0;
      const int32_t* __restrict__ gx = c.x + my_n0;
      const int32_t* __restrict__ ge = c.eattr + rp0;
#pragma unroll 4
      for (int i = 0; i < N; ++i) xs[i * 32] = (unsigned char)__ldg(gx + i);
#pragma unroll 4
      for (int e = 0; e < nnz; ++e) eas[e * 32] = (unsigned char)__ldg(ge + e);
    }
    // label of CSR entry (cur -> v): rank of v inside cur's row; rp = rowptr of cur, fetched together with its adjacency
    auto edge_label = [&](int rp, int v, unsigned long long au) -> int {
      const unsigned long long below = au & ((1ull << v) - 1ull);
      const int e = rp + __popcll(below);
      const int a = (LABELED && STAGE_LABELS) ? (int)eas[(e - rp0) * 32] : __ldg(c.eattr + e);
      return edge_off + a + cfg.edge_label_shift;
    };
    auto node_label = [&](int v) -> int {
      return node_off + ((LABELED && STAGE_LABELS) ? (int)xs[v * 32] : __ldg(c.x + my_n0 + v));
    };

    // random words
    const uint64_t gid = (uint64_t)(ch.graph_id_base + g);
    unsigned b0 = 0, b1 = 0, b2 = 0, b3 = 0;
    unsigned d = 0;
    int64_t s_off = 0;
    unsigned s_len = 0;
    if (ch.stream && valid) {
      s_off = __ldg(ch.stream_off + g);
      s_len = (unsigned)(__ldg(ch.stream_off + g + 1) - s_off);
    }
    auto pick = [&](int cnt) -> int {
      unsigned r;
      if (ch.stream) {
        r = d < s_len ? __ldg(ch.stream + s_off + d) : 0u;
      } else {
        if ((d & 3u) == 0u) philox10((unsigned)gid, (unsigned)(gid >> 32), d >> 2, 0u, k0, k1, b0, b1, b2, b3);
        const unsigned w01 = (d & 1u) ? b1 : b0, w23 = (d & 1u) ? b3 : b2;
        r = (d & 2u) ? w23 : w01;
      }
      ++d;
      return (int)__umulhi(r, (unsigned)cnt);
    };

    if (valid) emit(SOS);
    const unsigned long long nodemask = N >= 64 ? ~0ull : ((1ull << N) - 1ull);
    unsigned long long visited = 0ull, open = 0ull;     // open: visited nodes that still have an unvisited neighbour
    int nvis = 0, prev = -1, cur = 0;
    bool done = N == 0 || pos >= Lcap;
    if (!done) cur = pick(N);
    unsigned long long a_nxt = 0ull;                    // adjacency row of the node we are about to arrive at
    int rp_nxt = 0;                                      // (fetched as soon as the node is chosen)
    if (!done) {
      a_nxt = __ldg(arow + cur);
      if (LABELED) rp_nxt = __ldg(c.rowptr + my_n0 + cur);
    }
    while (true) {
      // ---- arrive at cur ---------------------------------------------------------------------
      if (!done) {
        const unsigned long long a_cur = a_nxt;
        const int rp_cur = rp_nxt;
        if (!((visited >> cur) & 1ull)) {
          const int k = nvis++;
          vis[cur * 32] = (unsigned char)k;
          if (LABELED) ord[k * 32] = (unsigned char)cur;
          emit(idx_off + k);
          if (LABELED) emit(node_label(cur));
          unsigned long long Mall = a_cur & visited;              // visited neighbours = bracket members + the trail
          if (prev >= 0) Mall &= ~(1ull << prev);                 // predecessor (already accounted for when we left it)
          visited |= 1ull << cur;
          const unsigned long long C0 = a_cur & ~visited;
          rem[cur * 32] = (unsigned char)__popcll(C0);
          if (C0) open |= 1ull << cur;
          // Mall: bracket members (visited neighbours except the trail predecessor).  Each one loses cur as an
          // unvisited neighbour; their visit indices, OR-ed into V, enumerate ascending without a sort.  The loops
          // run on 32-bit halves (a 64-bit find-first-set / clear-lowest pair costs twice the instructions).
          if (Mall) {
            emit(LADJ);
            auto close_member = [&](int u) -> int {             // counter update of member u -> its visit index
              const int r = (int)rem[u * 32] - 1;               // cur is no longer an unvisited neighbour of u
              rem[u * 32] = (unsigned char)r;
              if (r == 0) open &= ~(1ull << u);
              return (int)vis[u * 32];
            };
            if ((Mall & (Mall - 1)) == 0ull) {                  // one member (the common case): nothing to order
              const int u = __ffsll((long long)Mall) - 1;
              const int jv = close_member(u);
              if (LABELED) emit(edge_label(rp_cur, u, a_cur));
              emit(idx_off + jv);
            } else {
              unsigned vlo = 0u, vhi = 0u;
              unsigned mlo = (unsigned)Mall, mhi = (unsigned)(Mall >> 32);
              while (mlo) {
                const int u = __ffs((int)mlo) - 1;
                mlo &= mlo - 1;
                const int jv = close_member(u);
                if (jv < 32) vlo |= 1u << jv; else vhi |= 1u << (jv - 32);
              }
              while (mhi) {
                const int u = 32 + __ffs((int)mhi) - 1;
                mhi &= mhi - 1;
                const int jv = close_member(u);
                if (jv < 32) vlo |= 1u << jv; else vhi |= 1u << (jv - 32);
              }
              const bool longb = __popc(vlo) + __popc(vhi) > 10;  // only a long bracket can outrun the ring
              if (bracket_desc) {                                 // GTOK_SENT_BRACKET_VISIT_DESC: highest visit index first
                while (vhi) {
                  const int b = 31 - __clz((int)vhi);
                  vhi &= ~(1u << b);
                  if (LABELED) emit(edge_label(rp_cur, (int)ord[(32 + b) * 32], a_cur));
                  emit(idx_off + 32 + b);
                  if (longb) flush_guard();
                }
                while (vlo) {
                  const int b = 31 - __clz((int)vlo);
                  vlo &= ~(1u << b);
                  if (LABELED) emit(edge_label(rp_cur, (int)ord[b * 32], a_cur));
                  emit(idx_off + b);
                  if (longb) flush_guard();
                }
              }
              while (vlo) {
                const int jv = __ffs((int)vlo) - 1;
                vlo &= vlo - 1;
                if (LABELED) emit(edge_label(rp_cur, (int)ord[jv * 32], a_cur));
                emit(idx_off + jv);
                if (longb) flush_guard();
              }
              while (vhi) {
                const int jv = 32 + __ffs((int)vhi) - 1;
                vhi &= vhi - 1;
                if (LABELED) emit(edge_label(rp_cur, (int)ord[jv * 32], a_cur));
                emit(idx_off + jv);
                if (longb) flush_guard();
              }
            }
            emit(RADJ);
          }
        } else {
          emit(idx_off + (int)vis[cur * 32]);                    // head of a new segment at a visited node
        }
        // ---- leave: one pick per event -------------------------------------------------------------
        const unsigned long long C = a_cur & ~visited;
        if (!C && nvis == N) {
          done = true;
        } else {
          unsigned long long sel = C;
          if (!C) {                                               // stuck: restart at an open node, else at any unvisited node
            sel = (open && !restart_unvisited) ? open : (~visited & nodemask);
            emit(RESET);
          }
          const int nsel = __popcll(sel);
          const int kth = pick(nsel);                             // one word per event, also when there is no choice
          const int nxt = nsel == 1 ? __ffsll((long long)sel) - 1 : select64(sel, kth);
          a_nxt = __ldg(arow + nxt);
          if (LABELED) rp_nxt = __ldg(c.rowptr + my_n0 + nxt);
          if (C) {                                                // trail step: nxt stops being an unvisited neighbour of cur
            rem[cur * 32] = (unsigned char)(nsel - 1);            // (nsel == |C| == the counter's current value)
            if (nsel == 1) open &= ~(1ull << cur);
            if (LABELED) emit(edge_label(rp_cur, nxt, a_cur));
          }
          prev = C ? cur : -1;
          cur = nxt;
          if (pos >= Lcap) done = true;
        }
      }
      // ---- flush point (warp-synchronous) ----------------------------------------------------------
      if (__any_sync(kFull, pos - flushed >= 32)) {
        if (pos - flushed >= 16 && flushed + 16 <= Lcap) flush16();
        if (pos - flushed >= 32 && flushed + 16 <= Lcap) flush16();   // a labeled event can add more than 16 tokens:
      }                                                                   // pending must be below 32 when the next one starts
      if (__all_sync(kFull, done)) break;
    }
    if (valid) {
      if (!cfg.no_eos) emit(EOS);
      const int L0 = pos < Lcap ? pos : Lcap;
      // drain: whole pairs, then the unpaired last token
      for (int p = flushed; p + 1 < L0; p += 2) st_stream_i64x2(row + p, tok(p), tok(p + 1));
      if (L0 & 1) st_stream_i64(row + L0 - 1, tok(L0 - 1));
      int L = L0;
      if (cfg.append_query) {
        row[L] = idx_off + N;
        row[L + 1] = idx_off + __ldg(c.query_u + g);
        row[L + 2] = idx_off + __ldg(c.query_v + g);
        L += 3;
      }
      len_out[g] = L;
    }
    __syncwarp();
  }
}

// Called by gtok_sent_tokenize for batches whose graphs all have <= 64 nodes.
int sent_small_dispatch(const GtokCsrBatch* c, const GtokSentCfg* cfg, const GtokChoice* ch, int32_t* len_out,
                        int64_t* out, int64_t stride, cudaStream_t st, bool* handled) {
  *handled = false;
  if (c->max_nodes > 64 || (stride & 1) || c->adj64 == nullptr || c->adj_words > 1) return GTOK_OK;
  // largest token before the remap: node index / node label / edge label (needs the label bound)
  long long tmax = (long long)cfg->idx_offset + 64;
  bool narrow = tmax < 65536;
  if (cfg->labeled) {
    const long long lab = c->max_label > 0 ? c->max_label : (1ll << 40);
    const long long a = (long long)cfg->node_idx_offset + lab, b = (long long)cfg->edge_idx_offset + lab + cfg->edge_label_shift;
    narrow = narrow && a < 65536 && b < 65536 && cfg->node_idx_offset >= 0 && cfg->edge_idx_offset + cfg->edge_label_shift >= 0;
  }
  narrow = narrow && cfg->idx_offset >= 0;
  // labels travel as bytes through shared memory when their bound allows it and the rows are molecule sized
  const bool stage_labels = cfg->labeled && c->max_label > 0 && c->max_label < 256 && c->max_nnz <= 256;
  const SmallLayout lay = small_layout(c->max_nodes, c->max_nnz, cfg->labeled != 0, stage_labels, narrow ? 2 : 4);
  const size_t smem = (size_t)kSmallWarps * lay.warp_bytes;
  void (*k)(GtokCsrBatch, GtokSentCfg, GtokChoice, SmallLayout, int32_t*, int64_t*, int64_t);
  if (cfg->labeled) {
    if (stage_labels) k = narrow ? sent_small_kernel<true, true, uint16_t> : sent_small_kernel<true, true, int32_t>;
    else k = narrow ? sent_small_kernel<true, false, uint16_t> : sent_small_kernel<true, false, int32_t>;
  } else {
    k = narrow ? sent_small_kernel<false, false, uint16_t> : sent_small_kernel<false, false, int32_t>;
  }
  GTOK_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t warps = (c->num_graphs + 31) / 32;
  const int grid = persistent_grid(k, kSmallThreads, smem, (warps + kSmallWarps - 1) / kSmallWarps);
  k<<<grid, kSmallThreads, smem, st>>>(*c, *cfg, *ch, lay, len_out, out, stride);
  GTOK_CUDA_TRY(cudaGetLastError());
  *handled = true;
  return GTOK_OK;
}

}  // namespace gtok
