// gtok_sent_walk.cu - K3 for graphs of at most 64 nodes, TWO PHASES per warp (the default for this size class).
//
// Same function as gtok_sent_small.cu (Graph2TrailTokenizer.__call__ + remap_zinc_tokens + query append,
// trainer/train_agtt.py:171-273; spec oracle/sent_oracle.py), same outputs bit for bit, different work split.
// gtok_sent_small.cu walks AND emits in one thread-per-graph loop: ncu shows 9-16 of 32 lanes active per instruction
// (profiles/r1q_sent_er_full.txt, r2h): the bracket of every first visit is a data-dependent loop (collect the visited
// neighbours, order them by visit index, emit them) inside the sequential walk, so a warp pays max-over-lanes of every
// loop while most lanes idle.  But the WALK never looks at a bracket: its decisions need the adjacency row of the
// current node, the visited set and (restart policy "open first") the set of visited nodes that still have an
// unvisited neighbour.  So:
//   phase 1  walk     lane j <-> graph j, lockstep: one event (arrival at a node) per iteration, a dozen straight-line
//                     instructions + one pick.  Nothing is emitted: the event (node, first-visit flag, RESET flag) goes
//                     to a byte in shared memory, the visit index to vis[node].  The exact restart set is computed only
//                     when a lane is stuck (adjacency rows of the open-set candidates, L1 resident).
//   phase 2  emit     the warp's 32 event lists are flattened, lane <-> (graph, event): every lane rebuilds ITS event's
//                     bracket from the adjacency row and vis[] (neighbour visited earlier <=> vis[w] < vis[u]), counts
//                     its tokens, a segmented warp scan gives the position inside the graph's row, and the tokens go
//                     straight to HBM.  All lanes carry an event each (graphs are not padded to the longest walk), and
//                     consecutive lanes write consecutive runs of the same row.
// Random words, decisions and token order are exactly those of the oracle; tests compare both kernels with it.
#include "gtok_sent_common.cuh"

namespace gtok {

namespace {

struct WalkLayout {
  int o_ev, o_vis, o_ord, o_pref, o_gpos;   // byte offsets inside a warp slice
  int ev_cap;                               // events per graph (2 * max_nodes: every revisit is followed by a first visit)
  int warp_bytes;
};

constexpr int kEvFirst = 0x40, kEvReset = 0x80;

}  // namespace

static WalkLayout walk_layout(int nmax, bool labeled) {
  WalkLayout L;
  int o = 0;
  L.ev_cap = 2 * nmax;
  L.o_ev = o; o += 32 * L.ev_cap;             // u8 [e][lane]  node | first << 6 | reset << 7
  L.o_vis = o; o += 32 * nmax;                // u8 [v][lane]  node -> visit index
  L.o_ord = o; o += labeled ? 32 * nmax : 0;  // u8 [k][lane]  visit index -> node (edge labels of bracket members)
  o = (o + 15) & ~15;
  L.o_pref = o; o += 33 * 4;                  // int [33]      exclusive prefix of the 32 event counts
  L.o_gpos = o; o += 32 * 4;                  // int [32]      tokens written so far per graph
  L.warp_bytes = (o + 15) & ~15;
  return L;
}

template <bool LABELED>
__global__ void __launch_bounds__(kSmallThreads)
sent_walk_kernel(const GtokCsrBatch c, const GtokSentCfg cfg, const GtokChoice ch, const WalkLayout lay,
                 int32_t* __restrict__ len_out, int64_t* __restrict__ out, int64_t stride) {
  extern __shared__ __align__(16) unsigned char w_smem[];
  const int lane = lane_id();
  const int w = threadIdx.x >> 5;
  unsigned char* base = w_smem + (size_t)w * lay.warp_bytes;
  unsigned char* ev = base + lay.o_ev;
  unsigned char* vis = base + lay.o_vis;
  unsigned char* ord = base + lay.o_ord;
  int* pref = reinterpret_cast<int*>(base + lay.o_pref);
  int* gpos = reinterpret_cast<int*>(base + lay.o_gpos);
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
  __shared__ unsigned short s_remap[1024];
  if (remap) {
    for (int t = threadIdx.x; t < 1024; t += blockDim.x) s_remap[t] = (unsigned short)remap_zinc_tok(t, idx_off, node_off, edge_off);
    __syncthreads();
  }

  for (int64_t g0 = wid0 * 32; g0 < G; g0 += nw * 32) {
    const int Tc = (int)((G - g0 < 32) ? (G - g0) : 32);
    const bool valid = lane < Tc;
    const int64_t g = valid ? (c.order ? (int64_t)__ldg(c.order + g0 + lane) : g0 + lane) : g0;
    const int64_t my_n0 = __ldg(c.node_off + g);
    const int N = valid ? (int)(__ldg(c.node_off + g + 1) - my_n0) : 0;
    const unsigned long long* __restrict__ arow = adj64 + my_n0;

    // ---- phase 1: the walk (lane <-> graph) ----------------------------------------------------------------------
    const uint64_t gid = (uint64_t)(ch.graph_id_base + g);
    unsigned b0 = 0, b1 = 0, b2 = 0, b3 = 0, d = 0;
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
    const unsigned long long nodemask = N >= 64 ? ~0ull : ((1ull << N) - 1ull);
    unsigned long long visited = 0ull, open = 0ull;     // open: SUPERSET of the visited nodes with an unvisited neighbour
    int nvis = 0, ne = 0, cur = 0;
    bool done = N == 0, reset = false;
    unsigned long long a_nxt = 0ull;
    if (!done) {
      cur = pick(N);
      a_nxt = __ldg(arow + cur);
    }
    while (!__all_sync(kFull, done)) {
      if (!done) {
        const unsigned long long a_cur = a_nxt;
        const unsigned long long bit = 1ull << cur;
        const bool first = !(visited & bit);
        if (first) {
          vis[cur * 32 + lane] = (unsigned char)nvis;
          if (LABELED) ord[nvis * 32 + lane] = (unsigned char)cur;
          ++nvis;
          visited |= bit;
        }
        ev[ne * 32 + lane] = (unsigned char)(cur | (first ? kEvFirst : 0) | (reset ? kEvReset : 0));
        ++ne;
        const unsigned long long C = a_cur & ~visited;
        if (first && C) open |= bit;
        if (!C && nvis == N) {
          done = true;
        } else {
          unsigned long long sel = C;
          reset = !C;
          if (!C) {                                       // stuck: restart at an open node, else at any unvisited node
            if (!restart_unvisited) {
              unsigned long long cand = open & ~bit, keep = 0ull;      // (cur itself has no unvisited neighbour left)
              while (cand) {                              // exact open set: candidates that still see an unvisited node
                const int u = __ffsll((long long)cand) - 1;
                cand &= cand - 1;
                if (__ldg(arow + u) & ~visited) keep |= 1ull << u;
              }
              open = keep;
              sel = keep;
            }
            if (!sel) sel = ~visited & nodemask;
          }
          const int nsel = __popcll(sel);
          const int kth = pick(nsel);                     // one word per event, also when there is no choice
          const int nxt = nsel == 1 ? __ffsll((long long)sel) - 1 : select64(sel, kth);
          a_nxt = __ldg(arow + nxt);
          cur = nxt;
        }
      }
    }
    __syncwarp();

    // ---- phase 2: emission (lane <-> (graph, event)) -------------------------------------------------------------
    int incl = ne;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
      const int up = __shfl_up_sync(kFull, incl, dd);
      if (lane >= dd) incl += up;
    }
    pref[lane + 1] = incl;
    if (lane == 0) pref[0] = 0;
    gpos[lane] = 1;                                       // SOS
    if (valid && Lcap > 0) out[g * stride] = remap ? (int64_t)remap_zinc_tok(SOS, idx_off, node_off, edge_off) : (int64_t)SOS;
    __syncwarp();
    const int T = __shfl_sync(kFull, incl, 31);
    auto tokv = [&](int t) -> long long {
      if (!remap) return t;
      return (unsigned)t < 1024u ? (int)s_remap[t] : remap_zinc_tok(t, idx_off, node_off, edge_off);
    };
    for (int t0 = 0; t0 < T; t0 += 32) {
      const int t = t0 + lane;
      const bool act = t < T;
      // graph of item t: largest gi with pref[gi] <= t (pref is non-decreasing; graphs without events are skipped)
      int gi = 0;
      if (act) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1)
          if (gi + s <= 31 && pref[gi + s] <= t) gi += s;
      }
      const int e = act ? t - pref[gi] : 0;
      const int64_t gg = __shfl_sync(kFull, g, gi);
      const int64_t n0g = __shfl_sync(kFull, my_n0, gi);
      int cnt = 0, k = 0, u = 0, m = 0, rp_u = 0, pre_lab = -1;
      bool first = false, rs = false;
      unsigned long long V = 0ull, a_u = 0ull;
      if (act) {
        const unsigned evb = ev[e * 32 + gi];
        u = evb & 63;
        first = (evb & kEvFirst) != 0;
        rs = (evb & kEvReset) != 0;
        k = vis[u * 32 + gi];
        const int prev = (e > 0 && !rs) ? (ev[(e - 1) * 32 + gi] & 63) : -1;
        if (first || (LABELED && prev >= 0)) a_u = __ldg(adj64 + n0g + u);
        if (LABELED && (first || prev >= 0)) rp_u = __ldg(c.rowptr + n0g + u);
        if (first) {
          unsigned long long rest = a_u;
          while (rest) {                                  // neighbours visited before u, in visit-index space
            const int wv = __ffsll((long long)rest) - 1;
            rest &= rest - 1;
            const int kv = vis[wv * 32 + gi];
            if (kv < k) V |= 1ull << kv;
          }
          if (prev >= 0) V &= ~(1ull << vis[prev * 32 + gi]);      // the trail predecessor is not a bracket member
          m = __popcll(V);
        }
        if (LABELED && prev >= 0)                         // label of the trail edge prev -> u (rows are symmetric)
          pre_lab = edge_off + __ldg(c.eattr + rp_u + __popcll(a_u & ((1ull << prev) - 1ull))) + cfg.edge_label_shift;
        cnt = (rs ? 1 : 0) + (pre_lab >= 0 ? 1 : 0) + 1 + ((first && LABELED) ? 1 : 0) + (m ? 2 + m * (LABELED ? 2 : 1) : 0);
      }
      // position inside the graph's row: segmented inclusive scan (items of one graph are consecutive lanes)
      int sc = cnt;
#pragma unroll
      for (int dd = 1; dd < 32; dd <<= 1) {
        const int up = __shfl_up_sync(kFull, sc, dd);
        if (lane >= dd) sc += up;
      }
      int head = act ? pref[gi] - t0 : 0;                 // lane of the graph's first item in this round (< 0: earlier round)
      if (head < 0) head = 0;
      const int before = __shfl_sync(kFull, sc, head > 0 ? head - 1 : 0);
      int p = 0;
      if (act) p = gpos[gi] + sc - cnt - (head > 0 ? before : 0);
      __syncwarp();
      if (act && (t + 1 == pref[gi + 1] || lane == 31)) gpos[gi] = p + cnt;     // last item of the graph in this round
      if (act) {
        int64_t* __restrict__ row = out + gg * stride;
        auto put = [&](int tk) {
          if (p < Lcap) __stcs(row + p, (int64_t)tokv(tk));
          ++p;
        };
        if (rs) put(RESET);
        if (pre_lab >= 0) put(pre_lab);
        put(idx_off + k);
        if (first && LABELED) put(node_off + __ldg(c.x + n0g + u));
        if (m) {
          put(LADJ);
          while (V) {
            int kv;
            if (bracket_desc) { kv = 63 - __clzll((long long)V); V &= ~(1ull << kv); }
            else { kv = __ffsll((long long)V) - 1; V &= V - 1; }
            if (LABELED) {
              const int wn = ord[kv * 32 + gi];
              put(edge_off + __ldg(c.eattr + rp_u + __popcll(a_u & ((1ull << wn) - 1ull))) + cfg.edge_label_shift);
            }
            put(idx_off + kv);
          }
          put(RADJ);
        }
      }
      __syncwarp();
    }
    if (valid) {
      int L = gpos[lane];                                 // tokens before EOS
      int64_t* __restrict__ row = out + g * stride;
      if (L < Lcap) row[L] = tokv(EOS);
      L = (L + 1 < Lcap) ? L + 1 : Lcap;
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

// Called by gtok_sent_tokenize for batches whose graphs all have <= 64 nodes (GTOK_SENT_SMALL=1 selects the one-phase
// kernel of gtok_sent_small.cu instead).
int sent_walk_dispatch(const GtokCsrBatch* c, const GtokSentCfg* cfg, const GtokChoice* ch, int32_t* len_out, int64_t* out,
                       int64_t stride, cudaStream_t st, bool* handled) {
  *handled = false;
  if (c->max_nodes > 64 || c->adj64 == nullptr || c->adj_words > 1) return GTOK_OK;
  const WalkLayout lay = walk_layout(c->max_nodes, cfg->labeled != 0);
  const size_t smem = (size_t)kSmallWarps * lay.warp_bytes;
  auto k = cfg->labeled ? sent_walk_kernel<true> : sent_walk_kernel<false>;
  GTOK_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t warps = (c->num_graphs + 31) / 32;
  const int grid = persistent_grid(k, kSmallThreads, smem, (warps + kSmallWarps - 1) / kSmallWarps);
  k<<<grid, kSmallThreads, smem, st>>>(*c, *cfg, *ch, lay, len_out, out, stride);
  GTOK_CUDA_TRY(cudaGetLastError());
  *handled = true;
  return GTOK_OK;
}

}  // namespace gtok
