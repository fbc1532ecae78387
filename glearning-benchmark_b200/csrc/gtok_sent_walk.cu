// gtok_sent_walk.cu - K3 for graphs of at most 64 nodes as TWO kernels: the WALK and the EMISSION.
//
// Same function as gtok_sent_small.cu (Graph2TrailTokenizer.__call__ + remap_zinc_tokens + query append,
// trainer/train_agtt.py:171-273; spec oracle/sent_oracle.py), same outputs bit for bit, different work split.
// gtok_sent_small.cu walks AND emits in one thread-per-graph loop; ncu (profiles/r1q_sent_er_full.txt, r2h) shows 9-16 of
// 32 lanes active per instruction: ten emit sites, a token ring and its flush, and the enumeration of every bracket in
// visit order all sit inside the sequential walk, and every lane stores its own row 16 bytes at a time.  But the walk's
// DECISIONS never look at a token.  So:
//   sent_walk_kernel   thread per graph, 32 graphs of a warp in lockstep, one arrival and one pick per iteration (the Philox
//                      refresh falls on the same iteration for all lanes).  State: two masks in registers, visit index and
//                      open-neighbour counter per node as bytes in shared memory (3 KB per warp at 49 nodes).  A first
//                      visit appends ONE 8-byte item to the graph's LOG (the batch's workspace, slot = node_off[g] + visit
//                      index): the bracket as a mask in visit-index space, the node, how it was reached, and the
//                      position of its first token (a running sum of closed-form counts, which also gives max_length's
//                      early stop).  Nothing else is stored; row lengths go to len_out.
//   sent_emit_kernel   warp per row, lane per item: every lane expands its item into the warp's row buffer in shared
//                      memory (8- or 16-bit tokens; set bits of the bracket mask enumerate in visit order, the sorted
//                      neighbourhood set without a sort), then the warp streams the row: 32 lanes x 16 bytes = 64
//                      consecutive int64 tokens per store instruction, ZINC remap from a shared-memory table on the way
//                      out.  The same kernel writes the row straight into its padded batch (collate_fn,
//                      train_agtt.py:276-302): X row + attention-mask row at xoff[batch] + r * Lmax, pad beyond the
//                      length - the int64 rows at stride max_length never exist (gtok_sent_emit_collate).
// Random words, decisions and token order are exactly those of the oracle; tests compare all kernels with it.
#include "gtok_sent_common.cuh"


namespace gtok {

namespace {

constexpr int kPreReset = 0x80, kPreHead = 0x40;   // item reached by a restart / through a revisited head (low 6 bits: its visit index)
constexpr int kOrdEnd = 0xFF;                      // node byte of the slot after the last item of an early-stopped walk
constexpr int kItemMax = 2 + 2 + 2 + 2 * 63;       // most tokens one item can add (reset, head, edge label, index, node label,
                                                   // LADJ, 63 labeled members, RADJ) rounded up
constexpr int kEmitThreads = 256;
constexpr int kEmitWarps = kEmitThreads / 32;

template <typename M> struct Mk;
template <> struct Mk<unsigned> {
  static __device__ __forceinline__ int popc(unsigned m) { return __popc(m); }
  static __device__ __forceinline__ int first(unsigned m) { return __ffs((int)m) - 1; }
  static __device__ __forceinline__ int select(unsigned m, int k) { return select32(m, k); }
  static __device__ __forceinline__ unsigned bit(int b) { return 1u << b; }
  template <typename F> static __device__ __forceinline__ void each(unsigned m, F f) {
    while (m) { const int b = __ffs((int)m) - 1; m &= m - 1; f(b); }
  }
};
template <> struct Mk<unsigned long long> {
  static __device__ __forceinline__ int popc(unsigned long long m) { return __popcll(m); }
  static __device__ __forceinline__ int first(unsigned long long m) {
    const unsigned lo = (unsigned)m;
    return lo ? __ffs((int)lo) - 1 : 31 + __ffs((int)(unsigned)(m >> 32));
  }
  static __device__ __forceinline__ int select(unsigned long long m, int k) { return select64(m, k); }
  static __device__ __forceinline__ unsigned long long bit(int b) { return 1ull << b; }
  template <typename F> static __device__ __forceinline__ void each(unsigned long long m, F f) {   // 32-bit halves: a 64-bit
    unsigned lo = (unsigned)m, hi = (unsigned)(m >> 32);                                           // find-first / clear-lowest
    while (lo) { const int b = __ffs((int)lo) - 1; lo &= lo - 1; f(b); }                           // pair costs twice as much
    while (hi) { const int b = 31 + __ffs((int)hi); hi &= hi - 1; f(b); }
  }
};

// The walk of the warp's 32 graphs on masks of type MaskT (32-bit when every graph of the warp has at most 32 nodes).
template <typename MaskT, bool LABELED>
__device__ __forceinline__ void walk_body(const GtokCsrBatch& c, const GtokSentCfg& cfg, const GtokChoice& ch,
                                          unsigned char* vis, unsigned char* rem, uint2* __restrict__ log_lo,
                                          unsigned* __restrict__ log_hi, int32_t* __restrict__ len_out, bool valid,
                                          int64_t g, int64_t my_n0, int N) {
  using K = Mk<MaskT>;
  const unsigned long long* __restrict__ arow = reinterpret_cast<const unsigned long long*>(c.adj64) + my_n0;
  const int Lcap = cfg.max_length;
  const unsigned k0 = (unsigned)(ch.seed & 0xffffffffu), k1 = (unsigned)(ch.seed >> 32);
  const bool restart_unvisited = cfg.restart_policy == GTOK_SENT_RESTART_UNVISITED;
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
  const MaskT nodemask = N >= (int)(8 * sizeof(MaskT)) ? (MaskT)~(MaskT)0 : (MaskT)(K::bit(N) - 1);
  MaskT visited = 0, open = 0;                      // open: visited nodes that still have an unvisited neighbour
  int nvis = 0, cur = 0, prev = -1, pre = 0;
  int pos = 1;                                      // tokens so far (SOS)
  bool done = N == 0;
  MaskT a_nxt = 0;                                  // adjacency row of the node we are about to arrive at (fetched as soon as
  if (!done) {                                      //  the node is chosen)
    cur = pick(N);
    a_nxt = (MaskT)__ldg(arow + cur);
  }
  while (!__all_sync(kFull, done)) {
    if (!done) {
      const MaskT a_cur = a_nxt;
      const MaskT bit = K::bit(cur);
      if (!(visited & bit)) {                       // first visit: one item
        const int k = nvis++;
        vis[cur * 32] = (unsigned char)k;
        MaskT Mb = a_cur & visited;                 // visited neighbours except the trail predecessor = the bracket
        if (prev >= 0) Mb &= ~K::bit(prev);         // (the predecessor was accounted for when we left it)
        visited |= bit;
        const MaskT C0 = a_cur & ~visited;
        rem[cur * 32] = (unsigned char)K::popc(C0);
        if (C0) open |= bit;
        const int m = K::popc(Mb);
        MaskT V = 0;                                // the members' visit indices: set bits enumerate ascending, no sort
        K::each(Mb, [&](int u) {
          const int r = (int)rem[u * 32] - 1;       // cur is no longer an unvisited neighbour of u
          rem[u * 32] = (unsigned char)r;
          if (r == 0) open &= ~K::bit(u);
          V |= K::bit((int)vis[u * 32]);
        });
        int cnt = ((pre >> 7) & 1) + ((pre >> 6) & 1) + 1 + (m ? 2 + (LABELED ? 2 * m : m) : 0);
        if (LABELED) cnt += 1 + (prev >= 0 ? 1 : 0);
        log_lo[my_n0 + k] = make_uint2((unsigned)V, (unsigned)cur | ((unsigned)pre << 8) | ((unsigned)pos << 16));
        if (sizeof(MaskT) == 8 && k >= 32) log_hi[my_n0 + k] = (unsigned)((unsigned long long)V >> 32);
        pos += cnt;
        pre = 0;
      } else {
        pre = kPreReset | kPreHead | (int)vis[cur * 32];   // head of a new segment at a visited node (reached by a restart)
      }
      const MaskT C = a_cur & ~visited;
      if ((!C && nvis == N) || pos >= Lcap) {
        done = true;
      } else {
        MaskT sel = C;
        if (!C) {                                   // stuck: restart at an open node, else at any unvisited node
          sel = (open && !restart_unvisited) ? open : (MaskT)(~visited & nodemask);
          pre = kPreReset;
        }
        const int nsel = K::popc(sel);
        const int kth = pick(nsel);                 // one word per arrival, also when there is no choice
        const int nxt = nsel == 1 ? K::first(sel) : K::select(sel, kth);
        a_nxt = (MaskT)__ldg(arow + nxt);
        if (C) {                                    // trail step: nxt stops being an unvisited neighbour of cur
          rem[cur * 32] = (unsigned char)(nsel - 1);
          if (nsel == 1) open &= ~bit;
          prev = cur;
        } else {
          prev = -1;
        }
        cur = nxt;
      }
    }
  }
  if (valid) {
    if (nvis < N) log_lo[my_n0 + nvis] = make_uint2(0u, (unsigned)kOrdEnd);      // early stop: end of the item list
    const int Lfull = pos + (cfg.no_eos ? 0 : 1);   // + EOS
    len_out[g] = (Lfull < Lcap ? Lfull : Lcap) + (cfg.append_query ? 3 : 0);
  }
}

template <bool LABELED>
__global__ void __launch_bounds__(kSmallThreads, 10)   // 48 registers: 40 warps per SM (56+ registers: 0.72 -> 0.84 ms)
sent_walk_kernel(const GtokCsrBatch c, const GtokSentCfg cfg, const GtokChoice ch, const int nmax,
                 uint2* __restrict__ log_lo, unsigned* __restrict__ log_hi, int32_t* __restrict__ len_out) {
  extern __shared__ __align__(16) unsigned char w_smem[];
  const int lane = lane_id();
  const int w = threadIdx.x >> 5;
  unsigned char* vis = w_smem + (size_t)w * (64 * nmax) + lane;      // u8 [v][lane]  node -> visit index
  unsigned char* rem = vis + 32 * nmax;                               // u8 [v][lane]  unvisited neighbours left (visited nodes)
  const int64_t G = c.num_graphs;
  const int64_t wid0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g0 = wid0 * 32; g0 < G; g0 += nw * 32) {
    const int Tc = (int)((G - g0 < 32) ? (G - g0) : 32);
    const bool valid = lane < Tc;
    const int64_t g = valid ? (c.order ? (int64_t)__ldg(c.order + g0 + lane) : g0 + lane) : g0;
    const int64_t my_n0 = __ldg(c.node_off + g);
    const int N = valid ? (int)(__ldg(c.node_off + g + 1) - my_n0) : 0;
    if (__all_sync(kFull, N <= 32))
      walk_body<unsigned, LABELED>(c, cfg, ch, vis, rem, log_lo, log_hi, len_out, valid, g, my_n0, N);
    else
      walk_body<unsigned long long, LABELED>(c, cfg, ch, vis, rem, log_lo, log_hi, len_out, valid, g, my_n0, N);
    __syncwarp();
  }
}

// Emission.  COLLATE = false: row r of the batch goes to out + r * stride (r = graph index).  COLLATE = true: rows are
// grouped B at a time in batch_idx order; row r (graph batch_idx[r]) goes to X / attn + xoff[r / B] + (r % B) * lmax[r / B],
// padded to lmax[r / B] (gtok_pad_mask_batched's layout).
template <bool LABELED, typename TokT, bool COLLATE, bool REMAP>
__global__ void __launch_bounds__(kEmitThreads)
sent_emit_kernel(const GtokCsrBatch c, const GtokSentCfg cfg, const uint2* __restrict__ log_lo,
                 const unsigned* __restrict__ log_hi, const int32_t* __restrict__ len, int64_t* __restrict__ out,
                 int64_t stride, const int64_t* __restrict__ batch_idx, int64_t num_rows, int64_t B,
                 const int32_t* __restrict__ lmax, const int64_t* __restrict__ xoff, int64_t pad_id,
                 uint8_t* __restrict__ attn, int row_cap) {
  constexpr int kRemapN = sizeof(TokT) == 1 ? 256 : 1024;
  extern __shared__ __align__(16) unsigned char e_smem[];
  __shared__ unsigned short s_remap[kRemapN];
  __shared__ unsigned char s_ord[kEmitWarps][64];
  __shared__ ulonglong2 s_m16[17];
  if (COLLATE) {
    fill_mask_table(s_m16, threadIdx.x);
    __syncthreads();
  }
  const int lane = lane_id();
  const int w = threadIdx.x >> 5;
  TokT* rowb = reinterpret_cast<TokT*>(e_smem) + (size_t)w * row_cap;
  unsigned char* ordb = s_ord[w];
  const unsigned long long* __restrict__ adj64 = reinterpret_cast<const unsigned long long*>(c.adj64);
  const bool bracket_desc = cfg.bracket_order == GTOK_SENT_BRACKET_VISIT_DESC;
  const int idx_off = cfg.idx_offset, node_off = cfg.node_idx_offset, edge_off = cfg.edge_idx_offset;
  const int elab_off = edge_off + cfg.edge_label_shift;
  const int q3 = cfg.append_query ? 3 : 0;
  if (REMAP) {
    for (int t = threadIdx.x; t < kRemapN; t += blockDim.x) s_remap[t] = (unsigned short)remap_zinc_tok(t, idx_off, node_off, edge_off);
    __syncthreads();
  }
  const int64_t wid0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = wid0; r < num_rows; r += nw) {
    // (requesting the next row's extents and items one row ahead was tried: 0.42 -> 0.57 ms, the extra live registers cost
    // more occupancy than the shorter dependent chain wins)
    const int64_t g = (COLLATE && batch_idx) ? __ldg(batch_idx + r) : r;
    const int64_t n0 = __ldg(c.node_off + g);
    const int N = (int)(__ldg(c.node_off + g + 1) - n0);
    const int L = __ldg(len + g);                    // row length, query included
    const int L0 = L - q3;                           // tokens of the trail (clipped at max_length by the walk)
    if (lane == 0) rowb[0] = (TokT)SOS;
    int end_pos = (N == 0 && lane == 0) ? 1 : -1;    // position of EOS (known to the lane of the last item)
    // ---- items -> row buffer (lane <-> visit index, two rounds above 32 nodes) ----------------------------------------
    bool more = true;
    for (int kb = 0; kb < N && more; kb += 32) {
      const int k = kb + lane;
      uint2 it = make_uint2(0u, (unsigned)kOrdEnd);
      if (k < N) it = __ldg(log_lo + n0 + k);
      unsigned vhi = 0;
      if (kb && k < N) vhi = __ldg(log_hi + n0 + k);
      const int u = it.y & 0xff;
      const unsigned endm = __ballot_sync(kFull, u == kOrdEnd);       // first lane past the item list (early stop / k >= N)
      const int nk = endm ? __ffs((int)endm) - 1 : 32;                // items of this round: lanes [0, nk)
      more = endm == 0u;
      if (nk > 0) end_pos = -1;                                        // a later round holds the last item
      if (LABELED) {                                                   // visit index -> node: edge labels of predecessors / members
        if (lane < nk) ordb[k] = (unsigned char)u;
        __syncwarp();
      }
      if (lane < nk) {
        unsigned long long a_u = 0;
        int rp_u = 0, xl = 0;
        if (LABELED) {
          a_u = __ldg(adj64 + n0 + u);
          rp_u = __ldg(c.rowptr + n0 + u);
          xl = __ldg(c.x + n0 + u);
        }
        const int pr = (it.y >> 8) & 0xff;
        int p = (int)(it.y >> 16);
        int pv = -1;                                                   // trail predecessor (node id), -1 = none
        if (LABELED) {
          if (pr & kPreHead) pv = ordb[pr & 63];
          else if (!(pr & kPreReset) && k > 0) pv = ordb[k - 1];
        }
        if (pr & kPreReset) rowb[p++] = (TokT)RESET;
        if (pr & kPreHead) rowb[p++] = (TokT)(idx_off + (pr & 63));
        unsigned vlo = it.x;
        if (LABELED && pv >= 0) rowb[p++] = (TokT)(elab_off + __ldg(c.eattr + rp_u + __popcll(a_u & ((1ull << pv) - 1ull))));
        rowb[p++] = (TokT)(idx_off + k);
        if (LABELED) rowb[p++] = (TokT)(node_off + xl);
        if (vlo | vhi) {
          rowb[p++] = (TokT)LADJ;
          auto member = [&](int kv) {
            if (LABELED) {
              const int wn = ordb[kv];
              rowb[p++] = (TokT)(elab_off + __ldg(c.eattr + rp_u + __popcll(a_u & ((1ull << wn) - 1ull))));
            }
            rowb[p++] = (TokT)(idx_off + kv);
          };
          if (bracket_desc) {
            while (vhi) { const int b = 31 - __clz((int)vhi); vhi &= ~(1u << b); member(32 + b); }
            while (vlo) { const int b = 31 - __clz((int)vlo); vlo &= ~(1u << b); member(b); }
          } else {
            while (vlo) { const int b = __ffs((int)vlo) - 1; vlo &= vlo - 1; member(b); }
            while (vhi) { const int b = __ffs((int)vhi) - 1; vhi &= vhi - 1; member(32 + b); }
          }
          rowb[p++] = (TokT)RADJ;
        }
        if (lane == nk - 1) end_pos = p;             // the last item so far
      }
      __syncwarp();
    }
    if (end_pos >= 0) rowb[end_pos] = (TokT)EOS;     // (clipped by the stream below: at or beyond the row length it is never read)
    __syncwarp();
    // ---- stream the row: 64 tokens per store instruction ----------------------------------------------------------------
    auto rm = [&](int t) -> int64_t {
      if (REMAP) t = (sizeof(TokT) == 1 || t < kRemapN) ? (int)s_remap[t] : remap_zinc_tok(t, idx_off, node_off, edge_off);
      return (int64_t)t;
    };
    auto tokv = [&](int q) -> int64_t { return rm((int)rowb[q]); };
    int64_t qt = 0;
    if (q3 && lane < 3) qt = idx_off + (lane == 0 ? N : (lane == 1 ? __ldg(c.query_u + g) : __ldg(c.query_v + g)));
    if (!COLLATE) {
      int64_t* __restrict__ row = out + g * stride;  // rows are 16-byte aligned (even stride, aligned base)
      const int npair = L0 >> 1;
      for (int j = lane; j < npair; j += 32) {       // the row buffer is read two tokens at a time
        int t0, t1;
        if (sizeof(TokT) == 1) {
          const unsigned pr2 = reinterpret_cast<const unsigned short*>(rowb)[j];
          t0 = pr2 & 0xff; t1 = pr2 >> 8;
        } else {
          const unsigned pr2 = reinterpret_cast<const unsigned*>(rowb)[j];
          t0 = pr2 & 0xffff; t1 = pr2 >> 16;
        }
        st_stream_i64x2(row + 2 * j, rm(t0), rm(t1));
      }
      if ((L0 & 1) && lane == 0) st_stream_i64(row + L0 - 1, tokv(L0 - 1));
      if (q3 && lane < 3) row[L0 + lane] = qt;
    } else {
      const int64_t kbatch = row_batch(r, B);
      const int Lm = __ldg(lmax + kbatch);
      const int64_t cell0 = __ldg(xoff + kbatch) + (r - kbatch * B) * (int64_t)Lm;
      const int64_t q0 = q3 ? __shfl_sync(kFull, qt, 0) : 0, q1 = q3 ? __shfl_sync(kFull, qt, 1) : 0,
                    q2 = q3 ? __shfl_sync(kFull, qt, 2) : 0;
      auto full = [&](int q) -> int64_t {
        if (q < L0) return tokv(q);
        return q == L0 ? q0 : (q == L0 + 1 ? q1 : q2);
      };
      warp_pad_row_tbl(L, Lm, pad_id, out + cell0, attn + cell0, s_m16, full);
    }
    __syncwarp();
  }
}

struct WalkPlan {
  bool ok;
  bool tok8;
  bool wide;         // 65..256 nodes: the neighbour-list pair of gtok_sent_walkl.cu
  int row_cap;       // tokens per row buffer
};

}  // namespace

int sent_walkl_row_cap(const GtokCsrBatch* c, const GtokSentCfg* cfg);
int sent_walkl_launch(const GtokCsrBatch* c, const GtokSentCfg* cfg, const GtokChoice* ch, void* log, int32_t* len_out,
                      cudaStream_t st);
int sent_emitl_launch(const GtokCsrBatch* c, const GtokSentCfg* cfg, int row_cap, bool collate, const void* log,
                      const int32_t* len, int64_t* out, int64_t stride, const int64_t* batch_idx, int64_t num_rows, int64_t B,
                      const int32_t* lmax, const int64_t* xoff, int64_t pad_id, uint8_t* attn, cudaStream_t st);

namespace {

static bool cfg_force_mid() {       // GTOK_SENT_MID=1: 65..128 nodes through the single kernel of gtok_sent_mid.cu (A/B runs)
  static const bool f = [] { const char* e = getenv("GTOK_SENT_MID"); return e && atoi(e) != 0; }();
  return f;
}

static WalkPlan walk_plan(const GtokCsrBatch* c, const GtokSentCfg* cfg) {
  WalkPlan p{false, false, false, 0};
  if (c->max_nodes > 256 || c->max_nodes < 1) return p;
  p.wide = c->max_nodes > 64;
  // 65..128 nodes stay with the two-mask-word single kernel (gtok_sent_mid.cu: 2.9 ms per 32M nodes against 3.3 ms for the
  // neighbour-list pair); 129..256 nodes: 8.5 -> 5.5 ms (profiles/r2_large_graph_sweep.jsonl)
  if (p.wide && c->max_nodes <= 128 && c->adj64 != nullptr && cfg_force_mid()) return p;
  if (p.wide ? (c->col == nullptr || c->rowptr == nullptr) : (c->adj64 == nullptr || c->adj_words > 1)) return p;
  if (cfg->idx_offset < 0 || (!p.wide && cfg->max_length > 32768)) return p;
  long long tmax = (long long)cfg->idx_offset + c->max_nodes;   // largest token before the remap (needs the label bound)
  if (cfg->labeled) {
    if (c->max_label <= 0 || cfg->node_idx_offset < 0 || cfg->edge_idx_offset + cfg->edge_label_shift < 0) return p;
    const long long a = (long long)cfg->node_idx_offset + c->max_label;
    const long long b = (long long)cfg->edge_idx_offset + c->max_label + cfg->edge_label_shift;
    tmax = a > tmax ? a : tmax;
    tmax = b > tmax ? b : tmax;
  }
  if (tmax >= 65536) return p;
  if (p.wide) {
    p.row_cap = sent_walkl_row_cap(c, cfg);
    p.ok = p.row_cap > 0;
    return p;
  }
  p.tok8 = tmax < 256;
  p.row_cap = (cfg->max_length + kItemMax + 4 + 15) & ~15;
  if ((size_t)kEmitWarps * p.row_cap * (p.tok8 ? 1 : 2) > 160 * 1024) return p;
  p.ok = true;
  return p;
}

}  // namespace

size_t sent_log_bytes(int64_t total_nodes) { return (size_t)(total_nodes > 0 ? total_nodes : 0) * 12 + 16; }

static inline unsigned* log_hi_of(void* log, int64_t total_nodes) {
  return reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(log) + (size_t)total_nodes * 8);
}

int sent_walk_launch(const GtokCsrBatch* c, const GtokSentCfg* cfg, const GtokChoice* ch, int64_t total_nodes, void* log,
                     int32_t* len_out, cudaStream_t st) {
  if (c->max_nodes > 64) return sent_walkl_launch(c, cfg, ch, log, len_out, st);
  auto k = cfg->labeled ? sent_walk_kernel<true> : sent_walk_kernel<false>;
  const size_t smem = (size_t)kSmallWarps * 64 * c->max_nodes;
  GTOK_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t warps = (c->num_graphs + 31) / 32;
  const int grid = persistent_grid(k, kSmallThreads, smem, (warps + kSmallWarps - 1) / kSmallWarps);
  k<<<grid, kSmallThreads, smem, st>>>(*c, *cfg, *ch, c->max_nodes, reinterpret_cast<uint2*>(log), log_hi_of(log, total_nodes),
                                       len_out);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

template <bool COLLATE>
static int sent_emit_launch(const GtokCsrBatch* c, const GtokSentCfg* cfg, const WalkPlan& p, int64_t total_nodes, const void* log,
                            const int32_t* len, int64_t* out, int64_t stride, const int64_t* batch_idx, int64_t num_rows,
                            int64_t B, const int32_t* lmax, const int64_t* xoff, int64_t pad_id, uint8_t* attn,
                            cudaStream_t st) {
  if (p.wide)
    return sent_emitl_launch(c, cfg, p.row_cap, COLLATE, log, len, out, stride, batch_idx, num_rows, B, lmax, xoff, pad_id, attn, st);
  void (*k)(GtokCsrBatch, GtokSentCfg, const uint2*, const unsigned*, const int32_t*, int64_t*, int64_t, const int64_t*,
            int64_t, int64_t, const int32_t*, const int64_t*, int64_t, uint8_t*, int);
  if (cfg->remap_zinc) {
    if (cfg->labeled) k = p.tok8 ? sent_emit_kernel<true, uint8_t, COLLATE, true> : sent_emit_kernel<true, uint16_t, COLLATE, true>;
    else k = p.tok8 ? sent_emit_kernel<false, uint8_t, COLLATE, true> : sent_emit_kernel<false, uint16_t, COLLATE, true>;
  } else {
    if (cfg->labeled) k = p.tok8 ? sent_emit_kernel<true, uint8_t, COLLATE, false> : sent_emit_kernel<true, uint16_t, COLLATE, false>;
    else k = p.tok8 ? sent_emit_kernel<false, uint8_t, COLLATE, false> : sent_emit_kernel<false, uint16_t, COLLATE, false>;
  }
  const size_t smem = (size_t)kEmitWarps * p.row_cap * (p.tok8 ? 1 : 2);
  GTOK_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = persistent_grid(k, kEmitThreads, smem, (num_rows + kEmitWarps - 1) / kEmitWarps);
  k<<<grid, kEmitThreads, smem, st>>>(*c, *cfg, reinterpret_cast<const uint2*>(log),
                                      log_hi_of(const_cast<void*>(log), total_nodes), len, out, stride, batch_idx, num_rows, B, lmax,
                                      xoff, pad_id, attn, p.row_cap);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

// Called by gtok_sent_tokenize_ws for batches whose graphs all have <= 64 nodes when the caller passed a workspace of
// sent_log_bytes(total_nodes) (GTOK_SENT_SMALL=1 forces the one-kernel path of gtok_sent_small.cu, which also takes the
// shapes this one declines: odd strides, tokens beyond 16 bits, no workspace).
int sent_walk_dispatch(const GtokCsrBatch* c, const GtokSentCfg* cfg, const GtokChoice* ch, int32_t* len_out, int64_t* out,
                       int64_t stride, int64_t total_nodes, void* ws, size_t ws_bytes, cudaStream_t st, bool* handled) {
  *handled = false;
  const WalkPlan p = walk_plan(c, cfg);
  if (!p.ok || (stride & 1) || ws == nullptr || ws_bytes < sent_log_bytes(total_nodes)) return GTOK_OK;
  if ((reinterpret_cast<uintptr_t>(ws) & 7) != 0) return GTOK_OK;
  int rc = sent_walk_launch(c, cfg, ch, total_nodes, ws, len_out, st);
  if (rc != GTOK_OK) return rc;
  rc = sent_emit_launch<false>(c, cfg, p, total_nodes, ws, len_out, out, stride, nullptr, c->num_graphs, 1, nullptr, nullptr, 0,
                               nullptr, st);
  if (rc != GTOK_OK) return rc;
  *handled = true;
  return GTOK_OK;
}

bool sent_walk_supported(const GtokCsrBatch* c, const GtokSentCfg* cfg) { return walk_plan(c, cfg).ok; }

}  // namespace gtok

extern "C" size_t gtok_sent_log_bytes(const GtokCsrBatch* csr, const GtokSentCfg* cfg, int64_t total_nodes) {
  if (!csr || !cfg || !gtok::walk_plan(csr, cfg).ok) return 0;
  return gtok::sent_log_bytes(total_nodes);
}

extern "C" int gtok_sent_walk(const GtokCsrBatch* csr, const GtokSentCfg* cfg, const GtokChoice* choice, int64_t total_nodes,
                              void* log, size_t log_bytes, int32_t* len_out, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(csr && cfg && choice, GTOK_EINVAL, "sent_walk: null csr/cfg/choice");
  if (csr->num_graphs <= 0) return GTOK_OK;
  GTOK_REQUIRE(csr->node_off && len_out && log, GTOK_EINVAL, "sent_walk: null pointer");
  GTOK_REQUIRE(walk_plan(csr, cfg).ok, GTOK_EUNSUPPORTED, "sent_walk: needs graphs of at most 256 nodes (adjacency bit rows up to 64), a "
               "label bound and tokens below 65536 (gtok_sent_log_bytes returns 0 for anything else)");
  GTOK_REQUIRE(cfg->max_length >= 1, GTOK_EINVAL, "sent_walk: max_length must be >= 1");
  GTOK_REQUIRE(log_bytes >= sent_log_bytes(total_nodes) && (reinterpret_cast<uintptr_t>(log) & 7) == 0, GTOK_EINVAL,
               "sent_walk: log of %zu bytes (8-byte aligned) required", sent_log_bytes(total_nodes));
  GTOK_REQUIRE(!cfg->labeled || (csr->x && csr->rowptr && (csr->eattr || csr->max_nnz == 0)), GTOK_EINVAL,
               "sent_walk: labeled mode needs csr.x, csr.rowptr and csr.eattr");
  GTOK_REQUIRE(!cfg->append_query || (csr->query_u && csr->query_v), GTOK_EINVAL, "sent_walk: append_query needs query_u/query_v");
  GTOK_REQUIRE(choice->stream == nullptr || choice->stream_off != nullptr, GTOK_EINVAL, "sent_walk: choice stream without offsets");
  GTOK_REQUIRE(cfg->restart_policy == GTOK_SENT_RESTART_OPEN_FIRST || cfg->restart_policy == GTOK_SENT_RESTART_UNVISITED,
               GTOK_EUNSUPPORTED, "sent_walk: restart_policy %d is not implemented by the kernels", cfg->restart_policy);
  GTOK_REQUIRE(cfg->bracket_order == GTOK_SENT_BRACKET_VISIT_ASC || cfg->bracket_order == GTOK_SENT_BRACKET_VISIT_DESC,
               GTOK_EUNSUPPORTED, "sent_walk: bracket_order %d is not implemented by the kernels", cfg->bracket_order);
  return sent_walk_launch(csr, cfg, choice, total_nodes, log, len_out, (cudaStream_t)stream);
}

extern "C" int gtok_sent_emit_collate(const GtokCsrBatch* csr, const GtokSentCfg* cfg, int64_t total_nodes, const void* log,
                                      const int32_t* row_len, const int64_t* batch_idx, int64_t num_rows, int64_t B,
                                      const int32_t* lmax, const int64_t* xoff, int64_t pad_id, int64_t* X, uint8_t* attn,
                                      void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(csr && cfg, GTOK_EINVAL, "sent_emit_collate: null csr/cfg");
  if (num_rows <= 0) return GTOK_OK;
  GTOK_REQUIRE(log && row_len && lmax && xoff && X && attn && B > 0, GTOK_EINVAL, "sent_emit_collate: null pointer / batch size");
  const WalkPlan p = walk_plan(csr, cfg);
  GTOK_REQUIRE(p.ok, GTOK_EUNSUPPORTED, "sent_emit_collate: configuration not supported by the walk/emission kernels");
  return sent_emit_launch<true>(csr, cfg, p, total_nodes, log, row_len, X, 0, batch_idx, num_rows, B, lmax, xoff, pad_id, attn,
                                (cudaStream_t)stream);
}
