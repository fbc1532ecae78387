// gtok_sent_mid.cu - K3 for graphs of 65..256 nodes: ONE THREAD PER GRAPH on multi-word bit sets.
//
// Same function and the same walk as gtok_sent_small.cu (Graph2TrailTokenizer.__call__ + remap_zinc_tokens +
// query append, trainer/train_agtt.py:171-273; spec oracle/sent_oracle.py); the 64-bit masks become W = 2 or 4
// words (adjacency rows GtokCsrBatch.adj64 with adj_words = W per node, visited / open sets in registers,
// bracket members in visit-index space).  The warp-per-graph kernel (gtok_sent.cu) pays ~2000 cycles per walk
// step in ballots and shared-memory round trips, with 31 lanes idle in the scalar part of every step; a thread
// per graph advances 32 walks per warp instruction.  Per-lane state in shared memory: visit index, visit order
// and unvisited-neighbour counter per node (u8, [node][lane]: conflict free) and the token ring.
#include "gtok_sent_common.cuh"

namespace gtok {

namespace {

constexpr int kMidRing = 64;
constexpr int kMidRemap = 1024;

template <int W>
struct Bits {
  unsigned long long w[W];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < W; ++i) w[i] = 0ull;
  }
  __device__ __forceinline__ bool test(int b) const {
    unsigned long long v = w[0];
#pragma unroll
    for (int i = 1; i < W; ++i) v = (b >> 6) == i ? w[i] : v;
    return (v >> (b & 63)) & 1ull;
  }
  __device__ __forceinline__ void set(int b) {
    const unsigned long long m = 1ull << (b & 63);
#pragma unroll
    for (int i = 0; i < W; ++i) w[i] |= (b >> 6) == i ? m : 0ull;
  }
  __device__ __forceinline__ void clear(int b) {
    const unsigned long long m = 1ull << (b & 63);
#pragma unroll
    for (int i = 0; i < W; ++i) w[i] &= (b >> 6) == i ? ~m : ~0ull;
  }
  __device__ __forceinline__ bool any() const {
    unsigned long long v = 0ull;
#pragma unroll
    for (int i = 0; i < W; ++i) v |= w[i];
    return v != 0ull;
  }
  __device__ __forceinline__ int count() const {
    int c = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) c += __popcll(w[i]);
    return c;
  }
  __device__ __forceinline__ int pop_highest() {         // index of the highest set bit, cleared; any() must hold
    int r = 0;
    bool got = false;
#pragma unroll
    for (int i = W - 1; i >= 0; --i) {
      if (!got && w[i]) {
        const int b = 63 - __clzll((long long)w[i]);
        r = 64 * i + b;
        w[i] &= ~(1ull << b);
        got = true;
      }
    }
    return r;
  }
  __device__ __forceinline__ int pop_lowest() {          // index of the lowest set bit, cleared; any() must hold
    int r = 0;
    bool got = false;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      if (!got && w[i]) {
        r = 64 * i + __ffsll((long long)w[i]) - 1;
        w[i] &= w[i] - 1;
        got = true;
      }
    }
    return r;
  }
  __device__ __forceinline__ int select(int k) const {   // k-th (0-based) set bit; count() > k
    int r = 0;
    bool got = false;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const int c = __popcll(w[i]);
      if (!got) {
        if (k < c) { r = 64 * i + select64(w[i], k); got = true; }
        else k -= c;
      }
    }
    return r;
  }
  __device__ __forceinline__ int rank_below(int b) const {  // set bits at positions < b
    int c = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const int lo = b - 64 * i;                          // bits of word i below b
      const unsigned long long m = lo <= 0 ? 0ull : (lo >= 64 ? ~0ull : ((1ull << lo) - 1ull));
      c += __popcll(w[i] & m);
    }
    return c;
  }
};
template <int W>
__device__ __forceinline__ Bits<W> b_and(const Bits<W>& a, const Bits<W>& b) {
  Bits<W> r;
#pragma unroll
  for (int i = 0; i < W; ++i) r.w[i] = a.w[i] & b.w[i];
  return r;
}
template <int W>
__device__ __forceinline__ Bits<W> b_andnot(const Bits<W>& a, const Bits<W>& b) {
  Bits<W> r;
#pragma unroll
  for (int i = 0; i < W; ++i) r.w[i] = a.w[i] & ~b.w[i];
  return r;
}

struct MidLayout {
  int o_vis, o_rem, o_ord, o_ring;
  int warp_bytes;
};

MidLayout mid_layout(int nmax, bool labeled, int ring_elem) {
  MidLayout L;
  int o = 0;
  L.o_vis = o; o += 32 * nmax;
  L.o_rem = o; o += 32 * nmax;
  L.o_ord = o; o += labeled ? 32 * nmax : 0;
  o = (o + 15) & ~15;
  L.o_ring = o; o += kMidRing * 32 * ring_elem;
  L.warp_bytes = (o + 15) & ~15;
  return L;
}

template <int W, bool LABELED, typename RingT>
__global__ void __launch_bounds__(kSmallThreads)
sent_mid_kernel(const GtokCsrBatch c, const GtokSentCfg cfg, const GtokChoice ch, const MidLayout lay,
                int32_t* __restrict__ len_out, int64_t* __restrict__ out, int64_t stride) {
  extern __shared__ __align__(16) unsigned char m_smem[];
  __shared__ unsigned short s_remap[kMidRemap];
  const int lane = lane_id();
  const int w = threadIdx.x >> 5;
  unsigned char* base = m_smem + (size_t)w * lay.warp_bytes;
  unsigned char* vis = base + lay.o_vis + lane;
  unsigned char* rem = base + lay.o_rem + lane;
  unsigned char* ord = base + lay.o_ord + lane;
  RingT* ring = reinterpret_cast<RingT*>(base + lay.o_ring) + lane;
  const unsigned long long* __restrict__ adj = reinterpret_cast<const unsigned long long*>(c.adj64);

  const int64_t G = c.num_graphs;
  const int64_t wid0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int Lcap = cfg.max_length;
  const unsigned k0 = (unsigned)(ch.seed & 0xffffffffu), k1 = (unsigned)(ch.seed >> 32);
  const bool remap = cfg.remap_zinc != 0;
  const bool restart_unvisited = cfg.restart_policy == GTOK_SENT_RESTART_UNVISITED;
  const bool bracket_desc = cfg.bracket_order == GTOK_SENT_BRACKET_VISIT_DESC;
  const int idx_off = cfg.idx_offset, node_off = cfg.node_idx_offset, edge_off = cfg.edge_idx_offset;
  if (remap) {
    for (int t = threadIdx.x; t < kMidRemap; t += blockDim.x)
      s_remap[t] = (unsigned short)remap_zinc_tok(t, idx_off, node_off, edge_off);
    __syncthreads();
  }

  for (int64_t g0 = wid0 * 32; g0 < G; g0 += nw * 32) {
    const int Tc = (int)((G - g0 < 32) ? (G - g0) : 32);
    const bool valid = lane < Tc;
    const int64_t g = valid ? (c.order ? (int64_t)__ldg(c.order + g0 + lane) : g0 + lane) : g0;
    const int64_t my_n0 = __ldg(c.node_off + g);
    const int N = valid ? (int)(__ldg(c.node_off + g + 1) - my_n0) : 0;
    const unsigned long long* __restrict__ arow = adj + my_n0 * W;
    int64_t* __restrict__ row = out + g * stride;

    int pos = 0, flushed = 0;
    auto emit = [&](int t) {
      if (pos < Lcap) ring[(pos & (kMidRing - 1)) * 32] = (RingT)t;
      ++pos;
    };
    auto tok = [&](int p) -> long long {
      const int t = (int)ring[(p & (kMidRing - 1)) * 32];
      if (!remap) return t;
      return (unsigned)t < (unsigned)kMidRemap ? (int)s_remap[t] : remap_zinc_tok(t, idx_off, node_off, edge_off);
    };
    auto flush16 = [&]() {
#pragma unroll
      for (int i = 0; i < 16; i += 2) st_stream_i64x2(row + flushed + i, tok(flushed + i), tok(flushed + i + 1));
      flushed += 16;
    };
    auto flush_guard = [&]() {
      if (pos - flushed >= kMidRing - 8 && flushed + 16 <= Lcap) flush16();
    };
    auto load_row = [&](int v) -> Bits<W> {
      Bits<W> r;
#pragma unroll
      for (int i = 0; i < W; ++i) r.w[i] = __ldg(arow + (int64_t)v * W + i);
      return r;
    };
    auto edge_label = [&](int rp, int v, const Bits<W>& au) -> int {
      return edge_off + __ldg(c.eattr + rp + au.rank_below(v)) + cfg.edge_label_shift;
    };

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
    Bits<W> visited, open, nodemask;
    visited.zero(); open.zero();
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const int lo = N - 64 * i;
      nodemask.w[i] = lo <= 0 ? 0ull : (lo >= 64 ? ~0ull : ((1ull << lo) - 1ull));
    }
    int nvis = 0, prev = -1, cur = 0;
    bool done = N == 0 || pos >= Lcap;
    if (!done) cur = pick(N);
    Bits<W> a_nxt;
    a_nxt.zero();
    int rp_nxt = 0;
    if (!done) {
      a_nxt = load_row(cur);
      if (LABELED) rp_nxt = __ldg(c.rowptr + my_n0 + cur);
    }
    while (true) {
      if (!done) {
        const Bits<W> a_cur = a_nxt;
        const int rp_cur = rp_nxt;
        if (!visited.test(cur)) {
          const int k = nvis++;
          vis[cur * 32] = (unsigned char)k;
          if (LABELED) ord[k * 32] = (unsigned char)cur;
          emit(idx_off + k);
          if (LABELED) emit(node_off + __ldg(c.x + my_n0 + cur));
          Bits<W> Mall = b_and(a_cur, visited);                   // visited neighbours = bracket members + predecessor
          if (prev >= 0) Mall.clear(prev);
          visited.set(cur);
          const Bits<W> C0 = b_andnot(a_cur, visited);
          const int c0 = C0.count();
          rem[cur * 32] = (unsigned char)c0;
          if (c0) open.set(cur);
          Bits<W> V;
          V.zero();
          while (Mall.any()) {
            const int u = Mall.pop_lowest();
            const int r = (int)rem[u * 32] - 1;
            rem[u * 32] = (unsigned char)r;
            if (r == 0) open.clear(u);
            V.set((int)vis[u * 32]);
          }
          if (V.any()) {
            emit(LADJ);
            do {
              const int jv = bracket_desc ? V.pop_highest() : V.pop_lowest();
              if (LABELED) emit(edge_label(rp_cur, (int)ord[jv * 32], a_cur));
              emit(idx_off + jv);
              flush_guard();
            } while (V.any());
            emit(RADJ);
          }
        } else {
          emit(idx_off + (int)vis[cur * 32]);
        }
        const Bits<W> C = b_andnot(a_cur, visited);
        const bool hasC = C.any();
        if (!hasC && nvis == N) {
          done = true;
        } else {
          Bits<W> sel = C;
          if (!hasC) {
            sel = (open.any() && !restart_unvisited) ? open : b_andnot(nodemask, visited);
            emit(RESET);
          }
          const int nsel = sel.count();
          const int nxt = sel.select(pick(nsel));
          a_nxt = load_row(nxt);
          if (LABELED) rp_nxt = __ldg(c.rowptr + my_n0 + nxt);
          if (hasC) {
            rem[cur * 32] = (unsigned char)(nsel - 1);
            if (nsel == 1) open.clear(cur);
            if (LABELED) emit(edge_label(rp_cur, nxt, a_cur));
          }
          prev = hasC ? cur : -1;
          cur = nxt;
          if (pos >= Lcap) done = true;
        }
      }
      if (__any_sync(kFull, pos - flushed >= 32)) {
        if (pos - flushed >= 16 && flushed + 16 <= Lcap) flush16();
        if (pos - flushed >= 32 && flushed + 16 <= Lcap) flush16();   // a labeled event can add more than 16 tokens:
      }                                                                   // pending must be below 32 when the next one starts
      if (__all_sync(kFull, done)) break;
    }
    if (valid) {
      if (!cfg.no_eos) emit(EOS);
      const int L0 = pos < Lcap ? pos : Lcap;
      // the ring holds the last kMidRing tokens; everything older has been flushed
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

template <int W>
int mid_launch(const GtokCsrBatch* c, const GtokSentCfg* cfg, const GtokChoice* ch, int32_t* len_out, int64_t* out,
               int64_t stride, cudaStream_t st, bool narrow) {
  const MidLayout lay = mid_layout(c->max_nodes, cfg->labeled != 0, narrow ? 2 : 4);
  const size_t smem = (size_t)kSmallWarps * lay.warp_bytes;
  int dev = 0, max_optin = 0;
  GTOK_CUDA_TRY(cudaGetDevice(&dev));
  GTOK_CUDA_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  GTOK_REQUIRE(smem + 4096 <= (size_t)max_optin, GTOK_EUNSUPPORTED, "sent: thread-per-graph state needs %zu B of shared memory", smem);
  void (*k)(GtokCsrBatch, GtokSentCfg, GtokChoice, MidLayout, int32_t*, int64_t*, int64_t);
  if (cfg->labeled) k = narrow ? sent_mid_kernel<W, true, uint16_t> : sent_mid_kernel<W, true, int32_t>;
  else k = narrow ? sent_mid_kernel<W, false, uint16_t> : sent_mid_kernel<W, false, int32_t>;
  GTOK_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t warps = (c->num_graphs + 31) / 32;
  const int grid = persistent_grid(k, kSmallThreads, smem, (warps + kSmallWarps - 1) / kSmallWarps);
  k<<<grid, kSmallThreads, smem, st>>>(*c, *cfg, *ch, lay, len_out, out, stride);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

}  // namespace

// Called by gtok_sent_tokenize for batches of 65..256-node graphs that carry multi-word adjacency rows.
int sent_mid_dispatch(const GtokCsrBatch* c, const GtokSentCfg* cfg, const GtokChoice* ch, int32_t* len_out,
                      int64_t* out, int64_t stride, cudaStream_t st, bool* handled) {
  *handled = false;
  const int W = c->adj_words;
  if (c->adj64 == nullptr || (W != 2 && W != 4) || c->max_nodes > 64 * W || c->max_nodes > 256 || (stride & 1)) return GTOK_OK;
  // (a node's degree is at most 255, so the 8-bit unvisited-neighbour counters cannot overflow)
  long long tmax = (long long)cfg->idx_offset + 256;
  bool narrow = tmax < 65536 && cfg->idx_offset >= 0;
  if (cfg->labeled) {
    const long long lab = c->max_label > 0 ? c->max_label : (1ll << 40);
    const long long a = (long long)cfg->node_idx_offset + lab, b = (long long)cfg->edge_idx_offset + lab + cfg->edge_label_shift;
    narrow = narrow && a < 65536 && b < 65536 && cfg->node_idx_offset >= 0 && cfg->edge_idx_offset + cfg->edge_label_shift >= 0;
  }
  const int rc = W == 2 ? mid_launch<2>(c, cfg, ch, len_out, out, stride, st, narrow)
                        : mid_launch<4>(c, cfg, ch, len_out, out, stride, st, narrow);
  if (rc != GTOK_OK) return rc;
  *handled = true;
  return GTOK_OK;
}

}  // namespace gtok
