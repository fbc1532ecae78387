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
#include "gtok_common.cuh"

namespace gtok {

namespace {

constexpr int SOS = 0, RESET = 1, LADJ = 2, RADJ = 3, EOS = 4;
constexpr int kSmallThreads = 128;
constexpr int kSmallWarps = kSmallThreads / 32;

__device__ __forceinline__ void philox10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0,
                                         unsigned k1, unsigned (&o)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// trainer/train_agtt.py:193-242 (branch order preserved; oracle: remap_zinc_token)
__device__ __forceinline__ int remap_zinc_tok(int t, int idx_off, int node_off, int edge_off) {
  if (t == 0) return 0;
  if (t == 1 || t == 2 || t == 3 || t == 5) return 2;
  if (t == 4) return 1;
  if (t >= node_off && t < edge_off) { const int a = t - node_off; return (a >= 0 && a < 9) ? 8 + a : 22 + t; }
  if (t >= edge_off) { const int b = t - edge_off + 1; return (b >= 1 && b <= 4) ? 17 + (b - 1) : 22 + t; }
  if (t >= idx_off && t < node_off) return 22 + (t - idx_off);
  return 22 + t;
}

// k-th (0-based) set bit; popc(m) > k.  (__fns compiles to a long loop: 230 instr/graph in the first profile)
__device__ __forceinline__ int select32(unsigned m, int k) {
  int pos = 0, c;
  c = __popc(m & 0xFFFFu); if (k >= c) { k -= c; pos = 16; m >>= 16; }
  c = __popc(m & 0xFFu);   if (k >= c) { k -= c; pos += 8; m >>= 8; }
  c = __popc(m & 0xFu);    if (k >= c) { k -= c; pos += 4; m >>= 4; }
  c = __popc(m & 0x3u);    if (k >= c) { k -= c; pos += 2; m >>= 2; }
  if (k >= (int)(m & 1u)) pos += 1;
  return pos;
}
__device__ __forceinline__ int select64(unsigned long long m, int k) {
  const unsigned lo = (unsigned)m;
  const int cl = __popc(lo);
  return k < cl ? select32(lo, k) : 32 + select32((unsigned)(m >> 32), k - cl);
}

struct SmallLayout {
  int o_vis, o_ord;   // byte offsets inside a warp slice
  int warp_bytes;
};

}  // namespace

static SmallLayout small_layout(int nmax, bool labeled) {
  SmallLayout L;
  int o = 0;
  L.o_vis = o; o += 32 * nmax;                 // u8 [v][lane]  node -> visit index
  L.o_ord = o; o += labeled ? 32 * nmax : 0;   // u8 [k][lane]  visit index -> node
  L.warp_bytes = (o + 15) & ~15;
  return L;
}

template <bool LABELED>
__global__ void __launch_bounds__(kSmallThreads)
sent_small_kernel(const GtokCsrBatch c, const GtokSentCfg cfg, const GtokChoice ch, const SmallLayout lay,
                  int32_t* __restrict__ len_out, int64_t* __restrict__ out, int64_t stride) {
  extern __shared__ __align__(16) unsigned char s_smem[];
  const int lane = lane_id();
  const int w = threadIdx.x >> 5;
  unsigned char* base = s_smem + (size_t)w * lay.warp_bytes;
  unsigned char* vis = base + lay.o_vis;
  unsigned char* ord = base + lay.o_ord;
  const unsigned long long* __restrict__ adj64 = reinterpret_cast<const unsigned long long*>(c.adj64);

  const int64_t G = c.num_graphs;
  const int64_t wid0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int Lcap = cfg.max_length;
  const unsigned k0 = (unsigned)(ch.seed & 0xffffffffu), k1 = (unsigned)(ch.seed >> 32);

  for (int64_t g0 = wid0 * 32; g0 < G; g0 += nw * 32) {
    const int Tc = (int)((G - g0 < 32) ? (G - g0) : 32);
    // ---- one thread walks one graph -----------------------------------------------------------
    if (lane < Tc) {
      const int64_t g = c.order ? (int64_t)__ldg(c.order + g0 + lane) : g0 + lane;
      const int64_t my_n0 = __ldg(c.node_off + g);
      const int N = (int)(__ldg(c.node_off + g + 1) - my_n0);
      int64_t* __restrict__ row = out + g * stride;
      // pairwise output: rows are 16-byte aligned (even stride, aligned base)
      int pos = 0;
      long long held = 0;
      auto emit = [&](int t) {
        if (pos < Lcap) {
          const long long v = cfg.remap_zinc ? remap_zinc_tok(t, cfg.idx_offset, cfg.node_idx_offset, cfg.edge_idx_offset) : t;
          if (pos & 1) st_stream_i64x2(row + pos - 1, held, v);
          else held = v;
        }
        ++pos;
      };
      // random words
      const uint64_t gid = (uint64_t)(ch.graph_id_base + g);
      unsigned blk[4] = {0, 0, 0, 0};
      unsigned d = 0;
      int64_t s_off = 0;
      unsigned s_len = 0;
      if (ch.stream) {
        s_off = __ldg(ch.stream_off + g);
        s_len = (unsigned)(__ldg(ch.stream_off + g + 1) - s_off);
      }
      auto pick = [&](int cnt) -> int {
        unsigned r;
        if (ch.stream) {
          r = d < s_len ? __ldg(ch.stream + s_off + d) : 0u;
        } else {
          if ((d & 3u) == 0u) philox10((unsigned)gid, (unsigned)(gid >> 32), d >> 2, 0u, k0, k1, blk);
          const unsigned w01 = (d & 1u) ? blk[1] : blk[0], w23 = (d & 1u) ? blk[3] : blk[2];
          r = (d & 2u) ? w23 : w01;
        }
        ++d;
        return (int)__umulhi(r, (unsigned)cnt);
      };
      const unsigned long long* __restrict__ arow = adj64 + my_n0;
      auto adj = [&](int r) -> unsigned long long { return __ldg(arow + r); };
      auto edge_label = [&](int u, int v) -> int {     // label of CSR entry (u -> v): rank of v inside row u
        const unsigned long long below = adj(u) & ((1ull << v) - 1ull);
        return cfg.edge_idx_offset + __ldg(c.eattr + __ldg(c.rowptr + my_n0 + u) + __popcll(below)) + cfg.edge_label_shift;
      };

      emit(SOS);
      if (N > 0) {
        const unsigned long long nodemask = N >= 64 ? ~0ull : ((1ull << N) - 1ull);
        unsigned long long visited = 0ull;
        int nvis = 0, prev = -1;
        int cur = pick(N);
        while (pos < Lcap) {
          const unsigned long long a_cur = adj(cur);
          if (!((visited >> cur) & 1ull)) {
            const int k = nvis++;
            vis[cur * 32 + lane] = (unsigned char)k;
            if (LABELED) ord[k * 32 + lane] = (unsigned char)cur;
            emit(cfg.idx_offset + k);
            if (LABELED) emit(cfg.node_idx_offset + __ldg(c.x + my_n0 + cur));
            unsigned long long M = a_cur & visited;
            if (prev >= 0) M &= ~(1ull << prev);
            visited |= 1ull << cur;
            if (M) {
              unsigned long long V = 0ull;
              while (M) {
                const int u = __ffsll((long long)M) - 1;
                M &= M - 1;
                V |= 1ull << vis[u * 32 + lane];
              }
              emit(LADJ);
              while (V) {
                const int jv = __ffsll((long long)V) - 1;
                V &= V - 1;
                if (LABELED) emit(edge_label(cur, (int)ord[jv * 32 + lane]));
                emit(cfg.idx_offset + jv);
              }
              emit(RADJ);
            }
          } else {
            emit(cfg.idx_offset + (int)vis[cur * 32 + lane]);     // head of a new segment at a visited node
          }
          const unsigned long long C = a_cur & ~visited;
          if (C) {
            const int nxt = select64(C, pick(__popcll(C)));
            if (LABELED) emit(edge_label(cur, nxt));
            prev = cur;
            cur = nxt;
          } else {
            if (nvis == N) break;
            unsigned long long U = ~visited & nodemask, acc = 0ull;
            unsigned long long t = U;
            while (t) {
              const int wv = __ffsll((long long)t) - 1;
              t &= t - 1;
              acc |= adj(wv);
            }
            const unsigned long long R1 = acc & visited;
            const unsigned long long R = R1 ? R1 : U;
            emit(RESET);
            cur = select64(R, pick(__popcll(R)));
            prev = -1;
          }
        }
      }
      emit(EOS);
      int L = pos < Lcap ? pos : Lcap;
      if (L & 1) st_stream_i64(row + L - 1, held);          // flush the unpaired last token
      if (cfg.append_query) {
        row[L] = cfg.idx_offset + N;
        row[L + 1] = cfg.idx_offset + __ldg(c.query_u + g);
        row[L + 2] = cfg.idx_offset + __ldg(c.query_v + g);
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
  if (c->max_nodes > 64 || (stride & 1) || c->adj64 == nullptr) return GTOK_OK;
  const SmallLayout lay = small_layout(c->max_nodes, cfg->labeled != 0);
  const size_t smem = (size_t)kSmallWarps * lay.warp_bytes;
  auto k = cfg->labeled ? sent_small_kernel<true> : sent_small_kernel<false>;
  GTOK_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t warps = (c->num_graphs + 31) / 32;
  const int grid = persistent_grid(k, kSmallThreads, smem, (warps + kSmallWarps - 1) / kSmallWarps);
  k<<<grid, kSmallThreads, smem, st>>>(*c, *cfg, *ch, lay, len_out, out, stride);
  GTOK_CUDA_TRY(cudaGetLastError());
  *handled = true;
  return GTOK_OK;
}

}  // namespace gtok
