// gtok_tasks.cu - K6: the steps either side of the tokenization path.
//
//   find_token_kernel   extract_query_nodes' per-row host scan for the first <q> (trainer/train_ibtt.py:89-103,
//                       trainer/train_agtt.py:101-112) for every row of a padded epoch: a warp per row, 1 KB in
//                       flight per trip, stops after the trip with the first hit.  8 bytes per cell up to the hit.
//   labels_*_kernel     cycle_check / shortest_path labels from the canonical CSR (the reference's DFS label
//                       oracle, test/train_performance_test.py:27-67, as "listed edges > N - components"; BFS hop
//                       count for `lenK` / INF).  <= 64 nodes: a thread per graph on the 64-bit adjacency rows,
//                       the whole BFS in three registers; otherwise a warp per graph with visited / frontier /
//                       next bitmaps in shared memory.
//   er_*_kernel         on-device Erdos-Renyi graphs (stream defined in oracle/tasks_oracle.py er_graph): a warp
//                       per graph, one Philox-4x32-10 block = four node pairs per lane, ordered compaction of the
//                       kept pairs with a warp scan.  Two passes (count, fill) regenerate the same words instead of
//                       storing them: ALU is free next to the 8 bytes per edge that have to be written anyway.
#include "gtok_sent_common.cuh"

namespace gtok {

constexpr int kTaskThreads = 256;

// ------------------------------------------------------------------------------------------------- query positions
__global__ void __launch_bounds__(kTaskThreads)
find_token_kernel(const int64_t* __restrict__ X, const int64_t* __restrict__ xoff, const int32_t* __restrict__ lmax,
                  int64_t rows, int64_t B, const int64_t* __restrict__ tok_per_batch, int64_t token,
                  int32_t* __restrict__ pos_out) {
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < rows; r += nwarps) {
    const int64_t k = r / B;
    const int L = __ldg(lmax + k);
    const int64_t* __restrict__ row = X + __ldg(xoff + k) + (r - k * B) * (int64_t)L;
    const int64_t t = tok_per_batch ? __ldg(tok_per_batch + k) : token;
    int pos = -1;
    for (int p0 = 0; p0 < L && pos < 0; p0 += 4 * kWarp) {       // four 256-byte loads in flight per trip
      int64_t val[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = p0 + j * kWarp + lane;
        val[j] = p < L ? __ldg(row + p) : ~t;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned hit = __ballot_sync(kFull, val[j] == t);
        if (hit && pos < 0) pos = p0 + j * kWarp + __ffs(hit) - 1;
      }
    }
    if (lane == 0) pos_out[r] = pos;
  }
}

// ------------------------------------------------------------------------------------------------- task labels
// Thread per graph, adjacency rows as 64-bit masks (GtokCsrBatch.adj64, adj_words == 1).
__global__ void __launch_bounds__(kTaskThreads)
labels_small_kernel(GtokCsrBatch c, const int64_t* __restrict__ edge_off, int32_t* __restrict__ cyc,
                    int32_t* __restrict__ dist) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < c.num_graphs; g += stride) {
    const int64_t n0 = __ldg(c.node_off + g);
    const int N = (int)(__ldg(c.node_off + g + 1) - n0);
    const unsigned long long* __restrict__ adj = reinterpret_cast<const unsigned long long*>(c.adj64) + n0;
    const unsigned long long all = N >= 64 ? ~0ull : ((1ull << N) - 1ull);
    int u = 0, v = 0;
    bool q = false;
    if (dist) {
      u = __ldg(c.query_u + g); v = __ldg(c.query_v + g);
      q = (unsigned)u < (unsigned)N && (unsigned)v < (unsigned)N;
    }
    unsigned long long vis = 0;
    int comps = 0, d = -1;
    int start = q ? u : 0;
    while (start < N) {
      unsigned long long fr = 1ull << start;
      vis |= fr;
      for (int level = 0; fr; ++level) {
        if (q && comps == 0 && d < 0 && ((fr >> v) & 1ull)) d = level;
        unsigned long long nx = 0;
        while (fr) {
          const int b = __ffsll((long long)fr) - 1;
          fr &= fr - 1;
          nx |= __ldg(adj + b);
        }
        fr = nx & ~vis & all;
        vis |= fr;
      }
      ++comps;
      const unsigned long long rest = ~vis & all;
      start = rest ? __ffsll((long long)rest) - 1 : N;
    }
    if (cyc) {
      const long long M = edge_off ? (long long)(__ldg(edge_off + g + 1) - __ldg(edge_off + g))
                                   : (long long)(__ldg(c.rowptr + n0 + N) - __ldg(c.rowptr + n0)) / 2;
      cyc[g] = M > (long long)N - comps ? 1 : 0;
    }
    if (dist) dist[g] = d;
  }
}

// Warp per graph, any size whose three bitmaps fit the warp's shared-memory slice (`words` 32-bit words each).
__global__ void __launch_bounds__(kTaskThreads)
labels_warp_kernel(GtokCsrBatch c, const int64_t* __restrict__ edge_off, int32_t* __restrict__ cyc,
                   int32_t* __restrict__ dist, int words) {
  extern __shared__ unsigned t_smem[];
  const int lane = lane_id();
  unsigned* vis = t_smem + (size_t)(threadIdx.x >> 5) * 3 * words;
  unsigned* fr = vis + words;
  unsigned* nx = fr + words;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = warp0; g < c.num_graphs; g += nwarps) {
    const int64_t n0 = __ldg(c.node_off + g);
    const int N = (int)(__ldg(c.node_off + g + 1) - n0);
    const int Nw = (N + 31) >> 5;
    const int32_t* __restrict__ rp = c.rowptr + n0;
    for (int i = lane; i < Nw; i += kWarp) { vis[i] = 0u; fr[i] = 0u; nx[i] = 0u; }
    int u = 0, v = 0;
    bool q = false;
    if (dist) {
      u = __ldg(c.query_u + g); v = __ldg(c.query_v + g);
      q = (unsigned)u < (unsigned)N && (unsigned)v < (unsigned)N;
    }
    __syncwarp();
    int comps = 0, d = -1;
    int start = q ? u : 0;
    while (start < N) {
      if (lane == 0) { fr[start >> 5] = 1u << (start & 31); vis[start >> 5] |= 1u << (start & 31); }
      __syncwarp();
      for (int level = 0;; ++level) {
        if (q && comps == 0 && d < 0 && ((fr[v >> 5] >> (v & 31)) & 1u)) d = level;       // same word in every lane
        for (int i = lane; i < Nw; i += kWarp) {
          unsigned w = fr[i];
          while (w) {
            const int node = (i << 5) + __ffs(w) - 1;
            w &= w - 1;
            const int e1 = __ldg(rp + node + 1);
            for (int e = __ldg(rp + node); e < e1; ++e) {
              const int t = __ldg(c.col + e);
              const unsigned bit = 1u << (t & 31);
              if (!(vis[t >> 5] & bit)) atomicOr(&nx[t >> 5], bit);
            }
          }
        }
        __syncwarp();
        unsigned any = 0;
        for (int i = lane; i < Nw; i += kWarp) {
          const unsigned nw = nx[i] & ~vis[i];
          vis[i] |= nw; fr[i] = nw; nx[i] = 0u;
          any |= nw;
        }
        __syncwarp();
        if (!__any_sync(kFull, any != 0u)) break;
      }
      ++comps;
      start = N;                                             // lowest node not reached yet
      for (int base = 0; base < Nw; base += kWarp) {
        const int i = base + lane;
        unsigned m = i < Nw ? ~vis[i] : 0u;
        if (i == Nw - 1 && (N & 31)) m &= (1u << (N & 31)) - 1u;
        const unsigned ball = __ballot_sync(kFull, m != 0u);
        if (ball) {
          const int l = __ffs(ball) - 1;
          const unsigned mm = __shfl_sync(kFull, m, l);
          start = ((base + l) << 5) + __ffs(mm) - 1;
          break;
        }
      }
      __syncwarp();
    }
    if (lane == 0) {
      if (cyc) {
        const long long M = edge_off ? (long long)(__ldg(edge_off + g + 1) - __ldg(edge_off + g))
                                     : (long long)(__ldg(rp + N) - __ldg(rp)) / 2;
        cyc[g] = M > (long long)N - comps ? 1 : 0;
      }
      if (dist) dist[g] = d;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------- ER generator
struct ErHeader { int n; unsigned thr; int qu, qv; int m; };

__device__ __forceinline__ unsigned mulhi_u32(unsigned a, unsigned b) { return __umulhi(a, b); }

__device__ __forceinline__ ErHeader er_header(const GtokErSpec& s, unsigned long long gid) {
  const unsigned k0 = (unsigned)s.seed, k1 = (unsigned)(s.seed >> 32), g0 = (unsigned)gid, g1 = (unsigned)(gid >> 32);
  unsigned w0, w1, w2, w3, w4, x1, x2, x3;
  philox10(g0, g1, 0u, 1u, k0, k1, w0, w1, w2, w3);
  philox10(g0, g1, 1u, 1u, k0, k1, w4, x1, x2, x3);
  const int b = (int)mulhi_u32(w0, (unsigned)s.num_buckets);
  ErHeader h;
  h.n = s.n_lo[b] + (int)mulhi_u32(w1, (unsigned)(s.n_hi[b] - s.n_lo[b]));
  h.thr = s.thr_lo + mulhi_u32(w2, s.thr_hi - s.thr_lo);
  h.qu = (int)mulhi_u32(w3, (unsigned)h.n);
  h.qv = h.n > 1 ? (h.qu + 1 + (int)mulhi_u32(w4, (unsigned)(h.n - 1))) % h.n : h.qu;
  h.m = 0;
  if (s.kind == GTOK_GEN_BA || s.kind == GTOK_GEN_SFN) h.m = min(1 + (int)mulhi_u32(w2, (unsigned)s.num_blocks), h.n - 1);   // attachments per new node
  return h;
}

// pair index k (row-major over u < v) -> (u, v)
__device__ __forceinline__ void er_pair(long long k, int n, int& u, int& v) {
  const double t = 2.0 * n - 1.0;
  int r = (int)((t - sqrt(t * t - 8.0 * (double)k)) * 0.5);
  r = max(0, min(r, n - 2));
  auto first = [&](int a) { return (long long)a * (2ll * n - a - 1) / 2; };      // pairs in rows < a
  while (r + 1 <= n - 2 && first(r + 1) <= k) ++r;
  while (first(r) > k) --r;
  u = r;
  v = r + 1 + (int)(k - first(r));
}

template <bool FILL>
__global__ void __launch_bounds__(kTaskThreads)
er_kernel(GtokErSpec s, int64_t G, int32_t* __restrict__ nn_out, int32_t* __restrict__ ne_out,
          int32_t* __restrict__ qu_out, int32_t* __restrict__ qv_out, const int64_t* __restrict__ edge_off,
          int32_t* __restrict__ src, int32_t* __restrict__ dst) {
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const unsigned k0 = (unsigned)s.seed, k1 = (unsigned)(s.seed >> 32);
  for (int64_t g = warp0; g < G; g += nwarps) {
    const unsigned long long gid = (unsigned long long)(s.graph_id_base + g);
    const ErHeader h = er_header(s, gid);
    const long long P = (long long)h.n * (h.n - 1) / 2;
    const long long blocks = h.m ? 0 : (P + 3) >> 2;     // ba: closed-form count, filled by ba_fill_kernel
    long long base = FILL ? __ldg(edge_off + g) : (h.m ? (long long)h.m * (h.n - h.m) : 0);
    for (long long b0 = 0; b0 < blocks; b0 += kWarp) {
      const long long blk = b0 + lane;
      unsigned bits = 0;
      if (blk < blocks) {
        unsigned o[4];
        philox10((unsigned)gid, (unsigned)(gid >> 32), (unsigned)(2 + blk), 1u, k0, k1, o[0], o[1], o[2], o[3]);
        const long long kk = blk << 2;
        if (s.kind == GTOK_GEN_ER) {
          bits = (o[0] < h.thr ? 1u : 0u) | (kk + 1 < P && o[1] < h.thr ? 2u : 0u) | (kk + 2 < P && o[2] < h.thr ? 4u : 0u) |
                 (kk + 3 < P && o[3] < h.thr ? 8u : 0u);
        } else {                                         // the other families decide per (u, v)
          int u, v;
          er_pair(kk, h.n, u, v);
          const unsigned thr_out = (unsigned)(((unsigned long long)h.thr * s.out_ratio_q16) >> 16);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            bool keep;
            if (s.kind == GTOK_GEN_SBM)
              keep = o[i] < (((long long)u * s.num_blocks) / h.n == ((long long)v * s.num_blocks) / h.n ? h.thr : thr_out);
            else if (s.kind == GTOK_GEN_COMPLETE) keep = true;
            else if (s.kind == GTOK_GEN_STAR) keep = u == 0;
            else keep = v == u + 1;
            if (kk + i < P && keep) bits |= 1u << i;
            if (++v == h.n) { ++u; v = u + 1; }
          }
        }
      }
      const int cnt = __popc(bits);
      int incl = cnt;
#pragma unroll
      for (int dlt = 1; dlt < kWarp; dlt <<= 1) {
        const int o = __shfl_up_sync(kFull, incl, dlt);
        if (lane >= dlt) incl += o;
      }
      if (FILL && bits) {
        int u, v;
        er_pair(blk << 2, h.n, u, v);
        long long at = base + incl - cnt;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (bits & (1u << i)) { src[at] = u; dst[at] = v; ++at; }
          if (++v == h.n) { ++u; v = u + 1; }
        }
      }
      base += __shfl_sync(kFull, incl, kWarp - 1);
    }
    if (!FILL && lane == 0) {
      nn_out[g] = h.n;
      ne_out[g] = (int)base;
      if (qu_out) qu_out[g] = h.qu;
      if (qv_out) qv_out[g] = h.qv;
    }
  }
}

// Barabasi-Albert graphs (preferential attachment is sequential per graph): a thread per graph, adjacency as 64-bit
// rows in shared memory (so the edges come out sorted like every other family), degree-proportional targets drawn by
// walking the cumulative degrees.  Stream definition: oracle/tasks_oracle.py er_graph(kind="ba").
constexpr int kBaThreads = 64;
__global__ void __launch_bounds__(kBaThreads)
ba_fill_kernel(GtokErSpec s, int64_t G, const int64_t* __restrict__ edge_off, int32_t* __restrict__ src,
               int32_t* __restrict__ dst) {
  __shared__ unsigned long long rows[64][kBaThreads];
  __shared__ unsigned char deg[64][kBaThreads];
  const int t = threadIdx.x;
  const unsigned k0 = (unsigned)s.seed, k1 = (unsigned)(s.seed >> 32);
  for (int64_t g = (int64_t)blockIdx.x * kBaThreads + t; g < G; g += (int64_t)gridDim.x * kBaThreads) {
    const unsigned long long gid = (unsigned long long)(s.graph_id_base + g);
    const ErHeader h = er_header(s, gid);
    const int n = h.n, m = h.m;
    for (int i = 0; i < n; ++i) { rows[i][t] = 0ull; deg[i][t] = 0; }
    auto add_edge = [&](int a, int b) {
      rows[a][t] |= 1ull << b; rows[b][t] |= 1ull << a;
      ++deg[a][t]; ++deg[b][t];
    };
    for (int i = 1; i <= m; ++i) add_edge(0, i);                 // star on nodes 0..m
    unsigned e2 = 2u * (unsigned)m;                              // sum of degrees
    unsigned d = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0, have = 0xffffffffu;
    auto next_word = [&]() -> unsigned {
      const unsigned wi = 8u + d, blk = wi >> 2;
      ++d;
      if (blk != have) { philox10((unsigned)gid, (unsigned)(gid >> 32), blk, 1u, k0, k1, b0, b1, b2, b3); have = blk; }
      return (wi & 2u) ? ((wi & 1u) ? b3 : b2) : ((wi & 1u) ? b1 : b0);
    };
    for (int source = m + 1; source < n; ++source) {
      unsigned long long tm = 0ull;
      int cnt = 0;
      while (cnt < m) {
        const unsigned word = next_word();
        const unsigned r = __umulhi(word, e2);
        int x = 0;
        unsigned acc = deg[0][t];
        while (acc <= r) { ++x; acc += deg[x][t]; }
        if (!((tm >> x) & 1ull)) { tm |= 1ull << x; ++cnt; }
      }
      if (s.kind == GTOK_GEN_SFN) {                               // Holme-Kim: triad steps between the preferential ones
        int target = __ffsll((long long)tm) - 1;
        tm &= tm - 1;
        add_edge(target, source);
        for (int count = 1; count < m; ++count) {
          const unsigned a = next_word();
          bool done = false;
          if ((a >> 16) < s.out_ratio_q16) {
            const unsigned long long cand = rows[target][t] & ~rows[source][t] & ~(1ull << source);
            if (cand) {
              const unsigned b = next_word();
              add_edge(select64(cand, (int)__umulhi(b, (unsigned)__popcll(cand))), source);
              done = true;
            }
          }
          if (!done) {
            tm &= ~rows[source][t];
            target = __ffsll((long long)tm) - 1;
            tm &= tm - 1;
            add_edge(target, source);
          }
        }
      } else {
        while (tm) {
          const int x = __ffsll((long long)tm) - 1;
          tm &= tm - 1;
          add_edge(x, source);
        }
      }
      e2 += 2u * (unsigned)m;
    }
    long long at = __ldg(edge_off + g);
    for (int u = 0; u < n; ++u) {
      unsigned long long up = u < 63 ? (rows[u][t] >> (u + 1)) << (u + 1) : 0ull;
      while (up) {
        const int v = __ffsll((long long)up) - 1;
        up &= up - 1;
        src[at] = u; dst[at] = v; ++at;
      }
    }
  }
}

static int task_grid(int64_t warps_needed) {
  const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 8;
  int64_t blocks = (warps_needed + kTaskThreads / 32 - 1) / (kTaskThreads / 32);
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

static int er_check(const GtokErSpec* s, int64_t G, const char* who) {
  GTOK_REQUIRE(s != nullptr, GTOK_EINVAL, "%s: null spec", who);
  GTOK_REQUIRE(G >= 0, GTOK_EINVAL, "%s: negative graph count", who);
  GTOK_REQUIRE(s->num_buckets >= 1 && s->num_buckets <= 4, GTOK_EINVAL, "%s: 1..4 size buckets expected (got %d)", who,
               s->num_buckets);
  for (int b = 0; b < s->num_buckets; ++b)
    GTOK_REQUIRE(s->n_lo[b] >= 1 && s->n_lo[b] < s->n_hi[b] && s->n_hi[b] <= 32768, GTOK_EINVAL,
                 "%s: bucket %d needs 1 <= n_lo < n_hi <= 32768 (got [%d, %d))", who, b, s->n_lo[b], s->n_hi[b]);
  GTOK_REQUIRE(s->thr_lo <= s->thr_hi, GTOK_EINVAL, "%s: thr_lo > thr_hi", who);
  GTOK_REQUIRE(s->kind >= GTOK_GEN_ER && s->kind <= GTOK_GEN_SFN, GTOK_EINVAL, "%s: unknown generator kind %d", who, s->kind);
  if (s->kind == GTOK_GEN_BA || s->kind == GTOK_GEN_SFN) {
    GTOK_REQUIRE(s->num_blocks >= 1 && s->num_blocks <= 8, GTOK_EINVAL, "%s: ba / sfn need 1..8 attachments per node (num_blocks)", who);
    GTOK_REQUIRE(s->out_ratio_q16 <= 65536u, GTOK_EINVAL, "%s: sfn triad probability above 1", who);
    for (int b = 0; b < s->num_buckets; ++b)
      GTOK_REQUIRE(s->n_hi[b] <= 65, GTOK_EINVAL, "%s: ba / sfn graphs are limited to 64 nodes (bucket %d ends at %d)", who, b, s->n_hi[b]);
  }
  GTOK_REQUIRE(s->kind != GTOK_GEN_SBM || (s->num_blocks >= 1 && s->num_blocks <= 16 && s->out_ratio_q16 <= 65536u), GTOK_EINVAL,
               "%s: sbm needs 1..16 blocks and out_ratio_q16 <= 65536", who);
  return GTOK_OK;
}

}  // namespace gtok

extern "C" int gtok_find_token(const int64_t* X, const int64_t* xoff, const int32_t* lmax, int64_t num_rows,
                               int64_t batch_size, const int64_t* token_per_batch, int64_t token, int32_t* pos_out,
                               void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(num_rows >= 0, GTOK_EINVAL, "find_token: negative row count");
  if (num_rows == 0) return GTOK_OK;
  GTOK_REQUIRE(batch_size > 0, GTOK_EINVAL, "find_token: batch size must be positive");
  GTOK_REQUIRE(X && xoff && lmax && pos_out, GTOK_EINVAL, "find_token: null array");
  find_token_kernel<<<task_grid(num_rows), kTaskThreads, 0, (cudaStream_t)stream>>>(X, xoff, lmax, num_rows, batch_size,
                                                                                     token_per_batch, token, pos_out);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

extern "C" int gtok_graph_labels(const GtokCsrBatch* csr, const int64_t* edge_off, int32_t* has_cycle_out,
                                 int32_t* dist_out, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(csr != nullptr, GTOK_EINVAL, "graph_labels: null csr");
  GTOK_REQUIRE(csr->num_graphs >= 0, GTOK_EINVAL, "graph_labels: negative graph count");
  if (csr->num_graphs == 0 || (!has_cycle_out && !dist_out)) return GTOK_OK;
  GTOK_REQUIRE(csr->node_off && csr->rowptr, GTOK_EINVAL, "graph_labels: null csr arrays");   // col may be null without edges
  GTOK_REQUIRE(!dist_out || (csr->query_u && csr->query_v), GTOK_EINVAL, "graph_labels: distances need query_u/query_v");
  GTOK_REQUIRE(csr->max_nodes > 0, GTOK_EINVAL, "graph_labels: max_nodes bound required");
  cudaStream_t st = (cudaStream_t)stream;
  if (csr->adj64 && csr->adj_words <= 1 && csr->max_nodes <= 64) {
    const int64_t blocks = (csr->num_graphs + kTaskThreads - 1) / kTaskThreads;
    const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 8;
    labels_small_kernel<<<(unsigned)(blocks < cap ? blocks : cap), kTaskThreads, 0, st>>>(*csr, edge_off, has_cycle_out,
                                                                                         dist_out);
  } else {
    const int words = (csr->max_nodes + 31) / 32;
    int dev = 0, optin = 0;
    GTOK_CUDA_TRY(cudaGetDevice(&dev));
    GTOK_CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    int threads = kTaskThreads;
    while (threads > 32 && (size_t)(threads / 32) * 3 * words * 4 > (size_t)optin) threads >>= 1;
    const size_t smem = (size_t)(threads / 32) * 3 * words * 4;
    GTOK_REQUIRE(smem <= (size_t)optin, GTOK_EUNSUPPORTED, "graph_labels: graphs of %d nodes exceed the shared-memory bitmaps",
                 csr->max_nodes);
    GTOK_CUDA_TRY(cudaFuncSetAttribute(labels_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_cta = threads / 32;
    int64_t blocks = (csr->num_graphs + per_cta - 1) / per_cta;
    const int grid = persistent_grid(labels_warp_kernel, threads, smem, blocks);
    labels_warp_kernel<<<grid, threads, smem, st>>>(*csr, edge_off, has_cycle_out, dist_out, words);
  }
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

extern "C" int gtok_er_sizes(const GtokErSpec* spec, int64_t num_graphs, int32_t* num_nodes_out, int32_t* num_edges_out,
                             int32_t* query_u_out, int32_t* query_v_out, void* stream) {
  using namespace gtok;
  if (int rc = er_check(spec, num_graphs, "er_sizes")) return rc;
  if (num_graphs == 0) return GTOK_OK;
  GTOK_REQUIRE(num_nodes_out && num_edges_out, GTOK_EINVAL, "er_sizes: null output");
  er_kernel<false><<<task_grid(num_graphs), kTaskThreads, 0, (cudaStream_t)stream>>>(
      *spec, num_graphs, num_nodes_out, num_edges_out, query_u_out, query_v_out, nullptr, nullptr, nullptr);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

extern "C" int gtok_er_edges(const GtokErSpec* spec, int64_t num_graphs, const int64_t* edge_off, int32_t* src_out,
                             int32_t* dst_out, void* stream) {
  using namespace gtok;
  if (int rc = er_check(spec, num_graphs, "er_edges")) return rc;
  if (num_graphs == 0) return GTOK_OK;
  GTOK_REQUIRE(edge_off != nullptr, GTOK_EINVAL, "er_edges: null edge_off");          // src / dst may be null without edges
  if (spec->kind == GTOK_GEN_BA || spec->kind == GTOK_GEN_SFN) {
    const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 4;
    const int64_t blocks = (num_graphs + kBaThreads - 1) / kBaThreads;
    ba_fill_kernel<<<(unsigned)(blocks < cap ? blocks : cap), kBaThreads, 0, (cudaStream_t)stream>>>(*spec, num_graphs, edge_off,
                                                                                                   src_out, dst_out);
  } else {
    er_kernel<true><<<task_grid(num_graphs), kTaskThreads, 0, (cudaStream_t)stream>>>(
        *spec, num_graphs, nullptr, nullptr, nullptr, nullptr, edge_off, src_out, dst_out);
  }
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}
