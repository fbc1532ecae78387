// gtok_sent.cu - K3: SENT (Segmented Eulerian Neighborhood Trail) tokenization.
//
// Replaces (reference, per graph, EVERY epoch, in the training process):
//   Graph2TrailTokenizer.__call__   external BorgwardtLab/AutoGraph, call site
//                                   trainer/train_agtt.py:250 (ctor :514-542)
//   remap_zinc_tokens               trainer/train_agtt.py:171-244   (fused, optional)
//   <q> u v append                  trainer/train_agtt.py:257-267   (fused, optional)
// The walk itself follows the normative spec in oracle/sent_oracle.py (AutoGraph
// is not vendored by the reference: parity to AutoGraph is UNPINNED, parity to
// the oracle is bit-exact under the same choice words).
//
// One warp per graph; the graph's canonical CSR, visit indices, per-node
// "unvisited neighbours left" counters, a visit-index bitmap and the staged token
// row all live in that warp's slice of shared memory, so the dependent walk makes
// no global access between the coalesced CSR load and the coalesced 16-byte
// write-out.  Neighbour selection is ballot + popc ranking; the bracketed
// neighbourhood set is emitted in ascending visit index by setting bits in the
// visit-index bitmap and enumerating them with a warp prefix sum (no sort).
// Random words: Philox-4x32-10 keyed by the seed, counter = (global graph id,
// decision >> 2), or an explicit per-graph choice stream (deterministic replay).
// Algorithmic bytes / graph: 4(N+1) + 4 nnz (CSR) [+4N + 4 nnz labels] + 8 L.
#include <stdlib.h>

#include "gtok_sent_common.cuh"

namespace gtok {

__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0,
                                              unsigned k1, unsigned (&o)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// trainer/train_agtt.py:193-242 (branch order preserved; oracle: remap_zinc_token)
__device__ __forceinline__ int remap_zinc(int t, int idx_off, int node_off, int edge_off) {
  if (t == 0) return 0;
  if (t == 1 || t == 2 || t == 3 || t == 5) return 2;
  if (t == 4) return 1;
  if (t >= node_off && t < edge_off) {
    const int a = t - node_off;
    return (a >= 0 && a < 9) ? 8 + a : 22 + t;
  }
  if (t >= edge_off) {
    const int b = t - edge_off + 1;
    return (b >= 1 && b <= 4) ? 17 + (b - 1) : 22 + t;
  }
  if (t >= idx_off && t < node_off) return 22 + (t - idx_off);
  return 22 + t;
}

__device__ __forceinline__ int ld_keep_l2(const int32_t* p, unsigned long long policy) {
  int v;
  asm volatile("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
  return v;
}

struct SentSmemLayout {
  int rp, stage, ea, ent, bits, openb, visb, vis, rem, col, total;  // byte offsets inside one warp's slice
};

// csr_global: rowptr / col / eattr stay in HBM (read through L1/L2) - for graphs whose CSR would leave room for
// only a couple of warps per SM (a 4096-node, 32k-entry graph needs 100 KB staged, 18 KB this way).
static SentSmemLayout sent_layout(int nmax, int zmax, int lcap, bool labeled, bool staged, bool csr_global) {
  SentSmemLayout L;
  int o = 0;
  const int nwords = (nmax + 31) / 32;
  L.rp = o;    o += csr_global ? 0 : 4 * (nmax + 1);
  L.stage = o; o += staged ? 4 * lcap : 0;
  L.ea = o;    o += (labeled && !csr_global) ? 4 * zmax : 0;
  L.ent = o;   o += labeled ? 4 * nmax : 0;
  L.bits = o;  o += 4 * nwords;
  L.openb = o; o += 4 * nwords;     // visited nodes that still have an unvisited neighbour
  L.visb = o;  o += 4 * nwords;     // visited nodes
  L.vis = o;   o += 2 * nmax;
  L.rem = o;   o += 2 * nmax;
  L.col = o;   o += csr_global ? 0 : 2 * zmax;
  L.total = (o + 15) & ~15;
  return L;
}

template <bool LABELED, bool STAGED, bool CSRG>
__global__ void sent_kernel(const GtokCsrBatch c, const GtokSentCfg cfg, const GtokChoice ch,
                            const SentSmemLayout lay, int32_t* __restrict__ len_out, int64_t* __restrict__ out,
                            int64_t stride) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = lane_id();
  const bool restart_unvisited = cfg.restart_policy == GTOK_SENT_RESTART_UNVISITED;
  const bool bracket_desc = cfg.bracket_order == GTOK_SENT_BRACKET_VISIT_DESC;
  const int w = threadIdx.x >> 5;
  unsigned char* base = smem + (size_t)w * lay.total;
  int* rp = reinterpret_cast<int*>(base + lay.rp);
  int* stage = reinterpret_cast<int*>(base + lay.stage);
  int* ea = reinterpret_cast<int*>(base + lay.ea);
  int* ent = reinterpret_cast<int*>(base + lay.ent);
  unsigned* bits = reinterpret_cast<unsigned*>(base + lay.bits);
  unsigned* openb = reinterpret_cast<unsigned*>(base + lay.openb);
  unsigned* visb = reinterpret_cast<unsigned*>(base + lay.visb);
  short* vis = reinterpret_cast<short*>(base + lay.vis);
  unsigned short* rem = reinterpret_cast<unsigned short*>(base + lay.rem);
  unsigned short* col = reinterpret_cast<unsigned short*>(base + lay.col);

  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int Lcap = cfg.max_length;
  const int per = LABELED ? 2 : 1;
  const unsigned k0 = (unsigned)(ch.seed & 0xffffffffu), k1 = (unsigned)(ch.seed >> 32);
  unsigned long long l2_keep = 0;
  if (CSRG) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(l2_keep));

  for (int64_t g = warp0; g < c.num_graphs; g += nwarps) {
    const int64_t n0 = __ldg(c.node_off + g);
    const int N = (int)(__ldg(c.node_off + g + 1) - n0);
    int64_t* __restrict__ row = out + g * stride;

    // ---- stage the graph (coalesced) -----------------------------------
    const int z0 = __ldg(c.rowptr + n0);
    const int32_t* __restrict__ g_rp = c.rowptr + n0;
    const int32_t* __restrict__ g_col = c.col + z0;
    const int32_t* __restrict__ g_ea = LABELED ? c.eattr + z0 : nullptr;
    // CSR in HBM: the row pointers of the resident graphs (4 bytes per node) are the one array the walk revisits
    // at random - keep them in L2 (evict_last) and let neighbour lists / labels stream through (evict_first)
    auto RP = [&](int i) -> int { return CSRG ? ld_keep_l2(g_rp + i, l2_keep) - z0 : rp[i]; };
    auto COL = [&](int e) -> int { return CSRG ? __ldcs(g_col + e) : (int)col[e]; };
    auto EA = [&](int e) -> int { return CSRG ? __ldcs(g_ea + e) : ea[e]; };
    if (!CSRG) for (int i = lane; i <= N; i += kWarp) rp[i] = __ldg(g_rp + i) - z0;
    for (int i = lane; i < N; i += kWarp) { vis[i] = -1; }
    for (int i = lane; i < ((N + 31) >> 5); i += kWarp) { openb[i] = 0u; visb[i] = 0u; }
    __syncwarp();
    if (!CSRG) {
      const int Z = rp[N];
      for (int i = lane; i < Z; i += kWarp) {
        col[i] = (unsigned short)__ldg(g_col + i);
        if (LABELED) ea[i] = __ldg(g_ea + i);
      }
    }
    for (int i = lane; i < N; i += kWarp) rem[i] = (unsigned short)(RP(i + 1) - RP(i));
    __syncwarp();
    int nopen = 0;                     // popcount of openb (warp-uniform)

    // ---- random words ----------------------------------------------------
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
        if ((d & 3u) == 0u) philox4x32_10((unsigned)gid, (unsigned)(gid >> 32), d >> 2, 0u, k0, k1, blk);
        const unsigned w01 = (d & 1u) ? blk[1] : blk[0], w23 = (d & 1u) ? blk[3] : blk[2];
        r = (d & 2u) ? w23 : w01;
      }
      ++d;
      return (int)__umulhi(r, (unsigned)cnt);
    };
    auto put = [&](int pos, int t) {
      if (pos < Lcap) {
        if (STAGED) stage[pos] = t;
        else __stcs(row + pos, (int64_t)(cfg.remap_zinc ? remap_zinc(t, cfg.idx_offset, cfg.node_idx_offset, cfg.edge_idx_offset) : t));
      }
    };
    // k-th (0-based) set bit of an N-bit bitmap (or of its complement), ascending node id; the caller knows
    // there are more than k.  128 words for 4096 nodes = 4 trips (the O(N) scans over vis[]/rem[] this replaces
    // were most of the time on large graphs).
    auto locate = [&](const unsigned* bm, bool invert, int k) -> int {
      const int nwords = (N + 31) >> 5;
      int seen = 0, found = -1;
      for (int w0 = 0; w0 < nwords && found < 0; w0 += kWarp) {
        const int wi = w0 + lane;
        unsigned word = wi < nwords ? bm[wi] : 0u;
        if (invert) {
          word = wi < nwords ? ~word : 0u;
          if (wi == nwords - 1 && (N & 31)) word &= (1u << (N & 31)) - 1u;
        }
        const int cw = __popc(word);
        int incl = cw;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
          const int o = __shfl_up_sync(kFull, incl, dd);
          if (lane >= dd) incl += o;
        }
        const int excl = seen + incl - cw;
        const bool mine = k >= excl && k < excl + cw;
        const unsigned who = __ballot_sync(kFull, mine);
        if (who) {
          int pos = -1;
          if (mine) {
            int kk = k - excl;
            unsigned t = word;
            while (kk--) t &= t - 1;
            pos = (wi << 5) + __ffs(t) - 1;
          }
          found = __shfl_sync(kFull, pos, __ffs(who) - 1);
        }
        seen += __shfl_sync(kFull, incl, 31);
      }
      return found;
    };

    int len = 0;
    if (lane == 0) put(0, SOS);
    len = 1;
    if (N > 0) {
      int nvis = 0, prev = -1;
      int cur = pick(N);
      int lo = RP(cur), hi = RP(cur + 1);
      while (len < Lcap) {
        const int vc = vis[cur];
        if (hi - lo <= kWarp) {
          // ---- fast path: the whole row sits in the lanes (lane <-> neighbour).  The row extents of every
          // neighbour are requested right away and the rows of the candidates are prefetched, so the next step
          // finds its (lo, hi) in a register and its row in L1 / L2: one dependent HBM round trip per step
          // instead of three (rowptr, col, col again).
          const int e = lo + lane;
          const bool valid = e < hi;
          const int u = valid ? COL(e) : 0;
          int rlo = 0, rhi = 0, eav = 0;
          if (valid) {
            rlo = RP(u); rhi = RP(u + 1);
            if (LABELED) eav = EA(e);
          }
          const int vu = valid ? (int)vis[u] : 0;
          if (vc < 0) {
            // first visit: node token [+ label], then the bracketed set of visited neighbours (not the one we came
            // from) in ascending visit index - ranked inside the warp, no bitmap
            const int nv = nvis;
            if (lane == 0) {
              vis[cur] = (short)nv;
              visb[cur >> 5] |= 1u << (cur & 31);
              put(len, cfg.idx_offset + nv);
              if (LABELED) put(len + 1, cfg.node_idx_offset + __ldg(c.x + n0 + cur));
            }
            len += per;
            ++nvis;
            bool z = false;
            if (valid) {
              const int r = (int)rem[u] - 1;                      // cur stops being an unvisited neighbour of u
              rem[u] = (unsigned short)r;
              z = vu >= 0 && r == 0;
              if (z) atomicAnd(&openb[u >> 5], ~(1u << (u & 31)));
            }
            nopen -= __popc(__ballot_sync(kFull, z));
            const bool member = valid && vu >= 0 && u != prev;
            const unsigned mm = __ballot_sync(kFull, member);
            if (mm) {
              const int M = __popc(mm);
              // rank among the members by ascending visit index: one warp min-reduction per member (visit indices
              // are distinct), the lane that holds the minimum takes the next rank and drops out
              int rank = 0;
              unsigned key = member ? (unsigned)vu : 0xffffffffu;
              for (int j = 0; j < M; ++j) {
                const unsigned mn = __reduce_min_sync(kFull, key);
                if (key == mn) { rank = j; key = 0xffffffffu; }
              }
              if (member) {
                const int pos = len + 1 + (bracket_desc ? M - 1 - rank : rank) * per;
                if (LABELED) {
                  put(pos, cfg.edge_idx_offset + eav + cfg.edge_label_shift);
                  put(pos + 1, cfg.idx_offset + vu);
                } else {
                  put(pos, cfg.idx_offset + vu);
                }
              }
              if (lane == 0) { put(len, LADJ); put(len + 1 + M * per, RADJ); }
              len += 2 + M * per;
            }
          } else {
            if (lane == 0) put(len, cfg.idx_offset + vc);        // head of a new segment at a visited node
            len += 1;
          }
          const unsigned fm = __ballot_sync(kFull, valid && vu < 0);   // unvisited neighbours, CSR order
          const int cnt = __popc(fm);                                  // == rem[cur]
          if (cnt > 0) {
            if (vc < 0) {                                              // cur joins the restart candidates
              if (lane == 0) atomicOr(&openb[cur >> 5], 1u << (cur & 31));   // (other lanes update this map with atomics in the same step)
              ++nopen;
            }
            if (CSRG && ((fm >> lane) & 1u)) {
              asm volatile("prefetch.global.L1 [%0];" ::"l"(g_col + rlo));
              if (LABELED) asm volatile("prefetch.global.L1 [%0];" ::"l"(g_ea + rlo));
            }
            const int kth = pick(cnt);                                  // lane holding the kth candidate
            const int sl = __ffs(__ballot_sync(kFull, ((fm >> lane) & 1u) && __popc(fm & ((1u << lane) - 1u)) == kth)) - 1;
            if (LABELED) {
              const int eas = __shfl_sync(kFull, eav, sl);
              if (lane == 0) put(len, cfg.edge_idx_offset + eas + cfg.edge_label_shift);
              len += 1;
            }
            prev = cur;
            cur = __shfl_sync(kFull, u, sl);
            lo = __shfl_sync(kFull, rlo, sl);
            hi = __shfl_sync(kFull, rhi, sl);
            __syncwarp();
            continue;
          }
          __syncwarp();
          if (nvis == N) break;
          // ---- restart: visited nodes with unvisited neighbours, else unvisited nodes
          int nxt;
          if (nopen > 0 && !restart_unvisited) nxt = locate(openb, false, pick(nopen));
          else nxt = locate(visb, true, pick(N - nvis));
          if (lane == 0) put(len, RESET);
          len += 1;
          cur = nxt;
          prev = -1;
          lo = RP(cur); hi = RP(cur + 1);
          continue;
        }
        // ---- rows longer than a warp: chunked row scans, bracket members ranked through a visit-index bitmap
        {
          if (vc < 0) {
            // ---- first visit: node token [+ label], then bracketed neighbourhood set
            const int nv = nvis;
            const int nwords = (nv + 31) >> 5;
            for (int i = lane; i < nwords; i += kWarp) bits[i] = 0u;
            if (lane == 0) {
              vis[cur] = (short)nv;
              visb[cur >> 5] |= 1u << (cur & 31);
              put(len, cfg.idx_offset + nv);
              if (LABELED) put(len + 1, cfg.node_idx_offset + __ldg(c.x + n0 + cur));
            }
            len += per;
            ++nvis;
            __syncwarp();
            int closed = 0;
            for (int e0 = lo; e0 < hi; e0 += kWarp) {
              const int e = e0 + lane;
              bool z = false;
              if (e < hi) {
                const int u = COL(e);
                const int vu = vis[u];
                const int r = (int)rem[u] - 1;                    // cur stops being an unvisited neighbour of u
                rem[u] = (unsigned short)r;
                if (vu >= 0 && u != cur) {
                  if (r == 0) { z = true; atomicAnd(&openb[u >> 5], ~(1u << (u & 31))); }
                  if (u != prev) {
                    atomicOr(&bits[vu >> 5], 1u << (vu & 31));
                    if (LABELED) ent[vu] = e;
                  }
                }
              }
              closed += __popc(__ballot_sync(kFull, z));
            }
            nopen -= closed;
            __syncwarp();
            if (rem[cur] > 0) {                                   // cur joins the restart candidates
              if (lane == 0) atomicOr(&openb[cur >> 5], 1u << (cur & 31));   // (other lanes update this map with atomics in the same step)
              ++nopen;
            }
            int M = 0;
            int Mtot = 0;                                         // descending brackets: positions count from the end
            if (bracket_desc) {
              for (int i = lane; i < nwords; i += kWarp) Mtot += __popc(bits[i]);
              Mtot = __reduce_add_sync(kFull, Mtot);
            }
            for (int w0 = 0; w0 < nwords; w0 += kWarp) {
              unsigned word = (w0 + lane < nwords) ? bits[w0 + lane] : 0u;
              const int cw = __popc(word);
              int incl = cw;
  #pragma unroll
              for (int dd = 1; dd < 32; dd <<= 1) {
                const int o = __shfl_up_sync(kFull, incl, dd);
                if (lane >= dd) incl += o;
              }
              int j = M + incl - cw;
              while (word) {
                const int bpos = __ffs(word) - 1;
                word &= word - 1;
                const int vi = ((w0 + lane) << 5) + bpos;
                const int pos = len + 1 + (bracket_desc ? Mtot - 1 - j : j) * per;
                if (LABELED) {
                  put(pos, cfg.edge_idx_offset + EA(ent[vi]) + cfg.edge_label_shift);
                  put(pos + 1, cfg.idx_offset + vi);
                } else {
                  put(pos, cfg.idx_offset + vi);
                }
                ++j;
              }
              M += __shfl_sync(kFull, incl, 31);
            }
            if (M > 0) {
              if (lane == 0) { put(len, LADJ); put(len + 1 + M * per, RADJ); }
              len += 2 + M * per;
            }
          } else {
            if (lane == 0) put(len, cfg.idx_offset + vc);  // head of a new segment at a visited node
            len += 1;
          }
          __syncwarp();
          const int cnt = rem[cur];
          if (cnt > 0) {
            // ---- trail step to the k-th unvisited neighbour (CSR order)
            const int k = pick(cnt);
            int seen = 0, ne = -1;
            for (int e0 = lo; e0 < hi && ne < 0; e0 += kWarp) {
              const int e = e0 + lane;
              const bool f = e < hi && vis[COL(e)] < 0;
              const unsigned m = __ballot_sync(kFull, f);
              const int cch = __popc(m);
              if (k < seen + cch) {
                const unsigned sel = __ballot_sync(kFull, f && (__popc(m & ((1u << lane) - 1)) == k - seen));
                ne = e0 + __ffs(sel) - 1;
              }
              seen += cch;
            }
            if (LABELED) {
              if (lane == 0) put(len, cfg.edge_idx_offset + EA(ne) + cfg.edge_label_shift);
              len += 1;
            }
            prev = cur;
            cur = COL(ne);
            lo = RP(cur); hi = RP(cur + 1);
          } else {
            if (nvis == N) break;
            // ---- restart: visited nodes with unvisited neighbours, else unvisited nodes
            int nxt;
            if (nopen > 0 && !restart_unvisited) nxt = locate(openb, false, pick(nopen));
            else nxt = locate(visb, true, pick(N - nvis));
            if (lane == 0) put(len, RESET);
            len += 1;
            cur = nxt;
            prev = -1;
            lo = RP(cur); hi = RP(cur + 1);
          }
        }
      }
    }
    if (!cfg.no_eos) {
      if (lane == 0) put(len, EOS);
      len += 1;
    }
    const int L = len < Lcap ? len : Lcap;
    __syncwarp();

    // ---- write-out ---------------------------------------------------------
    const int Ltot = L + (cfg.append_query ? 3 : 0);
    int qn = 0, qu = 0, qv = 0;
    if (cfg.append_query) {
      qn = cfg.idx_offset + N;
      qu = cfg.idx_offset + __ldg(c.query_u + g);
      qv = cfg.idx_offset + __ldg(c.query_v + g);
    }
    if (STAGED) {
      warp_write_row(row, Ltot, [&](int p) -> int {
        if (p < L) {
          const int t = stage[p];
          return cfg.remap_zinc ? remap_zinc(t, cfg.idx_offset, cfg.node_idx_offset, cfg.edge_idx_offset) : t;
        }
        const int q = p - L;
        return q == 0 ? qn : (q == 1 ? qu : qv);
      });
    } else if (cfg.append_query && lane < 3) {
      row[L + lane] = lane == 0 ? qn : (lane == 1 ? qu : qv);
    }
    if (lane == 0) len_out[g] = Ltot;
    __syncwarp();
  }
}

template <bool LABELED, bool STAGED, bool CSRG>
static int sent_launch(const GtokCsrBatch* c, const GtokSentCfg* cfg, const GtokChoice* ch, int32_t* len_out,
                       int64_t* out, int64_t stride, cudaStream_t st) {
  const SentSmemLayout lay = sent_layout(c->max_nodes, c->max_nnz, cfg->max_length, LABELED, STAGED, CSRG);
  auto k = sent_kernel<LABELED, STAGED, CSRG>;
  int dev = 0, max_optin = 0;
  GTOK_CUDA_TRY(cudaGetDevice(&dev));
  GTOK_CUDA_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  int warps = 4;
  while (warps > 1 && (size_t)warps * lay.total > (size_t)max_optin) warps >>= 1;
  const size_t smem = (size_t)warps * lay.total;
  GTOK_REQUIRE(smem <= (size_t)max_optin, GTOK_EUNSUPPORTED,
               "sent: one graph needs %d B of shared memory (max_nodes=%d max_nnz=%d max_length=%d), device allows %d",
               lay.total, c->max_nodes, c->max_nnz, cfg->max_length, max_optin);
  GTOK_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t work = (c->num_graphs + warps - 1) / warps;
  const int grid = persistent_grid(k, warps * 32, smem, work);
  k<<<grid, warps * 32, smem, st>>>(*c, *cfg, *ch, lay, len_out, out, stride);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

int sent_small_dispatch(const GtokCsrBatch* c, const GtokSentCfg* cfg, const GtokChoice* ch, int32_t* len_out,
                        int64_t* out, int64_t stride, cudaStream_t st, bool* handled);
int sent_walk_dispatch(const GtokCsrBatch* c, const GtokSentCfg* cfg, const GtokChoice* ch, int32_t* len_out,
                       int64_t* out, int64_t stride, int64_t total_nodes, void* ws, size_t ws_bytes, cudaStream_t st,
                       bool* handled);
size_t sent_log_bytes(int64_t total_nodes);
int sent_mid_dispatch(const GtokCsrBatch* c, const GtokSentCfg* cfg, const GtokChoice* ch, int32_t* len_out,
                      int64_t* out, int64_t stride, cudaStream_t st, bool* handled);
int sent_big_dispatch(const GtokCsrBatch* c, const GtokSentCfg* cfg, const GtokChoice* ch, int32_t* len_out,
                      int64_t* out, int64_t stride, void* ws, size_t ws_bytes, int64_t total_nodes, cudaStream_t st,
                      bool* handled);

}  // namespace gtok

extern "C" size_t gtok_sent_workspace_bytes(int64_t total_nodes, int32_t max_nodes) {
  if (total_nodes <= 0) return 0;
  if (max_nodes <= 256) return gtok::sent_log_bytes(total_nodes);       // the walk's item log (gtok_sent_walk.cu / _walkl.cu)
  return max_nodes > 256 ? (size_t)total_nodes * 8 : 0;
}

extern "C" int gtok_sent_tokenize(const GtokCsrBatch* csr, const GtokSentCfg* cfg, const GtokChoice* choice,
                                  int32_t* len_out, int64_t* tok_out, int64_t stride, void* stream) {
  return gtok_sent_tokenize_ws(csr, cfg, choice, len_out, tok_out, stride, 0, nullptr, 0, stream);
}

extern "C" int gtok_sent_tokenize_ws(const GtokCsrBatch* csr, const GtokSentCfg* cfg, const GtokChoice* choice,
                                     int32_t* len_out, int64_t* tok_out, int64_t stride, int64_t total_nodes, void* ws,
                                     size_t ws_bytes, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(csr && cfg && choice, GTOK_EINVAL, "sent: null csr/cfg/choice");
  GTOK_REQUIRE(csr->num_graphs >= 0, GTOK_EINVAL, "sent: negative graph count");
  if (csr->num_graphs == 0) return GTOK_OK;
  GTOK_REQUIRE(csr->node_off && csr->rowptr && len_out && tok_out, GTOK_EINVAL, "sent: null pointer");
  GTOK_REQUIRE(csr->col != nullptr || csr->max_nnz == 0, GTOK_EINVAL, "sent: null col");
  GTOK_REQUIRE(cfg->max_length >= 1, GTOK_EINVAL, "sent: max_length must be >= 1");
  GTOK_REQUIRE(stride >= (int64_t)cfg->max_length + (cfg->append_query ? 3 : 0), GTOK_EINVAL,
               "sent: stride %lld < max_length (+3 with append_query)", (long long)stride);
  GTOK_REQUIRE(csr->max_nodes > 0 && csr->max_nnz >= 0, GTOK_EINVAL, "sent: csr.max_nodes/max_nnz bounds required");
  GTOK_REQUIRE(csr->max_nodes <= 32767, GTOK_EUNSUPPORTED, "sent: more than 32767 nodes per graph not supported");
  GTOK_REQUIRE(!cfg->labeled || (csr->x && (csr->eattr || csr->max_nnz == 0)), GTOK_EINVAL,
               "sent: labeled mode needs csr.x and csr.eattr");
  GTOK_REQUIRE(!cfg->append_query || (csr->query_u && csr->query_v), GTOK_EINVAL, "sent: append_query needs query_u/query_v");
  GTOK_REQUIRE(choice->stream == nullptr || choice->stream_off != nullptr, GTOK_EINVAL, "sent: choice stream without offsets");
  GTOK_REQUIRE(cfg->restart_policy == GTOK_SENT_RESTART_OPEN_FIRST || cfg->restart_policy == GTOK_SENT_RESTART_UNVISITED,
               GTOK_EUNSUPPORTED, "sent: restart_policy %d is not implemented by the kernels", cfg->restart_policy);
  GTOK_REQUIRE(cfg->bracket_order == GTOK_SENT_BRACKET_VISIT_ASC || cfg->bracket_order == GTOK_SENT_BRACKET_VISIT_DESC,
               GTOK_EUNSUPPORTED, "sent: bracket_order %d is not implemented by the kernels", cfg->bracket_order);
  GTOK_REQUIRE((reinterpret_cast<uintptr_t>(tok_out) & 15) == 0, GTOK_EINVAL, "sent: tok_out must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  {
    // graphs of at most 64 nodes, 64-bit adjacency masks, workspace given: thread-per-graph walk kernel + warp-per-row
    // emission kernel (gtok_sent_walk.cu); otherwise, or with GTOK_SENT_SMALL=1, the one-kernel path (gtok_sent_small.cu)
    bool handled = false;
    static const int small_env = [] { const char* e = getenv("GTOK_SENT_SMALL"); return e ? atoi(e) : 2; }();
    int rc = GTOK_OK;
    if (small_env != 1) {
      rc = sent_walk_dispatch(csr, cfg, choice, len_out, tok_out, stride, total_nodes, ws, ws_bytes, st, &handled);
      if (rc != GTOK_OK || handled) return rc;
    }
    rc = sent_small_dispatch(csr, cfg, choice, len_out, tok_out, stride, st, &handled);
    if (rc != GTOK_OK || handled) return rc;
    // 65..256 nodes: the same scheme on 2 or 4 mask words (gtok_sent_mid.cu)
    rc = sent_mid_dispatch(csr, cfg, choice, len_out, tok_out, stride, st, &handled);
    if (rc != GTOK_OK || handled) return rc;
    // larger graphs with a caller-provided workspace: walk state in HBM scratch, 3-5x the occupancy (gtok_sent_big.cu)
    rc = sent_big_dispatch(csr, cfg, choice, len_out, tok_out, stride, ws, ws_bytes, total_nodes, st, &handled);
    if (rc != GTOK_OK || handled) return rc;
  }
  const bool staged = cfg->max_length <= 2048;
  // keep the CSR in HBM when staging it would leave fewer than ~8 warps per SM
  const bool csrg = sent_layout(csr->max_nodes, csr->max_nnz, cfg->max_length, cfg->labeled != 0, staged, false).total > 24 * 1024;
  if (cfg->labeled) {
    if (csrg) return staged ? sent_launch<true, true, true>(csr, cfg, choice, len_out, tok_out, stride, st)
                            : sent_launch<true, false, true>(csr, cfg, choice, len_out, tok_out, stride, st);
    return staged ? sent_launch<true, true, false>(csr, cfg, choice, len_out, tok_out, stride, st)
                  : sent_launch<true, false, false>(csr, cfg, choice, len_out, tok_out, stride, st);
  }
  if (csrg) return staged ? sent_launch<false, true, true>(csr, cfg, choice, len_out, tok_out, stride, st)
                          : sent_launch<false, false, true>(csr, cfg, choice, len_out, tok_out, stride, st);
  return staged ? sent_launch<false, true, false>(csr, cfg, choice, len_out, tok_out, stride, st)
                : sent_launch<false, false, false>(csr, cfg, choice, len_out, tok_out, stride, st);
}
