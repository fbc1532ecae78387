// gtok_sent_big.cu - K3 for graphs above 256 nodes when the caller provides a workspace: warp per graph, the
// per-node walk state in an HBM scratch word instead of shared memory.
//
// The shared-memory kernel (gtok_sent.cu) keeps a visit index and an unvisited-neighbour counter per node in the
// warp's slice: 4 bytes per node = 16 KB at 4096 nodes, 12 warps per SM, and every walk step is a dependent chain
// of ~300 warp instructions that 3 warps per scheduler cannot overlap (ncu: 27 % issue utilisation, 2760 cycles
// per step).  Here the two fields share ONE 32-bit word per node in a scratch array,
//     state[u] = (visit index + 1) << 16 | unvisited neighbours left,
// and the first visit of a node updates all its neighbours with one atomicSub each: the returned old value carries
// both fields, so "is u visited, what is its index, did it just lose its last unvisited neighbour" cost a single
// L2 round trip per neighbour, issued for the whole row at once (lane <-> neighbour).  Shared memory holds only the
// three bitmaps (3 * N / 8 bytes per warp), so up to 64 warps per SM hide each other's latency.
// Up to 4096 nodes per graph the word is 64 bits wide and also carries the node's CSR row:
//     state[u] = row start << 37 | degree << 25 | (visit index + 1) << 12 | unvisited neighbours left,
// so the one atomic per neighbour returns the extents of the next row as well (ncu on the 32-bit variant: all
// graphs of a batch walk concurrently, the working set is the whole batch, L2 hit rate 26 %, and the separate
// rowptr reads were 38 % of all stall samples and a third of the 1.2 KB of DRAM sectors touched per step).
// Row extents of all neighbours are requested with the row and the rows of the candidates are prefetched, as in
// the fast path of gtok_sent.cu.  Rows longer than a warp take a chunked path (bracket ranks through a visit-index
// bitmap + word prefix).  Walk semantics, random words, remap and query append: identical to gtok_sent.cu / the
// oracle (bit-exact under the same choice words).
#include "gtok_sent_common.cuh"

namespace gtok {
namespace {

constexpr int kBigWarps = 8;

struct BigLayout { int bits, pref, openb, visb, total; };

BigLayout big_layout(int nmax) {
  BigLayout L;
  const int nwords = (nmax + 31) / 32;
  int o = 0;
  L.bits = o;  o += 4 * nwords;
  L.openb = o; o += 4 * nwords;
  L.visb = o;  o += 4 * nwords;
  L.pref = o;  o += 2 * nwords;
  L.total = (o + 15) & ~15;
  return L;
}

__device__ __forceinline__ unsigned ld_cg(const unsigned* p) {
  unsigned v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_cg(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// field access of the per-node state word (32-bit: vis+1 | rem; 64-bit: row start | degree | vis+1 | rem)
template <bool P64> struct St;
template <> struct St<false> {
  using T = unsigned;
  static __device__ __forceinline__ int vis(T w) { return (int)(w >> 16) - 1; }
  static __device__ __forceinline__ bool last(T w) { return (w & 0xffffu) == 1u; }       // rem == 1 before the decrement
  static __device__ __forceinline__ bool unvisited(T w) { return (w >> 16) == 0u; }
  // (the graph belongs to this warp alone and the neighbours of a row are distinct: plain load + store, no atomic)
  static __device__ __forceinline__ T dec(T* p) { const T w = ld_cg(p); __stcg(p, w - 1u); return w; }
  static __device__ __forceinline__ void set_vis(T* p, int nv) { __stcg(p, ld_cg(p) + ((unsigned)(nv + 1) << 16)); }
};
template <> struct St<true> {
  using T = unsigned long long;
  static __device__ __forceinline__ int vis(T w) { return (int)((w >> 12) & 0x1fffu) - 1; }
  static __device__ __forceinline__ bool last(T w) { return (w & 0xfffu) == 1u; }
  static __device__ __forceinline__ bool unvisited(T w) { return ((w >> 12) & 0x1fffu) == 0u; }
  static __device__ __forceinline__ T dec(T* p) { const T w = ld_cg(p); __stcg(p, w - 1ull); return w; }   // rem >= 1: no borrow
  static __device__ __forceinline__ void set_vis(T* p, int nv) { __stcg(p, ld_cg(p) + ((unsigned long long)(nv + 1) << 12)); }
  static __device__ __forceinline__ int lo(T w) { return (int)(w >> 37); }
  static __device__ __forceinline__ int hi(T w) { return (int)(w >> 37) + (int)((w >> 25) & 0xfffu); }
};

// k-th (0-based) set bit of an N-bit bitmap (or of its complement); the caller knows there are more than k
__device__ __forceinline__ int locate_bit(const unsigned* bm, bool invert, int k, int N, int lane) {
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
      if (mine) pos = (wi << 5) + select32(word, k - excl);
      found = __shfl_sync(kFull, pos, __ffs(who) - 1);
    }
    seen += __shfl_sync(kFull, incl, 31);
  }
  return found;
}

template <bool LABELED, bool P64>
__global__ void __launch_bounds__(kBigWarps * 32)
sent_big_kernel(const GtokCsrBatch c, const GtokSentCfg cfg, const GtokChoice ch, const BigLayout lay,
                typename St<P64>::T* __restrict__ state, int32_t* __restrict__ len_out, int64_t* __restrict__ out,
                int64_t stride) {
  using S = St<P64>;
  using W = typename S::T;
  extern __shared__ __align__(16) unsigned char b_smem[];
  const int lane = lane_id();
  const bool restart_unvisited = cfg.restart_policy == GTOK_SENT_RESTART_UNVISITED;
  const bool bracket_desc = cfg.bracket_order == GTOK_SENT_BRACKET_VISIT_DESC;
  unsigned char* base = b_smem + (size_t)(threadIdx.x >> 5) * lay.total;
  unsigned* bits = reinterpret_cast<unsigned*>(base + lay.bits);
  unsigned* openb = reinterpret_cast<unsigned*>(base + lay.openb);
  unsigned* visb = reinterpret_cast<unsigned*>(base + lay.visb);
  unsigned short* pref = reinterpret_cast<unsigned short*>(base + lay.pref);

  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int Lcap = cfg.max_length;
  const int per = LABELED ? 2 : 1;
  const unsigned k0 = (unsigned)(ch.seed & 0xffffffffu), k1 = (unsigned)(ch.seed >> 32);
  const bool remap = LABELED && cfg.remap_zinc;

  for (int64_t g = warp0; g < c.num_graphs; g += nwarps) {
    const int64_t n0 = __ldg(c.node_off + g);
    const int N = (int)(__ldg(c.node_off + g + 1) - n0);
    int64_t* __restrict__ row = out + g * stride;
    const int z0 = __ldg(c.rowptr + n0);
    const int32_t* __restrict__ g_rp = c.rowptr + n0;
    const int32_t* __restrict__ g_col = c.col + z0;
    const int32_t* __restrict__ g_ea = LABELED ? c.eattr + z0 : nullptr;
    W* __restrict__ st = state + n0;
    auto RP = [&](int i) -> int { return __ldg(g_rp + i) - z0; };
    auto COL = [&](int e) -> int { return __ldcs(g_col + e); };
    auto EA = [&](int e) -> int { return __ldcs(g_ea + e); };

    for (int i = lane; i < N; i += kWarp) {
      const unsigned r0 = (unsigned)(__ldg(g_rp + i) - z0), deg = (unsigned)(__ldg(g_rp + i + 1) - z0) - r0;
      if constexpr (P64) __stcg(st + i, (unsigned long long)r0 << 37 | (unsigned long long)deg << 25 | deg);
      else __stcg(st + i, deg);
    }
    // row extents of a node: from its state word (64-bit layout) or from rowptr
    auto extents = [&](int node, int& lo_, int& hi_, int* vis_) {
      if constexpr (P64) {
        const W w = ld_cg(st + node);
        lo_ = S::lo(w); hi_ = S::hi(w);
        if (vis_) *vis_ = S::vis(w);
      } else {
        lo_ = RP(node); hi_ = RP(node + 1);
        if (vis_) *vis_ = S::vis(ld_cg(st + node));
      }
    };
    for (int i = lane; i < ((N + 31) >> 5); i += kWarp) { openb[i] = 0u; visb[i] = 0u; }
    __syncwarp();

    const uint64_t gid = (uint64_t)(ch.graph_id_base + g);
    unsigned b0 = 0, b1 = 0, b2 = 0, b3 = 0, d = 0;
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
        if ((d & 3u) == 0u) philox10((unsigned)gid, (unsigned)(gid >> 32), d >> 2, 0u, k0, k1, b0, b1, b2, b3);
        const unsigned w01 = (d & 1u) ? b1 : b0, w23 = (d & 1u) ? b3 : b2;
        r = (d & 2u) ? w23 : w01;
      }
      ++d;
      return (int)__umulhi(r, (unsigned)cnt);
    };
    auto put = [&](int pos, int t) {
      if (pos < Lcap)
        __stcs(row + pos, (int64_t)(remap ? remap_zinc_tok(t, cfg.idx_offset, cfg.node_idx_offset, cfg.edge_idx_offset) : t));
    };

    int len = 1;
    if (lane == 0) put(0, SOS);
    if (N > 0) {
      int nvis = 0, prev = -1, nopen = 0;
      int cur = pick(N);
      int vc = -1;                                   // visit index of cur, -1 = first visit
      int lo, hi;
      extents(cur, lo, hi, nullptr);
      while (len < Lcap) {
        int cnt = 0;
        if (hi - lo <= kWarp) {
          // ---- the whole row sits in the lanes (lane <-> neighbour)
          const int e = lo + lane;
          const bool valid = e < hi;
          const int u = valid ? COL(e) : 0;
          int rlo = 0, rhi = 0, eav = 0;
          W old = 0;
          if (valid) {
            old = vc < 0 ? S::dec(st + u) : ld_cg(st + u);              // cur stops being an unvisited neighbour of u
            if constexpr (!P64) { rlo = RP(u); rhi = RP(u + 1); }
            if (LABELED) eav = EA(e);
          }
          if constexpr (P64) { rlo = S::lo(old); rhi = S::hi(old); }
          const int vu = S::vis(old);
          if (vc < 0) {
            const int nv = nvis++;
            if (lane == 0) {
              S::set_vis(st + cur, nv);
              visb[cur >> 5] |= 1u << (cur & 31);
              put(len, cfg.idx_offset + nv);
              if (LABELED) put(len + 1, cfg.node_idx_offset + __ldg(c.x + n0 + cur));
            }
            len += per;
            const bool z = valid && vu >= 0 && S::last(old);            // u just lost its last unvisited neighbour
            if (z) atomicAnd(&openb[u >> 5], ~(1u << (u & 31)));
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
            if (lane == 0) put(len, cfg.idx_offset + vc);               // head of a new segment at a visited node
            len += 1;
          }
          const unsigned fm = __ballot_sync(kFull, valid && vu < 0);    // unvisited neighbours, CSR order
          cnt = __popc(fm);
          if (cnt > 0) {
            if (vc < 0) {                                               // cur joins the restart candidates
              if (lane == 0) atomicOr(&openb[cur >> 5], 1u << (cur & 31));   // (other lanes update this map with atomics in the same step)
              ++nopen;
            }
            if ((fm >> lane) & 1u) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(g_col + rlo));
              if (LABELED) asm volatile("prefetch.global.L2 [%0];" ::"l"(g_ea + rlo));
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
            vc = -1;
            __syncwarp();
            continue;
          }
        } else {
          // ---- rows longer than a warp: chunked passes over the row
          if (vc < 0) {
            const int nv = nvis++;
            const int nwb = (nv + 31) >> 5;
            for (int i = lane; i < nwb; i += kWarp) bits[i] = 0u;
            if (lane == 0) {
              S::set_vis(st + cur, nv);
              visb[cur >> 5] |= 1u << (cur & 31);
              put(len, cfg.idx_offset + nv);
              if (LABELED) put(len + 1, cfg.node_idx_offset + __ldg(c.x + n0 + cur));
            }
            len += per;
            __syncwarp();
            int closed = 0;
            for (int e0 = lo; e0 < hi; e0 += kWarp) {
              const int e = e0 + lane;
              bool z = false, unv = false;
              if (e < hi) {
                const int u = COL(e);
                const W old = S::dec(st + u);
                const int vu = S::vis(old);
                unv = vu < 0;
                if (vu >= 0) {
                  z = S::last(old);
                  if (z) atomicAnd(&openb[u >> 5], ~(1u << (u & 31)));
                  if (u != prev) atomicOr(&bits[vu >> 5], 1u << (vu & 31));
                }
              }
              closed += __popc(__ballot_sync(kFull, z));
              cnt += __popc(__ballot_sync(kFull, unv));
            }
            nopen -= closed;
            __syncwarp();
            int M = 0;                                                  // exclusive popcount prefix per bitmap word
            for (int w0 = 0; w0 < nwb; w0 += kWarp) {
              const unsigned word = (w0 + lane < nwb) ? bits[w0 + lane] : 0u;
              const int cw = __popc(word);
              int incl = cw;
#pragma unroll
              for (int dd = 1; dd < 32; dd <<= 1) {
                const int o = __shfl_up_sync(kFull, incl, dd);
                if (lane >= dd) incl += o;
              }
              if (w0 + lane < nwb) pref[w0 + lane] = (unsigned short)(M + incl - cw);
              M += __shfl_sync(kFull, incl, 31);
            }
            __syncwarp();
            if (M > 0) {
              for (int e0 = lo; e0 < hi; e0 += kWarp) {
                const int e = e0 + lane;
                if (e < hi) {
                  const int u = COL(e);
                  const int vu = S::vis(ld_cg(st + u));
                  if (vu >= 0 && u != prev) {
                    const int rank = (int)pref[vu >> 5] + __popc(bits[vu >> 5] & ((1u << (vu & 31)) - 1u));
                    const int pos = len + 1 + (bracket_desc ? M - 1 - rank : rank) * per;
                    if (LABELED) {
                      put(pos, cfg.edge_idx_offset + EA(e) + cfg.edge_label_shift);
                      put(pos + 1, cfg.idx_offset + vu);
                    } else {
                      put(pos, cfg.idx_offset + vu);
                    }
                  }
                }
              }
              if (lane == 0) { put(len, LADJ); put(len + 1 + M * per, RADJ); }
              len += 2 + M * per;
            }
            if (cnt > 0) {
              if (lane == 0) atomicOr(&openb[cur >> 5], 1u << (cur & 31));   // (other lanes update this map with atomics in the same step)
              ++nopen;
            }
          } else {
            if (lane == 0) put(len, cfg.idx_offset + vc);
            len += 1;
            for (int e0 = lo; e0 < hi; e0 += kWarp) {
              const int e = e0 + lane;
              const bool unv = e < hi && S::unvisited(ld_cg(st + COL(e)));
              cnt += __popc(__ballot_sync(kFull, unv));
            }
          }
          if (cnt > 0) {
            // trail step to the k-th unvisited neighbour (CSR order)
            const int k = pick(cnt);
            int seen = 0, nxt = -1, ne = -1;
            for (int e0 = lo; e0 < hi && nxt < 0; e0 += kWarp) {
              const int e = e0 + lane;
              const int u = e < hi ? COL(e) : 0;
              const bool f = e < hi && S::unvisited(ld_cg(st + u));
              const unsigned m = __ballot_sync(kFull, f);
              const int cch = __popc(m);
              if (k < seen + cch) {
                const int sl = select32(m, k - seen);
                nxt = __shfl_sync(kFull, u, sl);
                ne = e0 + sl;
              }
              seen += cch;
            }
            if (LABELED) {
              if (lane == 0) put(len, cfg.edge_idx_offset + EA(ne) + cfg.edge_label_shift);
              len += 1;
            }
            prev = cur;
            cur = nxt;
            extents(cur, lo, hi, nullptr);
            vc = -1;
            __syncwarp();
            continue;
          }
        }
        __syncwarp();
        if (nvis == N) break;
        // ---- restart: visited nodes with unvisited neighbours, else unvisited nodes
        const bool reopen = nopen > 0 && !restart_unvisited;
        cur = reopen ? locate_bit(openb, false, pick(nopen), N, lane) : locate_bit(visb, true, pick(N - nvis), N, lane);
        extents(cur, lo, hi, &vc);
        if (!reopen) vc = -1;
        if (lane == 0) put(len, RESET);
        len += 1;
        prev = -1;
      }
    }
    if (!cfg.no_eos) {
      if (lane == 0) put(len, EOS);
      len += 1;
    }
    const int L = len < Lcap ? len : Lcap;
    int Ltot = L;
    if (cfg.append_query) {
      if (lane < 3)
        row[L + lane] = cfg.idx_offset + (lane == 0 ? N : (lane == 1 ? __ldg(c.query_u + g) : __ldg(c.query_v + g)));
      Ltot += 3;
    }
    if (lane == 0) len_out[g] = Ltot;
    __syncwarp();
  }
}

template <bool LABELED, bool P64>
int big_launch(const GtokCsrBatch* c, const GtokSentCfg* cfg, const GtokChoice* ch, int32_t* len_out, int64_t* out,
               int64_t stride, void* ws, cudaStream_t st, bool* handled) {
  const BigLayout lay = big_layout(c->max_nodes);
  auto k = sent_big_kernel<LABELED, P64>;
  auto* state = static_cast<typename St<P64>::T*>(ws);
  int dev = 0, max_optin = 0;
  GTOK_CUDA_TRY(cudaGetDevice(&dev));
  GTOK_CUDA_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t smem = (size_t)kBigWarps * lay.total;
  if (smem > (size_t)max_optin) return GTOK_OK;                      // not handled: the caller falls back
  GTOK_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t work = (c->num_graphs + kBigWarps - 1) / kBigWarps;
  const int grid = persistent_grid(k, kBigWarps * 32, smem, work);
  k<<<grid, kBigWarps * 32, smem, st>>>(*c, *cfg, *ch, lay, state, len_out, out, stride);
  GTOK_CUDA_TRY(cudaGetLastError());
  *handled = true;
  return GTOK_OK;
}

}  // namespace

int sent_big_dispatch(const GtokCsrBatch* c, const GtokSentCfg* cfg, const GtokChoice* ch, int32_t* len_out,
                      int64_t* out, int64_t stride, void* ws, size_t ws_bytes, int64_t total_nodes, cudaStream_t st,
                      bool* handled) {
  *handled = false;
  if (ws == nullptr || c->max_nodes <= 256) return GTOK_OK;
  GTOK_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 7) == 0, GTOK_EINVAL, "sent: workspace must be 8-byte aligned");
  GTOK_REQUIRE(total_nodes >= 0 && ws_bytes >= (size_t)total_nodes * 8, GTOK_EWORKSPACE,
               "sent: workspace %zu < %lld bytes (8 per node)", ws_bytes, (long long)total_nodes * 8);
  // 64-bit state words need: visit index + 1 <= 4096 (13 bits), degree <= 4095 (12 bits), row start < 2^27
  const bool p64 = c->max_nodes <= 4096 && c->max_nnz < (1 << 27);
  if (p64)
    return cfg->labeled ? big_launch<true, true>(c, cfg, ch, len_out, out, stride, ws, st, handled)
                        : big_launch<false, true>(c, cfg, ch, len_out, out, stride, ws, st, handled);
  return cfg->labeled ? big_launch<true, false>(c, cfg, ch, len_out, out, stride, ws, st, handled)
                      : big_launch<false, false>(c, cfg, ch, len_out, out, stride, ws, st, handled);
}

}  // namespace gtok
