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

__device__ __forceinline__ unsigned ld_cg_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

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

template <bool LABELED>
__global__ void __launch_bounds__(kBigWarps * 32)
sent_big_kernel(const GtokCsrBatch c, const GtokSentCfg cfg, const GtokChoice ch, const BigLayout lay,
                unsigned* __restrict__ state, int32_t* __restrict__ len_out, int64_t* __restrict__ out, int64_t stride) {
  extern __shared__ __align__(16) unsigned char b_smem[];
  const int lane = lane_id();
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
    unsigned* __restrict__ st = state + n0;
    auto RP = [&](int i) -> int { return __ldg(g_rp + i) - z0; };
    auto COL = [&](int e) -> int { return __ldcs(g_col + e); };
    auto EA = [&](int e) -> int { return __ldcs(g_ea + e); };

    for (int i = lane; i < N; i += kWarp) __stcg(st + i, (unsigned)(__ldg(g_rp + i + 1) - __ldg(g_rp + i)));
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
      int lo = RP(cur), hi = RP(cur + 1);
      while (len < Lcap) {
        int cnt = 0;
        if (hi - lo <= kWarp) {
          // ---- the whole row sits in the lanes (lane <-> neighbour)
          const int e = lo + lane;
          const bool valid = e < hi;
          const int u = valid ? COL(e) : 0;
          int rlo = 0, rhi = 0, eav = 0;
          unsigned old = 0;
          if (valid) {
            old = vc < 0 ? atomicSub(st + u, 1u) : ld_cg_u32(st + u);   // cur stops being an unvisited neighbour of u
            rlo = RP(u); rhi = RP(u + 1);
            if (LABELED) eav = EA(e);
          }
          const int vu = (int)(old >> 16) - 1;
          if (vc < 0) {
            const int nv = nvis++;
            if (lane == 0) {
              atomicAdd(st + cur, (unsigned)(nv + 1) << 16);
              visb[cur >> 5] |= 1u << (cur & 31);
              put(len, cfg.idx_offset + nv);
              if (LABELED) put(len + 1, cfg.node_idx_offset + __ldg(c.x + n0 + cur));
            }
            len += per;
            const bool z = valid && vu >= 0 && (old & 0xffffu) == 1u;   // u just lost its last unvisited neighbour
            if (z) atomicAnd(&openb[u >> 5], ~(1u << (u & 31)));
            nopen -= __popc(__ballot_sync(kFull, z));
            const bool member = valid && vu >= 0 && u != prev;
            const unsigned mm = __ballot_sync(kFull, member);
            if (mm) {
              const int M = __popc(mm);
              int rank = 0;
              for (unsigned t = mm; t;) {                               // four independent shuffles per trip
                int sl4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { sl4[j] = t ? __ffs(t) - 1 : lane; t &= t - 1; }
#pragma unroll
                for (int j = 0; j < 4; ++j) rank += (__shfl_sync(kFull, vu, sl4[j]) < vu) ? 1 : 0;
              }
              if (member) {
                const int pos = len + 1 + rank * per;
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
              if (lane == 0) openb[cur >> 5] |= 1u << (cur & 31);
              ++nopen;
            }
            if ((fm >> lane) & 1u) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(g_col + rlo));
              if (LABELED) asm volatile("prefetch.global.L2 [%0];" ::"l"(g_ea + rlo));
            }
            const int sl = select32(fm, pick(cnt));
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
              atomicAdd(st + cur, (unsigned)(nv + 1) << 16);
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
                const unsigned old = atomicSub(st + u, 1u);
                const int vu = (int)(old >> 16) - 1;
                unv = vu < 0;
                if (vu >= 0) {
                  z = (old & 0xffffu) == 1u;
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
                  const int vu = (int)(ld_cg_u32(st + u) >> 16) - 1;
                  if (vu >= 0 && u != prev) {
                    const int rank = (int)pref[vu >> 5] + __popc(bits[vu >> 5] & ((1u << (vu & 31)) - 1u));
                    const int pos = len + 1 + rank * per;
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
              if (lane == 0) openb[cur >> 5] |= 1u << (cur & 31);
              ++nopen;
            }
          } else {
            if (lane == 0) put(len, cfg.idx_offset + vc);
            len += 1;
            for (int e0 = lo; e0 < hi; e0 += kWarp) {
              const int e = e0 + lane;
              const bool unv = e < hi && (ld_cg_u32(st + COL(e)) >> 16) == 0u;
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
              const bool f = e < hi && (ld_cg_u32(st + u) >> 16) == 0u;
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
            lo = RP(cur); hi = RP(cur + 1);
            vc = -1;
            __syncwarp();
            continue;
          }
        }
        __syncwarp();
        if (nvis == N) break;
        // ---- restart: visited nodes with unvisited neighbours, else unvisited nodes
        int nxt;
        if (nopen > 0) {
          nxt = locate_bit(openb, false, pick(nopen), N, lane);
          vc = (int)(ld_cg_u32(st + nxt) >> 16) - 1;
        } else {
          nxt = locate_bit(visb, true, pick(N - nvis), N, lane);
          vc = -1;
        }
        if (lane == 0) put(len, RESET);
        len += 1;
        cur = nxt;
        prev = -1;
        lo = RP(cur); hi = RP(cur + 1);
      }
    }
    if (lane == 0) put(len, EOS);
    len += 1;
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

template <bool LABELED>
int big_launch(const GtokCsrBatch* c, const GtokSentCfg* cfg, const GtokChoice* ch, int32_t* len_out, int64_t* out,
               int64_t stride, unsigned* state, cudaStream_t st, bool* handled) {
  const BigLayout lay = big_layout(c->max_nodes);
  auto k = sent_big_kernel<LABELED>;
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
  GTOK_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 3) == 0, GTOK_EINVAL, "sent: workspace must be 4-byte aligned");
  GTOK_REQUIRE(total_nodes >= 0 && ws_bytes >= (size_t)total_nodes * 4, GTOK_EWORKSPACE,
               "sent: workspace %zu < %lld bytes (4 per node)", ws_bytes, (long long)total_nodes * 4);
  return cfg->labeled ? big_launch<true>(c, cfg, ch, len_out, out, stride, static_cast<unsigned*>(ws), st, handled)
                      : big_launch<false>(c, cfg, ch, len_out, out, stride, static_cast<unsigned*>(ws), st, handled);
}

}  // namespace gtok
