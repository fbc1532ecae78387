// gtok_sent_walkl.cu - K3 for graphs of 65..256 nodes: the walk / emission kernel pair of gtok_sent_walk.cu for graphs whose
// masks no longer fit one 64-bit word.
//
// Same function (Graph2TrailTokenizer.__call__ + remap_zinc_tokens + query append, trainer/train_agtt.py:171-273; spec
// oracle/sent_oracle.py), same outputs bit for bit.  gtok_sent_mid.cu walks AND emits these sizes thread-per-graph on 2 or 4
// mask words per node; ordering every bracket by visit index on 4-word masks inside the sequential walk made a 256-node graph
// cost 5.6x a 64-node one per node (large-graph sweep, profiles/r1r_large_graph_sweep.jsonl: 21 G nodes/s at 64 nodes, 10.9 at
// 128, 3.8 at 256).  Here:
//   sent_walkm_kernel  thread per graph, lockstep warp, one arrival and one pick per iteration, on the ADJACENCY BIT ROWS
//                      (GtokCsrBatch.adj64): one aligned 16- or 32-byte load per step, candidates / bracket members /
//                      restart sets as mask arithmetic on 4 or 8 32-bit words held in registers with the visited and open
//                      sets; shared memory holds one open-neighbour counter byte per node.  A first visit logs ONE 8-byte item
//                      (node, how it was reached, first token position).  Brackets are neither ordered nor recorded.
//   sent_walkl_kernel  the same walk on the CSR neighbour lists (for a CsrBatch without bit rows; GTOK_SENT_LISTS=1): the
//                      row is read 8 entries at a time.  2x slower than the bit-row walk: rowptr -> col is a dependent chain
//                      and the scattered 4-byte reads pull 7x their bytes from DRAM (profiles/r2y).
//   sent_emitl_kernel  warp per row, lane per item, 32 items per round: the node -> visit-index map is rebuilt from the items,
//                      every lane rescans ITS node's CSR row, keeps the neighbours visited before it and insertion-sorts them
//                      by visit index straight into the row buffer (shared memory, 16-bit tokens); then the warp streams the
//                      row with 64-token stores - or writes it into its padded batch (collate_fn) like the small-graph pair.
// Per 32M nodes (mean degree 8): 128 nodes 2.95 -> 2.1 ms, 256 nodes 8.5 -> 3.6 ms.
#include "gtok_sent_common.cuh"

namespace gtok {

namespace {

constexpr unsigned kFlagReset = 1u << 8, kFlagHead = 1u << 9, kFlagEnd = 1u << 31;
constexpr int kEmitLThreads = 256;
constexpr int kEmitLWarps = kEmitLThreads / 32;

template <bool LABELED>
__global__ void __launch_bounds__(kSmallThreads, 8)
sent_walkl_kernel(const GtokCsrBatch c, const GtokSentCfg cfg, const GtokChoice ch, const int nmax, const int words,
                  uint2* __restrict__ log_lo, int32_t* __restrict__ len_out) {
  extern __shared__ __align__(16) unsigned char wl_smem[];
  const int lane = lane_id();
  const int w_ = threadIdx.x >> 5;
  const int warp_bytes = words * 256 + nmax * 32;
  unsigned* vb = reinterpret_cast<unsigned*>(wl_smem + (size_t)w_ * warp_bytes) + lane;   // u32 [word][lane] visited bits
  unsigned* ob = vb + words * 32;                                                          // u32 [word][lane] open bits
  unsigned char* rem = reinterpret_cast<unsigned char*>(vb - lane + words * 64) + lane;    // u8  [v][lane] unvisited neighbours left
  const int64_t G = c.num_graphs;
  const int64_t wid0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int Lcap = cfg.max_length;
  const unsigned k0 = (unsigned)(ch.seed & 0xffffffffu), k1 = (unsigned)(ch.seed >> 32);
  const bool restart_unvisited = cfg.restart_policy == GTOK_SENT_RESTART_UNVISITED;
  const int32_t* __restrict__ colp = c.col;

  for (int64_t g0 = wid0 * 32; g0 < G; g0 += nw * 32) {
    const int Tc = (int)((G - g0 < 32) ? (G - g0) : 32);
    const bool valid = lane < Tc;
    const int64_t g = valid ? (c.order ? (int64_t)__ldg(c.order + g0 + lane) : g0 + lane) : g0;
    const int64_t my_n0 = __ldg(c.node_off + g);
    const int N = valid ? (int)(__ldg(c.node_off + g + 1) - my_n0) : 0;
    const int32_t* __restrict__ rp = c.rowptr + my_n0;
    for (int wd = 0; wd < words; ++wd) { vb[wd * 32] = 0u; ob[wd * 32] = 0u; }

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
    int nvis = 0, cur = 0, prev = -1;
    unsigned pre = 0;                                 // flags (+ head node << 16) of the next item
    int pos = 1;                                      // tokens so far (SOS)
    bool done = N == 0;
    int e0 = 0, e1 = 0;                               // CSR row of the node we are about to arrive at
    if (!done) {
      cur = pick(N);
      e0 = __ldg(rp + cur);
      e1 = __ldg(rp + cur + 1);
    }
    while (!__all_sync(kFull, done)) {
      if (!done) {
        const int cw = (cur >> 5) * 32;
        const unsigned bit = 1u << (cur & 31);
        const unsigned vw = vb[cw];
        const bool first = !(vw & bit);
        if (first) vb[cw] = vw | bit;
        int cntU = 0, cntM = 0;                       // unvisited neighbours / visited neighbours except the trail predecessor
        // The row is read 8 entries at a time: eight INDEPENDENT loads, one L2 latency per chunk (entry by entry the scan was
        // a chain of dependent loads: 10.7 ms per 32M nodes at 256 nodes).  The first chunk stays in registers for the pick.
        int w0[8];
        unsigned um0 = 0u;                            // which entries of the first chunk are unvisited
#pragma unroll
        for (int j = 0; j < 8; ++j) w0[j] = (e0 + j < e1) ? __ldg(colp + e0 + j) : -1;
        auto visit_nbr = [&](int w, bool& unv) {
          const int ww = (w >> 5) * 32;
          const unsigned wb = 1u << (w & 31);
          unv = !(vb[ww] & wb);
          if (unv) {
            ++cntU;
          } else if (first && w != prev) {            // a bracket member: cur is no longer an unvisited neighbour of w
            ++cntM;
            const int r = (int)rem[w * 32] - 1;
            rem[w * 32] = (unsigned char)r;
            if (r == 0) ob[ww] &= ~wb;
          }
        };
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (w0[j] >= 0) {
            bool unv;
            visit_nbr(w0[j], unv);
            if (unv) um0 |= 1u << j;
          }
        }
        for (int eb = e0 + 8; eb < e1; eb += 8) {
          int wv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) wv[j] = (eb + j < e1) ? __ldg(colp + eb + j) : -1;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (wv[j] >= 0) {
              bool unv;
              visit_nbr(wv[j], unv);
            }
          }
        }
        if (first) {
          const int k = nvis++;
          rem[cur * 32] = (unsigned char)cntU;
          if (cntU) ob[cw] |= bit;
          int cnt = ((pre & kFlagReset) ? 1 : 0) + ((pre & kFlagHead) ? 1 : 0) + 1 + (cntM ? 2 + (LABELED ? 2 * cntM : cntM) : 0);
          if (LABELED) cnt += 1 + (prev >= 0 ? 1 : 0);
          log_lo[my_n0 + k] = make_uint2((unsigned)pos, (unsigned)cur | pre);
          pos += cnt;
          pre = 0;
        } else {
          pre = kFlagReset | kFlagHead | ((unsigned)cur << 16);     // head of a new segment at a visited node
        }
        if ((cntU == 0 && nvis == N) || pos >= Lcap) {
          done = true;
        } else {
          int nxt = 0;
          if (cntU > 0) {                             // trail step to the kth unvisited neighbour (ascending node id)
            int kth = pick(cntU);
            const int c0 = __popc(um0);
            if (kth < c0) {                           // among the first eight entries: still in registers
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if ((um0 >> j) & 1u) {
                  if (kth == 0) nxt = w0[j];
                  --kth;
                }
              }
            } else {
              kth -= c0;
              bool found = false;
              for (int eb = e0 + 8; eb < e1 && !found; eb += 8) {
                int wv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) wv[j] = (eb + j < e1) ? __ldg(colp + eb + j) : -1;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  if (!found && wv[j] >= 0 && !(vb[(wv[j] >> 5) * 32] & (1u << (wv[j] & 31)))) {
                    if (kth == 0) { nxt = wv[j]; found = true; }
                    --kth;
                  }
                }
              }
            }
            rem[cur * 32] = (unsigned char)(cntU - 1);
            if (cntU == 1) ob[cw] &= ~bit;
            prev = cur;
          } else {                                    // stuck: restart at an open node, else at any unvisited node
            int nsel = 0;
            if (!restart_unvisited)
              for (int wd = 0; wd < words; ++wd) nsel += __popc(ob[wd * 32]);
            const bool use_open = nsel > 0;
            if (!use_open) nsel = N - nvis;
            int kth = pick(nsel);
            for (int wd = 0; wd < words; ++wd) {
              const int lim = N - wd * 32;
              const unsigned nm = lim >= 32 ? 0xffffffffu : (lim <= 0 ? 0u : ((1u << lim) - 1u));
              const unsigned m = use_open ? ob[wd * 32] : (~vb[wd * 32] & nm);
              const int cm = __popc(m);
              if (kth < cm) { nxt = wd * 32 + select32(m, kth); break; }
              kth -= cm;
            }
            pre = kFlagReset;
            prev = -1;
          }
          cur = nxt;
          e0 = __ldg(rp + cur);
          e1 = __ldg(rp + cur + 1);
        }
      }
    }
    if (valid) {
      if (nvis < N) log_lo[my_n0 + nvis] = make_uint2(0u, kFlagEnd);          // early stop: end of the item list
      const int Lfull = pos + (cfg.no_eos ? 0 : 1);
      len_out[g] = (Lfull < Lcap ? Lfull : Lcap) + (cfg.append_query ? 3 : 0);
    }
    __syncwarp();
  }
}

// The same walk on ADJACENCY BIT ROWS (GtokCsrBatch.adj64, 2 or 4 64-bit words per node): one aligned 16- or 32-byte load
// per step instead of the dependent rowptr -> col chain of the list walk (which also pulled 7.8 GB of DRAM sectors for 1.1
// GB of lists at 256 nodes, profiles/r2y), candidates and bracket members as mask arithmetic on NW 32-bit words held in
// registers together with the visited and open sets; only the per-node open counters live in shared memory.  Unlike
// gtok_sent_mid.cu nothing is emitted and no bracket is ordered here - the emission kernel does that with a lane per item.
template <bool LABELED, int NW>
__global__ void __launch_bounds__(kSmallThreads, NW == 8 ? 4 : 6)   // measured: 8 words 2.21 -> 2.09 ms with 4 CTAs (114 registers, no spills)
sent_walkm_kernel(const GtokCsrBatch c, const GtokSentCfg cfg, const GtokChoice ch, const int nmax,
                  uint2* __restrict__ log_lo, int32_t* __restrict__ len_out) {
  extern __shared__ __align__(16) unsigned char wm_smem[];
  const int lane = lane_id();
  const int w_ = threadIdx.x >> 5;
  unsigned char* rem = wm_smem + (size_t)w_ * nmax * 32 + lane;       // u8 [v][lane] unvisited neighbours left (visited nodes)
  const int64_t G = c.num_graphs;
  const int64_t wid0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int Lcap = cfg.max_length;
  const unsigned k0 = (unsigned)(ch.seed & 0xffffffffu), k1 = (unsigned)(ch.seed >> 32);
  const bool restart_unvisited = cfg.restart_policy == GTOK_SENT_RESTART_UNVISITED;

  for (int64_t g0 = wid0 * 32; g0 < G; g0 += nw * 32) {
    const int Tc = (int)((G - g0 < 32) ? (G - g0) : 32);
    const bool valid = lane < Tc;
    const int64_t g = valid ? (c.order ? (int64_t)__ldg(c.order + g0 + lane) : g0 + lane) : g0;
    const int64_t my_n0 = __ldg(c.node_off + g);
    const int N = valid ? (int)(__ldg(c.node_off + g + 1) - my_n0) : 0;
    const uint4* __restrict__ arow = reinterpret_cast<const uint4*>(c.adj64) + my_n0 * (NW / 4);

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
    unsigned vis[NW], opn[NW], an[NW];                // visited set, open set, adjacency row of the node we arrive at next
#pragma unroll
    for (int i = 0; i < NW; ++i) { vis[i] = 0u; opn[i] = 0u; an[i] = 0u; }
    auto load_row = [&](int node) {
#pragma unroll
      for (int q = 0; q < NW / 4; ++q) {
        const uint4 t = __ldg(arow + (size_t)node * (NW / 4) + q);
        an[4 * q] = t.x; an[4 * q + 1] = t.y; an[4 * q + 2] = t.z; an[4 * q + 3] = t.w;
      }
    };
    // k-th set bit (ascending) of the NW-word set `m`
    auto select_words = [&](const unsigned (&m)[NW], int kth) -> int {
      unsigned msel = 0u;
      int wsel = 0;
      bool found = false;
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        const int cm = __popc(m[i]);
        if (!found) {
          if (kth < cm) { msel = m[i]; wsel = i; found = true; }
          else kth -= cm;
        }
      }
      return wsel * 32 + select32(msel, kth);
    };
    int nvis = 0, cur = 0, prev = -1;
    unsigned pre = 0;
    int pos = 1;
    bool done = N == 0;
    if (!done) {
      cur = pick(N);
      load_row(cur);
    }
    while (!__all_sync(kFull, done)) {
      if (!done) {
        const int cw = cur >> 5;
        const unsigned bit = 1u << (cur & 31);
        unsigned a[NW], U[NW];
        bool first = false;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          a[i] = an[i];
          if (i == cw) { first = !(vis[i] & bit); vis[i] |= bit; }
        }
        int cntU = 0, cntM = 0;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          U[i] = a[i] & ~vis[i];
          cntU += __popc(U[i]);
        }
        if (first) {
          const int pw = prev >> 5;                   // (prev = -1: pw = -1, matches no word)
          const unsigned pbit = 1u << (prev & 31);
#pragma unroll
          for (int i = 0; i < NW; ++i) {
            unsigned m = a[i] & vis[i];               // visited neighbours ...
            if (i == pw) m &= ~pbit;                  // ... except the trail predecessor = the bracket members
            cntM += __popc(m);
            while (m) {                               // cur is no longer an unvisited neighbour of the member
              const int b = __ffs((int)m) - 1;
              m &= m - 1;
              const int u = i * 32 + b;
              const int r = (int)rem[u * 32] - 1;
              rem[u * 32] = (unsigned char)r;
              if (r == 0) opn[i] &= ~(1u << b);
            }
          }
          const int k = nvis++;
          rem[cur * 32] = (unsigned char)cntU;
          if (cntU) {
#pragma unroll
            for (int i = 0; i < NW; ++i)
              if (i == cw) opn[i] |= bit;
          }
          int cnt = ((pre & kFlagReset) ? 1 : 0) + ((pre & kFlagHead) ? 1 : 0) + 1 + (cntM ? 2 + (LABELED ? 2 * cntM : cntM) : 0);
          if (LABELED) cnt += 1 + (prev >= 0 ? 1 : 0);
          log_lo[my_n0 + k] = make_uint2((unsigned)pos, (unsigned)cur | pre);
          pos += cnt;
          pre = 0;
        } else {
          pre = kFlagReset | kFlagHead | ((unsigned)cur << 16);
        }
        if ((cntU == 0 && nvis == N) || pos >= Lcap) {
          done = true;
        } else {
          int nxt;
          if (cntU > 0) {                             // trail step to the kth unvisited neighbour (ascending node id)
            nxt = select_words(U, pick(cntU));
            rem[cur * 32] = (unsigned char)(cntU - 1);
            if (cntU == 1) {
#pragma unroll
              for (int i = 0; i < NW; ++i)
                if (i == cw) opn[i] &= ~bit;
            }
            prev = cur;
          } else {                                    // stuck: restart at an open node, else at any unvisited node
            int nsel = 0;
            if (!restart_unvisited) {
#pragma unroll
              for (int i = 0; i < NW; ++i) nsel += __popc(opn[i]);
            }
            const bool use_open = nsel > 0;
            if (!use_open) nsel = N - nvis;
            unsigned R[NW];
#pragma unroll
            for (int i = 0; i < NW; ++i) {
              const int lim = N - i * 32;
              const unsigned nm = lim >= 32 ? 0xffffffffu : (lim <= 0 ? 0u : ((1u << lim) - 1u));
              R[i] = use_open ? opn[i] : (~vis[i] & nm);
            }
            nxt = select_words(R, pick(nsel));
            pre = kFlagReset;
            prev = -1;
          }
          cur = nxt;
          load_row(cur);
        }
      }
    }
    if (valid) {
      if (nvis < N) log_lo[my_n0 + nvis] = make_uint2(0u, kFlagEnd);
      const int Lfull = pos + (cfg.no_eos ? 0 : 1);
      len_out[g] = (Lfull < Lcap ? Lfull : Lcap) + (cfg.append_query ? 3 : 0);
    }
    __syncwarp();
  }
}

template <bool LABELED, bool COLLATE, bool REMAP>
__global__ void __launch_bounds__(kEmitLThreads)
sent_emitl_kernel(const GtokCsrBatch c, const GtokSentCfg cfg, const uint2* __restrict__ log_lo,
                  const int32_t* __restrict__ len, int64_t* __restrict__ out, int64_t stride,
                  const int64_t* __restrict__ batch_idx, int64_t num_rows, int64_t B, const int32_t* __restrict__ lmax,
                  const int64_t* __restrict__ xoff, int64_t pad_id, uint8_t* __restrict__ attn, int row_cap) {
  constexpr int kRemapN = 1024;
  extern __shared__ __align__(16) unsigned char el_smem[];
  __shared__ unsigned short s_remap[kRemapN];
  __shared__ unsigned char s_ord[kEmitLWarps][256];
  __shared__ unsigned char s_vis[kEmitLWarps][256];
  __shared__ ulonglong2 s_m16[17];
  if (COLLATE) {
    fill_mask_table(s_m16, threadIdx.x);
    __syncthreads();
  }
  const int lane = lane_id();
  const int w_ = threadIdx.x >> 5;
  unsigned short* rowb = reinterpret_cast<unsigned short*>(el_smem) + (size_t)w_ * row_cap;
  unsigned char* ordb = s_ord[w_];
  unsigned char* visb = s_vis[w_];
  const bool desc = cfg.bracket_order == GTOK_SENT_BRACKET_VISIT_DESC;
  const int idx_off = cfg.idx_offset, node_off = cfg.node_idx_offset, edge_off = cfg.edge_idx_offset;
  const int elab_off = edge_off + cfg.edge_label_shift;
  const int q3 = cfg.append_query ? 3 : 0;
  const int32_t* __restrict__ colp = c.col;
  if (REMAP) {
    for (int t = threadIdx.x; t < kRemapN; t += blockDim.x) s_remap[t] = (unsigned short)remap_zinc_tok(t, idx_off, node_off, edge_off);
    __syncthreads();
  }
  const int64_t wid0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = wid0; r < num_rows; r += nw) {
    const int64_t g = (COLLATE && batch_idx) ? __ldg(batch_idx + r) : r;
    const int64_t n0 = __ldg(c.node_off + g);
    const int N = (int)(__ldg(c.node_off + g + 1) - n0);
    const int L = __ldg(len + g);                    // row length, query included
    const int L0 = L - q3;
    // ---- pass A: node of every item, node -> visit index (0xFF: not visited before the walk stopped) -------------------
    for (int v = lane; v < N; v += 32) visb[v] = 0xFF;
    __syncwarp();
    int nvis = N;
    for (int kb = 0; kb < N; kb += 32) {
      const int k = kb + lane;
      uint2 it = make_uint2(0u, kFlagEnd);
      if (k < N) it = __ldg(log_lo + n0 + k);
      const unsigned endm = __ballot_sync(kFull, (it.y & kFlagEnd) != 0u);
      const int nk = endm ? __ffs((int)endm) - 1 : 32;
      if (lane < nk) {
        ordb[k] = (unsigned char)(it.y & 0xff);
        visb[it.y & 0xff] = (unsigned char)k;
      }
      if (endm) { nvis = kb + nk; break; }
    }
    __syncwarp();
    if (lane == 0) rowb[0] = (unsigned short)SOS;
    int end_pos = (nvis == 0 && lane == 0) ? 1 : -1;
    // ---- pass B: lane <-> item, tokens into the row buffer --------------------------------------------------------------
    for (int kb = 0; kb < nvis; kb += 32) {
      const int k = kb + lane;
      if (k < nvis) {
        const uint2 it = __ldg(log_lo + n0 + k);
        const int u = it.y & 0xff;
        int p = (int)it.x;
        int pv = -1;                                 // trail predecessor (node id)
        if (it.y & kFlagHead) pv = (it.y >> 16) & 0xff;
        else if (!(it.y & kFlagReset) && k > 0) pv = ordb[k - 1];
        if (it.y & kFlagReset) rowb[p++] = (unsigned short)RESET;
        if (it.y & kFlagHead) rowb[p++] = (unsigned short)(idx_off + visb[pv]);
        int pe = -1;                                 // slot of the trail edge's label (filled when the scan meets pv)
        if (LABELED && pv >= 0) pe = p++;
        rowb[p++] = (unsigned short)(idx_off + k);
        if (LABELED) rowb[p++] = (unsigned short)(node_off + __ldg(c.x + n0 + u));
        unsigned short* mb = rowb + p + 1;           // members go after LADJ, kept sorted by visit index as they arrive
        int m = 0;
        auto nbr = [&](int w, int e) {               // neighbour w of u, CSR entry e
          if (w == pv) {
            if (LABELED) rowb[pe] = (unsigned short)(elab_off + __ldg(c.eattr + e));
            return;
          }
          const int kv = visb[w];
          if (kv >= k) return;                       // visited later (or never)
          const unsigned short key = (unsigned short)(idx_off + kv);     // appended in neighbour order, sorted below
          if (LABELED) {
            mb[2 * m] = (unsigned short)(elab_off + __ldg(c.eattr + e));
            mb[2 * m + 1] = key;
          } else {
            mb[m] = key;
          }
          ++m;
        };
        // (the members are found in the neighbour LIST of u: enumerating the bit row instead - one aligned load, no rowptr - was
        //  slower, 1.11 -> 1.37 ms per 32M nodes at 128 nodes: eight data-dependent loops per lane instead of one)
        const int e0 = __ldg(c.rowptr + n0 + u), e1 = __ldg(c.rowptr + n0 + u + 1);
        for (int eb = e0; eb < e1; eb += 8) {        // eight independent loads per chunk (see the walk kernel)
          int wv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) wv[j] = (eb + j < e1) ? __ldg(colp + eb + j) : -1;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (wv[j] >= 0) nbr(wv[j], eb + j);
        }
        // one insertion sort per item, by visit index (the sort sat inside the 8-way unrolled neighbour scan before: eight
        // copies of a data-dependent loop, 45 % of this kernel's instructions)
        for (int i = 1; i < m; ++i) {
          if (LABELED) {
            const unsigned short lab = mb[2 * i], key = mb[2 * i + 1];
            int jj = i;
            while (jj > 0 && (desc ? mb[2 * jj - 1] < key : mb[2 * jj - 1] > key)) {
              mb[2 * jj] = mb[2 * jj - 2];
              mb[2 * jj + 1] = mb[2 * jj - 1];
              --jj;
            }
            mb[2 * jj] = lab;
            mb[2 * jj + 1] = key;
          } else {
            const unsigned short key = mb[i];
            int jj = i;
            while (jj > 0 && (desc ? mb[jj - 1] < key : mb[jj - 1] > key)) { mb[jj] = mb[jj - 1]; --jj; }
            mb[jj] = key;
          }
        }
        if (m) {
          rowb[p] = (unsigned short)LADJ;
          p += 1 + (LABELED ? 2 * m : m);
          rowb[p++] = (unsigned short)RADJ;
        }
        if (k == nvis - 1) end_pos = p;
      }
    }
    if (end_pos >= 0) rowb[end_pos] = (unsigned short)EOS;       // (at or beyond the row length it is never read)
    __syncwarp();
    // ---- stream the row -----------------------------------------------------------------------------------------------------
    auto rm = [&](int t) -> int64_t {
      if (REMAP) t = t < kRemapN ? (int)s_remap[t] : remap_zinc_tok(t, idx_off, node_off, edge_off);
      return (int64_t)t;
    };
    auto tokv = [&](int q) -> int64_t { return rm((int)rowb[q]); };
    int64_t qt = 0;
    if (q3 && lane < 3) qt = idx_off + (lane == 0 ? N : (lane == 1 ? __ldg(c.query_u + g) : __ldg(c.query_v + g)));
    if (!COLLATE) {
      int64_t* __restrict__ row = out + g * stride;  // rows are 16-byte aligned (even stride, aligned base)
      const int npair = L0 >> 1;
      for (int j = lane; j < npair; j += 32) {
        const unsigned pr2 = reinterpret_cast<const unsigned*>(rowb)[j];
        st_stream_i64x2(row + 2 * j, rm((int)(pr2 & 0xffff)), rm((int)(pr2 >> 16)));
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

}  // namespace

// Row-buffer tokens for graphs of 65..256 nodes: the longest untruncated row (every node: reset + head + edge label + index +
// node label + LADJ + RADJ; every edge once as a bracket member) or max_length plus one item, whichever is smaller.
int sent_walkl_row_cap(const GtokCsrBatch* c, const GtokSentCfg* cfg) {
  const long long lab = cfg->labeled ? 1 : 0;
  const long long bound = 2 + (long long)c->max_nodes * (6 + lab) + ((long long)c->max_nnz / 2 + 1) * (1 + lab) + 8;
  const long long item = 8 + 2ll * c->max_nodes;
  long long cap = (long long)cfg->max_length + item;
  if (bound < cap) cap = bound;
  cap = (cap + 15) & ~15ll;
  if (cap * 2 * kEmitLWarps > 160 * 1024) return 0;
  return (int)cap;
}

template <int NW>
static int sent_walkm_launch(const GtokCsrBatch* c, const GtokSentCfg* cfg, const GtokChoice* ch, void* log, int32_t* len_out,
                             cudaStream_t st) {
  auto k = cfg->labeled ? sent_walkm_kernel<true, NW> : sent_walkm_kernel<false, NW>;
  const size_t smem = (size_t)kSmallWarps * c->max_nodes * 32;
  GTOK_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t warps = (c->num_graphs + 31) / 32;
  const int grid = persistent_grid(k, kSmallThreads, smem, (warps + kSmallWarps - 1) / kSmallWarps);
  k<<<grid, kSmallThreads, smem, st>>>(*c, *cfg, *ch, c->max_nodes, reinterpret_cast<uint2*>(log), len_out);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

int sent_walkl_launch(const GtokCsrBatch* c, const GtokSentCfg* cfg, const GtokChoice* ch, void* log, int32_t* len_out,
                      cudaStream_t st) {
  static const bool lists = [] { const char* e = getenv("GTOK_SENT_LISTS"); return e && atoi(e) != 0; }();
  if (!lists && c->adj64 != nullptr && (reinterpret_cast<uintptr_t>(c->adj64) & 15) == 0) {      // bit rows: one load per step
    if (c->adj_words == 2) return sent_walkm_launch<4>(c, cfg, ch, log, len_out, st);
    if (c->adj_words == 4) return sent_walkm_launch<8>(c, cfg, ch, log, len_out, st);
  }
  auto k = cfg->labeled ? sent_walkl_kernel<true> : sent_walkl_kernel<false>;
  const int words = (c->max_nodes + 31) / 32;
  const size_t smem = (size_t)kSmallWarps * (words * 256 + c->max_nodes * 32);
  GTOK_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t warps = (c->num_graphs + 31) / 32;
  const int grid = persistent_grid(k, kSmallThreads, smem, (warps + kSmallWarps - 1) / kSmallWarps);
  k<<<grid, kSmallThreads, smem, st>>>(*c, *cfg, *ch, c->max_nodes, words, reinterpret_cast<uint2*>(log), len_out);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

int sent_emitl_launch(const GtokCsrBatch* c, const GtokSentCfg* cfg, int row_cap, bool collate, const void* log,
                      const int32_t* len, int64_t* out, int64_t stride, const int64_t* batch_idx, int64_t num_rows, int64_t B,
                      const int32_t* lmax, const int64_t* xoff, int64_t pad_id, uint8_t* attn, cudaStream_t st) {
  void (*k)(GtokCsrBatch, GtokSentCfg, const uint2*, const int32_t*, int64_t*, int64_t, const int64_t*, int64_t, int64_t,
            const int32_t*, const int64_t*, int64_t, uint8_t*, int);
  const bool lab = cfg->labeled != 0, rmp = cfg->remap_zinc != 0;
  if (collate) {
    if (lab) k = rmp ? sent_emitl_kernel<true, true, true> : sent_emitl_kernel<true, true, false>;
    else k = rmp ? sent_emitl_kernel<false, true, true> : sent_emitl_kernel<false, true, false>;
  } else {
    if (lab) k = rmp ? sent_emitl_kernel<true, false, true> : sent_emitl_kernel<true, false, false>;
    else k = rmp ? sent_emitl_kernel<false, false, true> : sent_emitl_kernel<false, false, false>;
  }
  const size_t smem = (size_t)kEmitLWarps * row_cap * 2;
  GTOK_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = persistent_grid(k, kEmitLThreads, smem, (num_rows + kEmitLWarps - 1) / kEmitLWarps);
  k<<<grid, kEmitLThreads, smem, st>>>(*c, *cfg, reinterpret_cast<const uint2*>(log), len, out, stride, batch_idx, num_rows, B,
                                       lmax, xoff, pad_id, attn, row_cap);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

}  // namespace gtok
