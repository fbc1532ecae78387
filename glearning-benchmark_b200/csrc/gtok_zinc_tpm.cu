// gtok_zinc_tpm.cu - K2 v6: ONE THREAD PER MOLECULE inside a CTA tile, single pass over HBM.
//
// Same reference functions as gtok_zinc.cu (tokenize_molecule + __getitem__ truncation +
// TokenDataset, zinc_dataset_indexbase.py:143-227, data_loader.py:465-486), and - in the
// PADDED variant - collate as well (data_loader.py:488-497): molecules -> padded X / attn
// batches without the ragged stream ever touching HBM.
//
// Why thread per molecule: the warp-per-molecule kernels (gtok_zinc.cu and the fused single-pass kernel this
// file replaced, profiles/r1a..r1g) are ISSUE bound at ~500 warp instructions per molecule: the only sequential
// part of the job, the first-seen dedupe of unordered pairs, was paid with match_any /
// ballot / syncwarp per 32-column chunk.  Here a thread owns a molecule and walks its
// columns against a private triangular pair bitmap, so one warp instruction advances 32
// molecules; everything that touches HBM stays warp-cooperative and coalesced:
//   1. stage   the tile's columns are ONE contiguous run of (src,dst,attr): the CTA reads it
//              with 16-byte loads and packs each column into 15 bits (u6,v6,bond3) of a
//              16-bit word in shared memory; atoms become 4-bit codes;
//   2. walk    thread t: for each column of molecule t, test/set bit tri(min,max) of its
//              private bitmap (word-major layout [word][thread]: conflict free), compact the
//              first-seen columns in place -> number of bonds -> row length;
//   3. scan    block scan of the lengths; the tile aggregate is published for the decoupled
//              look-back (tokens for the ragged stream, rows*Lmax cells for padded batches);
//   4. expand  thread t writes row t as 16-bit tokens into the tile's staging run;
//   5. flush   the staging run leaves as int64 with coalesced 16-byte streaming stores: one
//              contiguous run (ragged) or row by row with pad fill + mask bytes (padded).
// HBM traffic = algorithmic minimum (12*Ed + 4N + 16 read, 8L written per molecule; padded:
// 9*B*Lmax written).  Shapes: <= 64 atoms, <= 256 columns per molecule.
#include <stdlib.h>

#include "gtok_common.cuh"

namespace gtok {

namespace {

constexpr int kTpmMaxT = 256;

struct TpmLayout {
  int T;          // molecules per tile = threads per CTA (32 per warp, lane j <-> molecule j)
  int nmax;       // node ids are clamped to nmax-1 (<= 63)
  int bm_words;   // 32-bit words of one molecule's triangular pair bitmap
  int col_cap;    // staged columns per WARP (32 * max_entries)
  int x_cap;      // staged atoms per WARP   (32 * max_nodes)
  int row_cap;    // tokens of one row buffer
  int o_col, o_x, o_bm, o_row;   // byte offsets inside a warp's slice
  int pair_cells; // int64 cells of one row-pair slot of the bulk-store flush (at o_row; 0 = no such slot)
  int warp_bytes;
  int smem;
};

struct TpmOut {
  int32_t* len_out;     // [G]
  // ragged
  int64_t* off_out;     // [G+1]
  int64_t* tok_out;
  int64_t tok_capacity;
  // padded (tile == batch)
  int32_t* lmax_out;    // [nb]
  int64_t* xoff_out;    // [nb+1]
  int64_t* X;
  uint8_t* attn;
  int64_t x_capacity;
  int64_t pad;
};

// two tokens at an EVEN index of a row buffer in one store
__device__ __forceinline__ void put2(uint16_t* rb, int idx, unsigned t0, unsigned t1) {
  *reinterpret_cast<unsigned*>(rb + idx) = t0 | (t1 << 16);
}
__device__ __forceinline__ void put2(int32_t* rb, int idx, unsigned t0, unsigned t1) {
  *reinterpret_cast<int2*>(rb + idx) = make_int2((int)t0, (int)t1);
}

// TMA (padded batches only): 0 = rows leave through registers (warp_pad_row); 1 / 2 = two consecutive rows are built as
// int64 cells in a shared-memory slot (1 or 2 slots per warp) and leave as ONE bulk asynchronous copy
// (cp.async.bulk.global.shared::cta, SASS UBLKCP): the X half of the flush - 8 of every 9 output bytes - costs one
// instruction per row pair instead of ~60, and the stores drain while the warp builds the next pair.
template <typename StageT, bool HAS_ATTR, bool PADDED, bool UNIQ = false, int TMA = 0>
__global__ void __launch_bounds__(kTpmMaxT)
zinc_tpm_kernel(const GtokGraphBatch b, const __grid_constant__ GtokZincVocab v, const TpmLayout lay, const TpmOut o,
                unsigned int* __restrict__ ticket, unsigned long long* __restrict__ status, const int flags,
                const int pf_dist, const uint16_t* __restrict__ p_col16, const uint8_t* __restrict__ p_x8) {
  constexpr bool unique_cols = UNIQ;                     // the columns are already the first-seen undirected edges
  const int vec_ok = flags & 1;                          // the int32 arrays are 16-byte aligned
  extern __shared__ __align__(16) unsigned char t_smem[];
  __shared__ StageT s_nt[64];
  __shared__ StageT s_atom[10];
  __shared__ StageT s_bond[5];
  __shared__ int s_wtot[2][kTpmMaxT / 32];
  __shared__ int s_wmax[2][kTpmMaxT / 32];
  __shared__ long long s_base[2];
  __shared__ unsigned int s_tile[2];
  __shared__ ulonglong2 s_m16[17];

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int T = lay.T, nwarps = T >> 5;
  unsigned char* wsm = t_smem + (size_t)w * lay.warp_bytes;
  uint16_t* s_col = reinterpret_cast<uint16_t*>(wsm + lay.o_col);
  uint8_t* s_x = wsm + lay.o_x;
  unsigned* s_bm = reinterpret_cast<unsigned*>(wsm + lay.o_bm);
  StageT* s_row = reinterpret_cast<StageT*>(wsm + lay.o_row);
  const int64_t G = b.num_graphs;
  const int item_max = v.item_max_len, ds_max = v.ds_max_len;
  const int32_t* __restrict__ g_src = b.src;
  const int32_t* __restrict__ g_dst = b.dst;
  const int32_t* __restrict__ g_attr = b.edge_attr;
  const int32_t* __restrict__ g_x = b.x;
  const unsigned nclamp = (unsigned)(lay.nmax - 1);

  for (int i = tid; i < 64; i += T) s_nt[i] = (StageT)(i < v.node_tok_len ? __ldg(v.node_tok + i) : v.oov);
  if (tid < 9) s_atom[tid] = (StageT)v.atom_tok[tid];
  if (tid == 9) s_atom[9] = (StageT)v.atom_oov;
  if (tid >= 16 && tid < 20) s_bond[tid - 16] = (StageT)v.bond_tok[tid - 16];
  if (tid == 20) s_bond[4] = (StageT)v.bond_oov;
  if (TMA) fill_mask_table(s_m16, tid);
  if (tid == 0) s_tile[0] = atomicAdd(ticket, 1u);
  const long long E_total = __ldg(b.edge_off + G), N_total = __ldg(b.node_off + G);
  const unsigned tk_bos = (unsigned)v.tok_bos, tk_atom = (unsigned)v.tok_atom, tk_bond = (unsigned)v.tok_bond;
  const int tail_tok = lane == 0 ? v.tok_q : (lane == 1 ? v.tok_regression : v.tok_p);
  __syncthreads();

  for (int par = 0;; par ^= 1) {
    const long long tile = s_tile[par];
    const int64_t t0 = tile * T;
    if (t0 >= G) break;                                  // every warp reads the same ticket: all leave together
    unsigned int my_next = 0;
    if (tid == 0) my_next = atomicAdd(ticket, 1u);       // next tile's ticket: latency hidden behind this tile
    const int nvalid = (int)((G - t0 < T) ? (G - t0) : T);
    const int64_t g0 = t0 + 32 * w;                      // this warp's first molecule
    const int nv = (int)(g0 >= G ? 0 : ((G - g0 < 32) ? (G - g0) : 32));

    // header of the tile `pf_dist` tickets ahead: its inputs are pulled into L2 while this tile computes.  The lead has
    // to cover the prefetch's own latency at the rate tickets are drawn (~11 tiles/us): 256-320 tickets measured best,
    // 128 lands too late (only 28 % of the staging sectors hit L2, profiles/r1l), 512+ is slower again
    long long pe_lo = 0, pe_hi = 0, pn_lo = 0, pn_hi = 0;
    const int64_t pt0 = (tile + pf_dist) * T;
    const bool do_pf = pf_dist > 0 && tid == T - 1 && pt0 < G && vec_ok;
    if (do_pf) {
      const int64_t pt1 = pt0 + T < G ? pt0 + T : G;
      pe_lo = __ldg(b.edge_off + pt0); pe_hi = __ldg(b.edge_off + pt1);
      pn_lo = __ldg(b.node_off + pt0); pn_hi = __ldg(b.node_off + pt1);
    }

    // ---- header: lane j holds molecule j ------------------------------------------------------
    long long e_lo = 0, e_hi = 0, n_lo = 0, n_hi = 0;
    if (lane < nv) {
      e_lo = __ldg(b.edge_off + g0 + lane); e_hi = __ldg(b.edge_off + g0 + lane + 1);
      n_lo = __ldg(b.node_off + g0 + lane); n_hi = __ldg(b.node_off + g0 + lane + 1);
    }
    if (!unique_cols) for (int k = 0; k < lay.bm_words; ++k) s_bm[k * 32 + lane] = 0u;
    const int last = nv > 0 ? nv - 1 : 0;
    const long long E0 = __shfl_sync(kFull, e_lo, 0), N0 = __shfl_sync(kFull, n_lo, 0);
    int Et = (int)(__shfl_sync(kFull, e_hi, last) - E0), Nt = (int)(__shfl_sync(kFull, n_hi, last) - N0);
    // the caller's per-molecule bounds size the staging areas; a batch that violates them yields
    // header-only rows instead of corrupting shared memory
    const bool fits = Et <= lay.col_cap && Nt <= lay.x_cap && Et >= 0 && Nt >= 0;
    if (!fits) { Et = 0; Nt = 0; }
    const bool packed = p_x8 != nullptr;                // columns arrive as 16-bit words (u6 | v6 << 6 | attr4 << 12), atoms as bytes
    const long long E0a = packed ? (E0 & ~7ll) : (vec_ok ? (E0 & ~3ll) : E0);           // staged index 0
    const long long N0a = packed ? (N0 & ~15ll) : (vec_ok ? (N0 & ~3ll) : N0);
    const int eo = (int)(e_lo - E0a), no = (int)(n_lo - N0a);
    const int Ed = fits ? (int)(e_hi - e_lo) : 0;
    const int N = fits ? (int)(n_hi - n_lo) : 0;

    // ---- 1. stage: the warp's columns / atoms are one contiguous run; 16-byte loads, packed stores ----
    {
      auto pack = [&](int a, int c, int t) -> unsigned {
        const unsigned aa = min((unsigned)a, nclamp), cc = min((unsigned)c, nclamp);
        const unsigned bb = HAS_ATTR ? min((unsigned)t, 15u) : 0u;   // raw edge_attr, 4 bits (0 / 5.. = 'unknown')
        return aa | (cc << 6) | (bb << 12);
      };
      if (packed) {
        // the wire / compact resident format IS the staged format: 16-byte copies, nothing to pack
        const long long j0 = E0a >> 3, j1 = (E0 + Et + 7) >> 3, j1v = min(j1, E_total >> 3);
        const uint4* __restrict__ c4g = reinterpret_cast<const uint4*>(p_col16);
        uint4* c4 = reinterpret_cast<uint4*>(s_col);
#pragma unroll 4
        for (long long j = j0 + lane; j < j1v; j += 32) c4[j - j0] = __ldcs(c4g + j);
        if (j1v < j1) {
          const long long e = (j1v << 3) + lane;
          if (lane < 8 && e < E_total) s_col[e - E0a] = p_col16[e];
        }
        const long long i0 = N0a >> 4, i1 = (N0 + Nt + 15) >> 4, i1v = min(i1, N_total >> 4);
        const uint4* __restrict__ x4g = reinterpret_cast<const uint4*>(p_x8);
        uint4* x4s = reinterpret_cast<uint4*>(s_x);
        for (long long j = i0 + lane; j < i1v; j += 32) x4s[j - i0] = __ldcs(x4g + j);
        if (i1v < i1) {
          const long long e = (i1v << 4) + lane;
          if (lane < 16 && e < N_total) s_x[e - N0a] = p_x8[e];
        }
      } else if (vec_ok) {
        const long long j0 = E0a >> 2, j1 = (E0 + Et + 3) >> 2, j1v = min(j1, E_total >> 2);
        const int4* __restrict__ s4 = reinterpret_cast<const int4*>(g_src);
        const int4* __restrict__ d4 = reinterpret_cast<const int4*>(g_dst);
        const int4* __restrict__ a4 = reinterpret_cast<const int4*>(g_attr);
        uint2* c2 = reinterpret_cast<uint2*>(s_col);
#pragma unroll 4
        for (long long j = j0 + lane; j < j1v; j += 32) {
          const int4 A = __ldcs(s4 + j);
          const int4 Cc = __ldcs(d4 + j);
          int4 Tt = make_int4(0, 0, 0, 0);
          if (HAS_ATTR) Tt = __ldcs(a4 + j);
          c2[j - j0] = make_uint2(pack(A.x, Cc.x, Tt.x) | (pack(A.y, Cc.y, Tt.y) << 16),
                                  pack(A.z, Cc.z, Tt.z) | (pack(A.w, Cc.w, Tt.w) << 16));
        }
        if (j1v < j1) {                                  // the array's last, partial group (last warp of the batch only)
          const long long e = (j1v << 2) + lane;
          if (lane < 4 && e < E_total)
            s_col[e - E0a] = (uint16_t)pack(__ldcs(g_src + e), __ldcs(g_dst + e), HAS_ATTR ? __ldcs(g_attr + e) : 0);
        }
        const long long i0 = N0a >> 2, i1 = (N0 + Nt + 3) >> 2, i1v = min(i1, N_total >> 2);
        const int4* __restrict__ x4 = reinterpret_cast<const int4*>(g_x);
        unsigned* x1 = reinterpret_cast<unsigned*>(s_x);
#pragma unroll 2
        for (long long j = i0 + lane; j < i1v; j += 32) {
          const int4 X4 = __ldcs(x4 + j);
          x1[j - i0] = min((unsigned)X4.x, 9u) | (min((unsigned)X4.y, 9u) << 8) | (min((unsigned)X4.z, 9u) << 16) |
                       (min((unsigned)X4.w, 9u) << 24);
        }
        if (i1v < i1) {
          const long long e = (i1v << 2) + lane;
          if (lane < 4 && e < N_total) s_x[e - N0a] = (uint8_t)min((unsigned)__ldcs(g_x + e), 9u);
        }
      } else {                                           // arrays not 16-byte aligned: scalar coalesced loads
#pragma unroll 4
        for (int i = lane; i < Et; i += 32)
          s_col[i] = (uint16_t)pack(__ldcs(g_src + E0 + i), __ldcs(g_dst + E0 + i), HAS_ATTR ? __ldcs(g_attr + E0 + i) : 0);
#pragma unroll 4
        for (int i = lane; i < Nt; i += 32) s_x[i] = (uint8_t)min((unsigned)__ldcs(g_x + N0 + i), 9u);
      }
    }
    if (do_pf) {
      auto pf = [&](const int32_t* p, long long lo, long long hi, long long total) {
        long long a = lo & ~3ll, z = (hi + 3) & ~3ll;
        if (z > total) z = total & ~3ll;
        if (z > a) {
          const unsigned bytes = (unsigned)((z - a) * 4);
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + a), "r"(bytes) : "memory");
        }
      };
      if (packed) {
        auto pfb = [&](const void* p, long long lo, long long hi, long long total, int esz) {   // byte ranges, 16-byte granules
          long long a = (lo * esz) & ~15ll, z = (hi * esz + 15) & ~15ll;
          if (z > total * esz) z = (total * esz) & ~15ll;
          if (z > a)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(static_cast<const char*>(p) + a), "r"((unsigned)(z - a)) : "memory");
        };
        pfb(p_col16, pe_lo, pe_hi, E_total, 2);
        pfb(p_x8, pn_lo, pn_hi, N_total, 1);
      } else {
        pf(g_src, pe_lo, pe_hi, E_total);
        pf(g_dst, pe_lo, pe_hi, E_total);
        if (HAS_ATTR) pf(g_attr, pe_lo, pe_hi, E_total);
        pf(g_x, pn_lo, pn_hi, N_total);
      }
    }
    __syncwarp();

    // ---- 2. walk: thread per molecule, first-seen dedupe, in-place compaction ------------------
    uint16_t* colp = s_col + eo;
    int cnt = 0;
    if (unique_cols) {
      cnt = Ed;                                          // the caller sent every undirected edge once, in first-seen order
    } else {
      unsigned* bm = s_bm + lane;
#pragma unroll 1
      for (int i = 0; i < Ed; ++i) {
        const unsigned wv = colp[i];
        const unsigned a = min(wv & 63u, nclamp), c = min((wv >> 6) & 63u, nclamp);   // (packed input is not clamped at staging)
        const unsigned mn = min(a, c), mx = max(a, c);
        const unsigned idx = ((mx * (mx + 1u)) >> 1) + mn;
        const unsigned wd = (idx >> 5) * 32u, bit = 1u << (idx & 31u);
        const unsigned old = bm[wd];
        if (!(old & bit)) {
          bm[wd] = old | bit;
          colp[cnt++] = (uint16_t)wv;
        }
      }
    }
    int my_len = 0;
    const int full_T = 2 * N + 4 * cnt + 6;              // text length incl. label + <eos> (oracle: zinc_length)
    if (lane < nv) {
      const int L = full_T <= item_max ? full_T - 2 : (full_T == item_max + 1 ? item_max - 1 : item_max);
      my_len = max(0, min(L, ds_max));
    }

    // ---- 3. warp scan, ONE CTA barrier per tile, look-back by warp 0 ----------------------------
    int incl = my_len, wmx = my_len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int up = __shfl_up_sync(kFull, incl, d);
      if (lane >= d) incl += up;
      wmx = max(wmx, __shfl_xor_sync(kFull, wmx, d));
    }
    if (lane == 31) { s_wtot[par][w] = incl; s_wmax[par][w] = wmx; }
    if (tid == 0) s_tile[par ^ 1] = my_next;
    __syncthreads();                                     // totals of every warp + next ticket visible
    int wexcl = 0, ttotal = 0, lmax = 0;
    for (int k = 0; k < nwarps; ++k) {
      const int tk = s_wtot[par][k];
      if (k < w) wexcl += tk;
      ttotal += tk;
      lmax = max(lmax, s_wmax[par][k]);
    }
    const int my_off = wexcl + incl - my_len;            // tile-local token offset of this lane's row
    const long long pub = PADDED ? (long long)nvalid * lmax : (long long)ttotal;
    // The other warps pick the tile's base up at named barrier 1: warp 0 ARRIVES there once the look-back is done
    // (it never waits), they SYNC on it right before their first store.  A warp parked at a hardware barrier
    // issues nothing - polling a shared flag instead cost 28 % of all issued instructions (profiles/r1l).
    long long base = -1;
    bool have_base = false;
    if (w == 0) {
      const long long tb = lookback_exclusive(status, tile, pub);
      if (lane == 0) {
        s_base[par] = tb;
        if (t0 + T >= G) {
          if (PADDED) o.xoff_out[tile + 1] = tb + pub;
          else o.off_out[G] = tb + pub;
        }
      }
      __threadfence_block();
      if (nwarps > 1) asm volatile("bar.arrive 1, %0;" ::"r"(T) : "memory");
      base = tb;
      have_base = true;
    }
    auto wait_base = [&]() {
      if (!have_base) {
        asm volatile("bar.sync 1, %0;" ::"r"(T) : "memory");
        base = *reinterpret_cast<volatile long long*>(&s_base[par]);
        have_base = true;
      }
    };

    // ---- 4. rows: the warp builds row j in a row buffer (lane i <-> atom i / bond i), then streams it out ----
    auto build = [&](int j, StageT* rb) {                // tokens of row j at rb[1 + r]
      const int rN = __shfl_sync(kFull, N, j), rcnt = __shfl_sync(kFull, cnt, j);
      const int reo = __shfl_sync(kFull, eo, j), rno = __shfl_sync(kFull, no, j);
      if (lane == 0) rb[1] = (StageT)tk_bos;
      for (int i = lane; i < rN; i += 32) put2(rb, 2 + 2 * i, tk_atom, (unsigned)s_atom[min((unsigned)s_x[rno + i], 9u)]);
      const int bb = 2 + 2 * rN;
      for (int k = lane; k < rcnt; k += 32) {
        const unsigned wv = s_col[reo + k];
        put2(rb, bb + 4 * k, tk_bond, (unsigned)s_bond[min((wv >> 12) - 1u, 4u)]);
        put2(rb, bb + 4 * k + 2, (unsigned)s_nt[wv & 63u], (unsigned)s_nt[(wv >> 6) & 63u]);
      }
      if (lane < 3) rb[bb + 4 * rcnt + lane] = (StageT)tail_tok;
    };
    auto fix_eos = [&](int j, StageT* rb) {              // truncated text keeps a trailing <eos> (:216-221)
      const int rfull = __shfl_sync(kFull, full_T, j);
      if (rfull >= item_max + 2) {
        __syncwarp();
        if (lane == 0) rb[item_max] = (StageT)v.tok_eos;  // row position item_max-1
      }
    };
    auto flush = [&](int j, const StageT* rb) {
      const int len = __shfl_sync(kFull, my_len, j);
      if (!PADDED) {
        const long long pos = base + __shfl_sync(kFull, my_off, j);
        if (pos + len > o.tok_capacity) return;
        int64_t* __restrict__ dst = o.tok_out + pos;
        const int head = (int)(pos & 1);
        if (head && lane == 0 && len > 0) st_stream_i64(dst, (int64_t)rb[1]);
        const int npair = (len - head) >> 1;
        for (int q = lane; q < npair; q += kWarp) {
          const int r = head + 2 * q;
          st_stream_i64x2(dst + r, (int64_t)rb[1 + r], (int64_t)rb[2 + r]);
        }
        const int tl = head + 2 * npair;
        if (tl < len && lane == 31) st_stream_i64(dst + tl, (int64_t)rb[1 + tl]);
      } else {
        if (base + pub > o.x_capacity) return;
        const long long cell = base + (long long)(32 * w + j) * lmax;
        warp_pad_row(len, lmax, o.pad, o.X + cell, o.attn + cell, [&](int p) -> int { return (int)rb[1 + p]; });
      }
    };
    StageT* rb0 = s_row;
    StageT* rb1 = s_row + lay.row_cap;
    int j_first = 0;
    if (TMA && PADDED && lmax > 0 && 2 * lmax <= lay.pair_cells) {
      // ---- 4'. row pairs as int64 cells in shared memory, one bulk store per pair --------------------------------
      // Rows 32w+j, 32w+j+1 (j even) are one contiguous run of 2*lmax cells that starts at an EVEN cell (every
      // earlier tile holds T rows, T even): 16-byte aligned, a multiple of 16 bytes long - what a bulk copy needs.
      int64_t* slots = reinterpret_cast<int64_t*>(s_row);
      const long long padv = o.pad;
      const bool l2_hint = (flags & 2) != 0;
      const unsigned long long l2_pol = l2_policy_evict_first();
      // `prev` = length of the row this sub-slot held before (lmax on first use in the tile): cells [prev, lmax) still
      // hold pad, so only [len, prev) has to be refilled
      auto build64 = [&](int j, int64_t* rb, int prev) -> int {   // cells [0, lmax) of row j: tokens below len, pad beyond
        const int rN = __shfl_sync(kFull, N, j), rcnt = __shfl_sync(kFull, cnt, j);
        const int reo = __shfl_sync(kFull, eo, j), rno = __shfl_sync(kFull, no, j);
        const int len = __shfl_sync(kFull, my_len, j), rfull = __shfl_sync(kFull, full_T, j);
        const int bb = 1 + 2 * rN;
        if (rfull <= item_max && len == rfull - 2) {       // untruncated text, whole row kept (the common case): unguarded stores
          if (lane == 0) rb[0] = (long long)tk_bos;
          for (int i = lane; i < rN; i += 32) {
            rb[1 + 2 * i] = (long long)tk_atom;
            rb[2 + 2 * i] = (long long)s_atom[min((unsigned)s_x[rno + i], 9u)];
          }
          for (int k = lane; k < rcnt; k += 32) {
            const unsigned wv = s_col[reo + k];
            int64_t* q = rb + bb + 4 * k;
            q[0] = (long long)tk_bond;
            q[1] = (long long)s_bond[min((wv >> 12) - 1u, 4u)];
            q[2] = (long long)s_nt[wv & 63u];
            q[3] = (long long)s_nt[(wv >> 6) & 63u];
          }
          if (lane < 3) rb[bb + 4 * rcnt + lane] = (long long)tail_tok;
        } else {                                           // truncated row: every store guarded, <eos> kept (:216-221)
          const int eos_pos = rfull >= item_max + 2 ? item_max - 1 : -1;
          auto put = [&](int p, unsigned t) { if (p < len && p != eos_pos) rb[p] = (long long)t; };
          if (lane == 0) {
            put(0, tk_bos);
            if (eos_pos >= 0 && eos_pos < len) rb[eos_pos] = (long long)v.tok_eos;
          }
          for (int i = lane; i < rN; i += 32) {
            put(1 + 2 * i, tk_atom);
            put(2 + 2 * i, (unsigned)s_atom[min((unsigned)s_x[rno + i], 9u)]);
          }
          for (int k = lane; k < rcnt; k += 32) {
            const unsigned wv = s_col[reo + k];
            put(bb + 4 * k, tk_bond);
            put(bb + 4 * k + 1, (unsigned)s_bond[min((wv >> 12) - 1u, 4u)]);
            put(bb + 4 * k + 2, (unsigned)s_nt[wv & 63u]);
            put(bb + 4 * k + 3, (unsigned)s_nt[(wv >> 6) & 63u]);
          }
          if (lane < 3) put(bb + 4 * rcnt + lane, (unsigned)tail_tok);
        }
        for (int p = len + lane; p < prev; p += 32) rb[p] = padv;
        return len;
      };
      int sl = 0;
      int pA0 = lmax, pB0 = lmax, pA1 = lmax, pB1 = lmax;
#pragma unroll 1
      for (; j_first + 1 < nv; j_first += 2) {
        const int j = j_first;
        int64_t* rb = slots + (TMA == 2 ? sl * lay.pair_cells : 0);
        if (lane == 0) {                                   // the slot's previous copy has read its bytes
          if (TMA == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
        }
        __syncwarp();
        const int lenA = build64(j, rb, sl ? pA1 : pA0);
        const int lenB = build64(j + 1, rb + lmax, sl ? pB1 : pB0);
        if (sl) { pA1 = lenA; pB1 = lenB; } else { pA0 = lenA; pB0 = lenB; }
        fence_proxy_async_smem();
        __syncwarp();
        wait_base();
        if (base + pub <= o.x_capacity) {
          const long long cell = base + (long long)(32 * w + j) * lmax;
          if (lane == 0) {
            if (l2_hint) bulk_store_s2g_hint(o.X + cell, rb, (unsigned)(2 * lmax) * 8u, l2_pol);
            else bulk_store_s2g(o.X + cell, rb, (unsigned)(2 * lmax) * 8u);
            bulk_commit();
          }
          warp_mask_run2(lenA, lenB, lmax, o.attn + cell, s_m16);
        }
        if (TMA == 2) sl ^= 1;
      }
      if (lane == 0) bulk_wait_read<0>();                  // the slots alias the row buffers / next tile's pair bitmaps
      __syncwarp();
    }
#pragma unroll 1
    for (int j = j_first; j < nv; j += 2) {
      const bool two = j + 1 < nv;
      build(j, rb0);
      if (two) build(j + 1, rb1);
      fix_eos(j, rb0);
      if (two) fix_eos(j + 1, rb1);
      __syncwarp();
      wait_base();
      flush(j, rb0);
      if (two) flush(j + 1, rb1);
      __syncwarp();                                      // row buffers drained before the next pair overwrites them
    }
    wait_base();                                         // every warp passes barrier 1 exactly once per tile
    if (nv > 0) {
      if (lane < nv) {
        o.len_out[g0 + lane] = my_len;
        if (!PADDED) o.off_out[g0 + lane] = base + my_off;
      }
      if (PADDED && tid == 0) { o.lmax_out[tile] = lmax; o.xoff_out[tile] = base; }
    }
  }
  if (TMA && lane == 0) bulk_wait<0>();                  // every bulk store of this warp has been performed
}

int tpm_prefetch_dist() {
  static const int d = [] { const char* e = getenv("GTOK_ZINC_PREFETCH"); return e ? atoi(e) : 288; }();
  return d;
}

// Shared-memory plan for tiles of T molecules; false if it does not fit.
template <typename StageT>
bool tpm_layout(int T, int max_nodes, int max_entries, int smem_optin, TpmLayout* out, int tma_slots = 0,
                int group_nodes = 0, int group_entries = 0) {
  TpmLayout L;
  L.T = T;
  L.nmax = max_nodes < 1 ? 1 : max_nodes;
  L.bm_words = (L.nmax * (L.nmax + 1) / 2 + 31) / 32;
  L.col_cap = 32 * max_entries;
  L.x_cap = 32 * L.nmax;
  // a warp stages the 32 molecules of one aligned group: the caller's per-group bounds, when given, are tighter than
  // 32 x the per-molecule bounds (ZINC-shaped data: 1.8k instead of 2.5k columns - one more resident CTA per SM)
  if (group_entries > 0 && group_entries < L.col_cap) L.col_cap = (group_entries + 7) & ~7;
  if (group_nodes > 0 && group_nodes < L.x_cap) L.x_cap = (group_nodes + 15) & ~15;
  L.row_cap = (2 * L.nmax + 4 * max_entries + 8 + 7) & ~7;      // untruncated row + slack, keeps buffers 16-byte aligned
  int o = 0;
  L.o_col = o; o += ((L.col_cap + 16) * 2 + 15) & ~15;          // slack: the aligned first / last 16-byte group may spill over
  L.o_x = o; o += (L.x_cap + 32 + 15) & ~15;
  // the pair bitmaps die with the walk, the row buffers are born after it (same warp, program order): same bytes
  const int bm_bytes = 32 * L.bm_words * 4, row_bytes = 2 * L.row_cap * (int)sizeof(StageT);
  L.o_bm = o;
  L.o_row = o;
  // bulk-store flush: a slot holds two rows as int64 cells.  Bounded at 512 cells (4 KB): a tile whose Lmax does not
  // fit takes the register flush, so the bound only decides which tiles get the fast path.
  L.pair_cells = 0;
  int pair_bytes = 0;
  if (tma_slots > 0) {
    L.pair_cells = 2 * L.row_cap < 512 ? 2 * L.row_cap : 512;
    pair_bytes = tma_slots * L.pair_cells * 8;
  }
  int un = bm_bytes > row_bytes ? bm_bytes : row_bytes;
  if (pair_bytes > un) un = pair_bytes;
  o += un;
  L.warp_bytes = (o + 15) & ~15;
  L.smem = (T / 32) * L.warp_bytes;
  if (L.smem + 2048 > smem_optin) return false;
  *out = L;
  return true;
}

template <typename StageT, bool PADDED>
int tpm_launch(const GtokGraphBatch* batch, const GtokZincVocab* vocab, const TpmOut& o, int T_fixed, void* ws,
               size_t ws_bytes, cudaStream_t st, bool* handled, const uint16_t* col16 = nullptr, const uint8_t* x8 = nullptr,
               int unique_cols = 0) {
  *handled = false;
  int dev = 0, optin = 0;
  GTOK_CUDA_TRY(cudaGetDevice(&dev));
  GTOK_CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  TpmLayout lay;
  bool ok = false;
  // bulk-store flush, one slot per warp (measured on 1M ZINC-shaped molecules, same box: register flush 0.705 ms,
  // one slot 0.594 ms, two slots 0.649 ms - the second slot costs resident CTAs); GTOK_ZINC_TMA=0/1/2 overrides
  static const int tma_env = [] { const char* e = getenv("GTOK_ZINC_TMA"); return e ? atoi(e) : -1; }();
  int tma = PADDED ? (tma_env < 0 ? 1 : (tma_env > 2 ? 2 : tma_env)) : 0;
  if (T_fixed > 0) {
    const int gn = batch->max_nodes_per_32, ge = batch->max_entries_per_32;
    ok = tma && tpm_layout<StageT>(T_fixed, batch->max_nodes, batch->max_entries, optin, &lay, tma, gn, ge);
    if (!ok) {
      tma = 0;                                           // (register flush: more resident CTAs only lengthen the look-back
      ok = tpm_layout<StageT>(T_fixed, batch->max_nodes, batch->max_entries, optin, &lay);   //  chain - measured slower)
    }
  } else {
    static const int t_env = [] { const char* e = getenv("GTOK_ZINC_TILE"); return e ? atoi(e) : 0; }();
    const int cand[4] = {(t_env >= 32 && t_env <= kTpmMaxT && t_env % 32 == 0) ? t_env : 128, 128, 64, 32};
    for (int i = 0; i < 4 && !ok; ++i)
      ok = tpm_layout<StageT>(cand[i], batch->max_nodes, batch->max_entries, optin, &lay);
  }
  if (!ok) return GTOK_OK;
  const int64_t tiles = (batch->num_graphs + lay.T - 1) / lay.T;
  const size_t need = 16 + (size_t)tiles * 8;
  GTOK_REQUIRE(ws_bytes >= need, GTOK_EWORKSPACE, "zinc tokenizer: workspace %zu < %zu bytes", ws_bytes, need);
  auto k = (batch->edge_attr || x8) ? zinc_tpm_kernel<StageT, true, PADDED> : zinc_tpm_kernel<StageT, false, PADDED>;
  if (PADDED && x8 && unique_cols) k = zinc_tpm_kernel<StageT, true, PADDED, PADDED>;   // (UNIQ only exists for the padded packed path)
  if (PADDED && tma) {
    const bool uq = x8 && unique_cols;
    if (tma == 1) {
      k = (batch->edge_attr || x8) ? zinc_tpm_kernel<StageT, true, PADDED, false, PADDED ? 1 : 0>
                                   : zinc_tpm_kernel<StageT, false, PADDED, false, PADDED ? 1 : 0>;
      if (uq) k = zinc_tpm_kernel<StageT, true, PADDED, PADDED, PADDED ? 1 : 0>;
    } else {
      k = (batch->edge_attr || x8) ? zinc_tpm_kernel<StageT, true, PADDED, false, PADDED ? 2 : 0>
                                   : zinc_tpm_kernel<StageT, false, PADDED, false, PADDED ? 2 : 0>;
      if (uq) k = zinc_tpm_kernel<StageT, true, PADDED, PADDED, PADDED ? 2 : 0>;
    }
  }
  GTOK_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.smem));
  GTOK_CUDA_TRY(cudaMemsetAsync(ws, 0, need, st));
  const int grid = persistent_grid(k, lay.T, (size_t)lay.smem, tiles);
  const int vec_ok = x8 ? 1 : ((reinterpret_cast<uintptr_t>(batch->src) | reinterpret_cast<uintptr_t>(batch->dst) |
                                   reinterpret_cast<uintptr_t>(batch->edge_attr) | reinterpret_cast<uintptr_t>(batch->x)) & 15) == 0;
  static const int hint_env = [] { const char* e = getenv("GTOK_ZINC_L2HINT"); return e ? atoi(e) : 1; }();
  k<<<grid, lay.T, lay.smem, st>>>(*batch, *vocab, lay, o, reinterpret_cast<unsigned int*>(ws),
                                   reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ws) + 16), vec_ok | (hint_env ? 2 : 0),
                                   tpm_prefetch_dist(), col16, x8);
  GTOK_CUDA_TRY(cudaGetLastError());
  *handled = true;
  return GTOK_OK;
}

}  // namespace

// Ragged single pass behind gtok_zinc_tokenize (below).
int zinc_tpm_ragged(const GtokGraphBatch* batch, const GtokZincVocab* vocab, int32_t* len_out, int64_t* off_out,
                    int64_t* tok_out, int64_t tok_capacity, void* ws, size_t ws_bytes, cudaStream_t st, bool* handled) {
  TpmOut o{};
  o.len_out = len_out; o.off_out = off_out; o.tok_out = tok_out; o.tok_capacity = tok_capacity;
  if (vocab->max_token_id > 0 && vocab->max_token_id < 65536)
    return tpm_launch<uint16_t, false>(batch, vocab, o, 0, ws, ws_bytes, st, handled);
  return tpm_launch<int32_t, false>(batch, vocab, o, 0, ws, ws_bytes, st, handled);
}

}  // namespace gtok

extern "C" size_t gtok_zinc_collate_workspace_bytes(int64_t num_graphs, int64_t batch_size) {
  if (batch_size < 1) batch_size = 1;
  const int64_t nb = (num_graphs + batch_size - 1) / batch_size;
  return 16 + (size_t)(nb > 0 ? nb : 1) * 8;
}

extern "C" int gtok_zinc_collate(const GtokGraphBatch* batch, const GtokZincVocab* vocab, int64_t batch_size,
                                 int64_t pad_id, int32_t* len_out, int32_t* lmax_out, int64_t* xoff_out, int64_t* X,
                                 uint8_t* attn, int64_t x_capacity, void* ws, size_t ws_bytes, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(batch && vocab, GTOK_EINVAL, "zinc_collate: null batch/vocab");
  GTOK_REQUIRE(batch->num_graphs >= 0, GTOK_EINVAL, "zinc_collate: negative graph count");
  GTOK_REQUIRE(xoff_out != nullptr, GTOK_EINVAL, "zinc_collate: null xoff_out");
  cudaStream_t st = (cudaStream_t)stream;
  if (batch->num_graphs == 0) {
    GTOK_CUDA_TRY(cudaMemsetAsync(xoff_out, 0, sizeof(int64_t), st));
    return GTOK_OK;
  }
  GTOK_REQUIRE(batch->node_off && batch->edge_off && batch->x, GTOK_EINVAL,
               "zinc_collate: null graph arrays (x is required; src / dst may be null without edges)");
  GTOK_REQUIRE(len_out && lmax_out && X && attn && ws, GTOK_EINVAL, "zinc_collate: null output/workspace");
  GTOK_REQUIRE(x_capacity >= 0, GTOK_EINVAL, "zinc_collate: negative capacity");
  GTOK_REQUIRE(vocab->item_max_len >= 2 && vocab->ds_max_len >= 0, GTOK_EINVAL, "zinc_collate: max_len must be >= 2");
  GTOK_REQUIRE(vocab->node_tok_len == 0 || vocab->node_tok != nullptr, GTOK_EINVAL, "zinc_collate: null node_tok table");
  GTOK_REQUIRE(batch->max_nodes > 0 && batch->max_entries > 0, GTOK_EINVAL, "zinc_collate: max_nodes/max_entries bounds required");
  GTOK_REQUIRE(batch_size >= 32 && batch_size <= kTpmMaxT && (batch_size & 31) == 0, GTOK_EUNSUPPORTED,
               "zinc_collate: batch size must be a multiple of 32 in [32, %d] (got %lld); use gtok_zinc_tokenize + "
               "gtok_pad_mask_batched", kTpmMaxT, (long long)batch_size);
  GTOK_REQUIRE(batch->max_nodes <= 64 && batch->max_entries <= 256, GTOK_EUNSUPPORTED,
               "zinc_collate: covers molecules with <= 64 atoms and <= 256 edge_index columns (got bounds %d, %d)",
               batch->max_nodes, batch->max_entries);
  GTOK_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(ws) & 15) == 0, GTOK_EINVAL,
               "zinc_collate: X and workspace must be 16-byte aligned");
  TpmOut o{};
  o.len_out = len_out; o.lmax_out = lmax_out; o.xoff_out = xoff_out; o.X = X; o.attn = attn;
  o.x_capacity = x_capacity; o.pad = pad_id;
  bool handled = false;
  int rc;
  if (vocab->max_token_id > 0 && vocab->max_token_id < 65536)
    rc = tpm_launch<uint16_t, true>(batch, vocab, o, (int)batch_size, ws, ws_bytes, st, &handled);
  else
    rc = tpm_launch<int32_t, true>(batch, vocab, o, (int)batch_size, ws, ws_bytes, st, &handled);
  if (rc != GTOK_OK) return rc;
  GTOK_REQUIRE(handled, GTOK_EUNSUPPORTED, "zinc_collate: a tile of %lld molecules with these bounds does not fit shared memory",
               (long long)batch_size);
  return GTOK_OK;
}

extern "C" size_t gtok_zinc_tokenize_workspace_bytes(int64_t num_graphs) {
  const int64_t tiles = (num_graphs + 31) / 32;           // smallest tile any ZINC kernel uses
  return 16 + (size_t)(tiles > 0 ? tiles : 1) * 8;
}

extern "C" int gtok_zinc_tokenize(const GtokGraphBatch* batch, const GtokZincVocab* vocab, int32_t* len_out,
                                  int64_t* off_out, int64_t* tok_out, int64_t tok_capacity, void* ws,
                                  size_t ws_bytes, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(batch && vocab, GTOK_EINVAL, "zinc_tokenize: null batch/vocab");
  GTOK_REQUIRE(batch->num_graphs >= 0, GTOK_EINVAL, "zinc_tokenize: negative graph count");
  GTOK_REQUIRE(off_out != nullptr, GTOK_EINVAL, "zinc_tokenize: null off_out");
  cudaStream_t st = (cudaStream_t)stream;
  if (batch->num_graphs == 0) {
    GTOK_CUDA_TRY(cudaMemsetAsync(off_out, 0, sizeof(int64_t), st));
    return GTOK_OK;
  }
  GTOK_REQUIRE(batch->node_off && batch->edge_off && batch->x, GTOK_EINVAL,
               "zinc_tokenize: null graph arrays (x is required; src / dst may be null without edges)");
  GTOK_REQUIRE(len_out && tok_out && ws, GTOK_EINVAL, "zinc_tokenize: null output/workspace");
  GTOK_REQUIRE(tok_capacity >= 0, GTOK_EINVAL, "zinc_tokenize: negative capacity");
  GTOK_REQUIRE(vocab->item_max_len >= 2 && vocab->ds_max_len >= 0, GTOK_EINVAL, "zinc_tokenize: max_len must be >= 2");
  GTOK_REQUIRE(vocab->node_tok_len == 0 || vocab->node_tok != nullptr, GTOK_EINVAL, "zinc_tokenize: null node_tok table");
  GTOK_REQUIRE(batch->max_nodes > 0 && batch->max_entries > 0, GTOK_EINVAL, "zinc_tokenize: max_nodes/max_entries bounds required");
  GTOK_REQUIRE(batch->max_nodes <= 64 && batch->max_entries <= 256, GTOK_EUNSUPPORTED,
               "zinc_tokenize: fused path covers molecules with <= 64 atoms and <= 256 edge_index columns "
               "(got bounds %d, %d); use gtok_zinc_lengths + gtok_scan_offsets + gtok_zinc_serialize",
               batch->max_nodes, batch->max_entries);
  GTOK_REQUIRE((reinterpret_cast<uintptr_t>(tok_out) & 15) == 0 && (reinterpret_cast<uintptr_t>(ws) & 15) == 0,
               GTOK_EINVAL, "zinc_tokenize: tok_out and workspace must be 16-byte aligned");
  const size_t need = gtok_zinc_tokenize_workspace_bytes(batch->num_graphs);
  GTOK_REQUIRE(ws_bytes >= need, GTOK_EWORKSPACE, "zinc_tokenize: workspace %zu < %zu bytes", ws_bytes, need);
  bool handled = false;
  const int rc = zinc_tpm_ragged(batch, vocab, len_out, off_out, tok_out, tok_capacity, ws, ws_bytes, st, &handled);
  if (rc != GTOK_OK) return rc;
  GTOK_REQUIRE(handled, GTOK_EUNSUPPORTED, "zinc_tokenize: these bounds (%d atoms, %d columns) do not fit shared memory; use "
               "gtok_zinc_lengths + gtok_scan_offsets + gtok_zinc_serialize", batch->max_nodes, batch->max_entries);
  return GTOK_OK;
}

extern "C" int gtok_zinc_collate_packed(const GtokPackedMolBatch* pk, const GtokZincVocab* vocab, int64_t batch_size,
                                        int64_t pad_id, int32_t* len_out, int32_t* lmax_out, int64_t* xoff_out, int64_t* X,
                                        uint8_t* attn, int64_t x_capacity, void* ws, size_t ws_bytes, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(pk && vocab, GTOK_EINVAL, "zinc_collate_packed: null batch/vocab");
  GTOK_REQUIRE(pk->num_graphs >= 0, GTOK_EINVAL, "zinc_collate_packed: negative graph count");
  GTOK_REQUIRE(xoff_out != nullptr, GTOK_EINVAL, "zinc_collate_packed: null xoff_out");
  cudaStream_t st = (cudaStream_t)stream;
  if (pk->num_graphs == 0) {
    GTOK_CUDA_TRY(cudaMemsetAsync(xoff_out, 0, sizeof(int64_t), st));
    return GTOK_OK;
  }
  GTOK_REQUIRE(pk->node_off && pk->edge_off && pk->x8, GTOK_EINVAL, "zinc_collate_packed: null graph arrays");   // col16 may be null without edges
  GTOK_REQUIRE(len_out && lmax_out && X && attn && ws, GTOK_EINVAL, "zinc_collate_packed: null output/workspace");
  GTOK_REQUIRE(x_capacity >= 0, GTOK_EINVAL, "zinc_collate_packed: negative capacity");
  GTOK_REQUIRE(vocab->item_max_len >= 2 && vocab->ds_max_len >= 0, GTOK_EINVAL, "zinc_collate_packed: max_len must be >= 2");
  GTOK_REQUIRE(vocab->node_tok_len == 0 || vocab->node_tok != nullptr, GTOK_EINVAL, "zinc_collate_packed: null node_tok table");
  GTOK_REQUIRE(pk->max_nodes > 0 && pk->max_entries > 0, GTOK_EINVAL, "zinc_collate_packed: max_nodes/max_entries bounds required");
  GTOK_REQUIRE(batch_size >= 32 && batch_size <= kTpmMaxT && (batch_size & 31) == 0, GTOK_EUNSUPPORTED,
               "zinc_collate_packed: batch size must be a multiple of 32 in [32, %d] (got %lld)", kTpmMaxT, (long long)batch_size);
  GTOK_REQUIRE(pk->max_nodes <= 64 && pk->max_entries <= 256, GTOK_EUNSUPPORTED,
               "zinc_collate_packed: covers molecules with <= 64 atoms and <= 256 edge_index columns (got bounds %d, %d)",
               pk->max_nodes, pk->max_entries);
  GTOK_REQUIRE(((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(ws) | reinterpret_cast<uintptr_t>(pk->col16) |
                 reinterpret_cast<uintptr_t>(pk->x8)) & 15) == 0, GTOK_EINVAL,
               "zinc_collate_packed: X, workspace, col16 and x8 must be 16-byte aligned");
  GtokGraphBatch b{};
  b.num_graphs = pk->num_graphs; b.node_off = pk->node_off; b.edge_off = pk->edge_off;
  b.max_nodes = pk->max_nodes; b.max_entries = pk->max_entries;
  b.max_nodes_per_32 = pk->max_nodes_per_32; b.max_entries_per_32 = pk->max_entries_per_32;
  TpmOut o{};
  o.len_out = len_out; o.lmax_out = lmax_out; o.xoff_out = xoff_out; o.X = X; o.attn = attn;
  o.x_capacity = x_capacity; o.pad = pad_id;
  bool handled = false;
  int rc;
  if (vocab->max_token_id > 0 && vocab->max_token_id < 65536)
    rc = tpm_launch<uint16_t, true>(&b, vocab, o, (int)batch_size, ws, ws_bytes, st, &handled, pk->col16, pk->x8, pk->unique_columns != 0);
  else
    rc = tpm_launch<int32_t, true>(&b, vocab, o, (int)batch_size, ws, ws_bytes, st, &handled, pk->col16, pk->x8, pk->unique_columns != 0);
  if (rc != GTOK_OK) return rc;
  GTOK_REQUIRE(handled, GTOK_EUNSUPPORTED, "zinc_collate_packed: a tile of %lld molecules with these bounds does not fit shared memory",
               (long long)batch_size);
  return GTOK_OK;
}
