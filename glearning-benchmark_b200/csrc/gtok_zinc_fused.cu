// gtok_zinc_fused.cu - K2 single pass: molecules -> (lengths, ragged offsets, int64 rows).
//
// Same reference functions as gtok_zinc.cu (zinc_dataset_indexbase.py:143-227 +
// data_loader.py:465-486), fused with K0 so every input byte is read from HBM once
// and nothing but the final rows / lengths / offsets is written.
//
// Design (the history v1..v4 with ncu evidence is in DESIGN.md / profiles/):
//   * a CTA (8 warps) takes the next tile of 256 CONSECUTIVE molecules from an atomic
//     ticket (so the look-back can never wait on a tile that has not started); warp w
//     owns molecules 32w..32w+31, lane j holds the header of molecule j;
//   * phase 1 (count): one atomicOr per edge_index column into a 64x64 adjacency bitmap
//     in shared memory + a popcount gives the number of distinct unordered pairs, hence
//     every row length.  The tile total is published for the decoupled look-back
//     IMMEDIATELY, long before any row is produced, so successors never wait on this
//     tile's token generation; warp 0 then resolves the tile's global offset while the
//     other warps already generate rows (they only need it for their first store);
//   * phase 2 (rows): first-seen dedupe in column order - __match_any_sync finds equal
//     pairs inside a 32-column chunk (lowest lane wins), the bitmap answers "seen in an
//     earlier chunk"; survivors are ranked with ballot/popc and write four tokens into
//     the warp's staging area (16-bit tokens when the vocabulary allows); four rows are
//     staged back to back and leave as ONE contiguous run of aligned 16-byte streaming
//     stores.  Columns are re-read from L2 (a tile's inputs are ~180 KB, read ~20 us
//     earlier), so HBM still sees each input byte once;
//   * columns go straight into registers, lane i <-> column i0+i (coalesced runs), and
//     the next molecule's loads are issued before the current one is processed.
// HBM traffic = algorithmic minimum: 12*Ed + 4N + 16 read, 8L + 12 written per molecule.
// Molecules only (max_nodes <= 64, <= 256 columns); other shapes use gtok_zinc.cu.
#include <stdlib.h>

#include "gtok_common.cuh"

namespace gtok {

constexpr int kWT = 4;                  // rows staged back to back per flush
constexpr int kFThreads = 256;
constexpr int kFWarps = kFThreads / 32;
constexpr int kFTile = 32 * kFWarps;    // 256 molecules per CTA tile, 32 per warp (lane j <-> molecule j)
constexpr int kFSmallN = 64;
constexpr int kFMinBlocks = 5;

struct ZincFusedLayout {
  int cap_row;        // staged positions per row
  int warp_stride;    // bytes per warp: bitmap + stage
};

template <typename StageT>
static ZincFusedLayout zinc_fused_layout(int max_nodes, int max_entries, int item_max, int ds_max) {
  ZincFusedLayout L;
  long long bound = 2ll * max_nodes + 4ll * max_entries + 4;
  const long long lim = (long long)(item_max > ds_max ? item_max : ds_max) + 1;
  if (bound > lim) bound = lim;
  L.cap_row = (int)bound;
  const int stage_bytes = (int)((kWT * bound * sizeof(StageT) + 15) & ~15ll);
  L.warp_stride = 4 * 2 * kFSmallN + stage_bytes;
  return L;
}

struct ZincTables {
  int atom[10];   // 0..8 = C..I, 9 = 'X'
  int bond[5];    // 0..3 = single..aromatic, 4 = 'unknown'
};

// One 32-column chunk of the first-seen dedupe + token staging.  `guard` is warp-uniform:
// false when the whole untruncated row fits the staging capacity (the common case).
template <typename StageT>
__device__ __forceinline__ int zinc_chunk(bool valid, int a, int c, int t, unsigned* rows, StageT* st, int pbase,
                                          int cap, bool guard, unsigned lt, int lane, const int* s_nt,
                                          const ZincTables& tb, int tok_bond, int oov) {
  const unsigned mn = (unsigned)min(a, c) & 63u, mx = (unsigned)max(a, c) & 63u;
  const unsigned key = valid ? (mn << 6 | mx) : (0x10000u | (unsigned)lane);
  const unsigned grp = __match_any_sync(kFull, key);
  const unsigned word = 2 * mn + (mx >> 5), bit = 1u << (mx & 31);
  // leaders of different groups touch different bits, so no barrier is needed between this read and the OR
  const bool keep = valid && ((grp & lt) == 0u) && ((rows[word] & bit) == 0u);
  const unsigned m = __ballot_sync(kFull, keep);
  if (keep) {
    atomicOr(&rows[word], bit);
    const int p = pbase + 4 * __popc(m & lt);
    const int k1 = tb.bond[min((unsigned)(t - 1), 4u)];
    const int k2 = (unsigned)a < (unsigned)kFSmallN ? s_nt[a] : oov;
    const int k3 = (unsigned)c < (unsigned)kFSmallN ? s_nt[c] : oov;
    if (!guard || p + 3 < cap) {
      st[p] = (StageT)tok_bond; st[p + 1] = (StageT)k1; st[p + 2] = (StageT)k2; st[p + 3] = (StageT)k3;
    } else {
      if (p < cap) st[p] = (StageT)tok_bond;
      if (p + 1 < cap) st[p + 1] = (StageT)k1;
      if (p + 2 < cap) st[p + 2] = (StageT)k2;
    }
  }
  return __popc(m);
}

template <typename StageT, bool HAS_ATTR>
__global__ void __launch_bounds__(kFThreads, kFMinBlocks)
zinc_fused_kernel(const GtokGraphBatch b, const __grid_constant__ GtokZincVocab v, const ZincFusedLayout lay,
                  int32_t* __restrict__ len_out, int64_t* __restrict__ off_out, int64_t* __restrict__ out,
                  int64_t out_cap, unsigned int* __restrict__ ticket, unsigned long long* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char f_smem[];
  __shared__ int s_nt[kFSmallN];     // node index -> token id
  __shared__ ZincTables s_tb;
  __shared__ int s_wtot[kFWarps];
  __shared__ long long s_base;
  __shared__ volatile unsigned int s_ready;   // tile+1 once s_base is valid
  __shared__ unsigned int s_tile;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  unsigned* rows = reinterpret_cast<unsigned*>(f_smem + (size_t)w * lay.warp_stride);
  StageT* stage = reinterpret_cast<StageT*>(rows + 2 * kFSmallN);
  const int cap = lay.cap_row;
  const unsigned lt = (1u << lane) - 1u;
  const int64_t G = b.num_graphs;
  const int item_max = v.item_max_len, ds_max = v.ds_max_len;
  const int tok_bond = v.tok_bond, tok_atom = v.tok_atom, oov = v.oov;
  const int tail_tok = lane == 0 ? v.tok_q : (lane == 1 ? v.tok_regression : v.tok_p);
  const int32_t* __restrict__ g_src = b.src;
  const int32_t* __restrict__ g_dst = b.dst;
  const int32_t* __restrict__ g_attr = b.edge_attr;
  const int32_t* __restrict__ g_x = b.x;

  if (tid < kFSmallN) s_nt[tid] = tid < v.node_tok_len ? __ldg(v.node_tok + tid) : v.oov;
  if (tid < 9) s_tb.atom[tid] = v.atom_tok[tid];
  if (tid == 9) s_tb.atom[9] = v.atom_oov;
  if (tid >= 16 && tid < 20) s_tb.bond[tid - 16] = v.bond_tok[tid - 16];
  if (tid == 20) s_tb.bond[4] = v.bond_oov;
  if (tid == 0) s_ready = 0u;

  while (true) {
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();                                   // (1) tile id visible; previous tile fully drained
    const long long tile = s_tile;
    const int64_t t0g = tile * kFTile;
    if (t0g >= G) break;
    const int64_t g0 = t0g + (int64_t)w * 32;          // this warp's first molecule
    const int Tc = (int)(g0 >= G ? 0 : ((G - g0 < 32) ? (G - g0) : 32));
    // ---- header: lane j holds molecule j's column range and atom range (relative to the warp's first) ----
    long long e_lo = 0, e_hi = 0, n_lo = 0, n_hi = 0;
    if (lane < Tc) {
      e_lo = __ldg(b.edge_off + g0 + lane); e_hi = __ldg(b.edge_off + g0 + lane + 1);
      n_lo = __ldg(b.node_off + g0 + lane); n_hi = __ldg(b.node_off + g0 + lane + 1);
    }
    const long long E0 = __shfl_sync(kFull, e_lo, 0), N0 = __shfl_sync(kFull, n_lo, 0);
    const int my_eo = (int)(e_lo - E0), my_Ed = (int)(e_hi - e_lo);
    const int my_no = (int)(n_lo - N0), my_N = (int)(n_hi - n_lo);
    const int32_t* __restrict__ ps = g_src + E0;
    const int32_t* __restrict__ pd = g_dst + E0;
    const int32_t* __restrict__ pa = HAS_ATTR ? g_attr + E0 : nullptr;
    const int32_t* __restrict__ px = g_x + N0;

    // ---- phase 1: distinct-pair counts -> row lengths (lane j keeps molecule j's) ----
    int my_len = 0;
    {
      int eo = __shfl_sync(kFull, my_eo, 0), Ed = __shfl_sync(kFull, my_Ed, 0);
      int a0 = 0, c0 = 0, a1 = 0, c1 = 0;
      if (lane < Ed) { a0 = __ldg(ps + eo + lane); c0 = __ldg(pd + eo + lane); }
      if (lane + 32 < Ed) { a1 = __ldg(ps + eo + lane + 32); c1 = __ldg(pd + eo + lane + 32); }
#pragma unroll 1
      for (int j = 0; j < Tc; ++j) {
        const int jn = min(j + 1, 31);
        const int neo = __shfl_sync(kFull, my_eo, jn);
        const int nEd = (j + 1 < Tc) ? __shfl_sync(kFull, my_Ed, jn) : 0;
        int na0 = 0, nc0 = 0, na1 = 0, nc1 = 0;
        if (lane < nEd) { na0 = __ldg(ps + neo + lane); nc0 = __ldg(pd + neo + lane); }
        if (lane + 32 < nEd) { na1 = __ldg(ps + neo + lane + 32); nc1 = __ldg(pd + neo + lane + 32); }
        reinterpret_cast<uint4*>(rows)[lane] = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        if (lane < Ed) {
          const unsigned mn = (unsigned)min(a0, c0) & 63u, mx = (unsigned)max(a0, c0) & 63u;
          atomicOr(&rows[2 * mn + (mx >> 5)], 1u << (mx & 31));
        }
        if (lane + 32 < Ed) {
          const unsigned mn = (unsigned)min(a1, c1) & 63u, mx = (unsigned)max(a1, c1) & 63u;
          atomicOr(&rows[2 * mn + (mx >> 5)], 1u << (mx & 31));
        }
#pragma unroll 1
        for (int i = 64 + lane; i < Ed; i += kWarp) {
          const int a = __ldg(ps + eo + i), c = __ldg(pd + eo + i);
          const unsigned mn = (unsigned)min(a, c) & 63u, mx = (unsigned)max(a, c) & 63u;
          atomicOr(&rows[2 * mn + (mx >> 5)], 1u << (mx & 31));
        }
        __syncwarp();
        const uint4 q = reinterpret_cast<const uint4*>(rows)[lane];
        int cnt = __popc(q.x) + __popc(q.y) + __popc(q.z) + __popc(q.w);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(kFull, cnt, d);
        if (lane == j) {
          const int T = 2 * my_N + 4 * cnt + 6;        // full text length incl. label + <eos> (oracle: zinc_length)
          const int L = T <= item_max ? T - 2 : (T == item_max + 1 ? item_max - 1 : item_max);
          my_len = max(0, min(L, ds_max));
        }
        eo = neo; Ed = nEd; a0 = na0; c0 = nc0; a1 = na1; c1 = nc1;
        __syncwarp();
      }
    }
    // warp-local exclusive offsets + warp total
    int incl = my_len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(kFull, incl, d);
      if (lane >= d) incl += o;
    }
    const int wtot = __shfl_sync(kFull, incl, 31);
    if (lane == 0) s_wtot[w] = wtot;
    __syncthreads();                                   // (2) all warp totals known
    int wexcl = 0;
    long long ttotal = 0;
#pragma unroll
    for (int k = 0; k < kFWarps; ++k) {
      const int tk = s_wtot[k];
      if (k < w) wexcl += tk;
      ttotal += tk;
    }
    if (w == 0) {
      // publish the tile aggregate NOW, resolve the global offset while the other warps build rows
      const long long tbase = lookback_exclusive(status, tile, ttotal);
      if (lane == 0) {
        s_base = tbase;
        if (t0g + kFTile >= G) off_out[G] = tbase + ttotal;
        __threadfence_block();
        s_ready = (unsigned)tile + 1u;
      }
    }
    const int my_off = wexcl + incl - my_len;          // tile-local offset of this lane's row

    // ---- phase 2: rows, kWT at a time -----------------------------------------------
    long long base = -1;
    int eo = __shfl_sync(kFull, my_eo, 0), Ed = __shfl_sync(kFull, my_Ed, 0);
    int no = 0, N = __shfl_sync(kFull, my_N, 0);
    int a0 = 0, a1 = 0, c0 = 0, c1 = 0, t0 = 0, t1 = 0, x0 = 0, x1 = 0;
    if (Tc > 0) {
      if (lane < Ed) { a0 = __ldcs(ps + lane); c0 = __ldcs(pd + lane); if (HAS_ATTR) t0 = __ldcs(pa + lane); }
      if (lane + 32 < Ed) { a1 = __ldcs(ps + lane + 32); c1 = __ldcs(pd + lane + 32); if (HAS_ATTR) t1 = __ldcs(pa + lane + 32); }
      if (lane < N) x0 = __ldcs(px + lane);
      if (lane + 32 < N) x1 = __ldcs(px + lane + 32);
    }
#pragma unroll 1
    for (int j0 = 0; j0 < Tc; j0 += kWT) {
      const int jend = min(j0 + kWT, Tc);
      int loc = 0;
#pragma unroll 1
      for (int j = j0; j < jend; ++j) {
        // prefetch molecule j+1
        const int jn = min(j + 1, 31);
        const int neo = __shfl_sync(kFull, my_eo, jn);
        const int nEd = (j + 1 < Tc) ? __shfl_sync(kFull, my_Ed, jn) : 0;
        const int nno = __shfl_sync(kFull, my_no, jn);
        const int nN = (j + 1 < Tc) ? __shfl_sync(kFull, my_N, jn) : 0;
        int na0 = 0, na1 = 0, nc0 = 0, nc1 = 0, nt0 = 0, nt1 = 0, nx0 = 0, nx1 = 0;
        if (lane < nEd) { na0 = __ldcs(ps + neo + lane); nc0 = __ldcs(pd + neo + lane); if (HAS_ATTR) nt0 = __ldcs(pa + neo + lane); }
        if (lane + 32 < nEd) { na1 = __ldcs(ps + neo + lane + 32); nc1 = __ldcs(pd + neo + lane + 32); if (HAS_ATTR) nt1 = __ldcs(pa + neo + lane + 32); }
        if (lane < nN) nx0 = __ldcs(px + nno + lane);
        if (lane + 32 < nN) nx1 = __ldcs(px + nno + lane + 32);
        // stage row j at stage[loc ..]
        StageT* st = stage + loc;
        const int bbase = 1 + 2 * N;
        const bool guard = bbase + 4 * Ed + 3 > cap;   // warp-uniform: can any staged position reach the capacity?
        reinterpret_cast<uint4*>(rows)[lane] = make_uint4(0u, 0u, 0u, 0u);
        if (lane == 0) st[0] = (StageT)v.tok_bos;
        if (lane < N) {
          const int p = 1 + 2 * lane;
          if (!guard || p < cap) st[p] = (StageT)tok_atom;
          if (!guard || p + 1 < cap) st[p + 1] = (StageT)s_tb.atom[min((unsigned)x0, 9u)];
        }
        if (lane + 32 < N) {
          const int p = 65 + 2 * lane;
          if (!guard || p < cap) st[p] = (StageT)tok_atom;
          if (!guard || p + 1 < cap) st[p + 1] = (StageT)s_tb.atom[min((unsigned)x1, 9u)];
        }
        __syncwarp();                                  // bitmap cleared before the first chunk reads it
        int Eu = 0;
        if (Ed > 0) Eu += zinc_chunk<StageT>(lane < Ed, a0, c0, t0, rows, st, bbase, cap, guard, lt, lane, s_nt, s_tb, tok_bond, oov);
        if (Ed > 32) {
          __syncwarp();                                // chunk 0's bits visible to chunk 1
          Eu += zinc_chunk<StageT>(lane + 32 < Ed, a1, c1, t1, rows, st, bbase + 4 * Eu, cap, guard, lt, lane, s_nt, s_tb, tok_bond, oov);
#pragma unroll 1
          for (int i0 = 64; i0 < Ed; i0 += kWarp) {     // rare: more than 64 columns
            const bool valid = i0 + lane < Ed;
            int a = 0, c = 0, t = 0;
            if (valid) {
              a = __ldcs(ps + eo + i0 + lane);
              c = __ldcs(pd + eo + i0 + lane);
              if (HAS_ATTR) t = __ldcs(pa + eo + i0 + lane);
            }
            __syncwarp();
            Eu += zinc_chunk<StageT>(valid, a, c, t, rows, st, bbase + 4 * Eu, cap, guard, lt, lane, s_nt, s_tb, tok_bond, oov);
          }
        }
        const int T = 2 * N + 4 * Eu + 6;
        const int tb = bbase + 4 * Eu;
        if (lane < 3 && (!guard || tb + lane < cap)) st[tb + lane] = (StageT)tail_tok;
        if (T >= item_max + 2) {                       // truncated text keeps a trailing <eos> (zinc_dataset_indexbase.py:216-221)
          __syncwarp();
          if (lane == 0 && item_max - 1 < cap) st[item_max - 1] = (StageT)v.tok_eos;
        }
        loc += __shfl_sync(kFull, my_len, j);          // == the length phase 1 derived for this molecule
        eo = neo; Ed = nEd; no = nno; N = nN;
        a0 = na0; a1 = na1; c0 = nc0; c1 = nc1; t0 = nt0; t1 = nt1; x0 = nx0; x1 = nx1;
        __syncwarp();                                  // row j complete before row j+1 may overwrite its slack
      }
      // ---- flush: one contiguous run of `loc` tokens ------------------------------------
      if (base < 0) {
        while (s_ready != (unsigned)tile + 1u) __nanosleep(40);
        __threadfence_block();
        base = s_base;
      }
      const long long run0 = base + __shfl_sync(kFull, my_off, j0);
      if (run0 + loc <= out_cap) {
        int64_t* __restrict__ dst = out + run0;
        const int head = (int)(run0 & 1);
        if (head && lane == 0 && loc > 0) st_stream_i64(dst, (int64_t)stage[0]);
        const int npair = (loc - head) >> 1;
#pragma unroll 1
        for (int q = lane; q < npair; q += kWarp) {
          const int p = head + 2 * q;
          st_stream_i64x2(dst + p, (int64_t)stage[p], (int64_t)stage[p + 1]);
        }
        const int tail = head + 2 * npair;
        if (tail < loc && lane == 31) st_stream_i64(dst + tail, (int64_t)stage[tail]);
      }
      __syncwarp();                                    // stage drained before the next group overwrites it
    }
    if (Tc > 0) {
      if (base < 0) {
        while (s_ready != (unsigned)tile + 1u) __nanosleep(40);
        __threadfence_block();
        base = s_base;
      }
      if (lane < Tc) {
        len_out[g0 + lane] = my_len;
        off_out[g0 + lane] = base + my_off;
      }
    }
  }
}

template <typename StageT>
static int zinc_fused_launch(const GtokGraphBatch* batch, const GtokZincVocab* vocab, int32_t* len_out,
                             int64_t* off_out, int64_t* tok_out, int64_t tok_capacity, void* ws, size_t need,
                             cudaStream_t st) {
  const ZincFusedLayout lay = zinc_fused_layout<StageT>(batch->max_nodes, batch->max_entries, vocab->item_max_len, vocab->ds_max_len);
  const size_t smem = (size_t)kFWarps * lay.warp_stride;
  int dev = 0, max_optin = 0;
  GTOK_CUDA_TRY(cudaGetDevice(&dev));
  GTOK_CUDA_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  GTOK_REQUIRE(smem <= (size_t)max_optin, GTOK_EUNSUPPORTED, "zinc_tokenize: needs %zu B shared memory > %d", smem, max_optin);
  auto k = batch->edge_attr ? zinc_fused_kernel<StageT, true> : zinc_fused_kernel<StageT, false>;
  GTOK_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GTOK_CUDA_TRY(cudaMemsetAsync(ws, 0, need, st));
  const int64_t tiles = (batch->num_graphs + kFTile - 1) / kFTile;
  const int grid = persistent_grid(k, kFThreads, smem, tiles);
  k<<<grid, kFThreads, smem, st>>>(*batch, *vocab, lay, len_out, off_out, tok_out, tok_capacity,
                                   reinterpret_cast<unsigned int*>(ws),
                                   reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ws) + 16));
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

// gtok_zinc_tpm.cu
int zinc_tpm_ragged(const GtokGraphBatch* batch, const GtokZincVocab* vocab, int32_t* len_out, int64_t* off_out,
                    int64_t* tok_out, int64_t tok_capacity, void* ws, size_t ws_bytes, cudaStream_t st, bool* handled);

}  // namespace gtok

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
  GTOK_REQUIRE(batch->node_off && batch->edge_off && batch->src && batch->dst && batch->x, GTOK_EINVAL,
               "zinc_tokenize: null graph arrays (x is required)");
  GTOK_REQUIRE(len_out && tok_out && ws, GTOK_EINVAL, "zinc_tokenize: null output/workspace");
  GTOK_REQUIRE(tok_capacity >= 0, GTOK_EINVAL, "zinc_tokenize: negative capacity");
  GTOK_REQUIRE(vocab->item_max_len >= 2 && vocab->ds_max_len >= 0, GTOK_EINVAL, "zinc_tokenize: max_len must be >= 2");
  GTOK_REQUIRE(vocab->node_tok_len == 0 || vocab->node_tok != nullptr, GTOK_EINVAL, "zinc_tokenize: null node_tok table");
  GTOK_REQUIRE(batch->max_nodes > 0 && batch->max_entries > 0, GTOK_EINVAL, "zinc_tokenize: max_nodes/max_entries bounds required");
  GTOK_REQUIRE(batch->max_nodes <= kFSmallN && batch->max_entries <= 256, GTOK_EUNSUPPORTED,
               "zinc_tokenize: fused path covers molecules with <= 64 atoms and <= 256 edge_index columns "
               "(got bounds %d, %d); use gtok_zinc_lengths + gtok_scan_offsets + gtok_zinc_serialize",
               batch->max_nodes, batch->max_entries);
  GTOK_REQUIRE((reinterpret_cast<uintptr_t>(tok_out) & 15) == 0 && (reinterpret_cast<uintptr_t>(ws) & 15) == 0,
               GTOK_EINVAL, "zinc_tokenize: tok_out and workspace must be 16-byte aligned");
  const size_t need = gtok_zinc_tokenize_workspace_bytes(batch->num_graphs);
  GTOK_REQUIRE(ws_bytes >= need, GTOK_EWORKSPACE, "zinc_tokenize: workspace %zu < %zu bytes", ws_bytes, need);
  {
    // v6 thread-per-molecule kernel first; GTOK_ZINC_IMPL=warp keeps the warp-per-molecule kernel (A/B runs)
    static const bool force_warp = [] { const char* e = getenv("GTOK_ZINC_IMPL"); return e && e[0] == 'w'; }();
    if (!force_warp) {
      bool handled = false;
      const int rc = zinc_tpm_ragged(batch, vocab, len_out, off_out, tok_out, tok_capacity, ws, ws_bytes, st, &handled);
      if (rc != GTOK_OK || handled) return rc;
    }
  }
  if (vocab->max_token_id > 0 && vocab->max_token_id < 65536)
    return zinc_fused_launch<uint16_t>(batch, vocab, len_out, off_out, tok_out, tok_capacity, ws, need, st);
  return zinc_fused_launch<int32_t>(batch, vocab, len_out, off_out, tok_out, tok_capacity, ws, need, st);
}
