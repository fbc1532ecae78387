// gtok_zinc_fused.cu - K2 single pass: molecules -> (lengths, ragged offsets, int64 rows).
//
// Same reference functions as gtok_zinc.cu (zinc_dataset_indexbase.py:143-227 +
// data_loader.py:465-486), fused with K0 so every input byte is read from HBM once
// and nothing but the final rows / lengths / offsets is written.
//
// Design (history of v1..v3 and their profiles in DESIGN.md / profiles/):
//   * a CTA (8 warps) takes the next tile of 64 CONSECUTIVE molecules from an atomic
//     ticket (so the look-back below can never wait on a tile that has not started);
//     warp w owns molecules 8w..8w+7 of the tile and works on them WITHOUT any CTA
//     barrier: header offsets and columns are loaded straight into registers, lane i
//     <-> column i0+i (coalesced runs), the next molecule's loads are issued before the
//     current one is processed;
//   * first-seen dedupe in column order: __match_any_sync finds equal unordered pairs
//     inside a 32-column chunk (lowest lane wins), a 64x64 adjacency bitmap in shared
//     memory answers "seen in an earlier chunk"; survivors are ranked with
//     ballot/popc and write their four tokens into the warp's staging area;
//   * the warp's eight rows are staged BACK TO BACK (16-bit tokens when the vocabulary
//     allows); lengths fall out of the single dedupe pass (no counting pre-pass);
//   * two CTA barriers per tile bracket the only serial step: warp 0 scans the eight
//     warp totals and gets the tile's global token offset by decoupled look-back (one
//     look-back per 64 molecules keeps the chain ~15 tiles/us; per-warp look-back was
//     tried and spent 27% of all instructions spinning, profiles/r1d_*);
//   * each warp then streams its contiguous run with aligned 16-byte stores.
// HBM traffic = algorithmic minimum: 12*Ed + 4N + 16 read, 8L + 12 written per molecule.
// Molecules only (max_nodes <= 64, <= 256 columns); other shapes use gtok_zinc.cu.
#include "gtok_common.cuh"

namespace gtok {

constexpr int kWT = 8;            // molecules per warp per tile
constexpr int kFThreads = 256;
constexpr int kFWarps = kFThreads / 32;
constexpr int kFTile = kWT * kFWarps;   // 64 molecules per CTA tile
constexpr int kFSmallN = 64;

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

template <typename StageT>
__global__ void __launch_bounds__(kFThreads)
zinc_fused_kernel(const GtokGraphBatch b, const __grid_constant__ GtokZincVocab v, const ZincFusedLayout lay,
                  int32_t* __restrict__ len_out, int64_t* __restrict__ off_out, int64_t* __restrict__ out,
                  int64_t out_cap, unsigned int* __restrict__ ticket, unsigned long long* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char f_smem[];
  __shared__ int s_nt[kFSmallN];     // node index -> token id
  __shared__ ZincTables s_tb;
  __shared__ int s_wtot[kFWarps];
  __shared__ long long s_wbase[kFWarps];
  __shared__ unsigned int s_tile;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  unsigned* rows = reinterpret_cast<unsigned*>(f_smem + (size_t)w * lay.warp_stride);
  StageT* stage = reinterpret_cast<StageT*>(rows + 2 * kFSmallN);
  const int cap = lay.cap_row;
  const unsigned lt = (1u << lane) - 1u;
  const int64_t G = b.num_graphs;
  const int item_max = v.item_max_len, ds_max = v.ds_max_len;
  const int tok_bond = v.tok_bond, tok_atom = v.tok_atom, oov = v.oov;
  const int32_t* __restrict__ g_src = b.src;
  const int32_t* __restrict__ g_dst = b.dst;
  const int32_t* __restrict__ g_attr = b.edge_attr;
  const int32_t* __restrict__ g_x = b.x;

  if (tid < kFSmallN) s_nt[tid] = tid < v.node_tok_len ? __ldg(v.node_tok + tid) : v.oov;
  if (tid < 9) s_tb.atom[tid] = v.atom_tok[tid];
  if (tid == 9) s_tb.atom[9] = v.atom_oov;
  if (tid >= 16 && tid < 20) s_tb.bond[tid - 16] = v.bond_tok[tid - 16];
  if (tid == 20) s_tb.bond[4] = v.bond_oov;

  while (true) {
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();                                   // (1) tile id visible; previous tile fully drained
    const long long tile = s_tile;
    const int64_t t0g = tile * kFTile;
    if (t0g >= G) break;
    const int64_t g0 = t0g + (int64_t)w * kWT;         // this warp's first molecule
    const int Tc = (int)(g0 >= G ? 0 : ((G - g0 < kWT) ? (G - g0) : kWT));
    // header: lanes 0..8 edge_off[g0..g0+8], lanes 16..24 node_off[g0..g0+8] (clamped to the warp's range)
    long long hv = 0;
    if (Tc > 0) {
      if (lane <= kWT) hv = __ldg(b.edge_off + g0 + (lane < Tc ? lane : Tc));
      else if (lane >= 16 && lane <= 16 + kWT) hv = __ldg(b.node_off + g0 + (lane - 16 < Tc ? lane - 16 : Tc));
    }
    const long long E0 = __shfl_sync(kFull, hv, 0), N0 = __shfl_sync(kFull, hv, 16);
    const int rel = (int)(hv - (lane < 16 ? E0 : N0));  // lane k: entry offset of molecule k; lane 16+k: node offset

    int eo = 0, Ed = __shfl_sync(kFull, rel, 1);
    int N = __shfl_sync(kFull, rel, 17);
    int a0 = 0, a1 = 0, c0 = 0, c1 = 0, t0 = 0, t1 = 0, x0 = 0, x1 = 0;
    if (Tc > 0) {
      const int32_t* ps = g_src + E0; const int32_t* pd = g_dst + E0; const int32_t* px = g_x + N0;
      if (lane < Ed) { a0 = __ldcs(ps + lane); c0 = __ldcs(pd + lane); if (g_attr) t0 = __ldcs(g_attr + E0 + lane); }
      if (lane + 32 < Ed) { a1 = __ldcs(ps + lane + 32); c1 = __ldcs(pd + lane + 32); if (g_attr) t1 = __ldcs(g_attr + E0 + lane + 32); }
      if (lane < N) x0 = __ldcs(px + lane);
      if (lane + 32 < N) x1 = __ldcs(px + lane + 32);
    }
    int loc = 0;          // tokens staged so far by this warp
    int my_len = 0, my_off = 0;
#pragma unroll 1
    for (int j = 0; j < Tc; ++j) {
      // ---- prefetch molecule j+1 ------------------------------------------------
      int neo = 0, nEd = 0, nN = 0;
      int na0 = 0, na1 = 0, nc0 = 0, nc1 = 0, nt0 = 0, nt1 = 0, nx0 = 0, nx1 = 0;
      if (j + 1 < Tc) {
        neo = __shfl_sync(kFull, rel, j + 1);
        nEd = __shfl_sync(kFull, rel, j + 2) - neo;
        const int nno = __shfl_sync(kFull, rel, 16 + j + 1);
        nN = __shfl_sync(kFull, rel, 16 + j + 2) - nno;
        const int32_t* ps = g_src + E0 + neo; const int32_t* pd = g_dst + E0 + neo; const int32_t* px = g_x + N0 + nno;
        if (lane < nEd) { na0 = __ldcs(ps + lane); nc0 = __ldcs(pd + lane); if (g_attr) nt0 = __ldcs(g_attr + E0 + neo + lane); }
        if (lane + 32 < nEd) { na1 = __ldcs(ps + lane + 32); nc1 = __ldcs(pd + lane + 32); if (g_attr) nt1 = __ldcs(g_attr + E0 + neo + lane + 32); }
        if (lane < nN) nx0 = __ldcs(px + lane);
        if (lane + 32 < nN) nx1 = __ldcs(px + lane + 32);
      }
      // ---- stage row j at stage[loc ..] ---------------------------------------------
      StageT* st = stage + loc;
      const int bbase = 1 + 2 * N;
      const bool guard = bbase + 4 * Ed + 3 > cap;     // warp-uniform: can any staged position reach the capacity?
      reinterpret_cast<uint4*>(rows)[lane] = make_uint4(0u, 0u, 0u, 0u);   // 32 x 16 B = the whole 64x64 bitmap
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
      __syncwarp();                                    // bitmap cleared before the first chunk reads it
      int Eu = 0;
      if (Ed > 0) Eu += zinc_chunk<StageT>(lane < Ed, a0, c0, t0, rows, st, bbase, cap, guard, lt, lane, s_nt, s_tb, tok_bond, oov);
      if (Ed > 32) {
        __syncwarp();                                  // chunk 0's bits visible to chunk 1
        Eu += zinc_chunk<StageT>(lane + 32 < Ed, a1, c1, t1, rows, st, bbase + 4 * Eu, cap, guard, lt, lane, s_nt, s_tb, tok_bond, oov);
#pragma unroll 1
        for (int i0 = 64; i0 < Ed; i0 += kWarp) {       // rare: more than 64 columns
          const bool valid = i0 + lane < Ed;
          int a = 0, c = 0, t = 0;
          if (valid) {
            a = __ldcs(g_src + E0 + eo + i0 + lane);
            c = __ldcs(g_dst + E0 + eo + i0 + lane);
            if (g_attr) t = __ldcs(g_attr + E0 + eo + i0 + lane);
          }
          __syncwarp();
          Eu += zinc_chunk<StageT>(valid, a, c, t, rows, st, bbase + 4 * Eu, cap, guard, lt, lane, s_nt, s_tb, tok_bond, oov);
        }
      }
      const int T = 2 * N + 4 * Eu + 6;                // full text length incl. label + <eos> (oracle: zinc_length)
      int L = T <= item_max ? T - 2 : (T == item_max + 1 ? item_max - 1 : item_max);
      L = max(0, min(L, ds_max));
      const int tb = bbase + 4 * Eu;
      if (lane < 3 && (!guard || tb + lane < cap))
        st[tb + lane] = (StageT)(lane == 0 ? v.tok_q : (lane == 1 ? v.tok_regression : v.tok_p));
      if (T >= item_max + 2) {                         // truncated text keeps a trailing <eos> (zinc_dataset_indexbase.py:216-221)
        __syncwarp();
        if (lane == 0 && item_max - 1 < cap) st[item_max - 1] = (StageT)v.tok_eos;
      }
      if (lane == j) { my_len = L; my_off = loc; }
      loc += L;
      // rotate the prefetched molecule in
      eo = neo; Ed = nEd; N = nN;
      a0 = na0; a1 = na1; c0 = nc0; c1 = nc1; t0 = nt0; t1 = nt1; x0 = nx0; x1 = nx1;
      __syncwarp();                                    // row j complete before row j+1 may overwrite its slack
    }
    // ---- warp totals -> tile total -> global offset (one decoupled look-back per CTA tile) ----
    if (lane == 0) s_wtot[w] = loc;
    __syncthreads();                                   // (2)
    if (w == 0) {
      const int mine = lane < kFWarps ? s_wtot[lane] : 0;
      int incl = mine;
#pragma unroll
      for (int d = 1; d < kFWarps; d <<= 1) {
        const int o = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += o;
      }
      const long long total = __shfl_sync(kFull, incl, kFWarps - 1);
      const long long tbase = lookback_exclusive(status, tile, total);
      if (lane < kFWarps) s_wbase[lane] = tbase + (incl - mine);
      if (lane == 0 && t0g + kFTile >= G) off_out[G] = tbase + total;
    }
    __syncthreads();                                   // (3)
    const long long base = s_wbase[w];
    if (lane < Tc) {
      len_out[g0 + lane] = my_len;
      off_out[g0 + lane] = base + my_off;
    }
    // ---- one contiguous run of `loc` tokens per warp, aligned 16-byte streaming stores ----
    if (base + loc <= out_cap) {
      int64_t* __restrict__ dst = out + base;
      const int head = (int)(base & 1);
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
  auto k = zinc_fused_kernel<StageT>;
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

}  // namespace gtok

extern "C" size_t gtok_zinc_tokenize_workspace_bytes(int64_t num_graphs) {
  const int64_t tiles = (num_graphs + gtok::kFTile - 1) / gtok::kFTile;
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
  if (vocab->max_token_id > 0 && vocab->max_token_id < 65536)
    return zinc_fused_launch<uint16_t>(batch, vocab, len_out, off_out, tok_out, tok_capacity, ws, need, st);
  return zinc_fused_launch<int32_t>(batch, vocab, len_out, off_out, tok_out, tok_capacity, ws, need, st);
}
