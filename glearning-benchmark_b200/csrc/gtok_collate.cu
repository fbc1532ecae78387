// gtok_collate.cu - K4: ragged token rows -> padded [B, Lmax] int64 batch + bool mask.
//
// Replaces (reference, Python row-copy loops run once per DataLoader step):
//   collate      graph_data_loader/data_loader.py:488-497   (IBTT; pad = vocab["<pad>"])
//   collate_fn   trainer/train_agtt.py:276-302              (AGTT; pad = Graph2TrailTokenizer.pad = 5)
//   X = full((B, Lmax), pad, long); attn = zeros(bool); X[i,:len]=row; attn[i,:len]=True
// Lmax is the longest row OF THAT BATCH (not max_len), which the batched entry
// point reproduces per batch.
//
// One warp per output row: 16-byte streaming stores for X (2 tokens / lane) and
// for the mask (16 flags / lane); source rows are read with 8-byte coalesced loads
// (source and destination 16-byte phases differ, so only the store side is vectorised).
// Algorithmic bytes: 8*sum(len) read + 9*B*Lmax written (+ 12 B/row of row metadata).
#include "gtok_common.cuh"

namespace gtok {

constexpr int kColThreads = 256;
constexpr int kColWarps = kColThreads / 32;

__device__ __forceinline__ unsigned long long mask_bytes(int nvalid) {
  // low `nvalid` bytes (0..8) = 0x01
  if (nvalid <= 0) return 0ull;
  if (nvalid >= 8) return 0x0101010101010101ull;
  return 0x0101010101010101ull >> (8 * (8 - nvalid));
}

__device__ __forceinline__ void warp_collate_row(const int64_t* __restrict__ src, int len, int Lmax, int64_t pad,
                                                 int64_t* __restrict__ xrow, uint8_t* __restrict__ arow) {
  const int lane = lane_id();
  // ---- X row: [0,len) copied, [len,Lmax) = pad; 16-byte stores on the aligned interior.
  const int head = (int)((reinterpret_cast<uintptr_t>(xrow) >> 3) & 1);
  if (head && lane == 0 && Lmax > 0) st_stream_i64(xrow, 0 < len ? __ldcs(src) : pad);
  const int npair = (Lmax - head) >> 1;
  int j = lane;
  // two pairs per lane per trip: four independent 8-byte loads in flight before the stores
  for (; j + kWarp < npair; j += 2 * kWarp) {
    const int p0 = head + 2 * j, p1 = p0 + 2 * kWarp;
    const int64_t a0 = p0 < len ? __ldcs(src + p0) : pad;
    const int64_t b0 = p0 + 1 < len ? __ldcs(src + p0 + 1) : pad;
    const int64_t a1 = p1 < len ? __ldcs(src + p1) : pad;
    const int64_t b1 = p1 + 1 < len ? __ldcs(src + p1 + 1) : pad;
    st_stream_i64x2(xrow + p0, a0, b0);
    st_stream_i64x2(xrow + p1, a1, b1);
  }
  if (j < npair) {
    const int p0 = head + 2 * j;
    const int64_t a0 = p0 < len ? __ldcs(src + p0) : pad;
    const int64_t b0 = p0 + 1 < len ? __ldcs(src + p0 + 1) : pad;
    st_stream_i64x2(xrow + p0, a0, b0);
  }
  const int tailx = head + 2 * npair;
  if (tailx < Lmax && lane == 31) st_stream_i64(xrow + tailx, tailx < len ? __ldcs(src + tailx) : pad);
  // ---- mask row: scalar head up to 16-byte alignment, 16-byte interior, scalar tail
  int mh = (int)((16 - (reinterpret_cast<uintptr_t>(arow) & 15)) & 15);
  if (mh > Lmax) mh = Lmax;
  if (lane < mh) arow[lane] = (uint8_t)(lane < len);
  const int nvec = (Lmax - mh) >> 4;
  for (int q = lane; q < nvec; q += kWarp) {
    const int p = mh + 16 * q;
    const int nv = len - p;
    __stcs(reinterpret_cast<ulonglong2*>(arow + p), make_ulonglong2(mask_bytes(nv), mask_bytes(nv - 8)));
  }
  const int tail = mh + 16 * nvec;
  if (tail + lane < Lmax && lane < 16) arow[tail + lane] = (uint8_t)(tail + lane < len);
}

__global__ void __launch_bounds__(kColThreads)
pad_mask_kernel(const int64_t* __restrict__ tok, const int64_t* __restrict__ row_start,
                const int32_t* __restrict__ row_len, const int64_t* __restrict__ batch_idx, int64_t B,
                int64_t Lmax, int64_t pad, int64_t* __restrict__ X, uint8_t* __restrict__ attn) {
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < B; r += nwarps) {
    const int64_t i = batch_idx ? __ldg(batch_idx + r) : r;
    int len = __ldg(row_len + i);
    if (len > Lmax) len = (int)Lmax;
    warp_collate_row(tok + __ldg(row_start + i), len, (int)Lmax, pad, X + r * Lmax, attn + r * Lmax);
  }
}

__global__ void __launch_bounds__(kColThreads)
batch_extents_kernel(const int32_t* __restrict__ row_len, const int64_t* __restrict__ batch_idx, int64_t G,
                     int64_t B, int32_t* __restrict__ lmax_out, int32_t* __restrict__ cells_out) {
  // one warp per batch: max over its rows
  const int lane = lane_id();
  const int64_t nb = (G + B - 1) / B;
  const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (k >= nb) return;
  const int64_t r0 = k * B, r1 = (r0 + B < G) ? r0 + B : G;
  int m = 0;
  for (int64_t r = r0 + lane; r < r1; r += kWarp) {
    const int64_t i = batch_idx ? __ldg(batch_idx + r) : r;
    m = max(m, __ldg(row_len + i));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(kFull, m, d));
  if (lane == 0) {
    const int64_t cells = (r1 - r0) * (int64_t)m;
    if (cells > 0x7fffffffll) {          // one batch of >= 2^31 cells (17 GB of int64): fail loudly, never wrap into a wrong xoff
      printf("gtok batch_extents: batch %lld has %lld rows x %d tokens >= 2^31 cells; use a smaller batch size\n",
             (long long)k, (long long)(r1 - r0), m);
      __trap();
    }
    lmax_out[k] = m;
    cells_out[k] = (int32_t)cells;
  }
}

__global__ void __launch_bounds__(kColThreads, 8)
pad_mask_batched_kernel(const int64_t* __restrict__ tok, const int64_t* __restrict__ row_start,
                        const int32_t* __restrict__ row_len, const int64_t* __restrict__ batch_idx, int64_t G,
                        int B, int64_t nb, const int32_t* __restrict__ lmax, const int64_t* __restrict__ xoff,
                        int64_t pad, int64_t* __restrict__ X, uint8_t* __restrict__ attn) {
  // one CTA per batch (grid-stride): batch extents are CTA-uniform, rows go round-robin to warps,
  // each warp fetches the metadata of its NEXT row before copying the current one.
  const int w = threadIdx.x >> 5;
  for (int64_t k = blockIdx.x; k < nb; k += gridDim.x) {
    const int Lm = __ldg(lmax + k);
    const int64_t base = __ldg(xoff + k);
    const int64_t r0 = k * B;
    const int rows = (int)((G - r0 < B) ? (G - r0) : B);
    int rr = w;
    int64_t i = 0, st = 0;
    int len = 0;
    if (rr < rows) {
      i = batch_idx ? __ldg(batch_idx + r0 + rr) : r0 + rr;
      len = __ldg(row_len + i);
      st = __ldg(row_start + i);
    }
    while (rr < rows) {
      const int nr = rr + kColWarps;
      int64_t ni = 0, nst = 0;
      int nlen = 0;
      if (nr < rows) {
        ni = batch_idx ? __ldg(batch_idx + r0 + nr) : r0 + nr;
        nlen = __ldg(row_len + ni);
        nst = __ldg(row_start + ni);
      }
      const int64_t o = base + (int64_t)rr * Lm;
      warp_collate_row(tok + st, len, Lm, pad, X + o, attn + o);
      rr = nr; i = ni; len = nlen; st = nst;
    }
  }
}

// Few, large batches (large-graph sweep: 8 batches of 1024 rows of 50k tokens): one CTA per batch would leave
// most SMs idle, so rows are dealt to ALL warps of the grid instead; each warp looks up its batch's extents.
__global__ void __launch_bounds__(kColThreads, 8)
pad_mask_rows_kernel(const int64_t* __restrict__ tok, const int64_t* __restrict__ row_start,
                     const int32_t* __restrict__ row_len, const int64_t* __restrict__ batch_idx, int64_t G, int B,
                     const int32_t* __restrict__ lmax, const int64_t* __restrict__ xoff, int64_t pad,
                     int64_t* __restrict__ X, uint8_t* __restrict__ attn) {
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < G; r += nwarps) {
    const int64_t k = r / B;
    const int rr = (int)(r - k * B);
    const int Lm = __ldg(lmax + k);
    const int64_t i = batch_idx ? __ldg(batch_idx + r) : r;
    const int64_t o = __ldg(xoff + k) + (int64_t)rr * Lm;
    warp_collate_row(tok + __ldg(row_start + i), __ldg(row_len + i), Lm, pad, X + o, attn + o);
  }
}

}  // namespace gtok

extern "C" int gtok_pad_mask(const int64_t* tok, const int64_t* row_start, const int32_t* row_len,
                             const int64_t* batch_idx, int64_t B, int64_t Lmax, int64_t pad_id, int64_t* X,
                             uint8_t* attn, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(B >= 0 && Lmax >= 0, GTOK_EINVAL, "pad_mask: negative shape");
  if (B == 0 || Lmax == 0) return GTOK_OK;
  GTOK_REQUIRE(Lmax < (1ll << 31), GTOK_EINVAL, "pad_mask: Lmax too large");
  GTOK_REQUIRE(tok && row_start && row_len && X && attn, GTOK_EINVAL, "pad_mask: null pointer");
  GTOK_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0, GTOK_EINVAL, "pad_mask: X must be 16-byte aligned");
  const int64_t work = (B + kColWarps - 1) / kColWarps;
  const int grid = persistent_grid(pad_mask_kernel, kColThreads, 0, work);
  pad_mask_kernel<<<grid, kColThreads, 0, (cudaStream_t)stream>>>(tok, row_start, row_len, batch_idx, B, Lmax, pad_id, X, attn);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

namespace gtok {
// Everything of a padded epoch that is NOT a token, plus where every row starts: row r of batch k = r / B begins at cell
// xoff[k] + (r - k*B) * lmax[k]; the cells [len, lmax) of the row get pad_id, the mask row gets len ones.  A tokenizer then
// writes the tokens straight at row_off (gtok_zinc_serialize_rows): the padded batches cost 9 * Lmax bytes per row, written
// once, for ANY batch size and molecule size (the single-pass fused kernel covers batches of 32..256 molecule-sized graphs).
__global__ void __launch_bounds__(kColThreads)
pad_tail_kernel(const int32_t* __restrict__ row_len, int64_t G, int64_t B, const int32_t* __restrict__ lmax,
                const int64_t* __restrict__ xoff, int64_t pad, int64_t* __restrict__ X, uint8_t* __restrict__ attn,
                int64_t* __restrict__ row_off) {
  const int lane = lane_id();
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < G; r += nw) {
    const int64_t k = r / B;
    const int Lm = __ldg(lmax + k);
    const int len = __ldg(row_len + r);
    const int64_t cell0 = __ldg(xoff + k) + (r - k * B) * (int64_t)Lm;
    if (lane == 0) {
      row_off[r] = cell0;
      if (r == G - 1) row_off[G] = cell0 + Lm;
    }
    int64_t* __restrict__ xr = X + cell0;
    int p0 = len;                                        // pad cells [len, Lm): scalar head to 16-byte alignment, then pairs
    if (p0 < Lm && ((reinterpret_cast<uintptr_t>(xr + p0) >> 3) & 1)) {
      if (lane == 0) st_stream_i64(xr + p0, pad);
      ++p0;
    }
    const int npair = (Lm - p0) >> 1;
    for (int j = lane; j < npair; j += kWarp) st_stream_i64x2(xr + p0 + 2 * j, pad, pad);
    if (((Lm - p0) & 1) && lane == 31) st_stream_i64(xr + Lm - 1, pad);
    uint8_t* __restrict__ ar = attn + cell0;
    int mh = (int)((16 - (reinterpret_cast<uintptr_t>(ar) & 15)) & 15);
    if (mh > Lm) mh = Lm;
    if (lane < mh) ar[lane] = (uint8_t)(lane < len);
    const int nvec = (Lm - mh) >> 4;
    for (int q = lane; q < nvec; q += kWarp) {
      const int p = mh + 16 * q;
      const int nv = len - p;
      __stcs(reinterpret_cast<ulonglong2*>(ar + p), make_ulonglong2(mask8(nv), mask8(nv - 8)));
    }
    const int tail = mh + 16 * nvec;
    if (tail + lane < Lm && lane < 16) ar[tail + lane] = (uint8_t)(tail + lane < len);
  }
}
}  // namespace gtok

extern "C" int gtok_pad_tail_batched(const int32_t* row_len, int64_t G, int64_t B, const int32_t* lmax, const int64_t* xoff,
                                     int64_t pad_id, int64_t* X, uint8_t* attn, int64_t* row_off_out, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(G >= 0 && B > 0, GTOK_EINVAL, "pad_tail_batched: need G >= 0 and B > 0");
  if (G == 0) return GTOK_OK;
  GTOK_REQUIRE(row_len && lmax && xoff && X && attn && row_off_out, GTOK_EINVAL, "pad_tail_batched: null pointer");
  const int64_t blocks = (G + kColWarps - 1) / kColWarps;
  const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 16;
  pad_tail_kernel<<<(unsigned)(blocks < cap ? blocks : cap), kColThreads, 0, (cudaStream_t)stream>>>(row_len, G, B, lmax, xoff, pad_id, X,
                                                                                                 attn, row_off_out);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

extern "C" int gtok_batch_extents(const int32_t* row_len, const int64_t* batch_idx, int64_t G, int64_t B,
                                  int32_t* lmax_out, int32_t* cells_out, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(G >= 0 && B > 0, GTOK_EINVAL, "batch_extents: need G >= 0 and B > 0");
  if (G == 0) return GTOK_OK;
  GTOK_REQUIRE(row_len && lmax_out && cells_out, GTOK_EINVAL, "batch_extents: null pointer");
  const int64_t nb = (G + B - 1) / B;
  const int64_t blocks = (nb + kColWarps - 1) / kColWarps;
  batch_extents_kernel<<<(unsigned)blocks, kColThreads, 0, (cudaStream_t)stream>>>(row_len, batch_idx, G, B, lmax_out, cells_out);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

extern "C" int gtok_pad_mask_batched(const int64_t* tok, const int64_t* row_start, const int32_t* row_len,
                                     const int64_t* batch_idx, int64_t G, int64_t B, const int32_t* lmax,
                                     const int64_t* xoff, int64_t pad_id, int64_t* X, uint8_t* attn, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(G >= 0 && B > 0, GTOK_EINVAL, "pad_mask_batched: need G >= 0 and B > 0");
  if (G == 0) return GTOK_OK;
  GTOK_REQUIRE(tok && row_start && row_len && lmax && xoff && X && attn, GTOK_EINVAL, "pad_mask_batched: null pointer");
  GTOK_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0, GTOK_EINVAL, "pad_mask_batched: X must be 16-byte aligned");
  GTOK_REQUIRE(B < (1ll << 31), GTOK_EINVAL, "pad_mask_batched: batch size too large");
  const int64_t nb = (G + B - 1) / B;
  // small batches: several CTAs per SM worth of batches per wave; large batches: one batch spans the CTA's warps anyway
  const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 8;      // resident CTAs of this kernel
  if (nb >= cap) {
    const int grid = persistent_grid(pad_mask_batched_kernel, kColThreads, 0, nb);
    pad_mask_batched_kernel<<<grid, kColThreads, 0, (cudaStream_t)stream>>>(tok, row_start, row_len, batch_idx, G, (int)B, nb, lmax, xoff, pad_id, X, attn);
  } else {
    const int grid = persistent_grid(pad_mask_rows_kernel, kColThreads, 0, (G + kColWarps - 1) / kColWarps);
    pad_mask_rows_kernel<<<grid, kColThreads, 0, (cudaStream_t)stream>>>(tok, row_start, row_len, batch_idx, G, (int)B, lmax, xoff, pad_id, X, attn);
  }
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}
