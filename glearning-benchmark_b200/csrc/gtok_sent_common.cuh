// gtok_sent_common.cuh - helpers shared by the thread-per-graph SENT kernels (gtok_sent_small.cu: <= 64 nodes,
// one 64-bit mask word; gtok_sent_mid.cu: <= 256 nodes, 2 or 4 mask words).
#pragma once

#include "gtok_common.cuh"

namespace gtok {
namespace {

constexpr int SOS = 0, RESET = 1, LADJ = 2, RADJ = 3, EOS = 4;
constexpr int kSmallThreads = 128;
constexpr int kSmallWarps = kSmallThreads / 32;

__device__ __forceinline__ void philox10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0,
                                         unsigned k1, unsigned& o0, unsigned& o1, unsigned& o2, unsigned& o3) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o0 = c0; o1 = c1; o2 = c2; o3 = c3;
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


// r / B for 0 <= r, 0 < B (both below 2^31 on the fast path): float estimate + fix-up instead of a 64-bit division
// (~70 instructions per row in the emission kernels, which are issue bound).
__device__ __forceinline__ int64_t row_batch(int64_t r, int64_t B) {
  if ((r | B) >> 31) return r / B;
  const unsigned ru = (unsigned)r, bu = (unsigned)B;
  unsigned q = (unsigned)__fdividef((float)ru, (float)bu);
  while ((unsigned long long)q * bu > ru) --q;                 // the estimate is off by at most a few units
  while ((unsigned long long)(q + 1u) * bu <= ru) ++q;
  return (int64_t)q;
}

// One warp writes a padded batch row from its narrow row buffer: X[0..Lm) = tok(q) for q < L, pad beyond; the mask bytes come
// from a 17-entry table in shared memory (tbl[k] = 16 bytes whose low k are 0x01, fill_mask_table).  Like warp_pad_row, with the
// mask half on the table.
template <typename F>
__device__ __forceinline__ void warp_pad_row_tbl(int len, int Lmax, int64_t pad, int64_t* __restrict__ xrow,
                                                 uint8_t* __restrict__ arow, const ulonglong2* __restrict__ tbl, F tok) {
  const int lane = lane_id();
  const int head = (int)((reinterpret_cast<uintptr_t>(xrow) >> 3) & 1);
  if (head && lane == 0 && Lmax > 0) st_stream_i64(xrow, 0 < len ? (int64_t)tok(0) : pad);
  const int npair = (Lmax - head) >> 1;
  const int nfull = (len - head) >> 1;                   // pairs that lie entirely inside the row
  for (int j = lane; j < npair; j += kWarp) {
    const int p = head + 2 * j;
    int64_t a = pad, b = pad;
    if (j < nfull) { a = (int64_t)tok(p); b = (int64_t)tok(p + 1); }
    else if (p < len) a = (int64_t)tok(p);
    st_stream_i64x2(xrow + p, a, b);
  }
  const int tailx = head + 2 * npair;
  if (tailx < Lmax && lane == 31) st_stream_i64(xrow + tailx, tailx < len ? (int64_t)tok(tailx) : pad);
  int mh = (int)((16 - (reinterpret_cast<uintptr_t>(arow) & 15)) & 15);
  if (mh > Lmax) mh = Lmax;
  if (lane < mh) arow[lane] = (uint8_t)(lane < len);
  const int nvec = (Lmax - mh) >> 4;
  for (int q = lane; q < nvec; q += kWarp) {
    const int p = mh + 16 * q;
    int nv = len - p;
    nv = nv < 0 ? 0 : (nv > 16 ? 16 : nv);
    __stcs(reinterpret_cast<ulonglong2*>(arow + p), tbl[nv]);
  }
  const int tail = mh + 16 * nvec;
  if (tail + lane < Lmax && lane < 16) arow[tail + lane] = (uint8_t)(tail + lane < len);
}

}  // namespace
}  // namespace gtok
