// gtok_common.cuh - shared device/host helpers for the sm_100a tokenization kernels.
//
// All kernels here are HBM-bound integer gathers/scatters (no FLOPs, no tensor
// cores).  The rules that matter (blackwell_cuda_programming.md G1, G2, G11, G13):
// full-sector coalesced 16-byte stores of the int64 token stream, grids sized as
// a multiple of the SM count with enough resident warps to cover DRAM latency, and
// small lookup tables served from L1/shared memory.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gtok.h"

namespace gtok {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int sm_count();  // cached, <0 on error

#define GTOK_CUDA_TRY(expr)                                   \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return ::gtok::cuda_fail(_e, #expr); \
  } while (0)

#define GTOK_REQUIRE(cond, code, ...)      \
  do {                                     \
    if (!(cond)) {                         \
      ::gtok::set_error(__VA_ARGS__);      \
      return (code);                       \
    }                                      \
  } while (0)

// Launch geometry for "one warp per graph, persistent grid-stride" kernels.
template <typename K>
inline int persistent_grid(K kernel, int block, size_t smem, int64_t work_blocks) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  int64_t cap = (int64_t)sm_count() * per_sm;
  if (cap < 1) cap = 148;
  int64_t g = work_blocks < cap ? work_blocks : cap;
  return (int)(g < 1 ? 1 : g);
}

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// streaming (evict-first) stores: the token stream is written once and consumed
// by a later kernel / the trainer, never re-read by the producer.
__device__ __forceinline__ void st_stream_i64x2(int64_t* p, int64_t a, int64_t b) {
  __stcs(reinterpret_cast<longlong2*>(p), make_longlong2(a, b));
}
__device__ __forceinline__ void st_stream_i64(int64_t* p, int64_t a) { __stcs(p, a); }

// Write row[0..L) = tok(p) with one warp: 16-byte vector stores on the 16-byte
// aligned interior, scalar stores on the (at most one) head and tail element.
// tok(p) must be callable for any 0 <= p < L by any lane.
template <typename F>
__device__ __forceinline__ void warp_write_row(int64_t* __restrict__ row, int L, F tok) {
  const int lane = lane_id();
  const int head = (int)((reinterpret_cast<uintptr_t>(row) >> 3) & 1);  // 1 if row is 8- but not 16-byte aligned
  if (head && lane == 0 && L > 0) st_stream_i64(row, (int64_t)tok(0));
  const int npair = (L - head) >> 1;
  for (int j = lane; j < npair; j += kWarp) {
    const int p = head + 2 * j;
    const int64_t a = (int64_t)tok(p);
    const int64_t b = (int64_t)tok(p + 1);
    st_stream_i64x2(row + p, a, b);
  }
  const int tail = head + 2 * npair;
  if (tail < L && lane == 31) st_stream_i64(row + tail, (int64_t)tok(tail));
}


// One warp writes one row of a padded batch: X[0..Lmax) = tok(p) for p < len else pad, attn[0..Lmax) =
// (p < len), both with 16-byte streaming stores on the aligned interior (collate, data_loader.py:488-497).
// tok(p) must be callable for any 0 <= p < len by any lane.
__device__ __forceinline__ unsigned long long mask8(int nvalid) {   // low `nvalid` bytes (0..8) = 0x01
  if (nvalid <= 0) return 0ull;
  if (nvalid >= 8) return 0x0101010101010101ull;
  return 0x0101010101010101ull >> (8 * (8 - nvalid));
}
template <typename F>
__device__ __forceinline__ void warp_pad_row(int len, int Lmax, int64_t pad, int64_t* __restrict__ xrow,
                                             uint8_t* __restrict__ arow, F tok) {
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
    const int nv = len - p;
    __stcs(reinterpret_cast<ulonglong2*>(arow + p), make_ulonglong2(mask8(nv), mask8(nv - 8)));
  }
  const int tail = mh + 16 * nvec;
  if (tail + lane < Lmax && lane < 16) arow[tail + lane] = (uint8_t)(tail + lane < len);
}

// Two consecutive rows of a padded batch are ONE contiguous run of 2*Lmax mask bytes: byte p is (p < len0) in the
// first row, (p - Lmax < len1) in the second.  Same 16-byte streaming stores as warp_pad_row's mask half.
__device__ __forceinline__ void warp_mask_run2(int len0, int len1, int Lmax, uint8_t* __restrict__ arow) {
  const int lane = lane_id();
  const int total = 2 * Lmax;
  auto val = [&](int p) -> uint8_t { return (uint8_t)(p < Lmax ? (p < len0) : (p - Lmax < len1)); };
  int mh = (int)((16 - (reinterpret_cast<uintptr_t>(arow) & 15)) & 15);
  if (mh > total) mh = total;
  if (lane < mh) arow[lane] = val(lane);
  const int nvec = (total - mh) >> 4;
  for (int q = lane; q < nvec; q += kWarp) {
    const int p = mh + 16 * q;
    const int a = len0 - p, b = Lmax + len1 - p, c = Lmax - p;       // ones below a, and in [c, b)
    const unsigned long long lo = mask8(a) | (mask8(b) & ~mask8(c));
    const unsigned long long hi = mask8(a - 8) | (mask8(b - 8) & ~mask8(c - 8));
    __stcs(reinterpret_cast<ulonglong2*>(arow + p), make_ulonglong2(lo, hi));
  }
  const int tail = mh + 16 * nvec;
  if (tail + lane < total && lane < 16) arow[tail + lane] = val(tail + lane);
}

// Table variant: tbl[k] (k = 0..16) is the 16-byte mask whose low k bytes are 0x01 (shared memory, filled once per CTA
// with fill_mask_table): three 16-byte loads replace ~30 ALU instructions per vector.
__device__ __forceinline__ void fill_mask_table(ulonglong2* tbl, int tid) {
  if (tid <= 16) tbl[tid] = make_ulonglong2(mask8(tid), mask8(tid - 8));
}
__device__ __forceinline__ void warp_mask_run2(int len0, int len1, int Lmax, uint8_t* __restrict__ arow,
                                               const ulonglong2* __restrict__ tbl) {
  const int lane = lane_id();
  const int total = 2 * Lmax;
  auto val = [&](int p) -> uint8_t { return (uint8_t)(p < Lmax ? (p < len0) : (p - Lmax < len1)); };
  int mh = (int)((16 - (reinterpret_cast<uintptr_t>(arow) & 15)) & 15);
  if (mh > total) mh = total;
  if (lane < mh) arow[lane] = val(lane);
  const int nvec = (total - mh) >> 4;
  for (int q = lane; q < nvec; q += kWarp) {
    const int p = mh + 16 * q;
    const ulonglong2 A = tbl[min(max(len0 - p, 0), 16)];
    const ulonglong2 B = tbl[min(max(Lmax + len1 - p, 0), 16)];
    const ulonglong2 C = tbl[min(max(Lmax - p, 0), 16)];
    __stcs(reinterpret_cast<ulonglong2*>(arow + p), make_ulonglong2(A.x | (B.x & ~C.x), A.y | (B.y & ~C.y)));
  }
  const int tail = mh + 16 * nvec;
  if (tail + lane < total && lane < 16) arow[tail + lane] = val(tail + lane);
}

// ---- bulk asynchronous copies (TMA engine, 1-D): shared memory -> global, tracked by bulk async-groups ---------------
// The copy is issued by ONE thread and moves a whole 16-byte aligned run without touching a register; the issuing
// thread's generic-proxy writes (and, after a __syncwarp / barrier, everybody else's) must be made visible to the async
// proxy with fence_proxy_async_smem() before the copy is issued.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
// same, with an L2 evict-first policy: the run is written once and never re-read by this kernel, so it must not push
// the L2-prefetched inputs of the coming tiles out of the cache (what __stcs does for register stores)
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_store_s2g_hint(void* gdst, const void* ssrc, unsigned bytes, unsigned long long pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
               "r"(smem_u32(ssrc)), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {   // at most N groups still READING their shared-memory source
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {        // at most N groups not yet complete (writes performed)
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---- decoupled look-back (single-pass chained scan) -------------------------
// One 64-bit word per tile carries {flag (top 2 bits) | value (62 bits)}, so flag
// and value can never be observed out of order; relaxed gpu-scope accesses suffice.
#ifndef GTOK_LOOKBACK_SLEEP_NS
#define GTOK_LOOKBACK_SLEEP_NS 64
#endif
constexpr unsigned kLookbackSleepNs = GTOK_LOOKBACK_SLEEP_NS;
constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagIncl = 2ull << 62;
constexpr unsigned long long kValMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by ONE full warp of the tile's CTA.  Publishes `total` for `tile`, walks
// back over predecessor tiles and returns the exclusive prefix (same value in all lanes).
// Tiles must be numbered by an atomic ticket so predecessors are already running.
__device__ __forceinline__ long long lookback_exclusive(unsigned long long* status, long long tile, long long total) {
  const int t = lane_id();
  if (t == 0) st_relaxed_u64(status + tile, (tile == 0 ? kFlagIncl : kFlagAgg) | (unsigned long long)total);
  long long excl = 0;
  long long j = tile - 1;
  while (j >= 0) {
    const long long idx = j - t;
    unsigned long long w;
    if (idx >= 0) {
      w = ld_relaxed_u64(status + idx);
      while ((w >> 62) == 0) {                          // predecessor not published yet: back off instead of hammering L2
        __nanosleep(kLookbackSleepNs);
        w = ld_relaxed_u64(status + idx);
      }
    } else {
      w = kFlagIncl;  // virtual tile -1: inclusive prefix 0
    }
    const unsigned incl_mask = __ballot_sync(kFull, (w >> 62) == 2);
    const int stop = incl_mask ? (__ffs(incl_mask) - 1) : 31;
    long long val = (t <= stop) ? (long long)(w & kValMask) : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(kFull, val, d);
    excl += val;
    if (incl_mask) break;
    j -= 32;
  }
  if (t == 0 && tile != 0) st_relaxed_u64(status + tile, kFlagIncl | (unsigned long long)(excl + total));
  return excl;
}

}  // namespace gtok
