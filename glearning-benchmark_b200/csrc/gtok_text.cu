// gtok_text.cu - ingest: graph-token TEXTS -> flat graph arrays, on the GPU.
//
// Replaces the per-text host parse in front of the tokenizers (the inverse of the external graph-token
// serializer): graph_data_loader/graph_token_dataset_autograph.py:14-54 parse_graph_from_text and the
// label/query parsers graph_data_loader/data_loader.py:19-74, i.e. what the drop-in TokenDataset has to do
// before any CUDA tokenizer can run when it is handed plain strings.  Grammar (docs/synthetic_data.md:46-68):
//     <bos> (u v <e>)*E <n> 0 1 .. N-1 <q> (has_cycle | shortest_distance u v) <p> LABEL <eos>
// Two passes over the concatenated ASCII bytes, one warp per text, one byte per lane and trip:
//   gtok_text_scan   token starts by ballot, positions of <n> / <q> / <p>  ->  N, E, task, query u v
//   gtok_text_edges  every edge token parsed as a decimal, every node-list token checked to be its own
//                    index, every third edge-region token checked to be <e>  ->  src, dst, status
// A text that is not of the grammar gets task = -1 / status = 0; the host raises (there is no CPU fallback).
#include "gtok_common.cuh"

namespace gtok {

constexpr int kTxtThreads = 256;
constexpr int kIntMax = 0x7fffffff;

__device__ __forceinline__ bool is_sp(unsigned c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r'; }

// decimal token starting at p (not a space); -1 unless it is 1..9 digits followed by a space / the end
__device__ __forceinline__ int parse_uint(const uint8_t* __restrict__ t, int64_t p, int64_t end) {
  int v = 0, n = 0;
  while (p < end) {
    const unsigned c = t[p];
    if (is_sp(c)) break;
    if (c < '0' || c > '9' || n >= 9) return -1;
    v = v * 10 + (int)(c - '0');
    ++n;
    ++p;
  }
  return n ? v : -1;
}

// does the token starting at p equal the literal s[0..len)?
__device__ __forceinline__ bool tok_is(const uint8_t* __restrict__ t, int64_t p, int64_t end, const char* s, int len) {
  if (p + len > end) return false;
  for (int i = 0; i < len; ++i)
    if (t[p + i] != (uint8_t)s[i]) return false;
  return p + len == end || is_sp(t[p + len]);
}

// next token start at or after p (end if none)
__device__ __forceinline__ int64_t skip_sp(const uint8_t* __restrict__ t, int64_t p, int64_t end) {
  while (p < end && is_sp(t[p])) ++p;
  return p;
}
__device__ __forceinline__ int64_t skip_tok(const uint8_t* __restrict__ t, int64_t p, int64_t end) {
  while (p < end && !is_sp(t[p])) ++p;
  return p;
}

__global__ void __launch_bounds__(kTxtThreads)
text_scan_kernel(const uint8_t* __restrict__ text, const int64_t* __restrict__ off, int64_t G, int32_t* __restrict__ nn,
                 int32_t* __restrict__ ne, int32_t* __restrict__ task, int32_t* __restrict__ qu, int32_t* __restrict__ qv) {
  const int lane = lane_id();
  const unsigned lt = (1u << lane) - 1u;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = warp0; g < G; g += nwarps) {
    const int64_t b0 = __ldg(off + g), b1 = __ldg(off + g + 1);
    int ntok = 0, tn = kIntMax, tq = kIntMax, tp = kIntMax, pq = kIntMax, bos = 0;
    unsigned carry = ' ';
    for (int64_t p0 = b0; p0 < b1; p0 += kWarp) {
      const int64_t p = p0 + lane;
      const unsigned c = p < b1 ? text[p] : ' ';
      unsigned prev = __shfl_up_sync(kFull, c, 1);
      if (lane == 0) prev = carry;
      const bool start = !is_sp(c) && is_sp(prev);
      const unsigned m = __ballot_sync(kFull, start);
      const int idx = ntok + __popc(m & lt);
      if (start && c == '<') {
        if (tok_is(text, p, b1, "<n>", 3)) tn = min(tn, idx);
        else if (tok_is(text, p, b1, "<q>", 3)) { if (idx < tq) { tq = idx; pq = (int)(p - b0); } }
        else if (tok_is(text, p, b1, "<p>", 3)) tp = min(tp, idx);
        else if (idx == 0 && tok_is(text, p, b1, "<bos>", 5)) bos = 1;
      }
      ntok += __popc(m);
      carry = __shfl_sync(kFull, c, 31);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      tn = min(tn, __shfl_xor_sync(kFull, tn, d));
      tp = min(tp, __shfl_xor_sync(kFull, tp, d));
      const int oq = __shfl_xor_sync(kFull, tq, d), op = __shfl_xor_sync(kFull, pq, d);
      if (oq < tq) { tq = oq; pq = op; }
      bos |= __shfl_xor_sync(kFull, bos, d);
    }
    // (that every third token of the edge region is <e>, and the node list itself, are pass 2's checks)
    if (lane == 0) {
      int t = -1, u = 0, v = 0, N = 0, E = 0;
      const bool shape = bos && tn != kIntMax && tq != kIntMax && tp != kIntMax && tn >= 1 && (tn - 1) % 3 == 0 && tq > tn && ntok >= 5;
      if (shape) {
        E = (tn - 1) / 3;
        N = tq - tn - 1;
        int64_t p = skip_sp(text, skip_tok(text, b0 + pq, b1), b1);            // token after <q>
        if (tok_is(text, p, b1, "has_cycle", 9) && tp == tq + 2) {
          t = GTOK_TASK_CYCLE_CHECK;
        } else if (tok_is(text, p, b1, "shortest_distance", 17) && tp == tq + 4) {
          p = skip_sp(text, skip_tok(text, p, b1), b1);
          u = parse_uint(text, p, b1);
          p = skip_sp(text, skip_tok(text, p, b1), b1);
          v = parse_uint(text, p, b1);
          if (u >= 0 && v >= 0) t = GTOK_TASK_SHORTEST_PATH;
        }
      }
      nn[g] = t >= 0 ? N : 0;
      ne[g] = t >= 0 ? E : 0;
      task[g] = t;
      if (qu) qu[g] = t == GTOK_TASK_SHORTEST_PATH ? u : 0;
      if (qv) qv[g] = t == GTOK_TASK_SHORTEST_PATH ? v : 0;
    }
  }
}

__global__ void __launch_bounds__(kTxtThreads)
text_edges_kernel(const uint8_t* __restrict__ text, const int64_t* __restrict__ off, int64_t G,
                  const int32_t* __restrict__ nn, const int32_t* __restrict__ task, const int64_t* __restrict__ edge_off,
                  int32_t* __restrict__ src, int32_t* __restrict__ dst, int32_t* __restrict__ status) {
  const int lane = lane_id();
  const unsigned lt = (1u << lane) - 1u;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = warp0; g < G; g += nwarps) {
    if (__ldg(task + g) < 0) {
      if (lane == 0) status[g] = 0;
      continue;
    }
    const int64_t b0 = __ldg(off + g), b1 = __ldg(off + g + 1);
    const int64_t e0 = __ldg(edge_off + g);
    const int E = (int)(__ldg(edge_off + g + 1) - e0), N = __ldg(nn + g);
    const int tn = 3 * E + 1, tq = tn + N + 1;
    int ntok = 0;
    bool ok = true;
    unsigned carry = ' ';
    for (int64_t p0 = b0; p0 < b1; p0 += kWarp) {
      const int64_t p = p0 + lane;
      const unsigned c = p < b1 ? text[p] : ' ';
      unsigned prev = __shfl_up_sync(kFull, c, 1);
      if (lane == 0) prev = carry;
      const bool start = !is_sp(c) && is_sp(prev);
      const unsigned m = __ballot_sync(kFull, start);
      const int idx = ntok + __popc(m & lt);
      if (start) {
        if (idx >= 1 && idx < tn) {
          const int e = (idx - 1) / 3, k = (idx - 1) - 3 * e;
          if (k == 2) {
            ok = ok && tok_is(text, p, b1, "<e>", 3);
          } else {
            const int v = parse_uint(text, p, b1);
            ok = ok && v >= 0;
            (k == 0 ? src : dst)[e0 + e] = v < 0 ? 0 : v;
          }
        } else if (idx > tn && idx < tq) {
          ok = ok && parse_uint(text, p, b1) == idx - tn - 1;           // the node list is 0 1 .. N-1
        }
      }
      ntok += __popc(m);
      carry = __shfl_sync(kFull, c, 31);
    }
    ok = __all_sync(kFull, ok);
    if (lane == 0) status[g] = ok ? 1 : 0;
  }
}

}  // namespace gtok

extern "C" int gtok_text_scan(const uint8_t* text, const int64_t* text_off, int64_t num_texts, int32_t* num_nodes,
                              int32_t* num_edges, int32_t* task, int32_t* query_u, int32_t* query_v, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(num_texts >= 0, GTOK_EINVAL, "text_scan: negative text count");
  if (num_texts == 0) return GTOK_OK;
  GTOK_REQUIRE(text && text_off && num_nodes && num_edges && task, GTOK_EINVAL, "text_scan: null pointer");
  const int64_t work = (num_texts + (kTxtThreads / 32) - 1) / (kTxtThreads / 32);
  const int grid = persistent_grid(text_scan_kernel, kTxtThreads, 0, work);
  text_scan_kernel<<<grid, kTxtThreads, 0, (cudaStream_t)stream>>>(text, text_off, num_texts, num_nodes, num_edges, task,
                                                                    query_u, query_v);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

extern "C" int gtok_text_edges(const uint8_t* text, const int64_t* text_off, int64_t num_texts, const int32_t* num_nodes,
                               const int32_t* task, const int64_t* edge_off, int32_t* src, int32_t* dst, int32_t* status,
                               void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(num_texts >= 0, GTOK_EINVAL, "text_edges: negative text count");
  if (num_texts == 0) return GTOK_OK;
  GTOK_REQUIRE(text && text_off && num_nodes && task && edge_off && status, GTOK_EINVAL, "text_edges: null pointer");
  const int64_t work = (num_texts + (kTxtThreads / 32) - 1) / (kTxtThreads / 32);
  const int grid = persistent_grid(text_edges_kernel, kTxtThreads, 0, work);
  text_edges_kernel<<<grid, kTxtThreads, 0, (cudaStream_t)stream>>>(text, text_off, num_texts, num_nodes, task, edge_off, src, dst,
                                                                     status);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

// ---- ids -> text bytes (the other direction): ZINCTokenizationDataset items carry their text because the unchanged
// trainer reads it to collect its vocabulary (trainer/train_ibtt.py:231-233, :364-369).  Rendering 220k molecules token
// by token in Python is hot loop #3 of the reference path; here every token of the ragged id stream writes its word (+ one
// blank) into one byte buffer, the host cuts an item's text out of it with one slice.
//   gtok_detok_lengths  chars[i] = word_len[tok[i]] + 1   (then K0 scan -> char_off)
//   gtok_detok_render   text[char_off[i] ..] = word(tok[i]) + ' '
//   gtok_first_occurrence  first[id] = smallest i with tok[i] == id (UINT64_MAX: never) - the first-seen order of the
//                          dynamic vocabulary tokens (extend_vocab_with_dynamic_tokens, zinc_vocab.py:154-179)
namespace gtok {

constexpr int kDetokThreads = 256;

__global__ void __launch_bounds__(kDetokThreads)
detok_lengths_kernel(const int64_t* __restrict__ tok, int64_t n, const int32_t* __restrict__ word_off, int32_t num_words,
                     int32_t* __restrict__ chars) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long t = __ldcs(tok + i);
    chars[i] = (t >= 0 && t < num_words) ? (__ldg(word_off + t + 1) - __ldg(word_off + t) + 1) : 2;   // unknown id: "? "
  }
}

__global__ void __launch_bounds__(kDetokThreads)
detok_render_kernel(const int64_t* __restrict__ tok, int64_t n, const uint8_t* __restrict__ words,
                    const int32_t* __restrict__ word_off, int32_t num_words, const int64_t* __restrict__ char_off,
                    uint8_t* __restrict__ text) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long t = __ldcs(tok + i);
    uint8_t* __restrict__ dst = text + __ldg(char_off + i);
    if (t >= 0 && t < num_words) {
      const int a = __ldg(word_off + t), b = __ldg(word_off + t + 1);
      for (int j = a; j < b; ++j) dst[j - a] = __ldg(words + j);
      dst[b - a] = ' ';
    } else {
      dst[0] = '?';
      dst[1] = ' ';
    }
  }
}

__global__ void __launch_bounds__(kDetokThreads)
first_occurrence_kernel(const int64_t* __restrict__ tok, int64_t n, int32_t num_ids, unsigned long long* __restrict__ first) {
  extern __shared__ unsigned long long s_first[];
  for (int i = threadIdx.x; i < num_ids; i += blockDim.x) s_first[i] = ~0ull;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long t = __ldcs(tok + i);
    if (t >= 0 && t < num_ids && s_first[t] > (unsigned long long)i) atomicMin(&s_first[t], (unsigned long long)i);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < num_ids; i += blockDim.x)
    if (s_first[i] != ~0ull) atomicMin(first + i, s_first[i]);
}

}  // namespace gtok

extern "C" int gtok_detok_lengths(const int64_t* tok, int64_t num_tokens, const int32_t* word_off, int32_t num_words,
                                  int32_t* chars_out, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(num_tokens >= 0 && num_words >= 0, GTOK_EINVAL, "detok_lengths: negative size");
  if (num_tokens == 0) return GTOK_OK;
  GTOK_REQUIRE(tok && word_off && chars_out, GTOK_EINVAL, "detok_lengths: null pointer");
  const int grid = persistent_grid(detok_lengths_kernel, kDetokThreads, 0, (num_tokens + kDetokThreads - 1) / kDetokThreads);
  detok_lengths_kernel<<<grid, kDetokThreads, 0, (cudaStream_t)stream>>>(tok, num_tokens, word_off, num_words, chars_out);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

extern "C" int gtok_detok_render(const int64_t* tok, int64_t num_tokens, const uint8_t* words, const int32_t* word_off,
                                 int32_t num_words, const int64_t* char_off, uint8_t* text_out, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(num_tokens >= 0 && num_words >= 0, GTOK_EINVAL, "detok_render: negative size");
  if (num_tokens == 0) return GTOK_OK;
  GTOK_REQUIRE(tok && words && word_off && char_off && text_out, GTOK_EINVAL, "detok_render: null pointer");
  const int grid = persistent_grid(detok_render_kernel, kDetokThreads, 0, (num_tokens + kDetokThreads - 1) / kDetokThreads);
  detok_render_kernel<<<grid, kDetokThreads, 0, (cudaStream_t)stream>>>(tok, num_tokens, words, word_off, num_words, char_off,
                                                                        text_out);
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}

extern "C" int gtok_first_occurrence(const int64_t* tok, int64_t num_tokens, int32_t num_ids, uint64_t* first_out, void* stream) {
  using namespace gtok;
  GTOK_REQUIRE(num_tokens >= 0 && num_ids >= 0, GTOK_EINVAL, "first_occurrence: negative size");
  GTOK_REQUIRE(num_ids <= 4096, GTOK_EUNSUPPORTED, "first_occurrence: at most 4096 distinct ids (got %d)", num_ids);
  if (num_ids == 0) return GTOK_OK;
  GTOK_REQUIRE(first_out != nullptr, GTOK_EINVAL, "first_occurrence: null output");
  cudaStream_t st = (cudaStream_t)stream;
  GTOK_CUDA_TRY(cudaMemsetAsync(first_out, 0xff, (size_t)num_ids * 8, st));
  if (num_tokens == 0) return GTOK_OK;
  GTOK_REQUIRE(tok != nullptr, GTOK_EINVAL, "first_occurrence: null tokens");
  int64_t blocks = (num_tokens + kDetokThreads * 8 - 1) / (kDetokThreads * 8);
  const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 4;
  if (blocks > cap) blocks = cap;
  first_occurrence_kernel<<<(unsigned)blocks, kDetokThreads, (size_t)num_ids * 8, st>>>(
      tok, num_tokens, num_ids, reinterpret_cast<unsigned long long*>(first_out));
  GTOK_CUDA_TRY(cudaGetLastError());
  return GTOK_OK;
}
