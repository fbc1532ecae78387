/*
 * gtok.h - C ABI of the B200 graph-tokenization engine (libgtok.so).
 *
 * Drop-in boundary for ONE hot path of KevinBian107/GLearning-Benchmark: batched
 * graph -> token-sequence -> padded batch, feeding its IBTT and AGTT transformers.
 * The reference has no FFI on this path (it is pure Python); each
 * entry point below states the reference Python it replaces (file:line under the
 * reference checkout).  The reference-side binding (ctypes) is shown in
 * INTEGRATION.md and shipped as glearning-benchmark_b200/gtok/_lib.py.
 *
 * Conventions
 *   - every pointer inside the structs and every buffer argument is a DEVICE
 *     pointer owned by the caller (the host side passes torch CUDA tensors'
 *     data_ptr()); the library allocates nothing persistent and keeps no state
 *     besides a thread-local error string.  The only exceptions are marked HOST.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *     All work is enqueued on it; no call synchronises the device.
 *   - return 0 on success, a negative GTOK_E* otherwise; gtok_last_error() gives
 *     the message for the calling thread.
 *   - token ids are int64 on output (torch.long, data_loader.py:483,491), graph
 *     structure is int32, offsets are int64.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     returns GTOK_ECUDA.
 */
#ifndef GTOK_H_
#define GTOK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GTOK_VERSION 101 /* 0.1.1 */

#if defined(__GNUC__)
#define GTOK_API __attribute__((visibility("default")))
#else
#define GTOK_API
#endif

enum {
  GTOK_OK = 0,
  GTOK_EINVAL = -1,       /* bad argument (null pointer, negative size, ...) */
  GTOK_ECUDA = -2,        /* CUDA runtime error / no device */
  GTOK_EUNSUPPORTED = -3, /* shape outside what the kernels cover (see message) */
  GTOK_EWORKSPACE = -4    /* workspace too small */
};

/* Disjoint union of G graphs in edge-index (COO) form, caller's column order.
 * Mirrors PyG Data(edge_index, x, edge_attr, num_nodes, query_u, query_v) as the
 * reference builds it (graph_token_dataset_autograph.py:385-394,
 * zinc_dataset_autograph.py:51-73). */
typedef struct GtokGraphBatch {
  int64_t num_graphs;
  const int64_t* node_off;   /* [G+1] */
  const int64_t* edge_off;   /* [G+1] directed entries of graph g: [edge_off[g], edge_off[g+1]) */
  const int32_t* src;        /* [E] local node ids */
  const int32_t* dst;        /* [E] */
  const int32_t* x;          /* [N] node type, or NULL */
  const int32_t* edge_attr;  /* [E] edge type, or NULL */
  const int32_t* query_u;    /* [G] or NULL (shortest_path) */
  const int32_t* query_v;    /* [G] or NULL */
  int32_t max_nodes;         /* upper bound on nodes per graph   (0 = unknown) */
  int32_t max_entries;       /* upper bound on entries per graph (0 = unknown) */
  int32_t max_nodes_per_32;  /* optional tighter staging bounds for the thread-per-molecule ZINC kernels: upper bound on */
  int32_t max_entries_per_32;/* the nodes / entries of any aligned group of 32 consecutive graphs [32k, 32k+32) (one warp's
                                staging run); 0 = unknown (32 * max_nodes / 32 * max_entries is assumed) */
} GtokGraphBatch;

/* graph-token rule set (docs/synthetic_data.md:46-68, docs/models.md:22) with the
 * vocabulary already resolved by the unchanged host code (build_vocab_from_texts,
 * data_loader.py:451-463): the kernels never assume id(str(k)) = base + k. */
enum { GTOK_TASK_CYCLE_CHECK = 0, GTOK_TASK_SHORTEST_PATH = 1 };
typedef struct GtokIbttRules {
  int32_t task;          /* GTOK_TASK_* */
  int32_t max_len;       /* TokenDataset max_len: ids[:max_len] (data_loader.py:482) */
  int32_t tok_bos, tok_e, tok_n, tok_q, tok_p;
  int32_t tok_task;      /* id("has_cycle") or id("shortest_distance") */
  int32_t oov;           /* vocab["<pad>"]: id of any token missing from the vocab */
  int32_t node_tok_len;
  const int32_t* node_tok; /* [node_tok_len] node_tok[k] = vocab.get(str(k), oov) */
  int32_t max_token_id;  /* upper bound on every id above incl. node_tok[] (0 = unknown); ids < 65536 let
                            gtok_ibtt_collate stage rows as 16-bit in shared memory */
  int32_t reserved_;
} GtokIbttRules;

/* ZINC molecular rule set (zinc_dataset_indexbase.py:143-227 + zinc_vocab.py:35-78). */
typedef struct GtokZincVocab {
  int32_t item_max_len;  /* ZINCTokenizationDataset.max_len (text truncation, :216-221) */
  int32_t ds_max_len;    /* TokenDataset max_len (data_loader.py:482) */
  int32_t tok_bos, tok_eos, tok_atom, tok_bond, tok_q, tok_regression, tok_p;
  int32_t atom_tok[9];   /* id of 'C','N','O','F','P','S','Cl','Br','I' */
  int32_t atom_oov;      /* id of 'X' (x outside 0..8, zinc_dataset_indexbase.py:104) */
  int32_t bond_tok[4];   /* id of single,double,triple,aromatic (edge_attr 1..4) */
  int32_t bond_oov;      /* id of 'unknown' (:138-139) */
  int32_t oov;           /* vocab["<pad>"] */
  int32_t node_tok_len;
  const int32_t* node_tok; /* [node_tok_len] id(str(k)) */
  int32_t max_token_id;  /* upper bound on every id above incl. node_tok[] (0 = unknown); ids < 65536
                            let the fused kernel stage rows as 16-bit in shared memory */
  int32_t reserved_;
} GtokZincVocab;

/* Canonical undirected CSR of the same G graphs, what SENT walks every epoch:
 * rows sorted by neighbour id, no self loops, no duplicate neighbours.  Built once
 * per dataset (host side: gtok/csr.py). */
typedef struct GtokCsrBatch {
  int64_t num_graphs;
  const int64_t* node_off;   /* [G+1] */
  const int32_t* rowptr;     /* [N_total+1] global: neighbours of global node i are col[rowptr[i] .. rowptr[i+1]) */
  const int32_t* col;        /* [nnz] LOCAL neighbour ids */
  const int32_t* eattr;      /* [nnz] edge type per CSR entry, or NULL */
  const int32_t* x;          /* [N_total] node type, or NULL */
  const int32_t* query_u;    /* [G] or NULL */
  const int32_t* query_v;    /* [G] or NULL */
  int32_t max_nodes;         /* upper bound on nodes per graph (required, > 0) */
  int32_t max_nnz;           /* upper bound on CSR entries per graph (required, > 0) */
  int32_t max_label;         /* every x[] and eattr[] value lies in [0, max_label]; 0 = unknown (the
                                thread-per-graph kernel stages labels as 8/16-bit and needs the bound) */
  int32_t adj_words;         /* 64-bit words per adjacency row in adj64: 0 or 1 (max_nodes <= 64), 2 (<= 128), 4 (<= 256) */
  const uint64_t* adj64;     /* [N_total * adj_words] optional: bit v of row node_off[g]+u set iff {u,v} is an edge of
                                graph g (enables the thread-per-graph SENT kernels; graphs up to 256 nodes) */
  const int32_t* order;      /* [G] optional processing order (a permutation of 0..G-1, e.g. graphs sorted by
                                size so the 32 graphs of a warp take similar time); rows are still written
                                at their own index g */
} GtokCsrBatch;

/* SENT configuration: the observable contract of AutoGraph's Graph2TrailTokenizer
 * (train_agtt.py:514-542) plus the fused post-processing of
 * TokenizedGraphDataset.__getitem__ (train_agtt.py:246-273). */
typedef struct GtokSentCfg {
  int32_t idx_offset;        /* 6 + len(dataset_names) */
  int32_t node_idx_offset;   /* idx_offset + max_num_nodes        (labeled) */
  int32_t edge_idx_offset;   /* node_idx_offset + num_node_types  (labeled) */
  int32_t max_length;        /* tokens[:max_length] */
  int32_t labeled;           /* labeled_graph */
  int32_t edge_label_shift;  /* edge label token = edge_idx_offset + attr + shift */
  int32_t remap_zinc;        /* apply remap_zinc_tokens (train_agtt.py:171-244) */
  int32_t append_query;      /* append [idx_offset+N, idx_offset+u, idx_offset+v] (:257-267) */
  /* Decisions the reference leaves open (AutoGraph's source is not in its checkout; oracle/sent_oracle.py SentSpec
   * holds the full list, tools/pin_autograph.py finds the setting that reproduces the real tokenizer).  0 = this
   * repo's normative choice. */
  int32_t restart_policy;    /* GTOK_SENT_RESTART_*: where a new segment starts when the walk is stuck */
  int32_t bracket_order;     /* GTOK_SENT_BRACKET_*: order of the neighbourhood set between LADJ and RADJ */
  int32_t no_eos;            /* 1 = Graph2TrailTokenizer(append_eos=False): the sequence ends without EOS */
  int32_t reserved0;         /* (keeps the struct a multiple of 8 bytes) */
} GtokSentCfg;
enum { GTOK_SENT_RESTART_OPEN_FIRST = 0, /* visited nodes that still have an unvisited neighbour, else unvisited nodes */
       GTOK_SENT_RESTART_UNVISITED = 1   /* any unvisited node */ };
enum { GTOK_SENT_BRACKET_VISIT_ASC = 0, GTOK_SENT_BRACKET_VISIT_DESC = 1 };

/* Where the walk's random words come from (oracle/sent_oracle.py "pick"). */
typedef struct GtokChoice {
  uint64_t seed;               /* Philox-4x32-10 key when stream == NULL */
  int64_t graph_id_base;       /* global id of graph 0 of this batch (rank sharding keeps streams invariant) */
  const uint32_t* stream;      /* explicit choice words, or NULL */
  const int64_t* stream_off;   /* [G+1] words of graph g are stream[stream_off[g] .. stream_off[g+1]) */
} GtokChoice;

GTOK_API int gtok_version(void);
GTOK_API const char* gtok_last_error(void);
/* HOST: number of SMs of the current device, or a negative GTOK_E*. */
GTOK_API int gtok_device_sm_count(void);

/* Ingest  compact host->device wire format.  The reference holds graphs as int64 PyG tensors on the host
 *      (zinc_dataset_autograph.py:51-73, graph_token_dataset_autograph.py:378-394); the engine's resident
 *      layout is the int32 GtokGraphBatch above.  The host side ships each array in the narrowest lossless
 *      unsigned type (local node ids, atom/bond types and per-graph counts of molecule-sized graphs fit 8/16
 *      bits) and these two calls restore the resident layout on the device:
 *      gtok_widen_i32: in[n] of elem_bytes (1 = uint8, 2 = uint16, 4 = int32 copy) -> out int32[n];
 *      gtok_counts_to_offsets: per-graph counts -> int64[G+1] exclusive offsets (tmp_i32 int32[G] scratch,
 *      ws as for gtok_scan_offsets).  All buffers 16-byte aligned. */
GTOK_API int gtok_widen_i32(const void* in, int32_t elem_bytes, int64_t n, int32_t* out, void* stream);
GTOK_API int gtok_counts_to_offsets(const void* counts, int32_t elem_bytes, int64_t num_items, int64_t* off_out,
                                    int32_t* tmp_i32, void* ws, size_t ws_bytes, void* stream);

/* Ingest  the reference's OWN storage -> resident layouts, on the device.  PyG keeps a dataset as int64 tensors on
 *      the host (InMemoryDataset: edge_index int64[2, E_total] with LOCAL node ids, x int64[N_total(, 1)], edge_attr
 *      int64[E_total], slices int64[G+1]; zinc_dataset_autograph.py:51-73, zinc_dataset_indexbase.py:63-70).  The
 *      slices ARE node_off / edge_off; once the bytes are in HBM these calls narrow the rest, so an end-to-end step
 *      that starts from the reference's tensors needs no host-side conversion:
 *      gtok_pyg_pack_mol: src64 / dst64 (rows 0 / 1 of edge_index) and attr64 (may be NULL) -> col16[num_entries]
 *      (GtokPackedMolBatch word), x64 -> x8[num_nodes]; gtok_pyg_narrow: int64[n] -> int32[n] (any graph size).
 *      A value that does not fit its field is never silently truncated: the matching GTOK_PYG_* bit is OR-ed into
 *      status[0] (device int32, zeroed by the caller).  Buffers 16-byte aligned; any sub-range of the arrays whose
 *      start is a multiple of 16 elements is a valid call (chunks of a copy in flight). */
enum { GTOK_PYG_BAD_NODE_ID = 1,    /* a local node id outside 0..63 */
       GTOK_PYG_BAD_EDGE_ATTR = 2,  /* an edge_attr outside 0..15 */
       GTOK_PYG_BAD_NODE_TYPE = 4,  /* an atom type outside 0..255 */
       GTOK_PYG_BAD_RANGE = 8       /* a value outside int32 (gtok_pyg_narrow) */ };
GTOK_API int gtok_pyg_pack_mol(const int64_t* src64, const int64_t* dst64, const int64_t* attr64, int64_t num_entries,
                               const int64_t* x64, int64_t num_nodes, uint16_t* col16, uint8_t* x8, int32_t* status,
                               void* stream);
GTOK_API int gtok_pyg_narrow(const int64_t* in64, int64_t n, int32_t* out, int32_t* status, void* stream);

/* Ingest  graph-token TEXTS -> graph arrays on the GPU: the inverse of the external serializer that the reference
 *      runs per text on the host (graph_token_dataset_autograph.py:14-54 parse_graph_from_text, data_loader.py:19-74
 *      query / label parsers) and that the drop-in TokenDataset needs before any tokenizer can run on plain
 *      strings.  text = the texts' ASCII bytes back to back, text_off int64[G+1].
 *      gtok_text_scan: num_nodes / num_edges int32[G], task int32[G] (GTOK_TASK_*, or -1 when the text is not
 *      of the grammar), query_u / query_v int32[G] (may be NULL).  The caller scans num_edges into edge_off
 *      (gtok_scan_offsets), then gtok_text_edges writes src / dst int32[E] and status int32[G] (1 = every edge
 *      token is a decimal, every third one is <e>, the node list is 0..N-1; 0 otherwise). */
GTOK_API int gtok_text_scan(const uint8_t* text, const int64_t* text_off, int64_t num_texts, int32_t* num_nodes,
                            int32_t* num_edges, int32_t* task, int32_t* query_u, int32_t* query_v, void* stream);
GTOK_API int gtok_text_edges(const uint8_t* text, const int64_t* text_off, int64_t num_texts, const int32_t* num_nodes,
                             const int32_t* task, const int64_t* edge_off, int32_t* src, int32_t* dst, int32_t* status,
                             void* stream);

/* Render  ids -> text bytes, the inverse of the text ingest: ZINCTokenizationDataset items carry their TEXT because the
 *      unchanged trainer reads it to collect its vocabulary (trainer/train_ibtt.py:231-233, :364-369;
 *      zinc_dataset_indexbase.py:143-227 renders it per molecule with .item() loops).  words = the vocabulary's strings
 *      back to back (ASCII), word_off int32[num_words+1].  gtok_detok_lengths: chars_out[i] = len(word(tok[i])) + 1; the
 *      caller scans them into char_off (gtok_scan_offsets); gtok_detok_render writes word(tok[i]) + ' ' at
 *      text_out + char_off[i].  gtok_first_occurrence: first_out[id] = smallest i with tok[i] == id (UINT64_MAX = never),
 *      num_ids <= 4096 - the first-seen order of the dynamic vocabulary tokens (zinc_vocab.py:154-179). */
GTOK_API int gtok_detok_lengths(const int64_t* tok, int64_t num_tokens, const int32_t* word_off, int32_t num_words,
                                int32_t* chars_out, void* stream);
GTOK_API int gtok_detok_render(const int64_t* tok, int64_t num_tokens, const uint8_t* words, const int32_t* word_off,
                               int32_t num_words, const int64_t* char_off, uint8_t* text_out, void* stream);
GTOK_API int gtok_first_occurrence(const int64_t* tok, int64_t num_tokens, int32_t num_ids, uint64_t* first_out, void* stream);

/* K5   token statistics for build_vocab_from_texts (data_loader.py:451-463) computed from the graphs instead
 *      of from their serialized texts: for every node index k < num_bins, counts[k] = occurrences of the
 *      string str(k) over all graph-token texts (edge endpoints + node list + query nodes) and first[k] =
 *      (graph << 32 | token position) of its first occurrence (UINT64_MAX if it never occurs) - the key that
 *      reproduces Counter.most_common()'s first-seen tie-break.  The remaining tokens (<e>, TASK words,
 *      labels) have closed-form counts; the host finishes the vocabulary on <= num_bins + a few entries. */
GTOK_API int gtok_vocab_histogram(const GtokGraphBatch* batch, int32_t task, int32_t num_bins, uint64_t* counts,
                                  uint64_t* first, void* stream);

/* K1a  lengths of TokenDataset sequences for graph-token graphs:
 *      min(max_len, 3E + N + 5|7)  (data_loader.py:471-484 applied to the
 *      serializer grammar).  len_out int32[G]. */
GTOK_API int gtok_ibtt_lengths(const GtokGraphBatch* batch, const GtokIbttRules* rules,
                      int32_t* len_out, void* stream);

/* K0   exclusive scan of per-graph lengths -> ragged offsets (implicit in the
 *      reference's Python list building).  off_out int64[G+1].  ws must hold
 *      gtok_scan_workspace_bytes(G) bytes. */
GTOK_API size_t gtok_scan_workspace_bytes(int64_t num_items);
GTOK_API int gtok_scan_offsets(const int32_t* len, int64_t num_items, int64_t* off_out,
                      void* ws, size_t ws_bytes, void* stream);

/* K1b  graph-token serializer + TokenDataset.__init__ id lookup, strip after
 *      <p>, truncation (external graph-token grammar + data_loader.py:465-486).
 *      tok_out int64[off[G]], row g at tok_out[off[g] .. off[g+1]). */
GTOK_API int gtok_ibtt_serialize(const GtokGraphBatch* batch, const GtokIbttRules* rules,
                        const int64_t* off, int64_t* tok_out, void* stream);

/* K1+K4 fused  the graph-token serializer + TokenDataset + collate in one pass (external grammar +
 *      data_loader.py:465-497): row r of the epoch is graph batch_idx[r] (or r when batch_idx is NULL - this
 *      is the DataLoader's per-epoch shuffle), batch k = rows [k*B, min(num_rows,(k+1)*B)), written row-major
 *      [rows_k, lmax[k]] at X + xoff[k] / attn + xoff[k] with pad_id beyond each row's length.  lmax / xoff
 *      come from gtok_ibtt_lengths -> gtok_batch_extents -> gtok_scan_offsets (closed-form lengths: no token
 *      is generated twice and the ragged stream never exists in HBM). */
GTOK_API int gtok_ibtt_collate(const GtokGraphBatch* batch, const GtokIbttRules* rules, const int64_t* batch_idx,
                               int64_t num_rows, int64_t batch_size, const int32_t* lmax, const int64_t* xoff,
                               int64_t pad_id, int64_t* X, uint8_t* attn, void* stream);

/* K2a/K2b  ZINCTokenizationDataset.tokenize_molecule + __getitem__ truncation +
 *      TokenDataset (zinc_dataset_indexbase.py:143-227, data_loader.py:465-486). */
GTOK_API int gtok_zinc_lengths(const GtokGraphBatch* batch, const GtokZincVocab* vocab,
                      int32_t* len_out, void* stream);
GTOK_API int gtok_zinc_serialize(const GtokGraphBatch* batch, const GtokZincVocab* vocab,
                        const int64_t* off, int64_t* tok_out, void* stream);

/* K2 fused  the same two reference functions in ONE pass over HBM: lengths, ragged
 *      offsets (int64[G+1]) and rows are produced together (CTA tiles of molecules
 *      staged in shared memory, one thread per molecule for the first-seen dedupe,
 *      decoupled look-back for the offsets).  Molecule
 *      shapes only: batch.max_nodes <= 64 and batch.max_entries <= 256, otherwise
 *      GTOK_EUNSUPPORTED (use the three calls above).  Rows that would end beyond
 *      tok_capacity are not written (lengths/offsets still are), so the caller can
 *      size tok_out from the bound 2N+4E+4G and check off_out[G] afterwards. */
GTOK_API size_t gtok_zinc_tokenize_workspace_bytes(int64_t num_graphs);
GTOK_API int gtok_zinc_tokenize(const GtokGraphBatch* batch, const GtokZincVocab* vocab, int32_t* len_out,
                                int64_t* off_out, int64_t* tok_out, int64_t tok_capacity,
                                void* ws, size_t ws_bytes, void* stream);

/* K2+K4 fused  tokenize_molecule + TokenDataset + collate in ONE pass (zinc_dataset_indexbase.py:143-227,
 *      data_loader.py:465-497): molecules -> padded batches; the ragged token stream never exists in HBM.
 *      Batch k = molecules [k*B, min(G,(k+1)*B)) in batch order, padded to ITS OWN Lmax_k with pad_id,
 *      written row-major [rows_k, Lmax_k] at X + xoff[k] / attn + xoff[k] - the layout
 *      gtok_pad_mask_batched produces.  Outputs: len_out int32[G], lmax_out int32[nb], xoff_out int64[nb+1],
 *      X int64 / attn uint8 (torch.bool storage) with x_capacity cells each; a batch that would end beyond
 *      x_capacity is not written (its extents still are), so the caller can size X from the bound
 *      G * min(max_len, 2*max_nodes + 4*max_entries + 4) and check xoff_out[nb] afterwards.
 *      B must be a multiple of 32 in [32, 256]; molecule shapes as for gtok_zinc_tokenize. */
GTOK_API size_t gtok_zinc_collate_workspace_bytes(int64_t num_graphs, int64_t batch_size);
GTOK_API int gtok_zinc_collate(const GtokGraphBatch* batch, const GtokZincVocab* vocab, int64_t batch_size,
                               int64_t pad_id, int32_t* len_out, int32_t* lmax_out, int64_t* xoff_out, int64_t* X,
                               uint8_t* attn, int64_t x_capacity, void* ws, size_t ws_bytes, void* stream);

/* Compact resident / wire layout of molecule-sized graphs: one 16-bit word per edge_index column
 *      (src | dst << 6 | edge_attr << 12: local ids < 64, edge_attr in 0..15; caller's column order) and one byte
 *      per atom type.  0.12 KB per ZINC molecule instead of 0.68 KB in the int32 layout (1.3 KB as the reference's
 *      int64 tensors), lossless for those ranges; it is also exactly what the thread-per-molecule kernel stages, so
 *      gtok_zinc_collate_packed reads it with plain 16-byte copies.  Same outputs as gtok_zinc_collate. */
typedef struct GtokPackedMolBatch {
  int64_t num_graphs;
  const int64_t* node_off;   /* [G+1] (gtok_counts_to_offsets from per-graph counts) */
  const int64_t* edge_off;   /* [G+1] */
  const uint16_t* col16;     /* [E] */
  const uint8_t* x8;         /* [N] */
  int32_t max_nodes;         /* bound on atoms per molecule AND on every local id + 1 (<= 64) */
  int32_t max_entries;       /* bound on columns per molecule (<= 256) */
  int32_t unique_columns;    /* 1: the columns already ARE the first-seen undirected edges, in order (what remains of
                                a canonical PyG edge_index - both directions, sorted by (src, dst), symmetric
                                edge_attr - when only its src < dst half is kept; gtok/batch.py verifies that before
                                it drops the other half).  The kernel then skips the dedupe.  0: any column list */
  int32_t max_nodes_per_32;  /* as in GtokGraphBatch (0 = unknown) */
  int32_t max_entries_per_32;
  int32_t reserved_;
} GtokPackedMolBatch;
GTOK_API int gtok_zinc_collate_packed(const GtokPackedMolBatch* batch, const GtokZincVocab* vocab, int64_t batch_size,
                                      int64_t pad_id, int32_t* len_out, int32_t* lmax_out, int64_t* xoff_out, int64_t* X,
                                      uint8_t* attn, int64_t x_capacity, void* ws, size_t ws_bytes, void* stream);

/* K7   ingest for K3: the canonical undirected CSR of GtokCsrBatch (+ its adjacency bit rows) from the caller's
 *      edge_index columns, graphs of at most 64 * adj_words <= 256 nodes - what to_undirected + coalesce do inside
 *      AutoGraph's tokenizer on every call (call site trainer/train_agtt.py:250), once per dataset here.
 *      gtok_csr_adjacency zeroes and fills adj64 [total_nodes * adj_words] (self loops dropped, repeats collapse) and
 *      writes the node degrees; the caller scans them into rowptr64 [total_nodes + 1] (gtok_scan_offsets) and
 *      gtok_csr_fill writes rowptr_out int32 [total_nodes + 1], col_out [nnz] (rows ascending) and, when eattr_out is
 *      given, the label of every entry = edge_attr of the FIRST column of its pair in caller order, either direction
 *      (first_col_ws: nnz * 8 bytes of scratch).  batch->max_nodes must be set. */
GTOK_API int gtok_csr_adjacency(const GtokGraphBatch* batch, int32_t adj_words, int64_t total_nodes, uint64_t* adj64,
                                int32_t* degree_out, void* stream);
GTOK_API int gtok_csr_fill(const GtokGraphBatch* batch, int32_t adj_words, int64_t total_nodes, const uint64_t* adj64,
                           const int64_t* rowptr64, int64_t nnz, int32_t* rowptr_out, int32_t* col_out,
                           int32_t* eattr_out, uint64_t* first_col_ws, void* stream);

/* K7 (any size)  the same canonical CSR for graphs beyond 256 nodes (up to the 32767 the SENT kernels accept), library
 *      kernels only: gtok_csrl_count writes degree upper bounds (repeated columns counted) -> the caller scans them
 *      into rowptr_ub int64[total_nodes+1]; gtok_csrl_sort scatters key = neighbour << 32 | column index into both
 *      endpoint rows of entries[rowptr_ub[total_nodes]], sorts every row and keeps the first key per neighbour (warp
 *      bitonic network <= 32 entries, CTA shared-memory bitonic <= 4096, neighbour-indexed table beyond: table_ws of
 *      gtok_csrl_table_bytes(max_nodes) bytes) and writes the unique degrees -> the caller scans them into rowptr64;
 *      gtok_csrl_emit writes rowptr_out int32[total_nodes+1], col_out[nnz] and eattr_out[nnz] (label = edge_attr of the
 *      pair's first column in caller order).  cursor_ws: total_nodes int32.  batch->max_nodes must be set. */
GTOK_API int gtok_csrl_count(const GtokGraphBatch* batch, int64_t total_nodes, int32_t* degree_ub, void* stream);
GTOK_API size_t gtok_csrl_table_bytes(int32_t max_nodes);
GTOK_API int gtok_csrl_sort(const GtokGraphBatch* batch, int64_t total_nodes, const int64_t* rowptr_ub, int32_t* cursor_ws,
                            uint64_t* entries, int32_t* degree_out, void* table_ws, size_t table_bytes, void* stream);
GTOK_API int gtok_csrl_emit(const GtokGraphBatch* batch, int64_t total_nodes, const int64_t* rowptr_ub, const uint64_t* entries,
                            const int64_t* rowptr64, int64_t nnz, int32_t* rowptr_out, int32_t* col_out, int32_t* eattr_out,
                            void* stream);

/* K3   Graph2TrailTokenizer.__call__ (external AutoGraph; spec in
 *      oracle/sent_oracle.py) + remap_zinc_tokens + query append
 *      (train_agtt.py:171-273).  Row g is written at tok_out[g*stride ..], its
 *      length to len_out[g]; stride >= max_length (+3 with append_query). */
GTOK_API int gtok_sent_tokenize(const GtokCsrBatch* csr, const GtokSentCfg* cfg, const GtokChoice* choice,
                       int32_t* len_out, int64_t* tok_out, int64_t stride, void* stream);
/*      Same, with a scratch workspace for graphs above 256 nodes (8 bytes per node of the batch: one word of walk
 *      state per node lives there instead of in shared memory, which lets 3-5x more graphs walk concurrently per
 *      SM).  total_nodes = node_off[num_graphs]; ws may be NULL (then, and for smaller graphs, this is
 *      gtok_sent_tokenize).  The workspace carries no state between calls.  Measured on B200 (32M nodes per call,
 *      mean degree 8): the shared-memory kernel is faster up to ~2048 nodes per graph (14.6 vs 15.8 ms), the
 *      workspace kernel beyond (18.2 vs 28.7 ms at 4096 nodes) - gtok/engine.py picks by that rule. */
/*      Graphs of at most 256 nodes: the workspace (12 bytes per node of the batch) is the walk's ITEM LOG - the walk kernel
 *      (thread per graph) records one 8-byte item per first visit and the emission kernel (warp per row) expands the
 *      items into tokens with full-sector stores; without it the one-kernel paths run (measured on B200 per 1M graphs:
 *      ZINC-shaped labeled 1.17 -> 0.85 ms, ER 1.22 -> 1.16 ms; per 32M nodes at 256 nodes per graph 8.5 -> 3.6 ms). */
GTOK_API size_t gtok_sent_workspace_bytes(int64_t total_nodes, int32_t max_nodes);
GTOK_API int gtok_sent_tokenize_ws(const GtokCsrBatch* csr, const GtokSentCfg* cfg, const GtokChoice* choice,
                                   int32_t* len_out, int64_t* tok_out, int64_t stride, int64_t total_nodes, void* ws,
                                   size_t ws_bytes, void* stream);

/* K3+K4 fused  TokenizedGraphDataset.__getitem__ + collate_fn of a whole epoch (train_agtt.py:246-302) for graphs of at
 *      most 64 nodes, without the int64 rows at stride max_length in between:
 *        gtok_sent_log_bytes   bytes of item log the configuration needs (0 = not supported by these kernels: use
 *                              gtok_sent_tokenize + gtok_pad_mask_batched);
 *        gtok_sent_walk        the walk: item log + row lengths len_out int32[G] (query included);
 *        gtok_batch_extents(len_out, batch_idx, ...) + gtok_scan_offsets give lmax / xoff as for K4b;
 *        gtok_sent_emit_collate  row r of the epoch (graph batch_idx[r], or r) is expanded from the log straight into
 *                              X / attn + xoff[r / B] + (r % B) * lmax[r / B], padded with pad_id (K4b's layout).
 *      Output identical to gtok_sent_tokenize followed by gtok_pad_mask_batched. */
GTOK_API size_t gtok_sent_log_bytes(const GtokCsrBatch* csr, const GtokSentCfg* cfg, int64_t total_nodes);
GTOK_API int gtok_sent_walk(const GtokCsrBatch* csr, const GtokSentCfg* cfg, const GtokChoice* choice, int64_t total_nodes,
                            void* log, size_t log_bytes, int32_t* len_out, void* stream);
GTOK_API int gtok_sent_emit_collate(const GtokCsrBatch* csr, const GtokSentCfg* cfg, int64_t total_nodes, const void* log,
                                    const int32_t* row_len, const int64_t* batch_idx, int64_t num_rows, int64_t B,
                                    const int32_t* lmax, const int64_t* xoff, int64_t pad_id, int64_t* X, uint8_t* attn,
                                    void* stream);

/* K4   collate / collate_fn (data_loader.py:488-497, train_agtt.py:276-302):
 *      X[b, :len] = row, X[b, len:] = pad_id, attn[b, :len] = 1.  Row b is
 *      tok[row_start[i] .. row_start[i] + row_len[i]) with i = batch_idx ? batch_idx[b] : b.
 *      X int64[B, Lmax], attn uint8[B, Lmax] (torch.bool storage). */
GTOK_API int gtok_pad_mask(const int64_t* tok, const int64_t* row_start, const int32_t* row_len,
                  const int64_t* batch_idx, int64_t B, int64_t Lmax, int64_t pad_id,
                  int64_t* X, uint8_t* attn, void* stream);

/* K4b  a whole epoch of batches in one launch (same reference functions, called
 *      once per DataLoader step there): rows are grouped B at a time in
 *      batch_idx order (batch k = rows [k*B, min(G,(k+1)*B))); batch k is padded
 *      to its OWN Lmax_k = max row_len in the batch, exactly as collate does.
 *      gtok_batch_extents writes lmax int32[nb] and cells int32[nb]
 *      (= rows_k * Lmax_k, which must stay below 2^31 per batch: the kernel traps otherwise; the caller scans cells into xoff int64[nb+1] with
 *      gtok_scan_offsets); gtok_pad_mask_batched then writes batch k at
 *      X + xoff[k] / attn + xoff[k], row-major [rows_k, Lmax_k]. nb = ceil(G/B). */
GTOK_API int gtok_batch_extents(const int32_t* row_len, const int64_t* batch_idx, int64_t G, int64_t B,
                       int32_t* lmax_out, int32_t* cells_out, void* stream);
GTOK_API int gtok_pad_mask_batched(const int64_t* tok, const int64_t* row_start, const int32_t* row_len,
                          const int64_t* batch_idx, int64_t G, int64_t B, const int32_t* lmax,
                          const int64_t* xoff, int64_t pad_id, int64_t* X, uint8_t* attn, void* stream);

/* K2+K4 for ANY batch size / molecule size (the single-pass gtok_zinc_collate covers batches of 32..256 molecule-sized
 *      graphs): gtok_zinc_lengths -> gtok_batch_extents -> gtok_scan_offsets give lmax / xoff; gtok_pad_tail_batched writes
 *      everything of the padded batches that is not a token (pad cells [len, lmax) of every row, the mask rows) and
 *      row_off_out int64[G+1] = first cell of every row; gtok_zinc_serialize_rows then writes row g's tokens at
 *      tok_out + row_off[g] (row_len[g] of them).  9 * Lmax bytes per row, written once; no ragged stream in between. */
GTOK_API int gtok_pad_tail_batched(const int32_t* row_len, int64_t G, int64_t B, const int32_t* lmax, const int64_t* xoff,
                                   int64_t pad_id, int64_t* X, uint8_t* attn, int64_t* row_off_out, void* stream);
GTOK_API int gtok_zinc_serialize_rows(const GtokGraphBatch* batch, const GtokZincVocab* vocab, const int64_t* row_off,
                                      const int32_t* row_len, int64_t* tok_out, void* stream);

/* Plumbing kernels (no torch op computes anything on the product path):
 *      gtok_compact_rows: strided rows [num_rows, stride] + row_len -> ragged stream, row r at out[off[r] ..]
 *      (off = gtok_scan_offsets(row_len)); the AGTT shim serves Graph2TrailTokenizer.__call__ results from it
 *      (trainer/train_agtt.py:250).
 *      gtok_segment_counts / gtok_gather_segments: gather the graphs idx[0..num_idx) of a batch (any order, repeats
 *      allowed) into a new batch - TokenDataset over a shuffled / subset view of a PyG dataset (ds._indices).
 *      cnt_out[i] = off[idx[i]+1] - off[idx[i]] (status |= 1 on an index outside [0, num_graphs)); the caller scans
 *      cnt_out into dst_off and gathers every int32 array with gtok_gather_segments.
 *      gtok_balance_order: order_out = graphs sorted by (nodes, CSR entries) inside windows of `window` (2..4096)
 *      consecutive graphs (GtokCsrBatch.order: the 32 lanes of a warp then walk graphs of similar work). */
GTOK_API int gtok_compact_rows(const int64_t* tok, int64_t stride, const int32_t* row_len, const int64_t* off,
                               int64_t num_rows, int64_t* out, void* stream);
GTOK_API int gtok_segment_counts(const int64_t* off, int64_t num_graphs, const int64_t* idx, int64_t num_idx,
                                 int32_t* cnt_out, int32_t* status, void* stream);
GTOK_API int gtok_gather_segments(const int32_t* src, const int64_t* src_off, int64_t num_graphs, const int64_t* idx,
                                  const int64_t* dst_off, int64_t num_idx, int32_t* dst, void* stream);
/*      gtok_ragged_append_lengths / gtok_ragged_append: TokenDataset(strip_label=False) (data_loader.py:478-482 without
 *      the strip): row r = (tok[off[r] .. +row_len[r]) ++ tail[r, :tail_len[r]])[:max_len]; the tails are the ids of the
 *      few tokens after <p> (label, <eos>).  new_len -> gtok_scan_offsets -> new_off -> out. */
GTOK_API int gtok_ragged_append_lengths(const int32_t* row_len, const int32_t* tail_len, int64_t num_rows, int32_t max_len,
                                        int32_t* new_len, void* stream);
GTOK_API int gtok_ragged_append(const int64_t* tok, const int64_t* off, const int32_t* row_len, const int64_t* tail,
                                int32_t tail_stride, const int64_t* new_off, int64_t num_rows, int64_t* out, void* stream);
GTOK_API int gtok_balance_order(const int64_t* node_off, const int32_t* rowptr, int64_t num_graphs, int32_t window,
                                int32_t* order_out, void* stream);

/* K6   the steps around the path: consumer-side query positions, task labels, and an
 *      on-device Erdos-Renyi generator for fully device-side benchmarks.
 *
 *      gtok_find_token: extract_query_nodes (trainer/train_ibtt.py:89-103, trainer/train_agtt.py:101-112) scans
 *      every row of a batch on the host for the FIRST position of the <q> id.  Here one launch does it for all
 *      rows of a padded epoch (layout of gtok_pad_mask_batched: batch k at X + xoff[k], [rows_k, lmax[k]]):
 *      pos_out[r] = first p with X[r, p] == t, else -1; t = token_per_batch[r / B] when given (AGTT: the id is
 *      idx_offset + num_nodes of the batch's first graph, train_agtt.py:131), else `token`. */
GTOK_API int gtok_find_token(const int64_t* X, const int64_t* xoff, const int32_t* lmax, int64_t num_rows,
                             int64_t batch_size, const int64_t* token_per_batch, int64_t token, int32_t* pos_out,
                             void* stream);

/*      gtok_graph_labels: the two graph-token task labels straight from the canonical CSR.
 *      has_cycle_out[g] (0/1) = the reference's DFS label oracle (test/train_performance_test.py:27-67) on the
 *      edge list: true iff M_g > N_g - C_g with C_g connected components and M_g listed edges = edge_off[g+1] -
 *      edge_off[g] (every extra list entry - a chord, a repeated edge, a self loop - closes a cycle for that
 *      DFS); edge_off NULL: M_g = CSR entries / 2 (simple graph).  dist_out[g] = BFS hops query_u -> query_v,
 *      -1 when unreachable / out of range (the generator's `lenK` / `INF` label, data_loader.py:19-41).
 *      Either output may be NULL; dist needs csr->query_u/query_v. */
GTOK_API int gtok_graph_labels(const GtokCsrBatch* csr, const int64_t* edge_off, int32_t* has_cycle_out,
                               int32_t* dist_out, void* stream);

/*      On-device graphs of the docs' generator families er / sbm / complete / star / path with the docs' size /
 *      sparsity distribution (docs/synthetic_data.md:93-94,130-138; ba too, up to 64 nodes; sfn is not provided);
 *      the stream definition is oracle/tasks_oracle.py er_graph (Philox-4x32-10 keyed by seed, counter = global
 *      graph id, so output is invariant to sharding).  gtok_er_sizes writes per-graph node / edge counts and a
 *      query pair; the caller scans the counts (gtok_counts_to_offsets) and gtok_er_edges fills src / dst
 *      (each edge once, u < v ascending - the order graph-token texts list them in). */
enum { GTOK_GEN_ER = 0,        /* every pair with probability p */
       GTOK_GEN_SBM = 1,       /* stochastic block model: p inside a block, p_out across */
       GTOK_GEN_COMPLETE = 2, GTOK_GEN_STAR = 3 /* centre 0 */, GTOK_GEN_PATH = 4 /* i - (i+1) */,
       GTOK_GEN_BA = 5,        /* Barabasi-Albert preferential attachment, m ~ U{1..num_blocks} new edges per node, N <= 64 */
       GTOK_GEN_SFN = 6        /* Holme-Kim scale-free graphs (docs "sfn"): BA + triad steps with probability
                                  out_ratio_q16 / 65536, N <= 64 */ };
typedef struct GtokErSpec {
  int32_t num_buckets;       /* 1..4 size buckets, one drawn uniformly per graph */
  int32_t n_lo[4], n_hi[4];  /* N uniform in [n_lo, n_hi), 1 <= n_lo < n_hi <= 32768 */
  uint32_t thr_lo, thr_hi;   /* edge probability p ~ U(thr_lo, thr_hi) / 2^32 */
  int32_t kind;              /* GTOK_GEN_*: which of the docs' generator families (docs/synthetic_data.md:93-94) */
  int32_t num_blocks;        /* GTOK_GEN_SBM: 1..16 contiguous blocks, block(i) = i * num_blocks / N;
                                GTOK_GEN_BA / GTOK_GEN_SFN: largest attachment count m (1..8) */
  uint32_t out_ratio_q16;    /* GTOK_GEN_SBM: p_out = p * out_ratio_q16 / 65536 between different blocks;
                                GTOK_GEN_SFN: probability of a triad step, in 1/65536 */
  uint64_t seed;
  int64_t graph_id_base;
} GtokErSpec;
GTOK_API int gtok_er_sizes(const GtokErSpec* spec, int64_t num_graphs, int32_t* num_nodes_out, int32_t* num_edges_out,
                           int32_t* query_u_out, int32_t* query_v_out, void* stream);
GTOK_API int gtok_er_edges(const GtokErSpec* spec, int64_t num_graphs, const int64_t* edge_off, int32_t* src_out,
                           int32_t* dst_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GTOK_H_ */
