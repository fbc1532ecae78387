"""ctypes binding of include/gtok.h (the C-ABI of libgtok.so).

This file is the "reference-side binding a maintainer would add" (INTEGRATION.md):
plain ctypes, no torch types cross the boundary - only raw device pointers,
sizes and a stream handle.  There is NO fallback: a missing library raises at
import time, a missing GPU makes every compute call raise GtokError(GTOK_ECUDA).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GTOK_LIB_PATH") or os.path.join(_HERE, "libgtok.so")   # (override: A/B builds of the same ABI)

GTOK_OK, GTOK_EINVAL, GTOK_ECUDA, GTOK_EUNSUPPORTED, GTOK_EWORKSPACE = 0, -1, -2, -3, -4
TASK_CYCLE_CHECK, TASK_SHORTEST_PATH = 0, 1
PYG_BAD_NODE_ID, PYG_BAD_EDGE_ATTR, PYG_BAD_NODE_TYPE, PYG_BAD_RANGE = 1, 2, 4, 8

i32, i64, u64, vp = C.c_int32, C.c_int64, C.c_uint64, C.c_void_p


class GtokError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libgtok error {code}: {msg}")
        self.code = code


class GraphBatchC(C.Structure):
    _fields_ = [("num_graphs", i64), ("node_off", vp), ("edge_off", vp), ("src", vp), ("dst", vp),
                ("x", vp), ("edge_attr", vp), ("query_u", vp), ("query_v", vp),
                ("max_nodes", i32), ("max_entries", i32), ("max_nodes_per_32", i32), ("max_entries_per_32", i32)]


class PackedMolBatchC(C.Structure):
    _fields_ = [("num_graphs", i64), ("node_off", vp), ("edge_off", vp), ("col16", vp), ("x8", vp),
                ("max_nodes", i32), ("max_entries", i32), ("unique_columns", i32), ("max_nodes_per_32", i32),
                ("max_entries_per_32", i32), ("reserved_", i32)]


class IbttRulesC(C.Structure):
    _fields_ = [("task", i32), ("max_len", i32), ("tok_bos", i32), ("tok_e", i32), ("tok_n", i32),
                ("tok_q", i32), ("tok_p", i32), ("tok_task", i32), ("oov", i32), ("node_tok_len", i32),
                ("node_tok", vp), ("max_token_id", i32), ("reserved_", i32)]


class ZincVocabC(C.Structure):
    _fields_ = [("item_max_len", i32), ("ds_max_len", i32), ("tok_bos", i32), ("tok_eos", i32),
                ("tok_atom", i32), ("tok_bond", i32), ("tok_q", i32), ("tok_regression", i32), ("tok_p", i32),
                ("atom_tok", i32 * 9), ("atom_oov", i32), ("bond_tok", i32 * 4), ("bond_oov", i32),
                ("oov", i32), ("node_tok_len", i32), ("node_tok", vp), ("max_token_id", i32), ("reserved_", i32)]


class CsrBatchC(C.Structure):
    _fields_ = [("num_graphs", i64), ("node_off", vp), ("rowptr", vp), ("col", vp), ("eattr", vp), ("x", vp),
                ("query_u", vp), ("query_v", vp), ("max_nodes", i32), ("max_nnz", i32), ("max_label", i32), ("adj_words", i32), ("adj64", vp), ("order", vp)]


class SentCfgC(C.Structure):
    _fields_ = [("idx_offset", i32), ("node_idx_offset", i32), ("edge_idx_offset", i32), ("max_length", i32),
                ("labeled", i32), ("edge_label_shift", i32), ("remap_zinc", i32), ("append_query", i32),
                ("restart_policy", i32), ("bracket_order", i32), ("no_eos", i32), ("reserved0", i32)]


class ErSpecC(C.Structure):
    _fields_ = [("num_buckets", i32), ("n_lo", i32 * 4), ("n_hi", i32 * 4), ("thr_lo", C.c_uint32), ("thr_hi", C.c_uint32),
                ("kind", i32), ("num_blocks", i32), ("out_ratio_q16", C.c_uint32), ("seed", u64), ("graph_id_base", i64)]


class ChoiceC(C.Structure):
    _fields_ = [("seed", u64), ("graph_id_base", i64), ("stream", vp), ("stream_off", vp)]


# name -> (restype, argtypes); must list EVERY prototype of include/gtok.h
# (tests/test_abi.py parses the header and checks both directions).
PROTOTYPES = {
    "gtok_version": (C.c_int, []),
    "gtok_last_error": (C.c_char_p, []),
    "gtok_device_sm_count": (C.c_int, []),
    "gtok_widen_i32": (C.c_int, [vp, i32, i64, vp, vp]),
    "gtok_counts_to_offsets": (C.c_int, [vp, i32, i64, vp, vp, vp, C.c_size_t, vp]),
    "gtok_pyg_pack_mol": (C.c_int, [vp, vp, vp, i64, vp, i64, vp, vp, vp, vp]),
    "gtok_pyg_narrow": (C.c_int, [vp, i64, vp, vp, vp]),
    "gtok_text_scan": (C.c_int, [vp, vp, i64, vp, vp, vp, vp, vp, vp]),
    "gtok_text_edges": (C.c_int, [vp, vp, i64, vp, vp, vp, vp, vp, vp, vp]),
    "gtok_detok_lengths": (C.c_int, [vp, i64, vp, i32, vp, vp]),
    "gtok_detok_render": (C.c_int, [vp, i64, vp, vp, i32, vp, vp, vp]),
    "gtok_first_occurrence": (C.c_int, [vp, i64, i32, vp, vp]),
    "gtok_vocab_histogram": (C.c_int, [C.POINTER(GraphBatchC), i32, i32, vp, vp, vp]),
    "gtok_ibtt_lengths": (C.c_int, [C.POINTER(GraphBatchC), C.POINTER(IbttRulesC), vp, vp]),
    "gtok_scan_workspace_bytes": (C.c_size_t, [i64]),
    "gtok_scan_offsets": (C.c_int, [vp, i64, vp, vp, C.c_size_t, vp]),
    "gtok_ibtt_serialize": (C.c_int, [C.POINTER(GraphBatchC), C.POINTER(IbttRulesC), vp, vp, vp]),
    "gtok_ibtt_collate": (C.c_int, [C.POINTER(GraphBatchC), C.POINTER(IbttRulesC), vp, i64, i64, vp, vp, i64, vp, vp, vp]),
    "gtok_zinc_lengths": (C.c_int, [C.POINTER(GraphBatchC), C.POINTER(ZincVocabC), vp, vp]),
    "gtok_zinc_serialize": (C.c_int, [C.POINTER(GraphBatchC), C.POINTER(ZincVocabC), vp, vp, vp]),
    "gtok_zinc_tokenize_workspace_bytes": (C.c_size_t, [i64]),
    "gtok_zinc_tokenize": (C.c_int, [C.POINTER(GraphBatchC), C.POINTER(ZincVocabC), vp, vp, vp, i64, vp, C.c_size_t, vp]),
    "gtok_zinc_collate_workspace_bytes": (C.c_size_t, [i64, i64]),
    "gtok_zinc_collate": (C.c_int, [C.POINTER(GraphBatchC), C.POINTER(ZincVocabC), i64, i64, vp, vp, vp, vp, vp, i64,
                                    vp, C.c_size_t, vp]),
    "gtok_zinc_collate_packed": (C.c_int, [C.POINTER(PackedMolBatchC), C.POINTER(ZincVocabC), i64, i64, vp, vp, vp, vp, vp, i64,
                                           vp, C.c_size_t, vp]),
    "gtok_sent_tokenize": (C.c_int, [C.POINTER(CsrBatchC), C.POINTER(SentCfgC), C.POINTER(ChoiceC), vp, vp, i64, vp]),
    "gtok_sent_workspace_bytes": (C.c_size_t, [i64, i32]),
    "gtok_sent_tokenize_ws": (C.c_int, [C.POINTER(CsrBatchC), C.POINTER(SentCfgC), C.POINTER(ChoiceC), vp, vp, i64, i64, vp,
                                        C.c_size_t, vp]),
    "gtok_sent_log_bytes": (C.c_size_t, [C.POINTER(CsrBatchC), C.POINTER(SentCfgC), i64]),
    "gtok_sent_walk": (C.c_int, [C.POINTER(CsrBatchC), C.POINTER(SentCfgC), C.POINTER(ChoiceC), i64, vp, C.c_size_t, vp, vp]),
    "gtok_sent_emit_collate": (C.c_int, [C.POINTER(CsrBatchC), C.POINTER(SentCfgC), i64, vp, vp, vp, i64, i64, vp, vp, i64, vp, vp,
                                         vp]),
    "gtok_pad_tail_batched": (C.c_int, [vp, i64, i64, vp, vp, i64, vp, vp, vp, vp]),
    "gtok_zinc_serialize_rows": (C.c_int, [C.POINTER(GraphBatchC), C.POINTER(ZincVocabC), vp, vp, vp, vp]),
    "gtok_pad_mask": (C.c_int, [vp, vp, vp, vp, i64, i64, i64, vp, vp, vp]),
    "gtok_batch_extents": (C.c_int, [vp, vp, i64, i64, vp, vp, vp]),
    "gtok_pad_mask_batched": (C.c_int, [vp, vp, vp, vp, i64, i64, vp, vp, i64, vp, vp, vp]),
    "gtok_csr_adjacency": (C.c_int, [C.POINTER(GraphBatchC), i32, i64, vp, vp, vp]),
    "gtok_csr_fill": (C.c_int, [C.POINTER(GraphBatchC), i32, i64, vp, vp, i64, vp, vp, vp, vp, vp]),
    "gtok_compact_rows": (C.c_int, [vp, i64, vp, vp, i64, vp, vp]),
    "gtok_segment_counts": (C.c_int, [vp, i64, vp, i64, vp, vp, vp]),
    "gtok_gather_segments": (C.c_int, [vp, vp, i64, vp, vp, i64, vp, vp]),
    "gtok_ragged_append_lengths": (C.c_int, [vp, vp, i64, i32, vp, vp]),
    "gtok_ragged_append": (C.c_int, [vp, vp, vp, vp, i32, vp, i64, vp, vp]),
    "gtok_balance_order": (C.c_int, [vp, vp, i64, i32, vp, vp]),
    "gtok_csrl_count": (C.c_int, [C.POINTER(GraphBatchC), i64, vp, vp]),
    "gtok_csrl_table_bytes": (C.c_size_t, [i32]),
    "gtok_csrl_sort": (C.c_int, [C.POINTER(GraphBatchC), i64, vp, vp, vp, vp, vp, C.c_size_t, vp]),
    "gtok_csrl_emit": (C.c_int, [C.POINTER(GraphBatchC), i64, vp, vp, vp, i64, vp, vp, vp, vp]),
    "gtok_find_token": (C.c_int, [vp, vp, vp, i64, i64, vp, i64, vp, vp]),
    "gtok_graph_labels": (C.c_int, [C.POINTER(CsrBatchC), vp, vp, vp, vp]),
    "gtok_er_sizes": (C.c_int, [C.POINTER(ErSpecC), i64, vp, vp, vp, vp, vp]),
    "gtok_er_edges": (C.c_int, [C.POINTER(ErSpecC), i64, vp, vp, vp, vp]),
}


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(`python -c 'import __graft_entry__ as g; g.build()'` or `make -C glearning-benchmark_b200/csrc`). "
            "gtok has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != GTOK_OK:
        raise GtokError(rc, (lib.gtok_last_error() or b"").decode("utf-8", "replace"))


def version() -> int:
    return lib.gtok_version()
