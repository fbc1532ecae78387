"""graph_data_loader - B200-native drop-in for the reference package of the same name.

Put ``glearning-benchmark_b200/`` (this directory's parent) on ``sys.path`` ahead of the
reference checkout and ``trainer/train_ibtt.py`` / ``trainer/train_agtt.py`` import these
names unchanged (reference graph_data_loader/__init__.py:69-104).  The hot-path classes run
on the gtok CUDA kernels; ``GraphTokenDataset`` / ``AddQueryEncoding`` (MPNN / GraphGPS
inputs, no tokenization) are outside this engine's scope and raise ImportError.
"""
from .data_loader import (                                            # noqa: F401
    SPECIAL, load_examples, load_examples_multi_algorithm, determine_num_classes,
    determine_num_classes_pyg, build_vocab_from_texts, TokenDataset, collate,
    resolve_split_globs, resolve_multi_algorithm_globs, parse_distance_label_from_text,
    parse_query_nodes_from_text, balance_classes, get_balanced_indices,
)
from .zinc_vocab import (                                             # noqa: F401
    build_fixed_zinc_vocab, get_atom_type_id, get_bond_type_id, get_atom_type_from_id,
    get_bond_type_from_id, extend_vocab_with_dynamic_tokens, map_autograph_token_to_fixed_id,
    ZINC_ATOM_TYPES, ZINC_BOND_TYPES, NUM_ATOM_TYPES, NUM_BOND_TYPES,
)
from .graph_token_dataset_autograph import GraphTokenDatasetForAutoGraph    # noqa: F401
from .zinc_dataset_indexbase import ZINCTokenizationDataset, collate_zinc_batch   # noqa: F401
from .zinc_dataset_autograph import ZINCDatasetForAutoGraph, get_zinc_num_types   # noqa: F401

_OUT_OF_SCOPE = ("GraphTokenDataset", "add_query_encoding_to_features", "AddQueryEncoding")


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        raise ImportError(f"{name} feeds the MPNN / GraphGPS models (no tokenization) and is outside the gtok "
                          "engine; import it from the reference's graph_data_loader instead")
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


__all__ = [
    'SPECIAL', 'load_examples', 'load_examples_multi_algorithm', 'determine_num_classes',
    'determine_num_classes_pyg', 'build_vocab_from_texts', 'TokenDataset', 'collate', 'resolve_split_globs',
    'resolve_multi_algorithm_globs', 'parse_distance_label_from_text', 'parse_query_nodes_from_text',
    'balance_classes', 'get_balanced_indices', 'GraphTokenDatasetForAutoGraph', 'ZINCTokenizationDataset',
    'collate_zinc_batch', 'ZINCDatasetForAutoGraph', 'get_zinc_num_types', 'build_fixed_zinc_vocab',
    'get_atom_type_id', 'get_bond_type_id', 'get_atom_type_from_id', 'get_bond_type_from_id',
    'extend_vocab_with_dynamic_tokens', 'map_autograph_token_to_fixed_id', 'ZINC_ATOM_TYPES', 'ZINC_BOND_TYPES',
    'NUM_ATOM_TYPES', 'NUM_BOND_TYPES',
]
