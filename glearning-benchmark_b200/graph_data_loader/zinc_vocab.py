"""Fixed ZINC vocabulary - drop-in for the reference's graph_data_loader/zinc_vocab.py.

Same names, ids and error behaviour (reference zinc_vocab.py:20-28, :35-78, :81-147,
:154-179, :186-238); the kernels consume these ids through GtokZincVocab, they never
hard-code anything but what this table says.
"""
from typing import Dict, Tuple

ZINC_ATOM_TYPES = ['C', 'N', 'O', 'F', 'P', 'S', 'Cl', 'Br', 'I']
NUM_ATOM_TYPES = len(ZINC_ATOM_TYPES)
ZINC_BOND_TYPES = ['single', 'double', 'triple', 'aromatic']
NUM_BOND_TYPES = len(ZINC_BOND_TYPES)
SPECIAL_TOKENS = ['<bos>', '<eos>', '<pad>', '<unk>', '<q>', '<p>', '<atom>', '<bond>']
_ATOM_BASE = len(SPECIAL_TOKENS)                  # 8
_BOND_BASE = _ATOM_BASE + NUM_ATOM_TYPES          # 17


def build_fixed_zinc_vocab() -> Tuple[Dict[str, int], Dict[int, str]]:
    """ids 0-7 specials, 8-16 atoms, 17-20 bonds, 21 'regression'."""
    vocab = {t: i for i, t in enumerate(SPECIAL_TOKENS + ZINC_ATOM_TYPES + ZINC_BOND_TYPES + ['regression'])}
    return vocab, {i: t for t, i in vocab.items()}


def get_atom_type_id(atom_type_idx: int) -> int:
    if not (0 <= atom_type_idx < NUM_ATOM_TYPES):
        raise ValueError(f"Invalid atom type index: {atom_type_idx} (expected 0-{NUM_ATOM_TYPES-1})")
    return _ATOM_BASE + atom_type_idx


def get_bond_type_id(bond_type_idx: int) -> int:
    if not (1 <= bond_type_idx <= NUM_BOND_TYPES):
        raise ValueError(f"Invalid bond type index: {bond_type_idx} (expected 1-{NUM_BOND_TYPES})")
    return _BOND_BASE + (bond_type_idx - 1)


def get_atom_type_from_id(token_id: int) -> str:
    if not (_ATOM_BASE <= token_id < _BOND_BASE):
        raise ValueError(f"Invalid atom type token ID: {token_id} (expected 8-16)")
    return ZINC_ATOM_TYPES[token_id - _ATOM_BASE]


def get_bond_type_from_id(token_id: int) -> str:
    if not (_BOND_BASE <= token_id < _BOND_BASE + NUM_BOND_TYPES):
        raise ValueError(f"Invalid bond type token ID: {token_id} (expected 17-20)")
    return ZINC_BOND_TYPES[token_id - _BOND_BASE]


def extend_vocab_with_dynamic_tokens(base_vocab: Dict[str, int], dynamic_tokens: list) -> Dict[str, int]:
    """Append unseen tokens from max(id)+1 in the GIVEN list order (zinc_vocab.py:154-179)."""
    vocab = dict(base_vocab)
    nxt = max(vocab.values()) + 1
    for tok in dynamic_tokens:
        if tok not in vocab:
            vocab[tok] = nxt
            nxt += 1
    return vocab


def map_autograph_token_to_fixed_id(autograph_token_id: int, tokenizer_node_idx_offset: int,
                                    tokenizer_edge_idx_offset: int, is_node_type: bool = False,
                                    is_edge_type: bool = False) -> int:
    """SOS/EOS/PAD -> 0/1/2, typed tokens -> fixed ids, everything else unchanged (zinc_vocab.py:186-238)."""
    special = {0: 0, 4: 1, 5: 2}
    if autograph_token_id in special:
        return special[autograph_token_id]
    if is_node_type:
        return get_atom_type_id(autograph_token_id - tokenizer_node_idx_offset)
    if is_edge_type:
        return get_bond_type_id(autograph_token_id - tokenizer_edge_idx_offset + 1)
    return autograph_token_id
