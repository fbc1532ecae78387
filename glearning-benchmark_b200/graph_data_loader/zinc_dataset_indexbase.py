"""Drop-in for the reference's graph_data_loader/zinc_dataset_indexbase.py.

``ZINCTokenizationDataset`` keeps the reference constructor and item contract
(``ds[i] -> {'text','label','graph_id'}``, reference :63-70, :197-227).  The item's
text is rendered on the host only because the unchanged trainer reads it to build
its vocabulary (trainer/train_ibtt.py:364-369); the token ids themselves come from
the K2 CUDA kernels when ``TokenDataset`` (data_loader.py) sees the hidden
``_zinc`` handle each item carries - the text is never parsed back.
"""
from __future__ import annotations

import torch
from torch.utils.data import Dataset

from .zinc_vocab import ZINC_ATOM_TYPES as ATOM_TYPE_LIST, ZINC_BOND_TYPES as BOND_TYPE_LIST

ZINC_ATOM_TYPES = {i: a for i, a in enumerate(ATOM_TYPE_LIST)}
ZINC_BOND_TYPES = {0: 'single', 1: 'single', 2: 'double', 3: 'triple', 4: 'aromatic'}   # reference :30-36


def _label_token(label: float) -> str:
    return f"val_{label:.2f}".replace('.', '_').replace('-', 'neg')            # reference :192


class ZINCTokenizationDataset(Dataset):
    """Args as the reference (:63-70); ``zinc_dataset`` may be passed to skip PyG (a sequence of
    objects with .x [N,1] or [N], .edge_index [2,E], .edge_attr [E], .y)."""

    def __init__(self, zinc_root: str = './data/ZINC', split: str = 'train', subset: bool = True,
                 max_vocab: int = 10000, max_len: int = 2048, zinc_dataset=None):
        super().__init__()
        self.zinc_root, self.split, self.subset, self.max_len = zinc_root, split, subset, max_len
        if zinc_dataset is None:
            try:
                from torch_geometric.datasets import ZINC
            except ImportError as e:                                   # reference __init__.py:46-52
                raise ImportError("ZINCTokenizationDataset requires torch_geometric to load ZINC. "
                                  "Install it with: pip install torch-geometric") from e
            zinc_dataset = ZINC(root=zinc_root, subset=subset, split=split)
        self.zinc_dataset = zinc_dataset
        self._gb = None
        print(f"Loaded ZINC {split} split: {len(self.zinc_dataset)} molecules")

    def __len__(self):
        return len(self.zinc_dataset)

    @property
    def graph_batch(self):
        """All molecules as one flat GraphBatch (host); built once, uploaded by TokenDataset."""
        if self._gb is None:
            from gtok.batch import graph_batch_from_dataset
            self._gb = graph_batch_from_dataset(self.zinc_dataset)      # InMemoryDataset storage converts without a Python loop
        return self._gb

    # ---- host rendering of the text (API compatibility only) ---------------------------------
    def decode_atom_features(self, node_features):
        return [ZINC_ATOM_TYPES.get(int(v), 'X') for v in torch.as_tensor(node_features).reshape(-1).tolist()]

    def decode_bond_features(self, edge_attr):
        names = {1: 'single', 2: 'double', 3: 'triple', 4: 'aromatic'}
        return [names.get(int(v), 'unknown') for v in torch.as_tensor(edge_attr).reshape(-1).tolist()]

    def tokenize_molecule(self, data, label):
        """`<bos> (<atom> SYM)* (<bond> TYPE u v)* <q> regression <p> val_.. <eos>` with first-seen
        undirected-pair dedupe in edge_index column order (reference :143-195)."""
        toks = ["<bos>"]
        for sym in self.decode_atom_features(data.x):
            toks += ["<atom>", sym]
        bonds = self.decode_bond_features(data.edge_attr)
        seen = set()
        for i, (u, v) in enumerate(data.edge_index.t().tolist()):
            key = (u, v) if u <= v else (v, u)
            if key in seen:
                continue
            seen.add(key)
            toks += ["<bond>", bonds[i] if i < len(bonds) else 'unknown', str(u), str(v)]
        toks += ["<q>", "regression", "<p>", _label_token(label), "<eos>"]
        return " ".join(toks)

    def __getitem__(self, idx):
        data = self.zinc_dataset[idx]
        label = data.y.item()
        toks = self.tokenize_molecule(data, label).split()
        if len(toks) > self.max_len:                                  # reference :216-221
            toks = toks[:self.max_len - 1] + ['<eos>']
        return {'text': " ".join(toks), 'label': label, 'graph_id': f"zinc_{self.split}_{idx}",
                '_zinc': (self, int(idx))}


def collate_zinc_batch(batch, pad_id):
    """Unused by the reference trainers ("not actually used", reference :235); same return."""
    return [b['text'] for b in batch], torch.tensor([b['label'] for b in batch], dtype=torch.float32)
