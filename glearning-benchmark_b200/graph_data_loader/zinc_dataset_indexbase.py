"""Drop-in for the reference's graph_data_loader/zinc_dataset_indexbase.py.

``ZINCTokenizationDataset`` keeps the reference constructor and item contract
(``ds[i] -> {'text','label','graph_id'}``, reference :63-70, :197-227).  The item's
text exists only because the unchanged trainer reads it to build its vocabulary
(trainer/train_ibtt.py:364-369): all texts are rendered at once on the GPU (K2 ids ->
gtok_detok_render) and an item is one slice of that buffer; ``dynamic_tokens()`` gives
the same token set without any text.  The token ids themselves come from the K2 CUDA
kernels when ``TokenDataset`` (data_loader.py) sees the hidden ``_zinc`` handle each
item carries - the text is never parsed back.
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

    # ---- all texts at once: tokenized and rendered on the GPU, items cut out of one byte buffer --------
    _RENDER_WORDS = ["<bos>", "<eos>", "<pad>", "<unk>", "<q>", "<p>", "<atom>", "<bond>"] + list(ATOM_TYPE_LIST) + \
        list(BOND_TYPE_LIST) + ["regression"]                        # the fixed vocabulary's ids 0..21 (zinc_vocab.py:35-78)

    def _bulk(self):
        """Every molecule's stripped token row (`<bos> ... <q> regression <p>`) from the K2 kernel under a private
        rendering vocabulary, turned into text bytes on the device; label tokens formatted in one vectorised pass.
        What `[ds[i] for i in range(len(ds))]` (trainer/train_ibtt.py:231-233) costs in the reference - one .item() per
        atom and bond - becomes three kernels and one slice per item."""
        st = getattr(self, "_bulk_state", None)
        if st is not None:
            return st
        import numpy as np
        from gtok import engine
        gb = self.graph_batch
        G = gb.num_graphs
        top = 0
        if gb.src.numel():
            top = max(int(gb.src.max()), int(gb.dst.max())) + 1
        words = self._RENDER_WORDS + [str(k) for k in range(max(top, 1))] + ["X", "unknown"]
        rv = {w_: i for i, w_ in enumerate(words)}
        big = 1 << 24
        rag = engine.tokenize_zinc(gb.to("cuda"), rv, big, item_max_len=big)
        buf, row_off = engine.render_texts(rag, words)
        y = gb.y.reshape(-1).to(torch.float32).to(torch.float64).numpy() if gb.y is not None else np.zeros(G)
        lab = np.char.mod("val_%.2f", y)
        lab = np.char.replace(np.char.replace(lab, ".", "_"), "-", "neg")              # reference :192
        st = {"buf": buf, "row_off": row_off, "ntok": rag.lengths.cpu().tolist(), "lab": lab.tolist(), "y": y.tolist(),
              "rag": rag, "words": words}
        self._bulk_state = st
        return st

    def dynamic_tokens(self, known=()):
        """The tokens of this dataset's texts that are not in the fixed vocabulary (node-index strings, 'X', 'unknown',
        the val_* labels), in order of first appearance - what the trainer's loop over every token of every text collects
        (trainer/train_ibtt.py:364-369; there into a set, whose order is arbitrary - SURVEY C5 - here deterministic).
        Computed from the device-resident ids (gtok_first_occurrence), labels from the formatted label column."""
        import numpy as np
        from gtok import engine
        st = self._bulk()
        rag, words = st["rag"], st["words"]
        nfix = len(self._RENDER_WORDS)
        first = engine.first_occurrence(rag.tokens, len(words)).cpu().numpy()
        off = rag.offsets.cpu().numpy()
        ev = []
        for i in range(nfix, len(words)):
            if first[i] >= 0:
                ev.append((int(np.searchsorted(off, first[i], side="right") - 1), 0, int(first[i]), words[i]))
        lab = np.asarray(st["lab"])
        if lab.size:
            _, idx = np.unique(lab, return_index=True)
            ev += [(int(i), 1, 0, str(lab[i])) for i in idx]
        seen = set(known)
        out = []
        for _, _, _, tok in sorted(ev):
            if tok not in seen:
                seen.add(tok)
                out.append(tok)
        return out

    def __getitem__(self, idx):
        if idx < 0:
            idx += len(self)
        if torch.cuda.is_available():
            st = self._bulk()
            a, b = st["row_off"][idx], st["row_off"][idx + 1]
            text = st["buf"][a:b - 1].decode("ascii") + " " + st["lab"][idx] + " <eos>"
            label = st["y"][idx]
            if st["ntok"][idx] + 2 > self.max_len:                        # reference :216-221 (rare: max_len below the row)
                toks = text.split()
                text = " ".join(toks[:self.max_len - 1] + ['<eos>'])
        else:                                                            # no device: the per-item host rendering (text only -
            data = self.zinc_dataset[idx]                                # TokenDataset still needs CUDA for the ids)
            label = data.y.item()
            toks = self.tokenize_molecule(data, label).split()
            if len(toks) > self.max_len:
                toks = toks[:self.max_len - 1] + ['<eos>']
            text = " ".join(toks)
        return {'text': text, 'label': label, 'graph_id': f"zinc_{self.split}_{idx}", '_zinc': (self, int(idx))}


def collate_zinc_batch(batch, pad_id):
    """Unused by the reference trainers ("not actually used", reference :235); same return."""
    return [b['text'] for b in batch], torch.tensor([b['label'] for b in batch], dtype=torch.float32)
