"""Drop-in for the reference's graph_data_loader/zinc_dataset_autograph.py (:13-100)."""
from __future__ import annotations

from torch.utils.data import Dataset


class ZINCDatasetForAutoGraph(Dataset):
    """Returns the PyG Data objects as they are (edge_attr flattened to 1-D, reference :70-71) and
    exposes the whole split as a flat GraphBatch for the epoch-prefetching SENT tokenizer."""

    def __init__(self, zinc_root: str = './data/ZINC', split: str = 'train', subset: bool = True, zinc_dataset=None):
        super().__init__()
        self.zinc_root, self.split, self.subset = zinc_root, split, subset
        if zinc_dataset is None:
            try:
                from torch_geometric.datasets import ZINC
            except ImportError as e:
                raise ImportError("ZINCDatasetForAutoGraph requires torch_geometric to load ZINC. "
                                  "Install it with: pip install torch-geometric") from e
            zinc_dataset = ZINC(root=zinc_root, subset=subset, split=split)
        self.zinc_dataset = zinc_dataset
        self._gb = None
        print(f"Loaded ZINC {split} split: {len(self.zinc_dataset)} molecules")

    def __len__(self):
        return len(self.zinc_dataset)

    @property
    def graph_batch(self):
        if self._gb is None:
            from gtok.batch import graph_batch_from_dataset
            self._gb = graph_batch_from_dataset(self.zinc_dataset)      # InMemoryDataset storage converts without a Python loop
        return self._gb

    def __getitem__(self, idx):
        data = self.zinc_dataset[idx]
        if data.edge_attr.dim() == 2 and data.edge_attr.size(1) == 1:
            data.edge_attr = data.edge_attr.flatten()
        if getattr(data, "num_nodes", None) is None:
            data.num_nodes = int(data.x.shape[0])
        data._gtok = (self, int(idx))          # handle for Graph2TrailTokenizer's epoch prefetch
        return data


def get_zinc_num_types():
    """(9 atom types, 4 bond types), reference :76-100."""
    return 9, 4
