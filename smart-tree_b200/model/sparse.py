"""/root/reference/smart_tree/model/sparse.py:9-19,40-61."""
import torch

from ..spconv.pytorch import SparseConvTensor


def sparse_from_batch(features, coordinates, device):
    batch_size = features.shape[0]                                  # sic (quirk C-2)
    features = features.to(device)
    coordinates = coordinates.to(device)
    values, _ = torch.max(coordinates, 0)                           # max, not max+1 (quirk C-3)
    return SparseConvTensor(features, coordinates.int(), values[1:], batch_size=batch_size)


def batch_collate(batch):
    feats, coords, masks, fn = zip(*batch)
    for i, c in enumerate(coords):
        c[:, 0] = i
    return [torch.cat(feats), torch.cat(coords), torch.cat(masks), fn]
