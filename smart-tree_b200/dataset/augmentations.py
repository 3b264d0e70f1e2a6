"""Inference-time preprocessing (/root/reference/smart_tree/dataset/augmentations.py:11-14,38-41,108-116).
The training-only augmentations of the reference are out of scope."""
from typing import Sequence

import torch

from ..data_types.cloud import Cloud


class Augmentation:
    def __call__(self, cloud: Cloud) -> Cloud:
        raise NotImplementedError


class CentreCloud(Augmentation):
    """x,z centred on the bounding box, lowest point at y=0; keeps only xyz and rgb."""

    def __call__(self, cloud: Cloud) -> Cloud:
        centre, (x, y, z) = cloud.bbox
        return cloud.translate(-centre + torch.tensor([0, y, 0], device=centre.device))


class FixedTranslate(Augmentation):
    def __init__(self, xyz):
        self.xyz = torch.tensor(xyz)

    def __call__(self, cloud):
        return cloud.translate(self.xyz)


class AugmentationPipeline(Augmentation):
    def __init__(self, augmentations: Sequence[Augmentation]):
        if not all(isinstance(a, Augmentation) for a in augmentations):
            raise TypeError("augmentations must be a sequence of Augmentation")
        self.augmentations = augmentations

    def __call__(self, cloud):
        for aug in self.augmentations:
            cloud = aug(cloud)
        return cloud
