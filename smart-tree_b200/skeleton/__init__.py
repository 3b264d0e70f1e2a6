from .skeletonize import Skeletonizer

__all__ = ["Skeletonizer"]
