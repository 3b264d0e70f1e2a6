"""Graph container (/root/reference/smart_tree/data_types/graph.py:14-66).  The cugraph
objects of the reference are replaced by plain tensors: a component is a `Component` record."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import torch

from .. import ops


@dataclass
class Component:
    vertex_ids: torch.Tensor      # ascending ids into the parent graph's vertices
    edges: torch.Tensor           # [E,2] int32, renumbered 0..n-1
    edge_weights: torch.Tensor    # [E]

    def __len__(self):
        return int(self.vertex_ids.shape[0])


@dataclass
class Graph:
    vertices: torch.Tensor        # [N,3]
    edges: torch.Tensor           # [E,2] (directed candidates; treated as undirected)
    edge_weights: torch.Tensor    # [E]

    def to_device(self, device):
        return Graph(self.vertices.to(device), self.edges.to(device), self.edge_weights.to(device))

    def component_labels(self):
        """label[v] = smallest vertex id of v's weakly connected component, size[v] = its size."""
        return ops.connected_components(self.edges.int().contiguous(), self.vertices.shape[0])

    def ranked_components(self, minimum_vertices=10):
        """Roots (= smallest vertex id) and sizes of the components with >= minimum_vertices
        vertices, ordered by size descending then root ascending (graph.py:32-51; the tie order
        of the reference is implementation-defined)."""
        label, size = self.component_labels()
        n = label.shape[0]
        is_root = (label == torch.arange(n, device=label.device, dtype=label.dtype)) & (size >= minimum_vertices)
        roots = torch.nonzero(is_root).flatten()
        sizes = size[roots].long()
        order = torch.argsort(sizes * n + (n - 1 - roots), descending=True, stable=True)   # (-size, root)
        return label, roots[order], sizes[order]

    def connected_cugraph_components(self, minimum_vertices=10) -> List[Component]:
        label, roots, sizes = self.ranked_components(minimum_vertices)
        comps = []
        e0 = self.edges[:, 0].long()
        e_label = label[e0] if len(e0) else label[:0]
        for root in roots.tolist():
            vids = torch.nonzero(label == root).flatten()
            local = torch.full((label.shape[0],), -1, dtype=torch.int32, device=label.device)
            local[vids] = torch.arange(len(vids), dtype=torch.int32, device=label.device)
            sel = e_label == root
            comps.append(Component(vids, local[self.edges[sel].long()].contiguous(), self.edge_weights[sel].contiguous()))
        return comps
