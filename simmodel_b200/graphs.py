"""Graph topologies and synthetic node-feature matrices of the named configs.

Topologies are the real ones the reference trains on:
NWB_AMS (975 nodes), NWB_UTR (1182), NWB_ROT (2602) from
``data/nwb_topologies.npz`` (extracted by ``oracle/extract_topologies.py`` from
the reference's datasets/G_nwb/4.GEPHI_to_SIM/*.bin), Manhattan5 (5x5 grid with
self-loops, modules/rl/rl_custom_worlds.py + make_reflexive,
simdata_utils.py:508-510) and the MemTask graph (modules/sim/sim_graphs.py:168).
edge_index follows the reference convention EI = W.nonzero().t()
(simdata_utils.py:519): row-major, [0] = row (gathering node), [1] = column.

Node features follow NFM_ev_ec_t_dt_at_um_us (modules/gnn/nfm_gen.py:245-306),
F = 7: visited, current position (one-hot), target, node score, hops to nearest
target, units on the move, units settled.
"""
from __future__ import annotations

import os

import numpy as np
import torch

_DATA = os.path.join(os.path.dirname(__file__), "data", "nwb_topologies.npz")
NODE_DIM = 7


class Topology:
    def __init__(self, name, num_nodes, edge_index, target_nodes, n_units):
        self.name = name
        self.num_nodes = int(num_nodes)
        self.edge_index = edge_index          # int64 (2,E), row-major
        self.target_nodes = target_nodes      # int64 (K,)
        self.n_units = int(n_units)

    @property
    def num_edges(self):
        return int(self.edge_index.shape[1])

    def neighbors(self, node: int):
        """Ascending neighbour list of a node (simdata_utils.py:632-642)."""
        ei = self.edge_index
        return sorted(ei[1][ei[0] == node].tolist())


def _row_major(ei: np.ndarray) -> torch.Tensor:
    order = np.lexsort((ei[1], ei[0]))
    return torch.from_numpy(ei[:, order].astype(np.int64))


def load_topology(name: str) -> Topology:
    if name in ("NWB_AMS", "NWB_UTR", "NWB_ROT"):
        with np.load(_DATA) as z:
            ei = torch.from_numpy(z[name + "/edge_index"].astype(np.int64))
            n = int(z[name + "/num_nodes"])
            tg = torch.from_numpy(z[name + "/target_nodes"].astype(np.int64))
        units = {"NWB_AMS": 10, "NWB_UTR": 10, "NWB_ROT": 20}[name]
        return Topology(name, n, ei, tg, units)
    if name == "Manhattan5":
        return manhattan(5, n_units=3)
    if name == "Manhattan3":
        return manhattan(3, n_units=2)
    if name == "MemTask":
        return memtask()
    raise KeyError(name)


def manhattan(side: int = 5, n_units: int = 3, reflexive: bool = True) -> Topology:
    """side x side grid, node id = row * side + col, optional self-loops; targets = top row."""
    e = []
    for r in range(side):
        for c in range(side):
            i = r * side + c
            if reflexive:
                e.append((i, i))
            if c + 1 < side:
                e += [(i, i + 1), (i + 1, i)]
            if r + 1 < side:
                e += [(i, i + side), (i + side, i)]
    ei = _row_major(np.array(e, dtype=np.int64).T)
    return Topology("Manhattan%d" % side, side * side, ei, torch.arange(side * (side - 1), side * side), n_units)


def memtask(reflexive: bool = True) -> Topology:
    """Directed 8-node MemGraph (sim_graphs.py:168-224): 0-1, 1-2, 3-4, 4-2, 4-5, 6-7, 7-5 plus self-loops."""
    e = [(0, 1), (1, 2), (3, 4), (4, 2), (4, 5), (6, 7), (7, 5)]
    if reflexive:
        e += [(i, i) for i in range(8)]
    return Topology("MemTask", 8, _row_major(np.array(e, dtype=np.int64).T), torch.tensor([2, 5]), 1)


def synth_nfm(topo: Topology, num_graphs: int, generator: torch.Generator, device="cpu") -> torch.Tensor:
    """(num_graphs, N, 7) float32 node features with the value ranges of the
    reference's NFM: integer-valued flags/counts, a score column in
    [0,1] with targets at 2.0, hop counts."""
    n, g = topo.num_nodes, generator
    dev = g.device
    x = torch.zeros(num_graphs, n, NODE_DIM, dtype=torch.float32, device=dev)
    ar = torch.arange(num_graphs, device=dev)
    x[:, :, 0] = (torch.rand(num_graphs, n, generator=g, device=dev) < 0.05).float()
    cur = torch.randint(0, n, (num_graphs,), generator=g, device=dev)
    x[ar, cur, 0] = 1.0
    x[ar, cur, 1] = 1.0
    tg = topo.target_nodes.to(dev)
    x[:, tg, 2] = 1.0
    x[:, :, 3] = torch.rand(num_graphs, n, generator=g, device=dev)
    x[:, tg, 3] = 2.0
    x[:, :, 4] = torch.randint(0, 9, (num_graphs, n), generator=g, device=dev).float()
    x[:, tg, 4] = 0.0
    upos = torch.randint(0, n, (num_graphs, topo.n_units), generator=g, device=dev)
    settled = torch.rand(num_graphs, topo.n_units, generator=g, device=dev) < 0.5
    ones = torch.ones(num_graphs, topo.n_units, device=dev)
    x[:, :, 5].scatter_add_(1, upos, ones * (~settled))
    x[:, :, 6].scatter_add_(1, upos, ones * settled)
    return x.to(device)


def batch_edge_index(topo: Topology, num_graphs: int, device="cpu") -> torch.Tensor:
    """PyG Batch.from_data_list edge_index of num_graphs copies: offsets by cumulative node counts."""
    ei = topo.edge_index.to(device)
    off = (torch.arange(num_graphs, device=device, dtype=torch.int64) * topo.num_nodes).view(-1, 1, 1)
    return (ei.unsqueeze(0) + off).permute(1, 0, 2).reshape(2, -1).contiguous()
