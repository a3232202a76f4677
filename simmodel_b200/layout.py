"""Graph layout in HBM: deterministic CSR / CSC over a PyG-style batched edge_index.

Replaces the dense padded adjacency the reference ships per call
(comb_opt.py:342 ``torch.stack(Ws).to(device)``, :538 ``F.pad(W,...)``;
models_basic_lstm.py:515 ``to_dense_adj(ei)``).  All index arrays are int32 device
tensors; the CSR is built on the GPU by ``s2v_csr_build`` (csrc/s2v_csr.cu).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _lib.S2VError("simmodel_b200: %s must be a CUDA tensor (got %s); there is no CPU path" % (name, t.device))


class GraphLayout:
    """CSR (row i -> columns j it gathers), its transpose, in-degrees and the
    per-graph structure of one batch.

    period == 0: batched -- the CSR spans all ``num_nodes`` rows (Batch.from_data_list).
    period  > 0: replicated -- the CSR describes one graph of ``period`` nodes and the
                 batch is ``num_graphs`` copies with different node features (PPO path).
    """

    def __init__(self, num_nodes, num_graphs, period, rowptr, col, rowptr_t, col_t, indeg, ptr, gid, npad, row_to_col=None):
        self.num_nodes = int(num_nodes)
        self.num_graphs = int(num_graphs)
        self.period = int(period)
        self.rowptr, self.col, self.rowptr_t, self.col_t, self.indeg = rowptr, col, rowptr_t, col_t, indeg
        self.ptr, self.gid = ptr, gid
        self.npad = npad                      # int or int32 tensor [B]
        self.row_to_col = row_to_col          # int32 [E] or None: col position -> col_t position of the same edge (GATv2)
        self._struct = None

    @property
    def device(self):
        return self.rowptr.device

    @property
    def num_edges(self):
        return int(self.col.numel())

    def c_struct(self) -> _lib.Graph:
        if self._struct is None:
            g = _lib.Graph()
            g.num_nodes, g.num_graphs, g.period, g.num_edges = self.num_nodes, self.num_graphs, self.period, self.num_edges
            g.rowptr, g.col, g.rowptr_t, g.col_t, g.indeg = (_p(self.rowptr), _p(self.col), _p(self.rowptr_t),
                                                              _p(self.col_t), _p(self.indeg))
            g.ptr, g.gid = _p(self.ptr), _p(self.gid)
            if isinstance(self.npad, torch.Tensor):
                g.npad, g.npad_scalar = _p(self.npad), 0
            else:
                g.npad, g.npad_scalar = C.c_void_p(0), int(self.npad)
            self._struct = g
        return self._struct

    def with_npad(self, npad) -> "GraphLayout":
        """Same topology, different padded size(s): the only way Npad enters the result is
        the theta4 term."""
        if isinstance(npad, torch.Tensor):
            npad = npad.to(device=self.device, dtype=torch.int32).contiguous()
        return GraphLayout(self.num_nodes, self.num_graphs, self.period, self.rowptr, self.col, self.rowptr_t,
                           self.col_t, self.indeg, self.ptr, self.gid, npad, self.row_to_col)

    # ------------------------------------------------------------------
    @staticmethod
    def build_csr(edge_index: torch.Tensor, n: int, validate: bool = True, edge_map: bool = False,
                  eptr: torch.Tensor | None = None, nptr: torch.Tensor | None = None):
        """edge_index int64 or int32 [2,E] on the GPU -> (rowptr, col, rowptr_t, col_t, indeg) int32
        (+ row_to_col int32 [E] when edge_map).  Ids are global in [0,n), or -- with eptr / nptr (int32 [B+1]: first
        edge / first node of every graph) -- local to their graph; the offsets are then added on the device."""
        _require_cuda(edge_index, "edge_index")
        if edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise ValueError("edge_index must have shape [2,E]")
        if edge_index.dtype not in (torch.int32, torch.int64):
            edge_index = edge_index.to(torch.int64)
        ei = edge_index.contiguous()
        E = int(ei.shape[1])
        dev = ei.device
        L = _lib.lib()
        rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        rowptr_t = torch.empty(n + 1, dtype=torch.int32, device=dev)
        col = torch.empty(E, dtype=torch.int32, device=dev)
        col_t = torch.empty(E, dtype=torch.int32, device=dev)
        indeg = torch.empty(n, dtype=torch.int32, device=dev)
        ws = torch.empty(L.s2v_csr_workspace_bytes(n, E), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            if eptr is None and ei.dtype == torch.int64:
                _lib.check(L.s2v_csr_build(_p(ei), E, n, _p(rowptr), _p(col), _p(rowptr_t), _p(col_t), _p(indeg),
                                           _p(ws), ws.numel(), _stream()), "s2v_csr_build")
            else:
                nb = 0 if eptr is None else int(eptr.numel() - 1)
                _lib.check(L.s2v_csr_build_local(_p(ei), ei.element_size(), E, n, _p(eptr), _p(nptr), nb, _p(rowptr), _p(col),
                                                 _p(rowptr_t), _p(col_t), _p(indeg), _p(ws), ws.numel(), _stream()),
                           "s2v_csr_build_local")
        _lib.launch_count += 12 if E > 0 else 6
        if validate:
            flags = int(ws[:4].view(torch.int32).item())
            if flags & 1:
                raise ValueError("edge_index contains node ids outside [0, %d)" % n)
            if flags & 2:
                raise ValueError("edge_index contains duplicate edges: the reference sums them into a weighted adjacency "
                                 "(to_dense_adj, models_basic_lstm.py:515), which this path does not implement -- coalesce "
                                 "the edge list first")
        if edge_map:
            r2c = torch.empty(E, dtype=torch.int32, device=dev)
            scratch = torch.empty(E, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(L.s2v_csr_row_to_col(_p(ws), n, E, _p(scratch), _p(r2c), _stream()), "s2v_csr_row_to_col")
            _lib.launch_count += 2 if E > 0 else 0
            return rowptr, col, rowptr_t, col_t, indeg, r2c
        return rowptr, col, rowptr_t, col_t, indeg

    @classmethod
    def from_batch(cls, edge_index: torch.Tensor, ptr: torch.Tensor, n_pad, batch: torch.Tensor | None = None,
                   validate: bool = True, edge_map: bool = False, eptr: torch.Tensor | None = None) -> "GraphLayout":
        """Batched layout from a PyG-style (edge_index, ptr[, batch]) triple.
        n_pad: int (xv.shape[1] of the reference call) or per-graph tensor.
        edge_index may be int64 or int32.  With ``eptr`` ([B+1], first edge of every graph) the ids are LOCAL to
        their graph (what a replay buffer holds before Batch.from_data_list adds the offsets, comb_opt.py:299,344):
        the node offsets ptr[g] are added inside the CSR kernels -- half the host-to-device bytes, no host pass."""
        _require_cuda(edge_index, "edge_index")
        dev = edge_index.device
        ptr32 = ptr.to(device=dev, dtype=torch.int32).contiguous()
        B = int(ptr32.numel() - 1)
        n = int(ptr[-1]) if ptr.numel() else 0
        if batch is None:
            sizes = (ptr32[1:] - ptr32[:-1]).to(torch.int64)
            gid = torch.repeat_interleave(torch.arange(B, device=dev, dtype=torch.int32), sizes)
        else:
            gid = batch.to(device=dev, dtype=torch.int32).contiguous()
        if gid.numel() != n:
            raise ValueError("batch vector length %d != ptr[-1] %d" % (gid.numel(), n))
        eptr32 = None if eptr is None else eptr.to(device=dev, dtype=torch.int32).contiguous()
        if eptr32 is not None and eptr32.numel() != B + 1:
            raise ValueError("eptr must have %d entries" % (B + 1))
        rowptr, col, rowptr_t, col_t, indeg, *r2c = cls.build_csr(edge_index, n, validate, edge_map, eptr32,
                                                                  ptr32 if eptr32 is not None else None)
        if isinstance(n_pad, torch.Tensor):
            n_pad = n_pad.to(device=dev, dtype=torch.int32).contiguous()
        return cls(n, B, 0, rowptr, col, rowptr_t, col_t, indeg, ptr32, gid, n_pad, r2c[0] if r2c else None)

    @classmethod
    def from_topologies(cls, topologies, topo_of_graph: torch.Tensor, n_pad=None, validate: bool = True,
                        edge_map: bool = False) -> "GraphLayout":
        """Batch described the way a trainer holds it: a table of DISTINCT topologies -- ``topologies[t] =
        (edge_index_local [2,E_t] int32|int64 on the GPU, num_nodes_t)`` -- and ``topo_of_graph`` ([B] ints, host or
        device) naming the topology of every graph.  A replay batch drawn from one world (the reference's DQN / PPO
        training on a single map, comb_opt.py:342: Ws of identical graphs) ships its 2 850 edges once instead of
        B times: one topology -> replicated layout (period = num_nodes, no per-graph edge list at all); several ->
        the batched edge list is expanded on the device and goes through the same CSR builder as from_batch."""
        if len(topologies) == 0:
            raise ValueError("no topology given")
        B = int(topo_of_graph.numel())
        if len(topologies) == 1:
            ei, n = topologies[0]
            return cls.replicated(ei, int(n), B, n_pad, validate, edge_map)
        dev = topologies[0][0].device
        tog = topo_of_graph.to(device=dev, dtype=torch.int64)
        n_t = torch.tensor([int(n) for _, n in topologies], device=dev, dtype=torch.int64)
        e_t = torch.tensor([int(ei.shape[1]) for ei, _ in topologies], device=dev, dtype=torch.int64)
        eoff_t = torch.cumsum(e_t, 0) - e_t
        ei_cat = torch.cat([ei.to(torch.int32) for ei, _ in topologies], dim=1)
        ptr = torch.zeros(B + 1, dtype=torch.int64, device=dev)
        ptr[1:] = torch.cumsum(n_t[tog], 0)
        eptr = torch.zeros(B + 1, dtype=torch.int64, device=dev)
        eptr[1:] = torch.cumsum(e_t[tog], 0)
        goe = torch.repeat_interleave(torch.arange(B, device=dev), e_t[tog])
        src = eoff_t[tog][goe] + (torch.arange(int(eptr[-1]), device=dev) - eptr[:-1][goe])
        ei_local = ei_cat[:, src].contiguous()                       # int32, graph-local ids
        if n_pad is None:
            n_pad = int(n_t.max())
        return cls.from_batch(ei_local, ptr, n_pad, validate=validate, edge_map=edge_map, eptr=eptr)

    @classmethod
    def replicated(cls, edge_index: torch.Tensor, num_nodes: int, copies: int, n_pad=None,
                   validate: bool = True, edge_map: bool = False) -> "GraphLayout":
        """``copies`` graphs sharing one topology (S time steps of a PPO rollout)."""
        rowptr, col, rowptr_t, col_t, indeg, *r2c = cls.build_csr(edge_index, num_nodes, validate, edge_map)
        return cls(num_nodes * copies, copies, num_nodes, rowptr, col, rowptr_t, col_t, indeg, None, None,
                   num_nodes if n_pad is None else n_pad, r2c[0] if r2c else None)

    def replicate(self, copies: int) -> "GraphLayout":
        """Re-use the CSR of a replicated layout for a different number of copies."""
        if self.period <= 0:
            raise ValueError("replicate() needs a replicated layout")
        return GraphLayout(self.period * copies, copies, self.period, self.rowptr, self.col, self.rowptr_t,
                           self.col_t, self.indeg, None, None, self.npad, self.row_to_col)
