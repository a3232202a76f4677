"""GATv2 embedding branch (``--qnet gat2``).

Mirrors what the reference builds from torch_geometric (third party, ~2.0.4):
``GATv2(BasicGNN)`` with ``GATv2Conv`` layers (modules/gnn/comb_opt.py:48-62, 156-182;
modules/ppo/models_basic_lstm.py:326-338, 414-426) -- same constructor arguments and the
same state-dict keys (``gat.convs.{k}.{att,bias,lin_l.weight,lin_l.bias,lin_r.weight,lin_r.bias}``,
checked against the reference's shipped checkpoint), so reference checkpoints load strict.

Per layer: the two node-wise linear layers are ONE plain GEMM over the concatenated
weights (cuBLAS through torch -- a library GEMM); everything edge-wise -- attention scores,
per-target segment softmax, weighted neighbour sum, bias, inter-layer ReLU, and the whole
backward of those -- runs in the sm_100a kernels of csrc/s2v_gat.cu behind the C ABI
(s2v_gat_attn_forward / _backward).  CUDA only; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import S2VError
from .layout import GraphLayout, _p, _require_cuda, _stream


def strip_self_loops(edge_index: torch.Tensor) -> torch.Tensor:
    """remove_self_loops of GATv2Conv.forward; the kernels add the self edge of every node themselves."""
    return edge_index[:, edge_index[0] != edge_index[1]].contiguous()


def gat_layout_from_batch(edge_index, ptr, batch=None) -> GraphLayout:
    """Batched layout for the GAT kernels (PyG Batch.from_data_list order)."""
    _require_cuda(edge_index, "edge_index")
    return GraphLayout.from_batch(strip_self_loops(edge_index), ptr, 0, batch, edge_map=True)


def gat_layout_replicated(edge_index, num_nodes: int, copies: int) -> GraphLayout:
    """S copies of one topology (models_basic_lstm.py:509-513 builds a Batch of S Data(e, ei))."""
    _require_cuda(edge_index, "edge_index")
    return GraphLayout.replicated(strip_self_loops(edge_index), num_nodes, copies, edge_map=True)


_CHUNK = 4096


def _rows_contract(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a^T b over the (long) row dimension: a [n,M], b [n,K] -> [M,K].  A library GEMM either way; for n >> M,K it
    is issued as a batched GEMM over row chunks plus a fixed-order sum, which cuBLAS runs several times faster than
    one GEMM with a 10^6-long contraction."""
    n = a.shape[0]
    m = (n // _CHUNK) * _CHUNK
    if m < 8 * _CHUNK:
        return a.t().mm(b)
    out = torch.bmm(a[:m].view(m // _CHUNK, _CHUNK, -1).transpose(1, 2), b[:m].view(m // _CHUNK, _CHUNK, -1)).sum(0)
    if m < n:
        out = out + a[m:].t().mm(b[m:])
    return out


TENSOR_MIN_NODES = 4096


def _tensor_lin(x_dim: int, HC: int, n: int, shared: bool) -> bool:
    """lin_l / lin_r (and their backward) on the tcgen05 kernels (s2v_rows_linear / s2v_rows_outer, bf16x3): every layer
    shape of the configurations the reference ships -- in = F (7, zero padded), 64 or 128; heads*channels 64 or 128;
    separate or shared weights -- for large batches.  S2V_B200_PATH=ffma keeps every GEMM in fp32 round-to-nearest
    (cuBLAS), small batches stay on the library GEMM (launch-bound either way)."""
    return (1 <= x_dim <= 128 and HC in (64, 128) and n >= TENSOR_MIN_NODES and _lib.path_flag() != _lib.PATHS["ffma"])


class _GATConvFn(torch.autograd.Function):
    """wr / br None => share_weights=True (x_r = x_l = lin_l(x); models_ngo.py:241-249).
    relu: 0 none, 1 ReLU on the output with the mask applied in this layer's backward, 2 ReLU on the output whose
    backward mask the CONSUMER applies (the next layer's dx GEMM epilogue: its input is this output).
    mask_dx: this layer's input is a ReLU output whose producer left the mask to us."""

    @staticmethod
    def forward(ctx, layout, heads, relu, mask_dx, x, wl, bl, wr, br, att, bias):
        L = _lib.lib()
        x = ops._f32c(x, "x")
        n, HC = layout.num_nodes, int(wl.shape[0])
        Cc = HC // heads
        shared = wr is None
        if x.shape[0] != n or x.shape[1] != wl.shape[1]:
            raise ValueError("x has shape %s, expected (%d, %d)" % (tuple(x.shape), n, wl.shape[1]))
        wcat = wl.detach() if shared else torch.cat((wl.detach(), wr.detach()), dim=0)      # [2HC, in]
        bcat = torch.cat((bl.detach(), bl.detach() if shared else br.detach()), dim=0)     # lin_l.bias | lin_r.bias
        tensor = _tensor_lin(int(x.shape[1]), HC, n, shared)
        if tensor:
            wcat = wcat.contiguous()
            m_out = int(wcat.shape[0])                   # 2 HC, or HC with shared weights
            xlr = torch.empty(n, m_out, dtype=torch.float32, device=x.device)
            with torch.cuda.device(x.device):
                _lib.check(L.s2v_rows_linear(_p(x), int(x.shape[1]), int(x.shape[1]), _p(wcat), int(wcat.shape[1]), m_out, 0, None,
                                             0, _p(xlr), m_out, n, _stream()), "s2v_rows_linear")
            _lib.launch_count += 1
        else:
            xlr = x.mm(wcat.t())                         # [n, 2HC]: x_l | x_r without biases (plain GEMM, no epilogue pass)
        ld = int(xlr.shape[1])
        xr_ptr = _p(xlr) if shared else C.c_void_p(xlr.data_ptr() + 4 * HC)
        att_f = ops._f32c(att.detach().reshape(-1), "att")
        bias_f = ops._f32c(bias.detach(), "bias")
        out = torch.empty(n, HC, dtype=torch.float32, device=x.device)
        # what the backward needs from the softmax: per-edge weights (specialised shapes, with the layout's edge map)
        # or 16 bytes of statistics per node and head (generic kernels)
        edge_alpha = bool(L.s2v_gat_uses_edge_alpha(heads, Cc, ld)) and (layout.row_to_col is not None or layout.num_edges == 0)
        if edge_alpha:
            edge_floats = int(L.s2v_gat_edge_floats(C.byref(layout.c_struct()), heads))
            stats = torch.empty(edge_floats, dtype=torch.float32, device=x.device)
        else:
            edge_floats = 0
            stats = torch.empty(n, heads, 4, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.s2v_gat_attn_forward(C.byref(layout.c_struct()), heads, Cc, ld, _p(xlr), xr_ptr, _p(att_f),
                                              _p(bias_f), _p(bcat), int(relu != 0), _p(out),
                                              None if edge_alpha else _p(stats), _p(stats) if edge_alpha else None, _stream()),
                       "s2v_gat_attn_forward")
        ctx.edge_floats = edge_floats
        _lib.launch_count += 1
        ctx.layout, ctx.heads, ctx.relu, ctx.shared, ctx.tensor, ctx.mask_dx = layout, heads, relu, shared, tensor, mask_dx
        ctx.save_for_backward(x, xlr, stats, out, att_f, wcat, bcat)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = _lib.lib()
        x, xlr, stats, out, att_f, wcat, bcat = ctx.saved_tensors
        layout, heads, shared = ctx.layout, ctx.heads, ctx.shared
        n, HC = layout.num_nodes, int(out.shape[1])
        Cc = HC // heads
        dout = ops._f32c(dout, "dout")
        with torch.cuda.device(x.device):
            if ctx.relu == 1:
                masked = torch.empty_like(dout)
                _lib.check(L.s2v_relu_mask(_p(dout), _p(out), _p(masked), dout.numel(), _stream()), "s2v_relu_mask")
                _lib.launch_count += 1
                dout = masked
            # shared weights: dx_l and dx_r land in two planes that are summed afterwards (both feed lin_l)
            dxlr = torch.empty((2, n, HC) if shared else (n, 2 * HC), dtype=torch.float32, device=x.device)
            ld = HC if shared else 2 * HC
            dxr_ptr = C.c_void_p(dxlr.data_ptr() + 4 * (n * HC if shared else HC))
            xr_ptr = _p(xlr) if shared else C.c_void_p(xlr.data_ptr() + 4 * HC)
            gab = torch.empty(4, HC, dtype=torch.float32, device=x.device)       # datt | dbias | dlin_l.bias | dlin_r.bias
            ef = ctx.edge_floats
            ws = torch.empty(L.s2v_gat_attn_backward_workspace_bytes(heads, Cc, ef), dtype=torch.uint8, device=x.device)
            _lib.check(L.s2v_gat_attn_backward(C.byref(layout.c_struct()), heads, Cc, ld, _p(xlr), xr_ptr, _p(att_f),
                                               _p(bcat), None if ef else _p(stats), _p(stats) if ef else None,
                                               _p(layout.row_to_col) if ef else None, _p(dout), _p(dxlr), dxr_ptr, _p(gab[0]), _p(gab[1]),
                                               _p(gab[2]), _p(ws), ws.numel(), _stream()), "s2v_gat_attn_backward")
        _lib.launch_count += 3
        def tensor_grads(dy):
            """dW = dy^T x and dx = dy W (* [x > 0] when the producer left its ReLU mask to us) on the tcgen05 kernels."""
            K, M = int(x.shape[1]), int(dy.shape[1])
            dw = torch.empty(M, K, dtype=torch.float32, device=x.device)
            dx = None
            with torch.cuda.device(x.device):
                ws2 = torch.empty(L.s2v_rows_reduce64_workspace_bytes(2), dtype=torch.uint8, device=x.device)
                _lib.check(L.s2v_rows_outer(_p(dy), M, M, _p(x), K, K, _p(dw), K, _p(ws2), ws2.numel(), n, _stream()),
                           "s2v_rows_outer")
                if ctx.needs_input_grad[4]:
                    if K % 64 != 0:
                        raise NotImplementedError("gradient w.r.t. a %d-wide layer input on the tensor path" % K)
                    dx = torch.empty_like(x)
                    _lib.check(L.s2v_rows_linear(_p(dy), M, M, _p(wcat), K, K, 1, _p(x) if ctx.mask_dx else None, K, _p(dx), K, n,
                                                 _stream()), "s2v_rows_linear")
            _lib.launch_count += 3
            return dw, dx
        if shared:
            dxs = dxlr[0] + dxlr[1]
            if ctx.tensor:
                dw, dx = tensor_grads(dxs)
            else:
                dw = _rows_contract(dxs, x)
                dx = dxs.mm(wcat) if ctx.needs_input_grad[4] else None
                if dx is not None and ctx.mask_dx:
                    with torch.cuda.device(x.device):
                        _lib.check(L.s2v_relu_mask(_p(dx), _p(x), _p(dx), dx.numel(), _stream()), "s2v_relu_mask")
            return (None, None, None, None, dx, dw, gab[2] + gab[3], None, None, gab[0].view(1, heads, Cc), gab[1])
        if ctx.tensor:
            dw, dx = tensor_grads(dxlr)
        else:
            dw = _rows_contract(dxlr, x)                                      # [2HC, in]   (plain GEMMs)
            dx = dxlr.mm(wcat) if ctx.needs_input_grad[4] else None
            if dx is not None and ctx.mask_dx:                                # the producer left its ReLU mask to us
                with torch.cuda.device(x.device):
                    _lib.check(L.s2v_relu_mask(_p(dx), _p(x), _p(dx), dx.numel(), _stream()), "s2v_relu_mask")
        return (None, None, None, None, dx, dw[:HC], gab[2], dw[HC:], gab[3], gab[0].view(1, heads, Cc), gab[1])


class GATv2Conv(nn.Module):
    """torch_geometric.nn.conv.GATv2Conv in the reference's configuration: concat=True, add_self_loops=True,
    share_weights=False, negative_slope 0.2, dropout 0, bias=True.  Init follows PyG (glorot weights and att,
    Linear bias U(-1/sqrt(in), 1/sqrt(in)), zero output bias)."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout=0.0,
                 add_self_loops=True, bias=True, share_weights=False, **kwargs):
        super().__init__()
        if not (concat and add_self_loops and bias) or dropout != 0.0 or negative_slope != 0.2:
            raise NotImplementedError("simmodel_b200.GATv2Conv covers the reference's configurations only "
                                      "(concat, self loops, slope 0.2, no dropout; share_weights either way)")
        if heads * out_channels > 128:
            raise NotImplementedError("GATv2Conv kernels cover heads*out_channels <= 128 (got %d)" % (heads * out_channels))
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.share_weights = bool(share_weights)
        self.lin_l = nn.Linear(in_channels, heads * out_channels, bias=True)
        # PyG: `self.lin_r = self.lin_l` when sharing -- one module under two names (both appear in the state dict)
        self.lin_r = self.lin_l if share_weights else nn.Linear(in_channels, heads * out_channels, bias=True)
        self.att = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(heads * out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        for lin in ((self.lin_l,) if self.share_weights else (self.lin_l, self.lin_r)):
            a = math.sqrt(6.0 / (lin.in_features + lin.out_features))
            nn.init.uniform_(lin.weight, -a, a)
            nn.init.uniform_(lin.bias, -1.0 / math.sqrt(lin.in_features), 1.0 / math.sqrt(lin.in_features))
        a = math.sqrt(6.0 / (self.heads + self.out_channels))
        nn.init.uniform_(self.att, -a, a)
        nn.init.zeros_(self.bias)

    def forward_layout(self, x, layout: GraphLayout, relu: bool = False, relu_mask_by_consumer: bool = False,
                       mask_dx: bool = False):
        """relu_mask_by_consumer / mask_dx: see _GATConvFn (GATv2.forward_layout pairs them between adjacent layers)."""
        if not self.att.is_cuda:
            raise S2VError("simmodel_b200.GATv2Conv runs on CUDA only: call .to('cuda') first (no CPU fallback)")
        mode = 0 if not relu else (2 if relu_mask_by_consumer else 1)
        if self.share_weights:
            return _GATConvFn.apply(layout, self.heads, mode, bool(mask_dx), x, self.lin_l.weight, self.lin_l.bias, None, None,
                                    self.att, self.bias)
        return _GATConvFn.apply(layout, self.heads, mode, bool(mask_dx), x, self.lin_l.weight, self.lin_l.bias,
                                self.lin_r.weight, self.lin_r.bias, self.att, self.bias)

    def uses_tensor_lin(self, n: int) -> bool:
        return _tensor_lin(self.in_channels, self.heads * self.out_channels, n, self.share_weights)

    def forward(self, x, edge_index):
        n = int(x.shape[0])
        ptr = torch.tensor([0, n], dtype=torch.int64, device=x.device)
        return self.forward_layout(x, gat_layout_from_batch(edge_index.to(x.device), ptr))


class GATv2(nn.Module):
    """The reference's ``GATv2(BasicGNN)`` (comb_opt.py:50-62 / models_basic_lstm.py:326-338): ``num_layers``
    GATv2Conv layers, widths hidden .. hidden, out; heads falls back to 1 for a layer whose width it does not
    divide; ReLU between layers, none after the last (BasicGNN, jk=None, norm=None, dropout 0)."""

    def __init__(self, in_channels, hidden_channels, num_layers, out_channels=None, heads=1, dropout=0.0,
                 share_weights=False, concat=True, **kwargs):
        super().__init__()
        if dropout != 0.0:
            raise NotImplementedError("dropout is 0 in every reference configuration")
        if num_layers < 2:
            raise NotImplementedError("BasicGNN with fewer than 2 layers is not used by the reference")
        self.in_channels, self.hidden_channels, self.num_layers = in_channels, hidden_channels, num_layers
        self.out_channels = hidden_channels if out_channels is None else out_channels
        self.supports_edge_weight = False
        self.supports_edge_attr = False
        widths = [hidden_channels] * (num_layers - 1) + [self.out_channels]
        self.convs = nn.ModuleList()
        cin = in_channels
        for w in widths:
            h = heads if w % heads == 0 else 1
            self.convs.append(GATv2Conv(cin, w // h, heads=h, concat=concat, share_weights=share_weights))
            cin = w

    def forward_layout(self, x, layout: GraphLayout):
        n = layout.num_nodes
        for k, conv in enumerate(self.convs):
            # a layer whose lin GEMMs run on the tensor-core kernels applies the ReLU mask of its INPUT in the epilogue
            # of its dx GEMM; the producing layer then skips its own masking pass
            fused_here = k > 0 and conv.uses_tensor_lin(n)
            fused_next = k + 1 < self.num_layers and self.convs[k + 1].uses_tensor_lin(n)
            x = conv.forward_layout(x, layout, relu=k < self.num_layers - 1, relu_mask_by_consumer=fused_next,
                                    mask_dx=fused_here)
        return x

    def forward(self, x, edge_index):
        n = int(x.shape[0])
        ptr = torch.tensor([0, n], dtype=torch.int64, device=x.device)
        return self.forward_layout(x, gat_layout_from_batch(edge_index.to(x.device), ptr))


class QNet_GATv2(nn.Module):
    """Drop-in for comb_opt.py:156-182 QNet_GATv2 (forward: QNet_GAT.forward, 84-135): GATv2 embedding
    (hidden = emb_dim, heads 2, 5 layers) + the s2v Q head (theta5/6/7 with mean or sum pooling)."""

    def __init__(self, config, verbose=True):
        super().__init__()
        self.emb_dim = config["emb_dim"]
        self.num_layers = config["emb_iter_T"]
        self.node_dim = config["node_dim"]
        self.norm_agg = config["norm_agg"]
        self.gat = GATv2(in_channels=self.node_dim, hidden_channels=self.emb_dim, heads=2, num_layers=5,
                         out_channels=self.emb_dim, share_weights=False, concat=True)
        self.theta5 = nn.Linear(2 * self.emb_dim, 1, True, dtype=torch.float32)
        self.theta6 = nn.Linear(self.emb_dim, self.emb_dim, True, dtype=torch.float32)
        self.theta7 = nn.Linear(self.emb_dim, self.emb_dim, True, dtype=torch.float32)
        self.numTrainableParameters(verbose)

    def head_weights(self):
        return [t for n in ops.HEAD_NAMES for t in (getattr(self, n).weight, getattr(self, n).bias)]

    @property
    def device(self):
        return self.theta6.weight.device

    def forward_graph(self, x, layout: GraphLayout):
        """x f32[sum N, F], GAT layout (self loops stripped) -> q f32[sum N]."""
        if self.device.type != "cuda":
            raise S2VError("simmodel_b200.QNet_GATv2 runs on CUDA only: call .to('cuda') first (no CPU fallback)")
        return ops.s2v_head(layout, self.embed_graph(x, layout), self.head_weights(), self.norm_agg)

    def embed_graph(self, x, layout: GraphLayout):
        return self.gat.forward_layout(x, layout)

    @torch.no_grad()
    def greedy_actions_graph(self, x, layout, reach_ptr, reach_idx):
        """Greedy action per graph over its neighbour list: GATv2 embedding, then Q head + list arg-max in one launch."""
        return ops.s2v_head_list_argmax(layout, self.embed_graph(x, layout), self.head_weights(), reach_ptr, reach_idx,
                                        self.norm_agg)[:3]

    def forward(self, xv, Ws, pyg_data):
        """Reference contract: xv and Ws are ignored (comb_opt.py:84-94 reads pyg_data.x / .edge_index / .ptr /
        .batch only); returns q over all nodes of the batch, squeezed."""
        x, layout = self.prepare(xv, Ws, pyg_data)
        return self.forward_graph(x, layout).squeeze()

    def prepare(self, xv, Ws, pyg_data):
        """The reference call's arguments -> (x f32[sum N, F], GAT layout)."""
        dev = self.device
        if dev.type != "cuda":
            raise S2VError("simmodel_b200.QNet_GATv2 runs on CUDA only: call .to('cuda') first (no CPU fallback)")
        x = pyg_data.x.to(dev, torch.float32).contiguous()
        layout = gat_layout_from_batch(pyg_data.edge_index.to(dev), pyg_data.ptr.to(dev), pyg_data.batch.to(dev))
        return x, layout

    def numTrainableParameters(self, verbose=True):
        total = 0
        lines = ["Qnet size:", "------------------------------------------"]
        for name, p in self.named_parameters():
            total += int(p.numel())
            lines.append("{:24s} {:12s} requires_grad={}".format(name, str(list(p.shape)), p.requires_grad))
        lines += ["Total number of parameters: {}".format(total), "------------------------------------------"]
        if verbose:
            print("\n".join(lines))
        assert total == sum(p.numel() for p in self.parameters() if p.requires_grad)
        return total
