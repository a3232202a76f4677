"""tcgen05 / TMEM path (p = 64): building blocks against fp64, single rounds against the FFMA
kernels, and the whole parity suite re-run with the tensor path forced on (S2V_B200_PATH=tc)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ptr(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("rows", [64, 200, 1000])
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_tc_building_blocks(mode, rows, cuda_device):
    """3xTF32 UMMA with hand-built SWIZZLE_128B tiles: A W^T, A W (K-major operands, M = 64) and
    A^T B over the rows (MN-major operands, stacked M = 128, accumulated in TMEM across tiles); mode 3 feeds A from
    TMEM."""
    from simmodel_b200 import _lib
    L = _lib.lib()
    dev = cuda_device
    g = torch.Generator().manual_seed(rows * 3 + mode)
    A = (torch.randn(rows, 64, generator=g) * torch.logspace(-2, 2, 64)).to(dev)
    B = torch.randn(64 if mode != 2 else rows, 64, generator=g).to(dev)
    out = torch.full((rows if mode != 2 else 64, 64), float("nan"), device=dev)
    _lib.check(L.s2v_debug_tc_gemm(mode, _ptr(A), _ptr(B), _ptr(out), rows, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
               "s2v_debug_tc_gemm")
    torch.cuda.synchronize()
    A64, B64 = A.double(), B.double()
    ref = A64 @ B64.t() if mode == 0 else A64 @ B64 if mode in (1, 3) else A64.t() @ B64
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    ref32 = (A @ B.t() if mode == 0 else A @ B if mode in (1, 3) else A.t() @ B).double()
    err32 = float((ref32 - ref).abs().max() / ref.abs().max())
    assert err < 2e-6, (mode, rows, err, err32)


@pytest.mark.parametrize("rows", [64, 200, 1000])
@pytest.mark.parametrize("mode", [6, 7])
def test_bf16x3_building_blocks(mode, rows, cuda_device):
    """bf16x3 split-precision UMMAs (six products): A W with K-major views (mode 6) and A^T B over the rows with
    MN-major views of the SAME tile layout (mode 7)."""
    from simmodel_b200 import _lib
    L = _lib.lib()
    dev = cuda_device
    g = torch.Generator().manual_seed(rows * 5 + mode)
    A = (torch.randn(rows, 64, generator=g) * torch.logspace(-2, 2, 64)).to(dev)
    B = torch.randn(64 if mode == 6 else rows, 64, generator=g).to(dev)
    out = torch.full((rows if mode == 6 else 64, 64), float("nan"), device=dev)
    _lib.check(L.s2v_debug_tc_gemm(mode, _ptr(A), _ptr(B), _ptr(out), rows, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
               "s2v_debug_tc_gemm")
    torch.cuda.synchronize()
    A64, B64 = A.double(), B.double()
    ref = A64 @ B64 if mode == 6 else A64.t() @ B64
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    assert err < 1e-6, (mode, rows, err)


@pytest.mark.parametrize("topo_name,B", [("NWB_AMS", 9), ("Manhattan5", 7), ("NWB_AMS", 64)])
def test_single_round_tc_vs_ffma(topo_name, B, cuda_device):
    import simmodel_b200 as sb
    from simmodel_b200 import _lib
    L = _lib.lib()
    dev = cuda_device
    topo = sb.graphs.load_topology(topo_name)
    n = B * topo.num_nodes
    lay = sb.GraphLayout.from_batch(sb.graphs.batch_edge_index(topo, B).to(dev),
                                    (torch.arange(B + 1) * topo.num_nodes).to(dev), topo.num_nodes)
    gen = torch.Generator().manual_seed(3)
    w2 = (torch.randn(64, 64, generator=gen) / 8).to(dev)
    base = torch.randn(n, 64, generator=gen).to(dev)
    mu = torch.randn(n, 64, generator=gen).clamp(min=0).to(dev)
    du = torch.randn(n, 64, generator=gen).to(dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = lay.c_struct()
    outs = {}
    for flag in (1, 2, 3, 4):
        o = torch.empty(n, 64, device=dev)
        _lib.check(L.s2v_embed_round_forward(C.byref(g), 64, flag, _ptr(w2), _ptr(base), _ptr(mu), _ptr(o), st), "fwd")
        d = torch.empty(n, 64, device=dev)
        ws = torch.zeros(L.s2v_embed_backward_workspace_bytes(7, 64), dtype=torch.uint8, device=dev)
        _lib.check(L.s2v_embed_round_backward(C.byref(g), 7, 64, flag, _ptr(w2), _ptr(du), _ptr(mu), _ptr(d), _ptr(ws),
                                              ws.numel(), st), "bwd")
        torch.cuda.synchronize()
        stride = 2 * 64 * 64 + 4 * 64 + 64 * 7
        part = ws.view(torch.float32)[: 1024 * stride].view(1024, stride)[:, : 64 * 64].sum(0)
        outs[flag] = (o, d, part)
    for flag in (2, 3, 4):                   # 2: 8-warp CTAs (3xTF32), 3: warp-specialised pipeline (bf16x3), 4: forward-shaped bf16x3 round
        for a, b, name in zip(outs[1], outs[flag], ("mu_next", "du_prev", "dtheta2")):
            err = float((a - b).abs().max() / a.abs().max())
            assert err < 3e-6, (flag, name, err)
        # exact zero pattern of the ReLU mask must agree
        assert torch.equal(outs[1][1] == 0, outs[flag][1] == 0)


@pytest.mark.parametrize("path", ["tc", "tcws", "tc16", "tc16s"])
def test_parity_suite_with_tensor_path_forced(path, cuda_device):
    """Every p = 64 parity case again with the tcgen05 kernels forced on for all sizes (tc: 8-warp 3xTF32 backward
    round, tcws: warp-specialised bf16x3 backward round -- the one auto picks for large batches, tc16 / tc16s: the
    forward-shaped bf16x3 backward round, 8-warp CTAs with the MMAs behind the next gather / 4-warp CTAs)."""
    env = dict(os.environ, S2V_B200_PATH="tc16" if path == "tc16s" else path)
    if path == "tc16s":
        env["S2V_B200_BWD16_KERNEL"] = "s"
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu",
                          "-q", "-x", "--tb=short", "-p", "no:cacheprovider"], env=env, capture_output=True, text=True,
                         timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-1000:]


@pytest.mark.parametrize("topo_name,B,kind,cfg", [("NWB_AMS", 37, "rep", "1,0,256"), ("NWB_AMS", 21, "bat", "1,0,64"),
                                                  ("NWB_UTR", 9, "rep", "2,0,128"), ("Manhattan5", 300, "bat", "1,200,96")])
def test_ring_forward_round_bit_identical(topo_name, B, kind, cfg, cuda_device):
    """The asynchronous-gather forward round (k_embed_round_ring: LDGSTS ring + cp.async.bulk base tiles + tcgen05, opt-in with
    S2V_B200_FWD_RING) computes the same bits as k_embed_round_tc16: same CSR-order sums, same split, same MMA order, same
    epilogue.  Replicated and batched layouts, one and several chunks per tile, ragged last tile, tiles that straddle graphs."""
    import simmodel_b200 as sb
    from simmodel_b200 import _lib
    from simmodel_b200.layout import _p, _stream
    L = _lib.lib()
    dev = cuda_device
    topo = sb.graphs.load_topology(topo_name)
    n = topo.num_nodes
    N = B * n
    if kind == "rep":
        lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, B)
    else:
        ptr = (torch.arange(B + 1, dtype=torch.int64) * n).to(dev)
        lay = sb.GraphLayout.from_batch(sb.graphs.batch_edge_index(topo, B, device=dev), ptr, n)
    gen = torch.Generator().manual_seed(B)
    w2 = (torch.randn(64, 64, generator=gen) / 8).to(dev)
    base = torch.randn(N, 64, generator=gen).to(dev)
    mu = torch.rand(N, 64, generator=gen).to(dev)
    g = lay.c_struct()

    def run(out):
        _lib.check(L.s2v_embed_round_forward(C.byref(g), 64, 3, _p(w2), _p(base), _p(mu), _p(out), _stream()), "round_forward")
        torch.cuda.synchronize()

    old = os.environ.pop("S2V_B200_FWD_RING", None)
    old_round = os.environ.get("S2V_B200_FWD_ROUND")
    os.environ["S2V_B200_FWD_ROUND"] = "tc16"          # the bf16x3 round the ring kernel reproduces (the default is fp16x2 now)
    try:
        ref = torch.full((N, 64), float("nan"), device=dev)
        run(ref)
        os.environ["S2V_B200_FWD_RING"] = cfg
        out = torch.full((N, 64), float("nan"), device=dev)
        run(out)
    finally:
        os.environ.pop("S2V_B200_FWD_RING", None)
        if old is not None:
            os.environ["S2V_B200_FWD_RING"] = old
        os.environ.pop("S2V_B200_FWD_ROUND", None)
        if old_round is not None:
            os.environ["S2V_B200_FWD_ROUND"] = old_round
    assert torch.equal(out, ref)
    agg = torch.zeros(N, 64, device=dev).index_add_(0, lay_edge_rows(lay, topo, B, dev)[0], mu[lay_edge_rows(lay, topo, B, dev)[1]])
    want = torch.relu(agg.double() @ w2.double().t() + base.double())
    assert float((out.double() - want).abs().max() / want.abs().max()) < 2e-6


def lay_edge_rows(lay, topo, B, dev):
    import simmodel_b200 as sb
    ei = sb.graphs.batch_edge_index(topo, B, device=dev)
    return ei[0], ei[1]


@pytest.mark.parametrize("T,topo_name,B", [(9, "NWB_AMS", 5), (2, "NWB_UTR", 4), (7, "Manhattan5", 170), (1, "NWB_AMS", 5)])
def test_last_stage_plane_ring_any_T(T, topo_name, B, cuda_device, monkeypatch):
    """The last backward stage takes the T gradient planes of a tile by cp.async.bulk into a 7-slot landing ring: more planes
    than slots (T = 9), fewer (T = 1, 2), ragged last tiles.  All 16 gradients of the tensor path against the fp64 oracle
    (1e-5 of scale, or twice the fp32 oracle's own error) and against the CUDA-core path."""
    import simmodel_b200 as sb
    from oracle import s2v_ref as R
    dev = cuda_device
    topo = sb.graphs.load_topology(topo_name)
    n = topo.num_nodes
    gen = torch.Generator().manual_seed(T * 31 + B)
    x = sb.graphs.synth_nfm(topo, B, gen).reshape(-1, sb.graphs.NODE_DIM)
    ei = sb.graphs.batch_edge_index(topo, B)
    ptr = torch.arange(B + 1, dtype=torch.int64) * n
    actions = torch.randint(0, n, (B,), generator=gen).tolist()
    targets = torch.randn(B, generator=gen).tolist()
    P0 = R.make_qnet_params(sb.graphs.NODE_DIM, 64, seed=T)

    def oracle(dtype):
        P = {k: v.to(dtype).clone().requires_grad_(True) for k, v in P0.items()}
        q = R.qnet_forward_sparse(P, x.to(dtype), ei, ptr, n, T, True)
        R.dqn_loss(q, actions, ptr, [float(v) for v in targets]).backward()
        return q.detach(), {k: v.grad for k, v in P.items()}

    q64, g64 = oracle(torch.float64)
    q32, g32 = oracle(torch.float32)

    def ours(path):
        monkeypatch.setenv("S2V_B200_PATH", path)
        torch.manual_seed(0)
        net = sb.QNet({"emb_dim": 64, "emb_iter_T": T, "norm_agg": True, "node_dim": sb.graphs.NODE_DIM}, verbose=False).to(dev)
        net.load_state_dict({k: v.to(dev) for k, v in P0.items()}, strict=True)
        lay = sb.GraphLayout.from_batch(ei.to(dev), ptr.to(dev), n)
        q = net.forward_graph(x.to(dev), lay)
        sel = torch.tensor(actions, device=dev) + ptr[:-1].to(dev)
        torch.nn.SmoothL1Loss()(q[sel], torch.tensor(targets, device=dev)).backward()
        torch.cuda.synchronize()
        return q.detach().cpu(), {k: p.grad.detach().cpu() for k, p in net.named_parameters()}

    q_tc, g_tc = ours("tcws")
    q_h2, g_h2 = ours("tc16")                  # fp16x2 rounds + the forward-shaped fp16x2 last stage (what auto runs at scale)
    q_ff, g_ff = ours("ffma")
    scale = float(q64.abs().max())
    assert float((q_tc.double() - q64).abs().max()) / scale < 1e-5
    assert float((q_h2.double() - q64).abs().max()) / scale < 1e-5
    for k in g64:
        s = float(g64[k].abs().max()) + 1e-30
        own = float((g32[k].double() - g64[k]).abs().max()) / s
        bound = max(1e-5, 2 * own)
        assert float((g_tc[k].double() - g64[k]).abs().max()) / s < bound, (k, T)
        assert float((g_h2[k].double() - g64[k]).abs().max()) / s < bound, (k, T, "h2")
        assert float((g_ff[k].double() - g64[k]).abs().max()) / s < bound, (k, T)


@pytest.mark.parametrize("du_scale,mu_scale,w_scale", [(1.0, 1.0, 1.0), (1e-30, 1.0, 1.0), (1e-8, 1e3, 1e-3), (1e25, 1e-6, 1.0),
                                                       (1.0, 1e-20, 1e12), (0.0, 1.0, 1.0)])
def test_fp16x2_rounds_dynamic_range(du_scale, mu_scale, w_scale, cuda_device):
    """The fp16x2 round kernels (two fp16 pieces per operand, per-tile power-of-two scales) against fp64 where fp16 alone would
    under- or overflow: gradients at 1e-30, activations at 1e-20 or 1e3, weights at 1e12, an all-zero gradient plane, and rows
    10^4 times larger than their tile mates.  Error bound: 2e-6 of the tensor's largest entry (measured: 2e-7); no inf / nan."""
    import simmodel_b200 as sb
    from simmodel_b200 import _lib
    L = _lib.lib()
    dev = cuda_device
    topo = sb.graphs.load_topology("NWB_UTR")
    B = 4
    n = B * topo.num_nodes
    ei = sb.graphs.batch_edge_index(topo, B).to(dev)
    lay = sb.GraphLayout.from_batch(ei, (torch.arange(B + 1) * topo.num_nodes).to(dev), topo.num_nodes)
    gen = torch.Generator().manual_seed(17)
    w2 = (torch.randn(64, 64, generator=gen) / 8 * w_scale).to(dev)
    big = (torch.rand(n, 1, generator=gen) < 0.01).float() * 1e4 + 1.0              # one row in a hundred is 10^4 times larger
    mu = (torch.randn(n, 64, generator=gen).clamp(min=0) * big * mu_scale).to(dev)
    du = (torch.randn(n, 64, generator=gen) * big.flip(0) * du_scale).to(dev)
    base = (torch.randn(n, 64, generator=gen) * mu_scale * w_scale).to(dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = lay.c_struct()
    o = torch.full((n, 64), float("nan"), device=dev)
    _lib.check(L.s2v_embed_round_forward(C.byref(g), 64, 4, _ptr(w2), _ptr(base), _ptr(mu), _ptr(o), st), "fwd")
    d = torch.full((n, 64), float("nan"), device=dev)
    ws = torch.zeros(L.s2v_embed_backward_workspace_bytes(7, 64), dtype=torch.uint8, device=dev)
    _lib.check(L.s2v_embed_round_backward(C.byref(g), 7, 64, 4, _ptr(w2), _ptr(du), _ptr(mu), _ptr(d), _ptr(ws), ws.numel(), st), "bwd")
    torch.cuda.synchronize()
    stride = 2 * 64 * 64 + 4 * 64 + 64 * 7
    dth = ws.view(torch.float32)[: 1024 * stride].view(1024, stride)[:, : 64 * 64].double().sum(0).view(64, 64)
    agg = torch.zeros(n, 64, dtype=torch.float64, device=dev).index_add_(0, ei[0], mu.double()[ei[1]])
    ref_o = (base.double() + agg @ w2.double().t()).clamp(min=0)
    g64 = torch.zeros(n, 64, dtype=torch.float64, device=dev).index_add_(0, ei[1], du.double()[ei[0]])
    ref_d = (g64 @ w2.double()) * (mu > 0)
    ref_t = g64.t() @ mu.double()
    for name, got, ref in (("mu_next", o.double(), ref_o), ("du_prev", d.double(), ref_d), ("dtheta2", dth, ref_t)):
        assert bool(torch.isfinite(got).all()), name
        scale = float(ref.abs().max())
        if scale == 0.0:
            assert float(got.abs().max()) == 0.0, name
        else:
            assert float((got - ref).abs().max()) / scale < 2e-6, (name, float((got - ref).abs().max()) / scale)
    assert torch.equal(d == 0, ref_d == 0)
