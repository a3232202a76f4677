"""Fused head launch (pool + theta6 + head + invalid-action mask + per-graph reduction, csrc/s2v_head.cu:k_head_fused)
against the three-launch sequence it replaces: every output bit-identical, gradients identical.
Reference ops replaced: models_basic_lstm.py:530-540 (pi), 548-564 (v), comb_opt.py:319-326 (greedy action)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(seed, p, dev):
    import simmodel_b200 as sb
    rng = np.random.default_rng(seed)
    sizes = rng.integers(1, 200, size=11).tolist() + [1, 64, 65, 128]
    n = int(sum(sizes))
    ptr = torch.zeros(len(sizes) + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(torch.tensor(sizes), 0)
    ei = torch.zeros(2, 0, dtype=torch.int64)
    layout = sb.GraphLayout.from_batch(ei.to(dev), ptr.to(dev), max(sizes))
    gen = torch.Generator().manual_seed(seed)
    mu = torch.randn(n, p, generator=gen).to(dev)
    w = [torch.randn(1, 2 * p, generator=gen) / 4, torch.randn(1, generator=gen), torch.randn(p, p, generator=gen) / 8,
         torch.randn(p, generator=gen) / 8, torch.randn(p, p, generator=gen) / 8, torch.randn(p, generator=gen) / 8]
    w = [t.to(dev) for t in w]
    mask = torch.rand(n, generator=gen) < 0.3
    mask[int(ptr[3]):int(ptr[4])] = False                       # a graph without any reachable node
    mask[int(ptr[5])] = True
    return layout, mu, w, mask.to(dev), ptr, sizes


@pytest.mark.parametrize("p", [16, 64])
@pytest.mark.parametrize("norm_agg", [True, False])
def test_fused_head_bit_identical_to_three_launches(p, norm_agg, cuda_device):
    import simmodel_b200 as sb
    from simmodel_b200 import ops
    dev = cuda_device
    layout, mu, w, mask, ptr, sizes = _case(3 + p, p, dev)
    assert ops.head_fused_eligible(layout, w)
    m8 = ops._mask_u8(mask, layout.num_nodes)
    pool0, g0, q0 = ops._head_forward(layout, norm_agg, mu, w)
    # mode q
    pool, gst, q = ops._head_fused_forward(layout, norm_agg, mu, w, "q")[:3]
    assert torch.equal(pool, pool0) and torch.equal(gst, g0) and torch.equal(q, q0)
    # softmax (NaN rows of the all-masked graph compare equal as NaN)
    prob0 = ops.masked_softmax(layout, q0, mask)
    prob = ops._head_fused_forward(layout, norm_agg, mu, w, "softmax", m8)[3]
    assert torch.equal(torch.isnan(prob), torch.isnan(prob0))
    assert torch.equal(torch.nan_to_num(prob), torch.nan_to_num(prob0))
    # masked max / argmax
    arg0, val0 = ops.masked_argmax(layout, q0, mask)
    _, _, _, _, arg, _, val = ops._head_fused_forward(layout, norm_agg, mu, w, "max", m8)
    assert torch.equal(arg, arg0) and torch.equal(val, val0)
    assert int(arg[3]) == -1 and float(val[3]) == float("-inf")
    # neighbour-list arg-max
    rng = np.random.default_rng(1)
    lists = [np.sort(rng.choice(s, size=rng.integers(0, min(s, 6) + 1), replace=False)) for s in sizes]
    rptr = torch.tensor(np.concatenate([[0], np.cumsum([len(l) for l in lists])]), dtype=torch.int32, device=dev)
    ridx = torch.tensor(np.concatenate(lists) if sum(len(l) for l in lists) else np.zeros(0), dtype=torch.int32, device=dev)
    pos0, node0, v0 = ops.list_argmax(layout, q0, rptr, ridx)
    pos, node, v, qq = ops.s2v_head_list_argmax(layout, mu, w, rptr, ridx, norm_agg)
    assert torch.equal(pos, pos0) and torch.equal(node, node0) and torch.equal(v, v0) and torch.equal(qq, q0)


def test_fused_head_gradients_equal_unfused(cuda_device):
    from simmodel_b200 import ops
    dev = cuda_device
    layout, mu, w, mask, ptr, sizes = _case(11, 64, dev)
    keep = torch.ones(layout.num_graphs, dtype=torch.bool, device=dev)
    keep[3] = False                                              # the all-masked graph is NaN on both sides

    def run(fused):
        mu_ = mu.clone().requires_grad_(True)
        w_ = [t.clone().requires_grad_(True) for t in w]
        if fused:
            prob = ops.s2v_head_softmax(layout, mu_, w_, mask)
            val, _ = ops.s2v_head_max(layout, mu_, w_, mask)
        else:
            prob = ops.masked_softmax(layout, ops.s2v_head(layout, mu_, w_, True), mask)
            val, _ = ops.masked_max(layout, ops.s2v_head(layout, mu_, w_, True), mask)
        gid = layout.gid.long()
        wts = torch.linspace(0.5, 1.5, layout.num_nodes, device=dev)
        loss = (prob * wts)[keep[gid]].sum() + val[keep].sum()
        loss.backward()
        return [mu_.grad] + [t.grad for t in w_]
    a, b = run(True), run(False)
    for x, y in zip(a, b):
        assert torch.equal(torch.nan_to_num(x), torch.nan_to_num(y))


def test_ppo_calls_use_one_head_launch(cuda_device):
    """pi(), v() and greedy_action() of the PPO model: exactly one k_head_fused launch and no stand-alone pool / head /
    mask kernel in the forward pass (counted from a CUPTI trace when the profiler is available)."""
    import types
    import simmodel_b200 as sb
    dev = cuda_device
    try:
        from torch.profiler import ProfilerActivity, profile
    except Exception:
        pytest.skip("torch.profiler unavailable")
    topo = sb.graphs.load_topology("Manhattan5")
    hp = types.SimpleNamespace(node_dim=sb.graphs.NODE_DIM, emb_dim=64, parallel_rollouts=1, batch_size=1, learning_rate=1e-3)
    model = sb.PPO_GNN_Single_LSTM({"lstm_type": "None", "qnet": "s2v", "critic": "q", "emb_dim": 64, "emb_iterT": 5}, hp,
                                   {"base_checkpoint_path": None}, device=dev)
    gen = torch.Generator().manual_seed(0)
    nfm = sb.graphs.synth_nfm(topo, 6, gen)
    reach = torch.rand(6, topo.num_nodes, generator=gen) < 0.4
    reach[:, 0] = True
    ei = topo.edge_index
    for fn in (model.pi, model.v, model.greedy_action):
        with torch.no_grad():
            fn(nfm, ei, reach, None)
            torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                fn(nfm, ei, reach, None)
                torch.cuda.synchronize()
        names = [e.name for e in prof.events() if "cuda" in str(getattr(e, "device_type", "")).lower()]
        if not names:
            pytest.skip("empty CUPTI trace")
        fused = [nm for nm in names if "k_head_fused" in nm]
        apart = [nm for nm in names if any(k in nm for k in ("k_pool", "k_head<", "k_head_tc", "k_masked", "k_list_argmax"))]
        assert len(fused) == 1 and not apart, (fn.__name__, names)


@pytest.mark.parametrize("topo_name,B,p", [("Manhattan5", 32, 64), ("MemTask", 50, 64), ("Manhattan3", 7, 64),
                                           ("MemTask", 1, 16), ("Manhattan5", 1, 32)])
def test_small_graph_single_launch_bit_identical(topo_name, B, p, cuda_device):
    """One cooperative launch (csrc/s2v_small.cu) == per-stage launches, bit for bit: embeddings of every round, pool,
    q, probabilities, arg-max; batched and replicated layouts; also when replayed from a CUDA graph."""
    import simmodel_b200 as sb
    from simmodel_b200 import ops
    from oracle import s2v_ref as R
    dev = cuda_device
    topo = sb.graphs.load_topology(topo_name)
    N, T = topo.num_nodes, 4
    gen = torch.Generator().manual_seed(B + p)
    x = sb.graphs.synth_nfm(topo, B, gen).reshape(-1, sb.graphs.NODE_DIM).to(dev)
    P = R.make_qnet_params(sb.graphs.NODE_DIM, p, seed=p)
    ew = [P[k + s].to(dev) for k in ops.EMBED_NAMES for s in (".weight", ".bias")]
    hw = [P[k + s].to(dev) for k in ops.HEAD_NAMES for s in (".weight", ".bias")]
    mask = (torch.rand(B * N, generator=gen) < 0.3).to(dev)
    mask[::N] = True
    m8 = ops._mask_u8(mask, B * N)
    ptr = torch.arange(B + 1, dtype=torch.int64) * N
    layouts = [sb.GraphLayout.replicated(topo.edge_index.to(dev), N, B, n_pad=N + 3),
               sb.GraphLayout.from_batch(sb.graphs.batch_edge_index(topo, B, device=dev), ptr.to(dev), N + 3)]
    big = sb.GraphLayout.replicated(sb.graphs.load_topology("NWB_AMS").edge_index.to(dev), 975, 2)
    assert not ops.head_fused_eligible(big, hw)              # few large graphs: the head is tile-parallel (next test)
    for lay in layouts:
        assert ops.qnet_small_eligible(lay, ew)
        base0, mus0 = ops._embed_forward(lay, T, x, ew)
        pool0, g0, q0, prob0 = ops._head_fused_forward(lay, True, mus0[T - 1], hw, "softmax", m8)[:4]
        arg0, val0 = ops._head_fused_forward(lay, True, mus0[T - 1], hw, "max", m8)[4::2]
        base, mus, pool, gst, q, prob = ops._qnet_small_forward(lay, T, x, ew, hw, True, "softmax", m8)[:6]
        for a, b in ((base, base0), (mus, mus0), (pool, pool0), (gst, g0), (q, q0), (prob, prob0)):
            assert torch.equal(a, b)
        out = ops._qnet_small_forward(lay, T, x, ew, hw, True, "max", m8)
        assert torch.equal(out[6], arg0) and torch.equal(out[8], val0)
        assert torch.equal(ops._qnet_small_forward(lay, T, x, ew, None, True, None)[1], mus0)
    # CUDA-graph replay of the cooperative launch
    lay = layouts[0]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ops.s2v_qnet_forward(lay, x, ew, hw, T)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        qg = ops.s2v_qnet_forward(lay, x, ew, hw, T)
    qg.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(qg, q0)


@pytest.mark.parametrize("B,p,layout_kind", [(1, 64, "rep"), (3, 64, "rep"), (2, 64, "bat"), (1, 32, "rep"), (2, 16, "var")])
def test_single_launch_large_graphs_bit_identical(B, p, layout_kind, cuda_device):
    """Acting on a road network (AMS, 975 rows per graph, B = 1..3): the cooperative launch runs the head as three
    tile-parallel stages (pool + theta6 | q per row tile | mask + reduction) behind grid barriers.  Everything equals the
    per-stage launches (s2v_embed_forward, s2v_head_forward, masked softmax / max, list arg-max) bit for bit."""
    import simmodel_b200 as sb
    from simmodel_b200 import ops
    from oracle import s2v_ref as R
    dev = cuda_device
    ams = sb.graphs.load_topology("NWB_AMS")
    T = 5
    gen = torch.Generator().manual_seed(7 * B + p)
    if layout_kind == "var":                                  # graphs of different sizes in one batch (AMS + UTR)
        utr = sb.graphs.load_topology("NWB_UTR")
        topos = [ams, utr][:B]
        sizes = [tp.num_nodes for tp in topos]
        ptr = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int64)
        ei = torch.cat([tp.edge_index + int(ptr[i]) for i, tp in enumerate(topos)], dim=1)
        lay = sb.GraphLayout.from_batch(ei.to(dev), ptr.to(dev), max(sizes) + 2)
        x = torch.cat([sb.graphs.synth_nfm(tp, 1, gen).reshape(-1, sb.graphs.NODE_DIM) for tp in topos]).to(dev)
    else:
        N = ams.num_nodes
        ptr = torch.arange(B + 1, dtype=torch.int64) * N
        lay = (sb.GraphLayout.replicated(ams.edge_index.to(dev), N, B) if layout_kind == "rep" else
               sb.GraphLayout.from_batch(sb.graphs.batch_edge_index(ams, B, device=dev), ptr.to(dev), N))
        x = sb.graphs.synth_nfm(ams, B, gen).reshape(-1, sb.graphs.NODE_DIM).to(dev)
    n = lay.num_nodes
    P = R.make_qnet_params(sb.graphs.NODE_DIM, p, seed=p)
    ew = [P[k + s].to(dev) for k in ops.EMBED_NAMES for s in (".weight", ".bias")]
    hw = [P[k + s].to(dev) for k in ops.HEAD_NAMES for s in (".weight", ".bias")]
    mask = (torch.rand(n, generator=gen) < 0.3).to(dev)
    mask[ptr[:-1].to(dev)] = True
    m8 = ops._mask_u8(mask, n)
    nb = len(ptr) - 1
    rptr = (torch.arange(nb + 1, dtype=torch.int32) * 3).to(dev)
    ridx = torch.tensor([5, 17, 2] * nb, dtype=torch.int32, device=dev)
    assert ops.qnet_small_eligible(lay, ew) and not ops.head_fused_eligible(lay, hw)
    base0, mus0 = ops._embed_forward(lay, T, x, ew)
    pool0, g0, q0 = ops._head_forward(lay, True, mus0[T - 1], hw)
    prob0 = ops.masked_softmax(lay, q0, mask)
    val0, arg0 = ops.masked_max(lay, q0, mask)
    lpos0, lnode0, lval0 = ops.list_argmax(lay, q0, rptr, ridx)
    base, mus, pool, gst, q, prob = ops._qnet_small_forward(lay, T, x, ew, hw, True, "softmax", m8)[:6]
    for a, b in ((base, base0), (mus, mus0), (pool, pool0), (gst, g0), (q, q0), (prob, prob0)):
        assert torch.equal(a, b)
    out = ops._qnet_small_forward(lay, T, x, ew, hw, True, "max", m8)
    assert torch.equal(out[4], q0) and torch.equal(out[6], arg0) and torch.equal(out[8], val0)
    out = ops._qnet_small_forward(lay, T, x, ew, hw, True, "list", None, rptr, ridx)
    assert torch.equal(out[4], q0) and torch.equal(out[6], lpos0) and torch.equal(out[7], lnode0) and torch.equal(out[8], lval0)
    assert torch.equal(ops._qnet_small_forward(lay, T, x, ew, hw, True, "q")[4], q0)
    # against the oracle (fp32 tolerance) -- the single launch is also RIGHT, not only equal to the other path
    if layout_kind != "var":
        ei_ref = sb.graphs.batch_edge_index(ams, B)
        q_ref = R.qnet_forward_sparse(P, x.cpu(), ei_ref, ptr, ams.num_nodes, T, True)
        assert float((q.cpu() - q_ref).abs().max() / q_ref.abs().max()) < 1e-5


def test_torch_library_ops_match_the_module_path(cuda_device):
    """torch.ops.s2v_b200.qnet_fwd (+ registered autograd) == QNet.forward_graph, values and all 16 gradients, and the op
    is traceable as one node by torch.compile(fullgraph=True) (eager backend: no code generation involved)."""
    import simmodel_b200 as sb
    from simmodel_b200 import torch_ops
    dev = cuda_device
    topo = sb.graphs.load_topology("NWB_UTR")
    B, T = 6, 5
    gen = torch.Generator().manual_seed(2)
    x = sb.graphs.synth_nfm(topo, B, gen).reshape(-1, sb.graphs.NODE_DIM).to(dev)
    torch.manual_seed(1)
    net = sb.QNet({"emb_dim": 64, "emb_iter_T": T, "norm_agg": True, "node_dim": sb.graphs.NODE_DIM}, verbose=False).to(dev)
    ptr = (torch.arange(B + 1) * topo.num_nodes).to(dev)
    lay = sb.GraphLayout.from_batch(sb.graphs.batch_edge_index(topo, B, device=dev), ptr, topo.num_nodes)
    wts = torch.randn(B * topo.num_nodes, generator=gen).to(dev)
    q0 = net.forward_graph(x, lay)
    (q0 * wts).sum().backward()
    g0 = [p.grad.clone() for p in net.parameters()]
    net.zero_grad()
    weights = net.embed_weights() + net.head_weights()
    q1 = torch_ops.qnet(x, weights, T, True, lay)
    (q1 * wts).sum().backward()
    assert torch.equal(q1, q0)
    for a, p in zip(g0, net.parameters()):
        assert torch.equal(a, p.grad)
    rp, col, rpt, colt, indeg = torch.ops.s2v_b200.csr_build(sb.graphs.batch_edge_index(topo, B, device=dev), B * topo.num_nodes)
    assert torch.equal(rp, lay.rowptr) and torch.equal(colt, lay.col_t)
    mask = (torch.rand(B * topo.num_nodes, generator=gen) < 0.2).to(dev)
    prob = torch.ops.s2v_b200.masked_softmax(q0.detach(), mask, lay.ptr, B, 0)
    assert torch.equal(prob, sb.ops.masked_softmax(lay, q0.detach(), mask))
    compiled = torch.compile(lambda xx: torch_ops.qnet(xx, weights, T, True, lay) * 2.0, backend="eager", fullgraph=True)
    with torch.no_grad():
        assert torch.equal(compiled(x), q0.detach() * 2.0)
