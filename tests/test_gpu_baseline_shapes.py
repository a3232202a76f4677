"""Parity of the BENCHMARKED arithmetic path at BASELINE.json shapes.

  * test_headline_shape_*       configs[2] as bench.py runs it: NWB_AMS, B = 4096 graphs (3 993 600 rows), p = 64,
                                T = 5, the default (`auto` = tcgen05 bf16x3, warp-specialised backward) kernels, DQN
                                loss -- q and all 16 parameter gradients against the sparse CPU oracle, adjudicated
                                in fp64 (bound: 1e-5 of the tensor scale, or 2x the fp32 oracle's own fp64 error).
  * test_dense_dmu_kinks_*      configs[3] shape: R rollouts x S = 50 steps of AMS, a gradient that reaches EVERY node
                                (dense dmu, what the PPO loss produces).  ReLU decisions are compared with the fp64
                                oracle one by one (count, distance of every differing element from the kink); with the
                                decisions held equal the gradients must agree to 1e-5.
  * layout / data-parallel      int32 + graph-local edge lists, topology tables, gradient of a batch = weighted sum of
                                the shard gradients through the flat in-place all-reduce buffer.
"""
import os

import numpy as np
import pytest
import torch

from test_gpu_parity import TOL, _check, _scale

pytestmark = pytest.mark.gpu


def _ams_batch(B, seed):
    import simmodel_b200 as sb
    topo = sb.graphs.load_topology("NWB_AMS")
    gen = torch.Generator().manual_seed(seed)
    x = sb.graphs.synth_nfm(topo, B, gen).reshape(-1, sb.graphs.NODE_DIM)
    actions = torch.randint(0, topo.num_nodes, (B,), generator=gen)
    targets = torch.randn(B, generator=gen)
    return topo, x, actions, targets


def test_headline_shape_vs_sparse_oracle_fp64(cuda_device, monkeypatch):
    """The bench.py workload itself.  The oracle runs on the host in chunks of 512 graphs (the gradient of the mean loss
    is the sum of the chunk gradients, weights 1/B), once in fp32 and once in fp64."""
    import simmodel_b200 as sb
    from oracle import s2v_ref as R
    monkeypatch.setenv("S2V_B200_PATH", "auto")
    dev = cuda_device
    B, T, CH = 4096, 5, 512
    topo, x, actions, targets = _ams_batch(B, 123)
    N = topo.num_nodes
    P = R.make_qnet_params(sb.graphs.NODE_DIM, 64, seed=0)            # default nn.Linear init, as bench.py
    net = sb.QNet({"emb_dim": 64, "emb_iter_T": T, "norm_agg": True, "node_dim": sb.graphs.NODE_DIM}, verbose=False)
    net.load_state_dict(P)
    net = net.to(dev)
    ptr = torch.arange(B + 1, dtype=torch.int64) * N
    layout = sb.GraphLayout.from_batch(sb.graphs.batch_edge_index(topo, B, device=dev), ptr.to(dev), N, validate=False)
    q = net.forward_graph(x.to(dev), layout)
    loss = torch.nn.SmoothL1Loss()(q[actions.to(dev) + ptr[:-1].to(dev)], targets.to(dev))
    loss.backward()
    q_gpu = q.detach().cpu().numpy()
    # oracle, chunked
    ei_c = sb.graphs.batch_edge_index(topo, CH)
    ptr_c = torch.arange(CH + 1, dtype=torch.int64) * N
    g32 = {k: np.zeros(tuple(v.shape), dtype=np.float64) for k, v in P.items()}
    g64 = {k: np.zeros(tuple(v.shape), dtype=np.float64) for k, v in P.items()}
    worst_q = 0.0
    for lo in range(0, B, CH):
        xs = x[lo * N:(lo + CH) * N]
        for dt, acc in ((torch.float32, g32), (torch.float64, g64)):
            Pc = {k: v.detach().clone().to(dt).requires_grad_(True) for k, v in P.items()}
            qc = R.qnet_forward_sparse(Pc, xs.to(dt), ei_c, ptr_c, N, T, True)
            sel = actions[lo:lo + CH] + ptr_c[:-1]
            # mean over the GLOBAL batch: chunk loss sums weighted 1/B
            (torch.nn.functional.smooth_l1_loss(qc[sel], targets[lo:lo + CH].to(dt), reduction="sum") / B).backward()
            for k in acc:
                acc[k] += Pc[k].grad.double().numpy()
            if dt == torch.float64:
                q64 = qc.detach().numpy()
            else:
                q32 = qc.detach().numpy()
        worst_q = max(worst_q, _check("q[%d:%d]" % (lo, lo + CH), q_gpu[lo * N:(lo + CH) * N], q32, q64))
    for k, prm in net.named_parameters():
        _check("grad " + k, prm.grad.cpu().numpy(), g32[k], g64[k])


@pytest.mark.parametrize("path", ["auto", "ffma"])
def test_dense_dmu_kinks_counted_and_excluded(path, cuda_device, monkeypatch):
    """configs[3] shape (AMS, S = 50, R = 2 -> 97 500 rows; every node receives a gradient).  (i) every ReLU decision
    of the five rounds is compared with the fp64 oracle: at most K_MAX of 31 M may differ and each of those must lie
    within fp32 resolution of the kink; (ii) with the decisions prescribed (oracle/s2v_ref.py:
    s2v_embed_sparse_masked) all embedding-parameter gradients agree to 1e-5 of scale (or 2x the fp32 oracle's own
    error against fp64) -- for positive and for random-sign weights."""
    import simmodel_b200 as sb
    from simmodel_b200 import ops
    from oracle import s2v_ref as R
    monkeypatch.setenv("S2V_B200_PATH", path)
    dev = cuda_device
    S, Rr, T = 50, 2, 5
    B = S * Rr
    topo, x, _, _ = _ams_batch(B, 321)
    N, n = topo.num_nodes, B * topo.num_nodes
    P = R.make_qnet_params(sb.graphs.NODE_DIM, 64, seed=4)
    names = [k for k in P if k.split(".")[0] in ("theta1a", "theta1b", "theta2", "theta3", "theta4")]
    w_dev = [P[k].detach().clone().to(dev).requires_grad_(True) for k in names]
    layout = sb.GraphLayout.replicated(topo.edge_index.to(dev), N, B)     # the PPO path: S x R copies of one topology
    ei = sb.graphs.batch_edge_index(topo, B)
    npad = torch.full((n,), N, dtype=torch.int64)
    gen = torch.Generator().manual_seed(9)
    K_MAX = 64
    # The first-stage ReLU h = relu(theta1a x + b1a) is recomputed inside the kernels and never exported, so its
    # decisions cannot be read back: rows with a first-stage pre-activation within 1e-5 of scale of its kink get their
    # (continuous) score feature re-drawn until none is left -- what is under test here are the kinks of the T rounds.
    w1a, b1a = P["theta1a.weight"].double(), P["theta1a.bias"].double()
    for _ in range(20):
        h64 = torch.nn.functional.linear(x.double(), w1a, b1a)
        bad = (h64.abs() < 1e-5 * float(h64.abs().max())).any(dim=1)
        if not bad.any():
            break
        x[bad, 3] = torch.rand(int(bad.sum()), generator=gen)
    assert not bad.any()
    with torch.no_grad():
        _, mus = ops._embed_forward(layout, T, x.to(dev), [t.detach() for t in w_dev])
    masks = [(mus[t] > 0).cpu() for t in range(T)]
    for signed in (False, True):
        wt = torch.randn(n, 64, generator=gen)
        if not signed:
            wt = wt.abs()
        for t_ in w_dev:
            t_.grad = None
        mu = ops.s2v_embed(layout, x.to(dev), w_dev, T)
        (mu * wt.to(dev)).sum().backward()
        res = {}
        for dt in (torch.float32, torch.float64):
            Pc = {k: P[k].detach().clone().to(dt).requires_grad_(True) for k in names}
            mu_o, zs = R.s2v_embed_sparse_masked(Pc, x.to(dt), ei, npad, T, masks)
            (mu_o * wt.to(dt)).sum().backward()
            res[dt] = ({k: Pc[k].grad.numpy() for k in names}, [z.detach() for z in zs], mu_o.detach())
        g32, _, _ = res[torch.float32]
        g64, zs64, mu64 = res[torch.float64]
        # (i) decisions
        flips, worst = 0, 0.0
        for t in range(T):
            diff = (zs64[t] > 0) != masks[t]
            flips += int(diff.sum())
            if diff.any():
                worst = max(worst, float(zs64[t][diff].abs().max()) / float(zs64[t].abs().max()))
        print("    [kinks] path %s: %d of %d ReLU decisions differ from the fp64 oracle, farthest from the kink: %.1e of scale"
              % (path, flips, T * n * 64, worst))
        assert flips <= K_MAX, flips
        assert worst <= 2e-6, worst            # every differing decision is within fp32 resolution of zero
        # (ii) values and gradients with equal decisions
        _check("mu^T", mus[T - 1].cpu().numpy(), res[torch.float32][2].numpy(), mu64.numpy())
        for k, t_ in zip(names, w_dev):
            _check("grad %s (%s)" % (k, "signed" if signed else "positive"), t_.grad.cpu().numpy(), g32[k], g64[k])


@pytest.mark.parametrize("dtype", [torch.int32, torch.int64])
def test_local_edge_lists_and_topology_tables_bit_exact(dtype, cuda_device):
    """int32 / graph-local edge lists (offsets added inside the CSR kernels) and topology tables build the same CSR,
    bit for bit, as the int64 batch-wide edge_index."""
    import simmodel_b200 as sb
    dev = cuda_device
    rng = np.random.default_rng(5)
    tops = [sb.graphs.load_topology(nm) for nm in ("Manhattan5", "MemTask", "NWB_UTR")]
    tog = rng.integers(0, len(tops), size=37)
    sizes = [tops[t].num_nodes for t in tog]
    ptr = torch.zeros(len(tog) + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(torch.tensor(sizes), 0)
    eptr = torch.zeros(len(tog) + 1, dtype=torch.int64)
    eptr[1:] = torch.cumsum(torch.tensor([tops[t].num_edges for t in tog]), 0)
    ei_local = torch.cat([tops[t].edge_index for t in tog], dim=1)
    ei_global = torch.cat([tops[t].edge_index + int(ptr[g]) for g, t in enumerate(tog)], dim=1)
    ref = sb.GraphLayout.from_batch(ei_global.to(dev), ptr.to(dev), 1200)
    cands = [sb.GraphLayout.from_batch(ei_local.to(dtype).to(dev), ptr.to(dev), 1200, eptr=eptr.to(dev)),
             sb.GraphLayout.from_batch(ei_global.to(dtype).to(dev), ptr.to(dev), 1200),
             sb.GraphLayout.from_topologies([(tp.edge_index.to(dtype).to(dev), tp.num_nodes) for tp in tops],
                                            torch.from_numpy(tog), 1200)]
    for lay in cands:
        assert lay.num_nodes == ref.num_nodes and lay.num_graphs == ref.num_graphs and lay.period == 0
        for a in ("rowptr", "col", "rowptr_t", "col_t", "indeg", "ptr", "gid"):
            assert torch.equal(getattr(lay, a), getattr(ref, a)), a
    one = sb.GraphLayout.from_topologies([(tops[2].edge_index.to(dtype).to(dev), tops[2].num_nodes)], torch.zeros(9))
    assert one.period == tops[2].num_nodes and one.num_graphs == 9
    with pytest.raises(ValueError):
        sb.GraphLayout.from_batch((ei_local + 5000).to(dtype).to(dev), ptr.to(dev), 1200, eptr=eptr.to(dev))


def test_shard_gradients_sum_to_global_batch_one_gpu(cuda_device):
    """Data parallelism without a second GPU: the gradient of the global-batch mean loss equals the
    local_count/global_count-weighted sum of the shard gradients, taken from the flat [embed | head] buffer the
    kernels write (the buffer dist.allreduce_gradients reduces in place)."""
    import simmodel_b200 as sb
    from simmodel_b200.dist import shard_graphs, shared_flat_grad
    dev = cuda_device
    B, world = 48, 3
    topo, x, actions, targets = _ams_batch(B, 77)
    N = topo.num_nodes
    torch.manual_seed(3)
    net = sb.QNet({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": sb.graphs.NODE_DIM}, verbose=False).to(dev)
    params = list(net.parameters())
    loss_fn = torch.nn.SmoothL1Loss()

    def flat_grad(lo, hi):
        net.zero_grad(set_to_none=True)
        lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), N, hi - lo)
        q = net.forward_graph(x[lo * N:hi * N].to(dev), lay)
        sel = actions[lo:hi].to(dev) + torch.arange(hi - lo, device=dev) * N
        loss_fn(q[sel], targets[lo:hi].to(dev)).backward()
        flat = shared_flat_grad(params)
        assert flat is not None and flat.numel() == sum(p.numel() for p in params), "gradients must alias ONE flat buffer"
        assert all(p.grad.data_ptr() >= flat.data_ptr() for p in params)
        return flat.double().clone()

    g_all = flat_grad(0, B)
    acc = torch.zeros_like(g_all)
    for r in range(world):
        lo, hi = shard_graphs(B, world, r)
        acc += flat_grad(lo, hi) * ((hi - lo) / B)
    assert float((acc - g_all).abs().max()) <= 1e-6 * float(g_all.abs().max())


def _dp_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import simmodel_b200 as sb
    from simmodel_b200.dist import allreduce_gradients, shard_graphs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    B = 64
    topo, x, actions, targets = _ams_batch(B, 77)
    N = topo.num_nodes
    torch.manual_seed(3)
    net = sb.QNet({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": sb.graphs.NODE_DIM}, verbose=False).to(dev)
    lo, hi = shard_graphs(B, world, rank)
    lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), N, hi - lo)
    q = net.forward_graph(x[lo * N:hi * N].to(dev), lay)
    sel = actions[lo:hi].to(dev) + torch.arange(hi - lo, device=dev) * N
    torch.nn.SmoothL1Loss()(q[sel], targets[lo:hi].to(dev)).backward()
    flat = allreduce_gradients(net.parameters(), hi - lo, B)
    torch.cuda.synchronize()
    torch.save(flat.cpu(), os.path.join(out_dir, "flat%d.pt" % rank))
    dist.destroy_process_group()


def test_two_gpu_step_equals_single_gpu_global_batch(cuda_device, tmp_path):
    """Two ranks on two GPUs (NCCL): kernels + sharding + in-place flat all-reduce reproduce the single-GPU gradient of
    the global batch to 1e-6.  Skipped on a one-GPU box (the driver's test box); run with `gpurun --gpus 2`."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import socket
    import torch.multiprocessing as mp
    import simmodel_b200 as sb
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    f0, f1 = torch.load(tmp_path / "flat0.pt"), torch.load(tmp_path / "flat1.pt")
    assert torch.equal(f0, f1)
    dev = cuda_device
    B = 64
    topo, x, actions, targets = _ams_batch(B, 77)
    N = topo.num_nodes
    torch.manual_seed(3)
    net = sb.QNet({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": sb.graphs.NODE_DIM}, verbose=False).to(dev)
    lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), N, B)
    q = net.forward_graph(x.to(dev), lay)
    sel = actions.to(dev) + torch.arange(B, device=dev) * N
    torch.nn.SmoothL1Loss()(q[sel], targets.to(dev)).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).cpu()
    assert float((f0 - ref).abs().max()) <= 1e-6 * float(ref.abs().max())


def test_csr_hub_rows_and_duplicate_edges(cuda_device):
    """A star graph with a 50 000-degree hub (both directions: one heavy row and one heavy column) builds the same CSR as
    the stable-sort oracle, in bounded time (CTA bitonic sort for rows with more than 32 entries); duplicate edges --
    which the reference would SUM into a weight of 2 (to_dense_adj) -- are rejected, for light and for heavy rows."""
    import time
    import simmodel_b200 as sb
    from oracle import s2v_ref as R
    dev = cuda_device
    n = 50001
    rng = np.random.default_rng(0)
    leaves = np.arange(1, n)
    star = np.concatenate([np.stack([np.zeros(n - 1, dtype=np.int64), leaves]), np.stack([leaves, np.zeros(n - 1, dtype=np.int64)])], axis=1)
    extra = np.stack([rng.integers(1, n, 5000), rng.integers(1, n, 5000)])
    extra = extra[:, extra[0] != extra[1]]
    ei = np.concatenate([star, extra], axis=1)
    ei = np.unique(ei, axis=1)                                   # no duplicates
    ei = ei[:, rng.permutation(ei.shape[1])]                     # arbitrary input order: the CSR must keep it per row
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rowptr, col, rowptr_t, col_t, indeg = sb.GraphLayout.build_csr(torch.from_numpy(ei).to(dev), n)
    torch.cuda.synchronize()
    assert time.perf_counter() - t0 < 5.0
    ref = R.csr_from_edge_index(ei, n)
    for got, want in zip((rowptr, col, rowptr_t, col_t, indeg), ref):
        assert np.array_equal(got.cpu().numpy(), want)
    dup_light = np.concatenate([ei, np.array([[7], [9]]), np.array([[7], [9]])], axis=1)
    with pytest.raises(ValueError, match="duplicate"):
        sb.GraphLayout.build_csr(torch.from_numpy(dup_light).to(dev), n)
    dup_heavy = np.concatenate([ei, ei[:, ei[0] == 0][:, :1]], axis=1)      # a second copy of one hub edge
    with pytest.raises(ValueError, match="duplicate"):
        sb.GraphLayout.build_csr(torch.from_numpy(dup_heavy).to(dev), n)
