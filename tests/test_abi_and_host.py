"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares,
the host-side mirror keeps the reference's interface, and nothing computes without a GPU."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "s2v_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(s2v_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from simmodel_b200 import _lib, build
    build.build()
    syms = _header_symbols()
    assert len(syms) >= 20
    L = _lib.lib()
    for s in syms:
        assert hasattr(L, s), "missing export " + s
        assert s in _lib.SIGNATURES, "no ctypes signature for " + s
    assert set(_lib.SIGNATURES) == set(syms)
    assert L.s2v_abi_version() == 5
    assert L.s2v_embed_grad_floats(7, 64) + L.s2v_head_grad_floats(64) == 21569      # comb_opt.py QNet size
    assert b"workspace" in L.s2v_error_string(-3)


def test_abi_argument_validation_without_gpu():
    """Entry points reject null / inconsistent arguments before touching the device."""
    import ctypes as C
    from simmodel_b200 import _lib
    L = _lib.lib()
    g = _lib.Graph()
    prm = _lib.EmbedParams()
    assert L.s2v_embed_forward(C.byref(g), C.byref(prm), None, None, None, None) == -1
    assert L.s2v_csr_build(None, -1, 4, None, None, None, None, None, None, 0, None) == -1
    assert L.s2v_relu_mask(None, None, None, 8, None) == -1
    assert L.s2v_csr_workspace_bytes(1000, 5000) > (2 * 1000 + 2 * 5000) * 4


def test_no_product_import_of_oracle():
    """The product package must never import the oracle (or the reference)."""
    pkg = os.path.join(ROOT, "simmodel_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_qnet_interface_matches_reference():
    import simmodel_b200 as sb
    from oracle import s2v_ref as R
    torch.manual_seed(3)
    net = sb.QNet({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": 7}, verbose=False)
    keys = list(net.state_dict().keys())
    assert keys == [n + s for n in R.QNET_LINEARS for s in (".weight", ".bias")]
    assert net.numTrainableParameters(verbose=False) == 21569
    # same construction order => same default init as the reference constructor under one seed
    P = R.make_qnet_params(7, 64, seed=3)
    for k, v in net.state_dict().items():
        assert torch.equal(v, P[k]), k
    # flat gradient layout == parameters() order
    shapes = [tuple(p.shape) for p in net.parameters()]
    assert shapes == [(64, 7), (64,), (64, 64), (64,), (64, 64), (64,), (64, 64), (64,), (64, 1), (64,),
                      (1, 128), (1,), (64, 64), (64,), (64, 64), (64,)]
    import copy
    clone = copy.deepcopy(net)                 # target-network sync (comb_opt.py:597)
    clone.load_state_dict(net.state_dict(), strict=True)


def test_cpu_module_raises_instead_of_falling_back():
    import simmodel_b200 as sb
    net = sb.QNet({"emb_dim": 16, "emb_iter_T": 2, "norm_agg": True, "node_dim": 7}, verbose=False)
    b = sb.Batch.from_data_list([sb.Data(torch.ones(3), [0, 0], num_nodes=3)])
    with pytest.raises(sb.S2VError):
        net(torch.zeros(1, 3, 7), torch.zeros(1, 3, 3), b)


def test_batch_shim_semantics():
    import simmodel_b200 as sb
    d = [sb.Data(torch.zeros(2, 7), torch.tensor([[0], [1]])), sb.Data(torch.zeros(3, 7), torch.tensor([[0, 2], [1, 0]]))]
    b = sb.Batch.from_data_list(d)
    assert b.ptr.tolist() == [0, 2, 5] and b.batch.tolist() == [0, 0, 1, 1, 1] and b.num_graphs == 2
    assert b.edge_index.tolist() == [[0, 2, 4], [1, 3, 2]] and b.x.shape == (5, 7)
    dummy = sb.Batch.from_data_list([sb.Data(torch.ones(4), [0, 0], num_nodes=4)])      # comb_opt.py:472
    assert dummy.edge_index is None and dummy.ptr.tolist() == [0, 4]


def test_graph_fixtures():
    from simmodel_b200 import graphs
    ams = graphs.load_topology("NWB_AMS")
    assert (ams.num_nodes, ams.num_edges, len(ams.target_nodes)) == (975, 2850, 113)
    rot = graphs.load_topology("NWB_ROT")
    assert (rot.num_nodes, rot.num_edges) == (2602, 7266)
    m5 = graphs.load_topology("Manhattan5")
    assert (m5.num_nodes, m5.num_edges) == (25, 105)
    for t in (ams, rot, m5, graphs.load_topology("MemTask")):
        ei = t.edge_index.numpy()
        order = np.lexsort((ei[1], ei[0]))
        assert np.array_equal(order, np.arange(ei.shape[1]))          # row-major like W.nonzero()
    assert ams.neighbors(970) == []                                    # the isolated node
    x = graphs.synth_nfm(ams, 3, torch.Generator().manual_seed(0))
    assert x.shape == (3, 975, 7) and float(x[:, :, 5:7].sum()) == 3 * 10
    ei = graphs.batch_edge_index(m5, 2)
    assert ei.shape == (2, 210) and int(ei.max()) == 49


def test_shard_helpers():
    from simmodel_b200.dist import shard_graphs, shard_graphs_by_nodes
    cover = [shard_graphs(10, 4, r) for r in range(4)]
    assert cover == [(0, 3), (3, 6), (6, 8), (8, 10)]
    parts = shard_graphs_by_nodes([975, 25, 25, 975, 9, 9, 975, 25], 2)
    assert parts[0][0] == 0 and parts[-1][1] == 8 and parts[0][1] == parts[1][0]
    loads = [sum([975, 25, 25, 975, 9, 9, 975, 25][a:b]) for a, b in parts]
    assert max(loads) - min(loads) <= 975


def test_bench_reference_arm_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "graphs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_sb3_mirror_state_dict_keys():
    """SB3 consumer: parameter names of the mirrors == the reference's (golden produced by the
    unmodified modules/ppo/models_sb3_s2v.py classes)."""
    import simmodel_b200 as sb
    from conftest import load_golden
    meta, G = load_golden("sb3_m3m5mix_pad25")
    ex = sb.sb3.Struc2VecExtractor(meta["n_pad"], meta["emb_dim"], meta["T"], 7)
    ac = sb.sb3.s2v_ACNetwork(meta["emb_dim"], 1, 1, meta["emb_dim"])
    assert ["extractor." + k for k in ex.state_dict()] == [k for k in G["P"] if k.startswith("extractor.")]
    assert ["acnet." + k for k in ac.state_dict()] == [k for k in G["P"] if k.startswith("acnet.")]
    assert ex.features_dim == meta["n_pad"] * (meta["emb_dim"] + 1)


def test_ngo_mirror_state_dict_keys():
    """models_ngo mirror: parameter names == the unmodified reference's MaskablePPOPolicy (incl. the
    duplicated lin_r entries of the shared-weight GATv2 layers)."""
    import types
    import simmodel_b200 as sb
    from conftest import load_golden
    for case in ("ngo_m3m5_seq3_nolstm_q", "ngo_m5m3mem_seq2_lstm_q", "ngo_m3m5_seq1_nolstm_v"):
        meta, G = load_golden(case)
        hp = types.SimpleNamespace(emb_dim=64, node_dim=7, lstm_on=meta["lstm_on"], hidden_size=64, recurrent_layers=1,
                                   max_possible_nodes=meta["max_nodes"], critic=meta["critic"])
        pol = sb.ngo.MaskablePPOPolicy(None, meta["max_nodes"], False, False, init_log_std_dev=0.0, hp=hp)
        assert list(pol.state_dict().keys()) == list(G["P"].keys())
        total, _ = pol.numTrainableParameters()
        assert total == sum(p.numel() for p in pol.parameters() if p.requires_grad)


def test_torch_library_ops_registered_with_fake_impls():
    """The kernels are dispatcher-level custom ops (namespace s2v_b200) with schemas and fake (meta) implementations:
    shapes propagate without a GPU, which is what torch.compile / export need."""
    import torch
    import simmodel_b200 as sb
    from torch._subclasses.fake_tensor import FakeTensorMode
    for name in sb.torch_ops.OPS:
        assert hasattr(torch.ops.s2v_b200, name), name
    n, B, F, p, T, E = 50, 2, 7, 64, 5, 120
    with FakeTensorMode():
        i32 = dict(dtype=torch.int32)
        x = torch.empty(n, F)
        shapes = [(p, F), (p,), (p, p), (p,), (p, p), (p,), (p, p), (p,), (p, 1), (p,), (1, 2 * p), (1,), (p, p), (p,), (p, p), (p,)]
        w = [torch.empty(s) for s in shapes]
        rp, col = torch.empty(n + 1, **i32), torch.empty(E, **i32)
        ptr, gid = torch.empty(B + 1, **i32), torch.empty(n, **i32)
        q, base, mus, pool, gst = torch.ops.s2v_b200.qnet_fwd(x, w, T, True, rp, col, rp, col, torch.empty(n, **i32), ptr, gid, None,
                                                             n, B, 0, 25)
        assert q.shape == (n,) and base.shape == (n, p) and mus.shape == (T, n, p) and pool.shape == (B, p)
        flat = torch.ops.s2v_b200.qnet_bwd(q, x, w, base, mus, pool, gst, T, True, rp, col, rp, col, torch.empty(n, **i32), ptr,
                                           gid, None, n, B, 0, 25)
        import math
        assert flat.numel() == sum(math.prod(s) for s in shapes)
        outs = torch.ops.s2v_b200.csr_build(torch.empty(2, E, dtype=torch.int64), n)
        assert outs[0].shape == (n + 1,) and outs[1].shape == (E,)
        prob = torch.ops.s2v_b200.masked_softmax(q, torch.empty(n, dtype=torch.bool), ptr, B, 0)
        assert prob.shape == q.shape
        arg, val = torch.ops.s2v_b200.masked_argmax(q, torch.empty(n, dtype=torch.bool), ptr, B, 0)
        assert arg.shape == (B,) and arg.dtype == torch.int64
