"""Multi-process data parallelism on CPU (gloo, world_size 2): the sharding + flat gradient
all-reduce of simmodel_b200.dist reproduce the single-process global-batch gradient.
The per-rank gradients come from the CPU oracle (the kernels need a GPU); what is under test
is the host-side DP logic that bench.py and a trainer use unchanged with NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, sizes, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import s2v_ref as R
    from simmodel_b200.dist import allreduce_gradients, broadcast_parameters, shard_graphs_by_nodes
    import simmodel_b200 as sb
    torch.manual_seed(100 + rank)                    # ranks start with DIFFERENT weights ...
    net = sb.QNet({"emb_dim": 16, "emb_iter_T": 3, "norm_agg": True, "node_dim": 7}, verbose=False)
    broadcast_parameters(net, src=0)                 # ... and are made identical here
    lo, hi = shard_graphs_by_nodes(sizes, world)[rank]
    data = torch.load(os.path.join(out_dir, "data.pt"))
    P = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    ptr = torch.zeros(hi - lo + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(torch.tensor(sizes[lo:hi]), 0)
    x = torch.cat(data["x"][lo:hi])
    ei = torch.cat([e + int(ptr[i]) for i, e in enumerate(data["ei"][lo:hi])], dim=1)
    q = R.qnet_forward_sparse(P, x, ei, ptr, 12, 3, True)
    loss = R.dqn_loss(q, data["actions"][lo:hi], ptr, data["targets"][lo:hi])      # LOCAL mean
    loss.backward()
    for k, prm in net.named_parameters():
        prm.grad = P[k].grad.clone()
    flat = allreduce_gradients(net.parameters(), hi - lo, len(sizes))
    torch.save({"flat": flat, "state": net.state_dict(), "range": (lo, hi)}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.destroy_process_group()


def test_dp_two_ranks_match_global_batch(tmp_path):
    from oracle import s2v_ref as R
    rng = np.random.default_rng(0)
    sizes = [5, 12, 3, 9, 7, 11, 2]
    eis = [torch.from_numpy(np.stack(np.nonzero(rng.random((s, s)) < 0.3)).astype(np.int64)) for s in sizes]
    xs = [torch.from_numpy(rng.random((s, 7)).astype(np.float32)) for s in sizes]
    actions = [int(rng.integers(0, s)) for s in sizes]
    targets = rng.standard_normal(len(sizes)).astype(np.float32).tolist()
    torch.save({"x": xs, "ei": eis, "actions": actions, "targets": targets}, tmp_path / "data.pt")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), sizes, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    assert r0["range"][1] == r1["range"][0] and r0["range"][0] == 0 and r1["range"][1] == len(sizes)
    assert torch.equal(r0["flat"], r1["flat"])
    for k in r0["state"]:
        assert torch.equal(r0["state"][k], r1["state"][k])
    # single-process global batch
    P = {k: v.clone().requires_grad_(True) for k, v in r0["state"].items()}
    bt = R.batch_from_sizes(sizes, eis)
    q = R.qnet_forward_sparse(P, torch.cat(xs), bt.edge_index, bt.ptr, 12, 3, True)
    R.dqn_loss(q, actions, bt.ptr, targets).backward()
    ref = torch.cat([P[k].grad.reshape(-1) for k in r0["state"]])
    assert float((r0["flat"] - ref).abs().max()) <= 1e-6 * float(ref.abs().max())
