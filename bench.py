#!/usr/bin/env python
"""Benchmark of the s2v hot path:  python bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the configuration `metric` is quoted on):
s2v GNN-DQN training step on NWB_AMS-shaped graphs (N = 975 nodes, 2850 directed edges,
F = 7, p = 64, T = 5, mean pool), `--batch` graphs per GPU (weak scaling), synthetic node
features, default-initialised weights.  One step = forward over the batch, loss gather +
SmoothL1, backward through every kernel, (N>1: one NCCL all-reduce of the flat gradient
buffer), Adam step.

Prints ONE JSON line (rank 0).  `value` is whole-job graphs/s with inputs resident in HBM
(CSR of the batch cached: topology is reused across steps); `e2e` is the same
step driven from pinned HOST buffers through the public API: H2D of node features and
edge_index, CSR build, forward, loss, backward, optimiser, D2H of the loss -- every step.

`--impl reference` times the reference's own CPU arithmetic for this path (the dense
restatement of QNet.forward, oracle/s2v_ref.py, kind "port": the reference is pure Python
whose third-party imports are absent on the GPU box) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "s2v_fwd_bwd_graphs_per_sec_N975"
UNIT = "graphs/s"
TOPO = "NWB_AMS"
P, T, F = 64, 5, 7


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="graphs per GPU")
    ap.add_argument("--topo", default="NWB_AMS", choices=["NWB_AMS", "NWB_ROT", "NWB_UTR", "Manhattan5", "MemTask"],
                    help="topology of the synthetic batch (default: the N=975 graph the metric is quoted on; "
                         "NWB_ROT = BASELINE configs[4] scale sweep)")
    ap.add_argument("--qnet", default="s2v", choices=["s2v", "gat2"],
                    help="s2v = the headline path; gat2 = the GATv2 branch (SURVEY.md 8f-1: QNet_GATv2 train step)")
    ap.add_argument("--workload", default="dqn", choices=["dqn", "ppo"],
                    help="dqn = the headline GNN-DQN train step; ppo = one PPO epoch over --batch rollouts of --ppo-steps "
                         "steps per GPU (BASELINE configs[3]: PPO-GNN-LSTM on NWB_AMS, lstm EMB 64, critic q)")
    ap.add_argument("--ppo-steps", type=int, default=50)
    ap.add_argument("--ppo-lstm", default="EMB", choices=["EMB", "None"],
                    help="lstm_type of the PPO model (EMB = configs[3]; None isolates the GNN + heads from the cuDNN LSTM)")
    ap.add_argument("--ppo-batched", action="store_true",
                    help="ppo workload through PPO_GNN_Single_LSTM.evaluate_rollouts (all rollouts of the rank in one GNN "
                         "pass, s' embedded once) instead of the reference's per-rollout v(s'), v(s), pi(s) calls")
    ap.add_argument("--ref-kind", default="dense", choices=["dense", "sparse"],
                    help="--impl reference: dense = the reference's own arithmetic (dense W, (B,N,N,p) theta4 tensor; B = 2 "
                         "graphs per step is what fits), sparse = the index_add restatement of the same math (B = 64 per step)")
    ap.add_argument("--ppo-graph", action="store_true",
                    help="ppo workload: capture one whole epoch (the per-rollout v(s'), v(s), pi(s) calls, loss, backward, Adam) "
                         "in a CUDA graph after warm-up and time its replays -- same kernels, no per-launch host cost")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


TOPO_SHAPES = {"NWB_AMS": (975, 2850), "NWB_ROT": (2602, 7266), "NWB_UTR": (1182, 3204), "Manhattan5": (25, 105),
               "MemTask": (8, 22)}


def workload_config(batch, n_gpus):
    n, e = TOPO_SHAPES[TOPO]
    return {"workload": "s2v GNN-DQN train step on %s (N=%d, E=%d, F=7, p=64, T=5, norm_agg)" % (TOPO, n, e),
            "graphs_per_gpu": batch, "global_batch": batch * n_gpus, "nodes_per_graph": n,
            "parallelism": "dp%d" % n_gpus,
            "l2": "inputs exceed L2: per-GPU node features %.0f MB, each embedding plane %.2f GB"
                  % (batch * n * F * 4 / 1e6, batch * n * P * 4 / 1e9)}


def reference_config(sample_b, kind):
    """Config of the --impl reference line: the same workload, with the sample that is REALLY timed per step
    (the dense reference arithmetic materialises a (B,N,N,p) tensor: 243 MB per AMS graph, so B = 2)."""
    n, e = TOPO_SHAPES[TOPO]
    return {"workload": "s2v GNN-DQN train step on %s (N=%d, E=%d, F=7, p=64, T=5, norm_agg)" % (TOPO, n, e),
            "graphs_per_step": sample_b, "global_batch": sample_b, "nodes_per_graph": n, "parallelism": "host cores, 1 process",
            "ref_kind": kind,
            "note": "bounded sample of the workload: %d graphs per step (the B200 arm runs %s graphs per GPU per step); "
                    "throughput in graphs/s is the comparable quantity" % (sample_b, "--batch")}


# --------------------------------------------------------------------------------------
# CPU reference arithmetic (oracle; timed as the baseline only)
# --------------------------------------------------------------------------------------
def cpu_reference_step_fn(kind, batch, device="cpu"):
    """Returns (step_fn, graphs_per_step, description). kind: 'dense' (reference arithmetic) or 'sparse'.
    device 'cuda': the same restatement on torch's CUDA ops (what the reference does when it finds a GPU,
    comb_opt.py:44) -- the caller runs step() inside `with torch.device(device)`."""
    import torch
    from oracle import s2v_ref as R
    from simmodel_b200 import graphs
    topo = graphs.load_topology(TOPO)
    gen = torch.Generator().manual_seed(0)
    N = topo.num_nodes
    xv = graphs.synth_nfm(topo, batch, gen).to(device)
    Pm = {k: v.to(device).requires_grad_(True) for k, v in R.make_qnet_params(F, P, seed=0).items()}
    sizes = [N] * batch
    bt = R.batch_from_sizes(sizes, [topo.edge_index] * batch)
    bt.ptr, bt.batch, bt.edge_index = bt.ptr.to(device), bt.batch.to(device), bt.edge_index.to(device)
    actions = torch.randint(0, N, (batch,), generator=gen).tolist()
    targets = torch.randn(batch, generator=gen).tolist()
    if kind == "dense":
        W = torch.zeros(N, N)
        W[topo.edge_index[0], topo.edge_index[1]] = 1.0
        Ws = W.unsqueeze(0).repeat(batch, 1, 1).to(device)

        def step():
            for v in Pm.values():
                v.grad = None
            q = R.qnet_forward_dense(Pm, xv, Ws, bt.ptr, bt.batch, T, True)
            R.dqn_loss(q.reshape(-1), actions, bt.ptr, targets).backward()
        desc = "dense reference arithmetic (comb_opt.py:217-275 restated), fwd+bwd, B=%d AMS graphs per step" % batch
    else:
        x = xv.reshape(-1, F)

        def step():
            for v in Pm.values():
                v.grad = None
            q = R.qnet_forward_sparse(Pm, x, bt.edge_index, bt.ptr, N, T, True)
            R.dqn_loss(q, actions, bt.ptr, targets).backward()
        desc = "sparse CPU restatement (index_add), fwd+bwd, B=%d AMS graphs per step" % batch
    return step, batch, desc


def time_cpu(step, min_steps, budget_s, warmup=1):
    for _ in range(warmup):
        step()
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < min_steps or (time.perf_counter() < t_end and len(times) < 50):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if len(times) >= min_steps and time.perf_counter() >= t_end:
            break
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_b = 2 if args.ref_kind == "dense" else 64
    step, gps, desc = cpu_reference_step_fn(args.ref_kind, sample_b)
    for _ in range(max(1, min(args.warmup, 3))):
        step()
    times = []
    for _ in range(max(1, args.steps)):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if sum(times) > 150:
            break
    ms = 1e3 * sum(times) / len(times)
    val = gps / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(times), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "node_updates_per_sec": val * TOPO_SHAPES[TOPO][0] * T,
            "config": reference_config(sample_b, args.ref_kind),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": desc + "; torch %s CPU, %d threads" % (torch.__version__, cores)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def count_kernel_launches(fn):
    """Kernels launched by one call of fn, COUNTED from a CUPTI trace (torch.profiler): own = kernels of
    libs2v_b200.so (every one is named k_*), library = everything else (torch elementwise / Adam / NCCL)."""
    import re
    import torch
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        own, other, names = 0, 0, {}
        for ev in prof.events():
            if getattr(ev, "device_type", None) is None or "cuda" not in str(ev.device_type).lower():
                continue
            nm = ev.name
            if nm.startswith("Memcpy") or nm.startswith("Memset"):
                continue
            if re.search(r"(^|::|\s)k_[a-z0-9_]+", nm):
                own += 1
                short = re.search(r"k_[a-z0-9_]+", nm).group(0)
                names[short] = names.get(short, 0) + 1
            else:
                other += 1
        if own + other > 0:
            return {"own": own, "library": other, "by_kernel": names, "source": "CUPTI trace of one step (torch.profiler)"}
    except Exception as exc:          # no CUPTI on this box: fall back to the count kept by the bindings
        err = str(exc)[:120]
    else:
        err = "empty trace"
    from simmodel_b200 import _lib
    c0 = _lib.launch_count
    fn()
    return {"own": _lib.launch_count - c0, "library": None, "by_kernel": None,
            "source": "count kept by the ctypes bindings (CUPTI unavailable: %s)" % err}


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [s.strip() for s in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import simmodel_b200 as sb
    from simmodel_b200 import _lib
    from simmodel_b200.dist import allreduce_gradients

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a CUDA device; there is no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    topo = sb.graphs.load_topology(TOPO)
    N, B = topo.num_nodes, args.batch
    gen = torch.Generator().manual_seed(1000 + rank)
    x_host = sb.graphs.synth_nfm(topo, B, gen).reshape(-1, F).contiguous().pin_memory()
    ei_host = sb.graphs.batch_edge_index(topo, B).contiguous().pin_memory()
    ptr_host = (torch.arange(B + 1, dtype=torch.int64) * N)
    actions_host = torch.randint(0, N, (B,), generator=gen).pin_memory()
    targets_host = torch.randn(B, generator=gen).pin_memory()

    torch.manual_seed(0)
    net = sb.QNet({"emb_dim": P, "emb_iter_T": T, "norm_agg": True, "node_dim": F}, verbose=False).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    loss_fn = torch.nn.SmoothL1Loss()
    params = list(net.parameters())

    # resident inputs
    x = x_host.to(dev)
    ptr = ptr_host.to(dev)
    ei = ei_host.to(dev)
    sel = actions_host.to(dev) + ptr[:-1]
    targets = targets_host.to(dev)
    torch.cuda.synchronize()
    layout = sb.GraphLayout.from_batch(ei, ptr, N, validate=False)          # cold call (module load, first launches)
    csr_times = []
    for _ in range(7):                                                       # warm builds: what a step really pays
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        layout = sb.GraphLayout.from_batch(ei, ptr, N, validate=False)
        t1.record()
        torch.cuda.synchronize()
        csr_times.append(t0.elapsed_time(t1))
    csr_ms = statistics.median(csr_times)
    del ei

    def step():
        opt.zero_grad(set_to_none=True)
        q = net.forward_graph(x, layout)
        loss = loss_fn(q[sel], targets)
        loss.backward()
        if world > 1:
            allreduce_gradients(params, B, B * world)
        opt.step()
        return loss

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms.item())

    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup):
        step()
    launch_info = count_kernel_launches(step)
    launches_per_step = launch_info["own"]
    if rank == 0:
        sampler.start()
    total_ms = timed(step, args.steps, 0)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = B * world / (ms_per_step / 1e3)

    # ---- end to end from pinned host buffers through the public API --------------------
    e2e = None
    if not args.no_e2e:
        from simmodel_b200.pipeline import HostBatchPrefetcher
        pf = HostBatchPrefetcher(dev, depth=2)
        # The batch as a replay buffer holds it: node features per graph + a table of the DISTINCT topologies (here one:
        # every graph of the batch is the NWB_AMS world, as in the reference's single-map DQN runs) with int32 local ids.
        topo_host = topo.edge_index.to(torch.int32).contiguous().pin_memory()
        topo_of_graph = torch.zeros(B, dtype=torch.int32)
        host_batch = {"x": x_host, "topology": topo_host, "actions": actions_host, "targets": targets_host}
        loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
        state = {"ticket": None, "n": 0, "ev": [None, None], "last": 0.0}

        def step_e2e():
            # batch k was submitted during step k-1 (or now, for the first call); batch k+1 is submitted now and
            # copies on the side stream while batch k runs.  Every step's H2D and D2H are inside the timed region.
            cur = state["ticket"] if state["ticket"] is not None else pf.submit(host_batch)
            state["ticket"] = pf.submit(host_batch)
            b = pf.acquire(cur)
            lay = sb.GraphLayout.from_topologies([(b["topology"], N)], topo_of_graph, N, validate=False)
            opt.zero_grad(set_to_none=True)
            q = net.forward_graph(b["x"], lay)
            loss = loss_fn(q[b["actions"] + ptr[:-1]], b["targets"])
            loss.backward()
            pf.release(cur)
            if world > 1:
                allreduce_gradients(params, B, B * world)
            opt.step()
            k = state["n"] & 1
            loss_host[k:k + 1].copy_(loss.detach().reshape(1), non_blocking=True)     # D2H read of the step's result
            ev = torch.cuda.Event()
            ev.record()
            state["ev"][k] = ev
            prev = state["ev"][k ^ 1]
            if prev is not None:                # the host consumes the previous step's loss: runs one step ahead at most
                prev.synchronize()
                state["last"] = float(loss_host[k ^ 1])
            state["n"] += 1
            return state["last"]
        e2e_steps = max(3, min(args.steps, 10))
        e2e_ms = timed(step_e2e, e2e_steps, 2) / e2e_steps
        h2d = x_host.numel() * 4 + topo_host.numel() * 4 + actions_host.numel() * 8 + targets_host.numel() * 4
        e2e = {"value": B * world / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
               "h2d_detail": {"x_f32": x_host.numel() * 4, "topology_table_int32": topo_host.numel() * 4,
                              "actions_int64": actions_host.numel() * 8, "targets_f32": targets_host.numel() * 4},
               "includes": "H2D(x, topology table (1 distinct topology, int32 local ids), actions, targets; double-buffered on a "
                           "copy stream, simmodel_b200.pipeline.HostBatchPrefetcher) + CSR build (GraphLayout.from_topologies) "
                           "+ fwd + loss + bwd + Adam + D2H(loss); r01 shipped the batch-wide int64 edge_index (187 MB of 299 MB)"}

    # ---- roofline of the dominant kernels, timed alone with CUDA events ------------------
    roofline, kernels = None, None
    if rank == 0:
        roofline, kernels = measure_roofline(sb, _lib, net, layout, x, dev, B, N)

    cpu_baseline, small_batch = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = measure_cpu_baseline()
        try:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                small_batch = measure_small_batch(sb, dev)
        except Exception as exc:      # informational only
            small_batch = {"error": str(exc)[:200]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "arithmetic": "fp32 in, fp32 out; tensor-core contractions on two fp16 pieces per operand with exact power-of-two "
                              "tile scales, fp32 accumulation: <= 2.5e-7 of scale against fp64 per kernel (bf16x3: 2.1e-7, "
                              "CUDA-core fp32: 3.9e-7), DESIGN.md 4f",
                "node_updates_per_sec": value * N * T,
                "config": workload_config(B, world), "csr_build_ms": csr_ms,
                "csr_build_note": "median of 7 warm s2v_csr_build calls over the batch's int64 edge_index (cold first call excluded)",
                "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
                "gpu_launches_detail": launch_info,
                "clocks": clocks, "e2e": e2e, "roofline": roofline, "kernels": kernels,
                "cpu_baseline": cpu_baseline, "small_batch_latency": small_batch}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_roofline(sb, _lib, net, layout, x, dev, B, N):
    """Dominant kernels = one propagation round forward / backward.  Algorithmic bytes per
    node-update (fused-epilogue variant): 4*(p*(dbar+1) + dbar + 1) for the gather
    (neighbour rows + output row + column ids + rowptr) + 4*p for the one extra row the fused
    epilogue reads (base forward, mu^{t-1} backward)."""
    import ctypes as C
    import torch
    from simmodel_b200.layout import _p, _stream
    L = _lib.lib()
    n = layout.num_nodes
    dbar = layout.num_edges / max(1, (layout.period or n))
    bytes_per_node = 4 * (P * (dbar + 1) + dbar + 1) + 4 * P
    alg_bytes = bytes_per_node * n
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    with torch.no_grad():
        mu = net.embed_graph(x, layout)
    base = torch.randn(n, P, device=dev)
    out = torch.empty(n, P, device=dev)
    du = torch.randn(n, P, device=dev)
    ws = torch.empty(L.s2v_embed_backward_workspace_bytes(F, P), dtype=torch.uint8, device=dev)
    g = layout.c_struct()
    w2 = net.theta2.weight.detach()

    flag = _lib.path_flag()
    tensor_f = flag in (2, 3, 4) or (flag == 0 and n >= 4096)   # same rules as csrc/s2v_embed.cu:use_tensor_path
    # kernel names as csrc/s2v_embed_tc.cu selects them (defaults: fp16x2 rounds; the env switches give the older kernels)
    fwd_env = os.environ.get("S2V_B200_FWD_ROUND", "h2")
    bwd_env = os.environ.get("S2V_B200_BWD_ROUND", "tc16")
    bwd_kern = os.environ.get("S2V_B200_BWD16_KERNEL", "h")
    if not tensor_f:
        sfx_f = sfx_b = " (FFMA)"
    else:
        sfx_f = ("_tc (tcgen05 3xTF32)" if os.environ.get("S2V_B200_FWD_TF32") else
                 "_h2 (tcgen05 fp16x2, per-warp power-of-two scales)" if fwd_env.startswith("h2") else "_tc16 (tcgen05 bf16x3)")
        if flag == 2:
            sfx_b = "_tc256 (tcgen05 3xTF32)"
        elif flag == 3 or (flag == 0 and not bwd_env.startswith("tc16")):
            sfx_b = "_ws (tcgen05 bf16x3, warp-specialised)"
        else:
            sfx_b = {"h": "_h2 (tcgen05 fp16x2, per-tile scales)", "p": "_tc16p (tcgen05 bf16x3, 8-warp pipelined)",
                     "s": "_tc16 (tcgen05 bf16x3, 4-warp)"}.get(bwd_kern[:1], "_h2 (tcgen05 fp16x2, per-tile scales)")

    def fwd():
        _lib.check(L.s2v_embed_round_forward(C.byref(g), P, flag, _p(w2), _p(base), _p(mu), _p(out), _stream()), "round fwd")

    def bwd():
        _lib.check(L.s2v_embed_round_backward(C.byref(g), F, P, flag, _p(w2), _p(du), _p(mu), _p(out), _p(ws), ws.numel(),
                                              _stream()), "round bwd")
    res = []
    for name, fn in (("k_embed_round" + sfx_f + ", forward round", fwd), ("k_embed_round_bwd" + sfx_b + ", backward round", bwd)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a, b in evs:
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        ms = statistics.mean(a.elapsed_time(b) for a, b in evs)
        ach = alg_bytes / (ms / 1e3) / 1e9
        res.append({"kernel": name, "ms": ms, "algorithmic_bytes": alg_bytes, "achieved_GBps": ach, "frac": ach / peak,
                    "launches_per_step": T - 1})
    traffic_tab = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic_tab = json.load(f)
    except Exception:
        pass
    for r in res:
        key = r["kernel"].split(" ")[0].split(",")[0]
        ent = traffic_tab.get(key)
        r["traffic"] = ent["bytes_per_node_update"] * n if ent else None
    dom = max(res, key=lambda r: r["ms"])
    for r in res:
        r["frac_nominal"] = r["achieved_GBps"] / 8000.0
        r["dram_frac"] = (r["traffic"] / (r["ms"] / 1e3) / 1e9 / peak) if r["traffic"] else None
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_GBps"], "peak": peak, "unit": "GB/s",
                "frac": dom["frac"], "frac_nominal": dom["frac_nominal"], "dram_frac": dom["dram_frac"],
                "frac_note": "frac = algorithmic bytes / time / measured copy peak; frac_nominal = the same against the 8 TB/s "
                             "nominal HBM3e figure; dram_frac = measured DRAM traffic (ncu) / time / measured copy peak",
                "traffic": dom["traffic"],
                "traffic_source": "ncu --set full dram__bytes_read+write per node-update (profiles/traffic.json) x rows of this launch",
                "peak_source": peak_src,
                "algorithmic_bytes_per_node_update": bytes_per_node, "node_updates_per_launch": n,
                "avg_launch_ms": dom["ms"]}
    return roofline, res


# --------------------------------------------------------------------------------------
# GATv2 branch (--qnet gat2): QNet_GATv2 train step on the same synthetic batch
# --------------------------------------------------------------------------------------
def run_gat(args):
    """One step = QNet_GATv2 forward (5 GATv2Conv layers, heads 2, hidden 64 + Q head), loss gather + SmoothL1,
    backward, Adam.  Single GPU.  The edge-attention kernels are timed alone for the roofline; the lin_l / lin_r
    GEMMs are cuBLAS (library) and reported as their share of the step."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    import simmodel_b200 as sb
    from simmodel_b200 import _lib
    from simmodel_b200.dist import allreduce_gradients
    from simmodel_b200.layout import _p, _stream
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a CUDA device; there is no CPU fallback")
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    topo = sb.graphs.load_topology(TOPO)
    N, B = topo.num_nodes, args.batch
    gen = torch.Generator().manual_seed(1000 + rank)
    x_host = sb.graphs.synth_nfm(topo, B, gen).reshape(-1, F).contiguous().pin_memory()
    ei_host = sb.graphs.batch_edge_index(topo, B).contiguous().pin_memory()
    ptr = (torch.arange(B + 1, dtype=torch.int64) * N).to(dev)
    actions_host = torch.randint(0, N, (B,), generator=gen).pin_memory()
    targets_host = torch.randn(B, generator=gen).pin_memory()
    torch.manual_seed(0)
    net = sb.QNet_GATv2({"emb_dim": P, "emb_iter_T": T, "norm_agg": True, "node_dim": F}, verbose=False).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    params = list(net.parameters())
    loss_fn = torch.nn.SmoothL1Loss()
    x = x_host.to(dev)
    sel = actions_host.to(dev) + ptr[:-1]
    targets = targets_host.to(dev)
    layout = sb.gat.gat_layout_from_batch(ei_host.to(dev), ptr)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(net.forward_graph(x, layout)[sel], targets)
        loss.backward()
        if world > 1:
            allreduce_gradients(params, B, B * world)
        opt.step()
        return loss

    def step_e2e():
        xb = x_host.to(dev, non_blocking=True)
        lay = sb.gat.gat_layout_from_batch(ei_host.to(dev, non_blocking=True), ptr)
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(net.forward_graph(xb, lay)[actions_host.to(dev, non_blocking=True) + ptr[:-1]],
                       targets_host.to(dev, non_blocking=True))
        loss.backward()
        if world > 1:
            allreduce_gradients(params, B, B * world)
        opt.step()
        return float(loss.detach())             # D2H read of the step's result

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms_t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms_t.item())
    for _ in range(args.warmup):
        step()
    launch_info = count_kernel_launches(step)
    launches_per_step = launch_info["own"]
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    ms = timed(step, args.steps, 0)
    clocks = sampler.stop()
    e2e_ms = None if args.no_e2e else timed(step_e2e, max(3, min(args.steps, 10)), 2)
    # ---- the edge-attention kernels alone: the model's layer shape (H=2, C=32) and the PPO gat2 shapes ----
    n = layout.num_nodes
    dbar = layout.num_edges / n
    g = layout.c_struct()
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    kernels = []
    shapes = [(2, P // 2, 5)] + ([(4, 32, 0), (4, 16, 0)] if n * 256 * 4 * 8 < 60e9 else [])    # (heads, channels, launches per step)
    for H, Cc, per_step in shapes:
        HC = H * Cc
        xlr = torch.randn(n, 2 * HC, device=dev)
        att, bias = torch.randn(HC, device=dev) * 0.1, torch.randn(HC, device=dev) * 0.1
        lin_bias = torch.randn(2 * HC, device=dev) * 0.1       # as in the step: bias-free GEMM output + biases added at load
        out = torch.empty(n, HC, device=dev)
        ef = int(L.s2v_gat_edge_floats(C.byref(g), H))
        alpha = torch.empty(ef, device=dev)
        dout, dxlr, gab = torch.randn(n, HC, device=dev), torch.empty(n, 2 * HC, device=dev), torch.empty(4, HC, device=dev)
        ws = torch.empty(L.s2v_gat_attn_backward_workspace_bytes(H, Cc, ef), dtype=torch.uint8, device=dev)
        xr_p, dxr_p = C.c_void_p(xlr.data_ptr() + 4 * HC), C.c_void_p(dxlr.data_ptr() + 4 * HC)

        def k_fwd():
            _lib.check(L.s2v_gat_attn_forward(C.byref(g), H, Cc, 2 * HC, _p(xlr), xr_p, _p(att), _p(bias), _p(lin_bias), 1,
                                              _p(out), None, _p(alpha), _stream()), "gat fwd")

        def k_bwd():
            _lib.check(L.s2v_gat_attn_backward(C.byref(g), H, Cc, 2 * HC, _p(xlr), xr_p, _p(att), _p(lin_bias), None, _p(alpha), _p(layout.row_to_col), _p(dout),
                                               _p(dxlr), dxr_p, _p(gab[0]), _p(gab[1]), _p(gab[2]), _p(ws), ws.numel(), _stream()),
                       "gat bwd")
        # forward: (dbar+1) x_l rows + x_r row read, out row written, ids + rowptr, statistics written
        fwd_bytes = 4 * (HC * (dbar + 3) + dbar + 1 + H * (dbar + 1))
        # backward (both kernels): target pass reads (dbar+1) x_l rows + x_r + dout + statistics, writes dx_r + D;
        # source pass reads x_l + (dbar+1) x (x_r, dout, statistics) rows, writes dx_l; ids + rowptr twice
        bwd_bytes = 4 * (HC * (dbar + 1 + 3) + 2 * H * (dbar + 1) + HC * (1 + 2 * (dbar + 1) + 1) + 2 * H * (dbar + 1) + 3 * (dbar + 1))
        for name, fn, bpn in (("k_gat_fwd4 (H=%d, C=%d)" % (H, Cc), k_fwd, fwd_bytes),
                              ("k_gat_bwd_target4 + k_gat_bwd_source4 (H=%d, C=%d)" % (H, Cc), k_bwd, bwd_bytes)):
            k_ms = timed(fn, 10, 3)
            ach = bpn * n / (k_ms / 1e3) / 1e9
            kernels.append({"kernel": name, "ms": k_ms, "algorithmic_bytes": bpn * n, "achieved_GBps": ach, "frac": ach / peak,
                            "launches_per_step": per_step})
        del xlr, out, alpha, dout, dxlr
    try:                                            # DRAM bytes per row from the ncu --set full capture (H*C = 64 kernels)
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tt = json.load(f)
        kernels[0]["traffic"] = tt["k_gat_fwd4"]["bytes_per_node_update"] * n
        kernels[1]["traffic"] = tt["k_gat_bwd_target4+k_gat_bwd_source4"]["bytes_per_node_update"] * n
    except Exception:
        pass
    dom = max(kernels[:2], key=lambda r: r["ms"])
    edge_ms = sum(k["ms"] * k["launches_per_step"] for k in kernels)
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_GBps"], "peak": peak, "unit": "GB/s",
                "frac": dom["frac"], "traffic": dom.get("traffic"),
                "traffic_source": "ncu --set full dram__bytes_read+write per row (profiles/traffic.json) x rows of this launch",
                "algorithmic_bytes_per_row": dom["algorithmic_bytes"] / n,
                "rows_per_launch": n, "avg_launch_ms": dom["ms"],
                "edge_kernels_share_of_step": edge_ms / ms,
                "note": "the rest of the step: lin_l/lin_r GEMMs (tcgen05 bf16x3 kernels for the 64-wide layers, cuBLAS for the first), Q head, Adam"}
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import gat_ref as G
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        b0 = 8
        Pm = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.state_dict().items()}
        xs = x_host[: b0 * N].clone()
        eis = sb.graphs.batch_edge_index(topo, b0)
        ptr0 = torch.arange(b0 + 1, dtype=torch.int64) * N
        sel0 = actions_host[:b0] + ptr0[:-1]

        def cpu_step():
            for v in Pm.values():
                v.grad = None
            q = G.qnet_gat_forward(Pm, xs, eis, ptr0, 5, True)
            loss_fn(q[sel0], targets_host[:b0]).backward()
        times = time_cpu(cpu_step, 3, 10.0)
        cpu_baseline = {"value": b0 / statistics.mean(times), "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "oracle/gat_ref.py (PyG GATv2Conv restated, torch CPU), fwd+bwd, B=%d graphs per step, %d steps"
                                  % (b0, len(times))}
    n_t, e_t = TOPO_SHAPES[TOPO]
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {"metric": "gat2_fwd_bwd_graphs_per_sec_N%d" % n_t, "value": B * world / (ms / 1e3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "QNet_GATv2 GNN-DQN train step on %s (N=%d, E=%d, F=7, emb 64, 5 GATv2Conv layers, heads 2)"
                                   % (TOPO, n_t, e_t), "graphs_per_gpu": B, "global_batch": B * world, "nodes_per_graph": n_t,
                       "parallelism": "dp%d" % world, "l2": "inputs exceed L2: every activation plane %.2f GB" % (B * n_t * P * 4 / 1e9)},
            "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
            "gpu_launches_detail": launch_info, "clocks": clocks,
            "e2e": None if e2e_ms is None else {
                "value": B * world / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": x_host.numel() * 4 + ei_host.numel() * 8 + actions_host.numel() * 8 + targets_host.numel() * 4,
                "d2h_bytes_per_step": 4},
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# PPO workload (--workload ppo): one PPO epoch of the reference's train_net over R rollouts per GPU
# --------------------------------------------------------------------------------------
def run_ppo(args):
    """BASELINE configs[3]: PPO-GNN-LSTM (lstm_type EMB, lstm_hdim 64, critic q) on NWB_AMS, R = --batch rollouts of
    S = --ppo-steps steps per GPU (whole rollouts per rank).  One step = one epoch of
    models_basic_lstm.py:595-652: per rollout v(s'), v(s), pi(s) through the public pi()/v() (three embedding passes,
    the reference's call pattern), clipped-surrogate + SmoothL1 loss, mean over all steps of all rollouts, backward,
    (all-reduce), Adam.  --qnet selects the s2v or the GATv2 embedding.  Metric unit: rollout steps (graphs) per second."""
    import types
    import torch
    import torch.distributed as dist
    import torch.nn.functional as Fn
    import simmodel_b200 as sb
    from simmodel_b200 import _lib
    from simmodel_b200.dist import allreduce_gradients
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a CUDA device; there is no CPU fallback")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    topo = sb.graphs.load_topology(TOPO)
    N, R, S = topo.num_nodes, args.batch, args.ppo_steps
    config = {"lstm_type": args.ppo_lstm, "qnet": args.qnet, "critic": "q", "emb_dim": P, "emb_iterT": T, "gat_concat": True,
              "gat_heads": 4, "gat_share_weights": False}
    hp = types.SimpleNamespace(node_dim=F, emb_dim=P, parallel_rollouts=R, batch_size=1, learning_rate=1e-4,
                               discount=0.99, gae_lambda=0.95, ppo_clip=0.2)
    torch.manual_seed(0)
    model = sb.PPO_GNN_Single_LSTM(config, hp, {"base_checkpoint_path": None}, device=dev)
    params = list(model.parameters())
    gen = torch.Generator().manual_seed(2000 + rank)
    ei = topo.edge_index.to(dev)
    host = []
    for _ in range(R):                                  # a rollout: S consecutive observations of one episode
        nfm = sb.graphs.synth_nfm(topo, S + 1, gen)
        reach = torch.zeros(S + 1, N, dtype=torch.bool)
        cur = nfm[:, :, 1].argmax(dim=1)
        for t in range(S + 1):
            reach[t, topo.neighbors(int(cur[t])) or [int(cur[t])]] = True
        a = torch.stack([reach[t].nonzero()[0] for t in range(S)])                   # a reachable action per step
        host.append(dict(nfm=nfm.pin_memory(), reach=reach.pin_memory(), a=a.pin_memory(),
                         r=torch.randn(S, 1, generator=gen).pin_memory(), prob_a=(torch.rand(S, 1, generator=gen) * 0.5 + 0.25).pin_memory(),
                         done=torch.ones(S, 1).pin_memory()))
    h2d = sum(v.numel() * v.element_size() for v in host[0].values()) * R

    def to_dev(b):
        return {k: v.to(dev, non_blocking=True) for k, v in b.items()}
    resident = [to_dev(b) for b in host]

    def epoch(batches):
        losses = []
        for b in batches:
            hidden = (torch.zeros(1, N, P, device=dev), torch.zeros(1, N, P, device=dev))
            nfm, nfm_p, reach, reach_p = b["nfm"][:-1], b["nfm"][1:], b["reach"][:-1], b["reach"][1:]
            v_prime = model.v(nfm_p, ei, reach_p, hidden)[0]
            td = b["r"] + hp.discount * v_prime * b["done"]
            v_s = model.v(nfm, ei, reach, hidden)[0]
            delta = (td - v_s).detach()
            k = torch.arange(S, device=dev, dtype=torch.float32)                    # GAE as one discounted suffix sum
            w = (hp.discount * hp.gae_lambda) ** k
            adv = torch.flip(torch.cumsum(torch.flip(delta[:, 0] * w, [0]), 0), [0]) / w
            pi, _ = model.pi(nfm, ei, reach, hidden)
            ratio = torch.exp(torch.log(pi.gather(1, b["a"])) - torch.log(b["prob_a"]))
            surr1, surr2 = ratio * adv[:, None], torch.clamp(ratio, 1 - hp.ppo_clip, 1 + hp.ppo_clip) * adv[:, None]
            losses.append((-torch.min(surr1, surr2) + Fn.smooth_l1_loss(v_s, td.detach())).squeeze(-1))
        loss = torch.cat(losses).mean()
        model.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        if world > 1:
            allreduce_gradients(params, R * S, R * S * world)
        model.optimizer.step()
        return loss

    def stack(batches):
        return {k: torch.stack([b[k] for b in batches]) for k in batches[0]}

    def epoch_batched(bt):
        hidden = (torch.zeros(1, R * N, P, device=dev), torch.zeros(1, R * N, P, device=dev))
        nfm, reach = bt["nfm"], bt["reach"]
        pi, v_s, v_prime = model.evaluate_rollouts(nfm[:, :-1], nfm[:, 1:], ei, reach[:, :-1], reach[:, 1:], hidden, hidden,
                                                   prime_is_shift=True)
        td = bt["r"] + hp.discount * v_prime * bt["done"]
        delta = (td - v_s).detach()
        w = ((hp.discount * hp.gae_lambda) ** torch.arange(S, device=dev, dtype=torch.float32))[None, :, None]
        adv = torch.flip(torch.cumsum(torch.flip(delta * w, [1]), 1), [1]) / w
        ratio = torch.exp(torch.log(pi.gather(2, bt["a"])) - torch.log(bt["prob_a"]))
        surr1, surr2 = ratio * adv, torch.clamp(ratio, 1 - hp.ppo_clip, 1 + hp.ppo_clip) * adv
        per_rollout_l2 = Fn.smooth_l1_loss(v_s, td.detach(), reduction="none").mean(dim=(1, 2), keepdim=True)
        loss = (-torch.min(surr1, surr2) + per_rollout_l2).mean()
        model.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        if world > 1:
            allreduce_gradients(params, R * S, R * S * world)
        model.optimizer.step()
        return loss
    resident_stacked = stack(resident) if args.ppo_batched else None

    def step():
        return epoch_batched(resident_stacked) if args.ppo_batched else epoch(resident)

    if args.ppo_graph:                      # whole-epoch capture: shapes, layouts and buffers are static across epochs
        if world > 1:
            raise SystemExit("--ppo-graph is a single-GPU option")
        model.optimizer = torch.optim.Adam(model.parameters(), lr=hp.learning_rate, capturable=True)
        eager_step = step
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                eager_step()
        torch.cuda.current_stream().wait_stream(side)
        cgraph = torch.cuda.CUDAGraph()
        model.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(cgraph):
            graph_loss = eager_step()

        def step():
            cgraph.replay()
            return graph_loss

    def step_e2e():
        if args.ppo_batched:
            return float(epoch_batched(stack([to_dev(b) for b in host])).detach())
        return float(epoch([to_dev(b) for b in host]).detach())

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms_t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms_t.item())
    for _ in range(args.warmup):
        step()
    launch_info = count_kernel_launches(eager_step if args.ppo_graph else step)
    launches = launch_info["own"]
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
    ms = timed(step, args.steps, 0)
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms = None if (args.no_e2e or args.ppo_graph) else timed(step_e2e, max(3, min(args.steps, 10)), 1)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.qnet == "s2v":
        # the reference's arithmetic (dense W, (S,N,N,p) theta4 tensor; oracle/s2v_ref.py restating models_basic_lstm.py:479-566)
        # on a bounded sample: v(s'), v(s), pi(s) forward + backward for the first steps of one rollout
        from oracle import s2v_ref as Rf
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        s_cpu = max(1, min(S, int(2.5e8 // (N * N * P))))       # keep the (S,N,N,p) tensor under ~1 GB
        Pm = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
        b0 = host[0]
        nfm_c, reach_c = b0["nfm"][:s_cpu + 1].clone(), b0["reach"][:s_cpu + 1].clone()
        ei_c = topo.edge_index
        lstm_c = torch.nn.LSTM(P, P) if args.ppo_lstm == "EMB" else None
        hid_c = (torch.zeros(1, N, P), torch.zeros(1, N, P)) if lstm_c is not None else None

        def cpu_step():
            for v in Pm.values():
                v.grad = None
            vp, _ = Rf.ppo_v_dense(Pm, nfm_c[1:], ei_c, reach_c[1:], T, "q", lstm_c, hid_c)
            vs, _ = Rf.ppo_v_dense(Pm, nfm_c[:-1], ei_c, reach_c[:-1], T, "q", lstm_c, hid_c)
            pi, _ = Rf.ppo_pi_dense(Pm, nfm_c[:-1], ei_c, reach_c[:-1], T, lstm_c, hid_c)
            (vp.sum() + vs.sum() + pi.gather(1, b0["a"][:s_cpu]).log().sum()).backward()
        times = time_cpu(cpu_step, 2, 10.0)
        cpu_baseline = {"value": s_cpu / statistics.mean(times), "unit": "rollout-steps/s", "cores": cores, "kind": "port",
                        "sample": "dense reference arithmetic (models_basic_lstm.py:479-566 restated): v(s'), v(s), pi(s) fwd+bwd "
                                  "for %d steps of one rollout on %s, %d repetitions; torch %s CPU" % (s_cpu, TOPO, len(times), torch.__version__)}
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    n_t, e_t = TOPO_SHAPES[TOPO]
    line = {"metric": "ppo_%s_epoch_rollout_steps_per_sec_N%d" % (args.qnet, n_t), "value": R * S * world / (ms / 1e3),
            "unit": "rollout-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "node_updates_per_sec": (R * (S + 1) if args.ppo_batched else 3 * R * S) * world * n_t * T / (ms / 1e3),
            "config": {"workload": "PPO-GNN-LSTM epoch (qnet %s, lstm %s 64, critic q) on %s (N=%d, E=%d): %s, loss, backward, Adam"
                                   % (args.qnet, args.ppo_lstm, TOPO, n_t, e_t,
                                      "evaluate_rollouts: one GNN pass over all rollouts (s' embedded once)" if args.ppo_batched
                                      else "per rollout v(s'), v(s), pi(s) through pi()/v() (the reference's call pattern)")
                                   + (" -- whole epoch replayed from a CUDA graph" if args.ppo_graph else ""),
                       "rollouts_per_gpu": R, "steps_per_rollout": S, "global_rollouts": R * world, "nodes_per_graph": n_t,
                       "parallelism": "dp%d (whole rollouts per rank)" % world,
                       "l2": "per rollout pass %d rows x 256 B = %.1f MB per plane; %d passes per epoch"
                             % (S * n_t, S * n_t * 256 / 1e6, 3 * R)},
            "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches, "gpu_launches_detail": launch_info,
            "clocks": clocks,
            "e2e": None if e2e_ms is None else {"value": R * S * world / (e2e_ms / 1e3), "unit": "rollout-steps/s",
                                                "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "cpu_baseline": cpu_baseline}
    print(json.dumps(line), flush=True)


def measure_small_batch(sb, dev):
    """The reference's own DQN call shapes (BASELINE configs[2]: batch 32 train step, B = 1 acting forward),
    eager and replayed from a CUDA graph (host wall clock per call, device synchronised)."""
    import torch
    topo = sb.graphs.load_topology(TOPO)
    N = topo.num_nodes
    torch.manual_seed(0)
    net = sb.QNet({"emb_dim": P, "emb_iter_T": T, "norm_agg": True, "node_dim": F}, verbose=False).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, capturable=True)
    loss_fn = torch.nn.SmoothL1Loss()

    def wall(fn, n=100):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3
    out = {}
    for B in (32, 1):
        gen = torch.Generator().manual_seed(B)
        x = sb.graphs.synth_nfm(topo, B, gen).reshape(-1, F).to(dev)
        lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), N, B)
        sel = (torch.randint(0, N, (B,), generator=gen) + torch.arange(B) * N).to(dev)
        tg = torch.randn(B, generator=gen).to(dev)

        def train_step():
            opt.zero_grad(set_to_none=False)
            loss = loss_fn(net.forward_graph(x, lay)[sel], tg)
            loss.backward()
            opt.step()
            return loss

        def act():
            with torch.no_grad():
                return net.forward_graph(x, lay)
        res = {"train_ms_eager": wall(train_step), "forward_ms_eager": wall(act)}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                train_step()
        torch.cuda.current_stream().wait_stream(side)
        g_train, g_act = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_train):
            train_step()
        with torch.cuda.graph(g_act):
            act()
        res["train_ms_graph"] = wall(g_train.replay)
        res["forward_ms_graph"] = wall(g_act.replay)
        out["B%d" % B] = res
    return out


def measure_cpu_baseline():
    """Reference CPU arithmetic on this box's host cores, bounded sample (~10-20 s)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dense_b = 2 if TOPO_SHAPES[TOPO][0] <= 1200 else 1          # the (B,N,N,p) tensor: 1.7 GB per ROT graph
    step, gps, desc = cpu_reference_step_fn("dense", dense_b)
    times = time_cpu(step, 3 if dense_b == 2 else 2, 8.0)
    dense = gps / statistics.mean(times)
    step2, gps2, desc2 = cpu_reference_step_fn("sparse", 64)
    times2 = time_cpu(step2, 3, 6.0)
    sparse = gps2 / statistics.mean(times2)
    out = {"value": dense, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": "%s, %d steps; torch %s CPU" % (desc, len(times), torch.__version__),
           "sparse_port": {"value": sparse, "unit": UNIT, "sample": "%s, %d steps" % (desc2, len(times2))}}
    out["torch_on_this_gpu"] = measure_torch_gpu_port()
    return out


def measure_torch_gpu_port():
    """The same restatements on torch's CUDA ops on this GPU (the reference picks device='cuda' when it finds one,
    comb_opt.py:44): what a user of the unmodified reference arithmetic would get from the same B200.  Bounded
    samples; reported next to the CPU figures, not instead of them."""
    import torch
    res = {}
    for kind, batch in (("dense", 8), ("sparse", 256)):
        try:
            step, gps, desc = cpu_reference_step_fn(kind, batch, "cuda")      # inputs built on the host, moved over
            with torch.device("cuda"):                                        # tensors the restatement creates itself
                for _ in range(2):
                    step()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                n = 0
                while n < 3 or (time.perf_counter() - t0 < 3.0 and n < 50):
                    step()
                    n += 1
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / n
            res[kind] = {"value": gps / dt, "unit": UNIT, "sample": desc.replace("CPU ", "") + ", torch CUDA ops, %d steps" % n}
        except Exception as e:                                  # an OOM or a device mismatch here must not cost the bench line
            res[kind] = {"error": "%s: %s" % (type(e).__name__, str(e)[:120])}
        torch.cuda.empty_cache()
    return res


def main():
    global TOPO, METRIC
    args = parse()
    TOPO = args.topo
    if TOPO != "NWB_AMS":
        METRIC = "s2v_fwd_bwd_graphs_per_sec_N%d" % TOPO_SHAPES[TOPO][0]
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "ppo":
        run_ppo(args)
    elif args.qnet == "gat2":
        run_gat(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
