"""Run / report the BASELINE.json configs that are not the headline line.

    python scripts/config_matrix.py run --gpus N [--out gpurun_out/config_matrix_nN.jsonl] [--only c0,c1,c3,c4]
        N = 1: configs[0] Manhattan5 B=32 DQN step, configs[1] PPO-GNN-LSTM (EMB) on MemTask, configs[3] PPO on AMS,
               configs[4] ROT sweep B = 512 / 2048 / 8192, each with its CPU baseline
        N > 1: configs[3] and configs[4] under torchrun (one rank per GPU, weak scaling: the per-GPU batch is fixed)
    python scripts/config_matrix.py report gpurun_out/config_matrix_n*.jsonl > profiles/r02_config_matrix.md

Every line is bench.py's own JSON line (same timing rules: CUDA events, max over ranks, clocks sampled under load)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bench_cmd(n, extra):
    base = [sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", str(n)] + extra
    if n == 1:
        return base
    return [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
            "--master-port", "29541"] + base[1:]


def configs(n):
    multi = ["--no-cpu-baseline"] if n > 1 else []
    out = []
    if n == 1:
        out.append(("c0 Manhattan5 B=32 DQN step", ["--topo", "Manhattan5", "--batch", "32", "--steps", "50", "--warmup", "10"]))
        out.append(("c1 PPO EMB MemTask R=4 S=50 (reference call pattern)", ["--workload", "ppo", "--topo", "MemTask", "--batch", "4",
                                                                              "--steps", "10", "--warmup", "3"]))
        out.append(("c1 PPO EMB MemTask R=4 S=50 (reference call pattern, CUDA-graph replay)",
                    ["--workload", "ppo", "--ppo-graph", "--topo", "MemTask", "--batch", "4", "--steps", "20", "--warmup", "3"]))
        out.append(("c1 PPO EMB MemTask R=64 S=50 (evaluate_rollouts)", ["--workload", "ppo", "--ppo-batched", "--topo", "MemTask",
                                                                         "--batch", "64", "--steps", "10", "--warmup", "3"]))
    out.append(("c3 PPO EMB AMS R=64 S=50 (evaluate_rollouts)", ["--workload", "ppo", "--ppo-batched", "--batch", "64", "--steps", "5",
                                                                 "--warmup", "3"] + multi))
    out.append(("c3 PPO (no LSTM) AMS R=64 S=50 (evaluate_rollouts): the GNN + heads alone",
                ["--workload", "ppo", "--ppo-batched", "--ppo-lstm", "None", "--batch", "64", "--steps", "5", "--warmup", "3",
                 "--no-cpu-baseline", "--no-e2e"]))
    for b in (512, 2048, 8192):
        out.append(("c4 ROT B=%d DQN step" % b, ["--topo", "NWB_ROT", "--batch", str(b), "--steps", "5", "--warmup", "3"] +
                    (multi if (n > 1 or b != 512) else []) + (["--no-cpu-baseline"] if (n == 1 and b != 512) else [])))
    return out


def run(n, out_path, only):
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "a") as f:
        for name, extra in configs(n):
            if only and not any(name.startswith(o) for o in only):
                continue
            print("[config_matrix] %s (N=%d)" % (name, n), flush=True)
            p = subprocess.run(bench_cmd(n, extra), capture_output=True, text=True, timeout=1500)
            line = None
            for ln in p.stdout.splitlines():
                if ln.startswith("{"):
                    line = ln
            if p.returncode != 0 or line is None:
                print("  FAILED rc=%d\n%s" % (p.returncode, p.stderr[-1500:]), flush=True)
                continue
            rec = json.loads(line)
            rec["_name"] = name
            f.write(json.dumps(rec) + "\n")
            f.flush()
            print("  %.4g %s  (%.2f ms/step)" % (rec["value"], rec["unit"], rec["ms_per_step"]), flush=True)


def report(paths):
    recs = []
    for p in paths:
        with open(p) as f:
            recs += [json.loads(ln) for ln in f if ln.strip()]
    names = []
    for r in recs:
        if r["_name"] not in names:
            names.append(r["_name"])
    print("# r02 config matrix: every BASELINE.json config, measured by `scripts/config_matrix.py` (bench.py JSON lines)\n")
    print("Weak scaling: the per-GPU batch is fixed, `value` is the whole-job aggregate, time = max over ranks (CUDA events).\n")
    print("| config | GPUs | value | unit | ms/step | e2e value | scaling vs 1 GPU | CPU baseline (cores) | SM MHz (max) / reasons |")
    print("|---|---|---|---|---|---|---|---|---|")
    for nm in names:
        rows = sorted([r for r in recs if r["_name"] == nm], key=lambda r: r["n_gpus"])
        one = next((r for r in rows if r["n_gpus"] == 1), None)
        for r in rows:
            eff = "%.3f" % (r["value"] / (one["value"] * r["n_gpus"])) if one else "-"
            e2e = "%.4g" % r["e2e"]["value"] if r.get("e2e") else "-"
            cb = r.get("cpu_baseline")
            cbs = "%.4g %s (%d)" % (cb["value"], cb["unit"], cb["cores"]) if cb else "-"
            ck = r.get("clocks") or {}
            print("| %s | %d | %.4g | %s | %.2f | %s | %s | %s | %s (%s) / %s |" % (
                nm, r["n_gpus"], r["value"], r["unit"], r["ms_per_step"], e2e, eff, cbs, ck.get("sm_mhz"), ck.get("sm_max_mhz"),
                ",".join(ck.get("reasons") or []) or "none"))
    print("\nFull JSON lines: `profiles/r02_config_matrix_n{1,2,4,8}.jsonl` (copies of the `gpurun_out/config_matrix_n*.jsonl` "
          "files this table was made from).")


if __name__ == "__main__":
    if sys.argv[1] == "run":
        n = int(sys.argv[sys.argv.index("--gpus") + 1])
        out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else os.path.join(ROOT, "gpurun_out", "config_matrix_n%d.jsonl" % n)
        only = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else None
        run(n, out, only)
    else:
        report(sys.argv[2:])
