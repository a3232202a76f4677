import json, sys
for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("%s: %.2f ms/step  %.0f graphs/s  %.3g node-upd/s | e2e %.2f ms %.0f g/s | launches/step %s | clocks %s" % (
        f, d["ms_per_step"], d["value"], d.get("node_updates_per_sec", 0), e.get("ms_per_step", 0), e.get("value", 0),
        d.get("gpu_launches_per_step"), d.get("clocks")))
    for k in d.get("kernels") or []:
        print("    %-58s %.3f ms  %.0f GB/s alg  frac %.3f" % (k["kernel"], k["ms"], k["achieved_GBps"], k["frac"]))
    if d.get("cpu_baseline"):
        print("    cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"].get("sparse_port", {}).get("value"))
