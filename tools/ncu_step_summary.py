"""Summarise `ncu --set full -k regex:gemm_umma_kernel ... bench.py` (one step's GEMM launches):
per-launch time, DRAM bytes, tensor-pipe activity -> markdown table + profiles/traffic.json.
usage: python tools/ncu_step_summary.py report.ncu-rep [--write-traffic]"""
import csv, io, json, os, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = rows[0]
col = {n: i for i, n in enumerate(h)}
def f(r, name):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return float("nan")
units = rows[1]
out = []
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    rd *= scale.get(units[col["dram__bytes_read.sum"]], 1.0)
    wr *= scale.get(units[col["dram__bytes_write.sum"]], 1.0)
    t = f(r, "gpu__time_duration.sum")
    t *= {"us": 1.0, "ns": 1e-3, "ms": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0)
    out.append(dict(name=name, us=t, rd=rd, wr=wr,
                    tensor=f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                    regs=f(r, "launch__registers_per_thread"), grid=f(r, "launch__grid_size")))
print("| # | kernel | time us | dram read MB | dram write MB | tensor pipe active % | regs |")
print("|---|---|---|---|---|---|---|")
for i, o in enumerate(out):
    short = o["name"].replace("void gm::gemm_umma_kernel", "gemm_umma").replace("(int)", "").replace("(bool)", "")
    short = short.split("(CUtensorMap")[0]
    print("| %d | `%s` | %.1f | %.1f | %.1f | %.1f | %d |" % (i, short, o["us"], o["rd"] / 1e6, o["wr"] / 1e6, o["tensor"], o["regs"]))
nt = [o for o in out if "gemm_umma_kernel<208" in o["name"].replace("(int)", "").replace(" ", "")]
if nt:
    avg = sum(o["rd"] + o["wr"] for o in nt) / len(nt)
    print("\nK-major 128x208 launches: %d, mean DRAM bytes per launch %.1f MB" % (len(nt), avg / 1e6))
    if "--write-traffic" in sys.argv:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        wl = "ns"
        for a in sys.argv[2:]:
            if a.startswith("--workload="):
                wl = a.split("=", 1)[1]
        path = os.path.join(root, "profiles", "traffic.json")
        cur = json.load(open(path)) if os.path.exists(path) else {}
        if "dram_bytes_per_launch" in cur:      # round-1 flat format
            cur = {}
        cur[wl] = {"kernel": "gemm_umma<208,0,K-major> (%d launches captured)" % len(nt),
                   "dram_bytes_per_launch": int(avg), "launches_captured": len(nt),
                   "source": "%s (ncu --set full --clock-control none, bench.py --workload %s --steps 2)" % (os.path.basename(rep), wl)}
        json.dump(cur, open(path, "w"), indent=1)
