"""Summarise an ncu --import-source report: top SASS instructions by stall samples.
usage: python tools/ncu_stalls.py report.ncu-rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = sum(int(r["# Samples"] or 0) for r in rows)
print(lines[0][:160]); print("total samples", tot, "instructions", len(rows))
stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
agg = {c: sum(int(r[c] or 0) for r in rows) for c in stall_cols}
print("by reason:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"] or 0))[:top]
for i in sorted(order):
    r = rows[i]
    why = sorted(((int(r[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r['# Samples']):6d} {100*int(r['# Samples'])/tot:5.1f}%  {r['Source'][:70]:70s} {why}")
