"""Per-kernel table from an .ncu-rep (ncu --set full of one proof): launches, total ms, share, DRAM GB/s and % of peak,
issue-active %, ALU / FMA pipe %, registers. Usage: python tools/ncu_table.py rep.ncu-rep out.md "title"."""
import collections, csv, io, subprocess, sys
rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return 0.0
def scale(k, v):  # to bytes / ms
    u = units[ix[k]]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1)
agg = collections.OrderedDict()
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    a = agg.setdefault(name, dict(n=0, ms=0.0, bytes=0.0, issue=0.0, alu=0.0, fma=0.0, regs=0, inst=0.0))
    ms = scale("gpu__time_duration.sum", f(r, "gpu__time_duration.sum"))
    a["n"] += 1; a["ms"] += ms
    a["bytes"] += scale("dram__bytes_read.sum", f(r, "dram__bytes_read.sum")) + scale("dram__bytes_write.sum", f(r, "dram__bytes_write.sum"))
    a["issue"] += ms * f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
    a["alu"] += ms * f(r, "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active")
    a["fma"] += ms * f(r, "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed")
    a["regs"] = int(f(r, "launch__registers_per_thread"))
    a["inst"] += f(r, "smsp__inst_executed.sum")
tot = sum(a["ms"] for a in agg.values())
lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full, cold caches, serialised; compare shares). total {tot:.1f} ms", "",
         "| kernel | launches | ms | share | DRAM GB/s | % of 6553 GB/s | issue-active % | ALU pipe % | FMA-heavy pipe % | regs | warp-instr |",
         "|---|---|---|---|---|---|---|---|---|---|---|"]
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    gbs = a["bytes"] / (a["ms"] * 1e-3) / 1e9 if a["ms"] else 0
    lines.append(f"| `{name[:60]}` | {a['n']} | {a['ms']:.3f} | {100*a['ms']/tot:.1f} % | {gbs:.0f} | {100*gbs/6553.3:.1f} | "
                 f"{a['issue']/a['ms']:.0f} | {a['alu']/a['ms']:.0f} | {a['fma']/a['ms']:.0f} | {a['regs']} | {a['inst']:.3g} |")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
