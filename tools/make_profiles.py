"""Turns the files a profiling gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/:
  launches_r1_final.csv    -> profiles/r1_launches_final.csv + r1_launches_final_summary.txt
  one_proof_raw.csv        -> profiles/r1_kernels_final.md (per-kernel DRAM throughput and pipe utilisation)
  prof_leafhash_r1_final   -> profiles/r1_leafhash_final.txt (tools/ncu_summary.py)
  bench_final.json         -> profiles/r1_bench_n1.json
"""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# ---- launch list ----
rows = list(csv.reader(open(os.path.join(G, "launches_r1_final.csv"))))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]; kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    agg[r[kn]][0] += 1; agg[r[kn]][1] += v
tot = sum(v[1] for v in agg.values())
out = ["# ncu --metrics gpu__time_duration.sum launch list of `python bench.py --steps 1 --warmup 1 --no-cpu-baseline` (4 proofs), per kernel",
       "# total %.1f ms over %d launches (cold-cache, serialised: compare shares)" % (tot / 1e6, sum(v[0] for v in agg.values()))]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:26]:
    out.append("%6.2f%% %5d launches %10.3f ms  %s" % (100 * v[1] / tot, v[0], v[1] / 1e6, k[:110]))
open(os.path.join(P, "r1_launches_final_summary.txt"), "w").write("\n".join(out) + "\n")
shutil.copy(os.path.join(G, "launches_r1_final.csv"), os.path.join(P, "r1_launches_final.csv"))
print("\n".join(out[:12]))

# ---- per-kernel table ----
rows = list(csv.reader(open(os.path.join(G, "one_proof_raw.csv"))))
hdr, units = rows[0], rows[1]; ix = {k: i for i, k in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]].replace(",", ""))
    except Exception: return 0.0
tu = units[ix["gpu__time_duration.sum"]]; ts = {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(tu, 1)
bu = units[ix["dram__bytes.sum.per_second"]]; bs = {"Gbyte/s": 1, "Tbyte/s": 1e3, "Mbyte/s": 1e-3, "Kbyte/s": 1e-6, "byte/s": 1e-9}.get(bu, 1)
agg = collections.OrderedDict()
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    a = agg.setdefault(name, dict(n=0, ms=0.0, gb=0.0, issue=0.0, alu=0.0, fma=0.0, regs=0, occ=0.0))
    ms = ts * f(r, "gpu__time_duration.sum")
    # units can differ per row in principle; ncu prints one unit row, values are already in it
    a["n"] += 1; a["ms"] += ms
    a["gb"] += ms * 1e-3 * f(r, "dram__bytes.sum.per_second") * bs
    a["issue"] += ms * f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
    a["alu"] += ms * f(r, "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active")
    a["fma"] += ms * f(r, "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed")
    a["occ"] += ms * f(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    a["regs"] = max(a["regs"], int(f(r, "launch__registers_per_thread")))
tot = sum(a["ms"] for a in agg.values())
lines = ["# Every major kernel of one config-2 proof (1024 G1 scalar-muls, 2^19 rows), end of round 1", "",
         "`ncu --section SpeedOfLight,ComputeWorkloadAnalysis,MemoryWorkloadAnalysis,LaunchStats,Occupancy,SchedulerStats --clock-control none` of `tools/profile_one_proof.py` (second proof, inside cudaProfilerStart/Stop). Per-launch times are cold-cache and serialised: compare shares. DRAM GB/s = `dram__bytes.sum.per_second` (read + write) averaged over the kernel's launches; peak = 6553 GB/s (MEASURED_PEAKS.json). Pipe columns are % of that pipe's peak while the kernel runs.", "",
         f"total of the listed kernels: {tot:.1f} ms", "",
         "| kernel | launches | ms | share | DRAM GB/s | % of HBM peak | issue-active % | ALU pipe % | FMA-heavy pipe % | warps active % | regs |",
         "|---|---|---|---|---|---|---|---|---|---|---|"]
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    gbs = a["gb"] / (a["ms"] * 1e-3) if a["ms"] else 0
    lines.append(f"| `{name[:70]}` | {a['n']} | {a['ms']:.2f} | {100*a['ms']/tot:.1f} % | {gbs:.0f} | {100*gbs/6553.3:.1f} | {a['issue']/a['ms']:.0f} | {a['alu']/a['ms']:.0f} | {a['fma']/a['ms']:.0f} | {a['occ']/a['ms']:.0f} | {a['regs']} |")
open(os.path.join(P, "r1_kernels_final.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[6:]))

# ---- leaf hash + bench ----
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, "prof_leafhash_r1_final.ncu-rep"),
                os.path.join(P, "r1_leafhash_final.txt"),
                "final round-1 leaf hash (merkle::k_leaf_hash, trace tree W=781, N=2^20, 102.8 M permutations), tools/profile_one_proof.py"],
               stdout=subprocess.DEVNULL)
shutil.copy(os.path.join(G, "bench_final.json"), os.path.join(P, "r1_bench_n1.json"))
d = json.load(open(os.path.join(P, "r1_bench_n1.json")))
print("bench:", d["value"], "proofs/s", d["ms_per_step"], "ms; e2e", d["e2e"]["value"], "; cpu", d["cpu_baseline"]["value"], "; roofline frac", d["roofline"]["frac"], "gperm/s", d["roofline"]["poseidon_gperm_per_s"], "; ntt GB/s", d["ntt"]["achieved"], d["clocks"])
