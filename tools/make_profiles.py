"""Turns the files a profiling gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/.

    python tools/make_profiles.py TAG ["note"]

reads (whichever exist)                         writes
  gpurun_out/launches_TAG.csv                   profiles/TAG_launches.csv, TAG_launches_summary.txt
  gpurun_out/one_proof_raw_TAG.csv              profiles/TAG_kernels.md   (per-kernel DRAM throughput / pipe utilisation)
  gpurun_out/prof_TAG*.ncu-rep  (--set full)    profiles/TAG_<name>.txt   (tools/ncu_summary.py)
  gpurun_out/bench_TAG.json                     profiles/TAG_bench_n1.json
and profiles/TAG_metrics.json: the measured per-kernel numbers bench.py quotes next to its live timings
(DRAM bytes per launch, thread-instructions, pipe utilisation), with the commit and date they were taken at.
bench.py reads the file profiles/current_metrics.json points to; nothing in the bench line is typed in by hand.
"""
import collections, csv, datetime, glob, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1]
NOTE = sys.argv[2] if len(sys.argv) > 2 else ""


def fnum(s):
    try:
        return float(str(s).replace(",", ""))
    except ValueError:
        return 0.0


metrics = {"tag": TAG, "note": NOTE, "date": datetime.date.today().isoformat(),
           "commit": subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True,
                                    text=True).stdout.strip(),
           "kernels": {}}

# ---- launch list ------------------------------------------------------------------------------------
src = os.path.join(G, f"launches_{TAG}.csv")
if os.path.exists(src):
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]; kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        agg[r[kn]][0] += 1; agg[r[kn]][1] += fnum(r[mv])
    tot = sum(v[1] for v in agg.values())
    out = ["# ncu --metrics gpu__time_duration.sum --clock-control none launch list of `python bench.py --steps 1 --warmup 1 "
           "--no-cpu-baseline`, per kernel",
           "# total %.1f ms over %d launches (cold-cache, serialised: compare shares)" % (tot / 1e6, sum(v[0] for v in agg.values()))]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        out.append("%6.2f%% %5d launches %10.3f ms  %s" % (100 * v[1] / tot, v[0], v[1] / 1e6, k[:110]))
    open(os.path.join(P, f"{TAG}_launches_summary.txt"), "w").write("\n".join(out) + "\n")
    shutil.copy(src, os.path.join(P, f"{TAG}_launches.csv"))
    metrics["launch_list_share"] = {k.split("(")[0][:60]: v[1] / tot for k, v in agg.items() if v[1] / tot > 0.01}
    print("\n".join(out[:14]))

# ---- per-kernel table -------------------------------------------------------------------------------
src = os.path.join(G, f"one_proof_raw_{TAG}.csv")
if os.path.exists(src):
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[h], rows[h + 1]; ix = {k: i for i, k in enumerate(hdr)}
    f = lambda r, k: fnum(r[ix[k]]) if k in ix and ix[k] < len(r) else 0.0
    tu = units[ix["gpu__time_duration.sum"]]; ts = {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(tu, 1)
    bu = units[ix["dram__bytes.sum.per_second"]]
    bs = {"Gbyte/s": 1, "Tbyte/s": 1e3, "Mbyte/s": 1e-3, "Kbyte/s": 1e-6, "byte/s": 1e-9}.get(bu, 1)
    agg = collections.OrderedDict()
    for r in rows[h + 2:]:
        if len(r) < len(hdr):
            continue
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, dict(n=0, ms=0.0, gb=0.0, issue=0.0, alu=0.0, fma=0.0, regs=0, occ=0.0))
        ms = ts * f(r, "gpu__time_duration.sum")
        a["n"] += 1; a["ms"] += ms
        a["gb"] += ms * 1e-3 * f(r, "dram__bytes.sum.per_second") * bs
        a["issue"] += ms * f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
        a["alu"] += ms * f(r, "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active")
        a["fma"] += ms * f(r, "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed")
        a["occ"] += ms * f(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
        a["regs"] = max(a["regs"], int(f(r, "launch__registers_per_thread")))
    tot = sum(a["ms"] for a in agg.values())
    lines = [f"# Every kernel of one config-2 proof (1024 G1 scalar-muls, 2^19 rows) - {TAG} {NOTE}", "",
             "`ncu --section SpeedOfLight,ComputeWorkloadAnalysis,MemoryWorkloadAnalysis,LaunchStats,Occupancy,SchedulerStats "
             "--clock-control none` of `tools/profile_one_proof.py` (second proof, inside cudaProfilerStart/Stop). Per-launch "
             "times are cold-cache and serialised: compare shares. DRAM GB/s = `dram__bytes.sum.per_second` (read + write) "
             "averaged over the kernel's launches; peak = 6553 GB/s (MEASURED_PEAKS.json). Pipe columns are % of that "
             "pipe's peak while the kernel runs.", "", f"total of the listed kernels: {tot:.1f} ms", "",
             "| kernel | launches | ms | share | DRAM GB/s | % of HBM peak | issue-active % | ALU pipe % | FMA-heavy pipe % | warps active % | regs |",
             "|---|---|---|---|---|---|---|---|---|---|---|"]
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        if a["ms"] <= 0:
            continue
        gbs = a["gb"] / (a["ms"] * 1e-3)
        lines.append(f"| `{name[:70]}` | {a['n']} | {a['ms']:.2f} | {100*a['ms']/tot:.1f} % | {gbs:.0f} | {100*gbs/6553.3:.1f} | "
                     f"{a['issue']/a['ms']:.0f} | {a['alu']/a['ms']:.0f} | {a['fma']/a['ms']:.0f} | {a['occ']/a['ms']:.0f} | {a['regs']} |")
        metrics["kernels"].setdefault(name[:70], {}).update(
            launches=a["n"], ms_under_ncu=a["ms"], share=a["ms"] / tot, dram_gb_s=gbs, issue_active=a["issue"] / a["ms"] / 100,
            alu_pipe=a["alu"] / a["ms"] / 100, fma_heavy_pipe=a["fma"] / a["ms"] / 100, warps_active=a["occ"] / a["ms"] / 100,
            registers=a["regs"])
    open(os.path.join(P, f"{TAG}_kernels.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[6:26]))

# ---- full captures -----------------------------------------------------------------------------------
for rep in sorted(glob.glob(os.path.join(G, f"prof_{TAG}*.ncu-rep"))):
    name = os.path.basename(rep)[len("prof_"):-len(".ncu-rep")]
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, os.path.join(P, f"{name}.txt"),
                    f"{NOTE} (ncu --set full --clock-control none, tools/profile_one_proof.py)"], stdout=subprocess.DEVNULL)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr = rows[0]; ix = {k: i for i, k in enumerate(hdr)}
    units = rows[1]

    def bytes_of(r, k):
        u = units[ix[k]]
        return fnum(r[ix[k]]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    for r in rows[2:]:
        kname = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")[:70]
        grid = r[ix["Grid Size"]] if "Grid Size" in ix else ""
        full = metrics["kernels"].setdefault(kname, {}).setdefault("full_captures", [])
        tu = units[ix["gpu__time_duration.sum"]]
        full.append({
            "grid": grid,
            "ms": fnum(r[ix["gpu__time_duration.sum"]]) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(tu, 1),
            "dram_bytes": bytes_of(r, "dram__bytes_read.sum") + bytes_of(r, "dram__bytes_write.sum"),
            "thread_instructions": fnum(r[ix["smsp__thread_inst_executed.sum"]]) if "smsp__thread_inst_executed.sum" in ix else None,
            "warp_instructions": fnum(r[ix["smsp__inst_executed.sum"]]) if "smsp__inst_executed.sum" in ix else None,
        })
    print("summarised", rep)

src = os.path.join(G, f"bench_{TAG}.json")
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, f"{TAG}_bench_n1.json"))
    d = json.loads(open(src).read().strip().splitlines()[-1])
    print("bench:", d["value"], d["unit"], d["ms_per_step"], "ms; e2e", d["e2e"]["value"], "; roofline frac", d["roofline"]["frac"],
          "; ntt", d["ntt"]["ms_per_proof"], "ms", d["clocks"])
json.dump(metrics, open(os.path.join(P, f"{TAG}_metrics.json"), "w"), indent=1)
json.dump({"file": f"{TAG}_metrics.json"}, open(os.path.join(P, "current_metrics.json"), "w"))
print("wrote", f"profiles/{TAG}_metrics.json")
