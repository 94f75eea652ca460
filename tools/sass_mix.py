"""Dynamic SASS instruction mix per kernel from an .ncu-rep captured with --import-source on.

    python tools/sass_mix.py gpurun_out/prof.ncu-rep [top_n]
"""
import collections, csv, io, re, subprocess, sys

ALU = ("IADD3", "LOP3", "SHF", "LEA", "PRMT", "SEL", "ISETP", "MOV", "IABS", "IMNMX", "VIADD", "VIMNMX", "SGXT", "BMSK", "FLO", "POPC", "PLOP3", "P2R", "R2P", "CS2R")
FMA = ("IMAD", "IDP", "FFMA", "FMUL", "FADD", "HFMA2", "HADD2", "HMUL2")


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    kernel, hdr = None, None
    mix = collections.OrderedDict()
    for row in csv.reader(io.StringIO(raw)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            kernel = row[1]
            mix[kernel] = collections.Counter()
            hdr = None
            continue
        if row[0] == "Address":
            hdr = row
            continue
        if hdr is None or kernel is None:
            continue
        src = row[hdr.index("Source")].strip()
        n = int(row[hdr.index("Instructions Executed")] or 0)
        m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", src)
        if m and n:
            mix[kernel][m.group(1)] += n
    for k, c in mix.items():
        tot = sum(c.values())
        alu = sum(v for o, v in c.items() if o.split(".")[0] in ALU)
        fma = sum(v for o, v in c.items() if o.split(".")[0] in FMA)
        print(f"{k[:70]}: {tot / 1e6:.1f} M warp-instr, ALU-pipe {alu / tot:.1%}, FMA-pipe {fma / tot:.1%}, other {(tot - alu - fma) / tot:.1%}")
        print("   " + ", ".join(f"{o} {v / tot:.1%}" for o, v in c.most_common(top)))


if __name__ == "__main__":
    main()
