"""GPU probe: full proofs, parity against the oracle (optional) and per-stage device times."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from plonky2_bn254_b200 import ffi, inputs as I

check = "--check" in sys.argv
cases = [(2, 3), (0, 2), (1, 1)] if check else []
cases += [(0, 128), (0, 1024)] if "--big" in sys.argv else []
ctx = ffi.Context(0)
if check:
    from oracle import pyoracle as O
for kind, k in cases:
    inp, ts = I.make_inputs(kind, k, I.config_seed(3))
    t = time.time(); pf = ctx.prove(kind, inp, ts, keep_debug=check and k <= 3); wall = time.time() - t
    if k > 3:
        t = time.time(); pf = ctx.prove(kind, inp, ts); wall = time.time() - t
    w = pf.words()
    print(f"kind={kind} k={k}: wall {wall*1e3:.1f} ms, proof {w.size*8/1e6:.2f} MB", flush=True)
    tot = 0
    for name, ms in ctx.timings():
        print("   %-22s %9.3f ms" % (name, ms)); tot += ms
    print("   %-22s %9.3f ms" % ("sum of stages", tot), flush=True)
    if check and k <= 3:
        opf, tt, tp = O.prove_inputs(kind, inp, ts, keep_debug=True)
        print(f"   oracle: trace {tt:.2f}s prove {tp:.2f}s on {O.num_threads()} threads")
        for which, name in [(2, "challenges"), (0, "aux values"), (1, "quotient chunks"), (3, "query indices")]:
            a, b = opf.debug(which), pf.debug(which)
            print("   ", name, "OK" if a.shape == b.shape and (a == b).all() else "MISMATCH")
        ow = opf.words()
        same = ow.size == w.size and (ow == w).all()
        print("    proof bytes", "IDENTICAL" if same else "DIFFERENT", flush=True)
        if not same and ow.size == w.size:
            bad = np.nonzero(ow != w)[0]; print("    first diffs", bad[:10], len(bad))
        print("    oracle verifies GPU proof:", O.verify(w, inp, ts), flush=True)
