"""GPU probe: BASELINE config 5 on ONE B200 - a single G1 proof of 8192 scalar-muls (2^22 rows, 24.4 GiB trace),
checked by the product's host verifier. Also 4096 scalar-muls (2^21 rows)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plonky2_bn254_b200 import ffi, inputs as I
ctx = ffi.Context(0)
for k in (4096, 8192):
    t = time.time(); inp, ts = I.make_inputs(0, k, I.config_seed(5)); tg = time.time() - t
    try:
        t = time.time(); pf = ctx.prove(0, inp, ts); wall = time.time() - t
    except Exception as e:
        print(k, "FAILED", e, flush=True); continue
    stages = ctx.timings()
    w = pf.words()
    t = time.time(); ok = ctx.L.verify(0, w, inp, ts); tv = time.time() - t
    print(f"G1 x {k}: rows 2^{int(w[2])}, wall {wall*1e3:.0f} ms (first call, includes the arena allocation), stages {sum(m for _, m in stages):.0f} ms, "
          f"proof {w.size*8/1e6:.2f} MB, verify {ok} in {tv:.1f} s, inputs generated in {tg:.1f} s", flush=True)
    print("   " + ", ".join(f"{n} {m:.1f}" for n, m in stages), flush=True)
