"""GPU probe: BASELINE configs 1-4 at full size (G2 x 2^10; fq_exp x 2^12 with blow-up 2 and 8),
verified by the product's host verifier."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from plonky2_bn254_b200 import ffi, inputs as I

ctx = ffi.Context(0)
cases = [("config 1: G1 x 1 (2^16 rows = min_rows)", 0, 1, None), ("config 2: G1 x 1024", 0, 1024, None),
         ("config 3: G2 x 1024", 1, 1024, None), ("config 4a: fq_exp x 4096, rate_bits 1", 2, 4096, None),
         ("config 4b: fq_exp x 4096, rate_bits 3, 28 queries", 2, 4096, (3, 28))]
for name, kind, k, rb in cases:
    t = time.time(); inp, ts = I.make_inputs(kind, k, I.config_seed(3 + kind)); tgen = time.time() - t
    cfg = None
    if rb:
        cfg = ctx.L.standard_fast_config(); cfg.rate_bits, cfg.num_query_rounds = rb
    try:
        ctx.prove(kind, inp, ts, config=cfg).close()
        t = time.time(); pf = ctx.prove(kind, inp, ts, config=cfg); wall = time.time() - t
    except Exception as e:
        print(name, "FAILED:", e, flush=True); continue
    w = pf.words()
    tot = sum(ms for _, ms in ctx.timings())
    t = time.time(); ok = ctx.L.verify(kind, w, inp, ts, config=cfg); tv = time.time() - t
    print(f"{name}: wall {wall*1e3:.1f} ms (stages {tot:.1f} ms), proof {w.size*8/1e6:.2f} MB, verify {ok} in {tv:.2f} s, inputs {tgen:.1f} s", flush=True)
    print("   " + ", ".join(f"{n} {ms:.1f}" for n, ms in ctx.timings()), flush=True)
