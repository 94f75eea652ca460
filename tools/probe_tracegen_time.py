"""GPU probe: device time of trace generation (second call, modules warm) for G1 and G2 at 1024 instances."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plonky2_bn254_b200 import ffi, inputs as I
ctx = ffi.Context(0)
for kind in (0, 1):
    inp, ts = I.make_inputs(kind, 1024, I.config_seed(4))
    for it in range(2):
        pf = ctx.prove(kind, inp, ts)
        t = dict(ctx.timings())
    print("kind", kind, "tracegen %.3f ms" % t["tracegen"], flush=True)
