import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from plonky2_bn254_b200 import ffi, inputs as I
ctx = ffi.Context(0)
for kind, k in [(0, 128), (0, 1024), (2, 1024), (1, 256)]:
    inp, ts = I.make_inputs(kind, k, 1)
    for _ in range(2):
        t = time.time(); tr = ctx.generate_trace(kind, inp, ts); wall = time.time() - t
    print(kind, k, tr.shape, ctx.timings(), "wall %.3f" % wall, flush=True)
