"""GPU probe: dump the per-group running totals of the quotient kernel at one point (debug artefact 5)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from plonky2_bn254_b200 import ffi, inputs as I
ctx = ffi.Context(0)
for kind in (0, 1, 2):
    inp, ts = I.make_inputs(kind, 1, I.config_seed(3))
    pf = ctx.prove(kind, inp, ts, keep_debug=True)
    np.save("gpurun_out/groups_kind%d.npy" % kind, pf.debug(5))
    np.save("gpurun_out/qvals_kind%d.npy" % kind, pf.debug(4))
    print("saved", kind, flush=True)
