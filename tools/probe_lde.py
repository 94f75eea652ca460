"""GPU probe / ncu target: LDE (K3) of a column slab; prints device time and algorithmic GB/s.

    python tools/probe_lde.py [cols] [log_n] [rate_bits]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from plonky2_bn254_b200 import ffi
from util import rand_field

cols = int(sys.argv[1]) if len(sys.argv) > 1 else 200
log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 19
r = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = ffi.Context(0)
v = rand_field(np.random.default_rng(0), (cols, 1 << log_n))
for it in range(3):
    ctx.commit(v, r, 4)
tm = dict(ctx.timings())
n = 1 << log_n
print(f"cols={cols} n=2^{log_n} r={r}: lde {tm['lde']:.3f} ms ({8 * cols * n * (2 + (1 << r)) / tm['lde'] / 1e6:.1f} GB/s algorithmic)")
