"""GPU probe: device time of LDE + Merkle for a slab of columns at a production height."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from plonky2_bn254_b200 import ffi
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import rand_field

ctx = ffi.Context(0)
rng = np.random.default_rng(0)
for cols, log_n in [(96, 19), (200, 19), (64, 21)]:
    v = rand_field(rng, (cols, 1 << log_n))
    for it in range(3):
        t = time.time()
        ctx.commit(v, 1, 4)
        wall = time.time() - t
    tm = dict(ctx.timings())
    n = 1 << log_n
    ntt_bytes = 8 * cols * n * 4
    mk_bytes = 8 * cols * n * 2 + 32 * (4 * n - 16)
    perms = ((cols + 7) // 8) * 2 * n + 2 * n
    print(f"cols={cols} n=2^{log_n}: lde {tm['lde']:.3f} ms ({ntt_bytes / tm['lde'] / 1e6:.1f} GB/s), "
          f"merkle {tm['merkle']:.3f} ms ({mk_bytes / tm['merkle'] / 1e6:.1f} GB/s, {perms / tm['merkle'] / 1e3:.1f} Mperm/s), "
          f"wall {wall * 1e3:.1f} ms", flush=True)
