"""GPU probe: proofs/s with 1, 2 or 3 prover contexts (own stream each) driven by host threads on one GPU."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from plonky2_bn254_b200 import ffi, inputs as I

K = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
batches = [I.make_inputs(0, K, I.config_seed(2) + b) for b in range(3)]
for nctx in (1, 2, 3):
    ctxs = [ffi.Context(0) for _ in range(nctx)]
    for c in ctxs:
        c.prove(0, *batches[0]).close()  # warm-up (arena allocation)
    per = 6
    def work(c):
        for i in range(per):
            c.prove(0, *batches[i % 3]).close()
    ths = [threading.Thread(target=work, args=(c,)) for c in ctxs]
    t = time.time()
    for th in ths: th.start()
    for th in ths: th.join()
    dt = time.time() - t
    print(f"contexts={nctx}: {nctx * per / dt:.3f} proofs/s ({dt / (nctx * per) * 1e3:.1f} ms per proof)", flush=True)
    for c in ctxs: c.close()
