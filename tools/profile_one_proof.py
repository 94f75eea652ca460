"""For ncu --profile-from-start off: one warm-up proof, then exactly one config-2 proof inside cudaProfilerStart/Stop."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plonky2_bn254_b200 import ffi, inputs as I
k = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = ffi.Context(0)
inp, ts = I.make_inputs(0, k, I.config_seed(2))
ctx.prove(0, inp, ts).close()
torch.cuda.synchronize()
torch.cuda.profiler.start()
ctx.prove(0, inp, ts).close()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
