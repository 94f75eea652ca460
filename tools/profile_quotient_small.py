import sys, os
sys.path.insert(0, "/root/repo")
import torch
from plonky2_bn254_b200 import ffi, inputs as I
ctx = ffi.Context(0)
inp, ts = I.make_inputs(0, 128, I.config_seed(2))
ctx.prove(0, inp, ts).close()
torch.cuda.synchronize(); torch.cuda.profiler.start()
ctx.prove(0, inp, ts).close()
torch.cuda.synchronize(); torch.cuda.profiler.stop()
