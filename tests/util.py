import numpy as np

GL_P = 0xFFFFFFFF00000001


def rand_field(rng, shape):
    """uniform canonical Goldilocks elements"""
    hi = rng.integers(0, 1 << 32, size=shape, dtype=np.uint64)
    lo = rng.integers(0, 1 << 32, size=shape, dtype=np.uint64)
    v = (hi << np.uint64(32)) | lo
    return np.where(v >= np.uint64(GL_P), v - np.uint64(GL_P), v).astype(np.uint64)
