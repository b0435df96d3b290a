import numpy as np


def unpack(bits, n):
    return np.unpackbits(bits, axis=1)[:, :n].astype(np.float32)
