import numpy as np


def unpack(bits, n):
    return np.unpackbits(bits, axis=1)[:, :n].astype(np.float32)


def decoder_fixture_layers(g):
    """(trunk, head_hits, head_act) lists of (W, b) from tests/golden/decoder_v3.npz."""
    def chain(name, n):
        return [(g["%s_W%d" % (name, i)], g["%s_b%d" % (name, i)]) for i in range(n)]
    return chain("trunk", int(g["n_trunk"])), chain("hits", int(g["n_head"])), chain("act", int(g["n_head"]))
