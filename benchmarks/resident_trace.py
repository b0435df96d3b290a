"""Latency trace of the resident kernel (debug build: RBM_B200_NVCC_DEFS=-DRBM_RES_TRACE python -m divae_b200.build).
Prints, for half-steps 200..215 of CTA 0, the SM-clock offsets of: words complete -> K blocks expanded -> MMA issue ->
accumulator complete -> epilogue phases -> exchange issued."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import divae_b200  # noqa: E402
from divae_b200 import _lib  # noqa: E402

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
K = 300
torch.manual_seed(32)
rbm = divae_b200.RBM(n, n)
with torch.no_grad():
    rbm._weights.mul_(5.0)
rbm = rbm.cuda()
z = rbm.estimate_logZ(n_chains=chains, betas=np.linspace(0.0, 1.0, K + 1), seed=0x5EED)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH) if hasattr(_lib, "LIB_PATH") else _lib.load()
buf = (ctypes.c_ulonglong * (64 * 16))()
lib.rbm_b200_debug_resident_trace.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
rc = lib.rbm_b200_debug_resident_trace(buf, 64 * 16)
a = np.array(buf[:], dtype=np.int64).reshape(16, 64)
names = {0: "words_in", 1: "exp_kb0", 2: "exp_kb1", 3: "exp_kb2", 4: "exp_kb3", 5: "exp_kb4", 6: "exp_kb5", 7: "exp_kb6", 8: "exp_kb7",
         10: "mma_kb0", 11: "mma_kb1", 12: "mma_kb2", 13: "mma_kb3", 14: "mma_kb4", 15: "mma_kb5", 16: "mma_kb6", 17: "mma_kb7",
         18: "mma_commit", 20: "epi4_acc", 21: "epi4_tmem_read", 22: "epi4_math", 23: "epi4_sent",
         24: "epi11_acc", 25: "epi11_tmem_read", 26: "epi11_math", 27: "epi11_sent"}
print("rc", rc, "log Z", z)
# events of slot s are recorded at index 32 s + event; offsets are SM cycles from the moment slot 0's words of half-step t arrived
for t in range(1, 15):
    base = a[t][0]
    ev = []
    for s in range(2):
        for k in names:
            for tt in (t, t + 1):
                if a[tt][32 * s + k]:
                    ev.append((int(a[tt][32 * s + k] - base), "t%d.s%d.%s" % (200 + tt, s, names[k])))
    ev = sorted(x for x in ev if -2000 <= x[0] <= 14000)
    print("t=%d (dir %d): " % (200 + t, (200 + t) & 1) + "  ".join("%s %d" % (nm, dt) for dt, nm in ev))
