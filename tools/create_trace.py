"""Where the time of gpv_create goes (GPV_TRACE_CREATE=1 prints the library's phases to stderr) next to the wall time of
the ctypes wrapper, whose column-major copies of the numpy arrays dominate (R holds its matrices column-major already).
    GPV_TRACE_CREATE=1 python tools/create_trace.py
B200, n = 1e6, m = 30: library 48 ms (ids / classes / obs 17, revCond 12, locality layer 7, locs 5), wrapper 400 ms."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import gpvecchia_b200 as G
from gpvecchia_b200 import harness as H
n, m = 1000000, 30
locs = H.make_locs(n, 2, stream=2)
NN = H.rev(H.ordered_nn_gpu(locs, m)).astype(np.int32)
Cond = np.zeros_like(NN, dtype=np.int32)
for k in range(3):
    t0 = time.perf_counter()
    h = G.UHandle(locs, NN, Cond, obs=np.ones(n, dtype=bool))
    print("UHandle total", (time.perf_counter() - t0) * 1e3, "ms", flush=True)
    t0 = time.perf_counter(); h.close(); print("close", (time.perf_counter() - t0) * 1e3, "ms", flush=True)
