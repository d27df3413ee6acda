"""Worker processes for the lock-step L-BFGS-B driver (batched.lockstep_lbfgsb, workers > 0).

SciPy's reverse-communication routine ``setulb`` costs ~13 us per call and holds the GIL; for the
5120 GPs of configuration C3 that is 1.5 s of a 1.9 s fit, against 0.3 s on the device.  The problems
are independent, so slices of the batch are advanced by separate Python processes, each running the
very same ``setulb`` sequence per problem as the in-process driver: the iterates stay bit-identical
to ``scipy.optimize.minimize(method="L-BFGS-B")``.

The workers import numpy and scipy only (no torch, no CUDA) and talk over pipes with length-prefixed
pickles; they exit when their stdin closes, so they cannot outlive the parent."""
from __future__ import annotations

import atexit
import os
import pickle
import struct
import subprocess
import sys
from typing import List

_WORKER_SRC = r'''
import pickle, struct, sys
import numpy as np
from scipy.optimize import _lbfgsb_py as _lb
_lbfgsb = _lb._lbfgsb
inp, out = sys.stdin.buffer, sys.stdout.buffer
def recv():
    h = inp.read(8)
    if len(h) < 8:
        sys.exit(0)
    (n,) = struct.unpack("<Q", h)
    return pickle.loads(inp.read(n))
def send(obj):
    b = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    out.write(struct.pack("<Q", len(b))); out.write(b); out.flush()
int_dtype = np.int64 if getattr(_lb, "HAS_ILP64", False) else np.int32
st = None
def advance(active):
    """run every problem in `active` until it asks for f, g (task 3) or stops; returns the waiting ones"""
    m, maxiter, maxfun = st["m"], st["maxiter"], st["maxfun"]
    args, NIT, NFEV = st["args"], st["NIT"], st["NFEV"]
    setulb = _lbfgsb.setulb
    waiting = []
    for b in active:
        a = args[b]
        task = a[11]
        while True:
            setulb(*a)
            t0 = task[0]
            if t0 == 3:
                waiting.append(b); break
            elif t0 == 1:
                NIT[b] += 1
                if NIT[b] >= maxiter:
                    task[0] = 5; task[1] = 504
                elif NFEV[b] > maxfun:
                    task[0] = 5; task[1] = 502
            else:
                break
    return waiting
while True:
    msg = recv()
    op = msg[0]
    if op == "init":
        _, X0, m, factr, gtol, maxls, maxiter, maxfun = msg
        B, n = X0.shape
        X = np.array(X0, dtype=np.float64)
        F = np.zeros((B,)); G = np.zeros((B, n))
        WA = np.zeros((B, 2 * m * n + 5 * n + 11 * m * m + 8 * m)); IWA = np.zeros((B, 3 * n), dtype=int_dtype)
        TASK = np.zeros((B, 2), dtype=int_dtype); LN = np.zeros((B, 2), dtype=int_dtype)
        LS = np.zeros((B, 4), dtype=int_dtype); IS = np.zeros((B, 44), dtype=int_dtype); DS = np.zeros((B, 29))
        nbd = np.zeros(n, dtype=int_dtype); low = np.zeros(n); up = np.zeros(n)
        args = [(m, X[b], low, up, nbd, F[b:b + 1].reshape(()), G[b], factr, gtol, WA[b], IWA[b], TASK[b], LS[b], IS[b], DS[b],
                 maxls, LN[b]) for b in range(B)]
        st = dict(m=m, maxiter=maxiter, maxfun=maxfun, args=args, X=X, F=F, G=G, TASK=TASK, NIT=np.zeros(B, np.int64),
                  NFEV=np.zeros(B, np.int64), waiting=[])
        st["waiting"] = advance(list(range(B)))
        w = np.asarray(st["waiting"], dtype=np.int64)
        send((w, X[w]))
    elif op == "step":
        _, fb, gb = msg
        w = np.asarray(st["waiting"], dtype=np.int64)
        st["F"][w] = fb; st["G"][w] = gb; st["NFEV"][w] += 1
        st["waiting"] = advance(st["waiting"])
        w = np.asarray(st["waiting"], dtype=np.int64)
        send((w, st["X"][w]))
    elif op == "finish":
        send((st["X"], st["F"], st["G"], st["TASK"], st["NIT"], st["NFEV"]))
        st = None
    elif op == "ping":
        send(("pong",))
    elif op == "quit":
        sys.exit(0)
'''


class _Worker:
    def __init__(self):
        # one BLAS / OpenMP thread per worker: the L-BFGS-B matrices are 2m x 2m, threads only contend
        env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
        self.p = subprocess.Popen([sys.executable, "-c", _WORKER_SRC], stdin=subprocess.PIPE, stdout=subprocess.PIPE, env=env)

    def send(self, obj):
        b = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
        self.p.stdin.write(struct.pack("<Q", len(b)))
        self.p.stdin.write(b)
        self.p.stdin.flush()

    def recv(self):
        h = self.p.stdout.read(8)
        if len(h) < 8:
            raise RuntimeError("L-BFGS-B worker process died (exit code %s)" % self.p.poll())
        (n,) = struct.unpack("<Q", h)
        return pickle.loads(self.p.stdout.read(n))

    def fileno(self) -> int:
        """for select(): a reply is waiting (replies are read whole, so nothing stays in Python's buffer)"""
        return self.p.stdout.fileno()

    def close(self):
        try:
            if self.p.poll() is None:
                self.send(("quit",))
                self.p.stdin.close()
                self.p.wait(timeout=2)
        except Exception:
            pass
        finally:
            if self.p.poll() is None:
                self.p.kill()


_pool: List[_Worker] = []


def get_workers(n: int) -> List[_Worker]:
    """A persistent pool (start-up costs an interpreter + numpy/scipy import per worker, ~0.5 s in parallel).
    New workers are pinged, so the pool is ready (imports done) when this returns."""
    fresh = []
    while len(_pool) < n:
        _pool.append(_Worker())
        fresh.append(_pool[-1])
    for wk in fresh:
        wk.send(("ping",))
    for wk in fresh:
        wk.recv()
    return _pool[:n]


def shutdown():
    while _pool:
        _pool.pop().close()


atexit.register(shutdown)
