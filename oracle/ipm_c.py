"""ORACLE (test infrastructure) -- ctypes wrapper of the C oracle (oracle/ipm_c.c + generated oracle/_gen/stage_gen.c).

Only tests/, __graft_entry__.smoke()/build() and bench.py's CPU-baseline legs may import this module."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libipm_c.so")
NX, NU = 28, 32
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, "ipm_c.c"), os.path.join(HERE, "_gen", "stage_gen.c")]
    if not os.path.exists(srcs[1]):
        raise ImportError("oracle/_gen/stage_gen.c is missing (python -m oracle.gen_c)")
    if force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=native", "-shared", "-fPIC", "-o", SO, srcs[0], "-lm"])
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_solve.restype = ctypes.c_int
    return _lib


ORACLE_R = dict(tol=1e-3, mu_final=1e-4, ipopt_termination=1, max_iter=1000)     # the reference's IPOPT options (MPC file :128), defaults otherwise


def solve_packed(N, x0, com_ref, foot_ref, gamma, mass, k1, warm=None, **opts):
    """Instance-major arrays of the C ABI -> dict(status, iters, cost, viol, x1, u0, X, U).  `warm` = (X, U) primal warm start."""
    L = lib()
    keys = ["tol", "mu_init", "mu_final", "relax", "max_iter", "ls_max", "eps_reg", "w_rate", "ipopt_termination"]
    o = np.full(len(keys), np.nan)
    for k, v in opts.items():
        o[keys.index(k)] = v
    x0 = np.ascontiguousarray(x0, float); com = np.ascontiguousarray(com_ref, float).reshape(N, 9)
    foot = np.ascontiguousarray(foot_ref, float).reshape(N, 8); gam = np.ascontiguousarray(gamma, float).reshape(N + 1, 2)
    X = np.zeros((N + 1, NX)); U = np.zeros((N, NU))
    if warm is not None:
        X[:, :20] = np.asarray(warm[0], float).reshape(N + 1, -1)[:, :20]
        U[:] = np.asarray(warm[1], float).reshape(N, NU)
    stats = np.zeros(8)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    L.oracle_solve(ctypes.c_int(N), dp(x0), dp(com), dp(foot), dp(gam), ctypes.c_double(mass), ctypes.c_double(k1), dp(o),
                   ctypes.c_int(0 if warm is None else 1), dp(X), dp(U), dp(stats))
    return {"status": int(stats[4]), "iters": int(stats[3]), "cost": stats[0], "viol": stats[1], "kkt": stats[2], "nfact": int(stats[5]),
            "x1": X[1, :20].copy(), "u0": U[0].copy(), "X": X[:, :20].copy(), "U": U}
