"""TEST AID ONLY: ctypes driver of tests/hostsim/libhostsim.so (the product solver core compiled by g++ with a
serial execution policy).  Used to debug/verify the algorithm without a GPU.  Not a product path."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, "online-non-linear-centroidal-mpc-with-stability-guarantees-for-robust-locomotion-of-legged-robots-_b200")
NX, NU, NR = 28, 32, 56


def build(force=False):
    so = os.path.join(HERE, "libhostsim.so")
    srcs = [os.path.join(HERE, "hostsim.cpp"), os.path.join(PKG, "csrc", "cmpc_solver.h"), os.path.join(PKG, "csrc", "cmpc_model.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(PKG, "csrc"),
                               srcs[0], "-o", so])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.hostsim_solve.restype = ctypes.c_int
        _lib.hostsim_work_doubles.restype = ctypes.c_int
    return _lib


def pack(prob):
    """oracle Problem -> the instance arrays of the C-ABI (instance-major)."""
    N = prob.N
    com = np.ascontiguousarray(prob.com_ref.T, float)                      # [N][9]
    foot = np.ascontiguousarray(np.concatenate([prob.pl_ref, prob.pr_ref, prob.al_ref[None], prob.ar_ref[None]], 0).T, float)
    gam = np.ascontiguousarray(np.stack([prob.gl, prob.gr], 1), float)     # [N+1][2]
    return np.ascontiguousarray(prob.x0, float), com, foot, gam


def solve(prob, work=None, warm=0, **over):
    L = lib()
    N = prob.N
    x0, com, foot, gam = pack(prob)
    keys = ["eps_reg", "relax", "mu_init", "mu_final", "tol", "max_iter", "ls_max", "w_rate", "mu_warm", "kappa_eps", "kappa_mu", "theta_mu", "tau_min", "warm_push", "warm_comp"] + ["xp%d" % j for j in range(8)] + ["stall_window", "stall_final", "jam_window", "crawl_window", "crawl_alpha", "warm_stall_window"]
    cfg = np.full(len(keys), np.nan)
    if getattr(prob, "eps_reg", None) is not None:
        over.setdefault("eps_reg", prob.eps_reg)       # None: the product default (1e-5; the oracle uses 1e-9)
    over.setdefault("w_rate", prob.w_rate)
    for k, v in over.items():
        cfg[keys.index(k)] = v
    if work is None:
        work = np.zeros(L.hostsim_work_doubles(N))
    stats = np.zeros(8)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    L.hostsim_solve(ctypes.c_int(N), dp(x0), dp(com), dp(foot), dp(gam), ctypes.c_double(prob.mass),
                    ctypes.c_double(prob.k1), dp(cfg), ctypes.c_int(warm), dp(work), dp(stats))
    X = work[:(N + 1) * NX].reshape(N + 1, NX).T.copy()
    U = work[(N + 1) * NX:(N + 1) * NX + N * NU].reshape(N, NU).T.copy()
    return dict(X=X, U=U, cost=stats[0], viol=stats[1], kkt=stats[2], mu=stats[3], iters=int(stats[4]),
                status=int(stats[5]), nfact=int(stats[6]), nreg=int(stats[7]), work=work)
