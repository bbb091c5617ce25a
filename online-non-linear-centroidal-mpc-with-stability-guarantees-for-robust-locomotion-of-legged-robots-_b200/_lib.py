"""ctypes binding of the C ABI declared in include/cmpc.h (libcmpc_b200.so).

The library is the product; this file only passes pointers.  Device buffers are torch CUDA tensors
(plumbing), host buffers are numpy arrays.  Nothing here falls back to a CPU implementation.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.environ.get("CMPC_LIB") or os.path.join(_HERE, "libcmpc_b200.so")   # CMPC_LIB: A/B builds of the same source (scripts/ab.sh)
NX, NU = 20, 32

STATUS_NAMES = {0: "converged", 1: "max_iter", 2: "line_search", 3: "regularization", 4: "infeasible_x0", 5: "nan"}
COLD, WARM_PRIMAL, WARM_FULL, WARM_SHIFTED = 0, 1, 2, 3


class CmpcError(RuntimeError):
    pass


class _Config(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int32), ("max_iter", ctypes.c_int32), ("ls_max", ctypes.c_int32), ("threads", ctypes.c_int32)] + \
               [(n, ctypes.c_double) for n in ("delta", "grav", "mu_fric", "foot_half_len", "foot_half_wid", "w_h", "w_xy",
                                               "w_zc", "w_foot", "w_sym", "w_swing", "w_rate", "eps_reg", "pz_max")] + \
               [("box", ctypes.c_double * 3)] + \
               [(n, ctypes.c_double) for n in ("relax", "mu_init", "mu_final", "mu_warm", "tol", "kappa_eps", "kappa_mu",
                                               "theta_mu", "tau_min", "bound_push")]


def library_path() -> str:
    return _SO


def nvcc_command(out: str = _SO, extra=()):
    return ["nvcc", *extra, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
            "--expt-relaxed-constexpr", "-diag-suppress", "170", "-Xptxas", "-v", "-shared", "-Xcompiler", "-fPIC",
            "-o", out, os.path.join(_HERE, "csrc", "cmpc_kernels.cu")]


def build_library(force: bool = False, verbose: bool = False, profile: bool = None) -> str:
    """Compile csrc/cmpc_kernels.cu for sm_100a into libcmpc_b200.so (in-tree)."""
    srcs = [os.path.join(_HERE, "csrc", f) for f in ("cmpc_kernels.cu", "cmpc_solver.h", "cmpc_model.h")]
    srcs.append(os.path.join(_ROOT, "include", "cmpc.h"))
    if not force and os.path.exists(_SO) and all(os.path.getmtime(s) <= os.path.getmtime(_SO) for s in srcs):
        return _SO
    if profile is None:
        profile = os.environ.get("CMPC_PROFILE", "0") == "1"
    res = subprocess.run(nvcc_command(extra=("-DCMPC_PROFILE",) if profile else ()), capture_output=True, text=True)
    if verbose or res.returncode:
        print(res.stdout + res.stderr)
    if res.returncode:
        raise CmpcError("nvcc failed building libcmpc_b200.so")
    return _SO


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise CmpcError("libcmpc_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                        "this package has no CPU fallback")
    L = ctypes.CDLL(_SO)
    dp, ip = ctypes.c_void_p, ctypes.c_void_p
    L.cmpc_last_error.restype = ctypes.c_char_p
    L.cmpc_version.restype = ctypes.c_char_p
    L.cmpc_default_config.argtypes = [ctypes.c_int32, ctypes.POINTER(_Config)]
    L.cmpc_create.argtypes = [ctypes.POINTER(_Config), ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p)]
    L.cmpc_destroy.argtypes = [ctypes.c_void_p]
    L.cmpc_solve_device.argtypes = [ctypes.c_void_p, ctypes.c_int32] + [dp] * 6 + [ctypes.c_int32] + [dp] * 5 + [ip, ip, ctypes.c_void_p]
    L.cmpc_solve_host.argtypes = [ctypes.c_void_p, ctypes.c_int32] + [dp] * 6 + [ctypes.c_int32] + [dp] * 5 + [ip, ip]
    L.cmpc_get_trajectory.argtypes = [ctypes.c_void_p, ctypes.c_int32, dp, dp]
    L.cmpc_set_warm.argtypes = [ctypes.c_void_p, ctypes.c_int32, dp, dp]
    L.cmpc_reset_warm.argtypes = [ctypes.c_void_p]
    L.cmpc_warm_save.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    L.cmpc_warm_restore.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
    L.cmpc_last_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64),
                                  ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)]
    L.cmpc_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
    L.cmpc_footprint.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]
    L.cmpc_measure_fp64_peak.argtypes = [ctypes.c_int32, ctypes.POINTER(ctypes.c_double)]
    _lib = L
    return L


EXPORTS = ["cmpc_default_config", "cmpc_create", "cmpc_destroy", "cmpc_last_error", "cmpc_version", "cmpc_solve_device",
           "cmpc_solve_host", "cmpc_get_trajectory", "cmpc_set_warm", "cmpc_reset_warm", "cmpc_warm_save", "cmpc_warm_restore", "cmpc_last_stats",
           "cmpc_footprint", "cmpc_phase_cycles", "cmpc_measure_fp64_peak"]


def _check(L, rc, what):
    if rc != 0:
        raise CmpcError("%s failed (%d): %s" % (what, rc, L.cmpc_last_error().decode()))


def measure_fp64_peak(device: int = 0) -> float:
    L = load()
    v = ctypes.c_double()
    _check(L, L.cmpc_measure_fp64_peak(device, ctypes.byref(v)), "cmpc_measure_fp64_peak")
    return v.value


def default_config(N: int) -> dict:
    L = load()
    c = _Config()
    _check(L, L.cmpc_default_config(N, ctypes.byref(c)), "cmpc_default_config")
    out = {}
    for name, _ in _Config._fields_:
        v = getattr(c, name)
        out[name] = list(v) if name == "box" else v
    return out


class BatchSolver:
    """B independent centroidal-MPC instances solved per call, one CTA each, on one GPU."""

    def __init__(self, N: int, batch_capacity: int, device: int = 0, **overrides):
        L = load()
        self._L = L
        self.N, self.capacity, self.device = int(N), int(batch_capacity), int(device)
        cfg = _Config()
        _check(L, L.cmpc_default_config(self.N, ctypes.byref(cfg)), "cmpc_default_config")
        for k, v in overrides.items():
            if not hasattr(cfg, k):
                raise CmpcError("unknown config field %r" % k)
            if k == "box":
                for j in range(3):
                    cfg.box[j] = float(v[j])
            else:
                setattr(cfg, k, v)
        self.config = cfg
        self._h = ctypes.c_void_p()
        _check(L, L.cmpc_create(ctypes.byref(cfg), self.capacity, self.device, ctypes.byref(self._h)), "cmpc_create")
        self._dev_out = None

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.cmpc_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host buffers (numpy): H2D + solve + D2H inside the call
    def solve_host(self, x0, com_ref, foot_ref, gamma, mass, k1, warm_mode=COLD):
        N = self.N
        x0 = np.ascontiguousarray(x0, np.float64).reshape(-1, NX)
        B = x0.shape[0]
        com_ref = np.ascontiguousarray(com_ref, np.float64).reshape(B, N, 9)
        foot_ref = np.ascontiguousarray(foot_ref, np.float64).reshape(B, N, 8)
        gamma = np.ascontiguousarray(gamma, np.float64).reshape(B, N + 1, 2)
        mass = np.ascontiguousarray(np.broadcast_to(np.asarray(mass, np.float64), (B,)))
        k1 = np.ascontiguousarray(np.broadcast_to(np.asarray(k1, np.float64), (B,)))
        out = {"x1": np.empty((B, NX)), "u0": np.empty((B, NU)), "xN": np.empty((B, NX)), "cost": np.empty(B),
               "viol": np.empty(B), "status": np.empty(B, np.int32), "iters": np.empty(B, np.int32)}
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        rc = self._L.cmpc_solve_host(self._h, B, p(x0), p(com_ref), p(foot_ref), p(gamma), p(mass), p(k1), int(warm_mode),
                                     p(out["x1"]), p(out["u0"]), p(out["xN"]), p(out["cost"]), p(out["viol"]),
                                     p(out["status"]), p(out["iters"]))
        _check(self._L, rc, "cmpc_solve_host")
        return out

    # ---- device buffers (torch CUDA tensors, float64 / int32, contiguous); asynchronous on `stream`
    def solve_device(self, x0, com_ref, foot_ref, gamma, mass, k1, warm_mode=COLD, out=None, stream=None):
        import torch
        B = x0.shape[0]
        for t in (x0, com_ref, foot_ref, gamma, mass, k1):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
                raise CmpcError("solve_device needs contiguous float64 CUDA tensors")
        if out is None:
            dev = x0.device
            out = {"x1": torch.empty((B, NX), dtype=torch.float64, device=dev), "u0": torch.empty((B, NU), dtype=torch.float64, device=dev),
                   "xN": torch.empty((B, NX), dtype=torch.float64, device=dev), "cost": torch.empty(B, dtype=torch.float64, device=dev),
                   "viol": torch.empty(B, dtype=torch.float64, device=dev), "status": torch.empty(B, dtype=torch.int32, device=dev),
                   "iters": torch.empty(B, dtype=torch.int32, device=dev)}
        if stream is None:
            stream = torch.cuda.current_stream(x0.device).cuda_stream
        if not stream:
            stream = 1          # cudaStreamLegacy: torch's default stream (NULL means "the handle's own stream" in the C ABI)
        rc = self._L.cmpc_solve_device(self._h, B, x0.data_ptr(), com_ref.data_ptr(), foot_ref.data_ptr(), gamma.data_ptr(),
                                       mass.data_ptr(), k1.data_ptr(), int(warm_mode), out["x1"].data_ptr(), out["u0"].data_ptr(),
                                       out["xN"].data_ptr(), out["cost"].data_ptr(), out["viol"].data_ptr(),
                                       out["status"].data_ptr(), out["iters"].data_ptr(), ctypes.c_void_p(stream))
        _check(self._L, rc, "cmpc_solve_device")
        return out

    def trajectory(self, batch: int):
        X = np.empty((batch, self.N + 1, NX))
        U = np.empty((batch, self.N, NU))
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _check(self._L, self._L.cmpc_get_trajectory(self._h, batch, p(X), p(U)), "cmpc_get_trajectory")
        return X, U

    def set_warm(self, X, U):
        X = np.ascontiguousarray(X, np.float64).reshape(-1, self.N + 1, NX)
        U = np.ascontiguousarray(U, np.float64).reshape(-1, self.N, NU)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _check(self._L, self._L.cmpc_set_warm(self._h, X.shape[0], p(X), p(U)), "cmpc_set_warm")

    def warm_save(self, batch: int):
        _check(self._L, self._L.cmpc_warm_save(self._h, batch), "cmpc_warm_save")

    def warm_restore(self, batch: int, stream=None):
        _check(self._L, self._L.cmpc_warm_restore(self._h, batch, ctypes.c_void_p(stream or 0)), "cmpc_warm_restore")


    def reset_warm(self):
        _check(self._L, self._L.cmpc_reset_warm(self._h), "cmpc_reset_warm")

    def last_stats(self) -> dict:
        it, nf, nr = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        ms, nl = ctypes.c_double(), ctypes.c_int32()
        _check(self._L, self._L.cmpc_last_stats(self._h, ctypes.byref(it), ctypes.byref(nf), ctypes.byref(nr),
                                                ctypes.byref(ms), ctypes.byref(nl)), "cmpc_last_stats")
        return {"iters": it.value, "nfact": nf.value, "nreg": nr.value, "kernel_ms": ms.value, "launches": nl.value}

    def phase_cycles(self) -> dict:
        arr = (ctypes.c_uint64 * 11)()
        _check(self._L, self._L.cmpc_phase_cycles(self._h, arr), "cmpc_phase_cycles")
        names = ["eval", "assemble", "pba", "factor", "store", "forward", "slack", "trial", "apply", "solve_total", "cta_total"]
        return dict(zip(names, [int(v) for v in arr]))

    def footprint(self) -> dict:
        a, b = ctypes.c_size_t(), ctypes.c_size_t()
        _check(self._L, self._L.cmpc_footprint(self._h, ctypes.byref(a), ctypes.byref(b)), "cmpc_footprint")
        return {"work_bytes_per_instance": a.value, "smem_bytes_per_cta": b.value}
