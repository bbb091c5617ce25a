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
# built to <repo>/lib/ (a short path: the package directory carries the reference repository's 100-character name)
_SO = os.environ.get("CMPC_LIB") or os.path.join(_ROOT, "lib", "libcmpc_b200.so")   # CMPC_LIB: A/B builds of the same source (scripts/ab.sh)
NX, NU = 20, 32

STATUS_NAMES = {0: "converged", 1: "max_iter", 2: "line_search", 3: "regularization", 4: "infeasible_x0", 5: "nan", 6: "stall"}
COLD, WARM_PRIMAL, WARM_FULL, WARM_SHIFTED, WARM_AUTO = 0, 1, 2, 3, 4


class CmpcError(RuntimeError):
    pass


class _Config(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int32), ("max_iter", ctypes.c_int32), ("ls_max", ctypes.c_int32), ("threads", ctypes.c_int32),
                ("stall_window", ctypes.c_int32), ("stall_final", ctypes.c_int32),
                ("jam_window", ctypes.c_int32), ("reserved0", ctypes.c_int32)] + \
               [(n, ctypes.c_double) for n in ("delta", "grav", "mu_fric", "foot_half_len", "foot_half_wid", "w_h", "w_xy",
                                               "w_zc", "w_foot", "w_sym", "w_swing", "w_rate", "eps_reg", "pz_max")] + \
               [("box", ctypes.c_double * 3)] + \
               [(n, ctypes.c_double) for n in ("relax", "mu_init", "mu_final", "mu_warm", "tol", "kappa_eps", "kappa_mu",
                                               "theta_mu", "tau_min", "bound_push")]


class _WalkTables(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("N", "rate", "first_swing_left", "n_steps", "T_ref", "T_plan")] + \
               [(n, ctypes.c_void_p) for n in ("com_tab", "foot_tab", "gamma_tab", "step_index")]


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
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
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
    L.cmpc_reset_warm.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
    L.cmpc_solve_host_traj.argtypes = [ctypes.c_void_p, ctypes.c_int32] + [dp] * 6 + [ctypes.c_int32] + [dp] * 5 + [ip, ip, ctypes.c_int32, dp, dp]
    L.cmpc_warm_save.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    L.cmpc_warm_restore.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
    L.cmpc_last_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64),
                                  ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)]
    L.cmpc_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
    L.cmpc_footprint.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_int32),
                                 ctypes.POINTER(ctypes.c_size_t)]
    L.cmpc_assemble_device.argtypes = [ctypes.POINTER(_WalkTables), ctypes.c_int32, ctypes.c_int32] + [ctypes.c_void_p] * 13
    L.cmpc_measure_fp64_peak.argtypes = [ctypes.c_int32, ctypes.POINTER(ctypes.c_double)]
    _lib = L
    return L


EXPORTS = ["cmpc_default_config", "cmpc_create", "cmpc_destroy", "cmpc_last_error", "cmpc_version", "cmpc_solve_device",
           "cmpc_solve_host", "cmpc_solve_host_traj", "cmpc_get_trajectory", "cmpc_set_warm", "cmpc_reset_warm", "cmpc_warm_save", "cmpc_warm_restore", "cmpc_last_stats",
           "cmpc_footprint", "cmpc_phase_cycles", "cmpc_measure_fp64_peak", "cmpc_assemble_device", "cmpc_qp_solve_device", "cmpc_qp_solve_host"]


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


class WalkTables:
    """Device tables of one walk for `cmpc_assemble_device` (torch CUDA tensors are the storage, the kernel does the gather)."""

    def __init__(self, plan_tables, ref_tables, N, rate, first_swing_left, n_steps, device=0):
        import torch
        self.device = int(device)
        dev = torch.device("cuda", self.device)
        T = min(min(len(c) for c in ref_tables.com), len(ref_tables.pos_l))
        self.com_tab = torch.as_tensor(np.stack([c[:T] for c in ref_tables.com], axis=1), dtype=torch.float64, device=dev).contiguous()
        self.foot_tab = torch.as_tensor(np.concatenate([ref_tables.pos_l[:T], ref_tables.pos_r[:T], ref_tables.yaw_l[:T, None],
                                                        ref_tables.yaw_r[:T, None]], 1), dtype=torch.float64, device=dev).contiguous()
        self.gamma_tab = torch.as_tensor(plan_tables.gamma, dtype=torch.float64, device=dev).contiguous()
        self.step_index = torch.as_tensor(np.asarray(plan_tables.step_index, np.int32), device=dev).contiguous()
        self.is_ss = torch.as_tensor(np.asarray(plan_tables.is_ss), device=dev)
        self.left_support = torch.as_tensor(np.asarray(plan_tables.left_support), device=dev)
        self.T_ref, self.T_plan, self.N, self.rate, self.n_steps = int(T), int(plan_tables.T), int(N), int(rate), int(n_steps)
        self.c = _WalkTables(self.N, self.rate, 1 if first_swing_left else 0, self.n_steps, self.T_ref, self.T_plan,
                             self.com_tab.data_ptr(), self.foot_tab.data_ptr(), self.gamma_tab.data_ptr(), self.step_index.data_ptr())

    def assemble(self, tick, com_pos, com_vel, hw, theta, yaw, plan, out=None, stream=None):
        """tick: int32 [B] CUDA tensor; the rest float64 CUDA tensors ([B,3] x 4, [B,2], [B,n_steps,3]).  Returns
        (x0, com_ref, foot_ref, gamma, err) as CUDA tensors."""
        import torch
        L = load()
        B, N = int(tick.shape[0]), self.N
        want = [(tick, (B,), torch.int32), (com_pos, (B, 3), torch.float64), (com_vel, (B, 3), torch.float64), (hw, (B, 3), torch.float64),
                (theta, (B, 3), torch.float64), (yaw, (B, 2), torch.float64), (plan, (B, self.n_steps, 3), torch.float64)]
        for t, shp, dt in want:
            if not (t.is_cuda and t.is_contiguous() and t.dtype == dt and tuple(t.shape) == shp and t.device.index == self.device):
                raise CmpcError("assemble: expected a contiguous %s CUDA tensor of shape %s on cuda:%d" % (dt, shp, self.device))
        if out is None:
            dev = tick.device
            out = (torch.empty((B, NX), dtype=torch.float64, device=dev), torch.empty((B, N, 9), dtype=torch.float64, device=dev),
                   torch.empty((B, N, 8), dtype=torch.float64, device=dev), torch.empty((B, N + 1, 2), dtype=torch.float64, device=dev),
                   torch.empty((B,), dtype=torch.int32, device=dev))
        if stream is None:
            stream = torch.cuda.current_stream(tick.device).cuda_stream
        rc = L.cmpc_assemble_device(ctypes.byref(self.c), self.device, B, tick.data_ptr(), com_pos.data_ptr(), com_vel.data_ptr(), hw.data_ptr(),
                                    theta.data_ptr(), yaw.data_ptr(), plan.data_ptr(), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                    out[3].data_ptr(), out[4].data_ptr(), ctypes.c_void_p(stream))
        _check(L, rc, "cmpc_assemble_device")
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
    def solve_host(self, x0, com_ref, foot_ref, gamma, mass, k1, warm_mode=COLD, traj_batch=0):
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
        if traj_batch:                                         # full primal trajectories on the same synchronisation
            out["X"] = np.empty((traj_batch, N + 1, NX)); out["U"] = np.empty((traj_batch, N, NU))
            rc = self._L.cmpc_solve_host_traj(self._h, B, p(x0), p(com_ref), p(foot_ref), p(gamma), p(mass), p(k1), int(warm_mode),
                                              p(out["x1"]), p(out["u0"]), p(out["xN"]), p(out["cost"]), p(out["viol"]),
                                              p(out["status"]), p(out["iters"]), int(traj_batch), p(out["X"]), p(out["U"]))
            _check(self._L, rc, "cmpc_solve_host_traj")
            return out
        rc = self._L.cmpc_solve_host(self._h, B, p(x0), p(com_ref), p(foot_ref), p(gamma), p(mass), p(k1), int(warm_mode),
                                     p(out["x1"]), p(out["u0"]), p(out["xN"]), p(out["cost"]), p(out["viol"]),
                                     p(out["status"]), p(out["iters"]))
        _check(self._L, rc, "cmpc_solve_host")
        return out

    # ---- device buffers (torch CUDA tensors, float64 / int32, contiguous); asynchronous on `stream`
    def solve_device(self, x0, com_ref, foot_ref, gamma, mass, k1, warm_mode=COLD, out=None, stream=None):
        import torch
        B, N = x0.shape[0], self.N
        want = {"x0": (B, NX), "com_ref": (B, N, 9), "foot_ref": (B, N, 8), "gamma": (B, N + 1, 2), "mass": (B,), "k1": (B,)}
        for name, t in zip(want, (x0, com_ref, foot_ref, gamma, mass, k1)):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
                raise CmpcError("solve_device needs contiguous float64 CUDA tensors (%s)" % name)
            if tuple(t.shape) != want[name]:
                raise CmpcError("solve_device: %s has shape %s, expected %s" % (name, tuple(t.shape), want[name]))
            if t.device.index != self.device:
                raise CmpcError("solve_device: %s lives on cuda:%s, the handle on cuda:%d" % (name, t.device.index, self.device))
        if B < 1 or B > self.capacity:
            raise CmpcError("solve_device: batch %d exceeds the handle's capacity %d" % (B, self.capacity))
        if out is not None:
            owant = {"x1": ((B, NX), torch.float64), "u0": ((B, NU), torch.float64), "xN": ((B, NX), torch.float64), "cost": ((B,), torch.float64),
                     "viol": ((B,), torch.float64), "status": ((B,), torch.int32), "iters": ((B,), torch.int32)}
            for name, (shp, dt) in owant.items():
                t = out.get(name)
                if t is None or tuple(t.shape) != shp or t.dtype != dt or not t.is_cuda or not t.is_contiguous() or t.device.index != self.device:
                    raise CmpcError("solve_device: out[%r] must be a contiguous %s CUDA tensor of shape %s on cuda:%d" % (name, dt, shp, self.device))
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


    def reset_warm(self, mask=None):
        """Forget the warm-start state of all instances, or of those flagged in `mask` (bool / uint8 array over the first len(mask) instances)."""
        if mask is None:
            _check(self._L, self._L.cmpc_reset_warm(self._h, None, 0), "cmpc_reset_warm")
        else:
            m = np.ascontiguousarray(np.asarray(mask) != 0, np.uint8)
            _check(self._L, self._L.cmpc_reset_warm(self._h, m.ctypes.data_as(ctypes.c_void_p), int(m.size)), "cmpc_reset_warm")

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
        a, b, d, n = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_int32()
        _check(self._L, self._L.cmpc_footprint(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(n), ctypes.byref(d)), "cmpc_footprint")
        return {"iterate_bytes_per_instance": a.value, "scratch_bytes_per_slot": b.value, "slots": n.value, "smem_bytes_per_cta": d.value}
