"""B200-native batched solver for the per-tick centroidal NMPC of the reference
(`code/centroidal_mpc_vertices.py`, `code/centroidal_mpc_vertices_payload.py`).

Layout
  csrc/            CUDA kernels + C ABI (include/cmpc.h) -> libcmpc_b200.so (built in-tree)
  _lib.py          ctypes binding of the C ABI (`BatchSolver`)
  assembly.py      per-tick parameter assembly of `solve` (MPC file :482-600), vectorised
  fleet.py         batched closed loop on the device: table-gather assembly, solve, plant, step adjustment
  idqp.py          batched whole-body inverse-dynamics QP (the consumer of the MPC's output; `utils.QPSolver` surface)
  com_reference.py CoM reference tables of `functions.references` without CasADi (minimum-norm quintic spline)
  parallel.py      instance sharding over GPUs + statistics reduction (no collective on the hot path)
  centroidal_mpc_vertices.py / centroidal_mpc_vertices_payload.py
                   drop-in modules with the reference's `centroidal_mpc` class surface

There is no CPU path: creating a solver without the CUDA library or without a GPU raises.
"""
from ._lib import BatchSolver, CmpcError, WalkTables, build_library, library_path, measure_fp64_peak  # noqa: F401
from .assembly import PlanTables, assemble_tick, pack_instances  # noqa: F401
from .parallel import gather_stats, shard_arrays, shard_range  # noqa: F401
from .fleet import Fleet  # noqa: F401
from .idqp import QPSolver, assemble_id_qp, joint_torques  # noqa: F401
from .com_reference import compute_knot, quintic_coefficients, references, references_from_knots, sample_tables  # noqa: F401

__all__ = ["BatchSolver", "CmpcError", "WalkTables", "build_library", "library_path", "measure_fp64_peak",
           "PlanTables", "assemble_tick", "pack_instances", "gather_stats", "shard_arrays", "shard_range", "Fleet",
           "quintic_coefficients", "references_from_knots", "sample_tables", "compute_knot", "references", "QPSolver", "assemble_id_qp", "joint_torques"]
