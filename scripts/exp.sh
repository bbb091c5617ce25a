cd "$GRAFT_REPO_ROOT"
for so in lib/variants/lib_prev.so lib/libcmpc_b200.so; do
CMPC_LIB=$PWD/$so python - <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import bench, cmpc_loader
pkg = cmpc_loader.load()
N, B = 20, 4096
(prev2, prev, cur), mass, k1, idx = bench.replay_workload(N, B, seed=0, back=2)
s = pkg.BatchSolver(N, B, device=0)
s.solve_host(*prev2, mass, k1, 0)
o1 = s.solve_host(*prev, mass, k1, 4)
s.warm_save(B)
for rep in range(2):
    s.warm_restore(B)
    o2 = s.solve_host(*cur, mass, k1, 4)
    st = s.last_stats()
    it = o2["iters"]
    big = np.argsort(-it)[:5]
    print(os.environ["CMPC_LIB"][-14:], "kernel %.2f ms  iters mean %.3f max %d nfact %.3f  top:" % (st["kernel_ms"], it.mean(), it.max(), st["nfact"] / B), [(int(idx[b]), int(it[b])) for b in big], flush=True)
PY
done
