cd "$GRAFT_REPO_ROOT"
# per-phase cycles with 1 / 2 / 3 / 4 resident CTAs per SM (shared-memory padding): which phases inflate when CTAs share an SM?
for pad in 120000 60000 20000 0; do
  echo "== pad $pad"
  CMPC_SMEM_PAD=$pad CMPC_LIB=$PWD/lib/variants/lib_prof_head.so timeout 200 python scripts/phase_profile.py 20 4096 | tail -1 | tee gpurun_out/phase_pad$pad.json | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print(j['kernel_ms'], j['nfact'], j['cta_cycles_mean'], j['cycles_per_fact'])"
done
