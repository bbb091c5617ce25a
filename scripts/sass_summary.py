"""Opcode histogram of the kernels in lib/libcmpc_b200.so (cuobjdump -sass) -> profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lib", "libcmpc_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEY = ['DFMA', 'DADD', 'DMUL', 'MUFU.RSQ', 'MUFU.RCP64H', 'BAR.SYNC', 'BAR.RED', 'SHFL.BFLY', 'SHFL.IDX', 'VOTE.ANY', 'VOTE.ALL', 'LDS', 'LDS.64', 'LDS.128',
       'STS', 'STS.64', 'STS.128', 'LDG.E', 'STG.E', 'LD.E', 'ST.E', 'LDL', 'LDL.64', 'STL', 'STL.64', 'LDGSTS.E', 'ATOMG.E', 'ATOMS', 'RED.E', 'CCTL', 'DMMA', 'UBLKCP', 'UTMALDG']
KEEP2 = ('LDG', 'STG', 'LDS', 'STS', 'LD.', 'ST.', 'LDL', 'STL', 'BAR', 'SHFL', 'VOTE', 'DFMA', 'DADD', 'DMUL', 'LDGSTS', 'ATOM', 'RED', 'MUFU', 'CCTL')
out = ["SASS opcode histogram of %s (cuobjdump -sass, sm_100a), per kernel; produced by scripts/sass_summary.py\n" % os.path.relpath(so, ROOT)]
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n', 1)[0].strip()
    ops = collections.Counter()
    for line in f.split('\n'):
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            op = m.group(1)
            ops['.'.join(op.split('.')[:2]) if op.startswith(KEEP2) else op.split('.')[0]] += 1
    agg = collections.Counter()
    for k, v in ops.items():
        for kk in KEY:
            if k == kk or k.startswith(kk + '.'):
                agg[kk] += v
    short = re.sub(r'^_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_cmpc_kernels_cu_[0-9a-f]+', '', name)
    out.append("== %s  (%d instructions)" % (short[:110], sum(ops.values())))
    out.append("   " + ", ".join("%s %d" % (k, agg[k]) for k in KEY if agg[k]))
    out.append("   top: " + ", ".join("%s %d" % (k, v) for k, v in ops.most_common(14)) + "\n")
open(os.path.join(ROOT, "profiles", "r02_sass_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[-8:]))
