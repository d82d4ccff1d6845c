import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, ffcuda, ff_cases as fc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = ffcuda.Context(0)
qp, qw = ffcuda.quadrature(3, 6)
mesh = ctx.mesh_cube(n, n, n); sp = mesh.space(2, 3); pat = sp.symbolic(); A = pat.matrix()
N, nnz = pat.info()
res = {}
for g in ("0", "1"):
    os.environ["FFCUDA_P2_PREGEOM"] = g
    A.assemble(fc.lame_terms(), qp, qw); ctx.sync()
    ctx.prof_enable(True); ctx.prof_reset()
    for _ in range(3): A.assemble(fc.lame_terms(), qp, qw)
    ctx.sync()
    ms, cnt = ctx.prof_get("asm_rows_p2"); mg, cg = ctx.prof_get("asm_p2_geom")
    ctx.prof_enable(False)
    res[g] = A.download().copy() if n <= 32 else None
    print(f"pregeom={g}: asm_rows_p2 {ms/3:.3f} ms per assembly ({cnt//3} launches) + geom {mg/3:.3f} ms; nnz {nnz} -> {nnz/((ms+mg)/3*1e-3)/1e9:.1f} G nnz/s", flush=True)
if res["0"] is not None:
    print("max rel diff", np.max(np.abs(res["0"]-res["1"]))/np.abs(res["0"]).max())
