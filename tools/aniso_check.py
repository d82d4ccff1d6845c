import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, ffcuda, ff_cases as fc
ctx = ffcuda.Context(0); ctx.set_option("tile_policy", 2)
qp, qw = ffcuda.quadrature(3, 6)
for dims in [(128,128,128),(128,128,256),(128,256,256),(256,256,64)]:
    mesh = ctx.mesh_cube(*dims); sp = mesh.space(1,1); pat = sp.symbolic(); A = pat.matrix()
    A.assemble(fc.LAP3, qp, qw); ctx.sync()
    ctx.prof_enable(True); ctx.prof_reset()
    for _ in range(5): A.assemble(fc.LAP3, qp, qw)
    b = ctx.vec(pat.info()[0])
    for _ in range(5): sp.assemble_linear(b, [(0,0,1.0)], qp, qw)
    ctx.sync()
    ms, cnt = ctx.prof_get("asm_rows_p1"); ms2, cnt2 = ctx.prof_get("rhs_rows")
    nt = mesh.info()[2]
    print(dims, f"asm {ms/cnt:.4f} ms = {ms/cnt/nt*12582912:.4f} ms per 12.58M tets; rhs {ms2/cnt2:.4f}", flush=True)
    ctx.prof_enable(False)
    del A, pat, sp, mesh, b
