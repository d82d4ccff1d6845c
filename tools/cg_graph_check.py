import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, ffcuda, ff_cases as fc
import torch
ctx = ffcuda.Context(0)
for kind, dims in (("square", (1000, 1000)), ("cube", (128, 128, 128))):
    dim = len(dims)
    qp, qw = ffcuda.quadrature(dim, 6)
    mesh = ctx.mesh_square(*dims) if kind == "square" else ctx.mesh_cube(*dims)
    sp = mesh.space(1, 1); pat = sp.symbolic(); A = pat.matrix()
    A.assemble(fc.LAP2 if dim == 2 else fc.LAP3, qp, qw)
    n = pat.info()[0]
    b = ctx.vec(n); sp.assemble_linear(b, [(0, 0, 1.0)], qp, qw)
    bc = sp.bc_from_labels([1, 2, 3, 4] if dim == 2 else [1, 2, 3, 4, 5, 6], 1, [0.0])
    A.apply_bc(bc, 1e30); b.apply_bc(bc, 1e30)
    x = ctx.vec(n)
    res = {}
    for g in ("0", "1"):
        os.environ["FFCUDA_CG_GRAPH"] = g
        for rep in range(3):
            x.fill(0.0); ctx.sync(); t0 = time.perf_counter()
            it, conv, gcg = A.cg(b, x, eps=1e-6, itmax=0, tgv=1e30)
            ctx.sync(); dt = time.perf_counter() - t0
        res[g] = (it, conv, dt, x.download().copy())
        print(f"{kind}{dims} graph={g}: iters {it} conv {conv} {dt*1e3:.2f} ms = {dt*1e3/it:.4f} ms/it", flush=True)
    print("   same iterate:", np.array_equal(res["0"][3], res["1"][3]), "same count:", res["0"][0] == res["1"][0])
    del A, pat, sp, mesh, b, x
