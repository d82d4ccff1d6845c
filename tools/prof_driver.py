#!/usr/bin/env python3
"""Small driver for ncu captures: one pass of the hot path on cube(n) with a CG cut to a few iterations, so that a
`ncu --set full` run (about 40 replays per kernel) stays short.  Usage: python tools/prof_driver.py [n] [cg_iters] [order] [ncomp]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
import ffcuda  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
order = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ncomp = int(sys.argv[4]) if len(sys.argv) > 4 else 1
ID, DX, DY, DZ = 0, 1, 2, 6
ctx = ffcuda.Context(0)
qp, qw = ffcuda.quadrature(3, 6)
mesh = ctx.mesh_cube(n, n, n)
sp = mesh.space(order, ncomp)
if ncomp == 1:
    terms = [(0, DX, 0, DX, 1.0), (0, DY, 0, DY, 1.0), (0, DZ, 0, DZ, 1.0)]
    rhs = [(0, ID, 1.0)]
    bcs = ([1, 2, 3, 4, 5, 6], 1, [0.0])
else:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ff_cases as fc

    terms, rhs, bcs = fc.lame_terms(), [(2, ID, -0.05)], ([1], 7, [0.0, 0.0, 0.0])
for rep in range(2):
    pat = sp.symbolic()
    A = pat.matrix()
    A.assemble(terms, qp, qw)
    N = pat.info()[0]
    b = ctx.vec(N)
    sp.assemble_linear(b, rhs, qp, qw)
    bc = sp.bc_from_labels(*bcs)
    A.apply_bc(bc, 1e30)
    b.apply_bc(bc, 1e30)
    x = ctx.vec(N)
    it, conv, g = A.cg(b, x, eps=1e-6, itmax=iters, tgv=1e30)
ctx.sync()
print("done", pat.info(), it, conv)
