#!/usr/bin/env python3
"""ncu driver for the shipped configuration: three passes of the P1 hot path on cube(n) (the third one runs the fused
symbolic kernel, the tile assembly and the tile right-hand side) with a CG cut to a few iterations, then one pass of
[P2,P2,P2] Lame on cube(n/4).  Usage: python tools/prof_final.py [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ffcuda  # noqa: E402
import ff_cases as fc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ID = 0
ctx = ffcuda.Context(0)
qp, qw = ffcuda.quadrature(3, 6)
mesh = ctx.mesh_cube(n, n, n)
sp = mesh.space(1, 1)
for rep in range(3):
    pat = sp.symbolic()
    A = pat.matrix()
    A.assemble(fc.LAP3, qp, qw)
    N = pat.info()[0]
    b = ctx.vec(N)
    sp.assemble_linear(b, [(0, ID, 1.0)], qp, qw)
    bc = sp.bc_from_labels(fc.ALL6, 1, [0.0])
    A.apply_bc(bc, 1e30)
    b.apply_bc(bc, 1e30)
    x = ctx.vec(N)
    it, conv, g = A.cg(b, x, eps=1e-6, itmax=3, tgv=1e30)
del A, pat, b, x, sp, mesh
m2 = ctx.mesh_cube(n // 4, n // 4, n // 4)
sp2 = m2.space(2, 3)
pat = sp2.symbolic()
A = pat.matrix()
A.assemble(fc.lame_terms(), qp, qw)
N = pat.info()[0]
b = ctx.vec(N)
sp2.assemble_linear(b, [(2, ID, -0.05)], qp, qw)
bc = sp2.bc_from_labels([1], 7, [0.0, 0.0, 0.0])
A.apply_bc(bc, 1e30)
b.apply_bc(bc, 1e30)
x = ctx.vec(N)
A.cg(b, x, eps=1e-6, itmax=3, tgv=1e30)
ctx.sync()
print("done", pat.info())
