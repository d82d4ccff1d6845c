#!/usr/bin/env python3
"""ncu driver: a few tile assemblies of 3-D P1 Poisson on cube(n).  Usage: python tools/prof_tiles.py [n] [rows]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
import ffcuda  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ID, DX, DY, DZ = 0, 1, 2, 6
LAP = [(0, DX, 0, DX, 1.0), (0, DY, 0, DY, 1.0), (0, DZ, 0, DZ, 1.0)]
ctx = ffcuda.Context(0)
ctx.set_option("tile_policy", 2)
ctx.set_option("tile_rows", rows)
qp, qw = ffcuda.quadrature(3, 6)
mesh = ctx.mesh_cube(n, n, n)
sp = mesh.space(1, 1)
pat = sp.symbolic()
A = pat.matrix()
for _ in range(3):
    A.assemble(LAP, qp, qw)
ctx.sync()
print("done", pat.info())
