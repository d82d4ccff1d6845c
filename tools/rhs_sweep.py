#!/usr/bin/env python3
"""Times the P1 right-hand side on cube(n): thread-per-row kernel against the tile kernel.  Usage: python tools/rhs_sweep.py [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
import ffcuda  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ID, DX, DY, DZ = 0, 1, 2, 6
LAP = [(0, DX, 0, DX, 1.0), (0, DY, 0, DY, 1.0), (0, DZ, 0, DZ, 1.0)]
ctx = ffcuda.Context(0)
qp, qw = ffcuda.quadrature(3, 6)
mesh = ctx.mesh_cube(n, n, n)
for policy, threads, lt, tag in [(0, 256, [(0, ID, 1.0)], "rows f"), (2, 128, [(0, ID, 1.0)], "tiles f"), (2, 256, [(0, ID, 1.0)], "tiles f"),
                                 (2, 512, [(0, ID, 1.0)], "tiles f"), (0, 256, [(0, ID, 1.0), (0, DX, 1.0)], "rows f+dx"),
                                 (2, 256, [(0, ID, 1.0), (0, DX, 1.0)], "tiles f+dx")]:
    os.environ["FFCUDA_RHS_THREADS"] = str(threads)
    ctx.set_option("tile_policy", policy)
    sp = mesh.space(1, 1)
    pat = sp.symbolic()
    A = pat.matrix()
    A.assemble(LAP, qp, qw)
    b = ctx.vec(pat.info()[0])
    for _ in range(3):
        sp.assemble_linear(b, lt, qp, qw)
    ctx.prof_enable(True)
    ctx.prof_reset()
    for _ in range(10):
        sp.assemble_linear(b, lt, qp, qw)
    ms, cnt = ctx.prof_get("rhs_rows")
    ctx.prof_enable(False)
    print(f"{tag}: policy={policy} threads={threads}: {ms / cnt * 1e3:.1f} us", flush=True)
