#!/usr/bin/env python3
"""One-off set-up costs on cube(n): incidence, tile set.  Usage: python tools/build_time.py [n]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
import ffcuda  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ID, DX, DY, DZ = 0, 1, 2, 6
LAP = [(0, DX, 0, DX, 1.0), (0, DY, 0, DY, 1.0), (0, DZ, 0, DZ, 1.0)]
ctx = ffcuda.Context(0)
qp, qw = ffcuda.quadrature(3, 6)
mesh = ctx.mesh_cube(n, n, n)
for rep in range(2):
    ctx.set_option("tile_policy", 2)
    sp = mesh.space(1, 1)
    ctx.prof_enable(True)
    ctx.prof_reset()
    ctx.sync()
    t0 = time.perf_counter()
    pat = sp.symbolic()
    ctx.sync()
    t1 = time.perf_counter()
    A = pat.matrix()
    A.assemble(LAP, qp, qw)
    ctx.sync()
    t2 = time.perf_counter()
    print(f"rep {rep}: first symbolic (incidence build included) {1e3 * (t1 - t0):.2f} ms wall, first assembly (tile build included) "
          f"{1e3 * (t2 - t1):.2f} ms wall; kernels: inc_ {ctx.prof_get('inc_')[0]:.2f} ms, tile_ {ctx.prof_get('tile_')[0]:.2f} ms "
          f"(sizes {ctx.prof_get('tile_sizes')[0]:.2f}, build {ctx.prof_get('tile_build')[0]:.2f})", flush=True)
    ctx.prof_enable(False)
