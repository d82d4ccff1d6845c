#!/usr/bin/env python3
"""Times the numeric assembly kernel of 3-D P1 Poisson / heat on cube(n) for several tile sizes and block sizes.
Usage: python tools/tile_sweep.py [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
import ffcuda  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ID, DX, DY, DZ = 0, 1, 2, 6
LAP = [(0, DX, 0, DX, 1.0), (0, DY, 0, DY, 1.0), (0, DZ, 0, DZ, 1.0)]
HEAT = [(0, ID, 0, ID, 100.0)] + LAP
os.environ["FFCUDA_VERBOSE"] = "1"
ctx = ffcuda.Context(0)
qp, qw = ffcuda.quadrature(3, 6)
mesh = ctx.mesh_cube(n, n, n)


def run(policy, rows, threads, terms, tag):
    os.environ["FFCUDA_TILE_THREADS"] = str(threads)
    ctx.set_option("tile_policy", policy)
    ctx.set_option("tile_rows", rows)
    sp = mesh.space(1, 1)
    pat = sp.symbolic()
    A = pat.matrix()
    for _ in range(3):
        A.assemble(terms, qp, qw)
    ctx.prof_enable(True)
    ctx.prof_reset()
    for _ in range(10):
        A.assemble(terms, qp, qw)
    ms, cnt = ctx.prof_get("asm_rows_p1")
    ctx.prof_enable(False)
    nnz = pat.info()[1]
    print(f"{tag}: policy={policy} rows={rows} threads={threads}: {ms / cnt * 1e3:.1f} us per assembly, {nnz / (ms / cnt * 1e-3) / 1e9:.1f} G nnz/s", flush=True)


run(0, 64, 256, LAP, "poisson thread-per-row")
for rows in (32, 48, 64, 96, 128):
    for threads in (256, 512):
        run(2, rows, threads, LAP, "poisson tiles")
run(0, 64, 256, HEAT, "heat thread-per-row")
for rows, threads in ((64, 512), (96, 256), (96, 512), (48, 256)):
    run(2, rows, threads, HEAT, "heat tiles")
