#!/usr/bin/env python3
"""Quick check + timing of the fan tile kernel: parity vs the oracle on small cubes, timing on cube(n).
usage: python tools/fan_check.py [n=128] [small_only]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import ffcuda  # noqa: E402
import ff_cases as fc  # noqa: E402
import oracle_lib as ol  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ctx = ffcuda.Context(0)
ctx.set_option("tile_policy", 2)
qp, qw = ffcuda.quadrature(3, 6)
for size in [(1, 1, 1), (2, 3, 1), (5, 4, 6), (11, 9, 13), (24, 24, 24)]:
    for rows in (96, 32, 256):
        ctx.set_option("tile_rows", rows)
        m = ol.cube(*size)
        N = m["xyz"].shape[0]
        ci, cj, ca = ol.assemble_coo(m, 1, 1, None, fc.LAP3, qp, qw)
        orp, ocol, oval = ol.coo_to_csr(N, ci, cj, ca)
        mesh = ctx.mesh_cube(*size)
        sp = mesh.space(1, 1)
        pat = sp.symbolic()
        A = pat.matrix()
        A.assemble(fc.LAP3, qp, qw)
        val = A.download()
        rp, col = pat.download()
        assert np.array_equal(rp, orp) and np.array_equal(col, ocol)
        err = np.max(np.abs(val - oval)) / np.abs(oval).max()
        A.assemble(fc.LAP3, qp, qw)
        same = np.array_equal(A.download(), val)
        heat = [(0, fc.ID, 0, fc.ID, 100.0)] + fc.LAP3
        hi, hj, ha = ol.assemble_coo(m, 1, 1, None, heat, qp, qw)
        _, _, hval = ol.coo_to_csr(N, hi, hj, ha)
        A.assemble(heat, qp, qw)
        hv = A.download()
        errh = np.max(np.abs(hv - hval)) / np.abs(hval).max()
        A.assemble(heat, qp, qw)
        assert errh <= 1e-12 and np.array_equal(A.download(), hv), ("heat form", errh)
        b = ctx.vec(N)
        sp.assemble_linear(b, [(0, fc.ID, 2.5)], qp, qw)
        ob = ol.assemble_rhs(m, 1, 1, None, N, [(0, fc.ID, 2.5)], qp, qw)
        hb = b.download()
        errb = np.max(np.abs(hb - ob)) / np.abs(ob).max()
        sp.assemble_linear(b, [(0, fc.ID, 2.5)], qp, qw)
        sameb = np.array_equal(b.download(), hb)
        print(f"cube{size} rows={rows}: rel err {err:.2e} reproducible={same}; rhs rel err {errb:.2e} reproducible={sameb}", flush=True)
        assert err <= 1e-12 and same and errb <= 1e-12 and sameb
if len(sys.argv) > 2:
    sys.exit(0)
ctx.set_option("tile_rows", int(os.environ.get("ROWS", "96")))
mesh = ctx.mesh_cube(n, n, n)
sp = mesh.space(1, 1)
pat = sp.symbolic()
A = pat.matrix()
t0 = time.time()
A.assemble(fc.LAP3, qp, qw)
ctx.sync()
print(f"first assembly incl. tile+fan build: {time.time() - t0:.3f} s", flush=True)
ctx.prof_enable(True)
ctx.prof_reset()
for _ in range(10):
    A.assemble(fc.LAP3, qp, qw)
ctx.sync()
ms, cnt = ctx.prof_get("asm_rows_p1")
nv, nt = mesh.info()[1], mesh.info()[2]
N, nnz = pat.info()
B = 16.0 * nt + 24.0 * nv + 12.0 * nnz + 4.0 * (N + 1)
print(f"asm_rows_p1: {ms / cnt:.4f} ms per launch, {B / (ms / cnt * 1e-3) / 1e9:.0f} GB/s algorithmic, frac {B / (ms / cnt * 1e-3) / 1e9 / 6454.3:.3f}")
ctx.prof_reset()
for _ in range(10):
    patx = sp.symbolic()
ctx.sync()
for nm in ("sym_p1_fused", "sym_p1_rows", "sym_p1_cols", "sym_", "scan_"):
    ms, cnt = ctx.prof_get(nm)
    print(f"{nm}: {ms / max(cnt, 1):.4f} ms per launch ({cnt} launches)")
rpx, colx = patx.download()
rp0, col0 = pat.download()
assert np.array_equal(rpx, rp0) and np.array_equal(colx, col0)
b = ctx.vec(N)
ctx.prof_reset()
for _ in range(10):
    sp.assemble_linear(b, [(0, fc.ID, 1.0)], qp, qw)
ctx.sync()
ms, cnt = ctx.prof_get("rhs_rows")
print(f"rhs_rows: {ms / cnt:.4f} ms per launch; sum b = {b.download().sum():.15f}")
import scipy.sparse as sps  # noqa: E402

rp, col = pat.download()
M = sps.csr_matrix((A.download(), col, rp), shape=(N, N))
print("row sums", np.max(np.abs(M @ np.ones(N))), "sym", abs(M - M.T).max())
heat = [(0, fc.ID, 0, fc.ID, 100.0)] + fc.LAP3
AH = pat.matrix()
AH.assemble(heat, qp, qw)
ctx.prof_reset()
for _ in range(5):
    AH.assemble(heat, qp, qw)
ctx.sync()
ms, cnt = ctx.prof_get("asm_rows_p1")
print(f"heat form on the fans: {ms / cnt:.4f} ms per launch")
hv = AH.download()
ctx.set_option("tile_fans", 0)
ctx.prof_reset()
for _ in range(3):
    AH.assemble(heat, qp, qw)
ctx.sync()
ms, cnt = ctx.prof_get("asm_rows_p1")
print(f"heat form, round-1 tile kernel: {ms / cnt:.4f} ms; max rel diff {np.max(np.abs(AH.download() - hv)) / np.abs(hv).max():.2e}")
ctx.prof_reset()
A2 = pat.matrix()
for _ in range(5):
    A2.assemble(fc.LAP3, qp, qw)
ctx.sync()
ms, cnt = ctx.prof_get("asm_rows_p1")
print(f"round-1 tile kernel: {ms / cnt:.4f} ms; max diff vs fans {np.max(np.abs(A2.download() - A.download())):.3e}")
