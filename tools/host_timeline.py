#!/usr/bin/env python3
"""Host-side timeline of one assembly step: wall time of every C-ABI call (with a stream sync after each)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
import ffcuda
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
sync_each = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ID, DX, DY, DZ = 0, 1, 2, 6
LAP = [(0, DX, 0, DX, 1.0), (0, DY, 0, DY, 1.0), (0, DZ, 0, DZ, 1.0)]
ctx = ffcuda.Context(0)
qp, qw = ffcuda.quadrature(3, 6)
mesh = ctx.mesh_cube(n, n, n)
sp = mesh.space(1, 1)
def T(name, f, acc):
    t0 = time.perf_counter(); r = f()
    if sync_each: ctx.sync()
    acc.append((name, (time.perf_counter() - t0) * 1e3)); return r
for rep in range(6):
    acc = []
    t00 = time.perf_counter()
    pat = T("symbolic", lambda: sp.symbolic(), acc)
    A = T("matrix_create", lambda: pat.matrix(), acc)
    T("assemble", lambda: A.assemble(LAP, qp, qw), acc)
    N = pat.info()[0]
    b = T("vec_create", lambda: ctx.vec(N), acc)
    T("rhs", lambda: sp.assemble_linear(b, [(0, ID, 1.0)], qp, qw), acc)
    bc = T("bc_from_labels", lambda: sp.bc_from_labels([1, 2, 3, 4, 5, 6], 1, [0.0]), acc)
    T("apply_bc_A", lambda: A.apply_bc(bc, 1e30), acc)
    T("apply_bc_b", lambda: b.apply_bc(bc, 1e30), acc)
    ctx.sync()
    tot = (time.perf_counter() - t00) * 1e3
    print(f"rep {rep}: total {tot:.2f} ms | " + " ".join(f"{k}={v:.2f}" for k, v in acc), flush=True)
    t0 = time.perf_counter(); del pat, A, b, bc; ctx.sync(); print(f"   free: {(time.perf_counter()-t0)*1e3:.2f} ms")
