#!/usr/bin/env python3
"""Bisect of the step-time instability: variants of the assembly loop (keep previous generation alive / torch stream)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
import ffcuda
mode = sys.argv[1]
n = 128
ID, DX, DY, DZ = 0, 1, 2, 6
LAP = [(0, DX, 0, DX, 1.0), (0, DY, 0, DY, 1.0), (0, DZ, 0, DZ, 1.0)]
if "torch" in mode:
    import torch
    torch.cuda.set_device(0)
ctx = ffcuda.Context(0)
if "tstream" in mode:
    st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
qp, qw = ffcuda.quadrature(3, 6)
mesh = ctx.mesh_cube(n, n, n)
sp = mesh.space(1, 1)
def step():
    pat = sp.symbolic(); A = pat.matrix(); A.assemble(LAP, qp, qw)
    b = ctx.vec(pat.info()[0]); sp.assemble_linear(b, [(0, ID, 1.0)], qp, qw)
    bc = sp.bc_from_labels([1, 2, 3, 4, 5, 6], 1, [0.0]); A.apply_bc(bc, 1e30); b.apply_bc(bc, 1e30)
    return pat, A, b
out = None
ts = []
for rep in range(12):
    ctx.sync(); t0 = time.perf_counter()
    if "keep" in mode:
        out = step()
    else:
        out = None
        out = step()
    ctx.sync(); ts.append((time.perf_counter() - t0) * 1e3)
print(mode, " ".join(f"{t:.1f}" for t in ts), flush=True)
os._exit(0)
