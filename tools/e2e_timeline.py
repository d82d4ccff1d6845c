#!/usr/bin/env python3
"""Wall time of every C-ABI call of the end-to-end step of bench.py (host mesh in pinned memory -> CSR + rhs back in pinned
memory), with a stream synchronisation after each call.  Usage: python tools/e2e_timeline.py [n]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
import ffcuda  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ID, DX, DY, DZ = 0, 1, 2, 6
LAP = [(0, DX, 0, DX, 1.0), (0, DY, 0, DY, 1.0), (0, DZ, 0, DZ, 1.0)]
ctx = ffcuda.Context(0)
qp, qw = ffcuda.quadrature(3, 6)
hm = ctx.mesh_cube(n, n, n).download()


def pin(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
    t.numpy()[...] = a
    return t.numpy()


hm = {k: (pin(v) if isinstance(v, np.ndarray) else v) for k, v in hm.items()}
nv = hm["xyz"].shape[0]
h_rp, h_ci, h_val, h_b = pin(np.zeros(nv + 1, np.int32)), None, None, pin(np.zeros(nv))


def T(name, f, acc):
    t0 = time.perf_counter()
    r = f()
    ctx.sync()
    acc.append((name, (time.perf_counter() - t0) * 1e3))
    return r


for rep in range(4):
    acc = []
    t00 = time.perf_counter()
    m2 = T("mesh_upload", lambda: ctx.mesh_upload(3, hm["xyz"], hm["conn"], hm["elab"], hm["bconn"], hm["blab"], hm["belem"], hm["bface"]), acc)
    sp = T("space", lambda: m2.space(1, 1), acc)
    pat = T("symbolic(+incidence)", lambda: sp.symbolic(), acc)
    N, nnz = pat.info()
    if h_ci is None:
        h_ci, h_val = pin(np.zeros(nnz, np.int32)), pin(np.zeros(nnz))
    A = T("matrix", lambda: pat.matrix(), acc)
    T("assemble", lambda: A.assemble(LAP, qp, qw), acc)
    b = T("vec", lambda: ctx.vec(N), acc)
    T("rhs", lambda: sp.assemble_linear(b, [(0, ID, 1.0)], qp, qw), acc)
    bc = T("bc_from_labels", lambda: sp.bc_from_labels([1, 2, 3, 4, 5, 6], 1, [0.0]), acc)
    T("apply_bc", lambda: (A.apply_bc(bc, 1e30), b.apply_bc(bc, 1e30)), acc)
    T("pattern_download", lambda: pat.download(h_rp, h_ci), acc)
    T("matrix_download", lambda: A.download(h_val), acc)
    T("vec_download", lambda: b.download(h_b), acc)
    tot = (time.perf_counter() - t00) * 1e3
    print(f"rep {rep}: total {tot:.2f} ms | " + " ".join(f"{k}={v:.2f}" for k, v in acc), flush=True)
    del m2, sp, pat, A, b, bc
