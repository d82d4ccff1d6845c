#!/usr/bin/env python3
"""Times the rows added after the headline path (round 1e) on one GPU through the C ABI: GMRES on a convection-diffusion
problem, boundary integrals (Neumann + Robin) and a right-hand side with data at the quadrature nodes, on cube(N).
One JSON line each; kernel times from the library's event profiler.  Usage: python tools/widen_run.py [N=96]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ffcuda  # noqa: E402

ID, DX, DY, DZ = 0, 1, 2, 6
LAP3 = [(0, DX, 0, DX, 1.0), (0, DY, 0, DY, 1.0), (0, DZ, 0, DZ, 1.0)]


def wall(ctx, fn):
    ctx.sync()
    t = time.perf_counter()
    r = fn()
    ctx.sync()
    return (time.perf_counter() - t) * 1e3, r


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    ctx = ffcuda.Context(0)
    qp, qw = ffcuda.quadrature(3, 6)
    fq3, fw3 = ffcuda.quadrature(2, 6)          # 7-point face rule (qf5pT), the default of int2d on a mesh3
    mesh = ctx.mesh_cube(N, N, N)
    sp = mesh.space(1, 1)
    pat = sp.symbolic()
    n, nnz = pat.info()
    nt = 6 * N ** 3
    # --- GMRES: convection-diffusion, Dirichlet on the six faces
    A = pat.matrix()
    A.assemble(LAP3 + [(0, DX, 0, ID, 20.0), (0, DY, 0, ID, -10.0), (0, ID, 0, ID, 1.0)], qp, qw)
    b = ctx.vec(n)
    sp.assemble_linear(b, [(0, ID, 1.0)], qp, qw)
    bc = sp.bc_from_labels([1, 2, 3, 4, 5, 6], 1, [0.0])
    A.apply_bc(bc, 1e30)
    b.apply_bc(bc, 1e30)
    for restart, coop in ((1000, 1), (50, 1), (50, 0)):
        ctx.set_option("gmres_coop", coop)
        x = ctx.vec(n)
        A.gmres(b, x, eps=1e-6, restart=restart)    # warm-up (SELL copy, allocator)
        x = ctx.vec(n)
        ctx.prof_enable(True)
        ctx.prof_reset()
        l0 = ctx.launch_count()
        t, (it, conv, rel) = wall(ctx, lambda: A.gmres(b, x, eps=1e-6, restart=restart))
        kern = {k: ctx.prof_get(k) for k in ("gmres_arnoldi", "gmres_mgs", "gmres_mgs_last", "gmres_scale", "gmres_precond_apply", "spmv", "gmres_update_x", "")}
        ctx.prof_enable(False)
        print(json.dumps({"what": "gmres convection-diffusion", "mesh": f"cube({N})", "n": n, "nnz": nnz, "restart": restart, "cooperative_arnoldi": coop, "iters": it,
                          "converged": conv, "relres": rel, "wall_ms": round(t, 2), "launches": int(ctx.launch_count() - l0),
                          "kernels_ms_and_launches": {k: [round(v[0], 3), int(v[1])] for k, v in kern.items()}}), flush=True)
    # --- boundary integrals: Neumann data on two faces, Robin term on two faces
    ctx.prof_enable(True)
    b2 = ctx.vec(n)
    A2 = pat.matrix()
    A2.assemble(LAP3, qp, qw)
    sp.assemble_linear_boundary(b2, [(0, ID, 2.5)], fq3, fw3, [2, 3], accumulate=False)   # first call builds the boundary incidence
    A2.assemble_boundary([(0, ID, 0, ID, 1.5)], fq3, fw3, [2, 3], accumulate=True)
    ctx.prof_reset()
    t1, _ = wall(ctx, lambda: sp.assemble_linear_boundary(b2, [(0, ID, 2.5)], fq3, fw3, [2, 3], accumulate=False))
    t2, _ = wall(ctx, lambda: A2.assemble_boundary([(0, ID, 0, ID, 1.5)], fq3, fw3, [2, 3], accumulate=True))
    kern = {k: ctx.prof_get(k) for k in ("bnd_measure", "bnd_gather", "bnd_bilinear")}
    print(json.dumps({"what": "boundary integrals (2 of 6 faces)", "mesh": f"cube({N})", "boundary_elements": 12 * N * N,
                      "linear_wall_ms": round(t1, 3), "bilinear_wall_ms": round(t2, 3),
                      "kernels_ms_and_launches": {k: [round(v[0], 4), int(v[1])] for k, v in kern.items()}}), flush=True)
    # --- right-hand side with data at the quadrature nodes (the table is what the plugin evaluates on the host)
    fq = np.random.default_rng(0).random((1, nt, len(qw)))
    b3 = ctx.vec(n)
    sp.assemble_linear_qvalues(b3, qp, qw, fq)
    ctx.prof_reset()
    t3, _ = wall(ctx, lambda: sp.assemble_linear_qvalues(b3, qp, qw, fq))
    kern = {k: ctx.prof_get(k) for k in ("rhs_qvalues_elem", "rhs_qvalues_gather")}
    print(json.dumps({"what": "rhs with data at the quadrature nodes", "mesh": f"cube({N})", "nt": nt, "table_MB": round(fq.nbytes / 1e6, 1),
                      "wall_ms_with_upload_from_pageable_memory": round(t3, 2),
                      "kernels_ms_and_launches": {k: [round(v[0], 4), int(v[1])] for k, v in kern.items()}}), flush=True)


if __name__ == "__main__":
    main()
