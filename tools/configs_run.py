#!/usr/bin/env python3
"""Times BASELINE.json configs 1, 3 and 4 on one GPU through the C ABI (bench.py carries config 2 / 5): per phase kernel
times from the library's event profiler and wall-clock per call, one JSON line per config.
Usage: python tools/configs_run.py [1] [3] [4] [--n3 64] [--n4 128] [--steps4 10]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ffcuda  # noqa: E402
import ff_cases as fc  # noqa: E402

ID = 0
NAMES_ASM = ("inc_", "sym_p1_rows", "sym_p1_cols", "sym_block_pattern", "sym_compact_cols", "sym_row_count", "sym_row_fill", "sym_diagpos",
             "sym_", "scan_", "asm_rows", "rhs_rows", "bc_", "vec_fill", "")
NAMES_CG = ("spmv_", "cg_diag_stats", "cg_precond", "cg_init", "cg_spmv_dots", "cg_update_g", "cg_update_xh", "")


def opt(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def wall(ctx, fn):
    ctx.sync()
    t = time.perf_counter()
    r = fn()
    ctx.sync()
    return (time.perf_counter() - t) * 1e3, r


def run(ctx, label, mesh, dim, order, ncomp, terms, rhs, bcs, reps=3, itmax=0):
    qp, qw = ffcuda.quadrature(dim, 6)
    t_space, sp = wall(ctx, lambda: mesh.space(order, ncomp))
    out = {"config": label, "space_ms": t_space}
    for rep in range(reps):
        ctx.prof_enable(True)
        ctx.prof_reset()
        t_sym, pat = wall(ctx, sp.symbolic)
        n, nnz = pat.info()
        A = pat.matrix()
        t_asm, _ = wall(ctx, lambda: A.assemble(terms, qp, qw))
        b = ctx.vec(n)
        t_rhs, _ = wall(ctx, lambda: sp.assemble_linear(b, rhs, qp, qw))
        bc = sp.bc_from_labels(*bcs)
        t_bc, _ = wall(ctx, lambda: (A.apply_bc(bc, 1e30), b.apply_bc(bc, 1e30)))
        prof_asm = {k: ctx.prof_get(k) for k in NAMES_ASM}
        ctx.prof_enable(False)
        x = ctx.vec(n)
        t_cg, (it, conv, gcg) = wall(ctx, lambda: A.cg(b, x, eps=1e-6, itmax=itmax, tgv=1e30))   # as a user runs it (CUDA-graph batches)
        prof_cg = {}
        if rep == reps - 1:   # kernel breakdown: the same solve once more under the event profiler (plain launches)
            ctx.prof_enable(True)
            ctx.prof_reset()
            x2 = ctx.vec(n)
            A.cg(b, x2, eps=1e-6, itmax=itmax, tgv=1e30)
            ctx.sync()
            prof_cg = {k: ctx.prof_get(k) for k in NAMES_CG}
            ctx.prof_enable(False)
            del x2
        if rep < reps - 1:
            del A, pat, b, x, bc
    kern = {k: [round(v[0], 4), int(v[1])] for k, v in prof_asm.items() if v[1]}
    kcg = {k: [round(v[0], 4), int(v[1])] for k, v in prof_cg.items() if v[1]}
    spmv = kcg.get("cg_spmv_dots", [0, 1])
    bytes_spmv = 12.0 * nnz + 20.0 * n
    out.update({"ndof": n, "nnz": nnz, "symbolic_ms": t_sym, "assembly_ms": t_asm, "rhs_ms": t_rhs, "bc_ms": t_bc,
                "assembly_nnz_per_s": nnz / (t_asm * 1e-3), "cg_iters": it, "cg_converged": conv, "cg_ms": t_cg,
                "cg_ms_per_iter": t_cg / max(it, 1), "spmv_ms": spmv[0] / max(spmv[1], 1),
                "spmv_gbs": bytes_spmv / (spmv[0] / max(spmv[1], 1) * 1e-3) / 1e9 if spmv[0] else None,
                "kernels_asm": kern, "kernels_cg": kcg})
    u = x.download()
    out["u_norm2"] = float(np.dot(u, u))
    print(json.dumps(out), flush=True)
    return out


def main():
    which = [a for a in sys.argv[1:] if a in ("1", "3", "4")] or ["1", "3", "4"]
    ctx = ffcuda.Context(0)
    if "1" in which:
        m = ctx.mesh_square(1000, 1000)
        run(ctx, "config1: 2-D P1 Laplace square(1000,1000)", m, 2, 1, 1, fc.LAP2, [(0, ID, 1.0)], ([1, 2, 3, 4], 1, [0.0]))
        del m
    if "3" in which:
        n3 = opt("--n3", 64)
        m = ctx.mesh_cube(n3, n3, n3)
        run(ctx, f"config3: 3-D [P2,P2,P2] Lame cube({n3})", m, 3, 2, 3, fc.lame_terms(), [(2, ID, -0.05)],
            ([1], 7, [0.0, 0.0, 0.0]), reps=2, itmax=opt("--it3", 200))
        del m
    if "4" in which:
        n4, steps = opt("--n4", 128), opt("--steps4", 10)
        m = ctx.mesh_cube(n4, n4, n4)
        sp = m.space(1, 1)
        qp, qw = ffcuda.quadrature(3, 6)
        dt = 0.01
        heat = [(0, ID, 0, ID, 1.0 / dt)] + fc.LAP3
        massf = [(0, ID, 0, ID, 1.0 / dt)]
        pat = sp.symbolic()
        n, nnz = pat.info()
        bc = sp.bc_from_labels(fc.ALL6, 1, [0.0])
        # Heat3d.idp shape: M = mass/dt and the load vector once; every step re-assembles A = M + K (symbolic + numeric +
        # Dirichlet), forms b = M u_old + f on the device (SpMV) and solves A u = b by CG started from u_old
        M = sp.symbolic().matrix()
        M.assemble(massf, qp, qw)
        f = ctx.vec(n)
        sp.assemble_linear(f, [(0, ID, 1.0)], qp, qw)
        hf = f.download()
        u = ctx.vec(n)
        u.fill(0.0)
        b = ctx.vec(n)
        t_asm, t_rhs, t_cg, iters = [], [], [], []
        for s in range(steps):
            ta, (pat, A) = wall(ctx, lambda: (lambda p: (p, p.matrix()))(sp.symbolic()))
            ta2, _ = wall(ctx, lambda: (A.assemble(heat, qp, qw), A.apply_bc(bc, 1e30)))
            tr, _ = wall(ctx, lambda: M.spmv(u, b))
            b.upload(b.download() + hf)  # host add of the load vector: not timed (FreeFEM does it on its own arrays)
            b.apply_bc(bc, 1e30)
            tc, (it, conv, _) = wall(ctx, lambda: A.cg(b, u, eps=1e-6, itmax=0, tgv=1e30))
            t_asm.append(round(ta + ta2, 4)); t_rhs.append(round(tr, 4)); t_cg.append(round(tc, 3)); iters.append(it)
        hu = u.download()
        print(json.dumps({"config": f"config4: 3-D P1 heat cube({n4}), {steps} time steps dt={dt}, A = M/dt + K re-assembled every step",
                          "ndof": n, "nnz": nnz, "reassembly_ms": t_asm, "rhs_spmv_ms": t_rhs, "cg_ms": t_cg, "cg_iters": iters,
                          "u_max": float(hu.max())}), flush=True)

if __name__ == "__main__":
    main()
