#!/usr/bin/env python3
"""Multi-GPU parity check, launched with torchrun (one process per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dist_check.py
Every rank assembles and solves its slab of 3-D P1 Poisson on cube(nx,ny,nz); the owned rows are gathered on rank 0 with
global column ids and compared with the CPU oracle on the whole mesh: pattern bit-exact, values / rhs 1e-12, CG
iteration count equal and solution 1e-12 (eps = 1e-6) — the same bars as on one GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ffcuda  # noqa: E402

ID, DX, DY, DZ = 0, 1, 2, 6
LAP = [(0, DX, 0, DX, 1.0), (0, DY, 0, DY, 1.0), (0, DZ, 0, DZ, 1.0)]
RHS = [(0, ID, 1.0)]
ALL6 = [1, 2, 3, 4, 5, 6]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ffcuda.Context(local)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.tensor(list(ffcuda.Context.comm_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    ctx.comm_init(rank, world, bytes(idt.cpu().tolist()))
    ok = True
    # every mesh twice: thread-per-row kernels (tile_policy 1: one assembly per fespace never builds the tiles) and row
    # tiles for the matrix and the right-hand side (tile_policy 2)
    for dims, policy in [((6, 5, 7), 1), ((16, 12, 21), 1), ((6, 5, 7), 2), ((16, 12, 21), 2)]:
        ctx.set_option("tile_policy", policy)
        nx, ny, nz = dims
        qp, qw = ffcuda.quadrature(3, 6)
        mesh = ctx.mesh_cube(nx, ny, nz, distributed=True)
        nown, gid = mesh.local_to_global()
        sp = mesh.space(1, 1)
        pat = sp.symbolic()
        n, nnz = pat.info()
        assert n == nown
        A = pat.matrix()
        A.assemble(LAP, qp, qw)
        b = ctx.vec(n)
        sp.assemble_linear(b, RHS, qp, qw)
        bc = sp.bc_from_labels(ALL6, 1, [0.0])
        A.apply_bc(bc, 1e30)
        b.apply_bc(bc, 1e30)
        x = ctx.vec(len(gid))           # owned + ghost entries
        it, conv, _ = A.cg(b, x, eps=1e-6, itmax=0, tgv=1e30)
        # SpMV with halo exchange on x_i = sin(global id)
        xs = ctx.vec_from(np.sin(gid.astype(np.float64)))
        ys = ctx.vec(n)
        A.spmv(xs, ys)
        rp, ci = pat.download()
        pack = dict(rp=rp, cols=gid[ci], vals=A.download(), b=b.download(), u=x.download()[:n], y=ys.download(), gid=gid[:n], it=it,
                    conv=conv)
        allp = [None] * world
        dist.all_gather_object(allp, pack)
        if rank == 0:
            import oracle_lib as ol

            m = ol.cube(nx, ny, nz)
            N = m["xyz"].shape[0]
            oi, oj, oa = ol.assemble_coo(m, 1, 1, None, LAP, qp, qw)
            d, v = ol.bc_pairs(m, 1, 1, None, ALL6, 1, [0.0])
            oa = ol.bc_matrix_coo(oi, oj, oa, N, d, 1e30)
            ob = ol.bc_rhs(ol.assemble_rhs(m, 1, 1, None, N, RHS, qp, qw), d, v, 1e30)
            orp, ocol, oval = ol.coo_to_csr(N, oi, oj, oa)
            ox, oit, _, _ = ol.cg(N, oi, oj, oa, ob, np.zeros(N), eps=1e-6, itmax=0, tgv=1e30)
            oy = ol.spmv_coo(N, oi, oj, oa, np.sin(np.arange(N, dtype=np.float64)))
            # concatenation of the owned-row blocks = the global CSR (rows are owned in contiguous global ranges)
            g_rows = np.concatenate([p["gid"] for p in allp])
            assert np.array_equal(g_rows, np.arange(N)), "owned rows do not tile the global numbering"
            lens = np.concatenate([np.diff(p["rp"]) for p in allp])
            assert np.array_equal(lens, np.diff(orp)), "row lengths differ"
            # a rank's rows are sorted by LOCAL column id (owned first, then ghosts): bring every row to the global order
            cols = np.concatenate([p["cols"] for p in allp]).astype(np.int64)
            vals = np.concatenate([p["vals"] for p in allp])
            rowid = np.repeat(np.arange(N, dtype=np.int64), lens)
            o = np.argsort(rowid * N + cols, kind="stable")
            cols, vals = cols[o], vals[o]
            assert np.array_equal(cols, ocol), "column indices differ"
            reg = np.abs(oval) < 1e29
            assert np.array_equal(vals[~reg], oval[~reg])
            assert np.max(np.abs(vals - oval)[reg]) <= 1e-12 * np.abs(oval[reg]).max()
            bb = np.concatenate([p["b"] for p in allp])
            breg = np.abs(ob) < 1e20
            assert np.max(np.abs(bb - ob)[breg]) <= 1e-12 * np.abs(ob[breg]).max()
            yy = np.concatenate([p["y"] for p in allp])
            yreg = np.abs(oy) < 1e20
            assert np.max(np.abs(yy - oy)[yreg]) <= 1e-12 * np.abs(oy[yreg]).max()
            its = {p["it"] for p in allp}
            assert its == {oit} and all(p["conv"] == 1 for p in allp), (its, oit)
            uu = np.concatenate([p["u"] for p in allp])
            assert np.max(np.abs(uu - ox)) <= 1e-12 * np.abs(ox).max()
            print(f"dist_check cube{dims} tile_policy={policy} on {world} GPUs: n={N} nnz={len(ocol)} cg_iters={oit} OK", flush=True)
    dist.barrier()
    ctx.comm_finalize()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK_PASSED", flush=True)


if __name__ == "__main__":
    try:
        main()
    except BaseException:  # a failed rank must not leave the others waiting in a collective until the launcher's timeout
        import traceback

        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
