#!/usr/bin/env python3
"""Multi-GPU parity on an UNSTRUCTURED partition, launched with torchrun (one process per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631 tests/dist_check_rcb.py
A cube mesh with shuffled vertex / element numbering and warped coordinates is split by recursive coordinate bisection
(ffcuda_partition_rcb: the vertex -> rank vector a METIS call would give, plugin/seq/metis.cpp); every rank builds ITS
local problem (ffcuda_partition_local), uploads it (ffcuda_mesh_upload_distributed), assembles its owned rows without
communication and runs the distributed CG (halo exchange with gather lists and any number of neighbours).  Rank 0
gathers the owned rows with global column ids and compares with the CPU oracle on the whole mesh: pattern bit-exact,
values / rhs / SpMV 1e-12, CG iteration count equal and solution 1e-12 — the same bars as on one GPU.  Also a
[P1,P1,P1] Lame matrix (vector space on a distributed mesh: node-blocked halo) and, on the small mesh, a non-symmetric
convection-diffusion matrix solved by the distributed GMRES(30) (restarts included) against the oracle's fgmres."""
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
LAP = [(0, DX, 0, DX, 1.0), (0, DY, 0, DY, 1.0), (0, DZ, 0, DZ, 1.0), (0, ID, 0, ID, 0.5)]
RHS = [(0, ID, 1.0)]
ALL6 = [1, 2, 3, 4, 5, 6]
CONV = LAP + [(0, DX, 0, ID, 8.0), (0, DY, 0, ID, 3.0), (0, DZ, 0, ID, -2.0)]


def scrambled_cube(dims, seed):
    import oracle_lib as ol

    m = ol.cube(*dims)
    rng = np.random.default_rng(seed)
    nv, nt, nbe = m["xyz"].shape[0], m["conn"].shape[0], m["bconn"].shape[0]
    pv, pe = rng.permutation(nv), rng.permutation(nt)
    inv = np.empty(nv, np.int64)
    inv[pv] = np.arange(nv)
    inve = np.empty(nt, np.int64)
    inve[pe] = np.arange(nt)
    xyz = m["xyz"][pv].copy()
    xyz[:, 0] += 0.15 * xyz[:, 1] ** 2
    return dict(dim=3, xyz=xyz, conn=inv[m["conn"]][pe].astype(np.int32), elab=m["elab"][pe].astype(np.int32),
                bconn=inv[m["bconn"]].astype(np.int32), blab=m["blab"].astype(np.int32), belem=inve[m["belem"]].astype(np.int32),
                bface=m["bface"].astype(np.int32))


def local_problem(ctx, m, rank, world):
    dim, xyz, conn = m["dim"], m["xyz"], m["conn"]
    nv = xyz.shape[0]
    part = ffcuda.partition_rcb(xyz, world)
    me = ffcuda.partition_local(dim, nv, conn, part, rank, world)
    no, l2g, elems = me["nowned"], me["l2g"], me["elems"]
    g2l = -np.ones(nv, np.int64)
    g2l[l2g] = np.arange(len(l2g))
    e2l = -np.ones(conn.shape[0], np.int64)
    e2l[elems] = np.arange(len(elems))
    keep = e2l[m["belem"]] >= 0                      # boundary elements whose element is local
    mesh = ctx.mesh_upload_distributed(dim, no, xyz[l2g], g2l[conn[elems]], m["elab"][elems], g2l[m["bconn"][keep]], m["blab"][keep],
                                       e2l[m["belem"][keep]], m["bface"][keep], l2g, me["nbr"], me["recv_off"], me["recv_cnt"],
                                       me["send_ptr"], me["send_idx"])
    return mesh, me


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
    import ff_cases as fc
    import oracle_lib as ol

    for dims, policy in [((6, 5, 7), 1), ((14, 11, 13), 2)]:
        ctx.set_option("tile_policy", policy)
        m = scrambled_cube(dims, 7 + dims[0])
        N = m["xyz"].shape[0]
        qp, qw = ffcuda.quadrature(3, 6)
        mesh, me = local_problem(ctx, m, rank, world)
        nown, gid = mesh.local_to_global()
        assert nown == me["nowned"] and np.array_equal(gid, me["l2g"])
        sp = mesh.space(1, 1)
        pat = sp.symbolic()
        n, nnz = pat.info()
        assert n == nown
        A = pat.matrix()
        A.assemble(LAP, qp, qw)
        if policy == 2:
            A.assemble(LAP, qp, qw)      # tiles / fans of the local mesh
        b = ctx.vec(n)
        sp.assemble_linear(b, RHS, qp, qw)
        bc = sp.bc_from_labels(ALL6, 1, [0.0])
        A.apply_bc(bc, 1e30)
        b.apply_bc(bc, 1e30)
        x = ctx.vec(len(gid))
        it, conv, _ = A.cg(b, x, eps=1e-6, itmax=0, tgv=1e30)
        xs = ctx.vec_from(np.sin(gid.astype(np.float64)))
        ys = ctx.vec(n)
        A.spmv(xs, ys)
        rp, ci = pat.download()
        # vector space on the same distributed mesh: [P1,P1,P1] Lame, SpMV with node-blocked halo
        sp3 = mesh.space(1, 3)
        pat3 = sp3.symbolic()
        A3 = pat3.matrix()
        A3.assemble(fc.lame_terms(), qp, qw)
        n3 = pat3.info()[0]
        g3 = (gid[:, None] * 3 + np.arange(3)[None, :]).reshape(-1)
        y3 = ctx.vec(n3)
        A3.spmv(ctx.vec_from(np.cos(0.37 * g3.astype(np.float64))), y3)
        pack = dict(rp=rp, cols=gid[ci], vals=A.download(), b=b.download(), u=x.download()[:n], y=ys.download(), gid=gid[:n], it=it,
                    conv=conv, y3=y3.download(), nbrs=len(me["nbr"]))
        if policy == 1:  # distributed GMRES on a non-symmetric matrix (same pattern object, values re-assembled)
            A.assemble(CONV, qp, qw)
            A.apply_bc(bc, 1e30)
            for eps, tag in ((1e-6, "g6"), (1e-14, "g14")):
                xg = ctx.vec(len(gid))
                git, gconv, _ = A.gmres(b, xg, eps=eps, itmax=0, restart=30, tgv=1e30)
                pack[tag] = (xg.download()[:n], git, gconv)
        allp = [None] * world
        dist.all_gather_object(allp, pack)
        if rank == 0:
            oi, oj, oa = ol.assemble_coo(m, 1, 1, None, LAP, qp, qw)
            d, v = ol.bc_pairs(m, 1, 1, None, ALL6, 1, [0.0])
            oa = ol.bc_matrix_coo(oi, oj, oa, N, d, 1e30)
            ob = ol.bc_rhs(ol.assemble_rhs(m, 1, 1, None, N, RHS, qp, qw), d, v, 1e30)
            orp, ocol, oval = ol.coo_to_csr(N, oi, oj, oa)
            ox, oit, _, _ = ol.cg(N, oi, oj, oa, ob, np.zeros(N), eps=1e-6, itmax=0, tgv=1e30)
            oy = ol.spmv_coo(N, oi, oj, oa, np.sin(np.arange(N, dtype=np.float64)))
            rows = np.concatenate([p["gid"] for p in allp])
            assert np.array_equal(np.sort(rows), np.arange(N)), "owned rows do not tile the global numbering"
            lens = np.concatenate([np.diff(p["rp"]) for p in allp])
            rowid = np.repeat(rows, lens)
            cols = np.concatenate([p["cols"] for p in allp]).astype(np.int64)
            vals = np.concatenate([p["vals"] for p in allp])
            o = np.argsort(rowid * N + cols, kind="stable")
            rowid, cols, vals = rowid[o], cols[o], vals[o]
            assert np.array_equal(np.bincount(rowid, minlength=N), np.diff(orp)), "row lengths differ"
            assert np.array_equal(cols, ocol), "column indices differ"
            reg = np.abs(oval) < 1e29
            assert np.array_equal(vals[~reg], oval[~reg])
            assert np.max(np.abs(vals - oval)[reg]) <= 1e-12 * np.abs(oval[reg]).max()
            inv = np.argsort(rows)
            bb = np.concatenate([p["b"] for p in allp])[inv]
            breg = np.abs(ob) < 1e20
            assert np.max(np.abs(bb - ob)[breg]) <= 1e-12 * np.abs(ob[breg]).max()
            yy = np.concatenate([p["y"] for p in allp])[inv]
            yreg = np.abs(oy) < 1e20
            assert np.max(np.abs(yy - oy)[yreg]) <= 1e-12 * np.abs(oy[yreg]).max()
            its = {p["it"] for p in allp}
            assert its == {oit} and all(p["conv"] == 1 for p in allp), (its, oit)
            uu = np.concatenate([p["u"] for p in allp])[inv]
            assert np.max(np.abs(uu - ox)) <= 1e-12 * np.abs(ox).max()
            if policy == 1:
                gi, gj, ga = ol.assemble_coo(m, 1, 1, None, CONV, qp, qw)
                ga = ol.bc_matrix_coo(gi, gj, ga, N, d, 1e30)
                for eps, tag, tol in ((1e-6, "g6", 1e-7), (1e-14, "g14", 1e-10)):
                    ogx, ogit, ogret, _ = ol.gmres(N, gi, gj, ga, ob, np.zeros(N), eps=eps, nbkrylov=30, tgv=1e30)
                    gits = {p[tag][1] for p in allp}
                    assert ogret == 1 and all(p[tag][2] == 1 for p in allp) and len(gits) == 1, (gits, ogit)
                    assert abs(gits.pop() - ogit) <= 1, (tag, ogit)
                    gu = np.concatenate([p[tag][0] for p in allp])[inv]
                    assert np.max(np.abs(gu - ogx)) <= tol * np.abs(ogx).max(), (tag, np.max(np.abs(gu - ogx)) / np.abs(ogx).max())
                print(f"dist_check_rcb cube{dims}: distributed GMRES(30) on {world} GPUs: {ogit} iterations OK", flush=True)
            # Lame
            li, lj, la = ol.assemble_coo(m, 1, 3, None, fc.lame_terms(), qp, qw)
            oy3 = ol.spmv_coo(3 * N, li, lj, la, np.cos(0.37 * np.arange(3 * N, dtype=np.float64)))
            yy3 = np.concatenate([p["y3"] for p in allp]).reshape(-1, 3)[inv].reshape(-1)
            assert np.max(np.abs(yy3 - oy3)) <= 1e-12 * np.abs(oy3).max()
            print(f"dist_check_rcb cube{dims} tile_policy={policy} on {world} GPUs: n={N} nnz={len(ocol)} cg_iters={oit} "
                  f"neighbours per rank {[p['nbrs'] for p in allp]} OK", flush=True)
    dist.barrier()
    ctx.comm_finalize()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK_RCB_PASSED", flush=True)


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback

        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
