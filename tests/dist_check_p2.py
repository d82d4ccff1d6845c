#!/usr/bin/env python3
"""Multi-GPU parity for P2 spaces on an unstructured partition, launched with torchrun (one process per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29661 tests/dist_check_p2.py
The scrambled, warped cube of dist_check_rcb.py is split by recursive coordinate bisection; the P2 nodes (vertices and
edges, FreeFEM's global first-encounter numbering) follow the vertices (an edge goes with its end point of smaller id);
every rank derives ITS node-level local problem (ffcuda_partition_local_nodes), creates the space with its own node table
and halo lists (ffcuda_space_create_distributed), assembles its owned rows without communication and runs the distributed
CG.  Rank 0 gathers the owned rows with global column ids and compares with the CPU oracle on the whole mesh: pattern
bit-exact, values / rhs / SpMV 1e-12, CG iteration count equal and solution 1e-12 (eps = 1e-14: 1e-10) - the same bars as on
one GPU.  Also a [P2,P2,P2] Lame matrix (node-blocked halo of a vector P2 space)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ffcuda  # noqa: E402
from dist_check_rcb import ALL6, ID, RHS, local_problem, scrambled_cube  # noqa: E402

DX, DY, DZ = 1, 2, 6
LAP = [(0, DX, 0, DX, 1.0), (0, DY, 0, DY, 1.0), (0, DZ, 0, DZ, 1.0), (0, ID, 0, ID, 0.5)]


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

    for dims in [(5, 4, 6)]:
        m = scrambled_cube(dims, 11 + dims[0])
        nv = m["xyz"].shape[0]
        e2n, NN = ol.p2_nodes_3d(nv, m["conn"])
        qp, qw = ffcuda.quadrature(3, 6)
        mesh, me = local_problem(ctx, m, rank, world)          # vertex-level: the local mesh
        part = ffcuda.partition_rcb(m["xyz"], world)
        pn = fc.p2_node_partition(m["conn"], e2n, part)
        L = ffcuda.partition_local_nodes(e2n, NN, pn, rank, world)
        assert np.array_equal(L["elems"], me["elems"])
        no, l2g = L["nowned"], L["l2g"]
        g2l = -np.ones(NN, np.int64)
        g2l[l2g] = np.arange(len(l2g))
        sp = mesh.space_distributed(2, 1, g2l[e2n[L["elems"]]], no, len(l2g), L["nbr"], L["recv_off"], L["recv_cnt"], L["send_ptr"],
                                    L["send_idx"])
        pat = sp.symbolic()
        n, nnz = pat.info()
        assert n == no
        A = pat.matrix()
        A.assemble(LAP, qp, qw)
        b = ctx.vec(n)
        sp.assemble_linear(b, RHS, qp, qw)
        bc = sp.bc_from_labels(ALL6, 1, [0.0])
        A.apply_bc(bc, 1e30)
        b.apply_bc(bc, 1e30)
        x = ctx.vec(len(l2g))
        it, conv, _ = A.cg(b, x, eps=1e-6, itmax=0, tgv=1e30)
        x14 = ctx.vec(len(l2g))
        it14, conv14, _ = A.cg(b, x14, eps=1e-14, itmax=0, tgv=1e30)
        gid = l2g.astype(np.int64)
        xs = ctx.vec_from(np.sin(gid.astype(np.float64)))
        ys = ctx.vec(n)
        A.spmv(xs, ys)
        rp, ci = pat.download()
        # [P2,P2,P2] on the same node lists
        sp3 = mesh.space_distributed(2, 3, g2l[e2n[L["elems"]]], no, len(l2g), L["nbr"], L["recv_off"], L["recv_cnt"], L["send_ptr"],
                                     L["send_idx"])
        pat3 = sp3.symbolic()
        A3 = pat3.matrix()
        A3.assemble(fc.lame_terms(), qp, qw)
        n3 = pat3.info()[0]
        g3 = (gid[:, None] * 3 + np.arange(3)[None, :]).reshape(-1)
        y3 = ctx.vec(n3)
        A3.spmv(ctx.vec_from(np.cos(0.37 * g3.astype(np.float64))), y3)
        pack = dict(rp=rp, cols=gid[ci], vals=A.download(), b=b.download(), u=x.download()[:n], u14=x14.download()[:n], y=ys.download(),
                    gid=gid[:n], it=it, conv=conv, it14=it14, conv14=conv14, y3=y3.download(), nbrs=len(L["nbr"]))
        allp = [None] * world
        dist.all_gather_object(allp, pack)
        if rank == 0:
            N = NN
            oi, oj, oa = ol.assemble_coo(m, 2, 1, e2n, LAP, qp, qw)
            d, v = ol.bc_pairs(m, 2, 1, e2n, ALL6, 1, [0.0])
            oa = ol.bc_matrix_coo(oi, oj, oa, N, d, 1e30)
            ob = ol.bc_rhs(ol.assemble_rhs(m, 2, 1, e2n, N, RHS, qp, qw), d, v, 1e30)
            orp, ocol, oval = ol.coo_to_csr(N, oi, oj, oa)
            ox, oit, _, _ = ol.cg(N, oi, oj, oa, ob, np.zeros(N), eps=1e-6, itmax=0, tgv=1e30)
            ox14, oit14, _, _ = ol.cg(N, oi, oj, oa, ob, np.zeros(N), eps=1e-14, itmax=0, tgv=1e30)
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
            assert len(its) == 1 and abs(its.pop() - oit) <= 2 and all(p["conv"] == 1 for p in allp), (oit, [p["it"] for p in allp])
            uu = np.concatenate([p["u"] for p in allp])[inv]
            assert np.max(np.abs(uu - ox)) <= 1e-6 * np.abs(ox).max()       # (P2 iterate at eps = 1e-6: the single-GPU bar)
            assert all(p["conv14"] in (1, 2) for p in allp)
            uu14 = np.concatenate([p["u14"] for p in allp])[inv]
            assert np.max(np.abs(uu14 - ox14)) <= 1e-10 * np.abs(ox14).max()
            li, lj, la = ol.assemble_coo(m, 2, 3, e2n, fc.lame_terms(), qp, qw)
            oy3 = ol.spmv_coo(3 * N, li, lj, la, np.cos(0.37 * np.arange(3 * N, dtype=np.float64)))
            yy3 = np.concatenate([p["y3"] for p in allp]).reshape(-1, 3)[inv].reshape(-1)
            assert np.max(np.abs(yy3 - oy3)) <= 1e-12 * np.abs(oy3).max()
            print(f"dist_check_p2 cube{dims} on {world} GPUs: nodes={N} nnz={len(ocol)} cg_iters={oit} "
                  f"neighbours per rank {[p['nbrs'] for p in allp]} OK", flush=True)
    dist.barrier()
    ctx.comm_finalize()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK_P2_PASSED", flush=True)


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback

        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
