"""Multi-process host logic on CPU (gloo, world_size 2 and 3): every rank computes ITS slab of the cube partition with
the library's own arithmetic (ffcuda_partition_cube, the function the device mesh generator uses), the ranks exchange
their layouts over torch.distributed and check that they fit together: every vertex layer owned exactly once, every
cell present where a vertex it touches is owned, halo send/receive ranges of neighbours mirror each other.  Also: the
reference arm of bench.py under a multi-rank launch prints exactly one line (rank 0) and the other ranks exit 0."""
import json
import os
import subprocess
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, dims, q):
    import torch.distributed as dist

    sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
    import ffcuda

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nx, ny, nz = dims
        mine = ffcuda.partition_cube(nx, ny, nz, rank, world)
        allp = [None] * world
        dist.all_gather_object(allp, mine)
        nk = (nx + 1) * (ny + 1)
        # every vertex layer owned exactly once, in rank order
        owner = {}
        for r, p in enumerate(allp):
            for L in range(p["L0"], p["L0"] + p["nown"]):
                assert L not in owner
                owner[L] = r
        assert sorted(owner) == list(range(nz + 1))
        # my cells = exactly the cell layers touching one of my owned vertex layers
        need = {c for c in range(nz) if owner[c] == rank or owner[c + 1] == rank}
        assert need == set(range(mine["c_lo"], mine["c_lo"] + mine["ncl"]))
        assert mine["nt_local"] == 6 * nx * ny * mine["ncl"]
        assert mine["nv_owned"] == nk * mine["nown"]
        assert mine["nv_local"] == nk * (mine["nown"] + mine["has_lower"] + mine["has_upper"])
        assert mine["layer"] == nk
        # halo: what I send up is what my upper neighbour receives from below, and vice versa
        if mine["has_upper"]:
            up = allp[mine["nbr_hi"]]
            assert up["nbr_lo"] == rank and up["has_lower"] == 1
            assert mine["L0"] + mine["send_off_hi"] // nk == up["L0"] - 1           # my last owned layer = its lower ghost
            assert up["recv_off_lo"] == up["nv_owned"]
            assert up["L0"] + up["send_off_lo"] // nk == mine["L0"] + mine["nown"]  # its first layer = my upper ghost
            assert mine["recv_off_hi"] == mine["nv_owned"] + (nk if mine["has_lower"] else 0)
        else:
            assert mine["nbr_hi"] == -1 and mine["L0"] + mine["nown"] == nz + 1
        if not mine["has_lower"]:
            assert mine["nbr_lo"] == -1 and mine["L0"] == 0
        # global sums over the ranks
        import torch

        t = torch.tensor([mine["nv_owned"], mine["nown"]], dtype=torch.int64)
        dist.all_reduce(t)
        assert t.tolist() == [nk * (nz + 1), nz + 1]
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dims", [(2, (4, 3, 9)), (2, (128, 128, 256)), (3, (5, 5, 7)), (2, (3, 3, 1))])
def test_slab_partition_is_consistent_across_ranks(world, dims):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + hash(dims)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, dims, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def test_partition_rejects_more_ranks_than_layers():
    sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
    import ffcuda

    with pytest.raises(ffcuda.FfcudaError):
        ffcuda.partition_cube(4, 4, 1, 0, 3)


def test_reference_arm_under_multi_rank_launch():
    """torchrun-style environment with 2 ranks: rank 0 alone runs and prints the JSON line, rank 1 exits 0 silently."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "FreeFem++-nw")) and not os.path.exists(
            os.path.join(ROOT, "oracle", "liboracle.so")):
        pytest.skip("no CPU implementation built")
    outs = []
    for rank in (0, 1):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                            "--warmup", "0", "--ref-n", "8"], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip())
    line = json.loads(outs[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0 and line["unit"] == "nnz/s"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["cores"] == 1
    assert outs[1] == ""


# ------------------------------------------------------------------------------------------------------------------
# general (unstructured) meshes: recursive coordinate bisection + the local problem of every rank (host arithmetic only)
# ------------------------------------------------------------------------------------------------------------------
def _general_mesh(kind):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib as ol

    m = ol.cube(6, 5, 4) if kind == "cube" else ol.square(11, 9)
    # an unstructured-looking numbering: vertices and elements shuffled, coordinates warped
    rng = np.random.default_rng(7)
    nv, nt = m["xyz"].shape[0], m["conn"].shape[0]
    pv, pe = rng.permutation(nv), rng.permutation(nt)
    inv = np.empty(nv, np.int64)
    inv[pv] = np.arange(nv)
    xyz = m["xyz"][pv].copy()
    xyz[:, 0] += 0.15 * xyz[:, 1] ** 2
    conn = inv[m["conn"]][pe].astype(np.int32)
    return dict(dim=m["dim"], xyz=xyz, conn=conn, elab=np.zeros(nt, np.int32))


def _worker_general(rank, world, port, kind, q):
    import numpy as np
    import torch.distributed as dist

    sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ffcuda
    import ff_cases as fc
    import oracle_lib as ol

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = _general_mesh(kind)
        dim, xyz, conn = m["dim"], m["xyz"], m["conn"]
        nv, nt = xyz.shape[0], conn.shape[0]
        part = ffcuda.partition_rcb(xyz, world)                      # the same on every rank (deterministic)
        sizes = np.bincount(part, minlength=world)
        assert sizes.sum() == nv and sizes.max() - sizes.min() <= 2
        me = ffcuda.partition_local(dim, nv, conn, part, rank, world)
        no, l2g = me["nowned"], me["l2g"]
        # owned = my part, ascending; local elements = exactly those touching an owned vertex; ghosts = their other vertices
        assert np.array_equal(l2g[:no], np.flatnonzero(part == rank))
        touch = (part[conn] == rank).any(axis=1)
        assert np.array_equal(me["elems"], np.flatnonzero(touch))
        gh = np.setdiff1d(np.unique(conn[touch]), l2g[:no])
        assert np.array_equal(np.sort(l2g[no:]), gh)
        key = part[l2g[no:]].astype(np.int64) * nv + l2g[no:]
        assert np.all(np.diff(key) > 0)                              # grouped by owner, ascending id inside
        for x, r in enumerate(me["nbr"]):
            rng_ = l2g[me["recv_off"][x]:me["recv_off"][x] + me["recv_cnt"][x]]
            assert np.all(part[rng_] == r)
        allp = [None] * world
        dist.all_gather_object(allp, me)
        # halo lists mirror each other: what I gather for r is r's contiguous ghost range owned by me, in order
        for x, r in enumerate(me["nbr"]):
            other = allp[r]
            assert rank in other["nbr"]
            y = list(other["nbr"]).index(rank)
            mine_send = l2g[me["send_idx"][me["send_ptr"][x]:me["send_ptr"][x + 1]]]
            their_recv = other["l2g"][other["recv_off"][y]:other["recv_off"][y] + other["recv_cnt"][y]]
            assert np.array_equal(mine_send, their_recv)
        # every vertex owned exactly once, every element local somewhere
        owned_all = np.concatenate([p["l2g"][:p["nowned"]] for p in allp])
        assert np.array_equal(np.sort(owned_all), np.arange(nv))
        assert np.array_equal(np.unique(np.concatenate([p["elems"] for p in allp])), np.arange(nt))
        # a halo exchange carried out over gloo: ghosts end up with the owners' values
        xg = np.sin(np.arange(nv, dtype=np.float64))
        xl = np.zeros(len(l2g))
        xl[:no] = xg[l2g[:no]]
        payload = {int(r): xl[me["send_idx"][me["send_ptr"][x]:me["send_ptr"][x + 1]]] for x, r in enumerate(me["nbr"])}
        allpay = [None] * world
        dist.all_gather_object(allpay, payload)
        for x, r in enumerate(me["nbr"]):
            xl[me["recv_off"][x]:me["recv_off"][x] + me["recv_cnt"][x]] = allpay[r][rank]
        assert np.array_equal(xl, xg[l2g])
        # assembly without communication: the P1 Laplace matrix of the LOCAL mesh has, in its owned rows, exactly the rows
        # of the global matrix (pattern and values), and the local product with the exchanged vector is the global one
        qp, qw = ol.quadrature(dim, "qfV5" if dim == 3 else "qf5pT")
        lap = fc.LAP3 if dim == 3 else fc.LAP2
        g2l = -np.ones(nv, np.int64)
        g2l[l2g] = np.arange(len(l2g))
        ml = dict(dim=dim, xyz=xyz[l2g], conn=g2l[conn[me["elems"]]].astype(np.int32), elab=np.zeros(len(me["elems"]), np.int32))
        li, lj, la = ol.assemble_coo(ml, 1, 1, None, lap, qp, qw)
        gi, gj, ga = ol.assemble_coo(m, 1, 1, None, lap, qp, qw)
        import scipy.sparse as sps
        Ag = sps.coo_matrix((ga, (gi, gj)), shape=(nv, nv)).tocsr()
        rows = li < no
        Al = sps.coo_matrix((la[rows], (li[rows], l2g[lj[rows]])), shape=(no, nv)).tocsr()
        Aref = Ag[l2g[:no]]
        Al.sort_indices()
        Aref.sort_indices()
        assert np.array_equal(Al.indptr, Aref.indptr) and np.array_equal(Al.indices, Aref.indices)     # structural zeros included
        assert np.max(np.abs(Al.data - Aref.data)) <= 1e-13 * np.abs(Aref.data).max()
        yl = sps.coo_matrix((la[rows], (li[rows], lj[rows])), shape=(no, len(l2g))).tocsr() @ xl
        assert np.max(np.abs(yl - (Ag @ xg)[l2g[:no]])) <= 1e-12 * np.abs(Ag @ xg).max()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, repr(e) + traceback.format_exc()[-600:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,kind", [(2, "cube"), (3, "cube"), (2, "square"), (4, "square")])
def test_general_partition_is_consistent_across_ranks(world, kind):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + 17 * world + len(kind)) % 2000
    procs = [ctx.Process(target=_worker_general, args=(r, world, port, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def test_general_partition_edge_cases():
    """single process: one part, as many parts as vertices, identical coordinates (ties by vertex id), and the local lists
    of an arbitrary (random, non-geometric) vertex partition checked against a brute-force construction."""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
    import ffcuda

    m = _general_mesh("square")
    xyz, conn, nv = m["xyz"], m["conn"], m["xyz"].shape[0]
    assert np.all(ffcuda.partition_rcb(xyz, 1) == 0)
    p = ffcuda.partition_rcb(xyz, nv)
    assert np.array_equal(np.sort(p), np.arange(nv))                         # one vertex per part
    p = ffcuda.partition_rcb(np.zeros((10, 3)), 4)                            # all points equal: split by id, sizes 2 3 2 3
    assert np.array_equal(p, [0, 0, 1, 1, 1, 2, 2, 3, 3, 3])
    p5 = ffcuda.partition_rcb(xyz, 5)
    assert np.bincount(p5, minlength=5).max() - np.bincount(p5, minlength=5).min() <= 2
    # parts of an RCB partition are boxes in the split direction: the first split separates the point set by a plane
    p2 = ffcuda.partition_rcb(xyz, 2)
    ext = xyz.max(axis=0) - xyz.min(axis=0)
    ax = int(np.argmax(ext))
    assert xyz[p2 == 0, ax].max() <= xyz[p2 == 1, ax].min()
    rng = np.random.default_rng(1)
    part = rng.integers(0, 3, nv).astype(np.int32)
    for rank in range(3):
        L = ffcuda.partition_local(2, nv, conn, part, rank, 3)
        touch = (part[conn] == rank).any(axis=1)
        assert np.array_equal(L["elems"], np.flatnonzero(touch))
        assert np.array_equal(L["l2g"][:L["nowned"]], np.flatnonzero(part == rank))
        gh = np.setdiff1d(np.unique(conn[touch]), np.flatnonzero(part == rank))
        assert np.array_equal(np.sort(L["l2g"][L["nowned"]:]), gh)
        assert L["send_ptr"][0] == 0 and L["send_ptr"][-1] == len(L["send_idx"]) and np.all(L["send_idx"] < L["nowned"])
        for x, r in enumerate(L["nbr"]):
            sent = L["l2g"][L["send_idx"][L["send_ptr"][x]:L["send_ptr"][x + 1]]]
            other = ffcuda.partition_local(2, nv, conn, part, int(r), 3)
            y = list(other["nbr"]).index(rank)
            assert np.array_equal(sent, other["l2g"][other["recv_off"][y]:other["recv_off"][y] + other["recv_cnt"][y]])
    with pytest.raises(ffcuda.FfcudaError):
        ffcuda.partition_rcb(xyz, 0)
    # malformed inputs come back as errors through the ABI: a partition vector made for more parts than ranks,
    # a connectivity entry outside the mesh, more parts than vertices
    with pytest.raises(ffcuda.FfcudaError):
        ffcuda.partition_local(2, nv, conn, ffcuda.partition_rcb(xyz, 5), 0, 3)
    bad = conn.copy()
    bad[3, 1] = nv
    with pytest.raises(ffcuda.FfcudaError):
        ffcuda.partition_local(2, nv, bad, part, 0, 3)
    with pytest.raises(ffcuda.FfcudaError):
        ffcuda.partition_rcb(xyz[:4], 5)


def test_node_partition_p2_local_problem():
    """P2 spaces on a distributed mesh (ffcuda_partition_local_nodes): for every rank of a 3-way RCB partition of a cube the
    node-level local problem has the elements of the vertex-level one, owns each node exactly once, its send lists are the
    neighbours' ghost ranges in order, and the P2 Laplace matrix of the LOCAL mesh with the LOCAL node table holds, in its
    owned rows, exactly the rows of the global matrix (structural zeros included): assembly needs no communication."""
    import numpy as np
    import scipy.sparse as sps

    sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ffcuda
    import ff_cases as fc
    import oracle_lib as ol

    m = _general_mesh("cube")
    xyz, conn, nv = m["xyz"], m["conn"], m["xyz"].shape[0]
    e2n, nn = ol.p2_nodes_3d(nv, conn)
    world = 3
    part = ffcuda.partition_rcb(xyz, world)
    pn = fc.p2_node_partition(conn, e2n, part)
    qp, qw = ol.quadrature(3, "qfV5")
    gi, gj, ga = ol.assemble_coo(m, 2, 1, e2n, fc.LAP3, qp, qw)
    Ag = sps.coo_matrix((ga, (gi, gj)), shape=(nn, nn)).tocsr()
    Ls = [ffcuda.partition_local_nodes(e2n, nn, pn, r, world) for r in range(world)]
    assert np.array_equal(np.sort(np.concatenate([L["l2g"][:L["nowned"]] for L in Ls])), np.arange(nn))
    for rank, L in enumerate(Ls):
        Lv = ffcuda.partition_local(3, nv, conn, part, rank, world)
        assert np.array_equal(L["elems"], Lv["elems"])
        no, l2g = L["nowned"], L["l2g"]
        assert np.array_equal(l2g[:no], np.flatnonzero(pn == rank))
        for x, r in enumerate(L["nbr"]):
            other = Ls[int(r)]
            y = list(other["nbr"]).index(rank)
            sent = l2g[L["send_idx"][L["send_ptr"][x]:L["send_ptr"][x + 1]]]
            assert np.array_equal(sent, other["l2g"][other["recv_off"][y]:other["recv_off"][y] + other["recv_cnt"][y]])
        g2l_v = -np.ones(nv, np.int64)
        g2l_v[Lv["l2g"]] = np.arange(len(Lv["l2g"]))
        g2l_n = -np.ones(nn, np.int64)
        g2l_n[l2g] = np.arange(len(l2g))
        ml = dict(dim=3, xyz=xyz[Lv["l2g"]], conn=g2l_v[conn[L["elems"]]].astype(np.int32), elab=np.zeros(len(L["elems"]), np.int32))
        e2n_l = g2l_n[e2n[L["elems"]]].astype(np.int32)
        assert e2n_l.min() >= 0
        li, lj, la = ol.assemble_coo(ml, 2, 1, e2n_l, fc.LAP3, qp, qw)
        rows = li < no
        Al = sps.coo_matrix((la[rows], (li[rows], l2g[lj[rows]])), shape=(no, nn)).tocsr()
        Aref = Ag[l2g[:no]]
        Al.sort_indices()
        Aref.sort_indices()
        assert np.array_equal(Al.indptr, Aref.indptr) and np.array_equal(Al.indices, Aref.indices)
        assert np.max(np.abs(Al.data - Aref.data)) <= 1e-13 * np.abs(Aref.data).max()


@pytest.mark.parametrize("world,structure", [(2, "sym"), (3, "sym"), (4, "nonsym")])
def test_row_block_partition_of_a_host_matrix(world, structure):
    """ffcuda_partition_rows_local (what the plugin does with a MatriceMorse on several GPUs): the blocks tile the rows, ghost
    ranges are grouped by owner in ascending order, every send list is the neighbour's ghost range in its order, neighbours
    list each other even when the structure is not symmetric, and block products with exchanged ghosts give the global
    product."""
    import numpy as np
    import scipy.sparse as sps

    sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ffcuda
    import ff_cases as fc
    import oracle_lib as ol

    m = ol.cube(4, 3, 6)
    n = m["xyz"].shape[0]
    qp, qw = ol.quadrature(3, "qfV5")
    gi, gj, ga = ol.assemble_coo(m, 1, 1, None, fc.LAP3 + [(0, fc.DX, 0, fc.ID, 3.0)], qp, qw)
    A = sps.coo_matrix((ga, (gi, gj)), shape=(n, n)).tocsr()
    if structure == "nonsym":   # lower triangle only: a block needs nothing from the blocks after it, they need its values
        coo = A.tocoo()
        keep = coo.col <= coo.row
        A = sps.coo_matrix((coo.data[keep], (coo.row[keep], coo.col[keep])), shape=(n, n)).tocsr()
    A.sort_indices()
    rp, ci, va = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data
    Ls = [ffcuda.partition_rows_local(n, rp, ci, r, world) for r in range(world)]
    assert [L["first"] for L in Ls] == [n * r // world for r in range(world)]
    assert sum(L["nowned"] for L in Ls) == n
    xg = np.sin(np.arange(n, dtype=np.float64))
    y = A @ xg
    for r, L in enumerate(Ls):
        no, l2g, lo = L["nowned"], L["l2g"], L["first"]
        assert np.array_equal(l2g[:no], np.arange(lo, lo + no)) and np.all(np.diff(l2g[no:]) > 0)
        assert np.array_equal(l2g[L["colind"]], ci[rp[lo]:rp[lo + no]])
        assert np.array_equal(L["rowptr"], rp[lo:lo + no + 1] - rp[lo])
        off = no
        for x, o in enumerate(L["nbr"]):
            assert L["recv_off"][x] == off
            off += L["recv_cnt"][x]
            gh = l2g[L["recv_off"][x]:L["recv_off"][x] + L["recv_cnt"][x]]
            assert np.all((gh >= Ls[o]["first"]) & (gh < Ls[o]["first"] + Ls[o]["nowned"]))
            other = Ls[int(o)]
            assert r in other["nbr"]                                   # mutual, whatever the direction of the data
            yx = list(other["nbr"]).index(r)
            sent = other["l2g"][other["send_idx"][other["send_ptr"][yx]:other["send_ptr"][yx + 1]]]
            assert np.array_equal(sent, gh)
        assert off == len(l2g)
        # the block product with the exchanged ghost values
        Al = sps.csr_matrix((va[rp[lo]:rp[lo + no]], L["colind"], L["rowptr"]), shape=(no, len(l2g)))
        assert np.max(np.abs(Al @ xg[l2g] - y[lo:lo + no])) <= 1e-13 * np.abs(y).max()
    if structure == "nonsym":
        assert any(0 in L["recv_cnt"] or 0 in np.diff(L["send_ptr"]) for L in Ls)   # a one-way neighbour exists
