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
