#!/usr/bin/env python3
"""bench.py — the reference's headline workload on the ffcuda hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 128]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 3-D P1 Poisson on cube(n,n,n), n = 128 (12.58 M tets, 2.15 M dofs, 31.8 M nnz),
f = 1, u = 0 on the six faces (tgv = 1e30), FreeFEM's default 14-point rule, Jacobi-CG to eps = 1e-6 — the sequence
    matrix A = vLap(Vh,Vh,solver=CG);  real[int] b = vLap(0,Vh);  u[] = A^-1*b;
ONE STEP = one pass of the assembly path: symbolic sparsity -> numeric assembly -> Dirichlet rows -> right-hand side
(what `matrix A = ...; real[int] b = ...;` cost in FreeFEM, which rebuilds the pattern on every such statement).  The
CG solve that consumes A and b is timed right after, in the same run, with its own CUDA events, and reported in the
"cg"/"spmv" objects (BASELINE.json's metric names two figures; `value` can carry one).  At N > 1 the cube is stretched
so that every GPU keeps the n^3-cell share (weak scaling): N=2 cube(n,n,2n), N=4 cube(n,2n,2n), N=8 cube(2n,2n,2n) =
BASELINE.json configs[4] for n = 128; slab partition along z, NCCL halo exchange + all-reduce inside CG.

value   = nnz / t_step  [nnz/s]: assembly throughput with mesh and dof map already resident in HBM.  On the CPU this
          figure is size-independent (0.38-0.44 M nnz/s from cube(16) to cube(128)), so the bounded sample
          of the reference arm compares like with like.
spmv    = GB/s of the SpMV kernel inside CG (algorithmic bytes 12*nnz + 20*n per launch) and its fraction of the
          measured HBM roofline (MEASURED_PEAKS.json); cg = whole solve (iterations, ms, ms per iteration).
e2e     = the assembly step through the C ABI with HOST buffers, as the FreeFEM plugin calls it: the host mesh is
          uploaded, the CSR matrix and b are copied back to the host (FreeFEM's MatriceMorse); e2e.solve_ms is the
          solver entry with host b/x (u[] = A^-1*b).  All copies are inside the timed region, host buffers are pinned.
--impl reference: the unmodified FreeFem++ (oracle/_ref, one thread: the path is single-threaded by construction) runs
          the same .edp on a bounded sample of the workload (cube(m), m < n) and reports the same quantities.
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "freefem-sources_b200"))

ID, DX, DY, DZ = 0, 1, 2, 6
LAP3 = [(0, DX, 0, DX, 1.0), (0, DY, 0, DY, 1.0), (0, DZ, 0, DZ, 1.0)]
RHS = [(0, ID, 1.0)]
ALL6 = [1, 2, 3, 4, 5, 6]
TGV = 1e30
EPS = 1e-6
METRIC = ("assembly nnz/s + CG SpMV GB/s vs HBM roofline (value = assembly nnz/s: symbolic + numeric + Dirichlet + rhs; "
          "CG SpMV GB/s in 'spmv')")
FF_BIN = os.path.join(ROOT, "oracle", "_ref", "FreeFem++-nw")


def load_traffic():
    """per-kernel DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum per launch) written by tools/ncu_extract.py from
    the `ncu --set full` capture of the committed build: the newest profiles/*_traffic.json.  Nothing is hard-coded."""
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return {}, None
    try:
        with open(files[-1]) as f:
            return json.load(f), os.path.relpath(files[-1], ROOT)
    except Exception:
        return {}, None


CUBE_STENCIL = [(0, 0, 0)] + [(sx * a, sx * b, sx * c) for sx in (1, -1)
                              for (a, b, c) in ((1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (1, 1, 1))]


def cube_pattern_mismatches(np, nx, ny, nz, row_gid, rp, gcols):
    """rows of the P1 matrix on BuildCube's mesh in closed form (fflib/msh3.cpp:7683-7742: every cell is cut in 6 tets
    around the diagonal 0-7, so vertex (i,j,k) is coupled with itself and +-(1,0,0) (0,1,0) (0,0,1) (1,1,0) (0,1,1)
    (1,0,1) (1,1,1) when that vertex is in the box; HashMatrix keeps every couple of an element, zero or not).  Compares
    row lengths and an order-independent 64-bit hash of every row's global column set; returns the number of rows that differ."""
    g = row_gid.astype(np.int64)
    sx, sxy = nx + 1, (nx + 1) * (ny + 1)
    i, j, k = g % sx, (g // sx) % (ny + 1), g // sxy
    mul = np.uint64(0x9E3779B97F4A7C15)
    cnt = np.zeros(len(g), np.int64)
    hsh = np.zeros(len(g), np.uint64)
    for (a, b, c) in CUBE_STENCIL:
        ok = (i + a >= 0) & (i + a <= nx) & (j + b >= 0) & (j + b <= ny) & (k + c >= 0) & (k + c <= nz)
        nid = (g + a + b * sx + c * sxy).astype(np.uint64)
        cnt += ok
        hsh += np.where(ok, (nid + np.uint64(1)) * mul, np.uint64(0))
    lens = np.diff(rp.astype(np.int64))
    got = np.add.reduceat((gcols.astype(np.uint64) + np.uint64(1)) * mul, rp[:-1].astype(np.int64)) if len(gcols) else hsh * np.uint64(0)
    return int(np.count_nonzero((lens != cnt) | (got != hsh)))


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            p = [x.strip() for x in ln.split(",")]
            if len(p) >= 6 and p[0].isdigit():
                self.samples.append(p)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [int(s[0]) for s in self.samples]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": int(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the unmodified FreeFem++ on a bounded sample
# ------------------------------------------------------------------------------------------------------------------
EDP = """load "msh3"
%s
int n = %d;
exec("date +%%s.%%N >> ffstamps.txt");
real tm0 = clock();
mesh3 Th = cube(n,n,n);
real tm1 = clock();
exec("date +%%s.%%N >> ffstamps.txt");
fespace Vh(Th,P1);
varf va(u,v) = int3d(Th)(dx(u)*dx(v)+dy(u)*dy(v)+dz(u)*dz(v)) + int3d(Th)(1.*v) + on(1,2,3,4,5,6,u=0);
verbosity = 1;
exec("date +%%s.%%N >> ffstamps.txt");
real t0 = clock();
matrix A = va(Vh,Vh,solver=CG,eps=1e-6);
real t1 = clock();
exec("date +%%s.%%N >> ffstamps.txt");
real[int] b = va(0,Vh);
real t2 = clock();
exec("date +%%s.%%N >> ffstamps.txt");
Vh u; u[] = 0;
u[] = A^-1*b;
real t3 = clock();
exec("date +%%s.%%N >> ffstamps.txt");
cout.precision(12);
cout << "FFBENCH nt " << Th.nt << " n " << Vh.ndof << " nnz " << A.nnz << " tA " << t1-t0 << " tb " << t2-t1
     << " tcg " << t3-t2 << " uu " << u[]'*u[] << " tmesh " << tm1-tm0 << endl;
"""


def run_reference_once(m, plugin=False):
    """one run of the reference on cube(m): dict(nnz, n, iters, tA, tb, tcg).  plugin=True: the same script with
    `load "ffcuda"` (the drop-in as a script author sees it: FreeFEM's interpreter, host mesh, MatriceMorse hand-off)."""
    with tempfile.TemporaryDirectory() as td:
        edp = os.path.join(td, "bench.edp")
        with open(edp, "w") as f:
            f.write(EDP % ('load "ffcuda"' if plugin else "", m))
        env = dict(os.environ, FF_LOADPATH=os.path.join(ROOT, "freefem-sources_b200", "lib"))
        r = subprocess.run([FF_BIN, "-nw", "-v", "1", edp], capture_output=True, text=True, cwd=td, env=env)
        # wall clock of the three statements: time stamps the script writes with exec("date ...") (clock() is CPU time of the
        # whole process, time() has a resolution of one second)
        wall = [0.0, 0.0, 0.0, 0.0]
        try:
            st = [float(x) for x in open(os.path.join(td, "ffstamps.txt")).read().split()]
            if len(st) >= 6:  # mesh statement, then the three statements of the path
                wall = [st[3] - st[2], st[4] - st[3], st[5] - st[4], st[1] - st[0]]
        except Exception:
            pass
    mm = re.search(r"FFBENCH nt (\d+) n (\d+) nnz (\d+) tA (\S+) tb (\S+) tcg (\S+) uu (\S+) tmesh (\S+)", r.stdout)
    it = re.search(r"GC[^\n]*?converge after\s+(\d+)", r.stdout)
    if r.returncode != 0 or not mm or not it:
        raise RuntimeError("reference run failed: " + (r.stdout[-500:] + r.stderr[-500:]))
    return dict(nt=int(mm.group(1)), n=int(mm.group(2)), nnz=int(mm.group(3)), tA=float(mm.group(4)), tb=float(mm.group(5)),
                tcg=float(mm.group(6)), uu=float(mm.group(7)), tmesh=float(mm.group(8)), wA=wall[0], wb=wall[1], wcg=wall[2], wmesh=wall[3], iters=int(it.group(1)),
                gpu_path="assembled on the GPU" in r.stdout or "(ffcuda)" in r.stdout)


def run_port_once(m):
    """the oracle's C restatement on cube(m) (used only where the reference binary is absent)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib as ol

    mesh = ol.cube(m, m, m)
    qp, qw = ol.quadrature(3, "qfV5")
    n = mesh["xyz"].shape[0]
    t0 = time.perf_counter()
    ci, cj, ca = ol.assemble_coo(mesh, 1, 1, None, LAP3, qp, qw)
    d, v = ol.bc_pairs(mesh, 1, 1, None, ALL6, 1, [0.0])
    ca = ol.bc_matrix_coo(ci, cj, ca, n, d, TGV)
    t1 = time.perf_counter()
    b = ol.bc_rhs(ol.assemble_rhs(mesh, 1, 1, None, n, RHS, qp, qw), d, v, TGV)
    t2 = time.perf_counter()
    x, it, _, _ = ol.cg(n, ci, cj, ca, b, np.zeros(n), eps=EPS, itmax=0, tgv=TGV)
    t3 = time.perf_counter()
    return dict(nt=6 * m ** 3, n=n, nnz=len(ci), tA=t1 - t0, tb=t2 - t1, tcg=t3 - t2, uu=float(x @ x), iters=it, tmesh=0.0,
                wA=t1 - t0, wb=t2 - t1, wcg=t3 - t2, wmesh=0.0)


def cpu_figures(r):
    t = r["tA"] + r["tb"]
    return dict(value=r["nnz"] / t, assembly_nnz_per_s=r["nnz"] / r["tA"],
                spmv_gbs=(16.0 * r["nnz"] + 16.0 * r["n"]) / (r["tcg"] / max(r["iters"], 1)) / 1e9, seconds=t)


def cpu_run(m):
    if os.path.exists(FF_BIN):
        try:
            return run_reference_once(m), "reference"
        except Exception as e:  # e.g. binary not runnable on this host
            sys.stderr.write(f"[bench] reference binary failed ({e}); using the oracle port\n")
    return run_port_once(m), "port"


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # The workload itself (cube(n), n = 128: ~15 s of mesh generation + ~60-110 s of assembly and CG per run on one
    # core) when the run fits a few minutes, else a bounded sample cube(m): a calibration run on cube(40) decides.  The
    # reference has no warm state between statements (every `matrix A = ...` starts from nothing), so warm-up runs are
    # only made on the bounded sample.
    runs = []
    kind = "reference"
    cal, kind = cpu_run(40)
    scale = (args.n / 40.0) ** 3
    est_full = (cal["tA"] + cal["tb"] + cal["tcg"] * (args.n / 40.0) + cal["tmesh"]) * scale * 1.3
    budget = 240.0
    if args.ref_n > 0:
        m, nruns, nwarm = args.ref_n, args.steps, args.warmup
    elif est_full <= budget:
        m, nruns, nwarm = args.n, max(1, min(args.steps, int(budget // est_full))), 0
    else:
        m, nruns, nwarm = 64, max(1, min(args.steps, 3)), 0
    for _ in range(nwarm):
        _, kind = cpu_run(m)
    for _ in range(nruns):
        r, kind = cpu_run(m)
        runs.append(r)
    t = statistics.mean(r["tA"] + r["tb"] for r in runs)
    tcg = statistics.mean(r["tcg"] for r in runs)
    r0 = runs[0]
    value = r0["nnz"] / t
    fig = cpu_figures(dict(r0, tA=statistics.mean(r["tA"] for r in runs), tb=statistics.mean(r["tb"] for r in runs),
                           tcg=statistics.mean(r["tcg"] for r in runs)))
    sample = (f"3-D P1 Poisson cube({m}) ({r0['nt']} tets, nnz {r0['nnz']}, {r0['iters']} CG it): "
              + ("the workload itself" if m == args.n else "same .edp as the workload, bounded size")
              + f", {len(runs)} timed run(s) of {r0['tA'] + r0['tb'] + r0['tcg'] + r0['tmesh']:.0f} s each")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "nnz/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "steps_run": len(runs), "same_size_as_workload": m == args.n,
        "config": {"workload": workload_name(args.n, args.gpus), "sample": sample, "threads": 1,
                   "timed": "clock() deltas around `matrix A = va(Vh,Vh)` and `real[int] b = va(0,Vh)` inside FreeFem++"},
        "assembly": {"nnz_per_s": fig["assembly_nnz_per_s"]}, "spmv": {"gbs": fig["spmv_gbs"], "bytes_model": "16*nnz+16*n (COO)"},
        "cg": {"iters": r0["iters"], "ms": tcg * 1e3, "ms_per_iter": tcg * 1e3 / max(r0["iters"], 1)},
        "cpu_baseline": {"value": value, "unit": "nnz/s", "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def dims_for(n, gpus):
    f = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}.get(gpus)
    if f is None:
        return n, n, n * gpus
    return n * f[0], n * f[1], n * f[2]


def workload_name(n, gpus):
    nx, ny, nz = dims_for(n, gpus)
    return f"3-D P1 Poisson cube({nx},{ny},{nz}): varf -> CSR (symbolic + assembly + Dirichlet) + rhs; then Jacobi-CG eps=1e-6"


def ours(args):
    import numpy as np
    import torch
    import ffcuda

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the ffcuda path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ffcuda.Context(local)
    stream = torch.cuda.Stream()        # a real (non-default) stream: torch events recorded on it bracket the library's kernels
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(ffcuda.Context.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        ctx.comm_init(rank, world, bytes(idt.cpu().tolist()))

    nx, ny, nz = dims_for(args.n, world)
    qp, qw = ffcuda.quadrature(3, 6)
    mesh = ctx.mesh_cube(nx, ny, nz, distributed=(world > 1))
    space = mesh.space(1, 1)
    hbm, hbm_src = peaks()
    TRAFFIC, traffic_src = load_traffic()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def assemble_on(sp):
        pat = sp.symbolic()
        A = pat.matrix()
        A.assemble(LAP3, qp, qw)
        n_loc = pat.info()[0]
        b = ctx.vec(n_loc)
        sp.assemble_linear(b, RHS, qp, qw)
        bc = sp.bc_from_labels(ALL6, 1, [0.0])
        A.apply_bc(bc, TGV)
        b.apply_bc(bc, TGV)
        return pat, A, b

    def timed(fn, steps, warmup):
        out = None
        for _ in range(warmup):
            out = fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        e0.record(stream)
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps, ctx.launch_count() - l0, out

    # ---- what the resident data cost: the first statements on a fresh fespace (wall clock around a synchronised step)
    def wall_step(sp):
        barrier()
        t0 = time.perf_counter()
        out = assemble_on(sp)
        barrier()
        return (time.perf_counter() - t0) * 1e3, out

    cold_sp = mesh.space(1, 1)
    cold1_ms, _ = wall_step(cold_sp)        # node->element incidence + symbolic + thread-per-row kernels
    cold2_ms, _ = wall_step(cold_sp)        # second assembly on the fespace: the row tiles / fans are built here
    cold3_ms, _ = wall_step(cold_sp)        # steady state (wall clock, for comparison with the two above)
    del cold_sp

    # ---- the timed steps: assembly, device resident
    sampler = ClockSampler(local)
    if rank == 0 and not args.no_clocks:
        sampler.start()
    ms_step, launches, (pat, A, b) = timed(lambda: assemble_on(space), args.steps, args.warmup)
    n_loc, nnz_loc = pat.info()
    nv_loc, nt_loc = mesh.info()[1], mesh.info()[2]
    tot = torch.tensor([n_loc, nnz_loc], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tot)
    n_glob, nnz_glob = (int(v) for v in tot.tolist())
    value = nnz_glob / (ms_step * 1e-3)

    # ---- the CG solve that consumes it (own events, same run)
    x = ctx.vec(n_loc)

    def solve():
        x.fill(0.0)
        return A.cg(b, x, eps=EPS, itmax=0, tgv=TGV)

    ms_cg, launches_cg, (iters, conv, _) = timed(solve, max(1, min(3, args.steps)), 1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- kernel-level figures: CUDA events around every launch (the library's profiler), same workload, live
    ctx.prof_enable(True)
    ctx.prof_reset()
    nprof = 2
    for _ in range(nprof):
        pat, A, b = assemble_on(space)
        solve()
    torch.cuda.synchronize()
    prof = {}
    for key in ("inc_", "sym_", "scan_", "asm_rows", "bc_", "rhs_rows", "vec_fill", "cg_spmv_dots", "cg_update_g", "cg_update_xh", "cg_", "halo_", ""):
        ms, cnt = ctx.prof_get(key)
        prof[key] = (ms / nprof, cnt // nprof)
    kernels_ms = {}
    for nm in ("inc_count", "inc_blk_len", "inc_fill", "inc_sort", "inc_stage", "sym_p1_fused", "sym_p1_rows", "sym_p1_cols", "sym_p1_positions", "sym_block_pattern", "sym_compact_cols", "sym_row_count", "sym_row_fill", "sym_diagpos", "scan_tile_sums",
               "scan_tile_offsets", "scan_tiles", "scan_lookback", "asm_rows_p1", "rhs_rows", "bc_mark", "bc_compact", "bc_matrix", "bc_vec", "vec_fill",
               "halo_p2p", "halo_pack", "allreduce_p2p", "spmv_row_blocks", "spmv_sell_len", "spmv_sell_cols", "spmv_sell_vals", "cg_diag_stats", "cg_precond", "cg_init_spmv", "cg_init_h", "cg_spmv_dots", "cg_update_g", "cg_update_xh"):
        ms, cnt = ctx.prof_get(nm)
        if cnt:
            kernels_ms[nm] = [round(ms / nprof, 4), cnt // nprof]
    ctx.prof_enable(False)
    # algorithmic bytes per launch: SpMV 12*nnz + 20*n ; numeric assembly 16*nt + 24*nv + 12*nnz + 4*(n+1)
    B_spmv = 12.0 * nnz_loc + 20.0 * n_loc
    B_asm = 16.0 * nt_loc + 24.0 * nv_loc + 12.0 * nnz_loc + 4.0 * (n_loc + 1)
    t_spmv = prof["cg_spmv_dots"][0] / max(prof["cg_spmv_dots"][1], 1)  # ms per launch
    t_asmk = prof["asm_rows"][0]
    spmv_gbs = B_spmv / (t_spmv * 1e-3) / 1e9
    asm_gbs = B_asm / (t_asmk * 1e-3) / 1e9
    asm_step_kernels = prof["sym_"][0] + prof["scan_"][0] + prof["asm_rows"][0] + prof["bc_"][0] + prof["rhs_rows"][0]

    # standalone y = A x (x_i = sin(i)): the plain SpMV entry point
    xs = ctx.vec_from(np.sin(np.arange(nv_loc, dtype=np.float64)))  # owned + ghost columns
    ys = ctx.vec(n_loc)
    ms_spmv_plain, _, _ = timed(lambda: A.spmv(xs, ys), 50, 5)

    # ---- end to end through host buffers
    if world == 1:
        hm = mesh.download()

        def pin(a):
            t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
            t.numpy()[...] = a
            return t.numpy()

        hm = {k: (pin(v) if isinstance(v, np.ndarray) else v) for k, v in hm.items()}
        h_rp, h_ci = pin(np.zeros(n_loc + 1, np.int32)), pin(np.zeros(nnz_loc, np.int32))
        h_val, h_b, h_x = pin(np.zeros(nnz_loc)), pin(np.zeros(n_loc)), pin(np.zeros(n_loc))
        h2d = sum(hm[k].nbytes for k in ("xyz", "conn", "elab", "bconn", "blab", "belem", "bface"))
        d2h = h_rp.nbytes + h_ci.nbytes + h_val.nbytes + h_b.nbytes
        keep = {}

        def host_step():
            m2 = ctx.mesh_upload(3, hm["xyz"], hm["conn"], hm["elab"], hm["bconn"], hm["blab"], hm["belem"], hm["bface"])
            sp2 = m2.space(1, 1)
            p2 = sp2.symbolic()
            p2.download_async(h_rp, h_ci)   # the pattern travels to the host while the values are assembled
            A2 = p2.matrix()
            A2.assemble(LAP3, qp, qw)
            b2 = ctx.vec(n_loc)
            sp2.assemble_linear(b2, RHS, qp, qw)
            bc2 = sp2.bc_from_labels(ALL6, 1, [0.0])
            A2.apply_bc(bc2, TGV)
            b2.apply_bc(bc2, TGV)
            A2.download(h_val)              # what the plugin hands back to FreeFEM as its MatriceMorse
            b2.download(h_b)
            ctx.sync()                      # ... and the pattern copy
            keep["A"] = A2
            return p2

        ms_e2e, _, _ = timed(host_step, args.steps, args.warmup)

        def host_solve():
            h_x[:] = 0.0
            return keep["A"].cg_host(h_b, h_x, eps=EPS, itmax=0, tgv=TGV)   # u[] = A^-1*b on FreeFEM's host arrays

        ms_e2e_solve, _, (it2, _, _) = timed(host_solve, max(1, min(3, args.steps)), 1)
        e2e = {"value": nnz_glob / (ms_e2e * 1e-3), "unit": "nnz/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "pinned": True,
               "solve_ms": ms_e2e_solve, "solve_iters": it2, "solve_h2d_bytes": int(2 * h_b.nbytes), "solve_d2h_bytes": int(h_x.nbytes),
               "api": "ffcuda_mesh_upload -> space_create -> symbolic -> pattern_download_async -> assemble_bilinear/linear -> "
                      "bc_from_labels/apply -> matrix/vec_download -> ctx_sync ; solve: cg_host"}
    else:
        # N > 1: the partitioned mesh exists only as a device-side generator (no host mesh of 100 M tets is built); end to
        # end = generation + assembly + owned rows of A and b copied to pinned host memory
        h_val = torch.empty(nnz_loc, dtype=torch.float64, pin_memory=True).numpy()
        h_b = torch.empty(n_loc, dtype=torch.float64, pin_memory=True).numpy()

        def host_step():
            m2 = ctx.mesh_cube(nx, ny, nz, distributed=True)
            p2, A2, b2 = assemble_on(m2.space(1, 1))
            A2.download(h_val)
            b2.download(h_b)
            return p2

        ms_e2e, _, _ = timed(host_step, args.steps, args.warmup)
        e2e = {"value": nnz_glob / (ms_e2e * 1e-3), "unit": "nnz/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": int(h_val.nbytes + h_b.nbytes), "pinned": True,
               "api": "ffcuda_mesh_cube_distributed (inputs generated on the device) -> assembly -> matrix/vec_download"}

    # ---- parity of THIS run's result (every rank; size-independent properties + closed forms)
    def allsum(v):
        tt = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt)
        return tt.item()

    def allmax(v):
        tt = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt.item()

    parity = {}
    if not args.no_parity:
        rp_h, ci_h = pat.download()
        if world > 1:
            nown, gid = mesh.local_to_global()
        else:
            gid = np.arange(nv_loc, dtype=np.int64)
        bad_rows = cube_pattern_mismatches(np, nx, ny, nz, gid[:n_loc], rp_h, gid[ci_h])
        A0 = pat.matrix()
        A0.assemble(LAP3, qp, qw)                       # no Dirichlet rows: constants are in the kernel of the stiffness matrix
        ones = ctx.vec_from(np.ones(nv_loc))
        y0 = ctx.vec(n_loc)
        A0.spmv(ones, y0)
        a0max = allmax(np.abs(A0.download()).max())
        rowsum = allmax(np.abs(y0.download()).max()) / a0max
        b0 = ctx.vec(n_loc)
        space.assemble_linear(b0, RHS, qp, qw)
        sum_b = allsum(b0.download().sum())           # f = 1 on the unit cube: the measures of the elements add up to 1
        u_h = x.download()[:n_loc]
        uu = allsum(float(u_h @ u_h))
        del A0, ones, y0, b0
        parity = {"pattern_rows_differing_from_closed_form": int(allsum(bad_rows)), "rows_checked": n_glob,
                  "row_sum_max_over_amax": rowsum, "sum_b_minus_volume": abs(sum_b - 1.0), "cg_iters": iters, "uu": uu}
        ok = parity["pattern_rows_differing_from_closed_form"] == 0 and rowsum <= 1e-12 and abs(sum_b - 1.0) <= 1e-12 and conv == 1
        if world > 1:
            # the same mesh solved on ONE GPU (rank 0, its own context): iteration count and |u|^2 must agree
            if rank == 0:
                c1 = ffcuda.Context(local)
                m1 = c1.mesh_cube(nx, ny, nz)
                s1 = m1.space(1, 1)
                p1 = s1.symbolic()
                A1 = p1.matrix()
                A1.assemble(LAP3, qp, qw)
                n1 = p1.info()[0]
                b1 = c1.vec(n1)
                s1.assemble_linear(b1, RHS, qp, qw)
                bc1 = s1.bc_from_labels(ALL6, 1, [0.0])
                A1.apply_bc(bc1, TGV)
                b1.apply_bc(bc1, TGV)
                x1 = c1.vec(n1)
                it1, conv1, _ = A1.cg(b1, x1, eps=EPS, itmax=0, tgv=TGV)
                u1 = x1.download()
                parity["single_gpu_cg_iters"] = it1
                parity["single_gpu_uu_rel_diff"] = abs(float(u1 @ u1) - uu) / max(abs(uu), 1e-300)
                ok = ok and it1 == iters and parity["single_gpu_uu_rel_diff"] <= 1e-10
                del x1, b1, A1, p1, s1, m1, u1
                c1.close()
            barrier()
        parity["status"] = "ok" if ok else "FAIL"

    # ---- strong scaling: BASELINE.json configs[4] (cube(256), a fixed problem) on the N GPUs of this run
    strong = None
    if args.strong and args.n == 128 and world in (1, 2, 4, 8):
        if world == 8:      # the weak-scaling workload at N = 8 IS cube(256)
            strong = {"workload": workload_name(128, 8), "ms_per_step": ms_step, "value": value, "cg_iters": iters,
                      "cg_ms_per_iter": ms_cg / max(iters, 1), "note": "same run as the weak-scaling line"}
        else:
            sm_ = ctx.mesh_cube(256, 256, 256, distributed=(world > 1))
            ss_ = sm_.space(1, 1)
            ms_s, _, (sp_, sA_, sb_) = timed(lambda: assemble_on(ss_), max(2, min(args.steps, 5)), 3)
            sn_, snnz_ = sp_.info()
            sx_ = ctx.vec(sm_.info()[1])

            def ssolve():
                sx_.fill(0.0)
                return sA_.cg(sb_, sx_, eps=EPS, itmax=0, tgv=TGV)

            ms_scg, _, (sit_, sconv_, _) = timed(ssolve, 1, 1)
            strong = {"workload": workload_name(128, 8), "ms_per_step": ms_s, "value": allsum(snnz_) / (ms_s * 1e-3), "cg_iters": sit_,
                      "cg_converged": sconv_ == 1, "cg_ms_per_iter": ms_scg / max(sit_, 1)}
            del sx_, sb_, sA_, sp_, ss_, sm_

    # ---- CPU baseline beside it (rank 0, N=1 only): the unmodified reference on the workload itself (one run, ~1.5-2 min
    # on one core), or on a bounded sample with --ref-n; and the same script with `load "ffcuda"`: what a script author sees
    cpu = None
    e2e_plugin = None
    if rank == 0 and world == 1 and not args.no_cpu:
        m_cpu = args.ref_n if args.ref_n > 0 else args.n
        r, kind = cpu_run(m_cpu)
        fig = cpu_figures(r)
        cpu = {"value": fig["value"], "unit": "nnz/s", "cores": 1, "kind": kind, "same_size_as_workload": m_cpu == args.n,
               "sample": f"3-D P1 Poisson cube({m_cpu}) ({r['nt']} tets, nnz {r['nnz']}, {r['iters']} CG it, "
                         f"{r['tA'] + r['tb'] + r['tcg']:.1f} s + {r['tmesh']:.1f} s of mesh generation): "
                         + ("the workload itself, one run" if m_cpu == args.n else "same .edp as the workload at a bounded size"),
               "assembly_nnz_per_s": fig["assembly_nnz_per_s"], "spmv_gbs": fig["spmv_gbs"],
               "matrix_s": r["tA"], "rhs_s": r["tb"], "cg_s": r["tcg"], "cg_iters": r["iters"],
               "cg_ms_per_iter": r["tcg"] * 1e3 / max(r["iters"], 1)}
        if kind == "reference" and os.path.exists(os.path.join(ROOT, "freefem-sources_b200", "lib", "ffcuda.so")):
            try:
                del A, b, x, pat
                rp_ = run_reference_once(m_cpu, plugin=True)
                e2e_plugin = {"script": "the cpu_baseline .edp with `load \"ffcuda\"` as its second line, run by the unmodified FreeFem++",
                              "size": f"cube({m_cpu})", "gpu_path_taken": bool(rp_["gpu_path"]),
                              "matrix_s": rp_["wA"], "rhs_s": rp_["wb"], "cg_s": rp_["wcg"], "cg_iters": rp_["iters"],
                              "mesh_s": rp_["wmesh"], "reference_mesh_s": r["wmesh"],
                              "mesh_note": "`mesh3 Th = cube(n,n,n)`: with the plugin the arrays, the adjacency and the boundary links come from "
                                           "the device (FreeFEM spends this statement in BuildAdj); the device copy is adopted by the fespace",
                              "matrix_cpu_s": rp_["tA"], "reference_matrix_s": r["wA"], "reference_rhs_s": r["wb"], "reference_cg_s": r["wcg"],
                              "value": rp_["nnz"] / (rp_["wA"] + rp_["wb"]), "unit": "nnz/s",
                              "speedup_assembly": (r["wA"] + r["wb"]) / max(rp_["wA"] + rp_["wb"], 1e-9),
                              "speedup_cg": r["wcg"] / max(rp_["wcg"], 1e-9),
                              "uu_rel_diff_vs_reference": abs(rp_["uu"] - r["uu"]) / max(abs(r["uu"]), 1e-300),
                              "timed": "wall clock between time stamps the script writes around its three statements (exec date): mesh flattening, "
                                       "upload, kernels, download and the MatriceMorse hand-off are all inside; the plugin's host passes use up "
                                       "to 16 threads, so clock() (CPU time of the process, matrix_cpu_s) is larger than the wall time"}
            except Exception as e:  # the plugin leg must never take the bench line down
                e2e_plugin = {"error": str(e)[-300:]}

    if rank == 0:
        allk = max(prof[""][0], 1e-9)
        line = {
            "metric": METRIC, "value": value, "unit": "nnz/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(args.n, world), "nt": 6 * nx * ny * nz, "ndof": n_glob, "nnz": nnz_glob,
                       "quadrature": "14-point (qforder 6)", "tgv": TGV,
                       "resident": "mesh, element->dof map, its transpose (node->element incidence) and the row tiles / fans of "
                                   "the fespace (Morton clusters of rows with their elements grouped in fans, built once per fespace at "
                                   "its second assembly: `cold` gives what they cost)",
                       "l2": "inputs larger than L2 per GPU: connectivity %.0f MB + CSR %.0f MB vs 126 MB L2"
                             % (16.0 * nt_loc / 1e6, 12.0 * nnz_loc / 1e6),
                       "partition": "z-slabs, one process per GPU" if world > 1 else "single GPU"},
            "phases_ms": {"halo_exchange_in_cg": prof["halo_"][0], "incidence_once_per_fespace": prof["inc_"][0], "symbolic": prof["sym_"][0] + prof["scan_"][0], "numeric_assembly": prof["asm_rows"][0], "bc": prof["bc_"][0],
                          "rhs": prof["rhs_rows"][0], "assembly_step_kernels": asm_step_kernels, "cg_kernels": prof["cg_"][0]},
            "assembly": {"numeric_nnz_per_s": nnz_glob / (t_asmk * 1e-3), "numeric_ms": t_asmk, "algorithmic_GB": B_asm / 1e9,
                         "gbs": asm_gbs * world, "frac_of_hbm": asm_gbs / hbm},
            "cg": {"iters": iters, "converged": conv == 1, "ms": ms_cg, "ms_per_iter": ms_cg / max(iters, 1),
                   "launches": int(launches_cg) // max(1, min(3, args.steps))},
            "spmv": {"gbs": spmv_gbs * world, "ms": t_spmv, "algorithmic_GB": B_spmv / 1e9, "frac_of_hbm": spmv_gbs / hbm,
                     "kernel": "cg_spmv_dots (A*H fused with <G,H>, <H,AH>)",
                     "plain_spmv_gbs": B_spmv / (ms_spmv_plain * 1e-3) / 1e9 * world,
                     "plain_spmv_frac": B_spmv / (ms_spmv_plain * 1e-3) / 1e9 / hbm},
            "roofline": {"bound": "hbm", "kernel": "asm_rows_p1", "achieved": asm_gbs, "peak": hbm, "unit": "GB/s",
                         "frac": asm_gbs / hbm, "traffic": (TRAFFIC.get("asm_rows_p1") or {}).get("traffic") if args.n == 128 else None,
                         "traffic_source": traffic_src,
                         "peak_source": hbm_src, "share_of_step_kernels": t_asmk / max(asm_step_kernels, 1e-9), "per": "GPU"},
            "roofline_spmv": {"bound": "hbm", "kernel": "cg_spmv_dots", "achieved": spmv_gbs, "peak": hbm, "unit": "GB/s",
                              "frac": spmv_gbs / hbm, "traffic": (TRAFFIC.get("cg_spmv_dots") or {}).get("traffic") if args.n == 128 else None,
                              "peak_source": hbm_src, "share_of_cg_kernels": prof["cg_spmv_dots"][0] / max(prof["cg_"][0], 1e-9),
                              "per": "GPU"},
            "kernels_ms_and_launches_per_pass": kernels_ms,
            "cold": {"first_step_ms": cold1_ms, "second_step_ms": cold2_ms, "steady_step_wall_ms": cold3_ms,
                     "tile_build_ms": max(0.0, cold2_ms - cold3_ms), "incidence_and_first_kernels_extra_ms": max(0.0, cold1_ms - cold3_ms),

                     "note": "wall clock around synchronised steps on a fresh fespace of the resident mesh; `value` is the steady state"},
            "parity": parity, "strong_scaling": strong,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if e2e_plugin:
            line["e2e_plugin"] = e2e_plugin
        print(json.dumps(line), flush=True)
    if dist is not None:
        ctx.comm_finalize()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=128, help="cells per edge per GPU (128 = BASELINE.json configs[1])")
    ap.add_argument("--ref-n", type=int, default=0,
                    help="cube size of a bounded CPU sample (0: the workload itself; the reference arm falls back to a sample if a run does not fit)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity object")
    ap.add_argument("--no-strong", dest="strong", action="store_false", help="skip the strong-scaling leg (cube(256) on the N GPUs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample nvidia-smi clocks during the timed region")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
