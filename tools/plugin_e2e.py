"""The bench's .edp with `load "ffcuda"` at cube(n), FFCUDA_VERBOSE=1: where the script-level time goes."""
import os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
with tempfile.TemporaryDirectory() as td:
    edp = os.path.join(td, "b.edp")
    open(edp, "w").write(bench.EDP % ('load "ffcuda"', n))
    env = dict(os.environ, FF_LOADPATH=os.path.join(ROOT, "freefem-sources_b200", "lib"), FFCUDA_VERBOSE="1")
    r = subprocess.run([bench.FF_BIN, "-nw", "-v", "1", edp], capture_output=True, text=True, cwd=td, env=env)
    try:
        st = [float(x) for x in open(os.path.join(td, "ffstamps.txt")).read().split()]
        print("wall: mesh %.3f s, matrix %.3f s, rhs %.3f s, cg %.3f s" % (st[1] - st[0], st[3] - st[2], st[4] - st[3], st[5] - st[4]))
    except Exception as e:
        print("no stamps", e)
    for ln in (r.stdout + r.stderr).splitlines():
        if "ffcuda" in ln or "FFBENCH" in ln or "GC" in ln:
            print(ln[:260])
