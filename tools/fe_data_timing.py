"""Script-level cost of forms whose data are FE functions (heat step with a P1 conductivity: kap*grad u.grad v + u*v/dt,
right-hand side ff*v + uold*v/dt) at cube(n), P1: wall clock of `matrix A = va(Vh,Vh)` and `real[int] b = va(0,Vh)` under
  (1) FreeFEM alone (FFCUDA_DISABLE=1),
  (2) the plugin with the data evaluated by the interpreter at every quadrature node (FFCUDA_NO_FE_DOFS=1, the round-1 path),
  (3) the plugin with the data shipped as dof arrays (ffcuda_fe_table).
usage: python tools/fe_data_timing.py [n]"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FF = os.path.join(ROOT, "oracle", "_ref", "FreeFem++-nw")
EDP = """load "msh3"
load "ffcuda"
int n = %d;
mesh3 Th = cube(n,n,n);
fespace Vh(Th,P1);
Vh kap = 1+x*y+z*z, ff = x*y+sin(z), uold = x*x;
real dt = 0.1;
varf va(u,v) = int3d(Th)(kap*(dx(u)*dx(v)+dy(u)*dy(v)+dz(u)*dz(v)) + u*v/dt) + int3d(Th)(ff*v + uold*v/dt) + on(1,2,3,4,5,6,u=0);
exec("date +%%s.%%N >> ffstamps.txt");
matrix A = va(Vh,Vh,solver=CG,eps=1e-6);
exec("date +%%s.%%N >> ffstamps.txt");
real[int] b = va(0,Vh);
exec("date +%%s.%%N >> ffstamps.txt");
matrix A2 = va(Vh,Vh,solver=CG,eps=1e-6);
exec("date +%%s.%%N >> ffstamps.txt");
real[int] b2 = va(0,Vh);
exec("date +%%s.%%N >> ffstamps.txt");
Vh u; u[] = 0; u[] = A2^-1*b2;
cout.precision(12);
cout << "FEDATA nt " << Th.nt << " nnz " << A.nnz << " uu " << u[]'*u[] << " bb " << b'*b << endl;
"""


def run(n, env_extra):
    with tempfile.TemporaryDirectory() as td:
        with open(os.path.join(td, "t.edp"), "w") as f:
            f.write(EDP % n)
        env = dict(os.environ, FF_LOADPATH=os.path.join(ROOT, "freefem-sources_b200", "lib"), **env_extra)
        r = subprocess.run([FF, "-nw", "-v", "0", "t.edp"], capture_output=True, text=True, cwd=td, env=env)
        st = [float(x) for x in open(os.path.join(td, "ffstamps.txt")).read().split()]
        line = [ln for ln in (r.stdout + r.stderr).splitlines() if ln.startswith("FEDATA")]
        notes = [ln.strip() for ln in (r.stdout + r.stderr).splitlines() if "FE functions" in ln or "coefficient function" in ln]
        return st, (line[0] if line else (r.stdout + r.stderr)[-400:]), notes


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    for tag, env in (("FreeFEM alone", {"FFCUDA_DISABLE": "1"}),
                     ("plugin, interpreter tables", {"FFCUDA_NO_FE_DOFS": "1", "FFCUDA_VERBOSE": "1"}),
                     ("plugin, dof arrays", {"FFCUDA_VERBOSE": "1"})):
        st, line, notes = run(n, env)
        print("%-28s cube(%d): matrix %.3f s, rhs %.3f s | again (fespace on the device): matrix %.3f s, rhs %.3f s" %
              (tag, n, st[1] - st[0], st[2] - st[1], st[3] - st[2], st[4] - st[3]))
        print("    " + line)
        for ln in notes[:4]:
            print("    " + ln)
