"""The drop-in itself: the same .edp script is run by the unmodified FreeFem++ (oracle/_ref/FreeFem++-nw) twice —
once with `load "ffcuda"` doing its work (FFCUDA_STRICT=1: any delegation to FreeFEM's CPU operators is an error) and
once with the plugin disabled (FFCUDA_DISABLE=1) — and the dumps are compared: CSR pattern bit-exact, values / rhs
within 1e-12, CG iteration count and solution as in tests/test_gpu_parity.py.

CPU part (no GPU here): the plugin builds against the reference headers, loads, leaves out-of-scope forms to FreeFEM,
and refuses — loudly — to compute a claimed form without a CUDA device."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FF = os.path.join(ROOT, "oracle", "_ref", "FreeFem++-nw")
LIBDIR = os.path.join(ROOT, "freefem-sources_b200", "lib")
PLUGIN = os.path.join(LIBDIR, "ffcuda.so")

needs_ff = pytest.mark.skipif(not (os.path.exists(FF) and os.path.exists(PLUGIN)),
                              reason="reference binary / plugin not built (needs /root/reference at build time)")

LAP2 = "dx(u)*dx(v)+dy(u)*dy(v)"
LAP3 = "dx(u)*dx(v)+dy(u)*dy(v)+dz(u)*dz(v)"
LAME = ("lambda*(dx(u1)+dy(u2)+dz(u3))*(dx(v1)+dy(v2)+dz(v3))"
        "+2.*mu*(dx(u1)*dx(v1)+dy(u2)*dy(v2)+dz(u3)*dz(v3)"
        "+0.5*(dy(u1)+dx(u2))*(dy(v1)+dx(v2))+0.5*(dz(u1)+dx(u3))*(dz(v1)+dx(v3))"
        "+0.5*(dz(u2)+dy(u3))*(dz(v2)+dy(v3)))")
LAME_PRE = "real E=21.5e4, sigma=0.29; real mu=E/(2*(1+sigma)); real lambda=E*sigma/((1+sigma)*(1-2*sigma));"

DUMP = """
{ A.CSR; ofstream f("A.txt"); f.precision(17); f << A; }
{ ofstream f("b.txt"); f.precision(17); for(int i=0;i<b.n;++i) f << b[i] << endl; }
"""
SOLVE = """
verbosity=1; UU[] = 0; UU[] = A^-1*b; verbosity=0;
{ ofstream f("u.txt"); f.precision(17); for(int i=0;i<UU[].n;++i) f << UU[][i] << endl; }
"""


def script(dim, mesh, fe, bil, lin, bc, pre="", unk="u", tst="v", eps="1e-6", intopt="", tgv=None, sym=False, extra="", solver="CG"):
    mt, integ = ("mesh", "int2d") if dim == 2 else ("mesh3", "int3d")
    u0 = unk.strip("[]").split(",")[0]
    s = f'load "msh3"\nload "ffcuda"\n{pre}\n{mt} Th = {mesh};\nfespace Vh(Th,{fe});\n'
    s += f"varf va({unk},{tst}) = {integ}(Th{intopt})({bil}) + {integ}(Th{intopt})({lin}){extra}{('+' + bc) if bc else ''};\n"
    tg = "" if tgv is None else f",tgv={tgv}"
    sy = ",sym=1" if sym else ""
    s += f"matrix A = va(Vh,Vh,solver={solver},eps={eps}{tg}{sy});\nreal[int] b = va(0,Vh{tg});\n" + DUMP
    s += f"Vh {unk};\n" + SOLVE.replace("UU", u0)
    return s


CASES = {
    "poisson3d_p1": script(3, "cube(7,6,8)", "P1", LAP3, "1.*v", "on(1,2,3,4,5,6,u=0)"),
    "laplace2d_p1": script(2, "square(23,17)", "P1", LAP2, "1.*v", "on(1,2,3,4,u=0)"),
    "laplace2d_p1_warp_labels": script(2, "square(9,7,[x+0.2*y*y,y*(1+0.3*x)])", "P1", LAP2 + "+2.*u*v", "3.*v+dx(v)",
                                       "on(1,u=1)+on(3,u=2)"),
    "poisson3d_p2": script(3, "cube(3,4,3)", "P2", LAP3, "1.*v", "on(1,2,3,4,5,6,u=0)", eps="1e-14"),
    "laplace2d_p2": script(2, "square(6,5)", "P2", LAP2, "1.*v", "on(1,2,3,4,u=0)", eps="1e-14"),
    "heat3d_p1": script(3, "cube(5,5,5)", "P1", "u*v/dt+" + LAP3, "1.*v", "on(1,2,3,4,5,6,u=0)", pre="real dt=0.01;"),
    "lame3d_p2": script(3, "cube(2,3,2)", "[P2,P2,P2]", LAME, "-0.05*v3", "on(1,u1=0,u2=0,u3=0)", pre=LAME_PRE,
                        unk="[u1,u2,u3]", tst="[v1,v2,v3]", eps="1e-14"),
    # two components in 2-D: the elasticity of examples/tutorial/beam.edp on a structured beam, and its P2 version
    "beam2d_p1_vector": script(2, "square(20,5,[10*x,2*y])", "[P1,P1]",
                               "lambda*(dx(u1)+dy(u2))*(dx(v1)+dy(v2))+2.*mu*(dx(u1)*dx(v1)+dy(u2)*dy(v2)+0.5*(dy(u1)+dx(u2))*(dy(v1)+dx(v2)))",
                               "-0.05*v2", "on(2,4,u1=0,u2=0)", pre="real E=21.5, sigma=0.29; real mu=E/(2*(1+sigma)); real lambda=E*sigma/((1+sigma)*(1-2*sigma));",
                               unk="[u1,u2]", tst="[v1,v2]", eps="1e-14"),
    "beam2d_p2_vector": script(2, "square(8,3,[10*x,2*y])", "[P2,P2]",
                               "lambda*(dx(u1)+dy(u2))*(dx(v1)+dy(v2))+2.*mu*(dx(u1)*dx(v1)+dy(u2)*dy(v2)+0.5*(dy(u1)+dx(u2))*(dy(v1)+dx(v2)))",
                               "-0.05*v2", "on(2,4,u1=0,u2=0)", pre="real E=21.5, sigma=0.29; real mu=E/(2*(1+sigma)); real lambda=E*sigma/((1+sigma)*(1-2*sigma));",
                               unk="[u1,u2]", tst="[v1,v2]", eps="1e-14"),
    # exact elimination of the Dirichlet rows and columns (tgv = -2, HashMatrix::SetBC)
    "poisson3d_p1_tgvm2": script(3, "cube(5,6,4)", "P1", LAP3, "1.*v", "on(1,2,3,4,5,6,u=0)", tgv=-2),
    "laplace2d_p2_tgvm2": script(2, "square(6,5)", "P2", LAP2, "1.*v", "on(1,2,3,4,u=0)", eps="1e-14", tgv=-2),
    # half storage (sym=1): FreeFEM keeps the lower triangle, the solve runs on the full device matrix
    "poisson3d_p1_sym": script(3, "cube(6,5,4)", "P1", LAP3, "1.*v", "on(1,2,3,4,5,6,u=0)", sym=True),
    "lame3d_p2_sym": script(3, "cube(2,2,2)", "[P2,P2,P2]", LAME, "-0.05*v3", "on(1,u1=0,u2=0,u3=0)", pre=LAME_PRE,
                            unk="[u1,u2,u3]", tst="[v1,v2,v3]", eps="1e-14", sym=True),
    # boundary integrals: Neumann data and Robin terms (int2d on a mesh3, int1d on a mesh), 1-D rule chosen by the user
    "poisson3d_p1_robin": script(3, "cube(5,4,6,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", "P1", LAP3, "1.*v", "on(1,u=0)",
                                 extra="+int2d(Th,2,3)(1.5*u*v)+int2d(Th,2,3)(2.5*v)+int2d(Th,6)(-1.*v)"),
    "laplace2d_p2_robin": script(2, "square(6,5,[x+0.2*y*y,y*(1+0.3*x)])", "P2", LAP2, "1.*v", "on(4,u=0)", eps="1e-14",
                                 extra="+int1d(Th,2,3)(0.7*u*v)+int1d(Th,2,qfe=qf2pE)(1.5*v)"),
    "lame3d_p1_traction": script(3, "cube(3,4,3)", "[P1,P1,P1]", LAME, "-0.05*v3", "on(1,u1=0,u2=0,u3=0)", pre=LAME_PRE,
                                 unk="[u1,u2,u3]", tst="[v1,v2,v3]", eps="1e-14",
                                 extra="+int2d(Th,3)(1e3*(u1*v1+u2*v2+u3*v3))+int2d(Th,2)(0.3*v1-0.2*v3)"),
    # non-symmetric forms with solver=GMRES (fgmres; the second case restarts every 12 iterations)
    "convdiff3d_p1_gmres": script(3, "cube(6,5,7)", "P1", LAP3 + "+8.*dx(u)*v+3.*dy(u)*v-2.*dz(u)*v", "1.*v",
                                  "on(1,2,3,4,5,6,u=0)", solver="GMRES"),
    "convdiff2d_p2_gmres": script(2, "square(7,6)", "P2", LAP2 + "+5.*dx(u)*v+u*v", "1.*v", "on(1,3,u=0)", eps="1e-14",
                                  solver="GMRES,dimKrylov=40"),
    # right-hand side data depending on the mesh point (evaluated at the quadrature nodes by FreeFEM's evaluator, integrated on the GPU)
    "poisson3d_p1_fxyz": script(3, "cube(5,4,6,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", "P1", LAP3, "(x*y+sin(z))*v+2.*v",
                                "on(1,2,3,4,5,6,u=0)", eps="1e-14"),
    "lame3d_p2_fvec": script(3, "cube(2,3,2)", "[P2,P2,P2]", LAME, "x*v1-0.05*(1+y)*v3", "on(1,u1=0,u2=0,u3=0)", pre=LAME_PRE,
                             unk="[u1,u2,u3]", tst="[v1,v2,v3]", eps="1e-14"),
    # bilinear coefficients depending on the mesh point on P1 spaces (moments of the coefficient on every element)
    "diff3d_p1_kappa": script(3, "cube(5,4,6,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", "P1", "(1+x*y+z*z)*(" + LAP3 + ")+2.*u*v", "1.*v",
                              "on(1,2,u=0)", eps="1e-14"),
    "lame3d_p1_evar": script(3, "cube(3,4,3)", "[P1,P1,P1]", "(1+x)*(" + LAME + ")+0.5*(1+y*y)*(u1*v1+u2*v2+u3*v3)", "-0.05*v3",
                             "on(1,u1=0,u2=0,u3=0)", pre=LAME_PRE, unk="[u1,u2,u3]", tst="[v1,v2,v3]", eps="1e-14"),
    # Dirichlet data depending on the mesh point (evaluated at the boundary nodes by FreeFEM's evaluator, boundary element by
    # boundary element as AssembleBC does)
    "poisson3d_p1_dirichlet_g": script(3, "cube(5,4,6,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", "P1", LAP3, "1.*v",
                                       "on(1,2,3,u=x*y+z)+on(4,5,6,u=1+N.x)", eps="1e-14"),
    "laplace2d_p2_dirichlet_g": script(2, "square(6,5,[x+0.2*y*y,y*(1+0.3*x)])", "P2", LAP2, "1.*v", "on(1,3,u=sin(x)+y)+on(2,u=2.)",
                                       eps="1e-14", tgv=-2),
    "lame3d_p2_dirichlet_g": script(3, "cube(2,3,2)", "[P2,P2,P2]", LAME, "-0.05*v3", "on(1,u1=0.01*x,u2=0,u3=-0.02*z)", pre=LAME_PRE,
                                    unk="[u1,u2,u3]", tst="[v1,v2,v3]", eps="1e-14"),
    # boundary integrals whose data depend on the mesh point (Neumann g(x) v, Robin alpha(x) u v), P1 and P2
    "poisson3d_p2_bnd_g": script(3, "cube(3,3,4,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", "P2", LAP3, "1.*v", "on(1,u=0)", eps="1e-14",
                                 extra="+int2d(Th,2,3)((1+x*z)*u*v)+int2d(Th,2,3)((y+sin(z))*v)+int2d(Th,6)(0.5*v-N.z*x*v)"),
    "laplace2d_p1_bnd_g": script(2, "square(9,7,[x+0.2*y*y,y*(1+0.3*x)])", "P1", LAP2, "1.*v", "on(4,u=0)", eps="1e-14",
                                 extra="+int1d(Th,2,3,qfe=qf3pE)((1+x*y)*u*v)+int1d(Th,2,qfe=qf1pElump)(exp(y)*v)"),
    "lame3d_p1_bnd_g": script(3, "cube(3,4,3)", "[P1,P1,P1]", LAME, "-0.05*v3", "on(1,u1=0,u2=0,u3=0)", pre=LAME_PRE,
                              unk="[u1,u2,u3]", tst="[v1,v2,v3]", eps="1e-14",
                              extra="+int2d(Th,3)(1e3*(1+x)*(u1*v1+u2*v2+u3*v3))+int2d(Th,2)(0.3*z*v1-0.2*(1+y)*v3)"),
    "diff3d_p2_kappa": script(3, "cube(3,3,4,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", "P2", "(1+x*y+z*z)*(" + LAP3 + ")+2.*u*v", "1.*v",
                              "on(1,2,u=0)", eps="1e-14"),
    "lame3d_p2_evar": script(3, "cube(2,3,2)", "[P2,P2,P2]", "(1+x)*(" + LAME + ")+0.5*(1+y*y)*(u1*v1+u2*v2+u3*v3)", "-0.05*v3",
                             "on(1,u1=0,u2=0,u3=0)", pre=LAME_PRE, unk="[u1,u2,u3]", tst="[v1,v2,v3]", eps="1e-14"),
    "resid3d_p2_grad": script(3, "cube(3,3,4,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", "P2", LAP3, "(1+z)*dz(v)+x*v-y*z*dx(v)+2.*dy(v)",
                              "on(1,2,u=0)", eps="1e-14"),
    "mass3d_lumped": script(3, "cube(3,3,3)", "P1", "u*v+0.1*(" + LAP3 + ")", "1.*v", "on(1,u=0)", intopt=",qfV=qfV1lump"),
    # boundary integrals with derivatives of the unknown / of the test function (constant coefficients and the unit normal):
    # every node of the element behind the face is reached (Element_Op's border branch, problem.cpp:6518-6560)
    "poisson3d_p1_bnd_grad": script(3, "cube(5,4,6,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", "P1", LAP3, "1.*v", "on(1,u=0)", solver="GMRES",
                                    extra="+int2d(Th,2,3)(0.3*dx(u)*v+0.2*u*dy(v)+0.1*dz(u)*dx(v)+0.25*u*v)"
                                          "+int2d(Th,4)(-0.5*(N.x*dx(u)+N.y*dy(u)+N.z*dz(u))*v)+int2d(Th,6)(0.3*dz(v)-1.*v)"),
    "laplace2d_p2_bnd_grad": script(2, "square(6,5,[x+0.2*y*y,y*(1+0.3*x)])", "P2", LAP2, "1.*v", "on(4,u=0)", solver="GMRES", eps="1e-14",
                                    extra="+int1d(Th,2,3)(0.3*dx(u)*v+0.1*dy(u)*dy(v))+int1d(Th,2)(1.5*dx(v)+0.5*v)"),
    "lame3d_p1_bnd_grad": script(3, "cube(3,4,3)", "[P1,P1,P1]", LAME, "-0.05*v3", "on(1,4,5,u1=0,u2=0,u3=0)", pre=LAME_PRE,
                                 unk="[u1,u2,u3]", tst="[v1,v2,v3]", solver="GMRES", eps="1e-14",
                                 extra="+int2d(Th,3)(1e3*(dx(u1)*v2+u3*dz(v1)+u2*v2))+int2d(Th,2)(0.3*dy(v1)-0.2*v3)"),
}

# re-assembly in a time loop (configs[3] shape, idp/Heat3d.idp): `A = va(Vh,Vh)` on an existing matrix, rhs from the previous
# solution through a matrix-vector product, solve; the final state is dumped
HEAT_LOOP = f"""load "msh3"
load "ffcuda"
mesh3 Th = cube(5,4,5);
fespace Vh(Th,P1);
real dt = 0.01;
varf va(u,v) = int3d(Th)(u*v/dt+{LAP3}) + on(1,2,3,4,5,6,u=0);
varf vm(u,v) = int3d(Th)(u*v/dt);
varf vf(u,v) = int3d(Th)(1.*v) + on(1,2,3,4,5,6,u=0);
matrix A = va(Vh,Vh,solver=CG,eps=1e-10);
matrix M = vm(Vh,Vh);
real[int] f = vf(0,Vh);
Vh u; u[] = 0;
real[int] b(Vh.ndof);
for (int it = 0; it < 3; ++it) {{
  A = va(Vh,Vh,solver=CG,eps=1e-10);
  b = M*u[]; b += f;
  u[] = A^-1*b;
}}
{DUMP}
{{ ofstream g("u.txt"); g.precision(17); for(int i=0;i<u[].n;++i) g << u[][i] << endl; }}
"""


def run_ff(src, env_extra, want_fail=False):
    with tempfile.TemporaryDirectory() as td:
        with open(os.path.join(td, "case.edp"), "w") as f:
            f.write(src)
        env = dict(os.environ, FF_LOADPATH=LIBDIR, **env_extra)
        r = subprocess.run([FF, "-nw", "-v", "1", "case.edp"], capture_output=True, text=True, cwd=td, env=env, timeout=600)
        out = r.stdout + r.stderr
        if want_fail:
            return r.returncode, out, None
        assert r.returncode == 0, out[-3000:]
        res = {"out": out}
        with open(os.path.join(td, "A.txt")) as f:
            lines = [ln for ln in f if not ln.startswith("#")]
        hdr = lines[0].split()
        n, nnz = int(hdr[0]), int(hdr[3])
        a = np.array(" ".join(lines[1:]).split(), dtype=np.float64).reshape(-1, 3)
        assert a.shape[0] == nnz
        res.update(n=n, I=a[:, 0].astype(np.int64), J=a[:, 1].astype(np.int64), V=a[:, 2].copy())
        res["b"] = np.loadtxt(os.path.join(td, "b.txt"), ndmin=1)
        if os.path.exists(os.path.join(td, "u.txt")):
            res["u"] = np.loadtxt(os.path.join(td, "u.txt"), ndmin=1)
        res["iters"] = [int(x) for x in re.findall(r"(?:GC[^\n]*?after|fgmres[^\n]*?converged in)\s+(\d+)", out)]
        return r.returncode, out, res


def compare(gpu, cpu, tight):
    assert gpu["n"] == cpu["n"]
    assert np.array_equal(gpu["I"], cpu["I"]) and np.array_equal(gpu["J"], cpu["J"])          # bit-exact pattern
    big = np.abs(cpu["V"]) > 1e29
    assert np.array_equal(np.abs(gpu["V"]) > 1e29, big)
    scale = np.abs(cpu["V"][~big]).max()
    assert np.max(np.abs(gpu["V"] - cpu["V"])[~big]) <= 1e-12 * scale
    bbig = np.abs(cpu["b"]) > 1e20
    assert np.array_equal(np.abs(gpu["b"]) > 1e20, bbig)
    if (~bbig).any():
        assert np.max(np.abs(gpu["b"] - cpu["b"])[~bbig]) <= 1e-12 * max(np.abs(cpu["b"][~bbig]).max(), 1e-300)
    assert np.allclose(gpu["b"][bbig], cpu["b"][bbig], rtol=1e-15, atol=0)
    umax = np.abs(cpu["u"]).max()
    if tight:   # both converged to round-off
        assert np.max(np.abs(gpu["u"] - cpu["u"])) <= 1e-12 * umax
    else:       # the reference's own stopping point: same iteration count, same iterate
        assert gpu["iters"] == cpu["iters"]
        assert np.max(np.abs(gpu["u"] - cpu["u"])) <= 1e-12 * umax


@needs_ff
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_plugin_matches_freefem(name):
    src = CASES[name]
    _, out, gpu = run_ff(src, {"FFCUDA_STRICT": "1", "FFCUDA_VERBOSE": "1"})
    tag = "fgmres (ffcuda)" if "solver=GMRES" in src else "GC (ffcuda)"
    assert "assembled on the GPU" in out and tag in out                  # the GPU path ran, nothing was delegated
    _, out_cpu, cpu = run_ff(src, {"FFCUDA_DISABLE": "1"})
    assert "assembled on the GPU" not in out_cpu and "(ffcuda)" not in out_cpu and "ffcuda disabled" in out_cpu
    compare(gpu, cpu, tight="eps=1e-14" in src)


@needs_ff
@pytest.mark.gpu
def test_plugin_heat_time_loop():
    _, out, gpu = run_ff(HEAT_LOOP, {"FFCUDA_VERBOSE": "1"})
    assert out.count("assembled on the GPU") >= 6
    _, _, cpu = run_ff(HEAT_LOOP, {"FFCUDA_DISABLE": "1"})
    assert np.array_equal(gpu["I"], cpu["I"]) and np.array_equal(gpu["J"], cpu["J"])
    big = np.abs(cpu["V"]) > 1e29
    assert np.max(np.abs(gpu["V"] - cpu["V"])[~big]) <= 1e-12 * np.abs(cpu["V"][~big]).max()
    assert np.max(np.abs(gpu["u"] - cpu["u"])) <= 1e-9 * np.abs(cpu["u"]).max()


OUT_OF_SCOPE = """load "msh3"
load "ffcuda"
mesh3 Th = cube(3,3,3);
fespace Vh(Th,P1);
fespace Vh2(Th,P2);
varf vb(u,v) = int3d(Th)(x*dx(u)*dx(v)+u*v) + on(1,u=0);
matrix B = vb(Vh2,Vh);
fespace Wh(Th,P0);
varf vc(u,v) = int3d(Th)(u*v);
matrix C = vc(Wh,Wh);
varf vs(u,v) = int2d(Th,2)(u*v) + int2d(Th,2)(x*dx(v));
matrix S = vs(Vh,Vh);
real[int] r = vs(0,Vh);
cout << "NNZ " << B.nnz << " " << C.nnz << " " << S.nnz << endl;
"""

CLAIMED = """load "msh3"
load "ffcuda"
mesh3 Th = cube(3,3,3);
fespace Vh(Th,P1);
varf va(u,v) = int3d(Th)(dx(u)*dx(v)+dy(u)*dy(v)+dz(u)*dz(v)) + on(1,u=0);
matrix A = va(Vh,Vh);
cout << "NNZ " << A.nnz << endl;
"""


@needs_ff
def test_plugin_loads_and_leaves_out_of_scope_forms_to_freefem():
    """different unknown and test spaces, non-Lagrange element, boundary integral without a volume integral: not claimed, FreeFEM's own operators run
    (no GPU needed), and the plugin says so."""
    rc, out, _ = run_ff(OUT_OF_SCOPE, {}, want_fail=True)
    assert rc == 0, out[-2000:]
    assert re.search(r"^NNZ 2314 162 \d+", out, re.M)
    assert out.count("left to FreeFEM") >= 4
    rc, out, _ = run_ff(OUT_OF_SCOPE, {"FFCUDA_STRICT": "1"}, want_fail=True)
    assert rc != 0 and "FFCUDA_STRICT" in out


RECT_FORMS = """load "msh3"
load "ffcuda"
mesh3 Th = cube(3,3,3);
fespace Uh(Th,[P2,P2,P2]);
fespace Ph(Th,P1);
varf vb([u1,u2,u3],[q]) = int3d(Th)(-(dx(u1)+dy(u2)+dz(u3))*q) + int3d(Th)(0.5*u2*dx(q));
fespace Wh(Th,P2);
varf von(u,q) = int3d(Th)(dx(u)*q) + on(1,u=0);
varf vx([u1,u2,u3],[q]) = int3d(Th)(x*dx(u1)*q);
varf vbd([u1,u2,u3],[q]) = int3d(Th)(dx(u1)*q) + int2d(Th,2)(u1*q);
varf vq([u1,u2,u3],[q]) = int3d(Th)(dx(u1)*q) + int3d(Th,qfV=qfV1)(u2*q);
varf vr([u1,u2,u3],[q]) = int3d(Th,7)(dx(u1)*q);
varf vr0([u1,u2,u3],[q]) = int3d(Th,0)(dx(u1)*q);
try { matrix B = vb(Uh,Ph); cout << "B " << B.n << " " << B.m << " " << B.nnz << endl; } catch(...) { cout << "B: no device" << endl; }
matrix Bon = von(Wh,Ph);
matrix Bx = vx(Uh,Ph);
matrix Bbd = vbd(Uh,Ph);
matrix Bq = vq(Uh,Ph);
matrix Br = vr(Uh,Ph);
try { matrix Br0 = vr0(Uh,Ph); cout << "Br0 " << Br0.nnz << endl; } catch(...) { cout << "Br0: no device" << endl; }
cout << "NNZ " << Bon.nnz << " " << Bx.nnz << " " << Bbd.nnz << " " << Bq.nnz << endl;
fespace Xh(Th,[P2,P2,P2,P1]);
varf vs([u1,u2,u3,p],[v1,v2,v3,q]) = int3d(Th)(dx(u1)*dx(v1)+dy(u2)*dy(v2)+dz(u3)*dz(v3)-p*(dx(v1)+dy(v2)+dz(v3))-(dx(u1)+dy(u2)+dz(u3))*q)
   + on(1,2,u1=0,u2=0,u3=0);
try { matrix S = vs(Xh,Xh); cout << "S " << S.n << " " << S.nnz << endl; } catch(...) { cout << "S: no device" << endl; }
"""


@needs_ff
def test_plugin_rectangular_forms_are_recognised_without_a_device():
    """`matrix B = vb(Uh,Vh)` with two different spaces: left to FreeFEM unless FFCUDA_RECT=1; with it, volume integrals with
    constant coefficients and one rule are claimed (FFCUDA_EXPLAIN prints what was read before any device call), forms with
    on(...), mesh-dependent coefficients, boundary integrals or two rules are left to FreeFEM.  Checked where there is no GPU
    (the claimed statement then fails loudly: no CPU fallback); on a GPU box tests/test_zz_gpu_rect.py runs the statement."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    rc, out, _ = run_ff(RECT_FORMS, {}, want_fail=True)
    assert rc == 0 and re.search(r"^B 64 1029 ", out, re.M) and out.count("FFCUDA_RECT=1 takes such forms to the device") == 7
    assert re.search(r"^S 1093 \d+", out, re.M) and "boundary condition on some components only" in out  # (left to FreeFEM)
    rc, out, _ = run_ff(RECT_FORMS, {"FFCUDA_RECT": "1", "FFCUDA_EXPLAIN": "1"}, want_fail=True)
    assert rc == 0, out[-2000:]
    assert "rectangular matrix 64 x 1029: 4 term(s), 14 quadrature point(s), all regions" in out
    assert "B: no device" in out and not re.search(r"^B 64", out, re.M)
    assert "1 region label(s)" in out and "Br0: no device" in out  # (region 0 is the whole cube: claimed)
    for why in ("on(...) in a form with two different spaces", "coefficient depends on the mesh point",
                "boundary integral in a form with two different spaces", "different quadrature rules or regions",
                "do not visit every element (sub-pattern)"):
        assert why in out, why
    assert re.search(r"^NNZ \d+ \d+ \d+ \d+", out, re.M)
    # a mixed-order product space in one fespace: claimed as scalar blocks, on(...) left to FreeFEM's AssembleBC
    assert "mixed-order space [P2,P2,P2,P1], 1093 dofs: 16 scalar blocks, 9 term(s), 14 quadrature point(s), on(...) by FreeFEM's AssembleBC" in out
    assert "S: no device" in out and not re.search(r"^S 1093", out, re.M)


@needs_ff
def test_plugin_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    rc, out, _ = run_ff(CLAIMED, {}, want_fail=True)
    assert rc != 0
    assert "no CPU fallback" in out and not re.search(r"^NNZ \d", out, re.M)


# ---------------------------------------------------------------------------------------------------------------
# FE functions as data of a form: values / derivatives of P0 / P1 / P2 functions living on the mesh of
# the form, alone or in affine combinations with mesh-independent factors (uold/dt, -f, 2 + kappa), go to the device as dof
# arrays; the tables at the quadrature nodes are formed there (ffcuda_fe_table), no interpreter call per node
# ---------------------------------------------------------------------------------------------------------------
def fe_script(dim, mesh, fe, decl, bil, lin, bc, extra="", pre="", unk="u", tst="v"):
    mt, integ = ("mesh", "int2d") if dim == 2 else ("mesh3", "int3d")
    u0 = unk.strip("[]").split(",")[0]
    s = f'load "msh3"\nload "ffcuda"\n{pre}\n{mt} Th = {mesh};\nfespace Vh(Th,{fe});\n{decl}\n'
    s += f"varf va({unk},{tst}) = {integ}(Th)({bil}) + {integ}(Th)({lin}){extra}+{bc};\n"
    s += "matrix A = va(Vh,Vh,solver=CG,eps=1e-14);\nreal[int] b = va(0,Vh);\n" + DUMP
    s += f"Vh {unk};\n" + SOLVE.replace("UU", u0)
    return s


WARP3 = "[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)]"
FE_CASES = {
    # name: (script, bilinear terms on the dof-array path, linear terms on it, coefficient functions left to the interpreter)
    "fe3d_p1": (fe_script(3, f"cube(5,4,6,{WARP3})", "P1",
                          "fespace W1(Th,P1); W1 kap=1+x*y+z*z, ff=x*y+sin(z), uk=x*x+y*z; fespace W0(Th,P0); W0 rho=1+x+2*z; real dt=0.1;",
                          "kap*(" + LAP3 + ")+rho*u*v/dt+2.*u*v", "ff*v+dx(uk)*dx(v)+dy(uk)*dy(v)+dz(uk)*dz(v)+uk*v/dt", "on(1,u=0)",
                          extra="+int2d(Th,2,3)(kap*u*v)-int2d(Th,2,3)(ff*v)"), True, True, 0),
    "fe3d_p2": (fe_script(3, f"cube(3,3,4,{WARP3})", "P2",
                          "fespace W1(Th,P1); W1 kap=1+x*y+z*z; fespace W2(Th,P2); W2 uk=x*x+y*z+sin(x*z), m2=2+x*y*z;",
                          "kap*(" + LAP3 + ")+m2*u*v", "uk*v+dx(uk)*dx(v)+dz(uk)*dy(v)", "on(1,2,u=0)", extra="+int2d(Th,6)(uk*v)"),
                True, True, 0),
    "fe2d_p1": (fe_script(2, "square(9,7,[x+0.2*y*y,y*(1+0.3*x)])", "P1",
                          "fespace W2(Th,P2); W2 kap=1+sin(x)*y, uk=x*x*y-y*y; fespace W1(Th,P1); W1 ff=exp(x)*y;",
                          "kap*(" + LAP2 + ")+u*v", "ff*v+dx(uk)*dx(v)+dy(uk)*dy(v)", "on(4,u=0)",
                          extra="+int1d(Th,2,3)(ff*u*v)+int1d(Th,2)(dy(uk)*v)"), True, True, 0),
    "fe3d_lame": (fe_script(3, "cube(3,4,3)", "[P1,P1,P1]",
                            "fespace Wv(Th,[P1,P1,P1]); Wv [f1,f2,f3]=[x*y,sin(z),-0.05*(1+y)]; fespace W1(Th,P1); W1 ee=1+x;",
                            "ee*(" + LAME + ")", "f1*v1+f3*v3+dx(f2)*v2", "on(1,u1=0,u2=0,u3=0)", pre=LAME_PRE, unk="[u1,u2,u3]",
                            tst="[v1,v2,v3]"), True, True, 0),
    # products of FE functions and factors depending on x are NOT affine in the FE data: those terms stay with the interpreter,
    # the others of the same statement still go as dof arrays
    "fe3d_mixed": (fe_script(3, f"cube(4,4,4,{WARP3})", "P1", "fespace W1(Th,P1); W1 kap=1+x*y+z*z, ff=x*y+sin(z);",
                             "(1+x)*kap*dx(u)*dx(v)+kap*ff*dy(u)*dy(v)+kap*dz(u)*dz(v)+u*v", "ff*ff*v-kap*dx(v)", "on(1,u=0)"),
                   True, False, 2),
}


@needs_ff
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(FE_CASES))
def test_plugin_fe_data_go_as_dof_arrays(name):
    src, bil_fe, lin_fe, ninterp = FE_CASES[name]
    _, out, gpu = run_ff(src, {"FFCUDA_STRICT": "1", "FFCUDA_VERBOSE": "1"})
    assert "assembled on the GPU" in out and "GC (ffcuda)" in out
    assert ("that are FE functions" in out) == bil_fe and ("whose data are FE functions" in out) == lin_fe
    m = re.search(r"(\d+) coefficient function\(s\) depending on the mesh point", out)
    assert (int(m.group(1)) if m else 0) == ninterp
    _, _, cpu = run_ff(src, {"FFCUDA_DISABLE": "1"})
    compare(gpu, cpu, tight=True)
    if name == "fe3d_p1":   # the recognition switched off: the same statement through the interpreter tables, same result
        _, out2, gpu2 = run_ff(src, {"FFCUDA_STRICT": "1", "FFCUDA_VERBOSE": "1", "FFCUDA_NO_FE_DOFS": "1"})
        assert "FE functions" not in out2 and "coefficient function(s) depending on the mesh point" in out2
        compare(gpu2, cpu, tight=True)


EXPLAIN = """load "msh3"
load "ffcuda"
mesh3 Th = cube(3,3,3,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)]);
fespace Vh(Th,P1);
fespace W1(Th,P1); W1 kap=1+x*y+z*z, ff=x*y+sin(z), uk=x*x+y*z;
fespace W0(Th,P0); W0 rho=1+x+2*z;
fespace W2(Th,P2); W2 m2=2+x*y*z;
fespace Wv(Th,[P1,P1,P1]); Wv [f1,f2,f3]=[x*y,sin(z),x*z];
mesh3 Th2 = cube(2,2,2); fespace Z1(Th2,P1); Z1 other=x;
real dt = 0.1;
varf va(u,v) = int3d(Th)(2*u*v + kap*u*v + rho*u*v/dt + (1+x)*kap*dx(u)*dx(v) + kap*ff*dy(u)*dy(v) - m2*dz(u)*dz(v) + other*dx(u)*v)
   - int3d(Th)(ff*v + f2*v + uk*v/dt - 3*rho*v + 2*v + dz(f3)*dx(v))
   + int2d(Th,2,3)(kap*u*v/4) - int2d(Th,2,3)(ff*v) + on(1,u=0);
STATEMENT
"""


EXPLAIN_TRICKY = """load "msh3"
load "ffcuda"
mesh3 Th = cube(12,12,12);
fespace Vh(Th,P1);
Vh ff=x*y+sin(z), zero=0, bump=0;
bump[][777] = 1.;
fespace W0(Th,P0); W0 chi = (x>0.9)*(y>0.9)*(z>0.9);
real dt = 0.1;
varf va(u,v) = int3d(Th)(max(ff,0.9)*v + sin(ff)*dx(v) + x*ff*dy(v) + 2*bump*dz(v))
  + int3d(Th,qforder=2)(3*zero*v + abs(ff)*dx(v) + (1+chi)*ff*dy(v) + (ff>1.2 ? 2.:1.)*ff*dz(v))
  + int3d(Th,qforder=3)(2*chi*v - ff*dx(v)/dt + (ff+bump)*dy(v));
real[int] b = va(0,Vh);
"""


EXPLAIN_LOOP = """load "msh3"
load "ffcuda"
mesh3 Th = cube(4,4,4);
fespace Vh(Th,P1);
Vh ff=x*y+sin(z), uold=x;
real dt = 0.1;
varf va(u,v) = int3d(Th)(ff*v/dt + 2*uold*dx(v));
for (int it = 0; it < 3; ++it) {
  dt = 0.1*(it+1);
  try { real[int] b = va(0,Vh); } catch(...) { cout << "no device" << endl; }
}
"""


@needs_ff
def test_plugin_recognises_fe_data_without_a_device():
    """the FreeFEM side of the dof-array path, checked where there is no GPU: FFCUDA_EXPLAIN=1 prints, before any device call,
    how every term with mesh-dependent data will be treated (fe_affine: sub-expressions listed by E_F0::Optimize, factors
    fitted on the interpreter's own values and verified at every sampled node)"""
    rc, out, _ = run_ff(EXPLAIN.replace("STATEMENT", "matrix A = va(Vh,Vh);"), {"FFCUDA_EXPLAIN": "1"}, want_fail=True)
    ex = dict(re.findall(r"ffcuda explain: (.*? item \d+ term \d+): (.*)", out))
    assert len(ex) == 6, out[-3000:]
    # u*v: 2 + kap + rho/dt (LinearComb merged the three coefficients into one sum)
    t = ex["bilinear item 0 term 0"]
    assert t.startswith("FE data on the device: 2 + 1 * [function #0 (P1, 1 comp., 64 dofs) comp. 0 op 0] + 10 * [function #1 (P0, 1 comp., 162 dofs)")
    assert "interpreter" in ex["bilinear item 0 term 1"]          # (1+x)*kap
    assert "interpreter" in ex["bilinear item 0 term 2"]          # kap*ff
    assert re.search(r": 0 \+ -1 \* \[function #\d \(P2, 1 comp., 343 dofs, own node table\) comp. 0 op 0\]", ": " + ex["bilinear item 0 term 3"])
    assert "interpreter" in ex["bilinear item 0 term 4"]          # a function on another mesh
    assert re.search(r"0 \+ 0.25 \* \[function #0 ", ex["boundary bilinear item 1 term 0"])
    rc, out, _ = run_ff(EXPLAIN.replace("STATEMENT", "real[int] b = va(0,Vh);"), {"FFCUDA_EXPLAIN": "1"}, want_fail=True)
    ex = dict(re.findall(r"ffcuda explain: (.*? item \d+ term \d+): (.*)", out))
    assert len(ex) == 3, out[-3000:]
    t = ex["linear item 0 term 0"]    # -(ff + f2 + uk/dt - 3 rho + 2)
    assert t.startswith("FE data on the device: -2 + -1 * [function #0 (P1, 1 comp., 64 dofs) comp. 0 op 0] + -1 * [function #1 (P1, 3 comp., 192 dofs) comp. 1 op 0]")
    assert " + -10 * [function #2 (P1, 1 comp., 64 dofs) comp. 0 op 0] + 3 * [function #3 (P0, 1 comp., 162 dofs) comp. 0 op 0]" in t
    assert re.search(r"0 \+ -1 \* \[function #1 \(P1, 3 comp., 192 dofs\) comp. 2 op 6\]", ex["linear item 0 term 1"])   # -dz(f3) dx(v)
    assert re.search(r"0 \+ -1 \* \[function #0 ", ex["boundary linear item 1 term 0"])
    # what is NOT affine in the FE data on the range of the data is refused, however the sample of mesh nodes falls: kinks
    # (max, abs, ?:), products with an indicator supported on a few elements, functions of functions, factors depending on x;
    # a hat function that vanishes on almost every element keeps its exact factor, a function that is 0 everywhere too
    rc, out, _ = run_ff(EXPLAIN_TRICKY, {"FFCUDA_EXPLAIN": "1"}, want_fail=True)
    ex = [e for _, e in re.findall(r"ffcuda explain: (.*? item \d+ term \d+): (.*)", out)]
    assert len(ex) == 11, out[-3000:]
    for t in (0, 1, 2, 5, 6, 7):   # max(ff,0.9), sin(ff), x*ff | abs(ff), (1+chi)*ff, (ff>1.2 ? 2:1)*ff
        assert "interpreter" in ex[t], (t, ex[t])
    assert re.search(r": 0 \+ 2 \* \[function #\d \(P1, 1 comp., 2197 dofs\) comp. 0 op 0\]$", ": " + ex[3])      # 2*bump
    assert re.search(r": 0 \+ 3 \* \[function #\d ", ": " + ex[4])                                                  # 3*zero
    assert re.search(r": 0 \+ 2 \* \[function #\d \(P0, 1 comp., 10368 dofs\)", ": " + ex[8])                     # 2*chi
    assert re.search(r": 0 \+ -10 \* \[function #\d ", ": " + ex[9])                                                # -ff/dt
    assert ex[10].count("* [function") == 2 and " + 1 * [function" in ex[10]                                        # ff + bump
    # a statement that comes back in a loop: the flattened program is kept, the factors follow the script's variables
    rc, out, _ = run_ff(EXPLAIN_LOOP, {"FFCUDA_EXPLAIN": "1"}, want_fail=True)
    fac = re.findall(r"term 0: FE data on the device: 0 \+ ([0-9.]+) \* \[function #0 ", out)
    assert [round(float(f), 4) for f in fac] == [10.0, 5.0, 3.3333] and out.count("term 1: FE data on the device: 0 + 2 * [function #1") == 3
    # switched off: nothing is recognised
    rc, out, _ = run_ff(EXPLAIN.replace("STATEMENT", "real[int] b = va(0,Vh);"),
                        {"FFCUDA_EXPLAIN": "1", "FFCUDA_NO_FE_DOFS": "1"}, want_fail=True)
    assert out.count("evaluated by the interpreter") == 3 and "FE data on the device" not in out


# problem / solve statements (Problem::eval, fflib/problem.cpp:12198-12450): the tutorial shape `solve Poisson(u,v,...) = a - l + on`
SOLVE_DUMP = '{ ofstream f("u.txt"); f.precision(17); for(int i=0;i<UU[].n;++i) f << UU[][i] << endl; }\n'
SOLVE_CASES = {
    # examples/tutorial/Laplace.edp shape
    "laplace2d_p1": """mesh Th = square(24,19);
fespace Vh(Th,P1); Vh u,v;
solve Poisson(u,v,solver=CG,eps=1e-14) = int2d(Th)(dx(u)*dx(v)+dy(u)*dy(v)) - int2d(Th)(1.*v) + on(1,2,3,4,u=0);
""",
    "poisson3d_p2_robin_neumann": f"""mesh3 Th = cube(4,3,4,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)]);
fespace Vh(Th,P2); Vh u,v;
solve Pb(u,v,solver=CG,eps=1e-14) = int3d(Th)({LAP3}) + int2d(Th,2)(2.*u*v) - int3d(Th)(1.*v) - int2d(Th,3)(0.5*v) + on(1,u=1);
""",
    "lame3d_p1_vector": f"""{LAME_PRE}
mesh3 Th = cube(3,4,3);
fespace Vh(Th,[P1,P1,P1]); Vh [u1,u2,u3],[v1,v2,v3];
solve Lame([u1,u2,u3],[v1,v2,v3],solver=CG,eps=1e-14) = int3d(Th)({LAME}) - int3d(Th)(-0.05*v3) + on(1,u1=0,u2=0,u3=0);
""",
    "problem_reused_init": f"""mesh3 Th = cube(5,4,5);
fespace Vh(Th,P1); Vh u=0,v;
problem Pb(u,v,solver=CG,eps=1e-14,init=1) = int3d(Th)(u*v+{LAP3}) - int3d(Th)(1.*v) + on(1,2,u=0.5);
problem Pa(u,v,solver=CG,eps=1e-14) = int3d(Th)(3.*u*v+{LAP3}) - int3d(Th)(2.*v) + on(1,2,u=0.5);
for (int it = 0; it < 3; ++it) {{ Pa; }}
""",
    # examples/tutorial/Laplace.edp as it is written there: the data is a function of x and y
    "tutorial_laplace_fxy": """mesh Th = square(20,20);
fespace Vh(Th,P2); Vh u,v;
func f = x*y;
solve Poisson(u,v,solver=CG,eps=1e-14) = int2d(Th)(dx(u)*dx(v)+dy(u)*dy(v)) - int2d(Th)(f*v) + on(1,2,3,4,u=0);
""",
    # idp/Heat3d.idp shape: the previous time step enters the right-hand side as an FE function
    "heat3d_time_loop_uold": f"""mesh3 Th = cube(5,4,5);
fespace Vh(Th,P1); Vh u=0,v,uold;
real dt = 0.1;
int it = 0;
problem Heat(u,v,solver=CG,eps=1e-14,init=it) = int3d(Th)(u*v/dt+{LAP3}) - int3d(Th)(uold*v/dt) - int3d(Th)((1+x)*v) + on(1,2,u=0);
for (it = 0; it < 4; ++it) {{ uold = u; Heat; }}
""",
    # a P0 material coefficient and a P1 function in the reaction term
    "materials_p0_p1_coefficients": """mesh Th = square(14,12);
fespace Vh(Th,P1); Vh u,v,rho=1+x*y;
fespace Ph(Th,P0); Ph kappa = 1 + 9*(x>0.5)*(y<0.5);
solve Pb(u,v,solver=CG,eps=1e-14) = int2d(Th)(kappa*(dx(u)*dx(v)+dy(u)*dy(v))+rho*u*v) - int2d(Th)(rho*v) + on(1,u=0);
""",
    "dirichlet_data_function": """mesh Th = square(15,13);
fespace Vh(Th,P2); Vh u,v;
func g = cos(3*x)*y;
solve Pb(u,v,solver=CG,eps=1e-14) = int2d(Th)(dx(u)*dx(v)+dy(u)*dy(v)) - int2d(Th)(1.*v) + on(1,2,3,4,u=g);
""",
    # Newton iterations on -div((1+u^2) grad u) = 10: the Jacobian has three coefficient functions of the iterate, the
    # residual has derivatives of the test function times data depending on the mesh point
    "newton_nonlinear_diffusion": """mesh Th = square(12,11);
fespace Vh(Th,P1); Vh u=0,v,w,uk;
problem Newton(w,v,solver=GMRES,eps=1e-10) = int2d(Th)((1+uk*uk)*(dx(w)*dx(v)+dy(w)*dy(v)) + 2*uk*w*(dx(uk)*dx(v)+dy(uk)*dy(v)))
    - int2d(Th)((1+uk*uk)*(dx(uk)*dx(v)+dy(uk)*dy(v)) - 10.*v) + on(1,2,3,4,w=0);
for (int it = 0; it < 4; ++it) { uk = u; Newton; u[] -= w[]; }
""",
    "default_solver": """mesh Th = square(9,8);
fespace Vh(Th,P2); Vh u,v;
solve Pb(u,v) = int2d(Th)(dx(u)*dx(v)+dy(u)*dy(v)+u*v) - int2d(Th)(1.*v) - int1d(Th,2)(0.3*v) + on(4,u=0);
""",
    "convdiff2d_gmres": """mesh Th = square(12,11);
fespace Vh(Th,P1); Vh u,v;
solve Pb(u,v,solver=GMRES,eps=1e-14) = int2d(Th)(dx(u)*dx(v)+dy(u)*dy(v)+6.*dx(u)*v+2.*dy(u)*v) - int2d(Th)(1.*v) + on(1,2,3,4,u=0);
""",
}


def run_solve(body, u0, env_extra):
    src = 'load "msh3"\nload "ffcuda"\n' + body + SOLVE_DUMP.replace("UU", u0)
    with tempfile.TemporaryDirectory() as td:
        with open(os.path.join(td, "case.edp"), "w") as f:
            f.write(src)
        env = dict(os.environ, FF_LOADPATH=LIBDIR, **env_extra)
        r = subprocess.run([FF, "-nw", "-v", "1", "case.edp"], capture_output=True, text=True, cwd=td, env=env, timeout=600)
        out = r.stdout + r.stderr
        u = np.loadtxt(os.path.join(td, "u.txt"), ndmin=1) if os.path.exists(os.path.join(td, "u.txt")) else None
        return r.returncode, out, u


@needs_ff
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SOLVE_CASES))
def test_plugin_problem_solve_matches_freefem(name):
    body = SOLVE_CASES[name]
    u0 = "u1" if "u1" in body else "u"
    rc, out, gpu = run_solve(body, u0, {"FFCUDA_STRICT": "1", "FFCUDA_VERBOSE": "1"})
    assert rc == 0, out[-3000:]
    assert "problem matrix" in out and "problem right-hand side" in out and "assembled on the GPU" in out
    if "solver=CG" in body:
        assert "GC (ffcuda)" in out
    if "solver=GMRES" in body:
        assert "fgmres (ffcuda)" in out
    if name == "heat3d_time_loop_uold":  # init=it: the matrix is built once, the right-hand side at every step
        assert out.count("problem matrix") == 1 and out.count("problem right-hand side") == 4
    if name == "problem_reused_init":   # Pa rebuilds its matrix at every call (no init=): 3 matrices, 3 right-hand sides
        assert out.count("problem matrix") == 3 and out.count("problem right-hand side") == 3
    rc, out_cpu, cpu = run_solve(body, u0, {"FFCUDA_DISABLE": "1"})
    assert rc == 0 and "(ffcuda)" not in out_cpu and "assembled on the GPU" not in out_cpu
    assert np.max(np.abs(gpu - cpu)) <= (1e-9 if name.startswith("newton") else 1e-11) * np.abs(cpu).max()


SOLVE_FALLBACK = """mesh Th = square(10,9);
fespace Vh(Th,P1nc); Vh u,v;
solve Poisson(u,v,solver=LU) = int2d(Th)((1+x)*(dx(u)*dx(v)+dy(u)*dy(v))) - int2d(Th)(x*v) + on(1,2,3,4,u=0);
fespace Wh(Th,P1dc); Wh w,ww;
solve Proj(w,ww) = int2d(Th)(w*ww) - int2d(Th)(u*ww);
cout << "WW " << w[].sum << endl;
"""


@needs_ff
def test_plugin_problem_solve_left_to_freefem_when_not_claimed():
    """the re-pointed problem/solve types hand everything they do not claim (x-dependent data, other elements) to
    Problem::eval unchanged: same numbers with the plugin loaded (no GPU needed) and with the plugin disabled."""
    rc, out, u = run_solve(SOLVE_FALLBACK, "u", {})
    assert rc == 0, out[-3000:]
    assert out.count("problem / solve left to FreeFEM") >= 2
    rc, out2, u2 = run_solve(SOLVE_FALLBACK, "u", {"FFCUDA_DISABLE": "1"})
    assert rc == 0 and np.array_equal(u, u2)
    assert re.search(r"^WW (\S+)", out, re.M).group(1) == re.search(r"^WW (\S+)", out2, re.M).group(1)
    rc, out, _ = run_solve(SOLVE_CASES["laplace2d_p1"], "u", {}) if not _has_cuda() else (1, "no CPU fallback", None)
    assert rc != 0 and "no CPU fallback" in out


def _has_cuda():
    import torch

    return torch.cuda.is_available()


@needs_ff
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["poisson3d_p1", "lame3d_p2", "poisson3d_p2_bnd_g", "diff3d_p2_kappa", "laplace2d_p2_dirichlet_g"])
def test_plugin_check_mode(name):
    """FFCUDA_CHECK=1: every intercepted varf statement also runs FreeFEM's own operator and compares (pattern identical,
    values / right-hand side within 1e-12) inside the FreeFEM process - the drop-in validating itself on the user's script."""
    rc, out, res = run_ff(CASES[name], {"FFCUDA_CHECK": "1"})
    assert rc == 0
    m = re.search(r"ffcuda check: matrix .*pattern identical, max \|dA\| / max \|A\| = (\S+)", out)
    b = re.search(r"ffcuda check: right-hand side .* max \|db\| / max \|b\| = (\S+)", out)
    assert m and b and float(m.group(1)) <= 1e-12 and float(b.group(1)) <= 1e-12


@needs_ff
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["lame3d_p1_evar", "diff3d_p2_kappa"])
def test_plugin_coefficient_grouping_on_a_sample(name):
    """the grouping of proportional coefficient tables decided on a sparse sample of the elements and verified on the chunked
    full pass (forced here on small meshes: FFCUDA_SAMPLE_MIN=1, ~5 sampled elements) gives the same matrices"""
    src = CASES[name]
    _, out, gpu = run_ff(src, {"FFCUDA_STRICT": "1", "FFCUDA_VERBOSE": "1", "FFCUDA_SAMPLE_MIN": "1", "FFCUDA_SAMPLE_N": "5"})
    assert "coefficient function(s) depending on the mesh point" in out
    _, _, cpu = run_ff(src, {"FFCUDA_DISABLE": "1"})
    compare(gpu, cpu, tight=True)


@needs_ff
@pytest.mark.gpu
def test_plugin_coefficient_grouping_sample_misled():
    """a coefficient that vanishes on the sample but not on the mesh: the chunked pass notices, the exact grouping is taken"""
    body = """mesh Th = square(14,12);
fespace Vh(Th,P1); Vh u,v;
solve Pb(u,v,solver=CG,eps=1e-14) = int2d(Th)((1+x)*(dx(u)*dx(v)+dy(u)*dy(v)) + 50.*(x>0.93)*(y>0.9)*u*v) - int2d(Th)(1.*v) + on(1,u=0);
"""
    env = {"FFCUDA_STRICT": "1", "FFCUDA_VERBOSE": "1", "FFCUDA_SAMPLE_MIN": "1", "FFCUDA_SAMPLE_N": "3"}
    rc, out, gpu = run_solve(body, "u", env)
    assert rc == 0 and "problem matrix" in out
    rc, _, cpu = run_solve(body, "u", {"FFCUDA_DISABLE": "1"})
    assert rc == 0 and np.max(np.abs(gpu - cpu)) <= 1e-11 * np.abs(cpu).max()


# ---------------------------------------------------------------------------------------------------------------
# solver=CG with a user preconditioner, set(A,solver=CG) after the script changed the matrix, A'^-1
# (precon= must never be silently replaced by Jacobi; a stale device copy must never be
# adopted)
# ---------------------------------------------------------------------------------------------------------------
PRECON_NOT_CLAIMED = """mesh Th = square(12,11);
fespace Vh(Th,P1nc); Vh u,v;
varf va(u,v) = int2d(Th)(dx(u)*dx(v)+dy(u)*dy(v)+u*v) + int2d(Th)(1.*v) + on(1,2,3,4,u=0);
matrix M = va(Vh,Vh);
real[int] dm(Vh.ndof); dm = M.diag;
func real[int] Pre(real[int] &xx) { for (int i=0;i<xx.n;++i) xx[i] = (dm[i] > 1e20 ? 1. : 0.5)*xx[i]/dm[i]; return xx; }
matrix A = va(Vh,Vh,solver=CG,eps=1e-8,precon=Pre);
real[int] b = va(0,Vh);
verbosity=1; u[] = 0; u[] = A^-1*b; verbosity=0;
{ ofstream f("u.txt"); f.precision(17); for(int i=0;i<u[].n;++i) f << u[][i] << endl; }
"""


@needs_ff
def test_plugin_cg_user_preconditioner_runs_freefems_cg():
    """P1nc is not on the GPU path, so this runs without a device: with the plugin loaded `solver=CG` resolves to
    SolverCudaCG, which must hand a `precon=` solve to FreeFEM's own SolverCG (and say so) - same iterations, same bits."""
    rc, out, u = run_solve(PRECON_NOT_CLAIMED, "u", {})
    assert rc == 0, out[-3000:]
    assert "solver=CG left to FreeFEM (user preconditioner" in out
    rc2, out2, u2 = run_solve(PRECON_NOT_CLAIMED, "u", {"FFCUDA_DISABLE": "1"})
    assert rc2 == 0
    it = re.findall(r"GC[^\n]*?after\s+(\d+)", out)
    it2 = re.findall(r"GC[^\n]*?after\s+(\d+)", out2)
    assert it and it == it2
    assert np.array_equal(u, u2)
    rc, out, _ = run_solve(PRECON_NOT_CLAIMED, "u", {"FFCUDA_STRICT": "1"})
    assert rc != 0 and "FFCUDA_STRICT" in out


SET_SOLVER_CASES = {
    # the matrix is assembled on the GPU with the default solver, CHANGED by the script, and only then given to CG: the
    # solve must see the changed values (no stale device copy)
    "set_after_scaling": """mesh3 Th = cube(5,4,6);
fespace Vh(Th,P1); Vh u,v;
varf va(u,v) = int3d(Th)(dx(u)*dx(v)+dy(u)*dy(v)+dz(u)*dz(v)+u*v) + int3d(Th)(1.*v) + on(1,u=0);
matrix A = va(Vh,Vh);
real[int] b = va(0,Vh);
A = 3.*A;
set(A,solver=CG,eps=1e-12);
""",
    "set_after_diag_edit": """mesh Th = square(9,8);
fespace Vh(Th,P2); Vh u,v;
varf va(u,v) = int2d(Th)(dx(u)*dx(v)+dy(u)*dy(v)) + int2d(Th)(1.*v) + on(1,2,3,4,u=0);
matrix A = va(Vh,Vh,solver=CG,eps=1e-12);
real[int] b = va(0,Vh);
for (int i=0;i<Vh.ndof;i+=7) A(i,i) = 2.*A(i,i);
set(A,solver=CG,eps=1e-12);
""",
    "transposed_solve": """mesh3 Th = cube(4,5,4);
fespace Vh(Th,P1); Vh u,v;
varf va(u,v) = int3d(Th)(dx(u)*dx(v)+dy(u)*dy(v)+dz(u)*dz(v)) + int3d(Th)(1.*v) + on(1,2,3,4,5,6,u=0);
matrix A = va(Vh,Vh,solver=CG,eps=1e-12);
real[int] b = va(0,Vh);
u[] = A'^-1*b;
real[int] keep = u[];
""",
}


@needs_ff
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SET_SOLVER_CASES))
def test_plugin_set_solver_and_transposed(name):
    body = SET_SOLVER_CASES[name] + "verbosity=1; u[] = 0; u[] = A^-1*b; verbosity=0;\n"
    rc, out, gpu = run_solve(body, "u", {"FFCUDA_VERBOSE": "1"})
    assert rc == 0, out[-3000:]
    assert "assembled on the GPU" in out and "GC (ffcuda)" in out
    rc, out_cpu, cpu = run_solve(body, "u", {"FFCUDA_DISABLE": "1"})
    assert rc == 0 and "(ffcuda)" not in out_cpu
    assert np.max(np.abs(gpu - cpu)) <= 1e-10 * np.abs(cpu).max()


VEPS_CASE = """mesh3 Th = cube(5,4,6);
fespace Vh(Th,P1); Vh u,v;
varf va(u,v) = int3d(Th)(dx(u)*dx(v)+dy(u)*dy(v)+dz(u)*dz(v)) + int3d(Th)(1.*v) + on(1,2,3,4,5,6,u=0);
real ve = 1e-6;
matrix A = va(Vh,Vh,solver=CG,veps=ve);
real[int] b = va(0,Vh);
verbosity=1; u[] = 0; u[] = A^-1*b; verbosity=0;
cout.precision(15);
cout << "VEPS " << ve << endl;
"""


@needs_ff
@pytest.mark.gpu
def test_plugin_veps_is_the_stopping_threshold():
    """`veps=` comes back as FreeFEM's SolverCG leaves it: the ABSOLUTE threshold sqrt(eps^2 <g0,Cg0>) ConjugueGradient stopped
    on (femlib/CG.cpp:226, VirtualSolverCG.hpp:186), not the residual reached."""
    rc, out, gpu = run_solve(VEPS_CASE, "u", {"FFCUDA_VERBOSE": "1"})
    assert rc == 0 and "GC (ffcuda)" in out, out[-2000:]
    rc, out_cpu, cpu = run_solve(VEPS_CASE, "u", {"FFCUDA_DISABLE": "1"})
    assert rc == 0
    vg = float(re.search(r"^VEPS (\S+)", out, re.M).group(1))
    vc = float(re.search(r"^VEPS (\S+)", out_cpu, re.M).group(1))
    assert vc != 1e-6 and abs(vg - vc) <= 1e-10 * abs(vc)
    assert np.max(np.abs(gpu - cpu)) <= 1e-11 * np.abs(cpu).max()


# ---- mesh generators on the device: `cube(nx,ny,nz)` and `buildlayers(Th2,n,...)` ----------------------
MESH_DUMP = """
{ ofstream f("NAME.txt"); f.precision(17);
  f << Th.nv << " " << Th.nt << " " << Th.nbe << " " << Th.mesure << " " << Th.bordermesure << endl;
  for(int i=0;i<Th.nv;++i) f << Th(i).x << " " << Th(i).y << " " << Th(i).z << " " << Th(i).label << endl;
  for(int k=0;k<Th.nt;++k){ for(int i=0;i<4;++i) f << Th[k][i] << " "; f << Th[k].label << " " << Th[k].mesure << endl; }
  for(int e=0;e<Th.nbe;++e){ for(int i=0;i<3;++i) f << Th.be(e)[i] << " ";
     f << Th.be(e).label << " " << Th.be(e).Element << " " << Th.be(e).whoinElement << endl; }
  for(int k=0;k<Th.nt;++k) for(int e=0;e<4;++e){ int ee=e; int kk=Th[k].adj(ee); f << kk << " " << ee << endl; }
}
"""
MESHGEN = {
    "cube": 'mesh3 Th = cube(5,4,3);',
    "cube_region": 'mesh3 Th = cube(2,3,4,region=7);',
    "layers_heat3d": 'mesh Th2=square(6,6); int[int] refm=[1,1,2,1,3,1,4,1]; int[int] refu=[0,1];\n'
                     'mesh3 Th=buildlayers(Th2,6,zbound=[0.,1.],labelmid=refm,labelup=refu,labeldown=refu);',
    "layers_coef": 'mesh Th2=square(5,4,[x+0.1*y,y*(1+0.2*x)],flags=1); int[int] rr=[0,7]; int[int] rm=[1,11,3,13,1,21];\n'
                   'int[int] ru=[0,31]; int[int] rd=[0,41,5,6];\n'
                   'mesh3 Th=buildlayers(Th2,4,zbound=[0.1*x*y,1.+0.3*y-0.2*x],coef=1.02-x,region=rr,labelmid=rm,labelup=ru,labeldown=rd);',
    "layers_disk": 'border C(t=0,2*pi){x=cos(t);y=sin(t);label=9;}\nborder D(t=0,2*pi){x=0.3*cos(t)+0.1;y=0.3*sin(t);label=8;}\n'
                   'mesh Th2=buildmesh(C(24)+D(10)); int[int] rr=[0,3,1,4];\n'
                   'mesh3 Th=buildlayers(Th2,5,zbound=[-0.2*(1-x*x-y*y),0.5+0.5*(1-x*x-y*y)+0.1*x],coef=0.15+0.85*(x*x+0.5*y*y),reftet=rr);',
}


def _mesh_script(gen):
    s = 'load "msh3"\nload "ffcuda"\n' + gen + "\n" + MESH_DUMP.replace("NAME", "mesh")
    s += "fespace Vh(Th,P1);\nvarf va(u,v) = int3d(Th)(" + LAP3 + "+u*v) + int3d(Th)(1.*v);\n"
    s += "matrix A = va(Vh,Vh,solver=CG,eps=1e-14);\nreal[int] b = va(0,Vh);\n" + DUMP + "Vh u;\n" + SOLVE.replace("UU", "u")
    return s


def _run_mesh(src, env_extra):
    with tempfile.TemporaryDirectory() as td:
        with open(os.path.join(td, "case.edp"), "w") as f:
            f.write(src)
        env = dict(os.environ, FF_LOADPATH=LIBDIR, **env_extra)
        r = subprocess.run([FF, "-nw", "-v", "1", "case.edp"], capture_output=True, text=True, cwd=td, env=env, timeout=600)
        out = r.stdout + r.stderr
        assert r.returncode == 0, out[-3000:]
        return out, open(os.path.join(td, "mesh.txt")).read(), open(os.path.join(td, "A.txt")).read(), np.loadtxt(os.path.join(td, "u.txt"))


@needs_ff
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MESHGEN))
def test_plugin_device_mesh_generators(name):
    """the Mesh3 a script gets from `cube` / `buildlayers` with the plugin loaded is the one FreeFEM builds: vertices,
    labels, elements, measures, boundary elements with orientation and links, adjacency (Th[k].adj) — the dumps are equal
    as text; the first fespace on it adopts the device copy (no upload) and assembles the same matrix"""
    src = _mesh_script(MESHGEN[name])
    out, mesh, a, u = _run_mesh(src, {"FFCUDA_STRICT": "1", "FFCUDA_VERBOSE": "1"})
    assert ("cube(" if name.startswith("cube") else "buildlayers(") in out and "on the device" in out
    if name != "cube_region":
        assert "mesh already on the device" in out
    out0, mesh0, a0, u0 = _run_mesh(src, {"FFCUDA_DISABLE": "1"})
    assert "on the device" not in out0
    assert mesh == mesh0
    ga = np.array(" ".join(ln for ln in a.splitlines()[1:] if not ln.startswith("#")).split(), dtype=np.float64)
    ca = np.array(" ".join(ln for ln in a0.splitlines()[1:] if not ln.startswith("#")).split(), dtype=np.float64)
    assert ga.shape == ca.shape and np.max(np.abs(ga - ca)) <= 1e-12 * np.abs(ca).max()
    assert np.max(np.abs(u - u0)) <= 1e-12 * np.abs(u0).max()


@needs_ff
def test_plugin_mesh_generators_fall_back_without_a_device_or_for_options():
    """no device, or an option the device generator does not cover (label= of cube): FreeFEM's own generator runs and the
    plugin says so"""
    src = ('load "msh3"\nload "ffcuda"\nint[int] ll=[1,1,1,1,2,2]; mesh3 Th = cube(2,3,2,label=ll);\nmesh Th2=square(2,2);\n'
           'mesh3 T3=buildlayers(Th2,2,transfo=[x,y,2*z]);\ncout << "SIZES " << Th.nt << " " << T3.nt << endl;\n')
    rc, out, _ = run_ff(src, {"FFCUDA_VERBOSE": "1"}, want_fail=True)
    assert rc == 0, out[-2000:]
    assert "SIZES 72 48" in out
    assert out.count("left to FreeFEM") == 2


# ---- several GPUs driven from the one FreeFEM process (FFCUDA_NGPU): the solvers share the matrix out by row blocks ------------
NGPU_CASES = {
    "cg": script(3, "cube(9,8,10)", "P1", LAP3, "1.*v", "on(1,2,3,4,5,6,u=0)", eps="1e-14"),
    "cg_p2_vector": script(3, "cube(2,3,2)", "[P2,P2,P2]", LAME, "-0.05*v3", "on(1,u1=0,u2=0,u3=0)", pre=LAME_PRE,
                           unk="[u1,u2,u3]", tst="[v1,v2,v3]", eps="1e-14"),
    "gmres": script(3, "cube(8,7,9)", "P1", LAP3 + "+8.*dx(u)*v+3.*dy(u)*v-2.*dz(u)*v", "1.*v", "on(1,2,3,4,5,6,u=0)", eps="1e-14",
                    solver="GMRES,dimKrylov=30"),
}


@needs_ff
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(NGPU_CASES))
def test_plugin_solves_on_two_gpus(name):
    """FFCUDA_NGPU=2: `u[] = A^-1*b` runs the distributed CG / GMRES on two GPUs from the one FreeFem++ process (row blocks of the
    host MatriceMorse, one host thread per GPU, ghost exchange + all-reduced dot products); same converged solution as
    FreeFEM's own solver.  Self-skips on a box with one device."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    src = NGPU_CASES[name]
    _, out, gpu = run_ff(src, {"FFCUDA_NGPU": "2", "FFCUDA_VERBOSE": "1"})
    assert "2 GPUs driven from this process" in out and "shared out over 2 GPUs" in out
    assert ("fgmres (ffcuda, 2 GPUs)" if "GMRES" in src else "GC (ffcuda, 2 GPUs)") in out
    _, _, cpu = run_ff(src, {"FFCUDA_DISABLE": "1"})
    compare(gpu, cpu, tight=True)
