#!/usr/bin/env python3
"""Golden fixtures of RECTANGULAR matrices, `matrix B = vb(Uh,Vh)` with two different spaces on one mesh (rows = dofs of
the test space Vh, columns = dofs of the space of the unknown Uh; Element_Op with Ku != Kv, fflib/problem.cpp:6337-6437):
the unmodified reference (oracle/_ref, built by `make -C oracle ref`) runs the script and dumps the mesh, the dof tables
of both spaces and the matrix as HashMatrix holds it (insertion order) and as [I,J,C] gives it, 17 significant digits.

    python tests/golden/make_golden_rect.py [case ...]
"""
import os, subprocess, sys, tempfile
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import FF, HERE, toks  # noqa: E402

W3 = "[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)]"
W2 = "[x+0.2*y*y,y*(1+0.3*x)]"
# name -> dim, mesh, space / names of the unknown, space / names of the test function, the form, options of the integral
CASES = {
    # the divergence block of a Stokes problem: velocity [P2,P2,P2], pressure test function P1
    "rect3d_div_p2p1": dict(dim=3, mesh=f"cube(2,2,2,{W3})", ufe="[P2,P2,P2]", unk="[u1,u2,u3]", vfe="P1", tst="[q]",
                            bil="-(dx(u1)+dy(u2)+dz(u3))*q+0.5*u2*dx(q)"),
    # its transpose assembled as such: unknown P1, test functions [P2,P2,P2]
    "rect3d_grad_p1p2": dict(dim=3, mesh=f"cube(2,2,2,{W3})", ufe="P1", unk="[p]", vfe="[P2,P2,P2]", tst="[v1,v2,v3]",
                             bil="-p*(dx(v1)+dy(v2)+dz(v3))+dx(p)*v3"),
    # scalar spaces of different order in 2-D (the matrix of the L2 projection P1 -> P2 and a non-symmetric term)
    "rect2d_p1_to_p2": dict(dim=2, mesh=f"square(4,3,{W2})", ufe="P1", unk="[u]", vfe="P2", tst="[v]", bil="u*v+0.5*dx(u)*dy(v)"),
    "rect2d_div_p2p1": dict(dim=2, mesh=f"square(3,4,{W2})", ufe="[P2,P2]", unk="[u1,u2]", vfe="P1", tst="[q]", bil="-(dx(u1)+dy(u2))*q"),
    # same order, different number of components; u2 never appears (its columns are structural zeros of the element matrices)
    "rect3d_p1vec_p1": dict(dim=3, mesh="cube(3,2,3)", ufe="[P1,P1,P1]", unk="[u1,u2,u3]", vfe="P1", tst="[q]", bil="dx(u1)*q+2.*u3*dz(q)"),
    # two components against two components of another order, region-restricted rule left at its default, lumped quadrature
    "rect2d_p2vec_p1vec": dict(dim=2, mesh="square(3,3)", ufe="[P2,P2]", unk="[u1,u2]", vfe="[P1,P1]", tst="[v1,v2]",
                               bil="u1*v1+u2*v2+0.25*dy(u1)*v2", intopt=",qft=qf2pT"),
    # an integral restricted to one of two regions: only the visited elements put their couples into the matrix
    "rect3d_region": dict(dim=3, mesh="change(cube(3,2,3),fregion=(x<0.5)?1:2)", ufe="P2", unk="[u]", vfe="P1", tst="[q]", bil="dx(u)*q+u*q",
                          region=",2"),
    # mixed-order product spaces (Taylor-Hood Stokes matrices) in ONE fespace: assembled by the plugin as scalar blocks in the
    # global dof numbering of the space (tests/ff_cases.py MIXED_CASES); Uh and Vh are two objects of the same type here
    "mixed2d_stokes": dict(dim=2, mesh=f"square(3,3,{W2})", ufe="[P2,P2,P1]", unk="[u1,u2,p]", vfe="[P2,P2,P1]", tst="[v1,v2,q]",
                           bil="dx(u1)*dx(v1)+dy(u1)*dy(v1)+dx(u2)*dx(v2)+dy(u2)*dy(v2)-p*dx(v1)-p*dy(v2)-dx(u1)*q-dy(u2)*q-1e-10*p*q"),
    "mixed3d_stokes": dict(dim=3, mesh=f"cube(2,2,2,{W3})", ufe="[P2,P2,P2,P1]", unk="[u1,u2,u3,p]", vfe="[P2,P2,P2,P1]", tst="[v1,v2,v3,q]",
                           bil="dx(u1)*dx(v1)+dy(u1)*dy(v1)+dz(u1)*dz(v1)+dx(u2)*dx(v2)+dy(u2)*dy(v2)+dz(u2)*dz(v2)"
                               "+dx(u3)*dx(v3)+dy(u3)*dy(v3)+dz(u3)*dz(v3)-p*(dx(v1)+dy(v2)+dz(v3))-(dx(u1)+dy(u2)+dz(u3))*q-1e-10*p*q"),
}


def script(c, out):
    dim = c["dim"]
    mtype, integ = ("mesh", "int2d") if dim == 2 else ("mesh3", "int3d")
    nvk = dim + 1
    s = ['load "msh3"', f"{mtype} Th = {c['mesh']};", f"fespace Uh(Th,{c['ufe']});", f"fespace Vh(Th,{c['vfe']});",
         f"varf vb({c['unk']},{c['tst']}) = {integ}(Th{c.get('region', '')}{c.get('intopt', '')})({c['bil']});", "matrix B = vb(Uh,Vh);"]
    s.append(f'{{ ofstream f("{out}/mesh.txt"); f.precision(17);')
    s.append('  f << Th.nv << " " << Th.nt << " " << Th.nbe << endl;')
    if dim == 2:
        s.append('  for(int i=0;i<Th.nv;++i) f << Th(i).x << " " << Th(i).y << " " << Th(i).label << endl;')
    else:
        s.append('  for(int i=0;i<Th.nv;++i) f << Th(i).x << " " << Th(i).y << " " << Th(i).z << " " << Th(i).label << endl;')
    s.append("  for(int k=0;k<Th.nt;++k){ for(int i=0;i<%d;++i) f << Th[k][i] << \" \"; f << Th[k].label << endl; } }" % nvk)
    for w in ("Uh", "Vh"):
        s.append(f'{{ ofstream f("{out}/dof_{w}.txt"); f << {w}.ndof << " " << {w}.ndofK << endl;')
        s.append(f'  for(int k=0;k<Th.nt;++k){{ for(int i=0;i<{w}.ndofK;++i) f << {w}(k,i) << " "; f << endl; }} }}')
    s.append(f'{{ ofstream f("{out}/Bins.txt"); f.precision(17); f << B; }}')
    s.append("{ int[int] I(1),J(1); real[int] C(1); [I,J,C]=B;")
    s.append(f'  ofstream f("{out}/B.txt"); f.precision(17); f << B.n << " " << B.m << " " << B.nnz << endl;')
    s.append('  for(int k=0;k<I.n;++k) f << I[k] << " " << J[k] << " " << C[k] << endl; }')
    return "\n".join(s) + "\n"


def run_case(name):
    c = CASES[name]
    dim = c["dim"]
    with tempfile.TemporaryDirectory() as td:
        src = script(c, td)
        with open(os.path.join(td, "case.edp"), "w") as f:
            f.write(src)
        r = subprocess.run([FF, "-nw", "-v", "0", os.path.join(td, "case.edp")], capture_output=True, text=True, cwd=td)
        if r.returncode != 0:
            sys.stderr.write(r.stdout[-3000:] + r.stderr[-3000:])
            raise SystemExit(f"reference failed on {name}")
        t = toks(os.path.join(td, "mesh.txt"))
        nv, nt = int(t[0]), int(t[1])
        p = 3
        vt = np.array(t[p:p + nv * (dim + 1)], dtype=np.float64).reshape(nv, dim + 1); p += nv * (dim + 1)
        et = np.array(t[p:p + nt * (dim + 2)], dtype=np.int64).reshape(nt, dim + 2); p += nt * (dim + 2)
        assert p == len(t)
        out = dict(dim=np.int32(dim), xyz=np.ascontiguousarray(vt[:, :dim]), conn=et[:, :dim + 1].astype(np.int32),
                   elab=et[:, dim + 1].astype(np.int32), edp=np.array(src))
        for w in ("Uh", "Vh"):
            t = toks(os.path.join(td, f"dof_{w}.txt"))
            out["ndof_" + w] = np.int32(int(t[0]))
            out["dof_" + w] = np.array(t[2:], dtype=np.int32).reshape(nt, int(t[1]))
        t = toks(os.path.join(td, "B.txt"))
        n, m, nnz = int(t[0]), int(t[1]), int(t[2])
        a = np.array(t[3:], dtype=np.float64).reshape(-1, 3)
        with open(os.path.join(td, "Bins.txt")) as f:
            lines = [ln for ln in f if not ln.startswith("#")]
        ins = np.array(" ".join(lines[1:]).split(), dtype=np.float64).reshape(-1, 3)
        nnz = ins.shape[0]  # B.nnz is read after [I,J,C]=B, see below
        if a.shape[0] == nnz + 1:
            # `[I,J,C] = B` of a matrix without an entry in its last row and column ends with a zero at (n, m), one past the
            # last indices, that carries the shape of the matrix; it is not part of what the assembly produced
            assert a[-1, 0] == n and a[-1, 1] == m and a[-1, 2] == 0.0
            a = a[:-1]
        assert a.shape[0] == nnz and n == int(out["ndof_Vh"]) and m == int(out["ndof_Uh"])
        out.update(n=np.int32(n), m=np.int32(m), ins_i=ins[:, 0].astype(np.int32), ins_j=ins[:, 1].astype(np.int32),
                   coo_i=a[:, 0].astype(np.int32), coo_j=a[:, 1].astype(np.int32), coo_a=a[:, 2].copy())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(f"{name}: nv={nv} nt={nt} n={n} m={m} nnz={nnz}")


if __name__ == "__main__":
    if not os.path.exists(FF):
        raise SystemExit("oracle/_ref/FreeFem++-nw missing: run `make -C oracle ref -j8` (needs /root/reference)")
    for nm in (sys.argv[1:] or list(CASES)):
        run_case(nm)
