#!/usr/bin/env python3
"""Golden fixtures for `buildlayers` (fflib/msh3.cpp:895-1757, operator :4536-4760): the UNMODIFIED reference
(oracle/_ref/FreeFem++-nw) builds the layered mesh; the script dumps the 2-D mesh, the per-vertex data the operator
derives (zmin, zmax, coef -> number of layers) and the 3-D mesh, all to 17 digits.  Runs in the build container only
(needs oracle/_ref); the .npz files it writes are committed and travel to the GPU box.

    python tests/golden/make_golden_layers.py            # all cases
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
FF = os.path.join(ROOT, "oracle", "_ref", "FreeFem++-nw")

# zlo / zhi / coef are written in the variables X, Y: the script substitutes x, y for the buildlayers call and the
# vertex coordinates for the per-vertex dump (the operator evaluates them at the vertices, msh3.cpp:4566-4583)
CASES = {
    # the Heat3d.idp mesh (idp/Heat3d.idp:14-16)
    "layers_heat3d": dict(mesh2="square(4,4)", n=4, zlo="0.", zhi="1.", coef=None,
                          opts="labelmid=refm,labelup=refu,labeldown=refu",
                          pre="int[int] refm=[1,1,2,1,3,1,4,1]; int[int] refu=[0,1];",
                          maps=dict(mid=[1, 1, 2, 1, 3, 1, 4, 1], up=[0, 1], down=[0, 1], reg=[])),
    "layers_plain": dict(mesh2="square(3,2)", n=3, zlo="0.", zhi="1.", coef=None, opts="", pre="",
                         maps=dict(mid=[], up=[], down=[], reg=[])),
    # columns that lose layers towards x = 1 (pyramids, tetrahedra, merged lateral faces), curved bottom and top
    "layers_coef": dict(mesh2="square(4,3,[X+0.1*Y,Y*(1+0.2*X)],flags=1)".replace("X", "x").replace("Y", "y"), n=4,
                        zlo="0.1*X*Y", zhi="1.+0.3*Y-0.2*X", coef="1.02-X",
                        opts="region=rr,labelmid=rm,labelup=ru,labeldown=rd",
                        pre="int[int] rr=[0,7]; int[int] rm=[1,11,3,13,1,21]; int[int] ru=[0,31]; int[int] rd=[0,41,5,6];",
                        maps=dict(mid=[1, 11, 3, 13, 1, 21], up=[0, 31], down=[0, 41, 5, 6], reg=[0, 7])),
    # an unstructured 2-D mesh (two regions), every number of layers between 1 and 5
    "layers_disk": dict(mesh2=None, n=5, zlo="-0.2*(1-X*X-Y*Y)", zhi="0.5+0.5*(1-X*X-Y*Y)+0.1*X", coef="0.15+0.85*(X*X+0.5*Y*Y)",
                        opts="region=rr", pre="int[int] rr=[0,3,1,4];",
                        mesh2pre="border C(t=0,2*pi){x=cos(t);y=sin(t);label=9;}\n"
                                 "border D(t=0,2*pi){x=0.3*cos(t)+0.1;y=0.3*sin(t);label=8;}\n"
                                 "mesh Th2 = buildmesh(C(24)+D(10));",
                        maps=dict(mid=[], up=[], down=[], reg=[0, 3, 1, 4])),
    "layers_one": dict(mesh2="square(2,2)", n=1, zlo="-1.", zhi="2.", coef=None, opts="", pre="",
                       maps=dict(mid=[], up=[], down=[], reg=[])),
}


def script(c, out):
    sub_xy = lambda e: e.replace("X", "x").replace("Y", "y")
    sub_v = lambda e: e.replace("X", "xx").replace("Y", "yy")
    s = ['load "msh3"', c["pre"]]
    s.append(c.get("mesh2pre") or f"mesh Th2 = {c['mesh2']};")
    cf = f",coef={sub_xy(c['coef'])}" if c["coef"] else ""
    opts = ("," + c["opts"]) if c["opts"] else ""
    s.append(f"mesh3 Th = buildlayers(Th2,{c['n']},zbound=[{sub_xy(c['zlo'])},{sub_xy(c['zhi'])}]{cf}{opts});")
    s.append(f'{{ ofstream f("{out}/mesh2.txt"); f.precision(17);')
    s.append('  f << Th2.nv << " " << Th2.nt << " " << Th2.nbe << endl;')
    s.append("  for(int i=0;i<Th2.nv;++i){ real xx=Th2(i).x, yy=Th2(i).y; "
             f'f << xx << " " << yy << " " << Th2(i).label << " " << ({sub_v(c["zlo"])}) << " " << ({sub_v(c["zhi"])}) << " " '
             f'<< ({sub_v(c["coef"]) if c["coef"] else "1."}) << endl; }}')
    s.append('  for(int k=0;k<Th2.nt;++k){ for(int i=0;i<3;++i) f << Th2[k][i] << " "; f << Th2[k].label << endl; }')
    s.append('  for(int e=0;e<Th2.nbe;++e){ for(int i=0;i<2;++i) f << Th2.be(e)[i] << " "; '
             'f << Th2.be(e).label << " " << Th2.be(e).Element << " " << Th2.be(e).whoinElement << endl; } }')
    s.append(f'{{ ofstream f("{out}/mesh3.txt"); f.precision(17);')
    s.append('  f << Th.nv << " " << Th.nt << " " << Th.nbe << endl;')
    s.append('  for(int i=0;i<Th.nv;++i) f << Th(i).x << " " << Th(i).y << " " << Th(i).z << " " << Th(i).label << endl;')
    s.append('  for(int k=0;k<Th.nt;++k){ for(int i=0;i<4;++i) f << Th[k][i] << " "; f << Th[k].label << endl; }')
    s.append('  for(int e=0;e<Th.nbe;++e){ for(int i=0;i<3;++i) f << Th.be(e)[i] << " "; '
             'f << Th.be(e).label << " " << Th.be(e).Element << " " << Th.be(e).whoinElement << endl; } }')
    return "\n".join(s) + "\n"


def layers_of(nlayer, coef, zmin, zmax):
    """ni of BuildLayeMesh_Op (fflib/msh3.cpp:4581, :4647-4657)"""
    clayer = np.maximum(0.0, np.minimum(1.0, coef))
    ni = np.maximum(0, np.minimum(nlayer, np.rint(nlayer * clayer).astype(np.int64)))
    maxdz = np.abs(zmin - zmax).max()
    ni[np.abs(zmin - zmax) < maxdz * 1e-6] = 0
    return ni.astype(np.int32)


def run_case(name):
    c = CASES[name]
    with tempfile.TemporaryDirectory() as td:
        src = script(c, td)
        edp = os.path.join(td, "case.edp")
        open(edp, "w").write(src)
        r = subprocess.run([FF, "-nw", "-v", "0", edp], capture_output=True, text=True, cwd=td)
        if r.returncode != 0:
            sys.stderr.write(r.stdout[-3000:] + r.stderr[-3000:])
            raise SystemExit(f"reference failed on {name}")
        t = open(os.path.join(td, "mesh2.txt")).read().split()
        nv, nt, nbe = map(int, t[:3]); p = 3
        v2 = np.array(t[p:p + 6 * nv], dtype=np.float64).reshape(nv, 6); p += 6 * nv
        e2 = np.array(t[p:p + 4 * nt], dtype=np.int64).reshape(nt, 4); p += 4 * nt
        b2 = np.array(t[p:p + 5 * nbe], dtype=np.int64).reshape(nbe, 5); p += 5 * nbe
        assert p == len(t)
        t = open(os.path.join(td, "mesh3.txt")).read().split()
        nv3, nt3, nbe3 = map(int, t[:3]); p = 3
        v3 = np.array(t[p:p + 4 * nv3], dtype=np.float64).reshape(nv3, 4); p += 4 * nv3
        e3 = np.array(t[p:p + 5 * nt3], dtype=np.int64).reshape(nt3, 5); p += 5 * nt3
        b3 = np.array(t[p:p + 6 * nbe3], dtype=np.int64).reshape(nbe3, 6); p += 6 * nbe3
        assert p == len(t)
    zmin, zmax = v2[:, 3].copy(), v2[:, 4].copy()
    ni = layers_of(c["n"], v2[:, 5], zmin, zmax)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    out = dict(nlayer=np.int32(c["n"]), xy=np.ascontiguousarray(v2[:, :2]), vlab2=i32(v2[:, 2]), zmin=zmin, zmax=zmax, ni=ni,
               tri=i32(e2[:, :3]), trilab=i32(e2[:, 3]), bedge=i32(b2[:, :2]), bedge_lab=i32(b2[:, 2]),
               bedge_elem=i32(b2[:, 3]), bedge_face=i32(b2[:, 4]),
               regmap=i32(c["maps"]["reg"]), midmap=i32(c["maps"]["mid"]), upmap=i32(c["maps"]["up"]), downmap=i32(c["maps"]["down"]),
               xyz=np.ascontiguousarray(v3[:, :3]), vlab=i32(v3[:, 3]), conn=i32(e3[:, :4]), elab=i32(e3[:, 4]),
               bconn=i32(b3[:, :3]), blab=i32(b3[:, 3]), belem=i32(b3[:, 4]), bface=i32(b3[:, 5]), edp=np.array(src))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: 2-D nv={nv} nt={nt} nbe={nbe} layers {ni.min()}..{ni.max()} -> nv={nv3} nt={nt3} nbe={nbe3}")


if __name__ == "__main__":
    if not os.path.exists(FF):
        raise SystemExit("oracle/_ref/FreeFem++-nw missing: run `make -C oracle ref -j8` (needs /root/reference)")
    for nm in (sys.argv[1:] or list(CASES)):
        run_case(nm)
