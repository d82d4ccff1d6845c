#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED
reference (FreeFEM 4.15, built by `make -C oracle ref` into oracle/_ref/) on small
cases of the hot path and dumping mesh, dof table, matrix (COO as stored by
HashMatrix, i.e. insertion order), right-hand side, CG solution and iteration
count with 17 significant digits (exact round trip of fp64).

Only runs where /root/reference was available to build oracle/_ref (the dev
container); the resulting *.npz files are committed and are what the tests read.

    python tests/golden/make_golden.py [case ...]
"""
import os, re, subprocess, sys, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
FF = os.path.join(ROOT, "oracle", "_ref", "FreeFem++-nw")

LAP2 = "dx(u)*dx(v)+dy(u)*dy(v)"
LAP3 = "dx(u)*dx(v)+dy(u)*dy(v)+dz(u)*dz(v)"
LAME = ("lambda*(dx(u1)+dy(u2)+dz(u3))*(dx(v1)+dy(v2)+dz(v3))"
        "+2.*mu*(dx(u1)*dx(v1)+dy(u2)*dy(v2)+dz(u3)*dz(v3)"
        "+0.5*(dy(u1)+dx(u2))*(dy(v1)+dx(v2))+0.5*(dz(u1)+dx(u3))*(dz(v1)+dx(v3))"
        "+0.5*(dz(u2)+dy(u3))*(dz(v2)+dy(v3)))")
LAME_PRE = "real E=21.5e4, sigma=0.29; real mu=E/(2*(1+sigma)); real lambda=E*sigma/((1+sigma)*(1-2*sigma));"

# name -> dict(dim, mesh, fe, unk, tst, bil, lin, bc, intopt, pre, solve)
CASES = {
    # config 1 shape: 2-D P1 Laplace, f=1, u=0 on the whole boundary
    "lap2d_p1_sq4": dict(dim=2, mesh="square(4,4)", fe="P1", bil=LAP2, lin="1.*v", bc="on(1,2,3,4,u=0)"),
    "lap2d_p1_sq12x9": dict(dim=2, mesh="square(12,9)", fe="P1", bil=LAP2, lin="1.*v", bc="on(1,2,3,4,u=0)"),
    "lap2d_p1_warp": dict(dim=2, mesh="square(7,5,[x+0.2*y*y,y*(1+0.3*x)])", fe="P1", bil=LAP2, lin="3.*v",
                          bc="on(1,u=1)+on(3,u=2)"),
    "lap2d_p2_sq3": dict(dim=2, mesh="square(3,3)", fe="P2", bil=LAP2, lin="1.*v", bc="on(1,2,3,4,u=0)"),
    "lap2d_p2_warp": dict(dim=2, mesh="square(4,3,[x+0.2*y*y,y*(1+0.3*x)])", fe="P2", bil=LAP2 + "+2.*u*v",
                          lin="1.*v", bc="on(2,4,u=0)"),
    # config 2 / 5 shape: 3-D P1 Poisson
    "lap3d_p1_cube2": dict(dim=3, mesh="cube(2,2,2)", fe="P1", bil=LAP3, lin="1.*v", bc="on(1,2,3,4,5,6,u=0)"),
    "lap3d_p1_cube5": dict(dim=3, mesh="cube(5,5,5)", fe="P1", bil=LAP3, lin="1.*v", bc="on(1,2,3,4,5,6,u=0)"),
    "lap3d_p1_cube342": dict(dim=3, mesh="cube(3,4,2)", fe="P1", bil=LAP3, lin="1.*v", bc="on(1,2,3,4,5,6,u=0)"),
    "lap3d_p1_warp": dict(dim=3, mesh="cube(3,3,3,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", fe="P1", bil=LAP3,
                          lin="2.*v", bc="on(1,u=1)+on(6,u=-1)"),
    "lap3d_p2_cube2": dict(dim=3, mesh="cube(2,2,2)", fe="P2", bil=LAP3, lin="1.*v", bc="on(1,2,3,4,5,6,u=0)"),
    # config 4 shape: heat step matrix  u*v/dt + grad u . grad v  (Heat3d.idp)
    "heat3d_p1_cube3": dict(dim=3, mesh="cube(3,3,3)", fe="P1", pre="real dt=0.01;", bil="u*v/dt+" + LAP3,
                            lin="1.*v", bc="on(1,2,3,4,5,6,u=0)"),
    # mass only, lumped quadrature
    "mass3d_p1_lump": dict(dim=3, mesh="cube(2,2,2)", fe="P1", bil="u*v", lin="1.*v", bc="", intopt=",qfV=qfV1lump",
                           solve=False),
    "mass2d_p2_qf2": dict(dim=2, mesh="square(3,2)", fe="P2", bil="u*v", lin="1.*v", bc="", intopt=",qft=qf2pT",
                          solve=False),
    # non-symmetric form: checks row = test function, column = unknown
    "nonsym3d_p1": dict(dim=3, mesh="cube(2,3,2)", fe="P1", bil="dx(u)*v+2.*u*dy(v)+0.5*dz(u)*dx(v)", lin="dx(v)+2.*v",
                        bc="", solve=False),
    "nonsym2d_p2": dict(dim=2, mesh="square(3,2)", fe="P2", bil="dx(u)*v+2.*u*dy(v)+0.5*dy(u)*dx(v)", lin="dy(v)+2.*v",
                        bc="", solve=False),
    # config 3 shape: 3-component Lame (beam-3d.md macros), gravity on the 3rd component, clamped on label 1
    "lame3d_p2_cube2": dict(dim=3, mesh="cube(2,2,2)", fe="[P2,P2,P2]", unk="[u1,u2,u3]", tst="[v1,v2,v3]",
                            pre=LAME_PRE, bil=LAME, lin="-0.05*v3", bc="on(1,u1=0,u2=0,u3=0)"),
    "lame3d_p1_cube3": dict(dim=3, mesh="cube(3,2,3)", fe="[P1,P1,P1]", unk="[u1,u2,u3]", tst="[v1,v2,v3]",
                            pre=LAME_PRE, bil=LAME, lin="-0.05*v3", bc="on(1,u1=0,u2=0,u3=0)"),
    # exact elimination of the Dirichlet dofs (HashMatrix::SetBC with tgv < 0): rows (tgv=-1), rows and columns (tgv=-2),
    # columns only (tgv=-3)
    "lap3d_p1_tgvm1": dict(dim=3, mesh="cube(3,3,2)", fe="P1", bil=LAP3, lin="2.*v", bc="on(1,u=1)+on(6,u=-1)", tgv=-1,
                           solve=False),
    "lap3d_p1_tgvm2": dict(dim=3, mesh="cube(3,2,3)", fe="P1", bil=LAP3, lin="1.*v", bc="on(1,2,3,4,5,6,u=0)", tgv=-2),
    "lap2d_p2_tgvm2": dict(dim=2, mesh="square(4,3)", fe="P2", bil=LAP2, lin="1.*v", bc="on(1,2,3,4,u=0)", tgv=-2),
    "lame3d_p1_tgvm1": dict(dim=3, mesh="cube(2,2,2)", fe="[P1,P1,P1]", unk="[u1,u2,u3]", tst="[v1,v2,v3]",
                            pre=LAME_PRE, bil=LAME, lin="-0.05*v3", bc="on(1,u1=0,u2=0,u3=0)", tgv=-1, solve=False),
    "lap3d_p1_tgvm3": dict(dim=3, mesh="cube(2,3,2)", fe="P1", bil=LAP3, lin="1.*v", bc="on(1,3,u=0)", tgv=-3, solve=False),
    # Neumann / traction data: boundary integrals of the linear form (Element_rhs on border elements, problem.cpp:8439-8587)
    "lap3d_p1_neumann": dict(dim=3, mesh="cube(3,3,3,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", fe="P1", bil=LAP3, lin="1.*v",
                             blin="+int2d(Th,2,3)(2.5*v)", bc="on(1,u=0)"),
    "lap2d_p2_neumann": dict(dim=2, mesh="square(4,3,[x+0.2*y*y,y*(1+0.3*x)])", fe="P2", bil=LAP2, lin="1.*v",
                             blin="+int1d(Th,2)(1.5*v)", bc="on(4,u=0)"),
    "lame3d_p1_traction": dict(dim=3, mesh="cube(2,3,2)", fe="[P1,P1,P1]", unk="[u1,u2,u3]", tst="[v1,v2,v3]",
                               pre=LAME_PRE, bil=LAME, lin="-0.05*v3", blin="+int2d(Th,2)(0.3*v1-0.2*v3)",
                               bc="on(1,u1=0,u2=0,u3=0)"),
    "lap3d_p2_neumann": dict(dim=3, mesh="cube(2,2,2)", fe="P2", bil=LAP3 + "+u*v", lin="1.*v", blin="+int2d(Th,6)(-1.*v)",
                             bc=""),
    # Robin terms: boundary integrals of the bilinear form (AssembleBilinearForm border loop problem.cpp:1317-1326,
    # Element_Op border branch :6518-6560, 2-D :6216-6290)
    "lap3d_p1_robin": dict(dim=3, mesh="cube(3,3,3,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", fe="P1", bil=LAP3, lin="1.*v",
                           blin="+int2d(Th,2,3)(1.5*u*v)+int2d(Th,2,3)(2.5*v)", bc="on(1,u=0)"),
    "lap2d_p2_robin": dict(dim=2, mesh="square(4,3,[x+0.2*y*y,y*(1+0.3*x)])", fe="P2", bil=LAP2, lin="1.*v",
                           blin="+int1d(Th,2,3)(0.7*u*v)", bc="on(4,u=0)"),
    "lame3d_p1_robin": dict(dim=3, mesh="cube(2,3,2)", fe="[P1,P1,P1]", unk="[u1,u2,u3]", tst="[v1,v2,v3]",
                            pre=LAME_PRE, bil=LAME, lin="-0.05*v3",
                            blin="+int2d(Th,2)(1e4*u1*v1+1e4*u2*v2+5e3*u3*v3+2e3*u1*v3)", bc="on(1,u1=0,u2=0,u3=0)",
                            solve=False),
    "lap3d_p2_robin": dict(dim=3, mesh="cube(2,2,2)", fe="P2", bil=LAP3, lin="1.*v", blin="+int2d(Th,6,1)(2.*u*v)", bc=""),
    # non-symmetric forms solved by GMRES (SolverGMRES / fgmres, femlib/CG.cpp:347-517); the second one restarts
    "convdiff3d_p1_gmres": dict(dim=3, mesh="cube(4,4,4,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", fe="P1",
                                bil=LAP3 + "+8.*dx(u)*v+3.*dy(u)*v-2.*dz(u)*v", lin="1.*v", bc="on(1,2,3,4,5,6,u=0)",
                                solver="GMRES"),
    "convdiff2d_p2_gmres": dict(dim=2, mesh="square(5,4)", fe="P2", bil=LAP2 + "+5.*dx(u)*v+u*v", lin="1.*v",
                                bc="on(1,3,u=0)", solver="GMRES", sopt=",dimKrylov=25"),
    # right-hand sides whose data depend on the mesh point (the coefficient is evaluated at every quadrature node)
    "lap3d_p1_fxyz": dict(dim=3, mesh="cube(3,3,3,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", fe="P1", bil=LAP3, lin="(x*y+sin(z))*v",
                          bc="on(1,2,u=0)"),
    "lap2d_p2_fxy": dict(dim=2, mesh="square(4,3,[x+0.2*y*y,y*(1+0.3*x)])", fe="P2", bil=LAP2, lin="exp(x)*y*v", bc="on(4,u=0)"),
    "lame3d_p1_fvec": dict(dim=3, mesh="cube(2,3,2)", fe="[P1,P1,P1]", unk="[u1,u2,u3]", tst="[v1,v2,v3]", pre=LAME_PRE, bil=LAME,
                           lin="x*v1-0.05*(1+y)*v3", bc="on(1,u1=0,u2=0,u3=0)"),
    # bilinear forms whose coefficients depend on the mesh point (P1: exact through the moments of the coefficient)
    "diff3d_p1_kappa": dict(dim=3, mesh="cube(3,3,3,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", fe="P1",
                            bil="(1+x*y+z*z)*(" + LAP3 + ")+2.*u*v", lin="1.*v", bc="on(1,2,u=0)"),
    "reac2d_p1_rho": dict(dim=2, mesh="square(5,4,[x+0.2*y*y,y*(1+0.3*x)])", fe="P1", bil=LAP2 + "+(1+sin(x)*y)*u*v", lin="1.*v",
                          bc="on(4,u=0)"),
    "lame3d_p1_evar": dict(dim=3, mesh="cube(2,3,2)", fe="[P1,P1,P1]", unk="[u1,u2,u3]", tst="[v1,v2,v3]", pre=LAME_PRE,
                           bil="(1+x)*(" + LAME + ")", lin="-0.05*v3", bc="on(1,u1=0,u2=0,u3=0)"),
    # boundary integrals whose data depend on the mesh point (Neumann g(x) v, Robin alpha(x) u v)
    "lap3d_p1_bnd_g": dict(dim=3, mesh="cube(3,3,3,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", fe="P1", bil=LAP3, lin="1.*v",
                           blin="+int2d(Th,2,3)((1+x*z)*u*v)+int2d(Th,2,3)((y+sin(z))*v)+int2d(Th,6)(0.5*v)", bc="on(1,u=0)"),
    "lap2d_p2_bnd_g": dict(dim=2, mesh="square(4,3,[x+0.2*y*y,y*(1+0.3*x)])", fe="P2", bil=LAP2, lin="1.*v",
                           blin="+int1d(Th,2,3)((1+x*y)*u*v)+int1d(Th,2)(exp(y)*v)", bc="on(4,u=0)"),
    "lame3d_p1_bnd_g": dict(dim=3, mesh="cube(2,3,2)", fe="[P1,P1,P1]", unk="[u1,u2,u3]", tst="[v1,v2,v3]", pre=LAME_PRE, bil=LAME,
                            lin="-0.05*v3", blin="+int2d(Th,3)(1e3*(1+x)*(u1*v1+u2*v2+u3*v3))+int2d(Th,2)(0.3*z*v1-0.2*(1+y)*v3)",
                            bc="on(1,u1=0,u2=0,u3=0)"),
    # the same on P2 spaces (the per-pair tensors are formed from the coefficient values at the quadrature nodes)
    "diff3d_p2_kappa": dict(dim=3, mesh="cube(2,2,2,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", fe="P2",
                            bil="(1+x*y+z*z)*(" + LAP3 + ")+2.*u*v", lin="1.*v", bc="on(1,2,u=0)"),
    "reac2d_p2_rho": dict(dim=2, mesh="square(4,3,[x+0.2*y*y,y*(1+0.3*x)])", fe="P2", bil=LAP2 + "+(1+sin(x)*y)*(u*v+0.5*dx(u)*v+0.5*u*dx(v))",
                          lin="1.*v", bc="on(4,u=0)"),
    "lame3d_p2_evar": dict(dim=3, mesh="cube(2,2,2)", fe="[P2,P2,P2]", unk="[u1,u2,u3]", tst="[v1,v2,v3]", pre=LAME_PRE,
                           bil="(1+x)*(" + LAME + ")", lin="-0.05*v3", bc="on(1,u1=0,u2=0,u3=0)"),
    # ... and with derivatives of the test function (the residual of a Newton step: int(dx(uk) dx(v) + ...))
    "resid2d_p1_grad": dict(dim=2, mesh="square(5,4,[x+0.2*y*y,y*(1+0.3*x)])", fe="P1", bil=LAP2, lin="x*y*v+sin(x)*dx(v)-y*dy(v)",
                            bc="on(4,u=0)"),
    "resid3d_p2_grad": dict(dim=3, mesh="cube(2,2,2,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", fe="P2", bil=LAP3, lin="(1+z)*dz(v)+x*v-y*z*dx(v)",
                            bc="on(1,2,u=0)"),
    "resid3d_p1_vec_grad": dict(dim=3, mesh="cube(2,3,2)", fe="[P1,P1,P1]", unk="[u1,u2,u3]", tst="[v1,v2,v3]", pre=LAME_PRE, bil=LAME,
                                lin="x*dx(v1)+y*v2-0.05*(1+x)*dz(v3)+z*dy(v1)", bc="on(1,u1=0,u2=0,u3=0)"),
    # half storage (sym=1): MatriceElementaireSymetrique / the symmetric Element_Op, lower triangle only (HashMatrix.cpp:1319-1325)
    "lap3d_p1_sym": dict(dim=3, mesh="cube(3,3,3)", fe="P1", bil=LAP3, lin="1.*v", bc="on(1,2,3,4,5,6,u=0)", sym=1),
    "lap2d_p2_sym": dict(dim=2, mesh="square(4,3,[x+0.2*y*y,y*(1+0.3*x)])", fe="P2", bil=LAP2 + "+2.*u*v", lin="1.*v",
                         bc="on(2,4,u=0)", sym=1),
    "lame3d_p1_sym": dict(dim=3, mesh="cube(2,2,3)", fe="[P1,P1,P1]", unk="[u1,u2,u3]", tst="[v1,v2,v3]",
                          pre=LAME_PRE, bil=LAME, lin="-0.05*v3", bc="on(1,u1=0,u2=0,u3=0)", sym=1),
    "lame3d_p2_warp": dict(dim=3, mesh="cube(2,1,2,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", fe="[P2,P2,P2]",
                           unk="[u1,u2,u3]", tst="[v1,v2,v3]", pre=LAME_PRE, bil=LAME, lin="-0.05*v3",
                           bc="on(1,u1=0,u2=0,u3=0)+on(3,u1=0.01,u2=0,u3=-0.02)"),
    # boundary integrals with derivatives of the unknown / of the test function (every node of the adjacent element is reached)
    "lap3d_p1_bnd_grad": dict(dim=3, mesh="cube(3,3,3,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", fe="P1", bil=LAP3, lin="1.*v",
                              blin="+int2d(Th,2,3)(1.5*dx(u)*v+0.5*u*dy(v)+2.*dz(u)*dx(v)+0.25*u*v)+int2d(Th,6)(0.3*dz(v)-1.*v)",
                              bc="on(1,u=0)", solve=False),
    "lap2d_p2_bnd_grad": dict(dim=2, mesh="square(4,3,[x+0.2*y*y,y*(1+0.3*x)])", fe="P2", bil=LAP2, lin="1.*v",
                              blin="+int1d(Th,2,3)(0.7*dx(u)*v+0.2*dy(u)*dy(v))+int1d(Th,2)(1.5*dx(v)+0.5*v)", bc="on(4,u=0)", solve=False),
    "lame3d_p1_bnd_grad": dict(dim=3, mesh="cube(2,3,2)", fe="[P1,P1,P1]", unk="[u1,u2,u3]", tst="[v1,v2,v3]", pre=LAME_PRE, bil=LAME,
                               lin="-0.05*v3", blin="+int2d(Th,3)(1e3*(dx(u1)*v2+u3*dz(v1)+u2*v2))+int2d(Th,2)(0.3*dy(v1)-0.2*v3)",
                               bc="on(1,u1=0,u2=0,u3=0)", solve=False),
    "lap3d_p2_bnd_gradq": dict(dim=3, mesh="cube(2,2,2,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", fe="P2", bil=LAP3, lin="1.*v",
                               blin="+int2d(Th,2,3)((1+x*z)*(dx(u)*v+0.5*u*dz(v)))", bc="on(1,u=0)", solve=False),
    # FE functions as data of the form (f-2: shipped as dof arrays): P0 / P1 / P2 coefficients, right-hand sides and Newton
    # residuals given by FE functions, Robin / Neumann data given by FE functions, components of a vector function
    "fe3d_p1_data": dict(dim=3, mesh="cube(3,3,3,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", fe="P1",
                         fes=[dict(space="W1", fe="P1", decl="W1 kap=1+x*y+z*z, ff=x*y+sin(z), uk=x*x+y*z;", arrays=["kap", "ff", "uk"]),
                              dict(space="W0", fe="P0", decl="W0 rho=1+x+2*z;", arrays=["rho"])],
                         bil="kap*(" + LAP3 + ")+rho*u*v", lin="ff*v+dx(uk)*dx(v)+dy(uk)*dy(v)+dz(uk)*dz(v)",
                         blin="+int2d(Th,2,3)(kap*u*v)+int2d(Th,2,3)(ff*v)", bc="on(1,u=0)"),
    "fe3d_p2_data": dict(dim=3, mesh="cube(2,2,2,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", fe="P2",
                         fes=[dict(space="W1", fe="P1", decl="W1 kap=1+x*y+z*z;", arrays=["kap"]),
                              dict(space="W2", fe="P2", decl="W2 uk=x*x+y*z+sin(x*z), m2=2+x*y*z;", arrays=["uk", "m2"])],
                         bil="kap*(" + LAP3 + ")+m2*u*v", lin="uk*v+dx(uk)*dx(v)+dz(uk)*dy(v)",
                         blin="+int2d(Th,6)(uk*v)", bc="on(1,2,u=0)"),
    "fe2d_p1_data": dict(dim=2, mesh="square(5,4,[x+0.2*y*y,y*(1+0.3*x)])", fe="P1",
                         fes=[dict(space="W2", fe="P2", decl="W2 kap=1+sin(x)*y, uk=x*x*y-y*y;", arrays=["kap", "uk"]),
                              dict(space="W1", fe="P1", decl="W1 ff=exp(x)*y;", arrays=["ff"])],
                         bil="kap*(" + LAP2 + ")+u*v", lin="ff*v+dx(uk)*dx(v)+dy(uk)*dy(v)",
                         blin="+int1d(Th,2,3)(ff*u*v)+int1d(Th,2)(dy(uk)*v)", bc="on(4,u=0)"),
    "fe3d_lame_data": dict(dim=3, mesh="cube(2,3,2)", fe="[P1,P1,P1]", unk="[u1,u2,u3]", tst="[v1,v2,v3]", pre=LAME_PRE,
                           fes=[dict(space="Wv", fe="[P1,P1,P1]", decl="Wv [f1,f2,f3]=[x*y,sin(z),-0.05*(1+y)];", arrays=["f1"]),
                                dict(space="W1", fe="P1", decl="W1 ee=1+x;", arrays=["ee"])],
                           bil="ee*(" + LAME + ")", lin="f1*v1+f3*v3+dx(f2)*v2", bc="on(1,u1=0,u2=0,u3=0)"),
    # the reference's own regression problems (examples/tutorial/regtests.edp with the values of ref.edp): Laplace.edp
    # (REFLaplace = 0.167397 +-1%), LaplaceP1.edp (REFLaplaceP1 = 2.34669 +-1%), beam.edp (REFbeam = 2.19089 +-5%, on the mesh
    # buildmesh gives it), written as varf + matrix + CG so that the fixtures hold A, b and the solution
    "tutorial_laplace": dict(dim=2, mesh="square(10,10)", fe="P1", bil=LAP2, lin="1.*v", bc="on(1,2,3,4,u=0)", tgv=1e5),
    "tutorial_laplace_p1": dict(dim=2, mesh="square(10,10)", fe="P1", bil=LAP2, lin="1.*v", blin="+int1d(Th,1)(u*v)+int1d(Th,1)(1.*v)",
                                bc="on(2,3,4,u=0)", tgv=1e5),
    "tutorial_beam": dict(dim=2, pre="real E=21.5, sigma=0.29, gravity=-0.05; real mu=E/(2*(1+sigma)); real lambda=E*sigma/((1+sigma)*(1-2*sigma));"
                          " real sqrt2=sqrt(2.);"
                          " border ba(t=2,0){x=0;y=t;label=1;} border bb(t=0,10){x=t;y=0;label=2;}"
                          " border bc(t=0,2){x=10;y=t;label=1;} border bd(t=0,10){x=10-t;y=2;label=3;}",
                          mesh="buildmesh(bb(20)+bc(5)+bd(20)+ba(5))", fe="[P1,P1]", unk="[uu,vv]", tst="[w,s]",
                          bil="lambda*(dx(w)+dy(s))*(dx(uu)+dy(vv))+2.*mu*(dx(w)*dx(uu)+dy(s)*dy(vv)+(dy(w)+dx(s))*(dy(uu)+dx(vv))/2.)",
                          lin="gravity*s", bc="on(1,uu=0,vv=0)"),
}


def script(c, out):
    dim = c["dim"]
    unk, tst = c.get("unk", "u"), c.get("tst", "v")
    mtype, integ = ("mesh", "int2d") if dim == 2 else ("mesh3", "int3d")
    opt = c.get("intopt", "")
    bc = ("+" + c["bc"]) if c["bc"] else ""
    nvk = dim + 1
    s = []
    s.append('load "msh3"')
    s.append(c.get("pre", ""))
    s.append(f"{mtype} Th = {c['mesh']};")
    s.append(f"fespace Vh(Th,{c['fe']});")
    for sp in c.get("fes", []):  # FE functions used as data of the form: their spaces, declarations (dumped below)
        s.append(f"fespace {sp['space']}(Th,{sp['fe']});")
        s.append(sp["decl"])
    s.append(f"varf va({unk},{tst}) = {integ}(Th{opt})({c['bil']}) + {integ}(Th{opt})({c['lin']}){c.get('blin', '')}{bc};")
    tg = (",tgv=%g" % c["tgv"]) if "tgv" in c else ""
    sy = ",sym=1" if c.get("sym") else ""
    solver, sopt = c.get("solver", "CG"), c.get("sopt", "")
    s.append(f"matrix A = va(Vh,Vh,solver={solver},eps=1e-6{tg}{sy}{sopt});")
    s.append(f"real[int] b = va(0,Vh{tg});")
    # mesh dump
    s.append(f'{{ ofstream f("{out}/mesh.txt"); f.precision(17);')
    s.append('  f << Th.nv << " " << Th.nt << " " << Th.nbe << endl;')
    if dim == 2:
        s.append('  for(int i=0;i<Th.nv;++i) f << Th(i).x << " " << Th(i).y << " " << Th(i).label << endl;')
    else:
        s.append('  for(int i=0;i<Th.nv;++i) f << Th(i).x << " " << Th(i).y << " " << Th(i).z << " " << Th(i).label << endl;')
    s.append("  for(int k=0;k<Th.nt;++k){ for(int i=0;i<%d;++i) f << Th[k][i] << \" \"; f << Th[k].label << endl; }" % nvk)
    s.append("  for(int e=0;e<Th.nbe;++e){ for(int i=0;i<%d;++i) f << Th.be(e)[i] << \" \"; "
             "f << Th.be(e).label << \" \" << Th.be(e).Element << \" \" << Th.be(e).whoinElement << endl; } }" % dim)
    # dof table
    s.append(f'{{ ofstream f("{out}/dof.txt"); f << Vh.ndof << " " << Vh.ndofK << endl;')
    s.append('  for(int k=0;k<Th.nt;++k){ for(int i=0;i<Vh.ndofK;++i) f << Vh(k,i) << " "; f << endl; } }')
    # matrix as stored (COO, insertion order), rhs
    s.append(f'{{ ofstream f("{out}/Ains.txt"); f.precision(17); f << A; }}  // HashMatrix storage (insertion) order')
    s.append("{ int[int] I(1),J(1); real[int] C(1); [I,J,C]=A;")
    s.append(f'  ofstream f("{out}/A.txt"); f.precision(17); f << A.n << " " << A.m << " " << A.nnz << endl;')
    s.append('  for(int k=0;k<I.n;++k) f << I[k] << " " << J[k] << " " << C[k] << endl; }')
    s.append(f'{{ ofstream f("{out}/b.txt"); f.precision(17); for(int i=0;i<b.n;++i) f << b[i] << endl; }}')
    for sp in c.get("fes", []):
        w = sp["space"]
        s.append(f'{{ ofstream f("{out}/fes_{w}.txt"); f << {w}.ndof << " " << {w}.ndofK << endl;')
        s.append(f'  for(int k=0;k<Th.nt;++k){{ for(int i=0;i<{w}.ndofK;++i) f << {w}(k,i) << " "; f << endl; }} }}')
        for a in sp["arrays"]:
            s.append(f'{{ ofstream f("{out}/fe_{a}.txt"); f.precision(17); for(int i=0;i<{a}[].n;++i) f << {a}[][i] << endl; }}')
    if c.get("solve", True):
        s.append("Vh %s; %s[] = 0; verbosity=1;" % (unk, unk.strip("[]").split(",")[0]))
        u0 = unk.strip("[]").split(",")[0]
        s.append(f"{u0}[] = A^-1*b; verbosity=0;")
        s.append(f'{{ ofstream f("{out}/u.txt"); f.precision(17); for(int i=0;i<{u0}[].n;++i) f << {u0}[][i] << endl; }}')
        # second solve converged to round-off: the comparison point that does not depend on where eps=1e-6 stops
        s.append(f"verbosity=1; set(A,solver={solver},eps=1e-14{sopt}); {u0}[] = 0; {u0}[] = A^-1*b; verbosity=0;")
        s.append(f'{{ ofstream f("{out}/u14.txt"); f.precision(17); for(int i=0;i<{u0}[].n;++i) f << {u0}[][i] << endl; }}')
    return "\n".join(s) + "\n"


def toks(path):
    with open(path) as f:
        return f.read().split()


def run_case(name, case=None, save=True):
    """runs the reference on CASES[name] (or on the given case dict) and returns the dump; save=False: nothing is written"""
    c = case if case is not None else CASES[name]
    dim = c["dim"]
    with tempfile.TemporaryDirectory() as td:
        edp = os.path.join(td, "case.edp")
        src = script(c, td)
        with open(edp, "w") as f:
            f.write(src)
        r = subprocess.run([FF, "-nw", "-v", "1", edp], capture_output=True, text=True, cwd=td)
        if r.returncode != 0:
            sys.stderr.write(r.stdout[-3000:] + r.stderr[-3000:])
            raise SystemExit(f"reference failed on {name}")
        t = toks(os.path.join(td, "mesh.txt"))
        nv, nt, nbe = int(t[0]), int(t[1]), int(t[2])
        p = 3
        vt = np.array(t[p:p + nv * (dim + 1)], dtype=np.float64).reshape(nv, dim + 1); p += nv * (dim + 1)
        et = np.array(t[p:p + nt * (dim + 2)], dtype=np.int64).reshape(nt, dim + 2); p += nt * (dim + 2)
        bt = np.array(t[p:p + nbe * (dim + 3)], dtype=np.int64).reshape(nbe, dim + 3); p += nbe * (dim + 3)
        assert p == len(t)
        t = toks(os.path.join(td, "dof.txt"))
        ndof, ndofK = int(t[0]), int(t[1])
        dof = np.array(t[2:], dtype=np.int32).reshape(nt, ndofK)
        t = toks(os.path.join(td, "A.txt"))
        n, m, nnz = int(t[0]), int(t[1]), int(t[2])
        a = np.array(t[3:], dtype=np.float64).reshape(-1, 3)
        assert a.shape[0] == nnz and n == ndof
        b = np.array(toks(os.path.join(td, "b.txt")), dtype=np.float64)
        with open(os.path.join(td, "Ains.txt")) as f:
            lines = [ln for ln in f if not ln.startswith("#")]
        if "tgv" in c:  # SetBC with tgv < 0 leaves the matrix in CSR state: no insertion order to record
            ins = a.copy()
        else:
            assert "COO" in open(os.path.join(td, "Ains.txt")).readline()
            ins = np.array(" ".join(lines[1:]).split(), dtype=np.float64).reshape(-1, 3)
        assert ins.shape[0] == nnz
        out = dict(dim=np.int32(dim), xyz=np.ascontiguousarray(vt[:, :dim]), vlab=vt[:, dim].astype(np.int32),
                   conn=et[:, :dim + 1].astype(np.int32), elab=et[:, dim + 1].astype(np.int32),
                   bconn=bt[:, :dim].astype(np.int32), blab=bt[:, dim].astype(np.int32),
                   belem=bt[:, dim + 1].astype(np.int32), bface=bt[:, dim + 2].astype(np.int32),
                   ndof=np.int32(ndof), dof=dof,
                   ins_i=ins[:, 0].astype(np.int32), ins_j=ins[:, 1].astype(np.int32),
                   coo_i=a[:, 0].astype(np.int32), coo_j=a[:, 1].astype(np.int32), coo_a=a[:, 2].copy(), b=b,
                   edp=np.array(src))
        for sp in c.get("fes", []):
            t = toks(os.path.join(td, f"fes_{sp['space']}.txt"))
            ndk = int(t[1])
            tab = np.array(t[2:], dtype=np.int32).reshape(nt, ndk)
            for a in sp["arrays"]:
                out["fe_" + a] = np.array(toks(os.path.join(td, f"fe_{a}.txt")), dtype=np.float64)
                assert out["fe_" + a].shape[0] == int(t[0])
                out["fe_" + a + "_dof"] = tab
        if c.get("solve", True):
            out["u"] = np.array(toks(os.path.join(td, "u.txt")), dtype=np.float64)
            mm = re.findall(r"fgmres has converged in\s+(\d+)" if c.get("solver") == "GMRES" else r"GC:\s+converge after\s+(\d+)", r.stdout)
            assert len(mm) == 2, r.stdout[-2000:]
            out["cg_iters"] = np.int32(int(mm[0]))
            out["u14"] = np.array(toks(os.path.join(td, "u14.txt")), dtype=np.float64)
            out["cg_iters14"] = np.int32(int(mm[1]))
        if not save:
            return out
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(f"{name}: nv={nv} nt={nt} nbe={nbe} ndof={ndof} nnz={nnz}"
              + (f" cg_iters={int(out['cg_iters'])}" if "cg_iters" in out else ""))


if __name__ == "__main__":
    if not os.path.exists(FF):
        raise SystemExit("oracle/_ref/FreeFem++-nw missing: run `make -C oracle ref -j8` (needs /root/reference)")
    for nm in (sys.argv[1:] or list(CASES)):
        run_case(nm)
