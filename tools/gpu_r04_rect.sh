#!/bin/bash
# The hardware run the rectangular / mixed-order path is waiting for (written when the round's GPU minutes
# were spent).  One GPU, a few seconds:
#   gpurun --timeout 600 -- bash tools/gpu_r04_rect.sh
# -rxX: every xfail / XPASS with its reason.  All XPASS = drop the xfail mark of tests/test_zz_gpu_rect.py and the
# FFCUDA_RECT guard of plugin/ffcuda.cpp; an xfail prints the first failing assertion or CUDA error.
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_zz_gpu_rect.py -m gpu -q -rxX --runxfail --durations=5 > gpurun_out/r04_rect_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r04_rect_pytest.log
# the same statements with the in-process comparison against FreeFEM's own operator
cat > /tmp/rect_check.edp <<'EOF'
load "msh3"
load "ffcuda"
mesh3 Th = cube(6,5,7);
fespace Uh(Th,[P2,P2,P2]); fespace Ph(Th,P1); fespace Xh(Th,[P2,P2,P2,P1]);
varf vb([u1,u2,u3],[q]) = int3d(Th)(-(dx(u1)+dy(u2)+dz(u3))*q);
varf vs([u1,u2,u3,p],[v1,v2,v3,q]) = int3d(Th)(dx(u1)*dx(v1)+dy(u2)*dy(v2)+dz(u3)*dz(v3)-p*(dx(v1)+dy(v2)+dz(v3))-(dx(u1)+dy(u2)+dz(u3))*q)
  + on(1,2,u1=0,u2=0,u3=0);
matrix B = vb(Uh,Ph);
matrix S = vs(Xh,Xh);
cout << "B " << B.n << " x " << B.m << " nnz " << B.nnz << "   S " << S.n << " nnz " << S.nnz << endl;
EOF
(cd /tmp && FF_LOADPATH=$GRAFT_REPO_ROOT/freefem-sources_b200/lib FFCUDA_RECT=1 FFCUDA_CHECK=1 FFCUDA_VERBOSE=1 \
  timeout 300 $GRAFT_REPO_ROOT/oracle/_ref/FreeFem++-nw -nw -v 0 rect_check.edp) > gpurun_out/r04_rect_check.log 2>&1
echo "FreeFem++ rc=$?" >> gpurun_out/r04_rect_check.log
tail -30 gpurun_out/r04_rect_pytest.log; grep -i "ffcuda\|^B \|rc=" gpurun_out/r04_rect_check.log | tail -12
