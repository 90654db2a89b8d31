#!/bin/bash
# A/B of two versions of l96_energy.cu on the GPU box: timing, phase cycles and output agreement.
#   tools/energy_ab.sh [problems] [N]     (binaries tools/dbg/en_{old,new}_{0,1} built in the authoring container:
#   _0 plain, _1 with VGPA_EN_PROF phase timers; old = tools/dbg/l96_energy_r01.cu, new = vgpa_b200/csrc/l96_energy.cu)
P=${1:-888}; N=${2:-1001}
mkdir -p gpurun_out
for v in old new; do
  echo "== $v: full launch ($P x $N), then phase cycles loaded (148 x $N) and unloaded (1 x 148)"
  tools/dbg/en_${v}_0 $P $N gpurun_out/en_${v}.bin | tail -1
  tools/dbg/en_${v}_1 148 $N | tail -1
  tools/dbg/en_${v}_1 1 148 | tail -1
done
python - <<PY
import numpy as np
a=np.fromfile("gpurun_out/en_old.bin"); b=np.fromfile("gpurun_out/en_new.bin")
N=$N; D=40
for name,lo,hi in (("esde",0,N),("dEm",N,N+N*D),("dEs",N+N*D,a.size)):
    x,y=a[lo:hi],b[lo:hi]
    print(name,"max rel diff",float(np.abs(x-y).max()/np.abs(x).max()))
PY
