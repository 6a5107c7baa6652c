#!/bin/bash
# A/B builds: scripts/build_variant.sh NAME "-DMACRO=.. ..." [sizes]  ->  msm_b200/libmsm_b200_NAME.so
# Recompiles core.cu and the listed transform lengths (default 512) with the extra flags, links the rest from the
# regular build.  Select at run time with MSM_B200_LIB=msm_b200/libmsm_b200_NAME.so.
set -e
cd "$(dirname "$0")/../msm_b200/csrc"
NAME=$1; FLAGS=$2; SIZES=${3:-512}
make -j8 >/dev/null
D=build/var_$NAME; mkdir -p $D
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -fvisibility=default $FLAGS"
OBJS=""
for n in 2 4 8 16 32 64 128 256 512 1024; do
  if [[ " $SIZES " == *" $n "* ]]; then $NV -DMSM_FFT_N=$n -c fft_inst.cu -o $D/fft_$n.o & OBJS="$OBJS $D/fft_$n.o"; else OBJS="$OBJS build/fft_$n.o"; fi
done
$NV -c core.cu -o $D/core.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libmsm_b200_$NAME.so $OBJS $D/core.o build/fft_tma.o build/sim.o -ldl -lpthread
echo built msm_b200/libmsm_b200_$NAME.so
