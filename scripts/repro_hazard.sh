#!/bin/bash
# Reproducer of the ptxas store / back-edge hazard, to re-validate the work-around on a new toolkit:
#   1. builds the 64- and 128-point kernels WITHOUT the tile-boundary fence (-DMSM_NO_TILE_FENCE) as a variant library,
#   2. runs the static SASS checker on them (expected with CUDA 12.9: > 0 hazards; 0 in the shipping objects),
#   3. on a GPU, runs the blocked-layout 64^3 three-stream trajectory against the oracle with that library
#      (expected with CUDA 12.9: psi off by 1e-5 .. 1e-3; exact with the shipping library).
# If step 2 reports 0 hazards and step 3 passes, the toolkit no longer needs the fence.
set -u
cd "$(dirname "$0")/.."
bash scripts/build_variant.sh nofence "-DMSM_NO_TILE_FENCE" "64 128" || exit 2
echo "== static check, fence-less objects"
python scripts/check_war_hazard.py msm_b200/csrc/build/var_nofence/fft_64.o msm_b200/csrc/build/var_nofence/fft_128.o | tail -3
echo "== static check, shipping objects"
python scripts/check_war_hazard.py msm_b200/csrc/build/fft_64.o msm_b200/csrc/build/fft_128.o | tail -1
if python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" 2>/dev/null; then
  for lib in msm_b200/libmsm_b200_nofence.so msm_b200/libmsm_b200.so; do
    echo "== trajectories with $lib"
    MSM_B200_LIB=$PWD/$lib python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_trajectory_matches_oracle or test_blocked_device_layout" 2>&1 | tail -3
  done
fi
