"""msm_b200 -- B200-native (sm_100a CUDA) implementation of the MSM `simulator` time-evolution loop.

The product is `libmsm_b200.so` (C ABI, include/msm_b200.h).  This package is the Python host over it:
  api.py     object wrappers mirroring the reference's SimulationObject / FFT layer
  config.py  the reference's TOML format -> resolved parameters
  driver.py  `python -m msm_b200 --toml X`, mirror of msm-simulator's main loop
Importing the package loads the shared library and fails loudly if it has not been built.
"""
from ._lib import (COUPLING_INDEPENDENT, COUPLING_SUMMED, MsmError, LIB_PATH)
from .api import (Context, CosmologyParameters, FourierAliasing, SimulationObject, SimulationParameters, forward,
                  get_supercomoving_boxsize, get_tau, inverse, scale_factor_after, spec_grid)
from .config import RunConfig, StreamSpec, parse_seeds, read_toml

__all__ = ["COUPLING_INDEPENDENT", "COUPLING_SUMMED", "MsmError", "LIB_PATH", "Context", "CosmologyParameters",
           "FourierAliasing", "SimulationObject", "SimulationParameters", "forward", "inverse", "spec_grid",
           "get_tau", "get_supercomoving_boxsize", "scale_factor_after", "RunConfig", "StreamSpec", "parse_seeds",
           "read_toml"]
