"""The oracle reproduces the committed trajectories (tests/golden/traj_*.npz, made by make_golden.py)."""
import numpy as np
import pytest

from oracle import msm_oracle as o
from golden_util import GOLDEN, oracle_streams
from conftest import rel_l2

CONFIGS = ["spherical-tophat", "spherical-tophat-cosmo", "planeWave3d_e10_sym"]


@pytest.mark.parametrize("name", CONFIGS)
def test_oracle_matches_golden(name):
    z = np.load(f"{GOLDEN}/traj_{name}.npz")
    its = oracle_streams(name)
    names = [str(s) for s in z["streams"]]
    for j, sim_name in enumerate(names):
        p = next(q for q in its if q.sim_name == sim_name)
        sim = o.SimulationObject(p, z[f"s{j}_psi0"])
        scal = z[f"s{j}_scalars"]
        for step in range(1, 6):
            sim.update()
            row = scal[step - 1]
            assert abs(sim.last_dt - row[0]) <= 1e-13 * abs(row[0])
            assert abs(sim.last_potential_max - row[1]) <= 1e-12 * abs(row[1])
            assert abs(sim.parameters.time - row[3]) <= 1e-13 * max(abs(row[3]), 1e-300)
            if f"s{j}_psi_{step}" in z.files:
                assert rel_l2(sim.psi, z[f"s{j}_psi_{step}"]) < 1e-12


def test_golden_initial_conditions_are_reproducible():
    from golden_util import initial_wavefunction
    z = np.load(f"{GOLDEN}/traj_spherical-tophat.npz")
    its = oracle_streams("spherical-tophat")
    p = next(q for q in its if q.sim_name == str(z["streams"][0]))
    assert rel_l2(initial_wavefunction(p), z["s0_psi0"]) < 1e-14
