"""Generates the committed fixtures under tests/golden/.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

1. configs.json       the reference's own example TOMLs (examples/*.toml, tomls/*.toml) resolved by the oracle's
                      `read_toml` -- so tests and the GPU box never have to read /root/reference.
2. planeWave3d_e10_sym_ic.npz   the reference's IC fixture initial_conditions/planeWave3d_e10_sym.npz (f64 16^3).
3. planeWave1d_ic.npz           the reference's planeWave1d.npz (f64 256).
4. traj_<config>.npz  oracle trajectories (psi after selected update() counts, per-step scalars) for the 16^3
                      configs.  The oracle is a restatement of the reference (parity unpinned for the integrator,
                      see oracle/msm_oracle.py), so these pin the oracle against regressions and give the CUDA
                      path fixed vectors to hit; they are NOT outputs of the reference binary.
"""
import dataclasses
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import msm_oracle as o  # noqa: E402

REF = "/root/reference"
CONFIGS = {
    "gaussian-overdensity-mft": "examples/gaussian-overdensity-mft.toml",
    "spherical-tophat": "examples/spherical-tophat.toml",
    "spherical-tophat-cosmo": "examples/spherical-tophat-cosmo.toml",
    "planeWave3d_e10_sym": "tomls/planeWave3d_e10_sym.toml",
    "planeWave3d_e10_sym-mft": "tomls/planeWave3d_e10_sym-mft.toml",
    "repro-planeWave1d": "examples/repro.toml",
}
SNAP_STEPS = (1, 2, 5, 12, 25)


def main():
    cfgs = {}
    for name, rel in CONFIGS.items():
        t = o.read_toml(os.path.join(REF, rel))
        d = dataclasses.asdict(t)
        d["source"] = rel
        cfgs[name] = d
    with open(os.path.join(HERE, "configs.json"), "w") as f:
        json.dump(cfgs, f, indent=1, sort_keys=True)

    z = np.load(os.path.join(REF, "initial_conditions/planeWave3d_e10_sym.npz"))
    np.savez_compressed(os.path.join(HERE, "planeWave3d_e10_sym_ic.npz"), real=z["real"], imag=z["imag"])
    z = np.load(os.path.join(REF, "planeWave1d.npz"))
    np.savez_compressed(os.path.join(HERE, "planeWave1d_ic.npz"), real=z["real"], imag=z["imag"])

    git = subprocess.run(["git", "-C", ROOT, "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()
    for name in ("spherical-tophat", "spherical-tophat-cosmo", "planeWave3d_e10_sym"):
        t = o.read_toml(os.path.join(REF, CONFIGS[name]))
        its = list(o.simulation_iter(t))
        picks = [0, 1, len(its) - 1]                    # two sampled streams + the mean-field run
        out = {"oracle_git": np.array(git), "streams": np.array([its[i].sim_name for i in picks])}
        for j, i in enumerate(picks):
            p = its[i]
            psi0 = o.initial_wavefunction(p, REF)
            sim = o.SimulationObject(p, psi0)
            out[f"s{j}_psi0"] = psi0
            scal = []
            for step in range(1, max(SNAP_STEPS) + 1):
                sim.update()
                scal.append([sim.last_dt, sim.last_potential_max, sim.last_alias_mass, sim.parameters.time,
                             sim.parameters.tau, float(sim.parameters.current_dumps)])
                if step in SNAP_STEPS:
                    out[f"s{j}_psi_{step}"] = sim.psi.copy()
            out[f"s{j}_scalars"] = np.array(scal)
        np.savez_compressed(os.path.join(HERE, f"traj_{name}.npz"), **out)
        print(name, "done")


if __name__ == "__main__":
    main()
