"""Helpers shared by the tests: the committed reference configs as oracle / msm_b200 parameter objects."""
import json
import os

import numpy as np

from oracle import msm_oracle as o

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_toml(name: str, size=None) -> o.TomlParameters:
    """One of the reference's example TOMLs as resolved by make_golden.py (tests/golden/configs.json)."""
    with open(os.path.join(GOLDEN, "configs.json")) as f:
        d = dict(json.load(f)[name])
    d.pop("source")
    cosmo = d.pop("cosmology")
    t = o.TomlParameters(**d)
    if cosmo is not None:
        t.cosmology = o.CosmologyParameters(**cosmo)
    if size is not None:
        t.size = size
    # IC fixtures live next to this file
    if t.ics.get("type") == "UserSpecified":
        t.ics = dict(t.ics)
        t.ics["path"] = {"initial_conditions/planeWave3d_e10_sym.npz": "planeWave3d_e10_sym_ic.npz",
                         "planeWave1d.npz": "planeWave1d_ic.npz"}[t.ics["path"]]
    return t


def oracle_streams(name: str, size=None, expanding=None, limit=None):
    t = load_toml(name, size)
    its = list(o.simulation_iter(t, expanding=expanding))
    return its[:limit] if limit else its


def initial_wavefunction(p: o.SimulationParameters) -> np.ndarray:
    return o.initial_wavefunction(p, GOLDEN)


def to_msm_params(p: o.SimulationParameters):
    import msm_b200 as m
    cos = None
    if p.expanding:
        c = p.cosmo_params
        cos = m.CosmologyParameters(c.omega_matter_now, c.omega_radiation_now, c.h, c.z0, c.max_dloga)
    return m.SimulationParameters(axis_length=p.axis_length, final_sim_time=p.final_sim_time, cfl=p.cfl,
                                  num_data_dumps=p.num_data_dumps, total_mass=p.total_mass,
                                  particle_mass=p.particle_mass, hbar_=p.hbar_, k2_cutoff=p.k2_cutoff,
                                  alias_threshold=p.alias_threshold, dims=p.dims, size=p.size, time=p.time,
                                  cosmology=cos)
