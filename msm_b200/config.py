"""TOML configuration of a run, in the reference's own format (host driver parity).

Mirrors `common/src/parameters.rs` (TomlParameters :11-55, read_toml :96-107, parse_seeds :148-202,
determine_pmass_hbar_ :222-259), `common/src/ics.rs` and `SimulationIter` (simulator/src/utils/io.rs:115-246):
one stream per seed named "<sim>-stream%05d", then one un-sampled mean-field run named "<sim>".
The file format stays untouched; this module only resolves it into the scalars the C ABI takes.
"""
from __future__ import annotations

import re
import tomllib
from dataclasses import dataclass
from typing import List, Optional

from .api import CosmologyParameters, SimulationParameters

HBAR = 1.757e-90           # common/src/constants.rs:5


@dataclass
class StreamSpec:
    """One item of `SimulationIter` (utils/io.rs:164-245)."""
    sim_name: str
    seed: Optional[int]            # None = the un-sampled mean-field run
    scheme: Optional[str]


@dataclass
class RunConfig:
    parameters: SimulationParameters
    sim_name: str
    ics: dict
    streams: List[StreamSpec]
    n_tot: float
    output_potential: bool


def parse_seeds(s: str) -> List[int]:
    """parameters.rs:148-202: "a..=b", "a to b", "[s1, s2]" or "s1, s2"."""
    if re.search(r"\d+..=\d+", s):
        a, b = (int(x) for x in s.split("..="))
        return list(range(a, b + 1))
    if re.search(r"\d+ to \d+", s):
        a, b = (int(x) for x in s.split(" to "))
        return list(range(a, b + 1))
    found = re.findall(r"(\d+)[^,]?", s)
    if found:
        return [int(x) for x in found]
    raise ValueError("seeds did not match expected patterns: low..=high, low to high, [s1, s2, s3]")


def determine_pmass_hbar_(total_mass: float, ntot, particle_mass, hbar_):
    """parameters.rs:222-259."""
    if ntot is not None:
        pm = total_mass / ntot
        return pm, (hbar_ if hbar_ is not None else HBAR / pm)
    if particle_mass is not None:
        return particle_mass, (hbar_ if hbar_ is not None else HBAR / particle_mass)
    if hbar_ is not None:
        return HBAR / hbar_, hbar_
    raise ValueError("You must specify the total mass and one of ntot, particle_mass or hbar_")


def read_toml(path: str, expanding: Optional[bool] = None) -> RunConfig:
    """parameters.rs:96-107 + utils/io.rs:127-245.  `expanding` mirrors the cargo feature; default: on iff the
    file has a [cosmology] table."""
    with open(path, "rb") as f:
        d = tomllib.load(f)
    opt = lambda k: float(d[k]) if k in d else None
    total_mass = float(d["total_mass"])
    pm, hbar_ = determine_pmass_hbar_(total_mass, opt("ntot"), opt("particle_mass"), opt("hbar_"))
    cosmo = None
    if expanding is None:
        expanding = "cosmology" in d
    if expanding:
        c = d["cosmology"]
        cosmo = CosmologyParameters(float(c["omega_matter_now"]), float(c["omega_radiation_now"]), float(c["h"]),
                                    float(c["z0"]), float(c["max_dloga"]) if "max_dloga" in c else None)
    params = SimulationParameters(
        axis_length=float(d["axis_length"]), final_sim_time=float(d["final_sim_time"]), cfl=float(d["cfl"]),
        num_data_dumps=int(d["num_data_dumps"]), total_mass=total_mass, particle_mass=pm, hbar_=hbar_,
        k2_cutoff=float(d["k2_cutoff"]), alias_threshold=float(d["alias_threshold"]), dims=int(d["dims"]),
        size=int(d["size"]), time=float(d.get("time", 0.0)), cosmology=cosmo)
    name = str(d["sim_name"])
    streams: List[StreamSpec] = []
    if "sampling" in d:
        scheme = str(d["sampling"]["scheme"])
        for seed in parse_seeds(str(d["sampling"]["seeds"])):
            streams.append(StreamSpec(f"{name}-stream{seed:05d}", seed, scheme))      # io.rs:199
    streams.append(StreamSpec(name, None, None))                                      # io.rs:214-240
    return RunConfig(params, name, dict(d["ics"]), streams, total_mass / pm, bool(d.get("output_potential", False)))
