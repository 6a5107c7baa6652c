"""Host driver: `python -m msm_b200 --toml X` mirrors `msm-simulator --toml X` (simulator/src/main.rs:21-89).

The reference runs the streams of a TOML one after another (main.rs:43); here all streams of this rank are resident
on the GPU and advance together, each with its own adaptive time step and dump schedule.  Dumps use the reference's
on-disk layout `sim-data/<sim>[-stream%05d]/psi_%05d_real|_imag` (simulation_object.rs:1155-1158), so
`msm-synthesizer` and the plotting scripts keep working.
"""
from __future__ import annotations

import argparse
import os
import time
from typing import List, Optional

import numpy as np

from .api import COUPLING_INDEPENDENT, SimulationObject
from .config import RunConfig, read_toml


def shard_streams(n_streams: int, rank: int, nranks: int) -> List[int]:
    """Stream indices owned by `rank`: round-robin, so the trailing mean-field run lands on the least loaded rank."""
    return [s for s in range(n_streams) if s % nranks == rank]


def load_initial_conditions(sim: SimulationObject, cfg: RunConfig, local: List[int], base_dir: str = ".") -> None:
    """`new_from_params` (simulation_object.rs:404-435): build the IC of every local stream, then apply the sampler."""
    p = cfg.parameters
    ics = cfg.ics
    g = sim.grid
    kind = ics["type"]
    for li, s in enumerate(local):
        if li == 0:
            if kind == "UserSpecified":                              # ics.rs:650-730
                z = np.load(os.path.join(base_dir, ics["path"]))
                re_, im_ = np.asarray(z["real"], np.float64), np.asarray(z["imag"], np.float64)
                if re_.ndim != p.dims or any(v != p.size for v in re_.shape):
                    raise ValueError("user-provided data does not match dims/size of the toml")
                g.set_psi_planes(0, re_, im_)
            elif kind == "ColdGauss":                                # ics.rs:24-162
                g.ic_cold_gauss(0, [float(v) for v in ics["mean"]], [float(v) for v in ics["std"]])
            elif kind == "ColdGaussKSpace":                          # ics.rs:282-431
                g.ic_cold_gauss_kspace(0, [float(v) for v in ics["mean"]], [float(v) for v in ics["std"]],
                                       int(ics.get("phase_seed") or 0))
            elif kind == "SphericalTophat":                          # ics.rs:165-280
                g.ic_spherical_tophat(0, p.axis_length, float(ics["radius"]), float(ics["delta"]), float(ics["slope"]))
            else:
                raise NotImplementedError(f"ics type {kind} is not available on device")
        else:
            g.ic_copy(li, 0)
    for li, s in enumerate(local):                                   # ics.rs:434-648
        spec = cfg.streams[s]
        if spec.seed is not None:
            g.sample_perturbation(li, spec.scheme, spec.seed, cfg.n_tot)


def run(cfg: RunConfig, out_root: str = "sim-data", rank: int = 0, nranks: int = 1, device: int = 0,
        base_dir: str = ".", verbose: bool = False, write: bool = True, max_updates: Optional[int] = None) -> dict:
    local = shard_streams(len(cfg.streams), rank, nranks)
    sim = SimulationObject(cfg.parameters, n_streams=len(local), coupling=COUPLING_INDEPENDENT, device=device)
    load_initial_conditions(sim, cfg, local, base_dir)
    t0 = time.time()
    def dump(li, s, index):                                          # simulation_object.rs:1113-1180
        sim.dump(li, out_root, cfg.streams[s].sim_name, index)
        if cfg.output_potential:                                     # :1167-1180
            sim.dump_potential(li, out_root, cfg.streams[s].sim_name, index)

    if write:
        sim.reserve_dump_buffers(2)                                  # pin the dump staging before the loop, not inside it
        for li, s in enumerate(local):                               # main.rs:61 dump the initial condition
            dump(li, s, 0)
    updates = 0
    while sim.not_finished() and (max_updates is None or updates < max_updates):   # main.rs:65-69
        sim.update()
        updates += 1
        for li, s in enumerate(local):
            st = sim.state(li)
            if st.dumped and write:
                dump(li, s, st.current_dumps)
    sim.wait_io()
    steps = sum(int(sim.state(li).n_steps) for li in range(len(local)))
    wall = time.time() - t0
    if verbose:
        print(f"rank {rank}: {len(local)} streams, {steps} stream-steps in {wall:.2f} s")
    sim.close()
    return {"streams": len(local), "stream_steps": steps, "seconds": wall}


def export_params(cfg: RunConfig, path: str, base_dir: str = ".") -> None:
    """Write the resolved scalars of a reference TOML as the `key = value` file the native host
    (`msm_b200/msm-simulator-b200 --params`, csrc/msm_simulator.cpp) reads."""
    p = cfg.parameters
    lines = [f"sim_name = {cfg.sim_name}", f"dims = {p.dims}", f"size = {p.size}", f"axis_length = {p.axis_length!r}",
             f"time = {p.time!r}", f"final_sim_time = {p.final_sim_time!r}", f"cfl = {p.cfl!r}",
             f"num_data_dumps = {p.num_data_dumps}", f"total_mass = {p.total_mass!r}",
             f"particle_mass = {p.particle_mass!r}", f"hbar_ = {p.hbar_!r}", f"k2_cutoff = {p.k2_cutoff!r}",
             f"alias_threshold = {p.alias_threshold!r}"]
    c = p.cosmology
    if c is not None:
        lines += ["expanding = 1", f"omega_matter_now = {c.omega_matter_now!r}",
                  f"omega_radiation_now = {c.omega_radiation_now!r}", f"h = {c.h!r}", f"z0 = {c.z0!r}"]
        if c.max_dloga is not None:
            lines.append(f"max_dloga = {c.max_dloga!r}")
    ics = cfg.ics
    if ics["type"] == "ColdGauss":
        lines.append("ics = ColdGauss " + " ".join(repr(float(v)) for v in list(ics["mean"]) + list(ics["std"])))
    elif ics["type"] == "SphericalTophat":
        lines.append(f"ics = SphericalTophat {float(ics['radius'])!r} {float(ics['delta'])!r} {float(ics['slope'])!r}")
    elif ics["type"] == "UserSpecified":
        z = np.load(os.path.join(base_dir, ics["path"]))
        raw = path + ".ic.f64"
        (np.asarray(z["real"], np.float64) + 1j * np.asarray(z["imag"], np.float64)).astype(np.complex128).tofile(raw)
        lines.append(f"ics = File {raw}")
    else:
        raise NotImplementedError(ics["type"])
    if cfg.output_potential:
        lines.append("output_potential = 1")
    seeds = [s.seed for s in cfg.streams if s.seed is not None]
    if seeds:
        lines.append("seeds = " + ",".join(str(s) for s in seeds))
        lines.append(f"scheme = {cfg.streams[0].scheme}")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="msm_b200", description="B200-native msm-simulator time-evolution loop")
    ap.add_argument("--toml", "-t", required=True)
    ap.add_argument("--verbose", "-v", action="store_true")
    ap.add_argument("--test", action="store_true", help="construct the streams and exit (main.rs:59)")
    ap.add_argument("--out", default="sim-data")
    ap.add_argument("--static", action="store_true", help="ignore [cosmology] (cargo feature `expanding` off)")
    ap.add_argument("--export-params", metavar="FILE",
                    help="write the resolved parameters for the native host (msm-simulator-b200 --params FILE) and exit")
    args = ap.parse_args(argv)
    cfg = read_toml(args.toml, expanding=False if args.static else None)
    if args.export_params:
        tdir = os.path.dirname(os.path.abspath(args.toml))
        export_params(cfg, args.export_params, "." if os.path.exists(cfg.ics.get("path", "")) else os.path.dirname(tdir))
        return 0
    rank = int(os.environ.get("RANK", "0"))
    nranks = int(os.environ.get("WORLD_SIZE", "1"))
    device = int(os.environ.get("LOCAL_RANK", "0"))
    base_dir = os.path.dirname(os.path.abspath(args.toml))
    # the reference resolves IC paths relative to the working directory
    base_dir = "." if os.path.exists(cfg.ics.get("path", "")) else os.path.dirname(base_dir)
    res = run(cfg, args.out, rank, nranks, device, base_dir, args.verbose, write=not args.test,
              max_updates=0 if args.test else None)
    if args.verbose or len(cfg.streams) > 1:
        print(f"Finished all streams in {res['seconds']:.0f} seconds")
    return 0


def combine_streams(sim: SimulationObject, n_streams_global: int, dx: float, active=None, allreduce: bool = False) -> dict:
    """The synthesizer's per-dump products from the resident wavefunctions (synthesizer/src/lib.rs:106-342,
    main.rs:63-93,161-173): means over streams of psi, |psi|^2, psi_k, |psi_k|^2 and Qx = sum(<|psi|^2> - |<psi>|^2) dV.
    With several ranks (a context created with an nccl_unique_id), allreduce=True sums the four accumulators over the
    ranks on the device before the division, so every rank holds the ensemble means of ALL streams."""
    sums = sim.grid.ensemble_sums(active, allreduce=allreduce)
    means = {k: v / float(n_streams_global) for k, v in sums.items()}
    dv = dx ** sim.parameters.dims
    means["Qx"] = complex(np.sum(means["psi2"] - means["psi"] * np.conj(means["psi"])) * dv)
    return means


def write_combined(root: str, sim_name: str, dump_index: int, means: dict, dims: int, size: int) -> None:
    """`<root>/<sim>-combined/{psi,psi2,psik,psik2,Qx}_%05d_real|_imag`, NPY, the layout of `dump_complex`
    (synthesizer/src/lib.rs:70-102,282-330)."""
    d = os.path.join(root, f"{sim_name}-combined")
    os.makedirs(d, exist_ok=True)
    shape = (size, size if dims >= 2 else 1, size if dims == 3 else 1, 1)
    for name, arr in means.items():
        a = np.asarray(arr, dtype=np.complex128)
        a = a.reshape(shape) if a.size > 1 else a.reshape(1, 1, 1, 1)
        for part, data in (("real", a.real), ("imag", a.imag)):
            with open(os.path.join(d, f"{name}_{dump_index:05d}_{part}"), "wb") as f:
                np.lib.format.write_array(f, np.ascontiguousarray(data, dtype=np.float64), version=(1, 0))
