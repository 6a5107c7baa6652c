"""Host-side logic of the product (no GPU): cosmology scalars in libmsm_b200.so vs the oracle, TOML resolution,
stream sharding, and a world_size-2 gloo run of the multi-rank bookkeeping bench.py uses."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

import msm_b200 as m
from msm_b200 import config as mc
from msm_b200.driver import shard_streams
from oracle import msm_oracle as o
from conftest import ROOT
from golden_util import GOLDEN, load_toml

COSMOS = [(0.7, 0.0, 0.7, 1.0, 0.01), (1.0, 0.0, 1e-7, 99.0, 0.01), (0.3, 1e-4, 0.67, 20.0, None),
          (0.25, 0.05, 0.72, 3.0, 1e-3)]


@pytest.mark.parametrize("om,orad,h,z0,mdl", COSMOS)
def test_cosmology_scalars_match_oracle(om, orad, h, z0, mdl):
    c = m.CosmologyParameters(om, orad, h, z0, mdl)
    oc = o.CosmologyParameters(om, orad, h, z0, mdl)
    for t in (0.0, 0.3, 40.0, 2000.0):
        a, b = m.get_tau(t, c), o.get_tau(t, oc)               # simulation_object.rs:1408-1453
        assert abs(a - b) <= 1e-13 * max(abs(b), 1e-300)
    assert m.get_supercomoving_boxsize(0.05, c, 30.0) == o.get_supercomoving_boxsize(0.05, oc, 30.0)
    s = o.ScaleFactorSolver(oc)
    want = s.step(123.4)
    got = m.scale_factor_after(123.4, c)
    assert abs(got - want) <= 1e-14 * want


def test_toml_resolution_matches_oracle(tmp_path):
    text = """
axis_length = 60.0
final_sim_time = 2000.0
cfl = 0.1
num_data_dumps = 64
total_mass = 3e+16
hbar_ = 0.01
sim_name = "planeWave3d_e10_sym"
ntot = 10000000000.0
k2_cutoff = 0.95
alias_threshold = 0.001
dims = 3
size = 16
[ics]
type = "UserSpecified"
path = "initial_conditions/planeWave3d_e10_sym.npz"
[cosmology]
omega_matter_now = 1.0
omega_radiation_now = 0.0
h = 1e-07
z0 = 99.0
max_dloga = 0.01
[sampling]
num_streams = 16
seeds = "1 to 16"
scheme = "Wigner"
"""
    f = tmp_path / "p.toml"
    f.write_text(text)
    cfg = mc.read_toml(str(f))
    ot = o.read_toml(str(f))
    ops = list(o.simulation_iter(ot))
    assert [s.sim_name for s in cfg.streams] == [p.sim_name for p in ops]          # io.rs:199,214-240
    assert [s.seed for s in cfg.streams] == [p.sampling_parameters["seed"] if p.sampling_parameters else None for p in ops]
    p0 = ops[0]
    assert cfg.parameters.particle_mass == p0.particle_mass and cfg.parameters.hbar_ == p0.hbar_
    assert cfg.n_tot == p0.n_tot and cfg.parameters.cosmology.z0 == 99.0
    static = mc.read_toml(str(f), expanding=False)
    assert static.parameters.cosmology is None


def test_parse_seeds_product():
    assert mc.parse_seeds("0..=55") == list(range(56))            # common/src/parameters.rs:121-144
    assert mc.parse_seeds("0 to 55") == list(range(56))
    assert mc.parse_seeds("[1, 3]") == [1, 3] and mc.parse_seeds("1, 3") == [1, 3]
    with pytest.raises(ValueError):
        mc.parse_seeds("abc")


def test_pmass_hbar_resolution():
    # common/src/parameters.rs:222-259
    assert mc.determine_pmass_hbar_(1e10, 1e5, None, None) == (1e5, mc.HBAR / 1e5)
    assert mc.determine_pmass_hbar_(1e10, 1e5, None, 0.02) == (1e5, 0.02)
    assert mc.determine_pmass_hbar_(1e10, None, 3.0, None) == (3.0, mc.HBAR / 3.0)
    assert mc.determine_pmass_hbar_(1e10, None, None, 0.02) == (mc.HBAR / 0.02, 0.02)
    with pytest.raises(ValueError):
        mc.determine_pmass_hbar_(1e10, None, None, None)


def test_shard_streams_partition():
    for n, r in ((11, 2), (65, 8), (17, 4), (3, 8)):
        parts = [shard_streams(n, k, r) for k in range(r)]
        flat = sorted(s for p in parts for s in p)
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_two_rank_gloo_bookkeeping(tmp_path):
    """world_size 2 over gloo: each rank takes its shard of streams, the per-rank step times are MAX-reduced and
    rank 0 aggregates cell-updates/s -- the same reduction bench.py performs over NCCL."""
    script = tmp_path / "w.py"
    script.write_text(f"""
import os, sys, json
sys.path.insert(0, {ROOT!r})
import torch, torch.distributed as dist
from msm_b200.driver import shard_streams
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
mine = shard_streams(11, r, w)
t = torch.tensor([0.5 + 0.25 * r], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
cnt = torch.tensor([len(mine)], dtype=torch.int64)
dist.all_reduce(cnt)
got = [None] * w
dist.all_gather_object(got, mine)
if r == 0:
    print(json.dumps({{"max_t": float(t), "streams": int(cnt), "shards": got}}))
dist.destroy_process_group()
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["max_t"] == 0.75 and d["streams"] == 11
    assert sorted(d["shards"][0] + d["shards"][1]) == list(range(11))


def test_reference_arm_of_bench_runs_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-size", "32"], capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["metric"] == "cell-updates/s"


@pytest.mark.parametrize("n,chunk", [(64, 8), (8, 8), (1, 2), (3, 2), (11, 2), (16, 4), (17, 4), (33, 8), (5, 6), (100, 16), (2, 8)])
def test_run_streams_group_schedule(n, chunk):
    """Stream groups of msm_sim_run_streams (the pipelined outer loop of main.rs:43-85): a partition of the stream list
    in order, no group larger than the launch chunk, pairs never split (even group sizes except a ragged last piece),
    and short first / last groups when there is something to overlap them with."""
    import ctypes as C
    from msm_b200._lib import lib
    buf = (C.c_int32 * 128)()
    g = lib.msm_run_groups(n, chunk, buf, 128)
    b = list(buf[:g + 1])
    sizes = [b[i + 1] - b[i] for i in range(g)]
    assert b[0] == 0 and b[-1] == n and all(s > 0 for s in sizes) and max(sizes) <= max(chunk, 2)
    assert all(s % 2 == 0 for s in sizes[:-1]) or n % 2 == 1
    starts_even = all(x % 2 == 0 for x in b[:-1]) or n % 2 == 1
    assert starts_even
    if n >= 4 * chunk and chunk >= 4:
        assert sizes[0] == 2 and sizes[-1] == 2 and sizes[1] == chunk - 2
    if n >= 8:
        assert g >= 4                                             # enough groups to overlap transfers with compute
