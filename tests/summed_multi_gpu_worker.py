"""Coupled (summed-density) mode across ranks: torchrun --nproc-per-node R tests/summed_multi_gpu_worker.py (launched by tests/test_gpu_multi.py)
Each rank owns S/R streams; the density is ncclAllReduce'd inside libmsm_b200 (communicator from a unique id that
rank 0 creates and torch.distributed broadcasts); every rank checks its streams against the CPU oracle ensemble."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
import msm_b200 as m
from msm_b200._lib import lib
from oracle import msm_oracle as o
import golden_util as gu

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def fresh_unique_id():
    """a ncclUniqueId can initialise ONE communicator: rank 0 makes one per context, torch.distributed broadcasts it"""
    uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
    if rank == 0:
        buf = C.create_string_buffer(128)
        assert lib.msm_nccl_unique_id(buf) == 0
        uid.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    return bytes(uid.cpu().numpy().tobytes())


# (name, streams per rank, grid size, chunk_streams, MSM_B200_AR_SLABS): the last two cases force the slab-pipelined
# all-reduce (the default for >= 4 ranks), once with several chunks of local streams accumulating into the same plane
for name, per_rank, size, chunk, slabs in (("spherical-tophat", 3, None, 0, None), ("spherical-tophat-cosmo", 2, None, 0, None),
                                           ("spherical-tophat", 2, 64, 0, None), ("spherical-tophat", 3, 32, 2, "4"),
                                           ("spherical-tophat-cosmo", 2, None, 0, "2")):
    if slabs is None:
        os.environ.pop("MSM_B200_AR_SLABS", None)
    else:
        os.environ["MSM_B200_AR_SLABS"] = slabs
    S = per_rank * world
    ps = gu.oracle_streams(name, size, limit=S)
    psi0s = [gu.initial_wavefunction(p) for p in ps]
    ens = o.SummedEnsemble(ps[0], psi0s)
    uid_bytes = fresh_unique_id()
    sim = m.SimulationObject(gu.to_msm_params(ps[0]), n_streams=per_rank, coupling=m.COUPLING_SUMMED, device=local,
                             chunk_streams=chunk, rank=rank, nranks=world, n_streams_global=S, nccl_unique_id=uid_bytes)
    mine = list(range(rank * per_rank, (rank + 1) * per_rank))
    for li, s in enumerate(mine):
        sim.set_psi(li, psi0s[s])
    for _ in range(4):
        sim.update()
        ens.update()
    worst = 0.0
    for li, s in enumerate(mine):
        got = sim.get_psi(li)
        worst = max(worst, np.linalg.norm((got - ens.streams[s].psi).ravel()) / np.linalg.norm(ens.streams[s].psi.ravel()))
    st = sim.state(0)
    ok = worst < 1e-10 and abs(st.dt - ens.head.last_dt) <= 1e-12 * ens.head.last_dt
    print(f"rank {rank} {name} size {size or 16} S={S} chunk={chunk} slabs={slabs or 'default'}: psi rel-L2 {worst:.2e}, dt {st.dt:.6e} vs {ens.head.last_dt:.6e} -> {'OK' if ok else 'FAIL'}", flush=True)
    sim.close()
    t = torch.tensor([0 if ok else 1], device=f"cuda:{local}")
    dist.all_reduce(t)
    if int(t.item()) != 0:
        dist.destroy_process_group()
        sys.exit(1)
os.environ.pop("MSM_B200_AR_SLABS", None)

# an alias event on ONE rank stops the whole summed ensemble at the same step on EVERY rank (msm_allreduce_max)
p_alias = gu.oracle_streams("spherical-tophat", limit=1)[0]
p_alias.alias_threshold = 1e-6
good = gu.initial_wavefunction(p_alias)
rng = np.random.default_rng(3)
noisy = o.normalize(rng.standard_normal(good.shape) + 1j * rng.standard_normal(good.shape), p_alias.dx, 3)   # alias mass ~ 3.5e-5, the smooth stream ~ 1e-12
sim = m.SimulationObject(gu.to_msm_params(p_alias), n_streams=1, coupling=m.COUPLING_SUMMED, device=local, rank=rank,
                         nranks=world, n_streams_global=world, nccl_unique_id=fresh_unique_id())
sim.set_psi(0, noisy if rank == 0 else good)
try:
    sim.update()
    stopped = False
except m.FourierAliasing:
    stopped = True
ok = stopped and sim.state(0).aliased == 1 and not sim.not_finished()
print(f"rank {rank} alias on rank 0 only: update raised FourierAliasing = {stopped}, own alias mass {sim.state(0).alias_mass:.3e} -> {'OK' if ok else 'FAIL'}", flush=True)
sim.close()
t = torch.tensor([0 if ok else 1], device=f"cuda:{local}")
dist.all_reduce(t)
if int(t.item()) != 0:
    dist.destroy_process_group()
    sys.exit(1)

# ensemble statistics over ALL streams: per-rank accumulation + ncclAllReduce of the four grids inside
# the library, against the restatement of the synthesizer (synthesizer/src/lib.rs:106-342) on the host
from msm_b200 import driver
per_rank = 2
S = per_rank * world
ps = gu.oracle_streams("spherical-tophat", limit=S)
psi0s = [gu.initial_wavefunction(p) for p in ps]
sim = m.SimulationObject(gu.to_msm_params(ps[0]), n_streams=per_rank, device=local, rank=rank, nranks=world,
                         n_streams_global=S, nccl_unique_id=fresh_unique_id())       # independent streams + a communicator
refs = []
for li in range(per_rank):
    sim.set_psi(li, psi0s[rank * per_rank + li])
for s in range(S):
    refs.append(o.SimulationObject(ps[s], psi0s[s]))
for _ in range(3):
    sim.update()
    for r in refs:
        r.update()
got = driver.combine_streams(sim, S, ps[0].dx, allreduce=True)
want = o.synthesizer_combine([r.psi for r in refs], ps[0].dx ** 3)
worst = max(np.linalg.norm((got[k] - want[k]).ravel()) / np.linalg.norm(np.asarray(want[k]).ravel()) for k in ("psi", "psi2", "psik", "psik2"))
norm = float(np.sum(want["psi2"]).real) * ps[0].dx ** 3          # Qx is a difference of O(norm) sums: absolute bound
ok = worst < 1e-10 and abs(got["Qx"] - want["Qx"]) <= 1e-12 * norm
print(f"rank {rank} ensemble over {S} streams on {world} ranks: worst rel-L2 {worst:.2e}, Qx {got['Qx'].real:.6e} vs {want['Qx'].real:.6e} -> {'OK' if ok else 'FAIL'}", flush=True)
sim.close()
t = torch.tensor([0 if ok else 1], device=f"cuda:{local}")
dist.all_reduce(t)
code = int(t.item())
dist.destroy_process_group()
sys.exit(1 if code else 0)
