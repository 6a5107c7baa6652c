#!/usr/bin/env python
"""bench.py -- throughput of the MSM time-evolution loop on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W          (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W  (the reference's CPU algorithm on the host cores)

Metric (BASELINE.json): cell-updates/s = cells x streams x steps / seconds of the step loop, where one step is
one `SimulationObject::update()` (simulation_object.rs:475-661) of every stream: potential at t, adaptive dt (host
sync), drift, inverse FFT, potential, kick, forward FFT, drift, alias check (host sync).
Workload: BASELINE.json configs[4] -- synthetic 512^3 x 64 streams, fp64, physical scalars and ColdGauss base field
of examples/gaussian-overdensity-mft.toml, per-stream Wigner noise (seed = stream id), static box, independent
streams (the reference's semantics), stream-sharded across the ranks (64 / N each, no data-path collective).
Every stream is 2 GiB >> 126 MB of L2, so no L2 flush is needed between steps.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def claim_host_cores():
    """The CPU legs (--impl reference, cpu_baseline) use every host core.  torchrun exports OMP_NUM_THREADS=1 to its
    workers, which throttles NumPy's OpenBLAS pool (measured: the same oracle step ran 3x slower under torchrun in
    round 1), so the thread-count variables are set BEFORE NumPy is imported, and the affinity mask is widened to all
    cores where the container allows it.  Returns the number of cores this process may run on."""
    n = os.cpu_count() or 1
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(n)
    try:
        os.sched_setaffinity(0, range(n))
    except (OSError, AttributeError):
        pass
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return n


# rank 0 of the reference arm and the single-rank GPU arm (cpu_baseline leg) run CPU work; GPU ranks under torchrun do not
HOST_CORES = (claim_host_cores() if ("reference" in sys.argv or int(os.environ.get("WORLD_SIZE", "1")) == 1)
              else (os.cpu_count() or 1))

import numpy as np  # noqa: E402

ALG_BYTES_PER_CELL_UPDATE = 480.0       # SURVEY 8d / DESIGN.md section 3: 3-pass transforms, un-fused pass boundaries


def alg_bytes(coupling, s_local):
    """algorithmic HBM bytes per cell-update of the 3-pass model: 480 independent streams;
    272 + 216 / S_local with the summed density (one real-field solve per potential for all local streams)"""
    return 480.0 if coupling == "independent" else 272.0 + 216.0 / s_local


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="msm_b200", choices=["msm_b200", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--streams", type=int, default=64, help="total streams over all ranks")
    ap.add_argument("--chunk", type=int, default=8)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-summed", action="store_true", help="skip the summed-density (coupled) sub-record")
    ap.add_argument("--cpu-size", type=int, default=0, help="grid size of the CPU baseline sample (0 = auto)")
    ap.add_argument("--coupling", default="independent", choices=["independent", "summed"],
                    help="independent = the reference's semantics (headline); summed = north-star variant with one "
                         "ncclAllReduce of the density per potential")
    return ap.parse_args()


N_TOT_SAMPLER = 1e10        # Wigner noise with n_tot = 1e10


def workload_params(size):
    """examples/gaussian-overdensity-mft.toml resolved (tests/golden/configs.json holds the same numbers)."""
    import msm_b200 as m
    hbar_ = 0.02
    # (particle_mass only fixes n_tot = total_mass / particle_mass of the sampler: 1e10)
    return m.SimulationParameters(axis_length=30.0, final_sim_time=400.0, cfl=0.02, num_data_dumps=200,
                                  total_mass=1e10, particle_mass=1e10 / N_TOT_SAMPLER, hbar_=hbar_, k2_cutoff=0.95,
                                  alias_threshold=0.02, dims=3, size=size, time=0.0, cosmology=None)




class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                power.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_step_rate(size, steps, warmup):
    """The oracle's un-fused 7-FFT update() (the reference's algorithm) on the host cores: cell-updates/s."""
    from oracle import msm_oracle as o
    hbar_ = 0.02
    p = o.SimulationParameters(axis_length=30.0, time=0.0, final_sim_time=400.0, cfl=0.02, num_data_dumps=200,
                               total_mass=1e10, particle_mass=o.HBAR / hbar_, sim_name="bench", k2_cutoff=0.95,
                               alias_threshold=0.02, hbar_=hbar_, dims=3, size=size,
                               ics={"type": "ColdGauss", "mean": [15.0] * 3, "std": [10.0] * 3})
    psi0 = o.cold_gauss([15.0] * 3, [10.0] * 3, p)
    sim = o.SimulationObject(p, psi0)
    for _ in range(warmup):
        sim.update()
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.update()
    dt = time.perf_counter() - t0
    return size ** 3 * steps / dt, dt / steps


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm (oracle restatement: the Rust/ArrayFire binary cannot be
    built here) with all host threads.  Each step is a bounded sample of the workload: one update() of ONE of its
    streams on the workload's own grid (512^3; cell-updates/s is a per-cell rate, so one stream measures it).  If a
    512^3 step turns out too slow for the --steps/--warmup asked for (more than ~10 minutes in all), the sample drops
    to 256^3 and the line says so."""
    if rank != 0:
        return
    from oracle import msm_oracle as o
    o.set_workers(HOST_CORES)
    size = args.cpu_size or args.size
    note = ""
    if not args.cpu_size and size > 256:
        t0 = time.perf_counter()
        probe_rate, _ = cpu_step_rate(256, 1, 1)
        est = (size ** 3 / probe_rate) * (args.steps + args.warmup)
        if est > 600.0:
            note = f" (a {size}^3 step would take {size ** 3 / probe_rate:.0f} s: sample reduced to 256^3)"
            size = 256
    rate, sec = cpu_step_rate(size, args.steps, args.warmup)
    sample = (f"1 of the {args.streams} streams x {size}^3 per step{note}, NumPy/pocketfft restatement of the "
              f"reference's un-fused update(), {o.get_workers()} worker threads")
    line = {"impl": "reference", "metric": "cell-updates/s", "value": rate, "unit": "cell-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic {args.size}^3 x {args.streams} streams fp64 static box (BASELINE configs[4])",
                       "coupling": args.coupling, "sample": sample, "sample_grid": size, "sample_streams": 1,
                       "workers_used": o.get_workers(), "omp_num_threads": os.environ.get("OMP_NUM_THREADS")},
            "cpu_baseline": {"value": rate, "unit": "cell-updates/s", "cores": HOST_CORES, "kind": "port",
                             "sample": sample},
            "e2e": {"value": rate, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def build_streams(sim, n_local, first_global, size):
    g = sim.grid
    g.ic_cold_gauss(0, [15.0] * 3, [10.0] * 3)
    for s in range(1, n_local):
        g.ic_copy(s, 0)
    for s in range(n_local):
        g.sample_perturbation(s, "Wigner", first_global + s + 1, N_TOT_SAMPLER)
    g.synchronize()


def bind_to_gpu_numa_node(device):
    """Run this rank on the CPU cores next to its GPU (sysfs local_cpulist of the GPU's PCI function), so that the pinned
    host buffers of the end-to-end leg are allocated on that NUMA node: with 8 ranks on one box the transfers otherwise
    all cross the same socket.  Best effort; returns what was done for the JSON line."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return f"cpus {text} (gpu {bdf})"
        return f"unchanged ({text or 'no local_cpulist'})"
    except Exception as e:          # containers without sysfs access, older torch, ...
        return f"unavailable ({type(e).__name__})"


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line on the process's real stdout"""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # Libraries write banners to stdout at the file-descriptor level (NCCL prints "NCCL version ..." when NCCL_DEBUG is
    # set): everything but the JSON line goes to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import msm_b200 as m

    # (single rank: keep all cores -- the CPU baseline leg below runs on them)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "all cores (single rank)"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank
    n_total = args.streams
    if n_total % world:
        raise SystemExit("--streams must be divisible by the number of ranks")
    n_local = n_total // world
    size = args.size
    cells = size ** 3
    params = workload_params(size)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    coupling = m.COUPLING_SUMMED if args.coupling == "summed" else m.COUPLING_INDEPENDENT

    def comm_kwargs(cpl=None):
        """a fresh ncclUniqueId per context (rank 0 creates, torch.distributed broadcasts); summed mode only"""
        if (coupling if cpl is None else cpl) != m.COUPLING_SUMMED or world == 1:
            return {}
        import ctypes as C
        from msm_b200._lib import lib
        uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local_rank}")
        if rank == 0:
            buf = C.create_string_buffer(128)
            assert lib.msm_nccl_unique_id(buf) == 0
            uid.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        return dict(rank=rank, nranks=world, n_streams_global=n_total, nccl_unique_id=bytes(uid.cpu().numpy().tobytes()))

    # ---- resident run: streams generated on the device, timed with CUDA events on the library's stream ----------
    def resident(coupling_name):
        cpl = m.COUPLING_SUMMED if coupling_name == "summed" else m.COUPLING_INDEPENDENT
        sim, note, chunk = None, "", None
        for chunk in (args.chunk, 4, 2):
            try:
                sim = m.SimulationObject(params, n_streams=n_local, device=device, chunk_streams=chunk, coupling=cpl,
                                         **comm_kwargs(cpl))
                break
            except m.MsmError as e:
                if e.code != -7:
                    raise
                note = f"chunk {chunk} did not fit"
        if sim is None:
            raise SystemExit("streams do not fit in device memory: " + note)
        g = sim.grid
        build_streams(sim, n_local, rank * n_local, size)
        for _ in range(args.warmup):
            sim.update()
        g.synchronize()
        launches0 = g.launch_count()
        g.profile_enable(True)
        clocks = ClockSampler(device)
        clocks.start()
        barrier()
        g.timer_start()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            sim.update()
        ms = g.timer_stop()
        wall = time.perf_counter() - t0
        barrier()
        clk = clocks.stop()
        launches = g.launch_count() - launches0
        prof = g.profile_read()
        g.profile_enable(False)
        ms = max_over_ranks(ms)
        st = sim.state(0)
        assert st.n_steps == args.warmup + args.steps and not st.aliased
        return dict(sim=sim, ms=ms, wall=wall, clk=clk, launches=launches, prof=prof, chunk=chunk,
                    value=cells * n_total * args.steps / (ms * 1e-3))

    head = resident(args.coupling)
    sim, g, ms, wall, clk, launches, prof, chunk, value = (head[k] for k in ("sim", "sim", "ms", "wall", "clk", "launches",
                                                                             "prof", "chunk", "value"))
    g = sim.grid

    # ---- roofline of the dominant kernel (CUDA events around every launch, same timed region) ------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    fft = [r for r in prof if r["name"].startswith("fft_")]
    top = max(fft, key=lambda r: r["ms_total"])
    ach = top["algorithmic_bytes"] / (top["ms_total"] * 1e-3) / 1e9
    kernel_ms = sum(r["ms_total"] for r in prof)
    # DRAM traffic of that kernel per launch, from the committed `ncu --set full` capture of the same launch shape
    # (profiles/r02g_ncu_dram_traffic.json: 8 streams per launch, 512^3); null when the shape differs
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r02g_ncu_dram_traffic.json")
    if os.path.exists(tpath) and size == 512 and chunk == 8 and n_local >= 8:
        key = top["name"][:-3] + (",x>" if top["name"].endswith(",x>") else ",yz>")
        for k, v in json.load(open(tpath)).items():
            if k.startswith(key):
                traffic, traffic_src = v["dram_bytes_per_launch"], "profiles/r02g_ncu_dram_traffic.json"
    roofline = {"bound": "hbm", "kernel": top["name"], "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src, "traffic_measured_in_run": False,
                "algorithmic_bytes_per_launch": top["algorithmic_bytes"] / top["launches"], "peak_source": peak_src,
                "avg_launch_ms": top["ms_total"] / top["launches"], "share_of_step": top["ms_total"] / kernel_ms,
                "step": {"algorithmic_bytes_per_cell_update": alg_bytes(args.coupling, n_local),
                         "achieved_per_gpu": value / world * alg_bytes(args.coupling, n_local) / 1e9,
                         "frac": value / world * alg_bytes(args.coupling, n_local) / 1e9 / peak,
                         "frac_of_8TBps": value / world * alg_bytes(args.coupling, n_local) / 8e12},
                "kernels": [{"name": r["name"], "launches": r["launches"], "ms": round(r["ms_total"], 3),
                             "GBps": round(r["algorithmic_bytes"] / (r["ms_total"] * 1e-3) / 1e9, 1)}
                            for r in sorted(prof, key=lambda r: -r["ms_total"])]}

    # ---- end to end through the C ABI with host buffers --------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        host = torch.empty(2 * cells, dtype=torch.float64, pin_memory=True)
        out_re = torch.empty(cells, dtype=torch.float64, pin_memory=True)
        out_im = torch.empty(cells, dtype=torch.float64, pin_memory=True)
        hnp, renp, imnp = host.numpy(), out_re.numpy(), out_im.numpy()
        hnp[:] = g.get_psi(0).reshape(-1).view(np.float64)        # a realistic wavefunction as the host-side IC
    sim.close()

    # ---- the north-star's coupled mode beside the headline: |psi|^2 summed over ALL streams (local accumulate ->
    #      ncclAllReduce of the real density over NVLink -> replicated real-field Poisson solve), same streams, same steps
    def summed_record():
        r = resident("summed")
        r["sim"].close()
        sp = r["prof"]
        per = lambda pred: sum(x["ms_total"] for x in sp if pred(x["name"])) / args.steps       # noqa: E731
        sb = alg_bytes("summed", n_local)
        ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm = float(json.load(open(ppath))["hbm_gbs"]) if os.path.exists(ppath) else 6650.0
        return {"value": r["value"], "unit": "cell-updates/s", "ms_per_step": r["ms"] / args.steps, "steps": args.steps,
                "allreduce_ms_per_step": per(lambda nm: nm.startswith("nccl_")),
                "poisson_ms_per_step": per(lambda nm: ",half>" in nm or ",nyquist>" in nm or "poisson" in nm
                                           or nm.startswith(("pack_", "unpack_"))),
                "algorithmic_bytes_per_cell_update": sb, "step_frac": r["value"] / world * sb / 1e9 / hbm,
                "gpu_launches": int(r["launches"]), "clocks": r["clk"],
                "what": "rho = (A/S) sum over all streams of |psi_s|^2: accumulated over the local streams inside the "
                        "last inverse pass (real plane, 8 B/cell), ncclAllReduce of n^3 doubles in place, replicated "
                        "real-field (R2C/C2R half-spectrum) Poisson solve; twice per step (kick potential, dt potential)",
                "kernels": [{"name": x["name"], "launches": x["launches"], "ms": round(x["ms_total"], 3),
                             "GBps": round(x["algorithmic_bytes"] / (x["ms_total"] * 1e-3) / 1e9, 1)}
                            for x in sorted(sp, key=lambda x: -x["ms_total"])]}

    summed = None
    if args.coupling == "independent" and not args.no_summed:
        try:
            summed = summed_record()
        except Exception as exc:      # the headline line must survive a failure of the side record
            summed = {"error": f"{type(exc).__name__}: {exc}"}
    if not args.no_e2e:
        sim = m.SimulationObject(params, n_streams=n_local, device=device, chunk_streams=chunk, coupling=coupling,
                                 **comm_kwargs())
        g = sim.grid
        barrier()
        t0 = time.perf_counter()
        # msm_sim_run_streams = the reference's outer loop (main.rs:43-85) through the C ABI with host buffers: H2D of
        # every stream's IC from pinned memory, K update() per stream, D2H of every stream's final psi as re/im planes
        # (one dump).  Streams advance in groups; transfers of neighbouring groups overlap the step kernels.
        ids = list(range(n_local))
        if coupling == m.COUPLING_INDEPENDENT:
            sim.run_streams(ids, [hnp] * n_local, [renp] * n_local, [imnp] * n_local, max_updates=args.steps)
        else:       # coupled streams advance together: upload all, K x update(), pipelined dump of all
            for s in ids:
                sim.set_psi(s, hnp.view(np.complex128))
            for _ in range(args.steps):
                sim.update()
            g.get_psi_many(ids, [renp] * n_local, [imnp] * n_local)
        sec = time.perf_counter() - t0
        barrier()
        sec = max_over_ranks(sec)
        # per update: drift coefficient + stream id per stream read by the device from mapped host memory (the drift
        # tables themselves are built on the device), alias mass + max|phi| per stream written back the same way
        e2e = {"value": cells * n_total * args.steps / sec, "unit": "cell-updates/s",
               "h2d_bytes_per_step": int(16 * cells * n_local / args.steps + 12 * n_local),
               "d2h_bytes_per_step": int(16 * cells * n_local / args.steps + 16 * n_local),
               "seconds": sec, "what": "msm_sim_run_streams (upload of every stream's IC from pinned host memory, K x "
               "update() per stream, download of every stream's final psi as re/im planes) inside the timed region; "
               "transfers of neighbouring stream groups overlap the step kernels"}
        assert sim.state(n_local - 1).n_steps == args.steps and np.isfinite(renp[:16]).all()
        sim.close()
        def seeded_record():
            # second flavour: the same loop with the initial conditions built on the device from (IC spec, seed) -- what
            # the reference's own `new_from_params` does -- so that only the dump half crosses PCIe
            sim2 = m.SimulationObject(params, n_streams=n_local, device=device, chunk_streams=chunk, coupling=coupling)
            try:
                g2 = sim2.grid
                g2.ic_cold_gauss(0, [15.0] * 3, [10.0] * 3)
                g2.ic_store(0)
                g2.synchronize()
                barrier()
                t1 = time.perf_counter()
                sim2.run_streams_seeded(ids, "Wigner", [rank * n_local + s + 1 for s in ids], [renp] * n_local,
                                        [imnp] * n_local, max_updates=args.steps)
                sec2 = max_over_ranks(time.perf_counter() - t1)
                barrier()
                assert sim2.state(n_local - 1).n_steps == args.steps and np.isfinite(renp[:16]).all()
            finally:
                sim2.close()
            return {"value": cells * n_total * args.steps / sec2, "unit": "cell-updates/s", "seconds": sec2,
                    "h2d_bytes_per_step": int(20 * n_local + 12 * n_local),
                    "d2h_bytes_per_step": int(16 * cells * n_local / args.steps + 16 * n_local),
                    "what": "msm_sim_run_streams_seeded: un-sampled IC saved on the device + seeded Wigner sampler "
                            "per stream (device), K x update(), download of every stream's final psi"}

        if coupling == m.COUPLING_INDEPENDENT:
            try:
                e2e["seeded"] = seeded_record()
            except Exception as exc:      # the host-buffer figure above must survive a failure of this flavour
                e2e["seeded"] = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- the reference's CPU algorithm beside it (rank 0, N = 1 only) ------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import msm_oracle as o
        o.set_workers(HOST_CORES)
        csize = args.cpu_size or size
        rate, sec = cpu_step_rate(csize, 2, 1)
        cpu = {"value": rate, "unit": "cell-updates/s", "cores": HOST_CORES, "kind": "port",
               "sample": f"1 of the {n_total} streams x {csize}^3, 2 update() after 1 warm-up ({sec:.2f} s/step), "
                         "NumPy/pocketfft restatement of the reference's un-fused 7-FFT step, "
                         f"{o.get_workers()} worker threads"}

    if rank == 0:
        line = {"metric": "cell-updates/s", "value": value, "unit": "cell-updates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"synthetic {size}^3 x {n_total} streams fp64 static box (BASELINE configs[4])",
                           "streams_per_gpu": n_local, "coupling": args.coupling, "chunk_streams": chunk,
                           "l2": "inputs larger than L2 (2 GiB per stream vs 126 MB)",
                           "timing": "CUDA events on the library stream, max over ranks", "wall_s": wall,
                           "cpu_affinity": numa},
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "summed": summed}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
