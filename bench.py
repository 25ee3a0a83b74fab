#!/usr/bin/env python
"""Headline benchmark: photon packets/s of the photon shooting phases + the traversal (batched DustGrid::path) roofline.

    python bench.py --gpus N --steps K --warmup W              # this repository's engine (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's own CPU code (oracle/_ref)
    python bench.py --config C1|C2|C3|C4|C5 [--scaling weak|strong] ...

The default workload is C2, the configuration BASELINE.json's metric is quoted on (configs[1]: panchromatic Sersic bulge +
exponential disk, 50 wavelengths, InterstellarDustMix, absorption stored, SED + frame instruments, 1e8 packets).  The other
BASELINE configurations run through the same code: C1 (oligochromatic, Cartesian 100^3, 1e6 packets), C3 (adaptive octree
max level 8, spiral disk, 6 peel-off instruments, 1e9), C4 (Voronoi, 1e6 particles, 100 wavelengths, 1e9), C5 (dust emission
+ self-absorption cycles on an adaptive mesh, 1e10 over 8 GPUs).  Their grids are built by the product-side host library
(skirt_b200/libskirthost.so) -- nothing of oracle/ is touched by the engine arm outside the cpu_baseline leg.

One step = the complete shooting of the configuration: the stellar emission phase, and for C5 the self-absorption cycles
(until convergence) and the dust emission phase after it.  Weak scaling (default): every rank shoots `packages` packets per
wavelength; strong: that budget is split over the ranks.  Accumulators are summed over the ranks with NCCL at the points
the reference does it.  Timed on the device with CUDA events on the engine's stream, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "photon packets/sec"
UNIT = "packets/s"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """samples nvidia-smi clocks and throttle reasons while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index; self.proc = None; self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True); self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            w = [v.strip() for v in line.split(",")]
            if len(w) < 6:
                continue
            try:
                sm.append(float(w[0])); mx.append(float(w[1]))
            except ValueError:
                continue
            for nm, v in zip(names, w[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------------------------------
def make_params(args):
    """the parameter dict (skirt_b200.configs) of the selected configuration at the selected size"""
    from skirt_b200 import configs
    c = args.config
    if c == "C1":
        return configs.c1_params(n=args.grid, packages=args.packages)
    if c == "C2":
        return configs.c2_params(n=args.grid, nlambda=args.nlambda, packages=args.packages)
    if c == "C3":
        return configs.c3_params(maxlevel=args.maxlevel, packages=args.packages)
    if c == "C4":
        return configs.c4_params(particles=args.particles, nlambda=args.nlambda, packages=args.packages)
    return configs.c5_params(depth=args.depth, nlambda=args.nlambda, packages=args.packages)


DEFAULTS = {            # packets per wavelength (per GPU when weak-scaled) and wavelengths at BASELINE size
    "C1": dict(packages=1e6, nlambda=1), "C2": dict(packages=2e6, nlambda=50), "C3": dict(packages=1e9, nlambda=1),
    "C4": dict(packages=1e7, nlambda=100), "C5": dict(packages=2.5e7, nlambda=50)}

WORKLOADS = {
    "C1": "C1: OligoMonteCarloSimulation, edge-on ExpDisk stars + dust tau_V=1, CartesianDustGrid {grid}^3, 1 wavelength, FrameInstrument 800x200 at i=88deg",
    "C2": "C2: PanMonteCarloSimulation stellar emission phase, Sersic bulge + ExpDisk stars, ExpDisk dust tau_V=1, CartesianDustGrid {grid}^3, "
          "{nlambda}-point log wavelength grid 0.1-1000 micron, InterstellarDustMix, absorption stored, FrameInstrument 800x200 + SEDInstrument at i=88deg",
    "C3": "C3: adaptive OctTreeDustGrid (levels 2..{maxlevel}, maxMassFraction 1e-6, Neighbor search) of a two-armed spiral ExpDisk, forced scattering, "
          "6 peel-off FrameInstruments 400x400 at i=0,30,60,80,88,90deg, 1 wavelength",
    "C4": "C4: VoronoiDustGrid over {particles} synthetic SPH particles (Voro++ tessellation), ExpDisk stars + dust, {nlambda}-point log wavelength grid, "
          "absorption stored, FrameInstrument 400x400 + SEDInstrument at i=60deg",
    "C5": "C5: PanMonteCarloSimulation with dust emission and self-absorption cycles (until convergence) on a synthetic AdaptiveMesh (root 16^3, 2x2x2 "
          "refinement to depth {depth}), Sersic bulge + ExpDisk stars, {nlambda} wavelengths, FrameInstrument 800x200 + SEDInstrument at i=88deg"}


def workload_config(args, n, extra=None):
    per_gpu = args.packages if args.scaling == "weak" else args.packages / n
    cfg = {"workload": WORKLOADS[args.config].format(**vars(args)), "config": args.config,
           "packets_per_wavelength_per_gpu": per_gpu, "wavelengths": args.nlambda,
           "packets_per_phase": args.packages * args.nlambda * (n if args.scaling == "weak" else 1),
           "parallelism": f"packets sharded over {n} GPU(s) ({args.scaling} scaling); NCCL all-reduce of the stellar absorption table once, the dust "
                          "table per self-absorption cycle, the detector arrays once when read",
           "l2": "256 MiB memset between steps (inside the timed region); the accumulators exceed L2"}
    if extra:
        cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------------------------------------------
def reference_run(p, packages, threads, steps, warmup, dustsamples=10):
    """times the reference's own shooting phases (oracle/_ref: MonteCarloSimulation::runstellaremission and, for C5,
    PanMonteCarloSimulation's self-absorption cycles + dust emission) on `packages` packets per wavelength"""
    from oracle import skirtref as sr, refspec
    if not sr.available():
        raise RuntimeError("oracle/_ref/libskirtref.so is missing (build it where /root/reference exists: make -C oracle ref)")
    spec, L, mixes, extra = refspec.reference_spec(p, threads=threads, dustsamples=dustsamples, packages=packages, with_extra=True)
    S = sr.RefSim(spec, luminosities=L, mixes=mixes, particles=extra.get("particles"), amesh=extra.get("amesh")).setup()
    npp = S.packages_per_lambda(); nl = S.Nlambda
    times, packets = [], []
    for i in range(warmup + steps):
        S.reset(4357 + i)
        sec = S.run_stellar(); n = npp * nl
        if p.get("dustemission"):
            # three stages of one cycle each (a bounded sample of the cycles), then the emission phase
            for stage, factor in enumerate((0.1, 1. / 3., 1.0)):
                S.prepare_dust(stage == 0); sec += S.run_dust(True, factor); n += npp * factor * nl
            S.prepare_dust(False); sec += S.run_dust(False, 1.0); n += npp * nl
        if i >= warmup:
            times.append(sec); packets.append(n)
    total = float(np.sum(times))
    return dict(value=float(np.sum(packets)) / total if total > 0 else 0.0, seconds_per_step=total / max(len(times), 1),
                packets_per_step=float(np.mean(packets)), threads=threads, cells=S.Ncells)


def ref_sample_note(args, r):
    return (f"{r['packets_per_step']:.3g} packets per step ({args.ref_packages:g} per wavelength x {args.nlambda} wavelengths"
            + (", stellar phase + one self-absorption cycle per stage + dust emission" if args.config == "C5" else "")
            + f") of the {args.config} workload, the reference's own shooting phases from oracle/_ref on {r['threads']} threads, {r['seconds_per_step']:.1f} s")


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    p = make_params(args)
    threads = os.cpu_count() or 1
    try:
        r = reference_run(p, args.ref_packages, threads, args.steps, args.warmup)
    except Exception as ex:  # the oracle always exists in a built tree; report why it does not here
        print(json.dumps({"impl": "reference", "unavailable": str(ex).splitlines()[0][:200]}))
        return 0
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds_per_step"], "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, 1),
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "reference", "sample": ref_sample_note(args, r)},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------
def traversal_leg(engine, torch, ext, ncomp, nrays, reps=5, warm=3):
    """batched DustGrid::path()+fillOpticalDepth() on SURVEY.md 8d's synthetic rays (r uniform in 1.2 x the bounding box,
    k isotropic; device-resident inputs and outputs): the record kernel alone (skg_path_fill with given offsets), and
    through the API -- in one traversal per ray (skg_path_batch) and as count + scan + fill"""
    from skirt_b200 import configs
    g = torch.Generator(device="cuda"); g.manual_seed(0x5eed0001)
    box = torch.tensor(configs.C1_BOX, dtype=torch.float64, device="cuda")
    c = 0.5 * (box[0::2] + box[1::2]); w = box[1::2] - box[0::2]
    r = (c + (torch.rand((nrays, 3), generator=g, dtype=torch.float64, device="cuda") - 0.5) * w * 1.2).contiguous()
    k = torch.randn((nrays, 3), generator=g, dtype=torch.float64, device="cuda")
    k = (k / k.norm(dim=1, keepdim=True)).contiguous()
    if os.environ.get("SKG_BENCH_SORT_RAYS"):
        # experiment: the same rays handed over in Morton order of their start point (+ direction octant), the order a caller
        # that bins its rays by cell would use -- shows what ray coherence is worth to the walkers
        q = ((r - box[0::2]) / w).clamp(0, 1 - 1e-9).mul(64).to(torch.int64)
        def spread(v):
            o = torch.zeros_like(v)
            for b in range(6):
                o |= ((v >> b) & 1) << (3 * b)
            return o
        key = ((spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)) << 3) | ((k[:, 0] < 0).long() | ((k[:, 1] < 0).long() << 1) | ((k[:, 2] < 0).long() << 2))
        order = torch.argsort(key)
        r = r[order].contiguous(); k = k[order].contiguous()
    ell = torch.zeros(1, dtype=torch.int32, device="cuda")
    off = torch.zeros(nrays + 1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    total = engine.path_count_device(nrays, r.data_ptr(), k.data_ptr(), off.data_ptr())
    seg = torch.empty(total * 5, dtype=torch.float64, device="cuda")       # 40-byte DustGridPath::Segment records
    torch.cuda.synchronize()

    def timed(fn):
        out = []
        for _ in range(reps):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(ext); fn(); e1.record(ext); e1.synchronize()
            out.append(e0.elapsed_time(e1))
        return float(np.mean(out))

    def fill():
        engine.path_fill_device(nrays, r.data_ptr(), k.data_ptr(), ell.data_ptr(), 0, off.data_ptr(), seg.data_ptr())
    for _ in range(warm):
        fill()
    ms = timed(fill)
    ms_count = timed(lambda: engine.path_count_device(nrays, r.data_ptr(), k.data_ptr(), off.data_ptr()))
    starts = torch.zeros(nrays + 1, dtype=torch.int64, device="cuda"); lens = torch.zeros(nrays, dtype=torch.int32, device="cuda")
    need = engine.path_batch_device(nrays, r.data_ptr(), k.data_ptr(), ell.data_ptr(), 0, starts.data_ptr(), lens.data_ptr(), None, 0)
    del seg
    slab = torch.empty(need * 5, dtype=torch.float64, device="cuda")
    ms_one = timed(lambda: engine.path_batch_device(nrays, r.data_ptr(), k.data_ptr(), ell.data_ptr(), 0, starts.data_ptr(), lens.data_ptr(),
                                                    slab.data_ptr(), need))
    assert int(lens.sum().item()) == total, "one-pass and two-pass traversals disagree on the number of packet-steps"
    del slab
    # context for the roofline: a pure streaming write (memset of 1 GiB) on this GPU -- the path-record kernel is a
    # write-only stream, whereas the roofline denominator (MEASURED_PEAKS.json) is a copy, i.e. reads + writes
    big = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"); wtimes = []
    for _ in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); big.zero_(); e1.record(); e1.synchronize(); wtimes.append(e0.elapsed_time(e1))
    write_only_gbs = (1 << 30) / (min(wtimes[1:]) * 1e-3) / 1e9
    del big
    nbytes = 60.0 * nrays + total * (36.0 + 8.0 * ncomp)
    ms_two = ms + ms_count
    return dict(rays=nrays, packet_steps=int(total), ms=ms, ms_count_pass=ms_count, bytes=nbytes,
                through_api={"one_pass_ms": ms_one, "one_pass_gbs": nbytes / (ms_one * 1e-3) / 1e9, "slab_records": int(need),
                             "two_pass_ms": ms_two, "two_pass_gbs": nbytes / (ms_two * 1e-3) / 1e9,
                             "what": "skg_path_batch (slab capacities + scan + record kernel, one traversal per ray on Cartesian grids) vs skg_path_count + "
                                     "skg_path_fill; device-resident rays and records, same algorithmic bytes"},
                steps_per_s=total / (ms * 1e-3), gbs=nbytes / (ms * 1e-3) / 1e9, written_gbs=40.0 * total / (ms * 1e-3) / 1e9,
                write_only_memset_gbs=write_only_gbs)


def grid_upload_bytes(tabs):
    return sum(np.asarray(v).nbytes for v in tabs.values() if isinstance(v, np.ndarray))


def main_engine(args):
    import torch
    import torch.distributed as dist
    from skirt_b200 import configs

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    p = make_params(args)
    t0 = time.time()
    sim = configs.build(p, device=local, rank=rank, nranks=world)
    sim.packages = args.packages * (world if args.scaling == "weak" else 1)      # the host mirror block-splits this budget over the ranks
    sim.setup()
    e = sim.engine
    pan_dust = bool(p.get("dustemission"))
    if pan_dust:
        sim.setup_dust_library(sim.ds.grid.volumes())
    setup_s = time.time() - t0
    if world > 1:
        from skirt_b200.parallel import share_unique_id
        share_unique_id(e, dist, device="cuda")
    ext = torch.cuda.ExternalStream(e.stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def shoot():
        """the shooting phases of the configuration; returns the per-phase statistics"""
        sim.stats_log.clear(); sim.comm_ms.clear()
        sim.runstellaremission()
        cycles = 0
        if pan_dust:
            cycles = len(sim.rundustselfabsorption(None, cycles=args.cycles))
            sim.rundustemission(None)
        return list(sim.stats_log), cycles

    def step():
        e.reset_results()
        out = shoot()
        with torch.cuda.stream(ext):
            flush.zero_()
        return out

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    launches0 = e.launch_count
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    ev0.record(ext)
    logs = []
    for _ in range(args.steps):
        logs.append(step())
    ev1.record(ext)
    barrier()
    wall = time.perf_counter() - wall0
    ms = ev0.elapsed_time(ev1)
    launches = e.launch_count - launches0
    clocks = sampler.stop()
    comm_ms = {k: (float(np.mean(v)) if isinstance(v, list) else float(v)) for k, v in sim.comm_ms.items()}

    # statistics of one step, summed over its phases (this rank)
    def summed(log, keys):
        return {k: float(sum(st[k] for _, st in log)) for k in keys}
    KEYS = ("packets", "pathSegments", "paths", "scatterings", "absorbSegments", "detections", "peelSegments", "propagateSegments", "iterations",
            "kernel_ms", "launch_ms", "peel_ms", "absorb_ms", "propagate_ms")
    per_step = [summed(log, KEYS) for log, _ in logs]
    st = {k: float(np.mean([s_[k] for s_ in per_step])) for k in KEYS}
    cycles = [c for _, c in logs]
    packets_rank = st["packets"]

    # ---- end to end through the public API with host buffers: upload every table, shoot, read every result back ----
    tabs = sim.ds.grid.tables(); med = sim.ds.medium(); comps = [c.geometry.sampler() for c in sim.ss.comps]
    Lum = sim.ss.luminosities(); instr = [i.d for i in sim.isys.instruments]
    h2d = grid_upload_bytes(tabs) + sum(np.asarray(v).nbytes for v in (med["rho"], med["kext"], med["ksca"], med["g"], Lum))
    e2e_steps = max(1, args.steps if args.e2e_steps is None else args.e2e_steps)
    # the job's results are read back once, by the root process, like the reference does (Instrument::sumResults reduces to
    # the root, which alone writes the output; PeerToPeerCommunicator.cpp:36-50): the other ranks upload their tables, shoot
    # their share and take part in the reductions
    root = rank == 0
    # The read-back is pipelined (skg_results_snapshot + skg_fetch_snapshot_async): step i's arrays travel to page-locked host
    # memory on a second stream while step i+1 uploads and shoots; two sets of host buffers alternate, so a consumer can
    # still read the previous set.  Every transfer is complete before the clock stops (results_end after the last step).
    def read_back(slot):
        if root or world > 1:
            return sim.results_begin(slot)          # every rank takes part in the reduction of the detector arrays
        return None
    for slot in (0, 1):                             # allocates both sets of page-locked buffers (set-up, untimed)
        read_back(slot)
    if root:
        sim.results_end()
    d2h = 0
    e2e_times = []
    barrier()
    w_prev = time.perf_counter()
    for i in range(e2e_steps):
        e.set_grid(tabs); e.medium(med["rho"], med["kext"], med["ksca"], med["g"])
        e.sources(comps, Lum, sim.ss.emissionBias); e.instruments(instr)
        if pan_dust:
            sim.setup_dust_library(sim.ds.grid.volumes())
        e.reset_results()               # every simulation of the series starts from empty accumulators (the previous snapshot is already taken)
        shoot()
        bufs = read_back(i & 1)
        if root:
            d2h = sum(v.nbytes for v in bufs.values())
        if i == e2e_steps - 1:
            t_a = time.perf_counter()
            if root:
                sim.results_end()
            t_b = time.perf_counter()
            torch.cuda.synchronize()
            t_c = time.perf_counter()
            barrier()
            print(f"[bench] e2e tail: results_end {1e3 * (t_b - t_a):.1f} ms, sync {1e3 * (t_c - t_b):.1f} ms, barrier {1e3 * (time.perf_counter() - t_c):.1f} ms", file=sys.stderr)
        w_now = time.perf_counter()
        e2e_times.append(w_now - w_prev); w_prev = w_now

    # ---- reduce over ranks
    def maxr(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

    def sumr(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.SUM); return float(t.item())
    ms = maxr(ms); wall = maxr(wall)
    e2e_times = [maxr(t) for t in e2e_times]
    packets_step = sumr(packets_rank)                  # packets launched per step by all ranks, every phase
    value = packets_step * args.steps / (ms * 1e-3)
    e2e_s = float(np.mean(e2e_times))
    e2e_value = packets_step / e2e_s

    # ---- roofline of the dominant stage kernel of the step (this rank's statistics)
    peak, peak_src = measured_peak()
    ncomp = med["rho"].shape[1] if med["rho"].ndim > 1 else 1
    store = sim.storeabs
    stage_ms = {k: st[k] for k in ("launch_ms", "peel_ms", "absorb_ms", "propagate_ms")}
    absorb_segments_walked = st["pathSegments"] - st["peelSegments"] - st["propagateSegments"]
    absorb_paths = st["packets"] + st["scatterings"]
    peel_paths = st["paths"] - absorb_paths - st["scatterings"]            # propagation walks one path per scattering
    kind = {"cartesian": "GRID_CART", "octtree": "GRID_TREE", "bintree": "GRID_TREE", "amesh": "GRID_AMESH", "voronoi": "GRID_VORO"}[tabs["kind"]]
    if stage_ms["absorb_ms"] >= stage_ms["peel_ms"]:
        dom, dom_ms = f"absorbStage<{kind}> (scatter + traverse + absorb + terminate/sample)", stage_ms["absorb_ms"]
        alg_bytes = 8.0 * ncomp * absorb_segments_walked + 16.0 * st["absorbSegments"] + 192.0 * absorb_paths
        alg_note = "8*Ncomp B density gather per packet-step + 16 B read-modify-write of the absorption table per absorbing step + 192 B of packet record per path"
    else:
        dom, dom_ms = f"peelStage<{kind}> (peel-off towards every observer direction + detection)", stage_ms["peel_ms"]
        alg_bytes = 8.0 * ncomp * st["peelSegments"] + 16.0 * st["detections"] + 96.0 * max(peel_paths, 0.0)
        alg_note = "8*Ncomp B density gather per packet-step + 16 B per detector update + 96 B of packet record per peel-off path"
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "r02_stage_dram_traffic.json")
    if os.path.exists(tfile):
        t_ = json.load(open(tfile)).get(args.config)
        if t_ and t_.get("kernel", "").split("<")[0] == dom.split("<")[0] and abs(t_.get("packets_per_step", 0) / max(packets_rank, 1) - 1) < 0.01:
            traffic = t_["dram_bytes_read"] + t_["dram_bytes_write"]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_note": "DRAM bytes of all launches of this kernel in one step (ncu, profiles/r02_stage_dram_traffic.json) or null when not captured "
                                "for this workload; the density table and the wavelength-major absorption slices in flight stay L2-resident, so DRAM "
                                "traffic is far below the algorithmic bytes",
                "peak_source": peak_src, "bytes_per_step": alg_bytes, "bytes_what": alg_note, "kernel_ms_per_step": dom_ms,
                "launches_per_step": int(st["iterations"]), "share_of_step": dom_ms / max(st["kernel_ms"], 1e-9),
                "note": "not DRAM-bound: what limits the stage kernels is the rate at which the SM's load/store path and L2 retire scattered 8-byte "
                        "gathers and fp64 atomics (see `atomics`), and the latency of the dependent gather; the HBM-bound kernel is the path-record kernel "
                        "(`traversal_roofline`)"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, world, {"cells": int(e.Ncells), "packets_per_step_all_phases": packets_step}),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "s_per_step": e2e_times, "of_device_value": e2e_value / value,
                    "what": "skg_grid_*+skg_medium+skg_sources+skg_instruments from host arrays (every rank), the shooting phases, "
                            "skg_results_snapshot + skg_fetch_snapshot_async of frame/SED/absorption table into page-locked host arrays (root rank, after "
                            "the reduction), overlapped with the next step; the series ends with every transfer complete"},
            "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": wall, "setup_s": setup_s,
            "roofline": roofline,
            "stage_ms_per_step": dict(stage_ms, kernel_ms=st["kernel_ms"]),
            "packet_steps_per_s": st["pathSegments"] / (st["kernel_ms"] * 1e-3),
            "per_step_stats": {k: int(v) for k, v in st.items() if not k.endswith("_ms")}}
    if pan_dust:
        line["selfabs_cycles_per_step"] = cycles
    if world > 1:
        line["allreduce_ms"] = comm_ms
        if "labs_dust_cycles" in comm_ms:
            line["allreduce_ms_per_cycle"] = comm_ms["labs_dust_cycles"]

    if rank == 0 and store and not args.skip_atomics:
        # the ceiling of the escape + absorption stage: fp64 atomic adds to random cells of one wavelength slice
        rate = e.selftest_atomics(1 << 31, int(e.Ncells))
        line["atomics"] = {"absorb_atomics_per_s": st["absorbSegments"] / (stage_ms["absorb_ms"] * 1e-3), "measured_peak_per_s": rate,
                           "frac": st["absorbSegments"] / (stage_ms["absorb_ms"] * 1e-3) / rate,
                           "what": "fp64 atomicAdd (RED.E.ADD.F64) to pseudo-random cells of a table of Ncells doubles, nothing else in the kernel "
                                   "(skg_selftest_atomics), against the atomics the absorb stage retires per second of its own device time"}
    if rank == 0 and not args.skip_traversal:
        tr = traversal_leg(e, torch, ext, ncomp, args.rays)
        tr["frac"] = tr["gbs"] / peak; tr["peak"] = peak
        tr["through_api_frac"] = tr["through_api"]["one_pass_gbs"] / peak; tr["through_api_two_pass_frac"] = tr["through_api"]["two_pass_gbs"] / peak
        tr["kernel"] = f"pathFillKernel<{kind}> (batched DustGrid::path + fillOpticalDepth, path records)"
        tfile = os.path.join(ROOT, "profiles", "r02_path_dram_traffic.json")
        tr["traffic"] = None
        if os.path.exists(tfile):
            t_ = json.load(open(tfile)).get(kind)
            if t_:
                tr["traffic"] = t_["dram_bytes_per_packet_step"] * tr["packet_steps"]
                tr["traffic_note"] = t_.get("note")
        line["traversal_roofline"] = tr
    if rank == 0 and world == 1 and not args.skip_cpu:
        threads = os.cpu_count() or 1
        try:
            r = reference_run(p, args.ref_packages, threads, steps=1, warmup=0)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "reference", "sample": ref_sample_note(args, r)}
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": f"unavailable: {ex}"}
    if world > 1:
        dist.barrier()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(DEFAULTS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --packages per wavelength on EVERY GPU; strong: --packages per wavelength split over the GPUs")
    ap.add_argument("--packages", type=float, default=None, help="packets per wavelength (C2: 2e6 x 50 wavelengths = 1e8)")
    ap.add_argument("--nlambda", type=int, default=None)
    ap.add_argument("--grid", type=int, default=100, help="C1/C2: cells per axis of the Cartesian grid")
    ap.add_argument("--maxlevel", type=int, default=8, help="C3: maximum octree level")
    ap.add_argument("--particles", type=int, default=1000000, help="C4: SPH particles = Voronoi cells")
    ap.add_argument("--depth", type=int, default=6, help="C5: refinement depth of the adaptive mesh below its 16^3 root cells")
    ap.add_argument("--cycles", type=int, default=0, help="C5: self-absorption cycles per stage (0: until convergence, like the reference)")
    ap.add_argument("--rays", type=int, default=None, help="rays of the traversal-roofline leg (SURVEY.md 8d: 2^24 on the Cartesian grid)")
    ap.add_argument("--ref-packages", type=float, default=None, help="packets per wavelength of the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=None, help="end-to-end steps (default: --steps)")
    ap.add_argument("--skip-traversal", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-atomics", action="store_true")
    args = ap.parse_args()
    d = DEFAULTS[args.config]
    if args.packages is None:
        args.packages = d["packages"]
    if args.nlambda is None:
        args.nlambda = d["nlambda"]
    if args.config in ("C1", "C3"):
        args.nlambda = 1
    if args.rays is None:
        args.rays = 1 << 24 if args.config in ("C1", "C2") else 1 << 22
    if args.ref_packages is None:
        # about 10-30 s of CPU work per step: ~3e4 packets/s/core measured on this class of host for C2; the other grids
        # and the 6-instrument configuration are several times slower per packet
        cores = os.cpu_count() or 1
        slow = {"C1": 1.0, "C2": 1.0, "C3": 4.0, "C4": 8.0, "C5": 6.0}[args.config]
        args.ref_packages = float(min(args.packages, max(2e3, round(15.0 * 3.0e4 * cores / args.nlambda / slow, -3))))
    # stdout carries exactly one JSON line: libraries that write to fd 1 (NCCL prints its version banner there)
    # are sent to stderr for the whole run, and the line goes out on the saved descriptor
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(saved, "w", buffering=1)
    if args.impl == "reference":
        return main_reference(args)
    return main_engine(args)


if __name__ == "__main__":
    sys.exit(main())
