#!/usr/bin/env python
"""Headline benchmark: photon packets/s of the stellar-emission shooting phase on configuration C2
(BASELINE.json configs[1]: panchromatic Sersic bulge + exponential disk, 50 wavelengths, InterstellarDustMix,
absorption stored, SED + frame instruments, 1e8 packets) + the traversal (batched DustGrid::path) roofline.

    python bench.py --gpus N --steps K --warmup W              # this repository's engine (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's own CPU code (oracle/_ref)

One step = one complete stellar emission phase: every rank shoots `packages` packets per wavelength through
its replica of the grid (weak scaling), then the absorption table and the detector arrays are summed over the
ranks with NCCL.  Timed on the device with CUDA events on the engine's stream, max over ranks.
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
def reference_run(params, packages, threads, steps, warmup, dustsamples=10):
    """times the reference's own runstellaremission() (oracle/_ref) on `packages` packets per wavelength"""
    from oracle import skirtref as sr, refspec
    if not sr.available():
        raise RuntimeError("oracle/_ref/libskirtref.so is missing (build it where /root/reference exists: make -C oracle ref)")
    spec, L, mixes = refspec.reference_spec(params, threads=threads, dustsamples=dustsamples, packages=packages)
    S = sr.RefSim(spec, luminosities=L, mixes=mixes).setup()
    npp = S.packages_per_lambda(); nl = S.Nlambda
    times = []
    for i in range(warmup + steps):
        S.reset(4357 + i)
        sec = S.run_stellar()
        if i >= warmup:
            times.append(sec)
    total = float(np.sum(times))
    return dict(value=npp * nl * len(times) / total if total > 0 else 0.0, seconds_per_step=total / max(len(times), 1),
                packets_per_step=npp * nl, threads=threads)


def main_reference(args):
    from skirt_b200 import configs
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    params = configs.c2_params(n=args.grid, nlambda=args.nlambda, packages=args.packages)
    threads = os.cpu_count() or 1
    ppl = args.ref_packages
    try:
        r = reference_run(params, ppl, threads, args.steps, args.warmup)
    except Exception as ex:  # the oracle always exists in a built tree; report why it does not here
        print(json.dumps({"impl": "reference", "unavailable": str(ex).splitlines()[0][:200]}))
        return 0
    sample = (f"{r['packets_per_step']:.3g} packets per step ({ppl:g} per wavelength x {args.nlambda} wavelengths) of the "
              f"{args.packages * args.nlambda:.3g}-packet workload, reference runstellaremission() on {threads} threads")
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def workload_config(args, n):
    return {"workload": "C2: PanMonteCarloSimulation stellar emission phase, Sersic bulge + ExpDisk stars, ExpDisk dust tau_V=1, "
                        f"CartesianDustGrid {args.grid}^3, {args.nlambda}-point log wavelength grid 0.1-1000 micron, InterstellarDustMix, "
                        "absorption stored, FrameInstrument 800x200 + SEDInstrument at i=88deg",
            "packets_per_wavelength_per_gpu": args.packages, "wavelengths": args.nlambda,
            "packets_per_step": args.packages * args.nlambda * n, "parallelism": f"packets sharded over {n} GPU(s), NCCL allreduce per phase",
            "l2": "256 MiB memset between steps (inside the timed region); working set (Labs 400 MB + frames 64 MB) exceeds L2"}


# ---------------------------------------------------------------------------------------------------------------
def traversal_leg(engine, torch, ext, ncomp, nrays, reps=5, warm=3):
    """batched DustGrid::path()+fillOpticalDepth() (skg_path_count / skg_path_fill) on SURVEY.md 8d's synthetic rays:
    r uniform in 1.2 x the bounding box, k isotropic; device-resident inputs and outputs."""
    from skirt_b200 import configs
    g = torch.Generator(device="cuda"); g.manual_seed(0x5eed0001)
    box = torch.tensor(configs.C1_BOX, dtype=torch.float64, device="cuda")
    c = 0.5 * (box[0::2] + box[1::2]); w = box[1::2] - box[0::2]
    r = (c + (torch.rand((nrays, 3), generator=g, dtype=torch.float64, device="cuda") - 0.5) * w * 1.2).contiguous()
    k = torch.randn((nrays, 3), generator=g, dtype=torch.float64, device="cuda")
    k = (k / k.norm(dim=1, keepdim=True)).contiguous()
    ell = torch.zeros(1, dtype=torch.int32, device="cuda")
    off = torch.zeros(nrays + 1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    total = engine.path_count_device(nrays, r.data_ptr(), k.data_ptr(), off.data_ptr())
    seg = torch.empty(total * 5, dtype=torch.float64, device="cuda")       # 40-byte DustGridPath::Segment records
    torch.cuda.synchronize()

    def fill():
        engine.path_fill_device(nrays, r.data_ptr(), k.data_ptr(), ell.data_ptr(), 0, off.data_ptr(), seg.data_ptr())
    for _ in range(warm):
        fill()
    times = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(ext); fill(); e1.record(ext); e1.synchronize()
        times.append(e0.elapsed_time(e1))
    # count pass alone (geometry only, no output): the walker's own speed
    ctimes = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(ext); engine.path_count_device(nrays, r.data_ptr(), k.data_ptr(), off.data_ptr()); e1.record(ext); e1.synchronize()
        ctimes.append(e0.elapsed_time(e1))
    # through the API in ONE traversal per ray (skg_path_batch: closed-form slab capacities + scan + the same record kernel),
    # next to the two-pass sequence count + scan + fill that the CSR interface needs
    starts = torch.zeros(nrays + 1, dtype=torch.int64, device="cuda"); lens = torch.zeros(nrays, dtype=torch.int32, device="cuda")
    need = engine.path_batch_device(nrays, r.data_ptr(), k.data_ptr(), ell.data_ptr(), 0, starts.data_ptr(), lens.data_ptr(), None, 0)
    del seg
    slab = torch.empty(need * 5, dtype=torch.float64, device="cuda")
    otimes = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        engine.path_batch_device(nrays, r.data_ptr(), k.data_ptr(), ell.data_ptr(), 0, starts.data_ptr(), lens.data_ptr(), slab.data_ptr(), need)
        e1.record(ext); e1.synchronize()
        otimes.append(e0.elapsed_time(e1))
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
    ms = float(np.mean(times))
    nbytes = 60.0 * nrays + total * (36.0 + 8.0 * ncomp)
    ms_one = float(np.mean(otimes)); ms_two = ms + float(np.mean(ctimes))
    return dict(rays=nrays, packet_steps=int(total), ms=ms, ms_count_pass=float(np.mean(ctimes)), bytes=nbytes,
                through_api={"one_pass_ms": ms_one, "one_pass_gbs": nbytes / (ms_one * 1e-3) / 1e9, "slab_records": int(need),
                             "two_pass_ms": ms_two, "two_pass_gbs": nbytes / (ms_two * 1e-3) / 1e9,
                             "what": "skg_path_batch (capacity kernel + scan + record kernel, one traversal per ray) vs skg_path_count + skg_path_fill; device-resident rays and records, same algorithmic bytes"},
                steps_per_s=total / (ms * 1e-3), gbs=nbytes / (ms * 1e-3) / 1e9, written_gbs=40.0 * total / (ms * 1e-3) / 1e9,
                write_only_memset_gbs=write_only_gbs)


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
    n_gpus = world

    params = configs.c2_params(n=args.grid, nlambda=args.nlambda, packages=args.packages)
    t0 = time.time()
    sim = configs.build(params, device=local, rank=rank, nranks=world)
    sim.packages = args.packages * world                 # weak scaling: every rank shoots args.packages per wavelength
    sim.setup()
    e = sim.engine
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

    def step():
        e.reset_results()
        st = sim.runstellaremission()
        with torch.cuda.stream(ext):
            flush.zero_()
        return st

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    launches0 = e.launch_count
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    ev0.record(ext)
    stats = []
    for _ in range(args.steps):
        stats.append(step())
    ev1.record(ext)
    barrier()
    wall = time.perf_counter() - wall0
    ms = ev0.elapsed_time(ev1)
    launches = e.launch_count - launches0
    clocks = sampler.stop()
    kernel_ms = float(np.mean([s["kernel_ms"] for s in stats]))

    # ---- end-to-end through the public API with host buffers: upload every table, shoot, read every result back
    tabs = sim.ds.grid.tables(); med = sim.ds.medium(); comps = [c.geometry.sampler() for c in sim.ss.comps]
    Lum = sim.ss.luminosities(); instr = [i.d for i in sim.isys.instruments]
    h2d = sum(np.asarray(v).nbytes for v in (tabs["xv"], tabs["yv"], tabs["zv"], med["rho"], med["kext"], med["ksca"], med["g"], Lum))
    e2e_steps = max(1, min(args.steps, 3))
    # the job's results are read back once, by the root process, like the reference does (Instrument::sumResults reduces to
    # the root, which alone writes the output; PeerToPeerCommunicator.cpp:36-50): the other ranks upload their tables, shoot
    # their share and take part in the all-reduce
    root = rank == 0
    if root:
        sim.results(pinned=True)        # allocates the page-locked result buffers once (set-up, outside the timed steps)
    d2h = 0
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        e.set_grid(tabs); e.medium(med["rho"], med["kext"], med["ksca"], med["g"])
        e.sources(comps, Lum, sim.ss.emissionBias); e.instruments(instr)
        sim.runstellaremission()
        if root:
            res = sim.results(pinned=True)
            d2h = sum(v.nbytes for v in res.values())
        else:
            torch.cuda.synchronize()
    barrier()
    e2e_s = (time.perf_counter() - w0) / e2e_steps

    # ---- reduce over ranks
    def maxr(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())
    ms = maxr(ms); e2e_s = maxr(e2e_s); wall = maxr(wall)
    packets_per_step = args.packages * args.nlambda * world
    value = packets_per_step * args.steps / (ms * 1e-3)
    e2e_value = packets_per_step / e2e_s

    # ---- roofline of the dominant kernel, absorbStage (scatter + walk + absorb + terminate/sample): algorithmic bytes
    #      per phase = absorbing segments x (8*Ncomp rho gather + 16 Labs read-modify-write) + absorb paths x 128 (packet
    #      slot read + write), over the device time of all absorbStage launches of the phase (CUDA events in the engine)
    peak, peak_src = measured_peak()
    ncomp = med["rho"].shape[1] if med["rho"].ndim > 1 else 1
    st = stats[-1]
    absorb_ms = float(np.mean([s_["absorb_ms"] for s_ in stats]))
    absorb_paths = st["packets"] + st["scatterings"]
    alg_bytes = (8.0 * ncomp + 16.0) * st["absorbSegments"] + 128.0 * absorb_paths
    achieved = alg_bytes / (absorb_ms * 1e-3) / 1e9
    # DRAM traffic of the same kernel over one phase of this workload, from the committed ncu pass (profiles/)
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "r01_v6_absorbStage_dram_traffic.json")
    if os.path.exists(tfile) and args.packages == 2e6 and args.nlambda == 50 and args.grid == 100:
        t_ = json.load(open(tfile)); traffic = t_["dram_bytes_read"] + t_["dram_bytes_write"]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "what": "skg_grid_cartesian+skg_medium+skg_sources+skg_instruments from host arrays (every rank), skg_run_stellar, "
                                                 "skg_fetch_frame/sed/labs into page-locked host arrays (on the root rank, after the all-reduce)"},
            "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": wall, "setup_s": setup_s,
            "roofline": {"bound": "hbm", "kernel": "absorbStage<GRID_CART> (dominant kernel of the phase: scatter + traverse + absorb + terminate/sample)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_note": "DRAM bytes of all absorbStage launches of one phase (ncu, profiles/r01_v6_absorbStage_dram_traffic.json): far BELOW the algorithmic bytes because the density table and the wavelength-major Labs slices in flight stay L2-resident",
                         "peak_source": peak_src, "bytes_per_step": alg_bytes, "kernel_ms_per_step": absorb_ms,
                         "launches_per_step": int(st["iterations"]), "share_of_step": absorb_ms / kernel_ms,
                         "absorbing_packet_steps_per_s": st["absorbSegments"] / (absorb_ms * 1e-3),
                         "note": "fp64 issue/latency bound by design (3 IEEE divisions + expm1 + atomicAdd per 24 B); the HBM-bound kernel is the path-record kernel below"},
            "stage_ms_per_step": {k: float(np.mean([s_[k] for s_ in stats])) for k in ("launch_ms", "peel_ms", "absorb_ms", "propagate_ms", "kernel_ms")},
            "packet_steps_per_s": st["pathSegments"] / (kernel_ms * 1e-3),
            "per_step_stats": {k: int(v) for k, v in st.items() if not k.endswith("_ms")}}

    if rank == 0 and not args.skip_traversal:
        tr = traversal_leg(e, torch, ext, ncomp, args.rays)
        tr["frac"] = tr["gbs"] / peak; tr["peak"] = peak
        tr["through_api_frac"] = tr["through_api"]["one_pass_gbs"] / peak; tr["through_api_two_pass_frac"] = tr["through_api"]["two_pass_gbs"] / peak
        tr["kernel"] = "pathFillKernel<GRID_CART> (batched DustGrid::path + fillOpticalDepth, CSR path records)"
        # ncu --set full of this kernel on 1 Mi rays (profiles/r01_v6_path_kernels_ncu.txt): 1.995 GB written + 0.118 GB read
        # for 50 281 330 packet-steps = 42.0 B per step, against 44 B algorithmic (40 B of it written)
        tr["traffic"] = 42.0 * tr["packet_steps"]
        line["traversal_roofline"] = tr
    if rank == 0 and world == 1 and not args.skip_cpu:
        threads = os.cpu_count() or 1
        try:
            r = reference_run(params, args.ref_packages, threads, steps=1, warmup=0)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "reference",
                                    "sample": f"{r['packets_per_step']:.3g} packets ({args.ref_packages:g} per wavelength x {args.nlambda}) of the same "
                                              f"C2 workload, reference runstellaremission() from oracle/_ref, {r['seconds_per_step']:.1f} s"}
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
    ap.add_argument("--packages", type=float, default=2e6, help="packets per wavelength per GPU (C2: 2e6 x 50 = 1e8)")
    ap.add_argument("--nlambda", type=int, default=50)
    ap.add_argument("--grid", type=int, default=100)
    ap.add_argument("--rays", type=int, default=1 << 24, help="rays of the traversal-roofline leg (SURVEY.md 8d: 2^24; the records take 40 B x ~48 crossings per ray = 32 GB)")
    ap.add_argument("--ref-packages", type=float, default=None, help="packets per wavelength of the bounded CPU sample")
    ap.add_argument("--skip-traversal", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.ref_packages is None:
        # about 10-30 s of CPU work per step: ~3e4 packets/s/core measured on this class of host
        cores = os.cpu_count() or 1
        args.ref_packages = float(min(args.packages, max(2e3, round(15.0 * 3.0e4 * cores / args.nlambda, -3))))
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
