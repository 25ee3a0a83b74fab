"""times the pieces of the pipelined read-back (results_begin / results_end) on the C2 tables with few packets"""
import sys, time
sys.path.insert(0, ".")
import torch
from skirt_b200 import configs

sim = configs.build(configs.c2_params(packages=1e5))
sim.setup(); e = sim.engine
def T(label, f):
    t0 = time.perf_counter(); r = f(); t1 = time.perf_counter(); print(f"{label:40s} {1e3 * (t1 - t0):9.2f} ms", flush=True); return r
T("first shoot", sim.runstellaremission)
for slot in (0, 1):
    T(f"setup begin slot {slot}", lambda: sim.results_begin(slot))
T("setup end", sim.results_end)
tabs = sim.ds.grid.tables(); med = sim.ds.medium(); comps = [c.geometry.sampler() for c in sim.ss.comps]
Lum = sim.ss.luminosities(); instr = [i.d for i in sim.isys.instruments]
for i in range(4):
    T(f"step {i}: set_grid", lambda: e.set_grid(tabs))
    T(f"step {i}: medium", lambda: e.medium(med["rho"], med["kext"], med["ksca"], med["g"]))
    T(f"step {i}: sources", lambda: e.sources(comps, Lum, sim.ss.emissionBias))
    T(f"step {i}: instruments", lambda: e.instruments(instr))
    T(f"step {i}: shoot", sim.runstellaremission)
    T(f"step {i}: results_begin", lambda: sim.results_begin(i & 1))
T("results_end", sim.results_end)
T("torch sync", torch.cuda.synchronize)
T("blocking results(pinned)", lambda: sim.results(pinned=True))
T("blocking results(pinned) again", lambda: sim.results(pinned=True))
