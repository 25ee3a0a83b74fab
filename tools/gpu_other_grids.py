"""Throughput of the traversal and shooting kernels on the hierarchical / unstructured grids (tables built by the
reference's own grid classes through oracle/_ref).  Prints one JSON line per grid."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
import torch
import skirt_b200 as sk
from oracle import skirtref as sr

def bench(name, S, packages=2e6, nrays=1 << 21):
    t0 = time.time(); S.setup(); tset = time.time() - t0
    tables, medium, L = S.grid_tables(), S.medium(), S.luminosities()
    e = sk.Engine(0); e.set_grid(tables); e.medium(medium["rho"], medium["kext"], medium["ksca"], medium["g"])
    e.sources([dict(geometry=1, p=[4000 * common.PC, 350 * common.PC, 0, 0, 0])], L, 0.5)
    ins = [dict(kind=1, distance=1e7 * common.PC, inclination=float(np.radians(i)), Nxp=400, fovxp=50000 * common.PC, Nyp=400, fovyp=50000 * common.PC)
           for i in (0, 30, 60, 80, 88, 90)]
    e.instruments(ins)
    ext = torch.cuda.ExternalStream(e.stream)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    box = torch.tensor(common.C1_BOX, dtype=torch.float64, device="cuda"); c = 0.5 * (box[0::2] + box[1::2]); w = box[1::2] - box[0::2]
    r = (c + (torch.rand((nrays, 3), generator=g, dtype=torch.float64, device="cuda") - 0.5) * w * 1.2).contiguous()
    k = torch.randn((nrays, 3), generator=g, dtype=torch.float64, device="cuda"); k = (k / k.norm(dim=1, keepdim=True)).contiguous()
    ell = torch.zeros(1, dtype=torch.int32, device="cuda"); off = torch.zeros(nrays + 1, dtype=torch.int64, device="cuda")
    total = e.path_count_device(nrays, r.data_ptr(), k.data_ptr(), off.data_ptr())
    seg = torch.empty(total * 5, dtype=torch.float64, device="cuda")
    times = []
    for i in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(ext); e.path_fill_device(nrays, r.data_ptr(), k.data_ptr(), ell.data_ptr(), 0, off.data_ptr(), seg.data_ptr()); e1.record(ext); e1.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = min(times[1:])
    e.run_stellar(packages / 10, store_absorption=True, seed=1)
    st = e.run_stellar(packages, store_absorption=True, seed=2)
    print(json.dumps(dict(grid=name, cells=S.Ncells, setup_s=round(tset, 1), rays=nrays, packet_steps=int(total), fill_ms=ms,
                          steps_per_s=total / ms * 1e3, gbs=(60.0 * nrays + 44.0 * total) / ms * 1e3 / 1e9,
                          packets_per_s=st["packets"] / st["kernel_ms"] * 1e3, mc_steps_per_s=st["pathSegments"] / st["kernel_ms"] * 1e3,
                          stage_ms={k_: round(st[k_], 1) for k_ in ("launch_ms", "peel_ms", "absorb_ms", "propagate_ms")},
                          stuck=e.stuck_counts())), flush=True)
    e.close()

mk = lambda spec, **kw: sr.RefSim(spec, luminosities=[[1.0]], mixes=common.mix_v(), **kw)
thr = os.cpu_count() or 1
which = os.environ.get("SKG_GRIDS", "oct,bin,amesh,voronoi").split(",")
if "oct" in which:
    bench("octtree L7 neighbor", mk(common.spec_grid("octtree", search=1, minlevel=2, maxlevel=7, massfrac=2e-6, threads=thr)))
if "oct8" in which:
    bench("octtree L8 neighbor", mk(common.spec_grid("octtree", search=1, minlevel=2, maxlevel=8, massfrac=2e-7, threads=thr)))
if "bin" in which:
    bench("bintree L18 neighbor", mk(common.spec_grid("bintree", search=1, minlevel=6, maxlevel=18, massfrac=4e-6, threads=thr)))
if "amesh" in which:
    bench("amesh depth5", mk(common.spec_grid("amesh", threads=thr), amesh=common.make_amesh(root=(8, 8, 8), max_depth=5, frac=2e-5)))
if "voronoi" in which:
    bench("voronoi 1e5", mk(common.spec_grid("voronoi", threads=thr), particles=common.voronoi_particles(100000)))
