"""Run under torchrun on N GPUs.  Two checks of the multi-process data path (one engine per GPU, NCCL):

(A) stellar phase of a C1 run: every rank shoots its share, the engine sums Labs / frames / SEDs; all ranks hold identical
    arrays, the totals agree with the reference's golden runs and with a 1-rank run of the whole budget.
(B) the complete panchromatic flow of PanMonteCarloSimulation -- stellar emission -> self-absorption cycles -> dust
    emission (PanMonteCarloSimulation.cpp:105-264) -- on N ranks against the 1-rank run of the same Philox streams:
    stellar Labs summed once, the dust table after every cycle, Labsdusttot identical on all ranks (same cycle count),
    detector arrays summed once at the end (PanDustSystem.cpp:363-404, Instrument.cpp:57-65).  Labs, LabsDust, SED and
    frame must agree with the 1-rank run to 1e-9 (the packets are the same; only the order of the fp64 additions differs).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common                                   # noqa: E402
import skirt_b200 as sk                          # noqa: E402
from skirt_b200 import configs                   # noqa: E402
from skirt_b200.binding import EngineError, REDUCE_LABS_STELLAR, REDUCE_ALL    # noqa: E402
from skirt_b200.parallel import shard_packets, share_unique_id   # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).sum() / max(np.abs(b).sum(), 1e-300))


def pan_sim(local, rank, world, packages, cycles):
    p = configs.c2_params(n=16, nlambda=25, packages=packages)
    m = configs.build(p, device=local, rank=rank, nranks=world, storeAbsorption=True).setup()
    grid = m.ds.grid
    vol = (np.diff(grid.xv)[:, None, None] * np.diff(grid.yv)[None, :, None] * np.diff(grid.zv)[None, None, :]).ravel()
    m.setup_dust_library(vol)
    return m


def pan_flow(m, cycles):
    m.engine.reset_results()
    m.runstellaremission()
    hist = m.rundustselfabsorption(None, cycles=cycles)
    m.rundustemission(None)
    res = m.results()
    res["LabsDust"] = m.engine.fetch_labs_dust()
    return hist, res


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- (A) stellar phase ------------------------------------------------------------------------------------
    tables, medium, g = common.load_golden_mc()
    cfg = common.cfg_c1(n=24, packages=2e5, storeabs=1)
    e = common.setup_engine(sk.Engine(local), cfg, tables, medium, g["L"])
    share_unique_id(e, dist, device="cuda")
    npr, offset, total = shard_packets(float(g["Npp"][0]), rank, world)
    e.reset_results()
    st = e.run_stellar(npr, total_packages=total, store_absorption=True, seed=31, stream_offset=offset)
    e.allreduce(REDUCE_ALL)
    assert e.allreduce(REDUCE_ALL) == 0.0        # idempotent: everything already holds the sum
    sed, frame, labs = e.fetch_sed(1), e.fetch_frame(0), e.fetch_labs()
    sig = torch.tensor([sed.sum(), frame.sum(), labs.sum(), float(np.abs(frame).max())], dtype=torch.float64, device="cuda")
    allsig = [torch.zeros_like(sig) for _ in range(world)]
    dist.all_gather(allsig, sig)
    ok = all(torch.equal(allsig[0], a) for a in allsig)
    # rank-local additions on top of a summed table cannot be summed again in place: the engine refuses
    e.run_stellar(100, total_packages=100 * world, store_absorption=True, seed=32, stream_offset=100 * rank)
    try:
        e.allreduce(REDUCE_LABS_STELLAR); refused = False
    except EngineError as ex:
        refused = "already" in str(ex) or "after being summed" in str(ex)
    assert refused, "a second in-place sum of the stellar table must be refused"
    if rank == 0:
        assert ok, "reduced arrays differ between ranks"
        for name, a in (("sed", sed.sum()), ("frame", frame.sum()), ("labs", labs.sum())):
            b = float(g[name + "_total_mean"][0]); sem = float(g[name + "_total_sem"][0]) * 4.0
            assert abs(a - b) <= 5 * sem + 1e-3 * abs(b), f"{name}: {a} vs reference {b} +- {sem}"
        e1 = common.setup_engine(sk.Engine(local), cfg, tables, medium, g["L"])
        e1.run_stellar(total, store_absorption=True, seed=31)
        assert abs(e1.fetch_sed(1).sum() / sed.sum() - 1) < 1e-9 and abs(e1.fetch_labs().sum() / labs.sum() - 1) < 1e-9
        e1.close()
        print(f"multi-GPU check ok on {world} ranks: {st['packets']} packets per rank, totals match the reference and the 1-rank run")
    e.close()
    dist.barrier()

    # ---- (B) stellar -> self-absorption cycles -> dust emission -----------------------------------------------------
    for cycles, label in ((1, "3 fixed cycles"), (0, "cycles until convergence")):
        m = pan_sim(local, rank, world, 6e4, cycles)
        share_unique_id(m.engine, dist, device="cuda")
        hist, res = pan_flow(m, cycles)
        # every rank holds the same history (the convergence decisions are taken on identical numbers) and the same arrays
        sig = torch.tensor([len(hist)] + [h[2] for h in hist[:6]] + [res["Labs"].sum(), res["LabsDust"].sum(), res["sed88_sed"].sum(),
                           res["frame88_frame"].sum()], dtype=torch.float64, device="cuda")
        allsig = [torch.zeros_like(sig) for _ in range(world)]
        dist.all_gather(allsig, sig)
        same = all(torch.equal(allsig[0], a) for a in allsig)
        if rank == 0:
            assert same, f"pan flow ({label}): ranks disagree on the cycle history or the summed arrays"
            m1 = pan_sim(local, 0, 1, 6e4, cycles)
            hist1, res1 = pan_flow(m1, cycles)
            assert len(hist) == len(hist1), f"pan flow ({label}): {len(hist)} cycles on {world} ranks, {len(hist1)} on one"
            if cycles:
                assert len(hist) == 3
            for (s_, c_, tot, eps), (s1, c1, tot1, eps1) in zip(hist, hist1):
                assert (s_, c_) == (s1, c1) and abs(tot / tot1 - 1) < 1e-9, f"Labsdusttot of stage {s_} cycle {c_}: {tot} vs {tot1}"
            errs = {k: rel(res[k], res1[k]) for k in ("Labs", "LabsDust", "sed88_sed", "frame88_frame")}
            assert res["LabsDust"].sum() > 0 and res["sed88_sed"][-5:].sum() > 0
            assert all(v < 1e-9 for v in errs.values()), f"pan flow ({label}) on {world} ranks differs from the 1-rank run: {errs}"
            comm = {k: (float(np.mean(v)) if isinstance(v, list) else v) for k, v in m.comm_ms.items()}
            print(f"multi-GPU pan flow ok on {world} ranks ({label}): {len(hist)} cycles, relative L1 differences {errs}, allreduce ms {comm}")
            m1.engine.close()
        m.engine.close()
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
