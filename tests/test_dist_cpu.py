"""world_size-2 checks of the multi-process host logic on CPU (gloo): packet sharding, disjoint Philox stream blocks,
the hand-over of the communicator id, and that a sum over ranks of per-rank partial results equals the single-rank
result (what skg_allreduce_results does on the device with NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common
from skirt_b200.parallel import shard_packets, share_unique_id


def test_shard_packets_covers_the_budget():
    for packages, n in ((1e6, 1), (1e6, 8), (1000003, 4), (7, 8)):
        parts = [shard_packets(packages, r, n) for r in range(n)]
        assert len({p[2] for p in parts}) == 1 and parts[0][2] >= packages
        blocks = sorted((p[1], p[1] + p[0]) for p in parts)
        assert blocks[0][0] == 0 and all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
    with pytest.raises(ValueError):
        shard_packets(10, 3, 2)


class FakeEngine:
    """records what the host logic asks of the engine"""
    def __init__(self):
        self.calls = []

    def comm_unique_id(self):
        return np.arange(128, dtype=np.uint8)

    def comm_init(self, rank, nranks, uid):
        self.calls.append((rank, nranks, bytes(uid)))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        e = FakeEngine()
        uid = share_unique_id(e, dist)
        npr, offset, total = shard_packets(1001, rank, world)
        # every rank contributes the "packets" of its stream block; the reduction is a plain sum (ncclAllReduce(sum))
        part = torch.zeros(total, dtype=torch.float64); part[offset:offset + npr] = 1.0
        dist.all_reduce(part)
        out.put((rank, e.calls, uid.tolist(), npr, offset, total, part.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_two_ranks_share_the_id_and_partition_the_streams():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = [r[2] for r in res]
    assert ids[0] == ids[1] == list(range(128))
    for rank, calls, *_ in res:
        assert calls == [(rank, 2, bytes(range(128)))]
    assert res[0][3] == res[1][3] == 501 and res[0][5] == 1002
    assert (res[0][4], res[1][4]) == (0, 501)
    assert np.array_equal(res[0][6], np.ones(1002)) and np.array_equal(res[1][6], res[0][6])


# ---- the reduction choreography of the panchromatic flow (PanDustSystem.cpp:363-404, Instrument.cpp:57-65) -------------
class ModelEngine:
    """Stand-in for skirt_b200.binding.Engine in the host mirror: accumulators are numpy arrays, a "phase" adds a
    deterministic contribution per packet stream index, skg_allreduce is dist.all_reduce with the engine's rule that
    every accumulator is summed at most once between resets (skirtgpu.h).  What it checks is the HOST logic of
    skirt_b200/simulation.py: which accumulator is summed when, and that every rank takes the same decisions."""
    Ncells, Nlambda = 6, 3

    def __init__(self, dist_, world):
        self.dist, self.world = dist_, world
        self.acc = {k: np.zeros(self.Ncells * self.Nlambda) for k in ("labs", "labsdust")}
        self.acc["sed"] = np.zeros(self.Nlambda)
        self.state = {k: "zero" for k in ("labs", "labsdust", "instr")}
        self.log = []

    # set-up calls of MonteCarloSimulation.setup()
    def set_grid(self, t): pass
    def medium(self, *a): pass
    def sources(self, *a): pass
    def instruments(self, *a): pass
    def dust_library(self, *a): pass

    def _touch(self, key):
        self.state[key] = "mixed" if (self.world > 1 and self.state[key] in ("global", "mixed")) else "local"

    def _shoot(self, key, npr, offset, total, seed, weight):
        idx = np.arange(offset, offset + int(npr), dtype=np.float64)
        n = len(self.acc[key])
        for j in range(n):          # a smooth function of (stream index, bin): order-independent up to rounding
            self.acc[key][j] += weight * np.sum(np.cos(0.001 * idx * (j + 1) + seed) ** 2) / total

    def run_stellar(self, packages, total_packages=None, store_absorption=False, seed=0, stream_offset=0, **kw):
        self.log.append("stellar")
        if store_absorption:
            self._shoot("labs", packages, stream_offset, total_packages, seed, 1.0); self._touch("labs")
        self._shoot("sed", packages, stream_offset, total_packages, seed, 2.0); self._touch("instr")
        return {}

    def dust_cell_luminosities(self):
        # the spectra are made from the absorption tables as they are NOW: they must hold the sums over all ranks
        assert self.state["labs"] in ("global", "zero") or self.world == 1, "spectra made from a rank-local stellar table"
        assert self.state["labsdust"] in ("global", "zero") or self.world == 1, "spectra made from a rank-local dust table"
        self.log.append("spectra")
        return float(self.acc["labs"].sum() + self.acc["labsdust"].sum())

    def reset_labs_dust(self):
        self.acc["labsdust"][:] = 0; self.state["labsdust"] = "zero"; self.log.append("reboot")

    def run_dust_device(self, phase, d_L, packages, total_packages=None, seed=0, stream_offset=0, **kw):
        self.log.append("selfabs" if phase == 1 else "emission")
        if phase == 1:
            self._shoot("labsdust", packages, stream_offset, total_packages, seed, 0.3 * d_L); self._touch("labsdust")
        else:
            self._shoot("sed", packages, stream_offset, total_packages, seed, 0.1 * d_L); self._touch("instr")
        return {}

    def allreduce(self, which=7):
        ms = 0.0
        for bit, key, arrs in ((1, "labs", ["labs"]), (2, "labsdust", ["labsdust"]), (4, "instr", ["sed"])):
            if not which & bit or self.world == 1:
                continue
            assert self.state[key] != "mixed", f"{key} summed twice"
            if self.state[key] == "local":
                for a in arrs:
                    t = torch.from_numpy(self.acc[a]); self.dist.all_reduce(t)
                self.state[key] = "global"; self.log.append("sum:" + key); ms = 1.0
        return ms

    def labs_dust_total(self):
        t = torch.tensor([self.acc["labsdust"].sum()], dtype=torch.float64)
        if self.world > 1:
            if self.state["labsdust"] == "global":
                self.dist.broadcast(t, 0)
            else:
                self.dist.all_reduce(t)
        return float(t.item())

    def fetch_sed(self, i, out=None): return self.acc["sed"].copy()
    def fetch_labs(self, out=None): return self.acc["labs"].reshape(self.Ncells, self.Nlambda).copy()


def _pan_host_flow(dist_, rank, world, cycles):
    from skirt_b200 import simulation as sim
    lg = sim.LogWavelengthGrid(1e-7, 1e-3, 3)
    eng = ModelEngine(dist_, world)
    grid = sim.TreeTablesDustGrid(dict(kind="model"))
    ds = sim.DustSystem(grid, [sim.DustComp(None, sim.TableDustMix([1.] * 3, [1.] * 3, [0.] * 3), 1.0, 5e-7)], lg, rho=np.ones((6, 1)))
    ss = sim.StellarSystem([])
    ins = sim.InstrumentSystem([sim.SEDInstrument("s", 1.0, 0.5)])
    m = sim.MonteCarloSimulation(lg, ss, ds, ins, packages=6e4, storeAbsorption=True, rank=rank, nranks=world, engine=eng)
    m._setup = True; m._devlib = True
    m.runstellaremission()
    hist = m.rundustselfabsorption(None, cycles=cycles)
    m.rundustemission(None)
    res = m.results()
    res2 = m.results()          # reading twice must not sum twice
    assert np.array_equal(res["s_sed"], res2["s_sed"])
    return hist, res, eng.log


def _pan_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for cycles in (1, 0):
            hist, res, log = _pan_host_flow(dist, rank, world, cycles)
            out.put((rank, cycles, hist, res["s_sed"], res["Labs"], log))
    finally:
        dist.destroy_process_group()


def test_two_rank_pan_flow_equals_the_single_rank_flow():
    """stellar -> self-absorption cycles -> emission through the Python host mirror on two gloo ranks: the stellar table is
    summed once, the dust table once per cycle, the detector arrays once; both ranks see the same Labsdusttot history and
    the same cycle count, and the sums equal the one-rank run of the same stream indices"""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pan_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(4)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for cycles in (1, 0):
        a, b = sorted((g for g in got if g[1] == cycles), key=lambda t: t[0])
        assert a[2] == b[2], "ranks disagree on the self-absorption history"
        assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])
        hist1, res1, log1 = _pan_host_flow(None, 0, 1, cycles)
        assert len(a[2]) == len(hist1) and (cycles == 0 or len(hist1) == 3)
        np.testing.assert_allclose([h[2] for h in a[2]], [h[2] for h in hist1], rtol=1e-12)
        np.testing.assert_allclose(a[3], res1["s_sed"], rtol=1e-12)
        np.testing.assert_allclose(a[4], res1["Labs"], rtol=1e-12)
        log = a[5]
        assert log.count("sum:labs") == 1 and log.count("sum:instr") == 1
        assert log.count("sum:labsdust") == log.count("selfabs") == len(hist1)
        # order within a cycle: spectra (from summed tables) -> reboot -> shoot -> sum of the dust table
        i = log.index("selfabs")
        assert log[i - 2:i + 2] == ["spectra", "reboot", "selfabs", "sum:labsdust"]
        assert log.index("sum:labs") < log.index("spectra") and log[-1] == "sum:instr" and log[-2] == "emission"
