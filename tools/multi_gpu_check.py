"""Run under torchrun on N GPUs: every rank shoots its share of a C1 run, the engine all-reduces Labs / frames / SEDs
over NCCL, and rank 0 checks (1) all ranks hold identical reduced arrays, (2) the totals agree with the reference's
golden runs, (3) the N-rank result is statistically the same as a 1-rank run of the full budget."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common                                   # noqa: E402
import skirt_b200 as sk                          # noqa: E402
from skirt_b200.parallel import shard_packets, share_unique_id   # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tables, medium, g = common.load_golden_mc()
    cfg = common.cfg_c1(n=24, packages=2e5, storeabs=1)
    e = common.setup_engine(sk.Engine(local), cfg, tables, medium, g["L"])
    share_unique_id(e, dist, device="cuda")
    npr, offset, total = shard_packets(float(g["Npp"][0]), rank, world)
    e.reset_results()
    st = e.run_stellar(npr, total_packages=total, store_absorption=True, seed=31, stream_offset=offset)
    e.allreduce_results()
    sed, frame, labs = e.fetch_sed(1), e.fetch_frame(0), e.fetch_labs()
    # (1) identical on every rank
    sig = torch.tensor([sed.sum(), frame.sum(), labs.sum(), float(np.abs(frame).max())], dtype=torch.float64, device="cuda")
    allsig = [torch.zeros_like(sig) for _ in range(world)]
    dist.all_gather(allsig, sig)
    ok = all(torch.equal(allsig[0], a) for a in allsig)
    if rank == 0:
        assert ok, "reduced arrays differ between ranks"
        for name, a in (("sed", sed.sum()), ("frame", frame.sum()), ("labs", labs.sum())):
            b = float(g[name + "_total_mean"][0]); sem = float(g[name + "_total_sem"][0]) * 4.0
            assert abs(a - b) <= 5 * sem + 1e-3 * abs(b), f"{name}: {a} vs reference {b} +- {sem}"
        # (3) single-rank run of the whole budget with the same seed: the same packets (same Philox streams), so the
        # sums agree up to the order of the floating-point additions
        e1 = common.setup_engine(sk.Engine(local), cfg, tables, medium, g["L"])
        e1.run_stellar(total, store_absorption=True, seed=31)
        assert abs(e1.fetch_sed(1).sum() / sed.sum() - 1) < 1e-9 and abs(e1.fetch_labs().sum() / labs.sum() - 1) < 1e-9
        print(f"multi-GPU check ok on {world} ranks: {st['packets']} packets per rank, totals match the reference and the 1-rank run")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
