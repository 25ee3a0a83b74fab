"""Multi-process plumbing above the C ABI: one process per GPU, launched by torchrun.  torch.distributed is only
used to hand the NCCL unique id of the engines' communicator from rank 0 to the other ranks (the job that
PeerToPeerCommunicator's MPI bootstrap does in the reference) and for the bench's timing reductions; the data-path
collective itself is the engine's grouped ncclAllReduce (skg_allreduce_results)."""
import math

import numpy as np


def shard_packets(packages, rank, nranks):
    """Block split of the packet budget per wavelength over the processes, like IdenticalAssigner / SequentialAssigner
    split the chunks (IdenticalAssigner.cpp:37-58): every rank shoots ceil(packages/nranks) packets with its own block
    of Philox stream indices.  Returns (packets for this rank, first stream index, packets over all ranks)."""
    if nranks < 1 or not (0 <= rank < nranks):
        raise ValueError("rank out of range")
    npr = int(math.ceil(packages / nranks))
    return npr, rank * npr, npr * nranks


def share_unique_id(engine, dist, device=None):
    """rank 0 creates the NCCL unique id, everybody receives it through torch.distributed (any backend) and joins
    the engines' communicator"""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = torch.zeros(128, dtype=torch.uint8, device=device or "cpu")
    if rank == 0:
        uid.copy_(torch.from_numpy(np.asarray(engine.comm_unique_id(), dtype=np.uint8)))
    dist.broadcast(uid, 0)
    engine.comm_init(rank, world, uid.cpu().numpy())
    return uid.cpu().numpy()
