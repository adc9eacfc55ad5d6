"""Multi-GPU plumbing: one process per GPU (torch.distributed), chains partitioned over ranks,
graph replicated, ONE collective on the path -- the all-reduce (sum) of the per-node marginal
label histogram.  Chains never communicate; maximise / grid modes only gather
the best partition.  Works with backend "nccl" on device tensors and "gloo" on CPU tensors (the
latter is what the CPU tests exercise)."""
import numpy as np


def shard_chains(n_chains, rank, world):
    """Round-robin chain -> rank map (chain c lives on rank c % world), so a heterogeneous (Ka,Kb)
    grid spreads its cheap and expensive chains evenly.  Returns the global chain ids of `rank`."""
    return np.arange(rank, n_chains, world, dtype=np.int64)


def chain_seeds(base_seed, chain_ids):
    """Seed of a chain depends only on its GLOBAL id, so a run is reproducible for any world size."""
    ids = np.asarray(chain_ids, dtype=np.uint64)
    return (np.uint64(base_seed) * np.uint64(0x9E3779B97F4A7C15) + ids).astype(np.uint64)


def init_pool_comm(pool):
    """Give `pool`'s handle its own NCCL communicator over the torch.distributed world: rank 0 draws the unique id
    (bisbm_nccl_get_unique_id), torch.distributed only carries those 128 bytes to the other ranks, every rank joins
    with bisbm_nccl_init.  After this the marginal all-reduce runs entirely behind the C ABI."""
    import torch.distributed as dist
    from . import host
    world, rank = dist.get_world_size(), dist.get_rank()
    box = [host.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    pool.nccl_init(world, rank, box[0])
    pool._has_comm = True


def allreduce_marginals(hist):
    """In-place sum of the marginal histogram over ranks.  `hist` is a ChainPool whose handle joined a communicator
    (init_pool_comm): the library's own ncclAllReduce on the device-resident histogram over NVLink; or a torch tensor
    (CPU tensor over gloo in the host-logic tests, or a wrapped device histogram)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return hist
    if hasattr(hist, "marginals_allreduce"):
        if getattr(hist, "_has_comm", False):
            hist.marginals_allreduce()
        else:   # a pool without its own communicator: torch's collective on the wrapped device histogram
            from . import host
            dist.all_reduce(host.marginals_tensor(hist), op=dist.ReduceOp.SUM)
    else:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist


def gather_best(entropy, labels):
    """All ranks learn the globally best (lowest description length) chain: returns
    (entropy, labels, owner_rank).  entropy: float array of the local chains; labels: [local][n]."""
    import torch
    import torch.distributed as dist
    entropy = np.asarray(entropy, dtype=np.float64)
    i = int(np.argmin(entropy)) if entropy.size else -1
    best = float(entropy[i]) if i >= 0 else float("inf")
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return best, np.asarray(labels[i]), 0
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.tensor([best], dtype=torch.float64, device=dev)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    vals = [float(t.item()) for t in allv]
    owner = int(np.argmin(vals))
    lab = torch.from_numpy(np.ascontiguousarray(labels[i], dtype=np.int64)).to(dev) if rank == owner else \
        torch.zeros(len(labels[0]) if len(labels) else 0, dtype=torch.int64, device=dev)
    dist.broadcast(lab, src=owner)
    return vals[owner], lab.cpu().numpy(), owner
