r"""Row sharding over the GPUs of one box: one process per GPU, launched by ``torch.distributed.run``.

The determinant rows of the CI matrix are independent units for construction and for the gather-form SpMV
(SURVEY.md section 8e), so rank ``r`` owns a contiguous row block and no collective is needed to build it.
The eigen-solver all-gathers the trial vector once per SpMV and all-reduces a few scalars; ``compute_rdms``
all-reduces the dense tensors.  Those collectives run inside ``libpyci_b200.so`` on an NCCL communicator
whose unique id is created on rank 0 and handed to the other ranks here through ``torch.distributed``
(NCCL on the GPU box, gloo in the CPU tests)."""
import os

__all__ = ["env_ranks", "row_partition", "shard_of_row", "exchange_unique_id", "init_from_env", "concat_csr"]


def env_ranks():
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when not launched by it."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def row_partition(nrow, ncol, nranks):
    """[(row_begin, row_count)] per rank: uniform blocks of ``npad = ceil(max(nrow, ncol) / nranks)`` rows,
    the rule pyci_op_build (pyci_b200/csrc/api.cu) builds with.  A selected space may then be re-partitioned to equal
    stored entries per rank (pyci_b200/csrc/rebalance.cu): ``op.stats()["row_begin"/"row_count"]`` is what a rank
    holds."""
    npad = max(1, -(-max(nrow, ncol) // nranks))
    out = []
    for r in range(nranks):
        lo = min(nrow, npad * r)
        hi = min(nrow, npad * (r + 1))
        out.append((lo, hi - lo))
    return out


def shard_of_row(row, nrow, ncol, nranks):
    npad = max(1, -(-max(nrow, ncol) // nranks))
    return min(row // npad, nranks - 1)


def exchange_unique_id(make_id, rank, world_size):
    """Broadcast the 128-byte communicator id made by ``make_id()`` on rank 0 to every rank of the default
    ``torch.distributed`` process group."""
    if world_size == 1:
        return b"\0" * 128
    import torch.distributed as dist
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = bytes(box[0])
    if len(uid) != 128:
        raise ValueError("communicator id must be 128 bytes, got %d" % len(uid))
    return uid


def init_from_env(backend="nccl"):
    """Bind this process to its GPU, join the process group, and give the library its communicator.
    Returns (rank, world_size)."""
    import pyci_b200 as pyci
    rank, world, local = env_ranks()
    pyci.set_device(local)
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend == "nccl":
                torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
        uid = exchange_unique_id(pyci.nccl_unique_id, rank, world)
        pyci.init_comm(rank, world, uid)
    return rank, world


def concat_csr(shards):
    """Concatenate per-rank exports [(indptr, indices, data)] in rank order into one CSR: indptr of shard r is
    offset by the number of entries of the shards before it."""
    import numpy as np
    indptr = [np.zeros(1, dtype=np.int64)]
    off = 0
    for ip, _, _ in shards:
        indptr.append(np.asarray(ip[1:], dtype=np.int64) + off)
        off += int(ip[-1])
    return (np.concatenate(indptr), np.concatenate([s[1] for s in shards]),
            np.concatenate([s[2] for s in shards]))
