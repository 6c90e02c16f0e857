"""Particle sharding for the multi-GPU path.

Insertion is a sum over particles, so the particle list is cut into contiguous ranges, one per
rank; every rank accumulates private V/W volumes and a single reduce onto the root precedes the
normalisation and the inverse FFT.  This mirrors the reference's MPI program (private volumes per
worker, summed at the end: libraries/parallel/mpi_reconstruct_fourier.cpp:434-436, 484-512) without
its dispatcher rank and host staging.
"""


def shard_range(n_items, world_size, rank):
    """Contiguous [begin, end) of `rank`; sizes differ by at most one, earlier ranks get the extra."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size/rank")
    base, extra = divmod(int(n_items), int(world_size))
    begin = rank * base + min(rank, extra)
    end = begin + base + (1 if rank < extra else 0)
    return begin, end


def all_ranges(n_items, world_size):
    return [shard_range(n_items, world_size, r) for r in range(world_size)]
