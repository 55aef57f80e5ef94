"""Work partition for multi-GPU runs: one process per GPU, no collective on the data path.

Every output cell-timestep depends only on that cell's coordinates / elevation and the small
[T, nsta] station table, so the job shards two ways, both embarrassingly parallel:

  * ``time_blocks``  -- each rank distributes its own blocks of timesteps over the whole grid
                        (north_star wording; what bench.py uses, weak scaling);
  * ``row_slabs``    -- each rank owns a band of DEM rows for all timesteps (divides the per-cell
                        state -- DK kriging weights -- by the number of ranks).

Both return plain (begin, end) integer ranges; nothing here touches a GPU.
"""


def split_range(n, parts, index):
    """[begin, end) of the index-th of ``parts`` near-equal contiguous pieces of range(n)."""
    if parts <= 0 or not (0 <= index < parts):
        raise ValueError("bad partition %d/%d" % (index, parts))
    base, extra = divmod(n, parts)
    begin = index * base + min(index, extra)
    return begin, begin + base + (1 if index < extra else 0)


def time_blocks(T, block, world, rank):
    """Timestep blocks [t0, t1) of length <= ``block`` dealt round-robin to the ranks."""
    blocks = [(t0, min(t0 + block, T)) for t0 in range(0, T, block)]
    return blocks[rank::world]


def row_slabs(ny, world, rank, rows_per_slab=None):
    """Row band of this rank, optionally cut further into slabs of at most rows_per_slab rows."""
    r0, r1 = split_range(ny, world, rank)
    if rows_per_slab is None or rows_per_slab >= r1 - r0:
        return [(r0, r1)] if r1 > r0 else []
    return [(a, min(a + rows_per_slab, r1)) for a in range(r0, r1, rows_per_slab)]
