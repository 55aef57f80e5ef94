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


# ------------------------------------------------------------------------------------------------
# Product-level driver: a year of station rows -> distributed grids, sharded over the visible GPUs
# ------------------------------------------------------------------------------------------------
class FieldSpec:
    """One distributed variable: name, method ('idw' | 'dk' | 'grid'), the reference's per-variable config
    (detrend, detrend_slope, min, max, idw_power / grid_method / grid_local ..., as in CoreConfig.ini)."""

    def __init__(self, name, method, detrend=True, slope=0, vmin=None, vmax=None, **opts):
        self.name, self.method, self.detrend, self.slope = name, method, bool(detrend), int(slope)
        self.vmin, self.vmax, self.opts = vmin, vmax, opts


class BlockRunner:
    """Distribute [T, nsta] station tables of several variables over a DEM on THIS rank's share of the job and
    hand the grids to a sink block by block.

    partition='time' : rank r takes every world-th block of `block_steps` timesteps over the whole DEM
                       (north_star wording; weak scaling over the year);
    partition='rows' : rank r owns a band of DEM rows for every timestep (divides the per-cell state, e.g. the DK
                       kriging weights, by the number of ranks).
    Either way the rank's cells are walked in row slabs of `slab_rows` rows, one interpolator handle per slab and
    variable, so the page-locked staging stays bounded: two rotating buffers of [block_steps, slab_rows, nx]
    (float64, or float32 with out_dtype=np.float32 -- the type the reference's output files hold).  The sink is called
    as sink(field_name, (t0, t1), (r0, r1), block) with block = [t1-t0, r1-r0, nx]; the view is valid until the
    nbuf-th next call.  There is no collective on the data path: ranks share nothing but the input tables."""

    def __init__(self, fields, x, y, dem, stations, rank=0, world=1, partition="time", block_steps=64, slab_rows=1000,
                 device=None, out_dtype=None, handles=None, nbuf=2):
        import numpy as np
        from . import _lib
        self.fields = list(fields)
        self.x, self.y, self.dem = np.asarray(x, float), np.asarray(y, float), np.asarray(dem, float)
        self.ny, self.nx = self.dem.shape
        self.stations = stations            # dict: method -> (mx, my, mz) or a single (mx, my, mz)
        self.rank, self.world, self.partition = int(rank), int(world), partition
        self.block_steps, self.slab_rows = int(block_steps), int(slab_rows)
        self.device = _lib.default_device() if device is None else int(device)
        self.out_dtype = np.dtype(np.float64 if out_dtype is None else out_dtype)
        if partition not in ("time", "rows"):
            raise ValueError("partition must be 'time' or 'rows'")
        self.slabs = row_slabs(self.ny, self.world if partition == "rows" else 1, self.rank if partition == "rows" else 0,
                               rows_per_slab=self.slab_rows)
        self._handles = {}
        if handles:          # interpolators the caller already holds for the WHOLE grid (one slab): {field name: IDW | DK | GRID}
            if len(self.slabs) != 1 or self.slabs[0] != (0, self.ny):
                raise ValueError("handles= needs a single whole-grid slab (slab_rows >= ny, partition='time')")
            self._handles = {(name, self.slabs[0]): h for name, h in handles.items()}
        self._pinned = None
        self._nbuf = max(1, int(nbuf))    # rotating page-locked blocks: 2 lets the sink keep one while the next fills
        self._turn = 0

    def blocks(self, T):
        if self.partition == "time":
            return time_blocks(T, self.block_steps, self.world, self.rank)
        return [(t0, min(t0 + self.block_steps, T)) for t0 in range(0, T, self.block_steps)]

    def _sta(self, f):
        s = self.stations
        return s[f.name] if isinstance(s, dict) and f.name in s else (s[f.method] if isinstance(s, dict) else s)

    def handle(self, f, slab):
        """The interpolator of variable f over the row slab (r0, r1), created on first use."""
        import numpy as np
        from .spatial import DK, GRID, IDW
        key = (f.name, slab)
        h = self._handles.get(key)
        if h is None:
            r0, r1 = slab
            X = np.broadcast_to(self.x[None, :], (r1 - r0, self.nx))
            Y = np.broadcast_to(self.y[r0:r1, None], (r1 - r0, self.nx))
            Z = np.ascontiguousarray(self.dem[r0:r1])
            mx, my, mz = self._sta(f)
            if f.method == "idw":
                h = IDW(mx, my, X, Y, mz=mz, GridZ=Z, power=f.opts.get("idw_power", 2.0), device=self.device)
            elif f.method == "dk":
                h = DK(mx, my, mz, X, Y, Z, {"dk_ncores": 1, "detrend_slope": f.slope}, device=self.device)
            elif f.method == "grid":
                cfg = {"grid_local": bool(f.opts.get("grid_local", False)), "grid_mask": False,
                       "grid_local_n": f.opts.get("grid_local_n", 25)}
                h = GRID(cfg, mx, my, X, Y, mz=mz, GridZ=Z, metadata=f.opts.get("metadata"), device=self.device)
            else:
                raise ValueError("unknown method %r" % (f.method,))
            self._handles[key] = h
        return h

    def _buffer(self, nt, rows):
        from . import _lib
        import numpy as np
        if self._pinned is None:
            cap = self.block_steps * max(b - a for a, b in self.slabs) * self.nx
            self._pinned = [_lib.pinned_empty((cap,), self.out_dtype) for _ in range(self._nbuf)]
        buf = self._pinned[self._turn % self._nbuf]
        self._turn += 1
        return buf[: nt * rows * self.nx].reshape(nt, rows, self.nx)

    def distribute(self, f, rows, slab, out=None):
        """One variable, one block of rows [nt, nsta], one slab -> [nt, r1-r0, nx] (pinned, clamp fused)."""
        h = self.handle(f, slab)
        nt = rows.shape[0]
        if out is None:
            out = self._buffer(nt, slab[1] - slab[0])
        if f.method == "idw":
            return h.distribute_block(rows, detrend=f.detrend, flag=f.slope, vmin=f.vmin, vmax=f.vmax, out=out)
        if f.method == "dk":
            return h.distribute_block(rows, vmin=f.vmin, vmax=f.vmax, out=out)
        gm = f.opts.get("grid_method", "linear")
        if f.opts.get("grid_local") and f.detrend:
            return h.distribute_block_local(rows, flag=f.slope, vmin=f.vmin, vmax=f.vmax, out=out, grid_method=gm)
        return h.distribute_block(rows, detrend=f.detrend, flag=f.slope, vmin=f.vmin, vmax=f.vmax, out=out, grid_method=gm)

    def run(self, tables, sink=None, T=None):
        """tables: {field name: [T, nsta] array}.  Returns the number of cell-timesteps this rank distributed."""
        T = T if T is not None else min(len(v) for v in tables.values())
        done = 0
        for (t0, t1) in self.blocks(T):
            for slab in self.slabs:
                for f in self.fields:
                    blk = self.distribute(f, tables[f.name][t0:t1], slab)
                    if sink is not None:
                        sink(f.name, (t0, t1), slab, blk)
                    done += blk.size
        return done
