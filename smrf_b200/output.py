"""Output sink of the distributed grids (SURVEY.md 8f N-c): float32 ``[time, y, x]`` files, one per variable, fed from the
page-locked blocks the interpolators return, written by a background thread so that the run loop never waits for the
disk.  Mirrors what smrf/output/output_netcdf.py:19-188 produces -- variables ``time`` (hours since the start date), ``y``,
``x`` and the data variable in float32 (the reference's ``netcdf_output_precision`` default, CoreConfig.ini:1167-1170),
optional ``data * mask`` -- with two deliberate differences: the container is NetCDF-3 (64-bit offset, written directly by
``NetCDF3Writer`` below; this image has no HDF5, the reference writes NETCDF4 with (6, 10, 10) chunks), and the file stays
open and is appended block-wise (the reference re-opens the file for every timestep, output_netcdf.py:160-188).
Pure host I/O: nothing here touches the GPU."""
import queue
import struct
import threading

import numpy as np

NC_FLOAT, NC_CHAR = 5, 2
NC_DIMENSION, NC_VARIABLE, NC_ATTRIBUTE = 0x0A, 0x0B, 0x0C


def _pad4(b):
    return b + b"\x00" * (-len(b) % 4)


def _name(s):
    b = s.encode()
    return struct.pack(">i", len(b)) + _pad4(b)


def _att_list(atts):
    if not atts:
        return struct.pack(">ii", 0, 0)
    out = struct.pack(">ii", NC_ATTRIBUTE, len(atts))
    for k, v in atts.items():
        b = str(v).encode()
        out += _name(k) + struct.pack(">ii", NC_CHAR, len(b)) + _pad4(b)
    return out


class NetCDF3Writer:
    """One variable ``[time (record), y, x]`` float32 in a NetCDF-3 64-bit-offset file, appendable record by record and
    writable row-slab by row-slab (``write(t_index, r0, block)``)."""

    def __init__(self, path, variable, x, y, units="", time_units="hours since 1900-01-01", attrs=None):
        self.path, self.variable = path, variable
        self.ny, self.nx = len(y), len(x)
        self.nrec = 0
        dims = [("time", 0), ("y", self.ny), ("x", self.nx)]
        hdr_vars = [("time", [0], {"units": time_units, "calendar": "standard"}, 4),
                    ("y", [1], {"units": "meters"}, 4 * self.ny), ("x", [2], {"units": "meters"}, 4 * self.nx),
                    (variable, [0, 1, 2], dict({"units": units}, **(attrs or {})), 4 * self.ny * self.nx)]

        def header(begins):
            h = b"CDF\x02" + struct.pack(">i", self.nrec)
            h += struct.pack(">ii", NC_DIMENSION, len(dims)) + b"".join(_name(n) + struct.pack(">i", ln) for n, ln in dims)
            h += _att_list({"Conventions": "CF-1.6", "source": "smrf_b200"})
            h += struct.pack(">ii", NC_VARIABLE, len(hdr_vars))
            for (n, dimids, atts, vsize), beg in zip(hdr_vars, begins):
                h += _name(n) + struct.pack(">i", len(dimids)) + b"".join(struct.pack(">i", d) for d in dimids)
                h += _att_list(atts) + struct.pack(">ii", NC_FLOAT, vsize) + struct.pack(">q", beg)
            return h
        hlen = len(header([0, 0, 0, 0]))
        self.y_begin = hlen
        self.x_begin = self.y_begin + 4 * self.ny
        self.rec_begin = self.x_begin + 4 * self.nx            # records: [time float][data ny*nx float] each
        self.rec_size = 4 + 4 * self.ny * self.nx
        self._header = lambda: header([self.rec_begin, self.y_begin, self.x_begin, self.rec_begin + 4])
        self.f = open(path, "wb+")
        self.f.write(self._header())
        self.f.write(np.asarray(y, dtype=">f4").tobytes())
        self.f.write(np.asarray(x, dtype=">f4").tobytes())

    def write(self, t_index, time_value, r0, block):
        """block: [rows, nx] (any float dtype) of record ``t_index`` starting at row r0."""
        off = self.rec_begin + t_index * self.rec_size
        if t_index >= self.nrec:                               # new record(s): extend, fill the time values lazily
            self.nrec = t_index + 1
        self.f.seek(off)
        self.f.write(struct.pack(">f", float(time_value)))
        self.f.seek(off + 4 + 4 * r0 * self.nx)
        self.f.write(np.ascontiguousarray(block, dtype=">f4").tobytes())

    def close(self):
        end = self.rec_begin + self.nrec * self.rec_size
        self.f.seek(0, 2)
        if self.f.tell() < end:                                 # rows never written stay zero
            self.f.truncate(end)
        self.f.seek(0)
        self.f.write(self._header())
        self.f.close()


class BlockSink:
    """sink(field, (t0, t1), (r0, r1), block) for smrf_b200.sharding.BlockRunner / the run loop: queues the block and
    returns at once; a writer thread converts to float32, applies the mask (``mask_output``) and appends to the files.
    The caller's buffer must stay valid until the block is written: with BlockRunner(nbuf=2) ``max_pending=1``."""

    def __init__(self, out_dir, variables, x, y, hours, mask=None, mask_output=False, start_date="1900-01-01", units=None,
                 max_pending=1):
        import os
        os.makedirs(out_dir, exist_ok=True)
        self.hours = np.asarray(hours, dtype=np.float64)
        self.mask = None if (mask is None or not mask_output) else np.asarray(mask, dtype=np.float32)
        units = units or {}
        self.writers = {v: NetCDF3Writer(os.path.join(out_dir, v + ".nc"), v, x, y, units=units.get(v, ""),
                                         time_units="hours since %s" % start_date) for v in variables}
        self.q = queue.Queue(maxsize=max_pending)
        self.error = None
        self.bytes_written = 0
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def __call__(self, field, trange, rrange, block):
        if self.error:
            raise self.error
        self.q.put((field, trange, rrange, block))

    def _run(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            try:
                field, (t0, t1), (r0, r1), block = item
                w = self.writers[field]
                for k, t in enumerate(range(t0, t1)):
                    b = np.asarray(block[k], dtype=np.float32)
                    if self.mask is not None:
                        b = b * self.mask[r0:r1]                    # output_netcdf.py:182-183
                    w.write(t, self.hours[t], r0, b)
                    self.bytes_written += b.nbytes
            except Exception as e:            # surfaced on the next call / close
                self.error = e
            finally:
                self.q.task_done()

    def close(self):
        self.q.join()
        self.q.put(None)
        self.th.join()
        for w in self.writers.values():
            w.close()
        if self.error:
            raise self.error
