"""Mirror of the reference's dispatch layer ``smrf.distribute.image_data.image_data`` for the hot path.

Same attribute and method names as the reference base class (smrf/distribute/image_data.py:8-274) for the two
methods on the path -- ``_initialize(topo, metadata)`` (:127-208) and ``_distribute(data, other_attribute=None,
zeros=None)`` (:210-263) -- so a variable class written against the reference (``air_temp.ta`` etc.) can inherit
from this one unchanged.  Differences:

  * the interpolators are the CUDA-backed ``smrf_b200.spatial`` classes;
  * ``prefetch(frame)`` (optional) distributes a whole [time x station] DataFrame block in one fused launch and
    ``_distribute`` then serves the matching timestep from that block (clamp already applied) -- the run-loop façade
    of SURVEY.md 8f N-d.  Without ``prefetch`` every ``_distribute`` call is one launch, like the reference.
"""
import logging

import numpy as np

from ..spatial import dk, grid, idw


class image_data:
    def __init__(self, variable):
        self.variable = variable
        setattr(self, variable, None)
        self.gridded = False
        self._block = None
        self._block_index = None
        self._blocks = {}
        self._logger = logging.getLogger(__name__)

    def getConfig(self, cfg):
        # image_data.py:77-108
        if "stations" in cfg.keys() and cfg["stations"] is not None and not isinstance(cfg["stations"], list):
            cfg["stations"] = [cfg["stations"]]
        self.gridded = cfg.get("distribution") == "grid"
        self.getStations(cfg)
        self.config = cfg

    def getStations(self, config):
        # image_data.py:110-125: list.sort() returns None, so the station order always comes from the metadata;
        # reproduced on purpose (SURVEY.md 8a note)
        stations = None
        if "stations" in config.keys() and config["stations"] is not None:
            stations = config["stations"].sort()
        self.stations = stations

    def _initialize(self, topo, metadata):
        # image_data.py:127-208
        self.min = -np.inf if self.config.get("min") is None else self.config["min"]
        self.max = np.inf if self.config.get("max") is None else self.config["max"]
        if self.stations is not None:
            metadata = metadata.loc[self.stations]
        else:
            self.stations = metadata.index.values
        self.metadata = metadata
        try:
            self.mx = metadata.utm_x.values
            self.my = metadata.utm_y.values
        except Exception:
            self.mx = metadata.X.values
            self.my = metadata.Y.values
        self.mz = metadata.elevation.values
        dist = self.config.get("distribution")
        if dist == "idw":
            self.idw = idw.IDW(self.mx, self.my, topo.X, topo.Y, mz=self.mz, GridZ=topo.dem, power=self.config["idw_power"])
        elif dist == "dk":
            self.dk = dk.DK(self.mx, self.my, self.mz, topo.X, topo.Y, topo.dem, self.config)
        elif dist == "grid":
            self.grid = grid.GRID(self.config, self.mx, self.my, topo.X, topo.Y, mz=self.mz, GridZ=topo.dem,
                                  mask=getattr(topo, "mask", None), metadata=metadata)
        else:
            raise Exception("Could not determine the distribution method for {}".format(self.variable))

    # ------------------------------------------------------------------ block façade
    def prefetch(self, frame, other_attribute=None, clamp=True):
        """Distribute every row of ``frame`` ([time x station] DataFrame) now; clamp to min/max fused (clamp=False: not).
        With ``other_attribute`` the block serves later ``_distribute(row, other_attribute=...)`` calls."""
        if other_attribute is not None:
            if not hasattr(self, "_blocks"):
                self._blocks = {}
            keep = (self._block, self._block_index, self.min, self.max)
            try:
                if not clamp:
                    self.min, self.max = -np.inf, np.inf
                blk = self.prefetch(frame)
                self._blocks[other_attribute] = (blk, self._block_index)
            finally:
                self._block, self._block_index, self.min, self.max = keep
            return blk
        rows = frame[self.stations]
        vals = np.ascontiguousarray(rows.values, dtype=np.float64)
        if np.isnan(vals).all(axis=1).any():
            raise Exception("{}: All data values are NaN".format(self.variable))
        dist = self.config["distribution"]
        if dist == "idw":
            blk = self.idw.distribute_block(vals, detrend=bool(self.config["detrend"]), flag=self.config.get("detrend_slope", 0),
                                            vmin=self.min, vmax=self.max)
        elif dist == "dk":
            blk = self.dk.distribute_block(vals, vmin=self.min, vmax=self.max)
        elif self.config.get("grid_local") and self.config["detrend"]:     # grid.py:96-113 dispatches on grid_local
            blk = self.grid.distribute_block_local(vals, flag=self.config.get("detrend_slope", 0), vmin=self.min, vmax=self.max,
                                                   grid_method=self.config.get("grid_method", "linear"))
        else:
            blk = self.grid.distribute_block(vals, detrend=bool(self.config["detrend"]), flag=self.config.get("detrend_slope", 0),
                                             vmin=self.min, vmax=self.max, grid_method=self.config.get("grid_method", "linear"))
        self._block = blk
        self._block_index = {k: i for i, k in enumerate(rows.index)}
        return blk

    def _distribute(self, data, other_attribute=None, zeros=None, clamp=None):
        # image_data.py:210-263.  The variable classes clamp right after this call (utils.set_min_max(v, self.min,
        # self.max), e.g. air_temp.py:84); here the same NaN-preserving clamp is fused into the kernel's store, so the
        # later set_min_max finds nothing left to do (it is idempotent) and the grid crosses PCIe once.
        # clamp=None fuses it for the variable itself and not for an ``other_attribute`` (wind's direction components,
        # wind.py:168-172, must not be clamped to the wind-speed limits); True / False force it.
        if clamp is None:
            clamp = other_attribute is None
        key = getattr(data, "name", None)
        blk = self._blocks.get(other_attribute) if hasattr(self, "_blocks") else None
        if blk is not None and key in blk[1]:
            v = blk[0][blk[1][key]]
        elif other_attribute is None and self._block is not None and key in self._block_index:
            v = self._block[self._block_index[key]]
        else:
            data = data[self.stations]
            if np.sum(data.isnull()) == data.shape[0]:
                raise Exception("{}: All data values are NaN".format(self.variable))
            dist = self.config["distribution"]
            lo, hi = (self.min, self.max) if clamp else (None, None)
            if dist == "idw":
                if zeros is not None:          # rarely used branch of idw.py:96-110 (clips negatives itself)
                    v = self.idw.detrendedIDW(data.values, self.config["detrend_slope"], zeros=zeros)
                else:
                    v = self.idw.distribute_block(np.asarray(data.values, dtype=np.float64)[None, :],
                                                  detrend=bool(self.config["detrend"]),
                                                  flag=self.config.get("detrend_slope", 0), vmin=lo, vmax=hi)[0]
            elif dist == "dk":
                v = self.dk.distribute_block(np.asarray(data.values, dtype=np.float64)[None, :], vmin=lo, vmax=hi)[0]
            elif dist == "grid":
                method = self.config["grid_method"]
                if self.config["detrend"]:
                    if self.config.get("grid_local"):
                        vals = data.loc[self.metadata.index].values if hasattr(data, "loc") else np.asarray(data)
                        v = self.grid.distribute_block_local(np.asarray(vals, dtype=np.float64)[None, :],
                                                             flag=self.config["detrend_slope"], vmin=lo, vmax=hi, exact=True,
                                                             grid_method=method)[0]
                    else:
                        v = self.grid.distribute_block(np.asarray(data.values, dtype=np.float64)[None, :], detrend=True,
                                                       flag=self.config["detrend_slope"], vmin=lo, vmax=hi,
                                                       grid_method=method)[0]
                else:
                    v = self.grid.distribute_block(np.asarray(data.values, dtype=np.float64)[None, :], detrend=False,
                                                   vmin=lo, vmax=hi, grid_method=method)[0]
            else:
                raise Exception("unknown distribution")
        if other_attribute is not None:
            setattr(self, other_attribute, v)
        else:
            setattr(self, self.variable, v)
