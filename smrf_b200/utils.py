"""Host-facing helpers with the reference's names (smrf/utils/utils.py)."""
import numpy as np

from . import _lib


def set_min_max(data, min_val, max_val, device=None):
    """NaN-preserving clamp, in place, same contract as smrf/utils/utils.py:137-161.

    The distribute calls of this package already fuse the clamp (pass min/max there); this
    stand-alone form exists so `utils.set_min_max(grid, min, max)` call sites keep working.
    It runs on the GPU (smrfb_set_min_max); there is no NumPy fallback.
    """
    if max_val is None:
        max_val = np.inf
    if min_val is None:
        min_val = -np.inf
    if not (isinstance(data, np.ndarray) and data.dtype == np.float64 and data.flags["C_CONTIGUOUS"]):
        raise TypeError("set_min_max needs a C-contiguous float64 ndarray (it is modified in place)")
    dev = _lib.default_device() if device is None else device
    _lib.check(_lib.lib().smrfb_set_min_max(_lib.dptr(data), data.size, float(min_val), float(max_val), dev))
    return data


# ------------------------------------------------------------------------------------------------
# WindNinja-style precomputed interpolation (smrf/utils/utils.py:377-449, wind_ninja.py:171, 212-235)
# ------------------------------------------------------------------------------------------------
def interp_weights(xy, uv, d=2):
    """Vertices and barycentric weights of the LINEAR interpolation from the points ``xy`` [n, 2] to ``uv`` [m, 2]
    (utils.py:377-406): one-off host geometry (scipy Delaunay + find_simplex), exactly as the reference builds it.
    Returns (vertices [m, 3] int, weights [m, 3] float64)."""
    from scipy.spatial import Delaunay
    tri = Delaunay(xy)
    simplex = tri.find_simplex(uv)
    vertices = np.take(tri.simplices, simplex, axis=0)
    temp = np.take(tri.transform, simplex, axis=0)
    delta = uv - temp[:, d]
    bary = np.einsum('njk,nk->nj', temp[:, :d, :], delta)
    return vertices, np.hstack((bary, 1 - bary.sum(axis=1, keepdims=True)))


_interp_cache = []          # [(key, vtx, wts, handle)]: the last few (vertices, weights) pairs live on the device


def _interp_handle(vtx, wts, npts, device, fill_negative=True):
    key = (vtx.__array_interface__["data"][0], wts.__array_interface__["data"][0], vtx.shape, int(npts), int(device), bool(fill_negative))
    for k, v, w, h in _interp_cache:
        if k == key and v is vtx and w is wts:
            return h
    v32 = np.ascontiguousarray(vtx, dtype=np.int32)
    w64 = _lib.as_f64(wts)
    h = _lib.handle_t()
    _lib.check(_lib.lib().smrfb_interp_create(v32.shape[0], v32.ctypes.data_as(_lib.c_i32_p), _lib.dptr(w64.reshape(-1)),
                                              int(npts), int(bool(fill_negative)), int(device), h))
    _interp_cache.append((key, vtx, wts, h))
    while len(_interp_cache) > 4:
        _, _, _, old = _interp_cache.pop(0)
        _lib.lib().smrfb_destroy(old)
    return h


def grid_interpolate(values, vtx, wts, shp, fill_value=np.nan, device=None, _fill_negative=True):
    """utils.py:409-427 on the device: ``einsum('nj,nj->n', take(values, vtx), wts)`` with ``fill_value`` where a weight is
    negative, reshaped to ``shp``.  ``values`` may be [npts] (one field, as the reference) or [T, npts] (a block: returns
    [T, shp[0], shp[1]]).  The (vertices, weights) pair is uploaded once and kept for the following calls."""
    vals = _lib.as_f64(values)
    one = vals.ndim == 1
    if one:
        vals = vals[None, :]
    T, npts = vals.shape
    dev = _lib.default_device() if device is None else device
    h = _interp_handle(vtx, wts, npts, dev, _fill_negative)
    out = _lib.pinned_empty((T, int(shp[0]), int(shp[1])), np.float64)
    if out[0].size != vtx.shape[0]:
        raise ValueError("shape %r does not match %d interpolation targets" % (tuple(shp), vtx.shape[0]))
    _lib.check(_lib.lib().smrfb_interp_apply(h, _lib.dptr(vals), T, float(fill_value), _lib.vptr(out)))
    return out[0] if one else out


def grid_interpolate_deconstructed(tri, values, grid_points, method='linear', device=None):
    """utils.py:430-449 for ``method='linear'`` (LinearNDInterpolator over a cached triangulation): simplex search and
    barycentric weights on the host once per (tri, grid_points) pair, the interpolation on the device.  NaN outside the
    hull, like LinearNDInterpolator.  ``method='cubic'`` lives in ``smrf_b200.spatial.GRID`` (Clough-Tocher on the device)."""
    if method != 'linear':
        raise NotImplementedError("grid_interpolate_deconstructed: only 'linear' here; use smrf_b200.spatial.GRID for 'cubic'")
    gx, gy = np.asarray(grid_points[0], dtype=np.float64), np.asarray(grid_points[1], dtype=np.float64)
    key = (id(tri), gx.__array_interface__["data"][0], gy.__array_interface__["data"][0], gx.shape)
    cached = getattr(tri, "_smrfb_plan", None)
    if cached is None or cached[0] != key:
        uv = np.column_stack([gx.ravel(), gy.ravel()])
        simplex = tri.find_simplex(uv)
        vertices = np.take(tri.simplices, simplex, axis=0)
        temp = np.take(tri.transform, simplex, axis=0)
        bary = np.einsum('njk,nk->nj', temp[:, :2, :], uv - temp[:, 2])
        wts = np.hstack((bary, 1 - bary.sum(axis=1, keepdims=True)))
        wts[simplex < 0] = np.nan                              # outside the hull: NaN, whatever the values (LinearNDInterpolator)
        vertices = np.where(simplex[:, None] < 0, 0, vertices)
        cached = (key, np.ascontiguousarray(vertices, dtype=np.int32), np.ascontiguousarray(wts))
        try:
            tri._smrfb_plan = cached
        except AttributeError:
            pass
    _, vtx, wts = cached
    shp = gx.shape if gx.ndim == 2 else (1, gx.size)
    out = grid_interpolate(values, vtx, wts, shp, fill_value=np.nan, device=device, _fill_negative=False)
    return out if gx.ndim == 2 else out.reshape(out.shape[:-2] + (gx.size,))
