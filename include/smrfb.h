/*
 * smrfb.h -- C-ABI of libsmrfb.so: B200 (sm_100a) spatial distribution for SMRF.
 *
 * The drop-in boundary for the reference's per-timestep distribution path
 * (reference paths relative to USDA-ARS-NWRC/smrf):
 *
 *   smrfb_idw_*            replaces smrf/spatial/idw.py:14-144      (IDW.__init__, calculateIDW, detrendedIDW)
 *   smrfb_dk_*             replaces smrf/spatial/dk/dk.py:20-168    (DK.calculate, calculateWeights, detrend/retrend)
 *   smrfb_krige_grid       replaces smrf/spatial/dk/dk_header.h:23  (krige_grid(), reached through
 *                                   detrended_kriging.pyx:32-67 call_grid) -- same arguments, same in-place output
 *   smrfb_grid_*           replaces smrf/spatial/grid.py:23-231     (GRID.__init__, detrendedInterpolation[Mask],
 *                                   calculateInterpolation, grid_method='linear')
 *   vmin/vmax arguments    replace  smrf/utils/utils.py:137-161     (set_min_max, fused into the store)
 *   smrfb_set_min_max      replaces smrf/utils/utils.py:137-161     (stand-alone form)
 *
 * Conventions
 *   - plain C types only; every function returns an int status (0 = SMRFB_OK); the message of the
 *     last failure on the calling thread is smrfb_last_error().  Nothing here calls exit()
 *     (the reference's C does on a singular kriging matrix, krige.c:161-164).
 *   - station data are row-major double [T][nsta]; NaN marks a missing station (idw.py:88, dk.py:75).
 *   - outputs are row-major double [T][ncell], ncell = ny*nx, cell (i,j) at flat index i*nx + j.
 *   - grid_kind 0: separable mesh as np.meshgrid(x, y) gives (smrf/data/load_topo.py:65):
 *                  gx[nx], gy[ny]; cell (i,j) sits at (gx[j], gy[i]).
 *     grid_kind 1: arbitrary cell coordinates gx[ncell], gy[ncell] (GridX.ravel(), GridY.ravel()).
 *   - a multi-GPU run gives every rank its own handle over its own row slab (pass the slab's
 *     gy / dem rows); slabs are independent, there is no collective.
 *   - handles are independent: one device + one internal CUDA stream each; calls on different
 *     handles may run concurrently from different host threads, calls on one handle must be
 *     serialised by the caller (true of the reference: one thread per variable,
 *     smrf/framework/model_framework.py:615-632).
 *   - *_dev variants take DEVICE pointers and enqueue on the given cudaStream_t (passed as void*;
 *     NULL = the handle's stream) without synchronising.
 *   - slope_flag is the reference's detrend_slope: +1 / -1 constrain the sign of the elevation
 *     trend (both coefficients become 0 when violated, idw.py:124-127), 0 = unconstrained.
 *   - pv_in (nullable, [T][2] = slope, intercept) bypasses the device regression, e.g. to feed
 *     np.polyfit's coefficients; pv_out (nullable, [T][2]) returns the coefficients used.
 *   - vmin / vmax: pass -INFINITY / +INFINITY for "no bound" (image_data.py:148-154).
 */
#ifndef SMRFB_H
#define SMRFB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct smrfb_handle_s* smrfb_handle_t;

enum {
    SMRFB_OK = 0,
    SMRFB_ERR_ARG = 1,       /* bad argument */
    SMRFB_ERR_CUDA = 2,      /* CUDA runtime failure (message holds the cudaError string) */
    SMRFB_ERR_SINGULAR = 3,  /* singular kriging system (co-located stations) */
    SMRFB_ERR_LIMIT = 4,     /* size outside what the kernels support */
    SMRFB_ERR_NODATA = 5     /* a timestep has no valid station at all (image_data.py:229-231) */
};

const char* smrfb_last_error(void);
int smrfb_version(void);
int smrfb_device_count(int* count);
int smrfb_destroy(smrfb_handle_t h);
/* Block until everything enqueued on the handle's own stream has finished. */
int smrfb_synchronize(smrfb_handle_t h);
/* Number of kernels this library has launched since load (all handles); bench.py's gpu_launches. */
int64_t smrfb_launch_count(void);
/* Output element type of every later distribute call on this handle: 0 = float64 (default; the parity contract of
 * 1e-9), 1 = float32 (the type the reference's output files hold, smrf/output/output_netcdf.py:58,90-93 and
 * CoreConfig.ini:1167-1170; contract 1e-5).  The arithmetic stays fp64; the value is rounded once at the store.
 * With dtype 1 every `out` / `d_out` argument below points to float [T][ncell] although it is declared double*. */
int smrfb_set_output_dtype(smrfb_handle_t h, int dtype);
/* Page-locked host memory for outputs (D2H copies into pageable memory are staged by the driver at a fraction of the
 * PCIe rate).  The drop-in classes return NumPy arrays that live in such blocks and give them back when collected. */
int smrfb_host_alloc(size_t bytes, void** out);
int smrfb_host_free(void* p);

/* ------------------------------------------------------------------ IDW (idw.py) */
int smrfb_idw_create(int nsta, const double* mx, const double* my, const double* mz /* nullable */,
                     int ny, int nx, int grid_kind, const double* gx, const double* gy,
                     const double* dem /* nullable: required for detrend */,
                     double power, int device, smrfb_handle_t* out);
/* detrend = 0: calculateIDW (idw.py:82-94); detrend = 1: detrendedIDW (idw.py:96-110). Host buffers. */
int smrfb_idw_distribute(smrfb_handle_t h, const double* data, int T, int detrend, int slope_flag,
                         double vmin, double vmax, const double* pv_in, double* pv_out, double* out);
int smrfb_idw_distribute_dev(smrfb_handle_t h, const double* d_data, int T, int detrend, int slope_flag,
                             double vmin, double vmax, const double* d_pv_in, double* d_pv_out,
                             double* d_out, void* stream);

/* Same, restricted to the cell window [cell_begin, cell_begin + cell_count) of the handle's grid;
 * d_out is [T][cell_count].  Lets a caller walk a large grid slab by slab with long time blocks
 * (the per-tile weight generation amortises over T) while the device-resident output stays bounded. */
int smrfb_idw_distribute_window_dev(smrfb_handle_t h, const double* d_data, int T, int detrend, int slope_flag,
                                    double vmin, double vmax, const double* d_pv_in, double* d_pv_out,
                                    int64_t cell_begin, int64_t cell_count, double* d_out, void* stream);

/* Which kernel a whole-row call on this handle takes, and the plan of the near / far-field path (csrc/idw_far.cu):
 * info[0] = 1 if the near field + interpolated far field kernels apply (uniform mesh, power <= 4, > 16 stations), [1] =
 * Chebyshev nodes per direction, [2] = longest near-station list of a patch, [3] = patches, [4] = near-field radius in
 * patch half widths x 1000, [5] = 1 if the int8 tensor-core kernel serves the other calls, [6] = 1 if the register-weight
 * kernel does (<= 16 stations), [7] = mean near-station count per patch x 1000. */
int smrfb_idw_info(smrfb_handle_t h, int info[8]);
/* Per-launch device timing of the near / far-field path (CUDA events on the launch stream, for bench.py's roofline line).
 * enable = 1 starts collecting, 0 stops.  If ms_out is non-NULL the handle's stream(s) are synchronised and the events
 * collected so far are summed and dropped: ms_out[0] = node kernel ms, ms_out[1] = main kernel ms, ms_out[2] = launches. */
int smrfb_idw_far_timing(smrfb_handle_t h, int enable, double ms_out[3]);

/* ------------------------------------------------------------------ DK (dk.py, krige.c, lusolv.c) */
int smrfb_dk_create(int nsta, const double* mx, const double* my, const double* mz,
                    int ny, int nx, int grid_kind, const double* gx, const double* gy, const double* dem,
                    int slope_flag, int device, smrfb_handle_t* out);
/* DK.calculate for T rows; kriging weights are built on first sight of each NaN mask and cached. */
int smrfb_dk_distribute(smrfb_handle_t h, const double* data, int T, double vmin, double vmax,
                        const double* pv_in, double* pv_out, double* out);
int smrfb_dk_distribute_dev(smrfb_handle_t h, const double* d_data, const double* h_data /* same rows, host copy
                            used to read the NaN masks */, int T, double vmin, double vmax,
                            const double* d_pv_in, double* d_pv_out, double* d_out, void* stream);
int smrfb_dk_distribute_window_dev(smrfb_handle_t h, const double* d_data, const double* h_data, int T, double vmin,
                                   double vmax, const double* d_pv_in, double* d_pv_out, int64_t cell_begin,
                                   int64_t cell_count, double* d_out, void* stream);
/* Dense weights [ncell][nvalid] (columns = valid stations in station order) for one NaN mask,
 * exactly what DK.calculateWeights leaves in wg (dk.py:127-130).  nan_mask[nsta]: 1 = missing. */
int smrfb_dk_weights(smrfb_handle_t h, const uint8_t* nan_mask, double* weights_out);
/* Statistics of the weight builder: masks built, cells that needed the exact per-cell fallback, max active set. */
int smrfb_dk_stats(smrfb_handle_t h, int64_t* masks_built, int64_t* fallback_cells, int* max_active);

/* krige_grid() of dk_header.h:23 on the GPU: same argument order and meaning, weights[ngrid][nsta]
 * written in place. nthreads is accepted and ignored.  Returns a status instead of exit()ing. */
int smrfb_krige_grid(int nsta, int ngrid, const double* ad, const double* dgrid, const double* elevations,
                     int nthreads, double* weights);

/* ------------------------------------------------------------------ GRID linear (grid.py) */
/* The Delaunay triangulation and the per-cell simplex are built once on the host (scipy, as the
 * reference does): simplices[nsimp][3], transform[nsimp][3][2] (scipy.spatial.Delaunay.transform),
 * cell_simplex[ncell] (-1 = outside the hull -> NaN), point_mask[npts] (grid.py:80-94; NULL = all). */
int smrfb_grid_create(int npts, const double* mx, const double* my, const double* mz /* nullable */,
                      int ny, int nx, int grid_kind, const double* gx, const double* gy,
                      const double* dem /* nullable */,
                      int nsimp, const int32_t* simplices, const double* transform,
                      const int32_t* cell_simplex, const uint8_t* point_mask, int device, smrfb_handle_t* out);
/* Optional: the nearest source point of every cell (scipy.spatial.cKDTree.query, as griddata's
 * NearestNDInterpolator does it), enabling grid_method = 'nearest'.  cell_nearest[ncell] in [0, npts). */
int smrfb_grid_set_nearest(smrfb_handle_t h, const int32_t* cell_nearest);
/* detrend = 1: detrendedInterpolationMask (grid.py:179-212); 0: calculateInterpolation (grid.py:214-231).
 * method: 0 = 'linear', 1 = 'nearest' (the reference's grid_method; 'cubic' has its own entry below). */
int smrfb_grid_distribute(smrfb_handle_t h, const double* data, int T, int detrend, int slope_flag, int method,
                          double vmin, double vmax, const double* pv_in, double* pv_out, double* out);
int smrfb_grid_distribute_dev(smrfb_handle_t h, const double* d_data, int T, int detrend, int slope_flag, int method,
                              double vmin, double vmax, const double* d_pv_in, double* d_pv_out,
                              double* d_out, void* stream);

/* detrendedInterpolationLocal (grid.py:115-177) with grid_method = 'linear': fields3[T][3][npts] holds per step the
 * residual, the slope and the intercept of the per-point local fits (grid.py:125-141,164-166; computed by the
 * caller, smrf_b200/spatial/grid.py); out[t] = interp(residual) + interp(slope) * GridZ + interp(intercept), each
 * interpolation as griddata / LinearNDInterpolator evaluates it (utils.py:430-449), NaN outside the hull. */
/* grid_local arithmetic: exact = 1 (default) evaluates the three interpolations in scipy.interpolate.griddata's operation order
 * (bit-identical to grid.py:143-175 given the same per-point fits); exact = 0 folds residual + intercept into one field and uses
 * fused multiply-adds -- a few ulp away, 1.7 x faster; what a block caller with its own (closed-form) fits wants. */
int smrfb_grid_set_local_exact(smrfb_handle_t h, int exact);
int smrfb_grid_distribute_local(smrfb_handle_t h, const double* fields3, int T, double vmin, double vmax, double* out);
int smrfb_grid_distribute_local_dev(smrfb_handle_t h, const double* d_fields3, int T, double vmin, double vmax,
                                    double* d_out, void* stream);

/* grid_method = 'cubic' (the reference's default, CoreConfig.ini:183-186): scipy's CloughTocher2DInterpolator, reached
 * through griddata(method='cubic') at grid.py:204-207 / :226-229 and through utils.py:447 on the cached triangulation.
 * ct_geom[nsimp][9] = per simplex the edge vectors e12, e23, e31 (x, y each) and the three cross-edge coefficients g
 * (neighbour centroids in barycentric form; smrf_b200/spatial/grid.py builds them from scipy.spatial.Delaunay).
 * fields[T][nf][3][npts]: per step and field the point values and their gradients d/dx, d/dy (scipy's estimator,
 * run by the caller like the Delaunay triangulation).
 *   nf = 1: out = patch(field 0) [+ pv[t][0] * GridZ + pv[t][1] when pv != NULL: detrendedInterpolationMask with the
 *           residuals formed by the caller, grid.py:188-210; pv == NULL: calculateInterpolation]
 *   nf = 2: out = patch(field 0) + GridZ * patch(field 1)      (detrendedInterpolationLocal, grid.py:164-175, with
 *           field 0 = residual + intercept, field 1 = slope)
 * NaN outside the hull; clamp fused. */
int smrfb_grid_set_cubic(smrfb_handle_t h, const double* ct_geom);
int smrfb_grid_distribute_cubic(smrfb_handle_t h, const double* fields, int T, int nf, const double* pv /* nullable [T][2] */,
                                double vmin, double vmax, double* out);
int smrfb_grid_distribute_cubic_dev(smrfb_handle_t h, const double* d_fields, int T, int nf, const double* d_pv,
                                    double vmin, double vmax, double* d_out, void* stream);

/* ------------------------------------------------------------------ clamp (utils.py:137-161) */
/* In place on a host array: x<=vmin -> vmin, x>=vmax -> vmax, NaN stays NaN.  The distribute calls fuse this clamp
 * (vmin / vmax arguments); the stand-alone form is for call sites that clamp afterwards.  Arrays from smrfb_host_alloc
 * are clamped in place through their device mapping (no allocation, no staging copy); pageable arrays are staged. */
int smrfb_set_min_max(double* data, int64_t n, double vmin, double vmax, int device);

/* ------------------------------------------------------------------ consumers of a distributed grid (SURVEY.md 8f N-a), WindNinja-style weights (8a G5) */
/* Dew point (degC) from vapor pressure (Pa): smrf/envphys/core/dewpt.c:14-229 (envphys_c.cdewpt(vp, dpt, tolerance, nthreads),
 * called at smrf/distribute/vapor_pressure.py:136-141): Brent root of sati(T) = e, result truncated to float, K -> C. */
int smrfb_dew_point(const double* vp, int64_t n, double tol, double* dpt, int device);
int smrfb_dew_point_dev(const double* d_vp, int64_t n, double tol, double* d_dpt, void* stream);

/* Wet / ice bulb temperature (C) from air temperature (C), dew point (C) and elevation (m):
 * replaces envphys_c.cwbt -> iwbt (smrf/envphys/core/envphys_c.pyx:103-139, iwbt.c:84-134).  Where the reference prints and calls
 * exit(-1) -- a temperature below 0 K, no convergence within 10 Newton steps (iwbt.c:74-77, 123-126) -- the host form returns
 * SMRFB_ERR_ARG / SMRFB_ERR_NODATA after filling tw; the device form ORs 2 / 1 into *d_flags (an int the caller zeroed). */
int smrfb_wet_bulb(const double* ta, const double* td, const double* z, int64_t n, double tol, double* tw, int device);
int smrfb_wet_bulb_dev(const double* d_ta, const double* d_td, const double* d_z, int64_t n, double tol, double* d_tw, int* d_flags,
                       void* stream);
/* Precipitation phase and new-snow density from the precipitation temperature (the distributed dew point, C) and the distributed
 * precipitation (mm; the storm total for marks2017): replaces Snow.phase_and_density (smrf/envphys/snow.py:37-75, called at
 * smrf/distribute/precipitation.py:376-379, 428-431) with the models of smrf/envphys/nasde_model/: model 0 = susong1999
 * (susong_1999.py:29-116), 1 = piecewise_susong1999 (piecewise_suosong_1999.py:42-87), 2 = marks2017 (marks_2017.py:73-207).
 * pcs = fraction of the precipitation that is snow (0..1), rho_s = density of the new snow (kg/m^3).  The device form takes an int
 * of scratch the caller zeroed (marks2017 first asks whether any cell has precipitation). */
int smrfb_snow_phase(const double* tpp, const double* pp, int64_t n, int model, double* pcs, double* rho_s, int device);
int smrfb_snow_phase_dev(const double* d_tpp, const double* d_pp, int64_t n, int model, double* d_pcs, double* d_rho_s, int* d_any,
                         void* stream);
/* Hook a consumer onto the NEXT host-buffer distribute calls of a handle (any method): kind 1 = the dew point of every
 * distributed value, computed while the sub-batch is still in the handle's device ring and copied to out2 ([T][ncell], same
 * element type as the primary output) -- the vapor-pressure grid crosses PCIe once and is never re-read from host memory.
 * kind 0 removes the hook. */
int smrfb_set_consumer(smrfb_handle_t h, int kind, double tol, void* out2);
/* Wind azimuth from the two distributed direction components: az = arctan2(u, v) * 180 / pi, negative values + 360
 * (smrf/distribute/wind/wind.py:175-178). */
int smrfb_wind_direction(const double* u, const double* v, int64_t n, double* az, int device);
int smrfb_wind_direction_dev(const double* d_u, const double* d_v, int64_t n, double* d_az, void* stream);
/* utils.interp_weights / utils.grid_interpolate (smrf/utils/utils.py:377-427, used by smrf/distribute/wind/wind_ninja.py:171,
 * 212-235): vertices [n][3] (indices into the npts source values) and weights [n][3] as interp_weights returns them;
 * apply: out[t][i] = sum_j values[t][vtx[i][j]] * wts[i][j]; with fill_negative = 1 cells with a negative weight take `fill`
 * (grid_interpolate, utils.py:424); with 0 the weights alone decide (NaN weights mark cells outside the hull: the
 * LinearNDInterpolator rule of grid_interpolate_deconstructed). */
int smrfb_interp_create(int64_t n, const int32_t* vtx, const double* wts, int npts, int fill_negative, int device, smrfb_handle_t* out);
int smrfb_interp_apply(smrfb_handle_t h, const double* values, int T, double fill, double* out);
int smrfb_interp_apply_dev(smrfb_handle_t h, const double* d_values, int T, double fill, double* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SMRFB_H */
