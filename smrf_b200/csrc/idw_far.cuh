// idw_far.cuh -- interface of the far-field / near-field IDW kernels, see idw_far.cu.
#pragma once
#include <vector>

#include "common.cuh"

namespace smrfb {

constexpr int FAR_P = 128;   // patch edge in cells (both directions)

// Per-handle plan (built once at create time on the host): the uniform mesh is cut into FAR_P x FAR_P-cell patches; per
// patch the stations closer than theta * (patch half width) to the patch box are its NEAR stations (summed directly per
// cell), all others are FAR and enter through an n x n tensor-product Chebyshev interpolant of their weight sums.
struct IdwFarPlan {
    bool enabled = false;
    int nn = 10;               // Chebyshev nodes per direction (10 for power <= 2, 12 for power <= 4)
    double theta = 6.0;        // near-field radius in patch half widths
    int npx = 0, npy = 0;      // patches per direction
    int max_near = 0;          // longest near list
    double x0 = 0, dx = 0, y0 = 0, dy = 0;   // mesh: x_j = x0 + j dx, y_i = y0 + i dy (dy < 0 for topo files)
    int* d_near_off = nullptr;   // [npx*npy + 1]
    int* d_near_idx = nullptr;   // [near_off[last]] station indices, ascending per patch
    double* d_tab = nullptr;     // [128][nn] Lagrange basis at the 128 cell positions, then [64][nn] folded (even | odd) halves
    std::vector<int> near_off;   // host copy (window maxima)
    int ncoin = 0;               // (cell, station) pairs with the station exactly on the cell centre, sorted by cell
    int* d_coin = nullptr;       // [3][ncoin]: row, column, station
    // scratch (grown on demand)
    double* d_G = nullptr;  size_t G_cap = 0;      // node sums [patch][chunk][ky][kx][num|den][16]
    double* d_rv = nullptr; size_t rv_cap = 0;     // [T_pad][S_pad][2] (residual, validity) and its station-major transpose
    double* d_wn = nullptr; size_t wn_cap = 0;     // near weights beyond the two kept in registers
    // optional per-launch timing (smrfb_idw_far_timing): three events per launch -- before the node kernel, between the
    // node and the main kernel, after the main kernel -- on the launch stream
    bool timing = false;
    std::vector<cudaEvent_t> tev;
    void release();
};

// Decide whether the far-field path applies (uniform mesh, power in (0, 4], more than 16 stations) and build the plan.
// gx / gy / mx / my are host arrays.
int idw_far_plan(IdwFarPlan* plan, int S, const double* mx, const double* my, int ny, int nx, int grid_kind,
                 const double* gx, const double* gy, double power, int device);

// Distribute T steps over whole rows [row0, row0 + nrows) of the mesh.  rec / valid as written by prep_steps_kernel.
int idw_far_run(cudaStream_t st, IdwFarPlan* plan, const Geom& g, const double* d_mx, const double* d_my, int S,
                double power, const double* d_rec, int RS, int S_pad, const uint8_t* d_valid, int T, int detrend,
                int clamp, double vmin, double vmax, double* d_out, int out_f32, int row0, int nrows, int device);

}  // namespace smrfb
