// grid.cu -- GRID linear interpolation (smrf/spatial/grid.py:179-231 over scipy.interpolate.griddata).
//
// The Delaunay triangulation and the simplex of every DEM cell are built once on the host with
// scipy (as the reference does, except that it redoes it on every call) and uploaded at create
// time.  Per batch of T timesteps one kernel runs:
//   per cell (once):  barycentric coordinates from the simplex transform
//                       c0 = T00*(x-r0) + T01*(y-r1);  c1 = T10*(x-r0) + T11*(y-r1);  c2 = (1-c0)-c1
//   per step:         out = ((0 + c0*v_a) + c1*v_b) + c2*v_c  [+ p0*Z + p1]  -> clamp -> streaming store
// evaluated with separate multiplies and adds (no FMA contraction), the order that reproduces
// griddata(method='linear') bit for bit.  Cells outside the hull are NaN.
//
// Residual values are read from the transposed block vt[point][T_pad] (written by a small
// transpose kernel after station prep) so that one thread reads its three vertices' values for
// four consecutive steps with 2 x 128-bit loads each; neighbouring cells share a simplex, so these
// are L1 broadcast hits and the kernel is bound by the output stream to HBM.
#include "common.cuh"

namespace smrfb {

constexpr int GRID_THREADS = 256;
constexpr int GRID_TU = 4;   // steps per inner iteration

struct GridHandle : HandleBase {
    int nsimp = 0;
    int32_t* d_simplices = nullptr;    // [nsimp][3]
    double* d_transform = nullptr;     // [nsimp][3][2]
    int32_t* d_cell_simplex = nullptr; // [ncell]
    int32_t* d_cell_nearest = nullptr; // [ncell], optional (grid_method 'nearest')
    RecLayout L;
    double* d_vt = nullptr;            // [S][T_pad] transposed residuals
    size_t vt_cap = 0;
    double* d_pvt = nullptr;           // [T_pad][2]
    size_t pvt_cap = 0;
    double* d_l4 = nullptr;            // grid_local: [T][4][S] residual, slope, intercept, 0 per step
    size_t l4_cap = 0;
    double* d_l3 = nullptr;            // grid_local: staging of the caller's [T][3][S] block
    int local_exact = 1;               // grid_local: 1 = griddata's operation order, bit for bit; 0 = folded + fused (smrfb_grid_set_local_exact)
    size_t l3_cap = 0;
    double* d_ctgeo = nullptr;         // cubic: [nsimp][9] edge vectors e12, e23, e31 and cross-edge coefficients g
    double* d_ctw = nullptr;           // cubic: [9][ncell] Clough-Tocher patch weights of every cell
    double* d_cf = nullptr;            // cubic: staging of the caller's [T][nf][3][S] block (+ [T][2] coefficients)
    size_t cf_cap = 0;
    double* d_cvt = nullptr;           // cubic: [S][T][4 nf] (value, d/dx, d/dy, 0) per point, step and field
    size_t cvt_cap = 0;
    GridHandle() : HandleBase(HK_GRID) {}
    ~GridHandle() override {
        cudaSetDevice(device);
        cudaFree(d_simplices);
        cudaFree(d_transform);
        cudaFree(d_cell_simplex);
        cudaFree(d_cell_nearest);
        cudaFree(d_vt);
        cudaFree(d_pvt);
        cudaFree(d_l4);
        cudaFree(d_l3);
        cudaFree(d_ctgeo);
        cudaFree(d_ctw);
        cudaFree(d_cf);
        cudaFree(d_cvt);
    }
};

// rec[t][s] -> vt[s][t] (+ pv[t] = p0,p1), tiled through shared memory.
__global__ void grid_transpose_kernel(const double* __restrict__ rec, RecLayout L, int T_pad, double* __restrict__ vt,
                                      double* __restrict__ pvt) {
    __shared__ double tile[32][33];
    int s0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int t = t0 + r, s = s0 + threadIdx.x;
        tile[r][threadIdx.x] = (t < T_pad && s < L.S) ? rec[(size_t)t * L.RS + s] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int s = s0 + r, t = t0 + threadIdx.x;
        if (s < L.S && t < T_pad) vt[(size_t)s * T_pad + t] = tile[threadIdx.x][r];
    }
    if (blockIdx.x == 0 && threadIdx.y == 0) {
        int t = t0 + threadIdx.x;
        if (t < T_pad) {
            pvt[2 * t] = rec[(size_t)t * L.RS + L.meta()];
            pvt[2 * t + 1] = rec[(size_t)t * L.RS + L.meta() + 1];
        }
    }
}

struct GridArgs {
    Geom g;
    const int32_t* simplices;
    const double* transform;
    const int32_t* cell_simplex;
    const int32_t* cell_nearest;   // non-null: nearest-point mode
    const double* vt;
    const double* pvt;
    int T, T_pad, detrend;
    double vmin, vmax;
    double* out;
    int out_f32;                   // 1: out is float32 [T][ncell]
};

__global__ void __launch_bounds__(GRID_THREADS) grid_distribute_kernel(GridArgs a) {
    const long long c = (long long)blockIdx.x * GRID_THREADS + threadIdx.x;
    if (c >= a.g.ncell) return;
    const size_t ncell = (size_t)a.g.ncell;
    const int f32 = a.out_f32;
    if (a.cell_nearest) {   // grid_method 'nearest': the value of the nearest source point, no hull, no arithmetic on it
        const double* vn = a.vt + (size_t)__ldg(a.cell_nearest + c) * a.T_pad;
        const double Zn = (a.detrend && a.g.dem) ? __ldg(a.g.dem + c) : 0.0;
        for (int t0 = 0; t0 < a.T; t0 += 2) {
            const double2 x = __ldg(reinterpret_cast<const double2*>(vn + t0));
            const double2 pa = __ldg(reinterpret_cast<const double2*>(a.pvt + 2 * t0));
            const double2 pb = __ldg(reinterpret_cast<const double2*>(a.pvt + 2 * t0 + 2));
            double v0 = x.x, v1 = x.y;
            if (a.detrend) {
                v0 = __dadd_rn(__dadd_rn(v0, __dmul_rn(pa.x, Zn)), pa.y);
                v1 = __dadd_rn(__dadd_rn(v1, __dmul_rn(pb.x, Zn)), pb.y);
            }
            store_out(a.out, (size_t)t0 * ncell + c, clamp_nan_keep(v0, a.vmin, a.vmax), f32);
            if (t0 + 1 < a.T) store_out(a.out, (size_t)(t0 + 1) * ncell + c, clamp_nan_keep(v1, a.vmin, a.vmax), f32);
        }
        return;
    }
    const int sx = __ldg(a.cell_simplex + c);
    if (sx < 0) {
        const double qnan = nan("");
        for (int t = 0; t < a.T; t++) store_out(a.out, (size_t)t * ncell + c, qnan, f32);
        return;
    }
    double X, Y;
    cell_xy(a.g, c, X, Y);
    const double* Tm = a.transform + (size_t)sx * 6;
    const double dx = __dsub_rn(X, __ldg(Tm + 4)), dy = __dsub_rn(Y, __ldg(Tm + 5));
    const double c0 = __dadd_rn(__dmul_rn(__ldg(Tm + 0), dx), __dmul_rn(__ldg(Tm + 1), dy));
    const double c1 = __dadd_rn(__dmul_rn(__ldg(Tm + 2), dx), __dmul_rn(__ldg(Tm + 3), dy));
    const double c2 = __dsub_rn(__dsub_rn(1.0, c0), c1);
    const double* va = a.vt + (size_t)__ldg(a.simplices + 3 * sx + 0) * a.T_pad;
    const double* vb = a.vt + (size_t)__ldg(a.simplices + 3 * sx + 1) * a.T_pad;
    const double* vc = a.vt + (size_t)__ldg(a.simplices + 3 * sx + 2) * a.T_pad;
    const double Z = (a.detrend && a.g.dem) ? __ldg(a.g.dem + c) : 0.0;

    for (int t0 = 0; t0 < a.T; t0 += GRID_TU) {   // T_pad is a multiple of GRID_TU: vector loads stay in bounds
        double xa[GRID_TU], xb[GRID_TU], xc[GRID_TU], pv[2 * GRID_TU];
#pragma unroll
        for (int u = 0; u < GRID_TU; u += 2) {
            double2 ta = __ldg(reinterpret_cast<const double2*>(va + t0 + u));
            double2 tb = __ldg(reinterpret_cast<const double2*>(vb + t0 + u));
            double2 tc = __ldg(reinterpret_cast<const double2*>(vc + t0 + u));
            xa[u] = ta.x; xa[u + 1] = ta.y;
            xb[u] = tb.x; xb[u + 1] = tb.y;
            xc[u] = tc.x; xc[u + 1] = tc.y;
        }
#pragma unroll
        for (int u = 0; u < GRID_TU; u++) {
            double2 p = __ldg(reinterpret_cast<const double2*>(a.pvt + 2 * (t0 + u)));
            pv[2 * u] = p.x;
            pv[2 * u + 1] = p.y;
        }
#pragma unroll
        for (int u = 0; u < GRID_TU; u++) {
            if (t0 + u < a.T) {
                double v = __dadd_rn(0.0, __dmul_rn(c0, xa[u]));
                v = __dadd_rn(v, __dmul_rn(c1, xb[u]));
                v = __dadd_rn(v, __dmul_rn(c2, xc[u]));
                if (a.detrend) v = __dadd_rn(__dadd_rn(v, __dmul_rn(pv[2 * u], Z)), pv[2 * u + 1]);  // grid.py:210
                store_out(a.out, (size_t)(t0 + u) * ncell + c, clamp_nan_keep(v, a.vmin, a.vmax), f32);
            }
        }
    }
}

// Two adjacent cells per thread, 128-bit streaming stores (needs an even cell count so every row stays 16-byte
// aligned): each (CTA, step) writes 4 KB contiguous instead of 2 KB.
struct GridCell {
    double c0, c1, c2, Z;
    const double *va, *vb, *vc;
    bool inside;
};

__device__ __forceinline__ GridCell grid_cell_setup(const GridArgs& a, long long c) {
    GridCell g;
    const int sx = __ldg(a.cell_simplex + c);
    g.inside = sx >= 0;
    const int s0 = g.inside ? sx : 0;
    double X, Y;
    cell_xy(a.g, c, X, Y);
    const double* Tm = a.transform + (size_t)s0 * 6;
    const double dx = __dsub_rn(X, __ldg(Tm + 4)), dy = __dsub_rn(Y, __ldg(Tm + 5));
    g.c0 = __dadd_rn(__dmul_rn(__ldg(Tm + 0), dx), __dmul_rn(__ldg(Tm + 1), dy));
    g.c1 = __dadd_rn(__dmul_rn(__ldg(Tm + 2), dx), __dmul_rn(__ldg(Tm + 3), dy));
    g.c2 = __dsub_rn(__dsub_rn(1.0, g.c0), g.c1);
    g.va = a.vt + (size_t)__ldg(a.simplices + 3 * s0 + 0) * a.T_pad;
    g.vb = a.vt + (size_t)__ldg(a.simplices + 3 * s0 + 1) * a.T_pad;
    g.vc = a.vt + (size_t)__ldg(a.simplices + 3 * s0 + 2) * a.T_pad;
    g.Z = (a.detrend && a.g.dem) ? __ldg(a.g.dem + c) : 0.0;
    return g;
}

__global__ void __launch_bounds__(GRID_THREADS) grid_distribute2_kernel(GridArgs a) {
    const long long c = ((long long)blockIdx.x * GRID_THREADS + threadIdx.x) * 2;
    if (c >= a.g.ncell) return;                       // ncell is even: c + 1 is a valid cell too
    const GridCell g0 = grid_cell_setup(a, c), g1 = grid_cell_setup(a, c + 1);
    const size_t ncell = (size_t)a.g.ncell;
    const int f32 = a.out_f32;
    const double qnan = nan("");
#pragma unroll 1
    for (int t0 = 0; t0 < a.T; t0 += 2) {             // T_pad is even: vector loads stay in bounds
        const double2 a0 = __ldg(reinterpret_cast<const double2*>(g0.va + t0));
        const double2 b0 = __ldg(reinterpret_cast<const double2*>(g0.vb + t0));
        const double2 q0 = __ldg(reinterpret_cast<const double2*>(g0.vc + t0));
        const double2 a1 = __ldg(reinterpret_cast<const double2*>(g1.va + t0));
        const double2 b1 = __ldg(reinterpret_cast<const double2*>(g1.vb + t0));
        const double2 q1 = __ldg(reinterpret_cast<const double2*>(g1.vc + t0));
        const double2 pa = __ldg(reinterpret_cast<const double2*>(a.pvt + 2 * t0));
        const double2 pb = __ldg(reinterpret_cast<const double2*>(a.pvt + 2 * t0 + 2));
        double v00 = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(g0.c0, a0.x)), __dmul_rn(g0.c1, b0.x)), __dmul_rn(g0.c2, q0.x));
        double v01 = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(g0.c0, a0.y)), __dmul_rn(g0.c1, b0.y)), __dmul_rn(g0.c2, q0.y));
        double v10 = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(g1.c0, a1.x)), __dmul_rn(g1.c1, b1.x)), __dmul_rn(g1.c2, q1.x));
        double v11 = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(g1.c0, a1.y)), __dmul_rn(g1.c1, b1.y)), __dmul_rn(g1.c2, q1.y));
        if (a.detrend) {                                // grid.py:210, left to right
            v00 = __dadd_rn(__dadd_rn(v00, __dmul_rn(pa.x, g0.Z)), pa.y);
            v10 = __dadd_rn(__dadd_rn(v10, __dmul_rn(pa.x, g1.Z)), pa.y);
            v01 = __dadd_rn(__dadd_rn(v01, __dmul_rn(pb.x, g0.Z)), pb.y);
            v11 = __dadd_rn(__dadd_rn(v11, __dmul_rn(pb.x, g1.Z)), pb.y);
        }
        v00 = g0.inside ? clamp_nan_keep(v00, a.vmin, a.vmax) : qnan;
        v01 = g0.inside ? clamp_nan_keep(v01, a.vmin, a.vmax) : qnan;
        v10 = g1.inside ? clamp_nan_keep(v10, a.vmin, a.vmax) : qnan;
        v11 = g1.inside ? clamp_nan_keep(v11, a.vmin, a.vmax) : qnan;
        store_out2(a.out, (size_t)t0 * ncell + c, v00, v10, f32);
        if (t0 + 1 < a.T) store_out2(a.out, (size_t)(t0 + 1) * ncell + c, v01, v11, f32);
    }
}

// ------------------------------------------------------------------------------------------ grid_local
// detrendedInterpolationLocal (grid.py:115-177, grid_method 'linear'): three point fields per step -- the residual,
// the slope and the intercept of the per-point local fits (host side, smrf_b200/spatial/grid.py) -- are interpolated
// with the same barycentric weights and combined per cell:
//   out = (interp(residual) + interp(slope) * Z) + interp(intercept)                                (grid.py:175)
// each interpolation in griddata's operation order.  vt[point][4 t + f]: f = 0 residual, 1 slope, 2 intercept.
struct GridLocalCell {
    double c0, c1, c2, Z;
    const double *va, *vb, *vc;
    bool inside;
};

__device__ __forceinline__ GridLocalCell grid_local_setup(const GridArgs& a, long long c) {
    GridLocalCell g;
    const int sx = __ldg(a.cell_simplex + c);
    g.inside = sx >= 0;
    const int s0 = g.inside ? sx : 0;
    double X, Y;
    cell_xy(a.g, c, X, Y);
    const double* Tm = a.transform + (size_t)s0 * 6;
    const double dx = __dsub_rn(X, __ldg(Tm + 4)), dy = __dsub_rn(Y, __ldg(Tm + 5));
    g.c0 = __dadd_rn(__dmul_rn(__ldg(Tm + 0), dx), __dmul_rn(__ldg(Tm + 1), dy));
    g.c1 = __dadd_rn(__dmul_rn(__ldg(Tm + 2), dx), __dmul_rn(__ldg(Tm + 3), dy));
    g.c2 = __dsub_rn(__dsub_rn(1.0, g.c0), g.c1);
    g.va = a.vt + (size_t)__ldg(a.simplices + 3 * s0 + 0) * a.T_pad;     // T_pad: row length of vt (>= 4 T)
    g.vb = a.vt + (size_t)__ldg(a.simplices + 3 * s0 + 1) * a.T_pad;
    g.vc = a.vt + (size_t)__ldg(a.simplices + 3 * s0 + 2) * a.T_pad;
    g.Z = a.g.dem ? __ldg(a.g.dem + c) : 0.0;
    return g;
}

__device__ __forceinline__ double grid_local_value(const GridLocalCell& g, int t, double vmin, double vmax) {
    const double2 a01 = __ldg(reinterpret_cast<const double2*>(g.va + 4 * t));
    const double2 b01 = __ldg(reinterpret_cast<const double2*>(g.vb + 4 * t));
    const double2 c01 = __ldg(reinterpret_cast<const double2*>(g.vc + 4 * t));
    const double a2 = __ldg(g.va + 4 * t + 2), b2 = __ldg(g.vb + 4 * t + 2), c2 = __ldg(g.vc + 4 * t + 2);
    const double res = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(g.c0, a01.x)), __dmul_rn(g.c1, b01.x)), __dmul_rn(g.c2, c01.x));
    const double slo = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(g.c0, a01.y)), __dmul_rn(g.c1, b01.y)), __dmul_rn(g.c2, c01.y));
    const double icp = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(g.c0, a2)), __dmul_rn(g.c1, b2)), __dmul_rn(g.c2, c2));
    const double v = __dadd_rn(__dadd_rn(res, __dmul_rn(slo, g.Z)), icp);
    return g.inside ? clamp_nan_keep(v, vmin, vmax) : nan("");
}

__global__ void __launch_bounds__(GRID_THREADS) grid_local_kernel(GridArgs a) {          // one cell per thread
    const long long c = (long long)blockIdx.x * GRID_THREADS + threadIdx.x;
    if (c >= a.g.ncell) return;
    const GridLocalCell g = grid_local_setup(a, c);
    const size_t ncell = (size_t)a.g.ncell;
    for (int t = 0; t < a.T; t++) store_out(a.out, (size_t)t * ncell + c, grid_local_value(g, t, a.vmin, a.vmax), a.out_f32);
}

__global__ void __launch_bounds__(GRID_THREADS) grid_local2_kernel(GridArgs a) {         // two adjacent cells, 128-bit stores
    const long long c = ((long long)blockIdx.x * GRID_THREADS + threadIdx.x) * 2;
    if (c >= a.g.ncell) return;                       // ncell is even: c + 1 is a valid cell too
    const GridLocalCell g0 = grid_local_setup(a, c), g1 = grid_local_setup(a, c + 1);
    const size_t ncell = (size_t)a.g.ncell;
    const int f32 = a.out_f32;
#pragma unroll 2
    for (int t = 0; t < a.T; t++)
        store_out2(a.out, (size_t)t * ncell + c, grid_local_value(g0, t, a.vmin, a.vmax), grid_local_value(g1, t, a.vmin, a.vmax), f32);
}

// ---- grid_local, folded form (smrfb_grid_set_local_exact(h, 0); what the block path uses with its closed-form fits, which are
// not the reference's bit for bit either).  ncu on grid_local2_kernel: L1 wavefronts 98 % of peak (nine 8-byte vertex values per
// cell-step: 18 wavefronts per 32 outputs), fp64 pipe 41 %.  Interpolation is linear, so
//     out = interp(residual) + interp(slope) Z + interp(intercept) = interp(residual + intercept) + Z interp(slope):
// two fields per vertex (ONE 128-bit load) instead of three, and with fused multiply-adds 7 fp64 instructions per output
// instead of 24.  The result differs from the three-interpolation order by a few ulp of the largest term (contract 1e-9).
// l2[2 t + f][s]: f = 0 residual + intercept, f = 1 slope
__global__ void grid_local_fold_kernel(const double* __restrict__ f3, int T, int S, double* __restrict__ l2) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)T * S) return;
    const int t = (int)(i / S), s = (int)(i - (long long)t * S);
    const double* p = f3 + (size_t)t * 3 * S;
    l2[(size_t)(2 * t) * S + s] = p[s] + p[2 * (size_t)S + s];
    l2[(size_t)(2 * t + 1) * S + s] = p[(size_t)S + s];
}

template <typename OutT>
__global__ void __launch_bounds__(GRID_THREADS) grid_local_fold2_kernel(GridArgs a) {    // two adjacent cells per thread
    const long long c = ((long long)blockIdx.x * GRID_THREADS + threadIdx.x) * 2;
    if (c >= a.g.ncell) return;                       // ncell is even: c + 1 is a valid cell too
    const GridLocalCell g0 = grid_local_setup(a, c), g1 = grid_local_setup(a, c + 1);    // va / vb / vc: rows of vt, here [2 t + f]
    const size_t ncell = (size_t)a.g.ncell;
    OutT* op = reinterpret_cast<OutT*>(a.out) + c;
    const double qnan = nan("");
#pragma unroll 2
    for (int t = 0; t < a.T; t++, op += ncell) {
        const double2 a0 = __ldg(reinterpret_cast<const double2*>(g0.va + 2 * t));
        const double2 b0 = __ldg(reinterpret_cast<const double2*>(g0.vb + 2 * t));
        const double2 c0 = __ldg(reinterpret_cast<const double2*>(g0.vc + 2 * t));
        const double2 a1 = __ldg(reinterpret_cast<const double2*>(g1.va + 2 * t));
        const double2 b1 = __ldg(reinterpret_cast<const double2*>(g1.vb + 2 * t));
        const double2 c1 = __ldg(reinterpret_cast<const double2*>(g1.vc + 2 * t));
        const double r0 = fma(g0.c2, c0.x, fma(g0.c1, b0.x, g0.c0 * a0.x)), s0 = fma(g0.c2, c0.y, fma(g0.c1, b0.y, g0.c0 * a0.y));
        const double r1 = fma(g1.c2, c1.x, fma(g1.c1, b1.x, g1.c0 * a1.x)), s1 = fma(g1.c2, c1.y, fma(g1.c1, b1.y, g1.c0 * a1.y));
        const double v0 = g0.inside ? clamp_nan_keep(fma(s0, g0.Z, r0), a.vmin, a.vmax) : qnan;
        const double v1 = g1.inside ? clamp_nan_keep(fma(s1, g1.Z, r1), a.vmin, a.vmax) : qnan;
        if constexpr (sizeof(OutT) == 4) __stcs(reinterpret_cast<float2*>(op), make_float2((float)v0, (float)v1));
        else __stcs(reinterpret_cast<double2*>(op), make_double2(v0, v1));
    }
}

template <typename OutT>
__global__ void __launch_bounds__(GRID_THREADS) grid_local_fold1_kernel(GridArgs a) {    // one cell per thread (odd cell counts)
    const long long c = (long long)blockIdx.x * GRID_THREADS + threadIdx.x;
    if (c >= a.g.ncell) return;
    const GridLocalCell g = grid_local_setup(a, c);
    const size_t ncell = (size_t)a.g.ncell;
    OutT* op = reinterpret_cast<OutT*>(a.out) + c;
    for (int t = 0; t < a.T; t++, op += ncell) {
        const double2 a0 = __ldg(reinterpret_cast<const double2*>(g.va + 2 * t));
        const double2 b0 = __ldg(reinterpret_cast<const double2*>(g.vb + 2 * t));
        const double2 c0 = __ldg(reinterpret_cast<const double2*>(g.vc + 2 * t));
        const double r = fma(g.c2, c0.x, fma(g.c1, b0.x, g.c0 * a0.x)), sl = fma(g.c2, c0.y, fma(g.c1, b0.y, g.c0 * a0.y));
        const double v = g.inside ? clamp_nan_keep(fma(sl, g.Z, r), a.vmin, a.vmax) : nan("");
        __stcs(op, (OutT)v);
    }
}

// Staged form of the folded kernel (default).  ncu on grid_local_fold2_kernel: L1 wavefronts 97 % of peak -- every lane asks for
// 96 bytes per step although the 64 cells of a warp sit in one to three triangles.  Along the linear cell index a CTA's 512
// cells fall into a few RUNS of equal simplex; the CTA lists its runs (a block scan of "my simplex differs from my left
// neighbour's"), stages the runs' vertex values for 16 steps at a time in shared memory (coalesced 256-byte reads per vertex), and
// the step loop reads them with broadcast 128-bit shared loads: ~12 wavefronts per warp-step instead of ~38.  More than GLS_MAXR
// runs in a CTA (a source lattice finer than ~13 cells): the direct loads of the kernel above.
constexpr int GLS_TC = 16;
constexpr int GLS_MAXR = 40;
constexpr int GLS_RS = GLS_TC + 1;       // row stride in double2: distinct rows start in different banks

template <typename OutT>
__global__ void __launch_bounds__(GRID_THREADS, 4) grid_local_fold_staged_kernel(GridArgs a) {
    __shared__ double2 vals[GLS_MAXR * 3 * GLS_RS];
    __shared__ int vid[GLS_MAXR * 3];
    __shared__ int wsum[GRID_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const long long c = ((long long)blockIdx.x * GRID_THREADS + tid) * 2;
    const bool valid = c < a.g.ncell;                 // ncell is even: c + 1 is a valid cell too
    const int sx0 = valid ? __ldg(a.cell_simplex + c) : -1, sx1 = valid ? __ldg(a.cell_simplex + c + 1) : -1;
    const bool f0 = valid && (tid == 0 || sx0 != __ldg(a.cell_simplex + c - 1)), f1 = valid && sx1 != sx0;
    // inclusive scan of the run starts over the CTA
    int inc = (int)f0 + (int)f1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) wsum[wp] = inc;
    __syncthreads();
    int base = 0, nruns = 0;
#pragma unroll
    for (int i = 0; i < GRID_THREADS / 32; i++) {
        const int v = wsum[i];
        if (i < wp) base += v;
        nruns += v;
    }
    const int r1 = base + inc - 1, r0 = r1 - (int)f1;       // run of cell c + 1, of cell c
    const bool staged = nruns <= GLS_MAXR;                   // uniform
    if (staged) {
        if (f0) {
#pragma unroll
            for (int v = 0; v < 3; v++) vid[3 * r0 + v] = sx0 >= 0 ? __ldg(a.simplices + 3 * sx0 + v) : -1;
        }
        if (f1) {
#pragma unroll
            for (int v = 0; v < 3; v++) vid[3 * r1 + v] = sx1 >= 0 ? __ldg(a.simplices + 3 * sx1 + v) : -1;
        }
    }
    GridLocalCell g0, g1;
    if (valid) {
        g0 = grid_local_setup(a, c);
        g1 = grid_local_setup(a, c + 1);
    }
    const size_t ncell = (size_t)a.g.ncell;
    OutT* op = reinterpret_cast<OutT*>(a.out) + c;
    const double qnan = nan("");
    auto finish = [&](const double2 a0, const double2 b0, const double2 c0, const double2 a1, const double2 b1, const double2 c1) {
        const double r0v = fma(g0.c2, c0.x, fma(g0.c1, b0.x, g0.c0 * a0.x)), s0v = fma(g0.c2, c0.y, fma(g0.c1, b0.y, g0.c0 * a0.y));
        const double r1v = fma(g1.c2, c1.x, fma(g1.c1, b1.x, g1.c0 * a1.x)), s1v = fma(g1.c2, c1.y, fma(g1.c1, b1.y, g1.c0 * a1.y));
        const double v0 = g0.inside ? clamp_nan_keep(fma(s0v, g0.Z, r0v), a.vmin, a.vmax) : qnan;
        const double v1 = g1.inside ? clamp_nan_keep(fma(s1v, g1.Z, r1v), a.vmin, a.vmax) : qnan;
        if constexpr (sizeof(OutT) == 4) __stcs(reinterpret_cast<float2*>(op), make_float2((float)v0, (float)v1));
        else __stcs(reinterpret_cast<double2*>(op), make_double2(v0, v1));
    };
    if (!staged) {
        if (!valid) return;
#pragma unroll 2
        for (int t = 0; t < a.T; t++, op += ncell)
            finish(__ldg(reinterpret_cast<const double2*>(g0.va + 2 * t)), __ldg(reinterpret_cast<const double2*>(g0.vb + 2 * t)),
                   __ldg(reinterpret_cast<const double2*>(g0.vc + 2 * t)), __ldg(reinterpret_cast<const double2*>(g1.va + 2 * t)),
                   __ldg(reinterpret_cast<const double2*>(g1.vb + 2 * t)), __ldg(reinterpret_cast<const double2*>(g1.vc + 2 * t)));
        return;
    }
    const double2* p0 = vals + (size_t)(3 * r0) * GLS_RS;
    const double2* p1 = vals + (size_t)(3 * r1) * GLS_RS;
    const int nitem = nruns * 3 * GLS_TC;
    for (int t0 = 0; t0 < a.T; t0 += GLS_TC) {
        __syncthreads();          // the previous chunk has been consumed (first pass: vid is complete)
        for (int i = tid; i < nitem; i += GRID_THREADS) {
            const int row = i / GLS_TC, tc = i - row * GLS_TC;
            const int v = vid[row];
            vals[row * GLS_RS + tc] = (v >= 0 && t0 + tc < a.T) ? __ldg(reinterpret_cast<const double2*>(a.vt + (size_t)v * a.T_pad) + t0 + tc)
                                                                 : make_double2(0.0, 0.0);
        }
        __syncthreads();
        if (valid) {
            const int tn = min(GLS_TC, a.T - t0);
#pragma unroll 2
            for (int tc = 0; tc < tn; tc++, op += ncell)
                finish(p0[tc], p0[GLS_RS + tc], p0[2 * GLS_RS + tc], p1[tc], p1[GLS_RS + tc], p1[2 * GLS_RS + tc]);
        }
    }
}

// d_f3: [T][3][S] on the device (residual, slope, intercept rows per step)
static int grid_local_run(GridHandle* h, const double* d_f3, int T, double vmin, double vmax, double* d_out, cudaStream_t st) {
    SMRFB_CHECK(T > 0, SMRFB_ERR_ARG, "T must be > 0");
    SMRFB_CHECK(h->g.dem, SMRFB_ERR_ARG, "grid_local needs GridZ");
    const size_t S = (size_t)h->S;
    const bool fold = !h->local_exact;
    const int R = (fold ? 2 : 4) * T, R_pad = (R + 31) / 32 * 32;
    SMRFB_PROPAGATE(ensure_cap((void**)&h->d_l4, &h->l4_cap, (size_t)R * S * sizeof(double)));
    SMRFB_PROPAGATE(ensure_cap((void**)&h->d_rec, &h->rec_cap, (size_t)R_pad * h->L.RS * sizeof(double)));
    SMRFB_PROPAGATE(ensure_cap((void**)&h->d_vt, &h->vt_cap, (size_t)R_pad * S * sizeof(double)));
    SMRFB_PROPAGATE(ensure_cap((void**)&h->d_pvt, &h->pvt_cap, (size_t)R_pad * 2 * sizeof(double)));
    if (fold) {
        const long long ne = (long long)T * (long long)S;
        grid_local_fold_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(d_f3, T, (int)S, h->d_l4);
        count_launch();
        SMRFB_CUDA(cudaGetLastError());
    } else {
        SMRFB_CUDA(cudaMemsetAsync(h->d_l4, 0, (size_t)R * S * sizeof(double), st));
        SMRFB_CUDA(cudaMemcpy2DAsync(h->d_l4, 4 * S * sizeof(double), d_f3, 3 * S * sizeof(double), 3 * S * sizeof(double), (size_t)T,
                                     cudaMemcpyDeviceToDevice, st));
    }
    SMRFB_PROPAGATE(launch_prep(st, h->d_l4, R, R_pad, h->S, h->L, nullptr, nullptr, 0, 0, nullptr, nullptr, h->d_rec, nullptr,
                                h->d_flags));
    dim3 tb(32, 8), tg((h->S + 31) / 32, R_pad / 32);
    grid_transpose_kernel<<<tg, tb, 0, st>>>(h->d_rec, h->L, R_pad, h->d_vt, h->d_pvt);
    count_launch();
    SMRFB_CUDA(cudaGetLastError());
    GridArgs a;
    a.g = h->g;
    a.simplices = h->d_simplices;
    a.transform = h->d_transform;
    a.cell_simplex = h->d_cell_simplex;
    a.cell_nearest = nullptr;
    a.vt = h->d_vt;
    a.pvt = h->d_pvt;
    a.T = T;
    a.T_pad = R_pad;
    a.detrend = 1;
    a.vmin = vmin;
    a.vmax = vmax;
    a.out = d_out;
    a.out_f32 = h->out_f32;
    const long long nblk = (h->g.ncell + GRID_THREADS - 1) / GRID_THREADS;
    SMRFB_CHECK(nblk < 2147483647LL, SMRFB_ERR_LIMIT, "too many cells for one launch");
    if (fold) {
        const long long nb2 = (h->g.ncell / 2 + GRID_THREADS - 1) / GRID_THREADS;
        const bool direct = getenv("SMRFB_GRID_FOLD_DIRECT") != nullptr;      // the kernel without staging, for comparison
        if ((h->g.ncell & 1) == 0 && !direct) {
            if (h->out_f32) grid_local_fold_staged_kernel<float><<<(unsigned)nb2, GRID_THREADS, 0, st>>>(a);
            else grid_local_fold_staged_kernel<double><<<(unsigned)nb2, GRID_THREADS, 0, st>>>(a);
        } else if ((h->g.ncell & 1) == 0) {
            if (h->out_f32) grid_local_fold2_kernel<float><<<(unsigned)nb2, GRID_THREADS, 0, st>>>(a);
            else grid_local_fold2_kernel<double><<<(unsigned)nb2, GRID_THREADS, 0, st>>>(a);
        } else {
            if (h->out_f32) grid_local_fold1_kernel<float><<<(unsigned)nblk, GRID_THREADS, 0, st>>>(a);
            else grid_local_fold1_kernel<double><<<(unsigned)nblk, GRID_THREADS, 0, st>>>(a);
        }
    } else if ((h->g.ncell & 1) == 0 && !getenv("SMRFB_GRID_ONE_CELL")) {
        const long long nb2 = (h->g.ncell / 2 + GRID_THREADS - 1) / GRID_THREADS;
        grid_local2_kernel<<<(unsigned)nb2, GRID_THREADS, 0, st>>>(a);
    } else
        grid_local_kernel<<<(unsigned)nblk, GRID_THREADS, 0, st>>>(a);
    count_launch();
    SMRFB_CUDA(cudaGetLastError());
    return SMRFB_OK;
}

// ------------------------------------------------------------------------------------------ grid_method 'cubic'
// scipy.interpolate.CloughTocher2DInterpolator (reached from grid.py:204-207, :226-229 through griddata(method='cubic')
// and from utils.py:447 on the cached triangulation): a C1 piecewise-cubic Bezier patch per Delaunay triangle, built
// from the three vertex values and three vertex gradients.  The vertex gradients are estimated on the host by SciPy's
// own iteration (its stopping point decides the low digits, so both sides must run the same code); the patch itself
// is LINEAR in those nine data, so a cell needs nine weights -- formed once per launch by evaluating the patch on
// the nine unit data, exactly SciPy's `_clough_tocher_2d_single` -- and a timestep costs nine multiply-adds:
//   out = sum_v ( wf[v] f_v + wx[v] df_v/dx + wy[v] df_v/dy )                     v = the simplex's three vertices
// nf = 1: out (+ p0 Z + p1: detrendedInterpolationMask, grid.py:210; the residuals and coefficients come from the host)
// nf = 2: out = patch(field 0) + Z * patch(field 1)      (detrendedInterpolationLocal, grid.py:175, with field 0 =
//         residual + intercept and field 1 = slope: the patch is linear, the three interpolations fold into two)
struct CtData {
    double f[3], d[6];
};

// SciPy `_clough_tocher_2d_single` for barycentrics b, geometry geo = {e12x, e12y, e23x, e23y, e31x, e31y, g0, g1, g2}
__device__ __forceinline__ double ct_patch(const double* b, const double* geo, const CtData& v) {
    const double e12x = geo[0], e12y = geo[1], e23x = geo[2], e23y = geo[3], e31x = geo[4], e31y = geo[5];
    const double f1 = v.f[0], f2 = v.f[1], f3 = v.f[2];
    const double df12 = +(v.d[0] * e12x + v.d[1] * e12y);
    const double df21 = -(v.d[2] * e12x + v.d[3] * e12y);
    const double df23 = +(v.d[2] * e23x + v.d[3] * e23y);
    const double df32 = -(v.d[4] * e23x + v.d[5] * e23y);
    const double df31 = +(v.d[4] * e31x + v.d[5] * e31y);
    const double df13 = -(v.d[0] * e31x + v.d[1] * e31y);
    const double c3000 = f1, c2100 = (df12 + 3 * c3000) / 3, c2010 = (df13 + 3 * c3000) / 3;
    const double c0300 = f2, c1200 = (df21 + 3 * c0300) / 3, c0210 = (df23 + 3 * c0300) / 3;
    const double c0030 = f3, c1020 = (df31 + 3 * c0030) / 3, c0120 = (df32 + 3 * c0030) / 3;
    const double c2001 = (c2100 + c2010 + c3000) / 3;
    const double c0201 = (c1200 + c0300 + c0210) / 3;
    const double c0021 = (c1020 + c0120 + c0030) / 3;
    const double c0111 = (geo[6] * (-c0300 + 3 * c0210 - 3 * c0120 + c0030) + (-c0300 + 2 * c0210 - c0120 + c0021 + c0201)) / 2;
    const double c1011 = (geo[7] * (-c0030 + 3 * c1020 - 3 * c2010 + c3000) + (-c0030 + 2 * c1020 - c2010 + c2001 + c0021)) / 2;
    const double c1101 = (geo[8] * (-c3000 + 3 * c2100 - 3 * c1200 + c0300) + (-c3000 + 2 * c2100 - c1200 + c2001 + c0201)) / 2;
    const double c1002 = (c1101 + c1011 + c2001) / 3;
    const double c0102 = (c1101 + c0111 + c0201) / 3;
    const double c0012 = (c1011 + c0111 + c0021) / 3;
    const double c0003 = (c1002 + c0102 + c0012) / 3;
    const double minval = fmin(b[0], fmin(b[1], b[2]));
    const double b1 = b[0] - minval, b2 = b[1] - minval, b3 = b[2] - minval, b4 = 3 * minval;
    return b1 * b1 * b1 * c3000 + 3 * b1 * b1 * b2 * c2100 + 3 * b1 * b1 * b3 * c2010 + 3 * b1 * b1 * b4 * c2001 +
           3 * b1 * b2 * b2 * c1200 + 6 * b1 * b2 * b4 * c1101 + 3 * b1 * b3 * b3 * c1020 + 6 * b1 * b3 * b4 * c1011 +
           3 * b1 * b4 * b4 * c1002 + b2 * b2 * b2 * c0300 + 3 * b2 * b2 * b3 * c0210 + 3 * b2 * b2 * b4 * c0201 +
           3 * b2 * b3 * b3 * c0120 + 6 * b2 * b3 * b4 * c0111 + 3 * b2 * b4 * b4 * c0102 + b3 * b3 * b3 * c0030 +
           3 * b3 * b3 * b4 * c0021 + 3 * b3 * b4 * b4 * c0012 + b4 * b4 * b4 * c0003;
}

struct GridCubicArgs {
    Geom g;
    const int32_t* simplices;
    const double* transform;
    const int32_t* cell_simplex;
    const double* ctgeo;     // [nsimp][9]
    const double* ctw;       // [9][ncell] patch weights per cell
    const double* cvt;       // [S][T][4 nf]
    const double* pv;        // [T][2] or null
    int T;
    double vmin, vmax;
    double* out;
    int out_f32;
};

// fields[t][q][s] (q = 3 f + {value, d/dx, d/dy}) -> cvt[s][t][4 f + {0, 1, 2}], pad slot = 0
__global__ void grid_cubic_pack_kernel(const double* __restrict__ fields, int T, int nf, int S, double* __restrict__ cvt) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
    if (s >= S) return;
    double* o = cvt + ((size_t)s * T + t) * (4 * nf);
    for (int f = 0; f < nf; f++) {
        const double* src = fields + ((size_t)t * nf + f) * 3 * S + s;
        reinterpret_cast<double2*>(o)[2 * f] = make_double2(src[0], src[S]);
        reinterpret_cast<double2*>(o)[2 * f + 1] = make_double2(src[2 * (size_t)S], 0.0);
    }
}

// Once per handle (smrfb_grid_set_cubic): the nine patch weights of every cell, ctw[i][ncell] (zeros outside the hull).
__global__ void __launch_bounds__(GRID_THREADS) grid_cubic_weights_kernel(GridCubicArgs a, double* __restrict__ ctw) {
    const long long c = (long long)blockIdx.x * GRID_THREADS + threadIdx.x;
    if (c >= a.g.ncell) return;
    const size_t ncell = (size_t)a.g.ncell;
    const int sx = __ldg(a.cell_simplex + c);
    double w[9];       // weights of f_0..2, then (d/dx, d/dy) of vertex 0, 1, 2
    if (sx < 0) {
#pragma unroll
        for (int i = 0; i < 9; i++) w[i] = 0.0;
    } else {
        double X, Y;
        cell_xy(a.g, c, X, Y);
        const double* Tm = a.transform + (size_t)sx * 6;
        const double dx = __dsub_rn(X, __ldg(Tm + 4)), dy = __dsub_rn(Y, __ldg(Tm + 5));
        double b[3];
        b[0] = __dadd_rn(__dmul_rn(__ldg(Tm + 0), dx), __dmul_rn(__ldg(Tm + 1), dy));
        b[1] = __dadd_rn(__dmul_rn(__ldg(Tm + 2), dx), __dmul_rn(__ldg(Tm + 3), dy));
        b[2] = __dsub_rn(__dsub_rn(1.0, b[0]), b[1]);
        double geo[9];
#pragma unroll
        for (int i = 0; i < 9; i++) geo[i] = __ldg(a.ctgeo + (size_t)sx * 9 + i);
#pragma unroll
        for (int i = 0; i < 9; i++) {     // the patch is linear in its nine data: evaluate it on the unit data
            CtData u;
#pragma unroll
            for (int j = 0; j < 3; j++) u.f[j] = (i == j) ? 1.0 : 0.0;
#pragma unroll
            for (int j = 0; j < 6; j++) u.d[j] = (i == 3 + j) ? 1.0 : 0.0;
            w[i] = ct_patch(b, geo, u);
        }
    }
#pragma unroll
    for (int i = 0; i < 9; i++) ctw[(size_t)i * ncell + c] = w[i];
}

template <int NF>
__global__ void __launch_bounds__(GRID_THREADS) grid_cubic_kernel(GridCubicArgs a) {
    const long long c = (long long)blockIdx.x * GRID_THREADS + threadIdx.x;
    if (c >= a.g.ncell) return;
    const size_t ncell = (size_t)a.g.ncell;
    const int f32 = a.out_f32;
    const int sx = __ldg(a.cell_simplex + c);
    if (sx < 0) {
        const double qnan = nan("");
        for (int t = 0; t < a.T; t++) store_out(a.out, (size_t)t * ncell + c, qnan, f32);
        return;
    }
    double w[9];
#pragma unroll
    for (int i = 0; i < 9; i++) w[i] = __ldg(a.ctw + (size_t)i * ncell + c);
    const size_t rowlen = (size_t)a.T * (4 * NF);
    const double* vp[3];
#pragma unroll
    for (int v = 0; v < 3; v++) vp[v] = a.cvt + (size_t)__ldg(a.simplices + 3 * sx + v) * rowlen;
    const double Z = a.g.dem ? __ldg(a.g.dem + c) : 0.0;
#pragma unroll 2
    for (int t = 0; t < a.T; t++) {
        double val[NF];
#pragma unroll
        for (int f = 0; f < NF; f++) {
            double acc = 0.0;
#pragma unroll
            for (int v = 0; v < 3; v++) {
                const double2 x0 = __ldg(reinterpret_cast<const double2*>(vp[v] + (size_t)t * (4 * NF) + 4 * f));
                const double2 x1 = __ldg(reinterpret_cast<const double2*>(vp[v] + (size_t)t * (4 * NF) + 4 * f + 2));
                acc = fma(w[v], x0.x, acc);
                acc = fma(w[3 + 2 * v], x0.y, acc);
                acc = fma(w[4 + 2 * v], x1.x, acc);
            }
            val[f] = acc;
        }
        double r = val[0];
        if (NF == 2) r = fma(Z, val[NF - 1], r);
        else if (a.pv) r = __dadd_rn(__dadd_rn(r, __dmul_rn(__ldg(a.pv + 2 * t), Z)), __ldg(a.pv + 2 * t + 1));   // grid.py:210
        store_out(a.out, (size_t)t * ncell + c, clamp_nan_keep(r, a.vmin, a.vmax), f32);
    }
}

// Staged form (default), as for grid_local above: a CTA's 256 consecutive cells fall into a few runs of equal simplex; the runs'
// vertex data (value and gradient, 4 NF doubles per vertex and step) are staged for 8 steps at a time in shared memory and read
// back with broadcast 128-bit loads.  The direct kernel asks L1 for 96 NF bytes per cell-step.
constexpr int GCS_TC = 8;
constexpr int GCS_MAXR = 24;

template <int NF, int CPT>      // CPT: adjacent cells per thread (2 when the cell count is even: they mostly share their simplex and its loads)
__global__ void __launch_bounds__(GRID_THREADS) grid_cubic_staged_kernel(GridCubicArgs a) {
    constexpr int Q = 2 * NF;                 // double2 per (vertex, step)
    constexpr int RS = GCS_TC * Q + 1;        // row stride in double2
    constexpr int MAXR = GCS_MAXR * CPT;
    __shared__ double2 vals[GCS_MAXR * 3 * RS * (CPT == 2 && NF == 2 ? 1 : CPT)];
    constexpr int MAXR_FIT = (int)(sizeof(vals) / sizeof(double2)) / (3 * RS);
    static_assert(MAXR_FIT >= GCS_MAXR, "stage too small");
    __shared__ int vid[MAXR * 3];
    __shared__ int wsum[GRID_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const long long c = ((long long)blockIdx.x * GRID_THREADS + tid) * CPT;
    const bool valid = c < a.g.ncell;         // CPT == 2: the cell count is even, c + 1 is a valid cell too
    int sx[CPT];
    bool fl[CPT];
#pragma unroll
    for (int j = 0; j < CPT; j++) sx[j] = valid ? __ldg(a.cell_simplex + c + j) : -1;
    fl[0] = valid && (tid == 0 || sx[0] != __ldg(a.cell_simplex + c - 1));
    if (CPT == 2) fl[CPT - 1] = valid && sx[CPT - 1] != sx[0];
    int inc = 0;
#pragma unroll
    for (int j = 0; j < CPT; j++) inc += (int)fl[j];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) wsum[wp] = inc;
    __syncthreads();
    int base = 0, nruns = 0;
#pragma unroll
    for (int i = 0; i < GRID_THREADS / 32; i++) {
        const int v = wsum[i];
        if (i < wp) base += v;
        nruns += v;
    }
    int r[CPT];
    r[CPT - 1] = base + inc - 1;
    if (CPT == 2) r[0] = r[CPT - 1] - (int)fl[CPT - 1];
    const bool staged = nruns <= MAXR_FIT;     // uniform
    if (staged) {
#pragma unroll
        for (int j = 0; j < CPT; j++)
            if (fl[j]) {
#pragma unroll
                for (int v = 0; v < 3; v++) vid[3 * r[j] + v] = sx[j] >= 0 ? __ldg(a.simplices + 3 * sx[j] + v) : -1;
            }
    }
    const size_t ncell = (size_t)a.g.ncell;
    const int f32 = a.out_f32;
    bool inside[CPT];
    double w[CPT][9], Z[CPT];
#pragma unroll
    for (int j = 0; j < CPT; j++) {
        inside[j] = valid && sx[j] >= 0;
#pragma unroll
        for (int i = 0; i < 9; i++) w[j][i] = inside[j] ? __ldg(a.ctw + (size_t)i * ncell + c + j) : 0.0;
        Z[j] = (inside[j] && a.g.dem) ? __ldg(a.g.dem + c + j) : 0.0;
    }
    const double qnan = nan("");
    auto patch = [&](int j, const double2 (&x)[3][Q], double (&val)[NF]) {
#pragma unroll
        for (int ff = 0; ff < NF; ff++) {
            double acc = 0.0;
#pragma unroll
            for (int v = 0; v < 3; v++) {
                acc = fma(w[j][v], x[v][2 * ff].x, acc);
                acc = fma(w[j][3 + 2 * v], x[v][2 * ff].y, acc);
                acc = fma(w[j][4 + 2 * v], x[v][2 * ff + 1].x, acc);
            }
            val[ff] = acc;
        }
    };
    auto combine = [&](int j, const double (&val)[NF], int t) -> double {
        double rr = val[0];
        if (NF == 2) rr = fma(Z[j], val[NF - 1], rr);
        else if (a.pv) rr = __dadd_rn(__dadd_rn(rr, __dmul_rn(__ldg(a.pv + 2 * t), Z[j])), __ldg(a.pv + 2 * t + 1));   // grid.py:210
        return inside[j] ? clamp_nan_keep(rr, a.vmin, a.vmax) : qnan;
    };
    auto emit = [&](const double (&o)[CPT], int t) {
        if (CPT == 2) store_out2(a.out, (size_t)t * ncell + c, o[0], o[CPT - 1], f32);
        else store_out(a.out, (size_t)t * ncell + c, o[0], f32);
    };
    if (!staged) {
        if (!valid) return;
        const size_t rowlen = (size_t)a.T * (4 * NF);
        const double2* vp[CPT][3];
#pragma unroll
        for (int j = 0; j < CPT; j++)
#pragma unroll
            for (int v = 0; v < 3; v++)
                vp[j][v] = reinterpret_cast<const double2*>(a.cvt + (size_t)(inside[j] ? __ldg(a.simplices + 3 * sx[j] + v) : 0) * rowlen);
        for (int t = 0; t < a.T; t++) {
            double o[CPT];
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                double2 x[3][Q];
#pragma unroll
                for (int v = 0; v < 3; v++)
#pragma unroll
                    for (int q = 0; q < Q; q++) x[v][q] = __ldg(vp[j][v] + (size_t)t * Q + q);
                double val[NF];
                patch(j, x, val);
                o[j] = combine(j, val, t);
            }
            emit(o, t);
        }
        return;
    }
    const double2* pr[CPT];
#pragma unroll
    for (int j = 0; j < CPT; j++) pr[j] = vals + (size_t)(3 * r[j]) * RS;
    const bool same = CPT == 2 && r[0] == r[CPT - 1];
    const int nitem = nruns * 3 * GCS_TC * Q;
    for (int t0 = 0; t0 < a.T; t0 += GCS_TC) {
        __syncthreads();          // the previous chunk has been consumed (first pass: vid is complete)
        for (int i = tid; i < nitem; i += GRID_THREADS) {
            const int row = i / (GCS_TC * Q), rem = i - row * (GCS_TC * Q), tc = rem / Q;
            const int v = vid[row];
            vals[row * RS + rem] = (v >= 0 && t0 + tc < a.T)
                                       ? __ldg(reinterpret_cast<const double2*>(a.cvt + ((size_t)v * a.T + t0) * (4 * NF)) + rem)
                                       : make_double2(0.0, 0.0);
        }
        __syncthreads();
        if (valid) {
            const int tn = min(GCS_TC, a.T - t0);
            for (int tc = 0; tc < tn; tc++) {
                double o[CPT];
                double2 x[3][Q];
#pragma unroll
                for (int v = 0; v < 3; v++)
#pragma unroll
                    for (int q = 0; q < Q; q++) x[v][q] = pr[0][v * RS + tc * Q + q];
                double val[NF];
                patch(0, x, val);
                o[0] = combine(0, val, t0 + tc);
                if (CPT == 2) {
                    if (!same) {       // the second cell starts a new run: its own vertex rows
#pragma unroll
                        for (int v = 0; v < 3; v++)
#pragma unroll
                            for (int q = 0; q < Q; q++) x[v][q] = pr[CPT - 1][v * RS + tc * Q + q];
                    }
                    patch(CPT - 1, x, val);
                    o[CPT - 1] = combine(CPT - 1, val, t0 + tc);
                }
                emit(o, t0 + tc);
            }
        }
    }
}

// d_fields: [T][nf][3][S] on the device; d_pv: [T][2] or null
static int grid_cubic_run(GridHandle* h, const double* d_fields, int T, int nf, const double* d_pv, double vmin, double vmax,
                          double* d_out, cudaStream_t st) {
    SMRFB_CHECK(T > 0, SMRFB_ERR_ARG, "T must be > 0");
    SMRFB_CHECK(nf == 1 || nf == 2, SMRFB_ERR_ARG, "cubic: nf must be 1 or 2");
    SMRFB_CHECK(h->d_ctgeo && h->d_ctw, SMRFB_ERR_ARG, "grid method 'cubic' needs smrfb_grid_set_cubic() first");
    SMRFB_CHECK((nf == 1 && !d_pv) || h->g.dem, SMRFB_ERR_ARG, "cubic with a trend needs GridZ");
    const size_t S = (size_t)h->S;
    SMRFB_PROPAGATE(ensure_cap((void**)&h->d_cvt, &h->cvt_cap, S * T * 4 * nf * sizeof(double)));
    dim3 pg((unsigned)((S + 127) / 128), (unsigned)T);
    grid_cubic_pack_kernel<<<pg, 128, 0, st>>>(d_fields, T, nf, (int)S, h->d_cvt);
    count_launch();
    SMRFB_CUDA(cudaGetLastError());
    GridCubicArgs a;
    a.g = h->g;
    a.simplices = h->d_simplices;
    a.transform = h->d_transform;
    a.cell_simplex = h->d_cell_simplex;
    a.ctgeo = h->d_ctgeo;
    a.ctw = h->d_ctw;
    a.cvt = h->d_cvt;
    a.pv = d_pv;
    a.T = T;
    a.vmin = vmin;
    a.vmax = vmax;
    a.out = d_out;
    a.out_f32 = h->out_f32;
    const long long nblk = (h->g.ncell + GRID_THREADS - 1) / GRID_THREADS;
    SMRFB_CHECK(nblk < 2147483647LL, SMRFB_ERR_LIMIT, "too many cells for one launch");
    if (getenv("SMRFB_GRID_CUBIC_DIRECT")) {       // the kernel without staging, for comparison
        if (nf == 1) grid_cubic_kernel<1><<<(unsigned)nblk, GRID_THREADS, 0, st>>>(a);
        else grid_cubic_kernel<2><<<(unsigned)nblk, GRID_THREADS, 0, st>>>(a);
    } else {
        // one cell per thread: two adjacent cells sharing their loads (CPT = 2) need 128 registers with two fields and were
        // measured slower (11.4 / 16.0 ms against 11.0 / 15.4 ms per 256 steps on the C5 grid, mask / local detrend)
        if (nf == 1) grid_cubic_staged_kernel<1, 1><<<(unsigned)nblk, GRID_THREADS, 0, st>>>(a);
        else grid_cubic_staged_kernel<2, 1><<<(unsigned)nblk, GRID_THREADS, 0, st>>>(a);
    }
    count_launch();
    SMRFB_CUDA(cudaGetLastError());
    return SMRFB_OK;
}

static int grid_run(GridHandle* h, const double* d_data, int T, int detrend, int slope_flag, int method, double vmin,
                    double vmax, const double* d_pv_in, double* d_pv_out, double* d_out, cudaStream_t st) {
    SMRFB_CHECK(T > 0, SMRFB_ERR_ARG, "T must be > 0");
    SMRFB_CHECK(method == 0 || method == 1, SMRFB_ERR_ARG, "grid method must be 0 (linear) or 1 (nearest)");
    SMRFB_CHECK(method == 0 || h->d_cell_nearest, SMRFB_ERR_ARG, "grid method 'nearest' needs smrfb_grid_set_nearest() first");
    SMRFB_CHECK(!detrend || (h->g.dem && h->d_mz), SMRFB_ERR_ARG, "detrended GRID needs mz and GridZ");
    const int T_pad = (T + 31) / 32 * 32;
    SMRFB_PROPAGATE(ensure_cap((void**)&h->d_rec, &h->rec_cap, (size_t)T_pad * h->L.RS * sizeof(double)));
    SMRFB_PROPAGATE(ensure_cap((void**)&h->d_vt, &h->vt_cap, (size_t)T_pad * h->S * sizeof(double)));
    SMRFB_PROPAGATE(ensure_cap((void**)&h->d_pvt, &h->pvt_cap, (size_t)T_pad * 2 * sizeof(double)));
    SMRFB_PROPAGATE(launch_prep(st, d_data, T, T_pad, h->S, h->L, h->d_mz, h->d_fitmask, detrend, slope_flag, d_pv_in,
                                d_pv_out, h->d_rec, nullptr, h->d_flags));
    dim3 tb(32, 8), tg((h->S + 31) / 32, T_pad / 32);
    grid_transpose_kernel<<<tg, tb, 0, st>>>(h->d_rec, h->L, T_pad, h->d_vt, h->d_pvt);
    count_launch();
    SMRFB_CUDA(cudaGetLastError());
    GridArgs a;
    a.g = h->g;
    a.simplices = h->d_simplices;
    a.transform = h->d_transform;
    a.cell_simplex = h->d_cell_simplex;
    a.cell_nearest = method == 1 ? h->d_cell_nearest : nullptr;
    a.vt = h->d_vt;
    a.pvt = h->d_pvt;
    a.T = T;
    a.T_pad = T_pad;
    a.detrend = detrend;
    a.vmin = vmin;
    a.vmax = vmax;
    a.out = d_out;
    a.out_f32 = h->out_f32;
    long long nblk = (h->g.ncell + GRID_THREADS - 1) / GRID_THREADS;
    SMRFB_CHECK(nblk < 2147483647LL, SMRFB_ERR_LIMIT, "too many cells for one launch");
    if (method == 0 && (h->g.ncell & 1) == 0 && !getenv("SMRFB_GRID_ONE_CELL")) {
        const long long nb2 = (h->g.ncell / 2 + GRID_THREADS - 1) / GRID_THREADS;
        grid_distribute2_kernel<<<(unsigned)nb2, GRID_THREADS, 0, st>>>(a);
    } else
    grid_distribute_kernel<<<(unsigned)nblk, GRID_THREADS, 0, st>>>(a);
    count_launch();
    SMRFB_CUDA(cudaGetLastError());
    return SMRFB_OK;
}

}  // namespace smrfb

using namespace smrfb;

extern "C" {

int smrfb_grid_create(int npts, const double* mx, const double* my, const double* mz, int ny, int nx, int grid_kind,
                      const double* gx, const double* gy, const double* dem, int nsimp, const int32_t* simplices,
                      const double* transform, const int32_t* cell_simplex, const uint8_t* point_mask, int device,
                      smrfb_handle_t* out) {
    SMRFB_CHECK(out, SMRFB_ERR_ARG, "out handle pointer is NULL");
    *out = nullptr;
    SMRFB_CHECK(nsimp > 0 && simplices && transform && cell_simplex, SMRFB_ERR_ARG, "triangulation arrays are required");
    for (int i = 0; i < 3 * nsimp; i++)
        SMRFB_CHECK(simplices[i] >= 0 && simplices[i] < npts, SMRFB_ERR_ARG, "simplex vertex %d out of range", simplices[i]);
    GridHandle* h = new GridHandle();
    h->device = device;
    int rc = upload_geom(h, npts, mx, my, mz, ny, nx, grid_kind, gx, gy, dem);
    auto up = [&](void** dst, const void* src, size_t bytes) {
        if (rc != SMRFB_OK) return;
        if (cudaMalloc(dst, bytes) != cudaSuccess || cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("GRID: upload of %zu bytes failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
            rc = SMRFB_ERR_CUDA;
        }
    };
    if (rc == SMRFB_OK) {
        const size_t ncell = (size_t)ny * nx;
        for (size_t c = 0; c < ncell && rc == SMRFB_OK; c++)
            if (cell_simplex[c] < -1 || cell_simplex[c] >= nsimp) {
                set_error("cell_simplex[%zu] = %d out of range", c, cell_simplex[c]);
                rc = SMRFB_ERR_ARG;
            }
        h->nsimp = nsimp;
        up((void**)&h->d_simplices, simplices, (size_t)nsimp * 3 * sizeof(int32_t));
        up((void**)&h->d_transform, transform, (size_t)nsimp * 6 * sizeof(double));
        up((void**)&h->d_cell_simplex, cell_simplex, ncell * sizeof(int32_t));
        if (point_mask) up((void**)&h->d_fitmask, point_mask, (size_t)npts);
        h->L = make_layout(npts, 2, 0, false);
        h->L.keep_nan = 1;
    }
    if (rc != SMRFB_OK) {
        delete h;
        return rc;
    }
    *out = reinterpret_cast<smrfb_handle_t>(h);
    return SMRFB_OK;
}

int smrfb_grid_set_nearest(smrfb_handle_t hh, const int32_t* cell_nearest) {
    SMRFB_CHECK(hh && cell_nearest, SMRFB_ERR_ARG, "NULL handle / array");
    GridHandle* h = reinterpret_cast<GridHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_GRID, SMRFB_ERR_ARG, "not a GRID handle");
    SMRFB_CUDA(cudaSetDevice(h->device));
    const size_t ncell = (size_t)h->g.ncell;
    for (size_t c = 0; c < ncell; c++)
        SMRFB_CHECK(cell_nearest[c] >= 0 && cell_nearest[c] < h->S, SMRFB_ERR_ARG, "cell_nearest[%zu] = %d out of range", c, cell_nearest[c]);
    if (!h->d_cell_nearest) SMRFB_CUDA(cudaMalloc((void**)&h->d_cell_nearest, ncell * sizeof(int32_t)));
    SMRFB_CUDA(cudaMemcpy(h->d_cell_nearest, cell_nearest, ncell * sizeof(int32_t), cudaMemcpyHostToDevice));
    return SMRFB_OK;
}

int smrfb_grid_set_cubic(smrfb_handle_t hh, const double* ct_geom) {
    SMRFB_CHECK(hh && ct_geom, SMRFB_ERR_ARG, "NULL handle / array");
    GridHandle* h = reinterpret_cast<GridHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_GRID, SMRFB_ERR_ARG, "not a GRID handle");
    SMRFB_CUDA(cudaSetDevice(h->device));
    if (!h->d_ctgeo) SMRFB_CUDA(cudaMalloc((void**)&h->d_ctgeo, (size_t)h->nsimp * 9 * sizeof(double)));
    SMRFB_CUDA(cudaMemcpy(h->d_ctgeo, ct_geom, (size_t)h->nsimp * 9 * sizeof(double), cudaMemcpyHostToDevice));
    if (!h->d_ctw) SMRFB_CUDA(cudaMalloc((void**)&h->d_ctw, (size_t)h->g.ncell * 9 * sizeof(double)));
    GridCubicArgs a = {};
    a.g = h->g;
    a.simplices = h->d_simplices;
    a.transform = h->d_transform;
    a.cell_simplex = h->d_cell_simplex;
    a.ctgeo = h->d_ctgeo;
    const long long nblk = (h->g.ncell + GRID_THREADS - 1) / GRID_THREADS;
    SMRFB_CHECK(nblk < 2147483647LL, SMRFB_ERR_LIMIT, "too many cells for one launch");
    grid_cubic_weights_kernel<<<(unsigned)nblk, GRID_THREADS, 0, h->stream>>>(a, h->d_ctw);
    count_launch();
    SMRFB_CUDA(cudaGetLastError());
    SMRFB_CUDA(cudaStreamSynchronize(h->stream));
    return SMRFB_OK;
}

int smrfb_grid_distribute_cubic_dev(smrfb_handle_t hh, const double* d_fields, int T, int nf, const double* d_pv, double vmin,
                                    double vmax, double* d_out, void* stream) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    GridHandle* h = reinterpret_cast<GridHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_GRID, SMRFB_ERR_ARG, "not a GRID handle");
    SMRFB_CHECK(d_fields && d_out, SMRFB_ERR_ARG, "NULL fields/out");
    SMRFB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    return grid_cubic_run(h, d_fields, T, nf, d_pv, vmin, vmax, d_out, st);
}

int smrfb_grid_distribute_cubic(smrfb_handle_t hh, const double* fields, int T, int nf, const double* pv, double vmin,
                                double vmax, double* out) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    GridHandle* h = reinterpret_cast<GridHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_GRID, SMRFB_ERR_ARG, "not a GRID handle");
    SMRFB_CHECK(fields && out && T > 0 && (nf == 1 || nf == 2), SMRFB_ERR_ARG, "NULL fields/out, T <= 0 or nf not 1 / 2");
    SMRFB_CUDA(cudaSetDevice(h->device));
    const size_t S = (size_t)h->S, per = (size_t)nf * 3 * S;
    SMRFB_PROPAGATE(ensure_cap((void**)&h->d_cf, &h->cf_cap, ((size_t)T * per + 2 * (size_t)T) * sizeof(double)));
    SMRFB_CUDA(cudaMemcpyAsync(h->d_cf, fields, (size_t)T * per * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    double* d_pv = nullptr;
    if (pv) {
        d_pv = h->d_cf + (size_t)T * per;
        SMRFB_CUDA(cudaMemcpyAsync(d_pv, pv, (size_t)T * 2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    return run_host_batches(h, T, 32, out, [&](int t0, int tn, double* d_out, cudaStream_t st) {
        return grid_cubic_run(h, h->d_cf + (size_t)t0 * per, tn, nf, d_pv ? d_pv + 2 * (size_t)t0 : nullptr, vmin, vmax, d_out, st);
    });
}

int smrfb_grid_distribute_dev(smrfb_handle_t hh, const double* d_data, int T, int detrend, int slope_flag, int method,
                              double vmin, double vmax, const double* d_pv_in, double* d_pv_out, double* d_out,
                              void* stream) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    GridHandle* h = reinterpret_cast<GridHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_GRID, SMRFB_ERR_ARG, "not a GRID handle");
    SMRFB_CHECK(d_data && d_out, SMRFB_ERR_ARG, "NULL data/out");
    SMRFB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    return grid_run(h, d_data, T, detrend, slope_flag, method, vmin, vmax, d_pv_in, d_pv_out, d_out, st);
}

int smrfb_grid_set_local_exact(smrfb_handle_t hh, int exact) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    GridHandle* h = reinterpret_cast<GridHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_GRID, SMRFB_ERR_ARG, "not a GRID handle");
    h->local_exact = exact ? 1 : 0;
    return SMRFB_OK;
}

int smrfb_grid_distribute_local_dev(smrfb_handle_t hh, const double* d_fields3, int T, double vmin, double vmax, double* d_out,
                                    void* stream) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    GridHandle* h = reinterpret_cast<GridHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_GRID, SMRFB_ERR_ARG, "not a GRID handle");
    SMRFB_CHECK(d_fields3 && d_out, SMRFB_ERR_ARG, "NULL fields/out");
    SMRFB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    return grid_local_run(h, d_fields3, T, vmin, vmax, d_out, st);
}

int smrfb_grid_distribute_local(smrfb_handle_t hh, const double* fields3, int T, double vmin, double vmax, double* out) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    GridHandle* h = reinterpret_cast<GridHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_GRID, SMRFB_ERR_ARG, "not a GRID handle");
    SMRFB_CHECK(fields3 && out && T > 0, SMRFB_ERR_ARG, "NULL fields/out or T <= 0");
    SMRFB_CUDA(cudaSetDevice(h->device));
    const size_t S = (size_t)h->S;
    SMRFB_PROPAGATE(ensure_cap((void**)&h->d_l3, &h->l3_cap, (size_t)T * 3 * S * sizeof(double)));
    SMRFB_CUDA(cudaMemcpyAsync(h->d_l3, fields3, (size_t)T * 3 * S * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return run_host_batches(h, T, 32, out, [&](int t0, int tn, double* d_out, cudaStream_t st) {
        return grid_local_run(h, h->d_l3 + (size_t)t0 * 3 * S, tn, vmin, vmax, d_out, st);
    });
}

int smrfb_grid_distribute(smrfb_handle_t hh, const double* data, int T, int detrend, int slope_flag, int method,
                          double vmin, double vmax, const double* pv_in, double* pv_out, double* out) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    GridHandle* h = reinterpret_cast<GridHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_GRID, SMRFB_ERR_ARG, "not a GRID handle");
    SMRFB_CHECK(data && out && T > 0, SMRFB_ERR_ARG, "NULL data/out or T <= 0");
    SMRFB_CUDA(cudaSetDevice(h->device));
    const size_t S = (size_t)h->S;
    SMRFB_PROPAGATE(ensure_cap((void**)&h->d_data, &h->data_cap, (size_t)T * S * sizeof(double)));
    SMRFB_PROPAGATE(ensure_cap((void**)&h->d_pv, &h->pv_cap, (size_t)T * 4 * sizeof(double)));
    SMRFB_CUDA(cudaMemcpyAsync(h->d_data, data, (size_t)T * S * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    double* d_pvin = nullptr;
    if (pv_in) {
        d_pvin = h->d_pv + (size_t)2 * T;
        SMRFB_CUDA(cudaMemcpyAsync(d_pvin, pv_in, (size_t)T * 2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    int rc = run_host_batches(h, T, 32, out, [&](int t0, int tn, double* d_out, cudaStream_t st) {
        return grid_run(h, h->d_data + (size_t)t0 * S, tn, detrend, slope_flag, method, vmin, vmax,
                        d_pvin ? d_pvin + (size_t)2 * t0 : nullptr, h->d_pv + (size_t)2 * t0, d_out, st);
    });
    if (rc == SMRFB_OK && pv_out &&
        cudaMemcpy(pv_out, h->d_pv, (size_t)T * 2 * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) {
        set_error("GRID: reading back coefficients failed");
        rc = SMRFB_ERR_CUDA;
    }
    return rc;
}

}  // extern "C"
