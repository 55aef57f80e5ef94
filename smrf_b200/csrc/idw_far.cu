// idw_far.cu -- IDW (smrf/spatial/idw.py:53-110) on a uniform mesh as near field + interpolated far field.
//
// The reference evaluates  v(c, t) = sum_s w(c, s) r(s, t) / sum_s w(c, s) valid(s, t),  w = 1 / dist(c, s)^power, over ALL
// stations for every cell: 2 S flop per output and term (400 + 400 at 200 stations), which puts the fp64 pipe, not HBM,
// in charge.  For the stations that are far from a patch of cells, w(., s) is an analytic function of the
// cell position with its singularity at the station, so over a 128 x 128-cell patch it is reproduced to ~1e-11 by its
// tensor-product Chebyshev interpolant on nn x nn nodes once the station is further than theta patch half widths from
// the patch (error per station <= eps(theta, nn, power) RELATIVE to its own weight, all weights positive, hence
// |error of v| <= 2 eps max|r|; eps(6, 10, 2) = 1.0e-10, eps(8, 12, 4) = 1.1e-12, measured over all positions on the
// boundary of the near region).  Interpolation is linear, so the far field of all far stations together is
//     F(c, t) = sum_ky sum_kx Ly[row][ky] Lx[col][kx] G[ky][kx][t],     G[node][t] = sum_{s far} w(node, s) r(s, t)
// (and the same with validity instead of residuals for the masked denominator -- no missing-station lists, no
// cancellation).  Per launch:
//   idw_far_pack_kernel   residual / validity pairs per (step, station) and their station-major transpose;
//   idw_far_node_kernel   G for every patch the row window touches: a [nn^2 x S] x [S x 2T] product per patch, the node
//                         weights generated in shared memory (near stations get weight 0 there);
//   idw_far_kernel        one CTA per 8 rows x 128 columns: the patch's G streams through shared memory in 16-step chunks
//                         (1-D TMA bulk copies, double buffered); each warp owns one row: it contracts the chunk with its
//                         row's Ly (phase A), then every lane finishes 4 cells per step (phase B): the column basis is
//                         folded into even / odd parts so that the mirror cells c and 127 - c share their products
//                         (nn FMA per pair and term instead of 2 nn), adds the patch's near stations directly, divides,
//                         retrends against the DEM, clamps and streams the row segment out with 256-byte stores.
// Per output: ~12 FMA of interpolation + 2 per near station (1.5 on average at 200 stations over 300 km) + the epilogue,
// against 800: the kernel is bound by the fp64 pipe at ~28 instructions per output.
#include "idw_far.cuh"

#include <array>
#include <type_traits>

namespace smrfb {

constexpr int FAR_RB = 8;            // mesh rows per CTA (one warp each)
constexpr int FAR_TC = 8;            // timesteps per staged chunk
constexpr int FAR_CW = 2 * FAR_TC;   // columns per chunk: numerator steps, then denominator steps
constexpr int FAR_THREADS = FAR_RB * 32;
constexpr int FAR_NQR = 2;           // near stations whose cell weights stay in registers; the rest go through global scratch
constexpr int FAR_NQM = 2;           // further near stations whose cell weights sit in shared memory (default; FarArgs::nqm, as many as fit)
constexpr int FAR_NQS = 7;           // near stations whose (residual, validity) pairs are staged in shared memory per chunk
constexpr int FAR_NODE_THREADS = 256;
constexpr int FAR_NODE_COLS = 128;   // (step, term) columns per node-kernel tile
constexpr int FAR_NODE_KB = 8;       // stations per staged block of the node kernel

void IdwFarPlan::release() {
    for (cudaEvent_t e : tev) cudaEventDestroy(e);
    tev.clear();
    cudaFree(d_near_off);
    cudaFree(d_near_idx);
    cudaFree(d_tab);
    cudaFree(d_G);
    cudaFree(d_rv);
    cudaFree(d_wn);
    cudaFree(d_coin);
    d_coin = nullptr;
    ncoin = 0;
    d_near_off = d_near_idx = nullptr;
    d_tab = d_G = d_rv = d_wn = nullptr;
    G_cap = rv_cap = wn_cap = 0;
    enabled = false;
}

struct FarXi {
    double xi[12];
};

// ------------------------------------------------------------------------------------------------------------------
// rv[t][s] = (residual, validity) and rvt[s][col], col = (t / TC) * 2 TC + term * TC + t % TC
__global__ void idw_far_pack_kernel(const double* __restrict__ rec, int RS, const uint8_t* __restrict__ valid, int T, int S,
                                    int N, double2* __restrict__ rv, double* __restrict__ rvt) {
    const int t = blockIdx.x;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        double r = 0.0, v = 0.0;
        if (t < T) {
            r = rec[(size_t)t * RS + s];
            v = valid[(size_t)t * S + s] ? 1.0 : 0.0;
        }
        rv[(size_t)t * S + s] = make_double2(r, v);
        const size_t col = (size_t)(t / FAR_TC) * (2 * FAR_TC) + (t % FAR_TC);
        rvt[(size_t)s * N + col] = r;
        rvt[(size_t)s * N + col + FAR_TC] = v;
    }
}

// ------------------------------------------------------------------------------------------------------------------
struct FarNodeArgs {
    const int* near_off;
    const int* near_idx;
    const double* mx;
    const double* my;
    const double* rvt;   // [S][N]
    double* G;           // [patch][chunk][node][FAR_CW]
    double x0, dx, y0, dy, power;
    int S, N, npx, pi0, SB;
    FarXi xi;
};

template <int NN>
__global__ void __launch_bounds__(FAR_NODE_THREADS, 1) idw_far_node_kernel(FarNodeArgs a) {
    constexpr int NNODE = NN * NN;
    constexpr int NPT = NNODE / 4;      // nodes per thread
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Wn = reinterpret_cast<double*>(smem_raw);                          // [SB][NNODE]
    double* Bs = Wn + (size_t)a.SB * NNODE;                                     // [2][8][128] staged (residual | validity) columns
    uint8_t* isnear = reinterpret_cast<uint8_t*>(Bs + 2 * FAR_NODE_KB * FAR_NODE_COLS);   // [S]
    const int tid = threadIdx.x;
    const int lp = blockIdx.x;
    const int pj = lp % a.npx, pi = a.pi0 + lp / a.npx;
    const int gp = pi * a.npx + pj;
    for (int s = tid; s < a.S; s += FAR_NODE_THREADS) isnear[s] = 0;
    __syncthreads();
    for (int i = a.near_off[gp] + tid; i < a.near_off[gp + 1]; i += FAR_NODE_THREADS) isnear[a.near_idx[i]] = 1;
    const double hx = 0.5 * (FAR_P - 1) * a.dx, hy = 0.5 * (FAR_P - 1) * a.dy;
    const double xc = a.x0 + ((double)pj * FAR_P + 0.5 * (FAR_P - 1)) * a.dx;
    const double yc = a.y0 + ((double)pi * FAR_P + 0.5 * (FAR_P - 1)) * a.dy;
    const bool p2 = (a.power == 2.0);
    const int ng = tid >> 6, cp = tid & 63;
    const int nchunk = a.N / (2 * FAR_TC);
    double* Gp = a.G + (size_t)lp * nchunk * NNODE * (2 * FAR_TC);

    for (int sb0 = 0; sb0 < a.S; sb0 += a.SB) {
        const int sbn = min(a.SB, a.S - sb0);
        __syncthreads();
        for (int i = tid; i < sbn * NNODE; i += FAR_NODE_THREADS) {
            const int sl = i / NNODE, nd = i - sl * NNODE;
            const int s = sb0 + sl;
            double w = 0.0;
            if (!isnear[s]) {
                const double X = xc + hx * a.xi.xi[nd % NN], Y = yc + hy * a.xi.xi[nd / NN];
                const double ddx = X - a.mx[s], ddy = Y - a.my[s];
                const double q = ddx * ddx + ddy * ddy;
                w = p2 ? 1.0 / q : 1.0 / pow(sqrt(q), a.power);
            }
            Wn[i] = w;
        }
        __syncthreads();
        for (int c0 = 0; c0 < a.N; c0 += FAR_NODE_COLS) {
            const int ca = c0 + cp, cb = c0 + 64 + cp;
            const bool la = ca < a.N, lb = cb < a.N;
            double acc0[NPT], acc1[NPT];
#pragma unroll
            for (int j = 0; j < NPT; j++) acc0[j] = acc1[j] = 0.0;
            const double* wrow = Wn + ng * NPT;
            // the (residual | validity) tile [8 stations][128 columns] goes through a double-buffered shared stage: the loads
            // of block i + 1 are in flight while block i is consumed
            double pre[4];
            auto fetch = [&](int s0) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int e = tid + FAR_NODE_THREADS * i, sr = e >> 7, col = c0 + (e & 127);
                    pre[i] = (s0 + sr < sbn && col < a.N) ? __ldg(a.rvt + (size_t)(sb0 + s0 + sr) * a.N + col) : 0.0;
                }
            };
            fetch(0);
            int buf = 0;
            for (int s0 = 0; s0 < sbn; s0 += FAR_NODE_KB, buf ^= 1) {
                double* bs = Bs + buf * FAR_NODE_KB * FAR_NODE_COLS;
#pragma unroll
                for (int i = 0; i < 4; i++) bs[tid + FAR_NODE_THREADS * i] = pre[i];
                __syncthreads();
                if (s0 + FAR_NODE_KB < sbn) fetch(s0 + FAR_NODE_KB);
                const int kn = min(FAR_NODE_KB, sbn - s0);
                for (int sr = 0; sr < kn; sr++) {
                    const double b0 = bs[sr * FAR_NODE_COLS + cp], b1 = bs[sr * FAR_NODE_COLS + 64 + cp];
                    const double* wr = wrow + (size_t)(s0 + sr) * NNODE;
#pragma unroll
                    for (int j = 0; j < NPT; j++) {
                        const double w = wr[j];
                        acc0[j] = fma(w, b0, acc0[j]);
                        acc1[j] = fma(w, b1, acc1[j]);
                    }
                }
            }
            __syncthreads();   // the stage is rewritten by the next tile's first block
#pragma unroll
            for (int j = 0; j < NPT; j++) {
                const int node = ng * NPT + j;
                if (la) {
                    double* p = Gp + ((size_t)(ca / FAR_CW) * NNODE + node) * FAR_CW + (ca % FAR_CW);
                    *p = sb0 ? *p + acc0[j] : acc0[j];
                }
                if (lb) {
                    double* p = Gp + ((size_t)(cb / FAR_CW) * NNODE + node) * FAR_CW + (cb % FAR_CW);
                    *p = sb0 ? *p + acc1[j] : acc1[j];
                }
            }
        }
    }
}

// The same product on the fp64 tensor cores (default): G[node][col] = sum_s Wn[s][node] B[s][col] as DMMA.8x8x4 tiles.  16 warps;
// a warp owns 8 of the tile's 128 columns and ALL node tiles (13 or 18 accumulator pairs in registers), so the B fragment of a
// k-step is loaded once and every A fragment feeds one DMMA; the node stride NP is = 8 (mod 16) doubles, which makes an A
// fragment load two wavefronts.  The DFMA kernel above (8 warps per SM, one shared load per two multiply-adds)
// reached 41 % of the fp64 peak.
constexpr int FAR_NODE_MMA_THREADS = 512;
template <int NN>
struct FarNodeMma {
    static constexpr int NNODE = NN * NN;
    static constexpr int MT = (NNODE + 7) / 8;                               // node tiles
    static constexpr int NP = MT * 8 + ((MT * 8) % 16 == 8 ? 0 : 8);         // padded node stride: 104 (NN = 10), 152 (NN = 12)
};

template <int NN>
__global__ void __launch_bounds__(FAR_NODE_MMA_THREADS, 1) idw_far_node_mma_kernel(FarNodeArgs a) {
    constexpr int NNODE = FarNodeMma<NN>::NNODE, MT = FarNodeMma<NN>::MT, NP = FarNodeMma<NN>::NP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Wn = reinterpret_cast<double*>(smem_raw);                          // [SB][NP], SB a multiple of 8
    uint8_t* isnear = reinterpret_cast<uint8_t*>(Wn + (size_t)a.SB * NP);      // [S]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int lp = blockIdx.x;
    const int pj = lp % a.npx, pi = a.pi0 + lp / a.npx;
    const int gp = pi * a.npx + pj;
    for (int s = tid; s < a.S; s += FAR_NODE_MMA_THREADS) isnear[s] = 0;
    __syncthreads();
    for (int i = a.near_off[gp] + tid; i < a.near_off[gp + 1]; i += FAR_NODE_MMA_THREADS) isnear[a.near_idx[i]] = 1;
    const double hx = 0.5 * (FAR_P - 1) * a.dx, hy = 0.5 * (FAR_P - 1) * a.dy;
    const double xc = a.x0 + ((double)pj * FAR_P + 0.5 * (FAR_P - 1)) * a.dx;
    const double yc = a.y0 + ((double)pi * FAR_P + 0.5 * (FAR_P - 1)) * a.dy;
    const bool p2 = (a.power == 2.0);
    const int nchunk = a.N / (2 * FAR_TC);
    double* Gp = a.G + (size_t)lp * nchunk * NNODE * (2 * FAR_TC);
    const int fk = lane & 3, fr = lane >> 2;          // fragment coordinates: k / row of A, k / column of B

    for (int sb0 = 0; sb0 < a.S; sb0 += a.SB) {
        const int sbn = min(a.SB, a.S - sb0), sbn8 = (sbn + 7) & ~7;
        __syncthreads();
        for (int i = tid; i < sbn8 * NP; i += FAR_NODE_MMA_THREADS) {      // padded nodes and padded stations get weight 0
            const int sl = i / NP, nd = i - sl * NP;
            const int s = sb0 + sl;
            double w = 0.0;
            if (sl < sbn && nd < NNODE && !isnear[s]) {
                const double X = xc + hx * a.xi.xi[nd % NN], Y = yc + hy * a.xi.xi[nd / NN];
                const double ddx = X - a.mx[s], ddy = Y - a.my[s];
                const double q = ddx * ddx + ddy * ddy;
                w = p2 ? 1.0 / q : 1.0 / pow(sqrt(q), a.power);
            }
            Wn[i] = w;
        }
        __syncthreads();
        for (int c0 = 0; c0 < a.N; c0 += FAR_NODE_COLS) {
            double acc[MT][2];
#pragma unroll
            for (int m = 0; m < MT; m++) acc[m][0] = acc[m][1] = 0.0;
            // A B element is used by exactly one warp of the CTA, so the B fragments come straight from global memory (the
            // [S][N] table is L2-resident, a fragment load is four 64-byte runs) through a register queue PF k-steps deep:
            // no staging, no block barrier inside a tile
            constexpr int PF = 4;
            const int bcol = c0 + warp * 8 + fr;
            const double* bp = a.rvt + (size_t)(sb0 + fk) * a.N + bcol;
            auto bload = [&](int k4) -> double {
                return (k4 * 4 + fk < sbn && bcol < a.N) ? __ldg(bp + (size_t)k4 * 4 * a.N) : 0.0;
            };
            double bq[PF];
#pragma unroll
            for (int i = 0; i < PF; i++) bq[i] = bload(i);
            const int nk = sbn8 / 4;
            for (int k0 = 0; k0 < nk; k0 += PF) {
#pragma unroll
                for (int i = 0; i < PF; i++) {
                    const int k4 = k0 + i;
                    if (k4 < nk) {           // uniform
                        const double b = bq[i];
                        bq[i] = bload(k4 + PF);
                        const double* wa = Wn + (size_t)(k4 * 4 + fk) * NP + fr;
#pragma unroll
                        for (int m = 0; m < MT; m++) dmma884(acc[m][0], acc[m][1], wa[m * 8], b);
                    }
                }
            }
            const int col = c0 + warp * 8 + 2 * fk;
            if (col < a.N) {   // N is a multiple of 16: col + 1 exists too
#pragma unroll
                for (int m = 0; m < MT; m++) {
                    const int node = m * 8 + fr;
                    if (node < NNODE) {
                        double2* p = reinterpret_cast<double2*>(Gp + ((size_t)(col / FAR_CW) * NNODE + node) * FAR_CW + (col % FAR_CW));
                        double2 v = make_double2(acc[m][0], acc[m][1]);
                        if (sb0) {
                            const double2 o = *p;
                            v.x += o.x;
                            v.y += o.y;
                        }
                        *p = v;
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
struct FarArgs {
    Geom g;
    const double* mx;
    const double* my;
    const int* near_off;
    const int* near_idx;
    const double* tab;      // [128][NN] basis, then [64][NN] folded halves
    const double* G;
    const double2* rv;      // [T_pad][S]
    const double* rec;
    double* wn;             // near-weight scratch
    double* out;
    double power, vmin, vmax;
    long long ostride;      // bytes between consecutive steps of the output window
    unsigned klo, kspan;    // clamp screen on the high word (see hi_key)
    int S, RS, S_pad, T, nchunk, clamp, out_f32;
    int row0, nrows, npx, pi0, qx, nqm;
};

// A valid station exactly on a cell centre: the reference's weight there is inf and the cell becomes NaN (inf / inf), or 0
// when that station's residual is exactly 0 (nansum drops inf * 0) -- idw.py:69-94.  The main kernel gives such a station
// weight 0 on that cell (right whenever the station is missing); this fix-up pass, over the at most S such cells, rewrites
// the steps at which the station reports.  One thread per (pair, step); pairs are sorted by cell.
struct FarCoinArgs {
    const int* crow;
    const int* ccol;
    const int* csta;
    const double2* rv;
    const double* rec;
    const double* dem;
    void* out;
    double vmin, vmax;
    int npair, S, RS, S_pad, nx, row0, nrows, T, clamp, out_f32;
};

__global__ void idw_far_coin_kernel(FarCoinArgs a) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int p = (int)(idx / a.T), t = (int)(idx % a.T);
    if (p >= a.npair) return;
    const int row = a.crow[p], col = a.ccol[p];
    if (p > 0 && a.crow[p - 1] == row && a.ccol[p - 1] == col) return;   // the first pair of a cell speaks for all of them
    if (row < a.row0 || row >= a.row0 + a.nrows) return;
    bool hit = false, bad = false;
    for (int q = p; q < a.npair && a.crow[q] == row && a.ccol[q] == col; q++) {
        const double2 x = a.rv[(size_t)t * a.S + a.csta[q]];
        if (x.y != 0.0) {
            hit = true;
            if (x.x != 0.0) bad = true;
        }
    }
    if (!hit) return;
    const double2 pp = *reinterpret_cast<const double2*>(a.rec + (size_t)t * a.RS + a.S_pad);
    const double Z = a.dem ? a.dem[(size_t)row * a.nx + col] : 0.0;
    double val = (bad ? nan("") : 0.0) + fma(pp.x, Z, pp.y);
    if (a.clamp) val = clamp_nan_keep(val, a.vmin, a.vmax);
    const size_t o = ((size_t)t * a.nrows + (row - a.row0)) * a.nx + col;
    if (a.out_f32) reinterpret_cast<float*>(a.out)[o] = (float)val;
    else reinterpret_cast<double*>(a.out)[o] = val;
}

// Phase A of idw_far_kernel for KN node-column pairs k0 .. k0 + KN - 1: g1 = sum_ky ly[ky] G[ky][k], g2 the same for the mirror
// column NN - 1 - k; stores E = g1 + g2, O = g1 - g2.
template <int NN, int KN>
__device__ __forceinline__ void far_phase_a(const double* __restrict__ gsl, const double* __restrict__ lyp, double* __restrict__ eo,
                                            int k0) {
    const double* ga = gsl + k0 * FAR_CW;
    const double* gb = gsl + (NN - 1 - k0) * FAR_CW;
    double g1[KN], g2[KN];
#pragma unroll
    for (int kk = 0; kk < KN; kk++) g1[kk] = g2[kk] = 0.0;
#pragma unroll
    for (int ky = 0; ky < NN; ky++) {
        const double ly = lyp[ky];
#pragma unroll
        for (int kk = 0; kk < KN; kk++) {
            g1[kk] = fma(ly, ga[ky * NN * FAR_CW + kk * FAR_CW], g1[kk]);
            g2[kk] = fma(ly, gb[ky * NN * FAR_CW - kk * FAR_CW], g2[kk]);
        }
    }
#pragma unroll
    for (int kk = 0; kk < KN; kk++) {
        eo[(k0 + kk) * 4] = g1[kk] + g2[kk];
        eo[(k0 + kk) * 4 + 1] = g1[kk] - g2[kk];
    }
}

// Monotone map of a double's high word to unsigned: key(a) < key(b) implies a < b for non-NaN a, b; NaNs map beyond +-inf.
__host__ __device__ __forceinline__ unsigned hi_key(int hi) { return (unsigned)hi ^ ((unsigned)(hi >> 31) | 0x80000000u); }

// CLAMP: 0 none; otherwise the screen that proves a value strictly inside (vmin, vmax) from its high word alone:
//   1 general (monotone key of the high word: 3 integer instructions per value), 2 vmin >= 0 (the high word itself is monotone
//   there; a negative value or a NaN falls outside as an unsigned number), 3 vmin < 0 < vmax (|value| below the smaller bound)
template <int NN, typename OutT, int CLAMP>
__global__ void __launch_bounds__(FAR_THREADS, 2) idw_far_kernel(FarArgs a) {
    constexpr int NNODE = NN * NN;
    constexpr int NH = NN / 2;
    constexpr int EOS = NH * 4 + 2;      // doubles per (row, step) in EO: NH x {E_num, O_num, E_den, O_den}, padded
    constexpr int SGS = 1 + FAR_NQS;     // double2 per step in the stage: (p0, p1), then (residual, validity) of the near stations
    constexpr uint32_t CHUNK_BYTES = NNODE * 2 * FAR_TC * sizeof(double);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Gs = reinterpret_cast<double*>(smem_raw);                    // [2][NNODE][FAR_CW]
    double* EO = Gs + 2 * NNODE * 2 * FAR_TC;                            // [2][RB][TC][EOS]
    double2* stg = reinterpret_cast<double2*>(EO + 2 * FAR_RB * FAR_TC * EOS);   // [2][TC][SGS]
    double* Lys = reinterpret_cast<double*>(stg + 2 * FAR_TC * SGS);     // [RB][NN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(Lys + FAR_RB * NN);     // [2]
    double* wns = reinterpret_cast<double*>(bars + 2);                   // [NQM][RB][128] near weights 2, 3 of the cells
    int* nearl = reinterpret_cast<int*>(wns + a.nqm * FAR_RB * FAR_P);   // [nq]

    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    // grid: x = 8-row band within the 128-row patch (the 16 CTAs that share a patch's node sums launch together and
    // meet in L2), y = patch column, z = patch row of the window
    const int pj = blockIdx.y;
    const int pi = a.pi0 + blockIdx.z;
    const int grow0 = pi * FAR_P + blockIdx.x * FAR_RB;
    if (grow0 + FAR_RB <= a.row0 || grow0 >= a.row0 + a.nrows) return;   // band outside the row window
    const int gp = pi * a.npx + pj;
    const int lp = blockIdx.z * a.npx + pj;
    const int grow = grow0 + w;
    const bool rowlive = grow >= a.row0 && grow < a.row0 + a.nrows;
    const int nq = a.near_off[gp + 1] - a.near_off[gp];
    const int nqs = min(nq, FAR_NQS);
    const double* Gp = a.G + (size_t)lp * a.nchunk * NNODE * (2 * FAR_TC);

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    for (int i = tid; i < nq; i += FAR_THREADS) nearl[i] = a.near_idx[a.near_off[gp] + i];
    for (int i = tid; i < FAR_RB * NN; i += FAR_THREADS) Lys[i] = a.tab[(blockIdx.x * FAR_RB + i / NN) * NN + i % NN];
    __syncthreads();
    if (tid == 0) {
        for (int b = 0; b < 2 && b < a.nchunk; b++) {
            mbar_arrive_expect_tx(&bars[b], CHUNK_BYTES);
            bulk_g2s(Gs + b * NNODE * 2 * FAR_TC, Gp + (size_t)b * NNODE * 2 * FAR_TC, CHUNK_BYTES, &bars[b]);
        }
    }

    // the lane's four cells: columns lane, 127 - lane (pair A) and 32 + lane, 95 - lane (pair B) of the patch
    const int cb = pj * FAR_P;
    const int coff[4] = {lane, FAR_P - 1 - lane, 32 + lane, FAR_P - 33 - lane};
    bool live[4];
    double Z[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        live[j] = rowlive && cb + coff[j] < a.g.nx;
        Z[j] = (live[j] && a.g.dem) ? __ldg(a.g.dem + (size_t)grow * a.g.nx + cb + coff[j]) : 0.0;
    }
    double cA[NN], cB[NN];   // folded column basis: [0, NH) even part, [NH, NN) odd part
    {
        const double* f = a.tab + FAR_P * NN;
#pragma unroll
        for (int k = 0; k < NN; k++) {
            cA[k] = __ldg(f + lane * NN + k);
            cB[k] = __ldg(f + (32 + lane) * NN + k);
        }
    }
    // near-station weights of the four cells
    double wr[FAR_NQR][4];
    const size_t blk = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    double* wn = a.wn + blk * (size_t)a.qx * (FAR_RB * FAR_P) + (size_t)w * FAR_P + lane;
    {
        const bool p2 = (a.power == 2.0);
        const double Y = rowlive ? __ldg(a.g.gy + grow) : 0.0;
        double X[4];
#pragma unroll
        for (int j = 0; j < 4; j++) X[j] = live[j] ? __ldg(a.g.gx + cb + coff[j]) : 0.0;
#pragma unroll
        for (int qi = 0; qi < FAR_NQR; qi++)
#pragma unroll
            for (int j = 0; j < 4; j++) wr[qi][j] = 0.0;
        for (int qi = 0; qi < nq; qi++) {
            const int s = nearl[qi];
            const double sx = __ldg(a.mx + s), sy = __ldg(a.my + s);
            double ws[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const bool lv = live[j];
                const double ddx = __dsub_rn(X[j], sx), ddy = __dsub_rn(Y, sy);
                const double q = __dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy));
                const bool zero = (q == 0.0);
                double v = p2 ? fast_rcp<2>(zero ? 1.0 : q) : 1.0 / pow(sqrt(zero ? 1.0 : q), a.power);   // idw.py:77
                ws[j] = (lv && !zero) ? v : 0.0;
            }
            if (qi == 0) {
#pragma unroll
                for (int j = 0; j < 4; j++) wr[0][j] = ws[j];
            } else if (qi == 1) {
#pragma unroll
                for (int j = 0; j < 4; j++) wr[1][j] = ws[j];
            } else if (qi < FAR_NQR + a.nqm) {
#pragma unroll
                for (int j = 0; j < 4; j++) wns[(qi - FAR_NQR) * (FAR_RB * FAR_P) + w * FAR_P + lane + 32 * j] = ws[j];
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) wn[(size_t)(qi - FAR_NQR - a.nqm) * (FAR_RB * FAR_P) + 32 * j] = ws[j];
            }
        }
    }
    const bool full = rowlive && cb + FAR_P <= a.g.nx;     // all four cells of every lane exist
    OutT* orow = reinterpret_cast<OutT*>(a.out) + (size_t)(rowlive ? grow - a.row0 : 0) * a.g.nx + cb;
    // the stage item this thread fetches per chunk: step tid / SGS, slot tid % SGS
    const int sg_t = tid / SGS, sg_s = tid - sg_t * SGS;
    const bool sg_on = sg_t < FAR_TC && sg_s <= nqs;
    const int sg_sta = (sg_on && sg_s > 0) ? nearl[sg_s - 1] : 0;

    for (int ch = 0; ch < a.nchunk; ch++) {
        const int buf = ch & 1;
        const double* gs = Gs + buf * NNODE * 2 * FAR_TC;
        double2 sgv = make_double2(0.0, 0.0);
        if (sg_on) {
            const int t = ch * FAR_TC + sg_t;
            sgv = sg_s ? __ldg(a.rv + (size_t)t * a.S + sg_sta)
                       : *reinterpret_cast<const double2*>(a.rec + (size_t)t * a.RS + a.S_pad);
        }
        mbar_wait(&bars[buf], (ch >> 1) & 1);
        // ---- phase A: rows x node columns of the chunk, E/O-folded over the mirrored column nodes.  Warps 2p and 2p + 1
        // share rows 2p, 2p + 1 and split the column-node pairs; a half warp owns one row and the chunk's 16 (term, step)
        // columns, so the two halves read the same 128 bytes of gs per load
        {
            const int arow = (w & ~1) + (lane >> 4), gl = lane & (FAR_CW - 1);
            double* eo = EO + (((size_t)buf * FAR_RB + arow) * FAR_TC + (gl & (FAR_TC - 1))) * EOS + (gl >> 3) * 2;
            static_assert(FAR_TC == 8, "the (term, step) split of gl assumes 8 steps per chunk");
            const double* lyp = Lys + arow * NN;
            constexpr int KA = (NH + 1) / 2, KB = NH - KA;      // node-column pairs of the even / odd warp of a row pair
            if (w & 1) far_phase_a<NN, KB>(gs + gl, lyp, eo, KA);
            else far_phase_a<NN, KA>(gs + gl, lyp, eo, 0);
        }
        if (sg_on) stg[(buf * FAR_TC + sg_t) * SGS + sg_s] = sgv;
        __syncthreads();   // every warp is done with gs: refill it; EO and the stage of this chunk are complete
        if (tid == 0 && ch + 2 < a.nchunk) {
            mbar_arrive_expect_tx(&bars[buf], CHUNK_BYTES);
            bulk_g2s(Gs + buf * NNODE * 2 * FAR_TC, Gp + (size_t)(ch + 2) * NNODE * 2 * FAR_TC, CHUNK_BYTES, &bars[buf]);
        }
        if (!rowlive) continue;
        // ---- phase B
        const int tend = min(FAR_TC, a.T - ch * FAR_TC);
        const double2* eo = reinterpret_cast<const double2*>(EO + ((size_t)buf * FAR_RB + w) * FAR_TC * EOS);
        const double2* sg = stg + buf * FAR_TC * SGS;
        const double* wsm = wns + w * FAR_P + lane;
        OutT* op = reinterpret_cast<OutT*>(reinterpret_cast<char*>(orow) + (long long)ch * FAR_TC * a.ostride);
        // one step in three parts: the interpolation sums (straight-line: 8 NH multiply-adds), the near stations (branches on the
        // patch's station count), the finish (reciprocal, retrend, clamp, stores)
        auto accumulate = [&](const double2* e, double (&num)[4], double (&den)[4]) {
            double enA = 0.0, onA = 0.0, edA = 0.0, odA = 0.0, enB = 0.0, onB = 0.0, edB = 0.0, odB = 0.0;
#pragma unroll
            for (int k = 0; k < NH; k++) {
                const double2 n2 = e[2 * k], d2 = e[2 * k + 1];     // (E_num, O_num), (E_den, O_den)
                enA = fma(cA[k], n2.x, enA);
                onA = fma(cA[NH + k], n2.y, onA);
                edA = fma(cA[k], d2.x, edA);
                odA = fma(cA[NH + k], d2.y, odA);
                enB = fma(cB[k], n2.x, enB);
                onB = fma(cB[NH + k], n2.y, onB);
                edB = fma(cB[k], d2.x, edB);
                odB = fma(cB[NH + k], d2.y, odB);
            }
            num[0] = enA + onA, num[1] = enA - onA, num[2] = enB + onB, num[3] = enB - onB;
            den[0] = edA + odA, den[1] = edA - odA, den[2] = edB + odB, den[3] = edB - odB;
        };
        auto near_add = [&](const double2* g, int tl, double (&num)[4], double (&den)[4]) {
            if (nq > 0) {
                const double2 x = g[1];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    num[j] = fma(wr[0][j], x.x, num[j]);
                    den[j] = fma(wr[0][j], x.y, den[j]);
                }
            }
            if (nq > 1) {
                const double2 x = g[2];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    num[j] = fma(wr[1][j], x.x, num[j]);
                    den[j] = fma(wr[1][j], x.y, den[j]);
                }
            }
            // further stations: weights from shared memory (the first nqm), then from the L1-resident global scratch; their
            // (residual, validity) pairs from the stage (the first NQS), then from global memory
            const int q1 = min(nq, FAR_NQR + a.nqm), q2 = min(nq, FAR_NQS);
            int qi = FAR_NQR;
            for (; qi < q1; qi++) {
                const double2 x = g[1 + qi];
                const double* wq = wsm + (qi - FAR_NQR) * (FAR_RB * FAR_P);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const double wv = wq[32 * j];
                    num[j] = fma(wv, x.x, num[j]);
                    den[j] = fma(wv, x.y, den[j]);
                }
            }
            for (; qi < nq; qi++) {
                const double2 x = qi < q2 ? g[1 + qi] : __ldg(a.rv + (size_t)(ch * FAR_TC + tl) * a.S + nearl[qi]);
                const double* wq = wn + (size_t)(qi - FAR_NQR - a.nqm) * (FAR_RB * FAR_P);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const double wv = wq[32 * j];      // written by this thread in the prologue: a plain (coherent) load
                    num[j] = fma(wv, x.x, num[j]);
                    den[j] = fma(wv, x.y, den[j]);
                }
            }
        };
        // quotient to ~4e-13 relative (reciprocal seed + two Newton steps) with the retrend fused in (idw.py:144)
        auto quotient = [&](const double2 pp, const double (&num)[4], const double (&den)[4], double (&val)[4]) {
#pragma unroll
            for (int j = 0; j < 4; j++) val[j] = fma(num[j], fast_rcp<2>(den[j]), fma(pp.x, Z[j], pp.y));
        };
        // pf / pm: the step's output row at the lane's forward (lane, 32 + lane) and mirror (127 - lane, 95 - lane) columns
        auto emit = [&](double (&val)[4], OutT* pf, OutT* pm, auto fullc) {
            if (CLAMP) {
                // all four strictly inside (vmin, vmax) by their high words alone (integer pipe, one warp-wide vote); otherwise,
                // NaNs included, the exact NaN-preserving clamp of utils.py:137-161
                bool out = false;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int hi = __double2hiint(val[j]);
                    const unsigned key = CLAMP == 2 ? (unsigned)hi : CLAMP == 3 ? (unsigned)hi & 0x7fffffffu : hi_key(hi);
                    out |= key - a.klo >= a.kspan;
                }
                if (__any_sync(0xffffffffu, out)) {
#pragma unroll
                    for (int j = 0; j < 4; j++) val[j] = clamp_nan_keep(val[j], a.vmin, a.vmax);
                }
            }
            if constexpr (decltype(fullc)::value) {
                __stcs(pf, (OutT)val[0]);
                __stcs(pm, (OutT)val[1]);
                __stcs(pf + 32, (OutT)val[2]);
                __stcs(pm - 32, (OutT)val[3]);
            } else {
                if (live[0]) __stcs(pf, (OutT)val[0]);
                if (live[1]) __stcs(pm, (OutT)val[1]);
                if (live[2]) __stcs(pf + 32, (OutT)val[2]);
                if (live[3]) __stcs(pm - 32, (OutT)val[3]);
            }
        };
        // Measured and dropped: a software pipeline over the steps (the sums of step t + 1 in one basic block with the reciprocal /
        // retrend chain of step t): 19.62 against 19.64 ms per C4 slab field; 3 CTAs per SM at 80 registers (~9 local loads per
        // step): 34.1 ms.  tools/dfma_issue_probe.cu shows why: a DFMA holds the sub-partition's issue port for both of its
        // pipe cycles (2.07 clk alone, + ~1 clk per integer / load instruction beside it), so a warp-step of F fp64 and O other
        // instructions costs 2 F + O issue cycles whatever the latency hiding: 2 x 91 + 94 = 276 here, 332 measured.
        // the step loop, specialised on whether all four cells of every lane exist (uniform per warp), two running output pointers
        auto steps = [&](auto fullc) {
            char* pf = reinterpret_cast<char*>(op + lane);
            char* pm = reinterpret_cast<char*>(op + (FAR_P - 1 - lane));
#pragma unroll 2
            for (int tl = 0; tl < tend; tl++, eo += EOS / 2, sg += SGS, pf += a.ostride, pm += a.ostride) {
                double num[4], den[4], val[4];
                accumulate(eo, num, den);
                near_add(sg, tl, num, den);
                quotient(sg[0], num, den, val);
                emit(val, reinterpret_cast<OutT*>(pf), reinterpret_cast<OutT*>(pm), fullc);
            }
        };
        if (full) steps(std::true_type{});
        else steps(std::false_type{});
    }
}

// ------------------------------------------------------------------------------------------------------------------
static void cheb_nodes(int nn, long double* xi) {
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int k = 0; k < nn; k++) xi[k] = cosl(pi * (2 * k + 1) / (2 * nn));
    for (int k = 0; k < nn / 2; k++) xi[nn - 1 - k] = -xi[k];      // exactly symmetric
}

int idw_far_plan(IdwFarPlan* plan, int S, const double* mx, const double* my, int ny, int nx, int grid_kind,
                 const double* gx, const double* gy, double power, int device) {
    plan->enabled = false;
    const char* k = getenv("SMRFB_IDW_KERNEL");
    const bool forced = k && !strcmp(k, "far");
    if (k && !forced && *k) return SMRFB_OK;                      // another kernel was asked for
    if (grid_kind != 0 || nx < 2 || ny < 2 || !(power > 0.0 && power <= 4.0)) return SMRFB_OK;
    if (S <= 16 && !forced) return SMRFB_OK;                      // the register-weight kernel is faster there
    const double dx = (gx[nx - 1] - gx[0]) / (nx - 1), dy = (gy[ny - 1] - gy[0]) / (ny - 1);
    if (dx == 0.0 || dy == 0.0 || !std::isfinite(dx) || !std::isfinite(dy)) return SMRFB_OK;
    for (int j = 0; j < nx; j++)
        if (fabs(gx[j] - (gx[0] + j * dx)) > 1e-7 * fabs(dx)) return SMRFB_OK;
    for (int i = 0; i < ny; i++)
        if (fabs(gy[i] - (gy[0] + i * dy)) > 1e-7 * fabs(dy)) return SMRFB_OK;
    plan->nn = power <= 2.0 ? 10 : 12;
    plan->theta = power <= 2.0 ? 6.0 : 8.0;
    plan->x0 = gx[0];
    plan->dx = dx;
    plan->y0 = gy[0];
    plan->dy = dy;
    plan->npx = (nx + FAR_P - 1) / FAR_P;
    plan->npy = (ny + FAR_P - 1) / FAR_P;
    const int np = plan->npx * plan->npy;
    const double hx = 0.5 * (FAR_P - 1) * fabs(dx), hy = 0.5 * (FAR_P - 1) * fabs(dy);
    std::vector<int> idx;
    auto near_lists = [&](double theta) {      // the stations closer than theta patch half widths to each patch
        const double rnear = theta * std::max(hx, hy);
        idx.clear();
        plan->near_off.assign(np + 1, 0);
        plan->max_near = 0;
        for (int pi = 0; pi < plan->npy; pi++)
            for (int pj = 0; pj < plan->npx; pj++) {
                const double xc = gx[0] + (pj * (double)FAR_P + 0.5 * (FAR_P - 1)) * dx;
                const double yc = gy[0] + (pi * (double)FAR_P + 0.5 * (FAR_P - 1)) * dy;
                const int p = pi * plan->npx + pj;
                for (int s = 0; s < S; s++) {
                    const double ex = std::max(0.0, fabs(mx[s] - xc) - hx), ey = std::max(0.0, fabs(my[s] - yc) - hy);
                    if (!(ex * ex + ey * ey >= rnear * rnear)) idx.push_back(s);   // NaN coordinates count as near
                }
                plan->near_off[p + 1] = (int)idx.size();
                plan->max_near = std::max(plan->max_near, plan->near_off[p + 1] - plan->near_off[p]);
            }
    };
    near_lists(plan->theta);
    // A dense network (many near stations per patch, e.g. 200 stations over a 40 km DEM) pays 4 flop per near station and
    // output: 12 x 12 nodes reach the same accuracy from theta = 4 (eps(4, 12, 2) = 4.9e-11 against eps(6, 10, 2) = 1.0e-10,
    // measured the same way), which halves the near region; the interpolation itself costs 6 flop per output more.
    if (power <= 2.0 && (double)idx.size() / np > 3.0 && !getenv("SMRFB_IDW_FAR_NN10")) {
        plan->nn = 12;
        plan->theta = 4.0;
        near_lists(plan->theta);
    }
    if (const char* e = getenv("SMRFB_IDW_FAR_THETA")) {                 // only ever wider
        const double th = atof(e);
        if (th > plan->theta) {
            plan->theta = th;
            near_lists(plan->theta);
        }
    }
    // Lagrange basis at the cell positions u_c = (c - 63.5) / 63.5 on the Chebyshev nodes, and its even / odd folding
    const int nn = plan->nn;
    long double xi[12];
    cheb_nodes(nn, xi);
    std::vector<double> tab((size_t)(FAR_P + FAR_P / 2) * nn);
    std::vector<long double> L((size_t)FAR_P * nn);
    for (int c = 0; c < FAR_P; c++) {
        const long double u = ((long double)c - 0.5L * (FAR_P - 1)) / (0.5L * (FAR_P - 1));
        for (int a = 0; a < nn; a++) {
            long double v = 1.0L;
            for (int m = 0; m < nn; m++)
                if (m != a) v *= (u - xi[m]) / (xi[a] - xi[m]);
            L[(size_t)c * nn + a] = v;
            tab[(size_t)c * nn + a] = (double)v;
        }
    }
    for (int c = 0; c < FAR_P / 2; c++)
        for (int a = 0; a < nn / 2; a++) {
            tab[(size_t)(FAR_P + c) * nn + a] = (double)(0.5L * (L[(size_t)c * nn + a] + L[(size_t)c * nn + nn - 1 - a]));
            tab[(size_t)(FAR_P + c) * nn + nn / 2 + a] = (double)(0.5L * (L[(size_t)c * nn + a] - L[(size_t)c * nn + nn - 1 - a]));
        }
    // stations exactly on a cell centre (the device tests (X - mx)^2 + (Y - my)^2 == 0, i.e. equality of the doubles)
    std::vector<std::array<int, 3>> coin;
    for (int sidx = 0; sidx < S; sidx++) {
        const long long j = llround((mx[sidx] - gx[0]) / dx), i = llround((my[sidx] - gy[0]) / dy);
        for (long long jj = std::max(0LL, j - 1); jj <= std::min<long long>(nx - 1, j + 1); jj++)
            for (long long ii = std::max(0LL, i - 1); ii <= std::min<long long>(ny - 1, i + 1); ii++)
                if (gx[jj] == mx[sidx] && gy[ii] == my[sidx]) coin.push_back({(int)ii, (int)jj, sidx});
    }
    std::sort(coin.begin(), coin.end());
    plan->ncoin = (int)coin.size();
    SMRFB_CUDA(cudaSetDevice(device));
    if (plan->ncoin) {
        std::vector<int> flat(3 * coin.size());
        for (size_t q = 0; q < coin.size(); q++)
            for (int c = 0; c < 3; c++) flat[c * coin.size() + q] = coin[q][c];
        SMRFB_CUDA(cudaMalloc((void**)&plan->d_coin, flat.size() * sizeof(int)));
        SMRFB_CUDA(cudaMemcpy(plan->d_coin, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    SMRFB_CUDA(cudaMalloc((void**)&plan->d_near_off, (np + 1) * sizeof(int)));
    SMRFB_CUDA(cudaMalloc((void**)&plan->d_near_idx, std::max<size_t>(1, idx.size()) * sizeof(int)));
    SMRFB_CUDA(cudaMalloc((void**)&plan->d_tab, tab.size() * sizeof(double)));
    SMRFB_CUDA(cudaMemcpy(plan->d_near_off, plan->near_off.data(), (np + 1) * sizeof(int), cudaMemcpyHostToDevice));
    if (!idx.empty()) SMRFB_CUDA(cudaMemcpy(plan->d_near_idx, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice));
    SMRFB_CUDA(cudaMemcpy(plan->d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    plan->enabled = true;
    return SMRFB_OK;
}

static void far_mark(IdwFarPlan* plan, cudaStream_t st) {
    if (!plan->timing) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    plan->tev.push_back(e);
}

template <int NN>
static int far_launch(cudaStream_t st, IdwFarPlan* plan, FarNodeArgs& na, FarArgs& fa, int npl, int nbands, int smem_max) {
    constexpr int NNODE = NN * NN;
    far_mark(plan, st);
    static const bool node_fma = getenv("SMRFB_FAR_NODE") && !strcmp(getenv("SMRFB_FAR_NODE"), "fma");
    if (node_fma) {
        const size_t node_smem = ((size_t)na.SB * NNODE + 2 * FAR_NODE_KB * FAR_NODE_COLS) * sizeof(double) + (size_t)na.S;
        SMRFB_CHECK(node_smem <= (size_t)smem_max, SMRFB_ERR_LIMIT, "IDW far field: node kernel needs %zu B of shared memory", node_smem);
        SMRFB_CUDA(cudaFuncSetAttribute(idw_far_node_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        idw_far_node_kernel<NN><<<npl, FAR_NODE_THREADS, node_smem, st>>>(na);
    } else {
        constexpr int NP = FarNodeMma<NN>::NP;
        const size_t fixed = (size_t)na.S + 16;
        const int sb_cap = (int)(((size_t)smem_max - fixed) / (NP * sizeof(double))) & ~7;      // stations per block, a multiple of 8
        na.SB = std::min((na.S + 7) & ~7, sb_cap);
        SMRFB_CHECK(na.SB >= 8, SMRFB_ERR_LIMIT, "IDW far field: no room for the node weights in shared memory");
        const size_t node_smem = (size_t)na.SB * NP * sizeof(double) + fixed;
        SMRFB_CUDA(cudaFuncSetAttribute(idw_far_node_mma_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        idw_far_node_mma_kernel<NN><<<npl, FAR_NODE_MMA_THREADS, node_smem, st>>>(na);
    }
    count_launch();
    SMRFB_CUDA(cudaGetLastError());
    far_mark(plan, st);
    const size_t smem = (size_t)(2 * NNODE * 2 * FAR_TC + 2 * FAR_RB * FAR_TC * (NN / 2 * 4 + 2) + FAR_RB * NN) * sizeof(double) +
                        (size_t)2 * FAR_TC * (1 + FAR_NQS) * sizeof(double2) + 2 * sizeof(uint64_t) +
                        (size_t)fa.nqm * FAR_RB * FAR_P * sizeof(double) +
                        (size_t)std::max(1, plan->max_near) * sizeof(int);
    SMRFB_CHECK(smem <= (size_t)smem_max, SMRFB_ERR_LIMIT, "IDW far field: %zu B of shared memory per CTA", smem);
    const dim3 grid(FAR_P / FAR_RB, plan->npx, nbands);
#define FAR_LAUNCH(OT, CL)                                                                                              \
    do {                                                                                                                \
        SMRFB_CUDA(cudaFuncSetAttribute(idw_far_kernel<NN, OT, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max)); \
        idw_far_kernel<NN, OT, CL><<<grid, FAR_THREADS, smem, st>>>(fa);                                                 \
    } while (0)
    if (fa.out_f32) {
        if (fa.clamp == 3) FAR_LAUNCH(float, 3);
        else if (fa.clamp == 2) FAR_LAUNCH(float, 2);
        else if (fa.clamp) FAR_LAUNCH(float, 1);
        else FAR_LAUNCH(float, 0);
    } else {
        if (fa.clamp == 3) FAR_LAUNCH(double, 3);
        else if (fa.clamp == 2) FAR_LAUNCH(double, 2);
        else if (fa.clamp) FAR_LAUNCH(double, 1);
        else FAR_LAUNCH(double, 0);
    }
#undef FAR_LAUNCH
    far_mark(plan, st);
    count_launch();
    SMRFB_CUDA(cudaGetLastError());
    return SMRFB_OK;
}

int idw_far_run(cudaStream_t st, IdwFarPlan* plan, const Geom& g, const double* d_mx, const double* d_my, int S,
                double power, const double* d_rec, int RS, int S_pad, const uint8_t* d_valid, int T, int detrend,
                int clamp, double vmin, double vmax, double* d_out, int out_f32, int row0, int nrows, int device) {
    const int nn = plan->nn, nnode = nn * nn;
    const int T_pad = (T + FAR_TC - 1) / FAR_TC * FAR_TC, nchunk = T_pad / FAR_TC, N = 2 * T_pad;
    const int pi0 = row0 / FAR_P, pi1 = (row0 + nrows - 1) / FAR_P;
    const int npl = (pi1 - pi0 + 1) * plan->npx;
    const int nbands = pi1 - pi0 + 1;          // grid z: patch rows of the window (x: the 16 bands of a patch row)
    SMRFB_CHECK(nbands <= 65535 && plan->npx <= 65535, SMRFB_ERR_LIMIT, "IDW far field: too many patches for one launch");
    int qmax = 0;
    for (int p = pi0 * plan->npx; p < (pi1 + 1) * plan->npx; p++) qmax = std::max(qmax, plan->near_off[p + 1] - plan->near_off[p]);
    // near stations 2 .. 2 + nqm - 1 keep their cell weights in shared memory (8 KB each), the rest go through a global scratch
    // that lives in L1.  More of them in shared memory was measured SLOWER on the dense C5 network (47.1 against 44.5 ms per
    // 512-step field at nqm = 5): the larger carve-out leaves the scratch of the remaining stations no L1 to live in.
    int nqm = FAR_NQM;
    if (const char* e = getenv("SMRFB_FAR_NQM")) nqm = std::max(0, atoi(e));
    const int qx = std::max(0, qmax - FAR_NQR - nqm);
    SMRFB_PROPAGATE(ensure_cap((void**)&plan->d_G, &plan->G_cap, (size_t)npl * nchunk * nnode * 2 * FAR_TC * sizeof(double)));
    SMRFB_PROPAGATE(ensure_cap((void**)&plan->d_rv, &plan->rv_cap, ((size_t)T_pad * S * 2 + (size_t)S * N) * sizeof(double)));
    SMRFB_PROPAGATE(ensure_cap((void**)&plan->d_wn, &plan->wn_cap,
                               std::max<size_t>(16, (size_t)plan->npx * nbands * (FAR_P / FAR_RB) * qx * FAR_RB * FAR_P * sizeof(double))));
    double2* d_rv = reinterpret_cast<double2*>(plan->d_rv);
    double* d_rvt = plan->d_rv + (size_t)T_pad * S * 2;
    idw_far_pack_kernel<<<T_pad, 128, 0, st>>>(d_rec, RS, d_valid, T, S, N, d_rv, d_rvt);
    count_launch();
    SMRFB_CUDA(cudaGetLastError());
    int smem_max = 0;
    SMRFB_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));

    FarNodeArgs na;
    na.near_off = plan->d_near_off;
    na.near_idx = plan->d_near_idx;
    na.mx = d_mx;
    na.my = d_my;
    na.rvt = d_rvt;
    na.G = plan->d_G;
    na.x0 = plan->x0;
    na.dx = plan->dx;
    na.y0 = plan->y0;
    na.dy = plan->dy;
    na.power = power;
    na.S = S;
    na.N = N;
    na.npx = plan->npx;
    na.pi0 = pi0;
    na.SB = std::min(S, (int)((size_t)(200 * 1024) / (nnode * sizeof(double))));
    long double xi[12];
    cheb_nodes(nn, xi);
    for (int k2 = 0; k2 < 12; k2++) na.xi.xi[k2] = k2 < nn ? (double)xi[k2] : 0.0;

    FarArgs fa;
    fa.g = g;
    fa.mx = d_mx;
    fa.my = d_my;
    fa.near_off = plan->d_near_off;
    fa.near_idx = plan->d_near_idx;
    fa.tab = plan->d_tab;
    fa.G = plan->d_G;
    fa.rv = d_rv;
    fa.rec = d_rec;
    fa.wn = plan->d_wn;
    fa.out = d_out;
    fa.power = power;
    fa.vmin = vmin;
    fa.vmax = vmax;
    fa.S = S;
    fa.RS = RS;
    fa.S_pad = S_pad;
    fa.T = T;
    fa.nchunk = nchunk;
    fa.clamp = clamp;
    fa.out_f32 = out_f32;
    fa.row0 = row0;
    fa.nrows = nrows;
    fa.npx = plan->npx;
    fa.pi0 = pi0;
    fa.qx = qx;
    fa.nqm = nqm;
    fa.ostride = (long long)nrows * g.nx * (out_f32 ? 4 : 8);
    {   // fast screen: key(vmin) < key(v) < key(vmax) on the high words proves vmin < v < vmax
        long long blo, bhi;
        memcpy(&blo, &vmin, 8);
        memcpy(&bhi, &vmax, 8);
        const int hlo = (int)(blo >> 32), hhi = (int)(bhi >> 32);
        if (clamp && hlo >= 0 && vmax > vmin) {                // vmin >= +0: the high word orders the non-negative doubles
            fa.clamp = 2;
            fa.klo = (unsigned)hlo + 1;
            fa.kspan = (unsigned)hhi > fa.klo ? (unsigned)hhi - fa.klo : 0;
        } else if (clamp && vmin < 0.0 && vmax > 0.0) {        // |v| < min(-vmin, vmax), compared on the high words
            fa.clamp = 3;
            fa.klo = 0;
            fa.kspan = (unsigned)std::min(hlo & 0x7fffffff, hhi);
        } else {
            const unsigned klo = hi_key(hlo), khi = hi_key(hhi);
            fa.klo = klo + 1;
            fa.kspan = khi > klo + 1 ? khi - (klo + 1) : 0;
        }
    }
    SMRFB_PROPAGATE(nn == 10 ? far_launch<10>(st, plan, na, fa, npl, nbands, smem_max)
                             : far_launch<12>(st, plan, na, fa, npl, nbands, smem_max));
    if (plan->ncoin) {
        FarCoinArgs ca;
        ca.crow = plan->d_coin;
        ca.ccol = plan->d_coin + plan->ncoin;
        ca.csta = plan->d_coin + 2 * plan->ncoin;
        ca.rv = d_rv;
        ca.rec = d_rec;
        ca.dem = g.dem;
        ca.out = d_out;
        ca.vmin = vmin;
        ca.vmax = vmax;
        ca.npair = plan->ncoin;
        ca.S = S;
        ca.RS = RS;
        ca.S_pad = S_pad;
        ca.nx = g.nx;
        ca.row0 = row0;
        ca.nrows = nrows;
        ca.T = T;
        ca.clamp = clamp;
        ca.out_f32 = out_f32;
        const long long n = (long long)plan->ncoin * T;
        idw_far_coin_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(ca);
        count_launch();
        SMRFB_CUDA(cudaGetLastError());
    }
    return SMRFB_OK;
}

}  // namespace smrfb
