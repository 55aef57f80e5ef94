// dk.cu -- detrended kriging (smrf/spatial/dk/dk.py + krige.c + lusolv.c) on sm_100a.
//
// Per NaN mask (set V of n valid stations) the reference solves, per cell, the ordinary-kriging
// system  [[ad,1],[1',0]] [w;mu] = [d;1]  over and over, each time dropping the highest-elevation
// station that got a negative weight (krige.c:134-198): up to n-1 full LU solves per cell, O(n^4).
//
// Here (validated against the reference on CPU):
//   1. host: B0 = inverse of the full (n+1)x(n+1) system, stations ordered by (elevation desc,
//      index asc) so that "highest-elevation negative weight" == "first negative weight".
//   2. dk_build_kernel -- one persistent CTA per SM, B (packed symmetric, <=162 KB) lives in shared
//      memory.  A tile of 32 neighbouring cells (one per warp) walks the elimination tree depth
//      first: cells that agree on the station to drop share ONE rank-1 downdate of B
//      (B -= b b'/beta, w -= (w_k/beta) b); where they disagree the tile forks: the inverse is
//      snapshotted once to global memory (L2) and every further child of the fork starts from that
//      exact copy.  Output: the final active set of every cell (bit mask).
//   3. dk_final_kernel -- one thread per cell: a fresh solve on the final active set in station
//      order with the reference's exact arithmetic (Crout LU, implicit scaling, >= pivot rule,
//      no FMA), which is precisely what the reference's last loop iteration computes; if a negative
//      weight shows up the reference loop simply continues here.  Weights go to an ELL table
//      [K][ncell] (station index uint8 + fp64 weight), K = largest final active set.
//   4. dk_distribute_kernel -- per timestep: sum_k w_k * residual[station_k] + p0*Z + p1, clamp,
//      streaming stores; HBM-write bound.
#include <algorithm>
#include <list>
#include <map>

#include "common.cuh"

namespace smrfb {

constexpr int DK_TILE = 32;
constexpr int DK_THREADS = 1024;
constexpr int DK_WARPS = DK_THREADS / 32;
constexpr int DK_CPW = DK_TILE / DK_WARPS;   // cells per warp
constexpr int DK_MAXN = 200;          // packed B for n+1 = 201 is 162 KB of shared memory
constexpr int DK_WORDS = 8;           // 256 active-set bits per cell
constexpr int DK_NSMAX = 31;          // largest final active set solved in thread-local memory
constexpr int DK_KREG = 16;           // ELL entries kept in registers by the distribute kernel
constexpr int DK_QMAX = (DK_MAXN + 1 + 31) / 32;  // 7 column slots per lane
constexpr int DK_MAXSNAP = 32;         // nested forks of a 32-cell tile
constexpr int DK_PEND = 4;             // rank-1 downdates of the inverse deferred and applied together (one pass over B)
constexpr int DK_TX = 8, DK_TY = 4;    // a tile is an 8 x 4 block of mesh cells (neighbours agree longest on what to drop)

struct DkPending {
    int k;          // station (rank index) this subgroup drops next
    unsigned mask;  // cells of the subgroup
    int depth;      // number of stations already dropped at the fork
    int snap;       // snapshot level holding the inverse at the fork
};

struct DkBuildArgs {
    int n, NN, NNP, nsel;
    int handoff;           // active-station count at which cells leave the shared walk for the exact per-cell solver
    long long ncell;
    Geom g;
    const double* sx;      // [n] station x in RANK order (coordinate mode)
    const double* sy;
    const double* dgrid;   // optional [ncell][n], COMPRESSED station order (krige_grid mode)
    const int* perm;       // [n] rank -> compressed station index
    const double* B0p;     // packed upper triangle of the inverse, rank order, NN*(NN+1)/2
    double* snap;          // [gridDim.x][DK_MAXSNAP][npk] fork snapshots of the packed inverse
    double* sglob;         // dk_sets_kernel with more than DK_MAXN stations: [gridDim.x][npk] the packed inverse itself
    uint32_t* final_act;   // [ncell][DK_WORDS], bits over RANK index
    unsigned long long* tile_counter;
    int* maxcount;
};

__device__ __forceinline__ double dist_rn(double X, double Y, double sx, double sy) {
    // dk.py:122  np.sqrt((Xa - mx[i])**2 + (Ya - my[i])**2), separate roundings
    double dx = __dsub_rn(X, sx), dy = __dsub_rn(Y, sy);
    return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// 1/x to ~1 ulp for any normal x (MUFU.RCP64H seed + two Newton steps): the IEEE divide's
// ~35-instruction sequence sat on the critical path of every elimination step.
__device__ __forceinline__ double dk_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

// Cells of a tile: an 8 x 4 block of the mesh (grid_kind 0, coordinate mode), else 32 consecutive cells.
struct DkTiling {
    int blocked, tiles_x;
    long long ntiles;
};
__host__ __device__ inline DkTiling dk_tiling(const Geom& g, long long ncell, bool has_dgrid) {
    DkTiling t;
    t.blocked = (!has_dgrid && g.kind == 0 && g.nx > 0 && (long long)g.ny * g.nx == ncell) ? 1 : 0;
    t.tiles_x = t.blocked ? (g.nx + DK_TX - 1) / DK_TX : 0;
    t.ntiles = t.blocked ? (long long)t.tiles_x * ((g.ny + DK_TY - 1) / DK_TY) : (ncell + DK_TILE - 1) / DK_TILE;
    return t;
}
// flat cell index of cell ci of a tile, -1 if it falls outside the grid
__device__ __forceinline__ long long dk_tile_cell(const DkTiling& tl, const Geom& g, long long ncell, long long tile, int ci) {
    if (tl.blocked) {
        const long long ty = tile / tl.tiles_x;
        const int tx = (int)(tile - ty * tl.tiles_x);
        const long long i = ty * DK_TY + (ci / DK_TX);
        const int j = tx * DK_TX + (ci % DK_TX);
        return (i < g.ny && j < g.nx) ? i * g.nx + j : -1;
    }
    const long long c = tile * DK_TILE + ci;
    return c < ncell ? c : -1;
}

__global__ void __launch_bounds__(DK_THREADS, 1) dk_build_kernel(DkBuildArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n, NN = a.NN, NNP = a.NNP;
    const int npk = NN * (NN + 1) / 2;
    double* Bp = reinterpret_cast<double*>(smem_raw);           // [npk] packed upper triangle of the inverse, WITHOUT the pending downdates
    double* w = Bp + npk;                                       // [32][NN] current weights of the tile's cells
    double* pcol = w + DK_TILE * NN;                            // [DK_PEND][NNP] extracted columns (zeros = inactive), pending ones first
    double* psm = pcol + DK_PEND * NNP;                         // [DK_PEND][NNP] column / beta
    int* rowoff = reinterpret_cast<int*>(psm + DK_PEND * NNP);  // [NNP]
    int* kstack = rowoff + NNP;                                 // [NN] stations dropped along the current path
    int* kc = kstack + NN;                                      // [32]
    int* ctrl = kc + DK_TILE;                                   // [8]
    DkPending* pend = reinterpret_cast<DkPending*>(ctrl + 8);   // [64]
    unsigned char* act = reinterpret_cast<unsigned char*>(pend + 64);  // [NNP]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* snap_g = a.snap + (size_t)blockIdx.x * DK_MAXSNAP * npk;   // fork snapshots of this CTA
    const DkTiling tiling = dk_tiling(a.g, a.ncell, a.dgrid != nullptr);
    const long long ntiles = tiling.ntiles;

    for (int i = tid; i < NNP; i += DK_THREADS) rowoff[i] = i * NN - (i * (i + 1)) / 2;   // idx(i,j>=i) = rowoff[i] + j
    __syncthreads();

    // The rank-1 downdates B[i][j] -= (b[i]/beta) * b[j] (j >= i) of the last np eliminations are DEFERRED: Bp holds
    // the inverse without them, and a column is corrected for them when it is extracted (np FMAs per element, in the
    // order and with the operand roles the eager update would have used: bit-identical).  flush() applies them to the
    // packed triangle in ONE pass -- B (162 KB at 200 stations) is read and written once per DK_PEND eliminations
    // instead of once per elimination, and that shared-memory traffic was what an elimination step cost.
    // Rows are dealt round-robin to the warps; zero entries (inactive stations) are no-ops, whole zero rows skipped.
    int np = 0;                                                // pending downdates (uniform across the CTA)
    auto flush = [&]() {
        if (np == 0) return;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            constexpr int QH = (DK_QMAX + 1) / 2;
            const int qa = half * QH, qb = half ? DK_QMAX : QH;
            double bj[DK_PEND][QH];
#pragma unroll
            for (int d = 0; d < DK_PEND; d++)
#pragma unroll
                for (int q = 0; q < QH; q++) {
                    const int j = lane + 32 * (qa + q);
                    bj[d][q] = (d < np && qa + q < qb && j < NN) ? pcol[d * NNP + j] : 0.0;
                }
            for (int i = warp; i < NN; i += DK_WARPS) {
                if ((i >> 5) >= qb) continue;                  // the row starts right of this half
                double bi[DK_PEND];
                bool any = false;
#pragma unroll
                for (int d = 0; d < DK_PEND; d++) {
                    bi[d] = d < np ? psm[d * NNP + i] : 0.0;
                    any |= bi[d] != 0.0;
                }
                if (!any) continue;
                double* row = Bp + rowoff[i];
                const int q0 = i >> 5;
#pragma unroll
                for (int q = 0; q < QH; q++) {
                    const int j = lane + 32 * (qa + q);
                    if (qa + q < qb && qa + q >= q0 && j >= i && j < NN) {
                        double v = row[j];
#pragma unroll
                        for (int d = 0; d < DK_PEND; d++) v = fma(-bi[d], bj[d][q], v);   // slots >= np hold zeros here
                        row[j] = v;
                    }
                }
            }
        }
        np = 0;
    };
    // element (i, k) of the CURRENT inverse: the stored value corrected for the pending downdates
    auto current = [&](int i, int k) {
        const int r = min(i, k), c = max(i, k);
        double v = Bp[rowoff[r] + c];
        for (int d = 0; d < np; d++) v = fma(-psm[d * NNP + r], pcol[d * NNP + c], v);
        return v;
    };

    for (;;) {
        if (tid == 0) ctrl[0] = (int)atomicAdd(a.tile_counter, 1ULL);
        __syncthreads();
        const long long tile = (unsigned)ctrl[0];
        if (tile >= ntiles) break;
        for (int p = tid; p < npk; p += DK_THREADS) Bp[p] = a.B0p[p];   // every tile starts from the pristine inverse (L2-resident)
        np = 0;
        long long mycell[DK_CPW];
        {   // right-hand side [dist to every station ; 1] in rank order
#pragma unroll
            for (int h = 0; h < DK_CPW; h++) {
                const int ci = warp + DK_WARPS * h;
                const long long cell = dk_tile_cell(tiling, a.g, a.ncell, tile, ci);
                mycell[h] = cell;
                const bool live = cell >= 0;
                double* wc = w + ci * NN;
                double X = 0.0, Y = 0.0;
                if (live && !a.dgrid) cell_xy(a.g, cell, X, Y);
                for (int j = lane; j < NN; j += 32) {
                    double v = 0.0;
                    if (live) {
                        if (j == n) v = 1.0;
                        else if (a.dgrid) v = a.dgrid[(size_t)cell * n + a.perm[j]];
                        else v = dist_rn(X, Y, a.sx[j], a.sy[j]);
                    }
                    wc[j] = v;
                }
            }
        }
        for (int i = tid; i < NNP; i += DK_THREADS) act[i] = (i < NN) ? 1 : 0;
        __syncthreads();
        {   // w = B0 * rhs (B symmetric, packed); each B element is loaded once for all of the warp's cells
            double acc[DK_CPW][DK_QMAX];
#pragma unroll
            for (int h = 0; h < DK_CPW; h++)
#pragma unroll
                for (int q = 0; q < DK_QMAX; q++) acc[h][q] = 0.0;
            for (int i = 0; i < NN; i++) {
                double r[DK_CPW];
#pragma unroll
                for (int h = 0; h < DK_CPW; h++) r[h] = w[(warp + DK_WARPS * h) * NN + i];
                const int ro = rowoff[i];
#pragma unroll
                for (int q = 0; q < DK_QMAX; q++) {
                    const int j = lane + 32 * q;
                    if (j < NN) {
                        const double bv = (j >= i) ? Bp[ro + j] : Bp[rowoff[j] + i];
#pragma unroll
                        for (int h = 0; h < DK_CPW; h++) acc[h][q] = fma(bv, r[h], acc[h][q]);
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int h = 0; h < DK_CPW; h++)
#pragma unroll
                for (int q = 0; q < DK_QMAX; q++) {
                    const int j = lane + 32 * q;
                    if (j < NN) w[(warp + DK_WARPS * h) * NN + j] = acc[h][q];
                }
        }
        unsigned tile_mask = 0;
        {
            const long long cl = dk_tile_cell(tiling, a.g, a.ncell, tile, lane);
            tile_mask = __ballot_sync(0xffffffffu, cl >= 0);
        }
        int depth = 0, npend = 0, nsnap = 0;
        unsigned n_apply = 0, n_restore = 0, n_fork = 0;
        __syncthreads();

        // SELECT for the cells in `mask`, excluding station kx (dropped in this very phase, act[kx] is still
        // being cleared): first (= highest elevation) active station with a negative weight; a cell with none
        // (or with few enough stations left) leaves the shared walk and publishes its active set.
        auto select = [&](unsigned mask, int kx, int ndrop) {
#pragma unroll
            for (int h = 0; h < DK_CPW; h++) {
                const int ci = warp + DK_WARPS * h;
                if ((mask >> ci) & 1u) {
                    const double* wc = w + ci * NN;
                    int k = -1;
                    for (int q = 0; q < DK_QMAX; q++) {
                        const int j = lane + 32 * q;
                        const bool c = (j < a.nsel) && (j != kx) && act[j] && (wc[j] < 0.0);
                        const unsigned bb = __ballot_sync(0xffffffffu, c);
                        if (bb) {
                            k = 32 * q + __ffs(bb) - 1;
                            break;
                        }
                    }
                    const bool hand = (k >= 0) && (n - ndrop <= a.handoff);
                    if (k < 0 || hand) {
                        const long long cell = mycell[h];
                        unsigned myword = 0;
                        for (int q = 0; q < DK_WORDS; q++) {
                            const int j = lane + 32 * q;
                            const unsigned bb = __ballot_sync(0xffffffffu, (j < n) && (j != kx) && act[j]);
                            if (lane == q) myword = bb;
                        }
                        if (hand && lane == DK_WORDS - 1) myword |= 0x80000000u;   // bit 255: handed off undecided
                        if (lane < DK_WORDS) a.final_act[(size_t)cell * DK_WORDS + lane] = myword;
                        k = -1;
                    }
                    if (lane == 0) kc[ci] = k;
                } else if (lane == 0) {
                    kc[ci] = -2;
                }
            }
        };
        select(tile_mask, -1, 0);

        for (;;) {   // ---- depth-first walk of the tile's elimination tree
            __syncthreads();   // kc[], w, B, act of the previous phase are in place
            // every warp partitions the group by chosen station (redundantly: no extra barrier in the common case)
            int k;
            unsigned mask;
            {
                const int kl = kc[lane];
                const unsigned livem = __ballot_sync(0xffffffffu, kl >= 0);
                const unsigned grp = __match_any_sync(0xffffffffu, kl);
                const bool leader = (kl >= 0) && (lane == __ffs(grp) - 1);
                const unsigned lead = __ballot_sync(0xffffffffu, leader);
                const int p = __popc(lead);
                if (p == 1) {          // the whole group drops the same station: apply directly
                    k = __shfl_sync(0xffffffffu, kl, __ffs(lead) - 1);
                    mask = livem;
                } else {
                    if (warp == 0 && leader) {   // fork: one pending entry per subgroup, all sharing snapshot level nsnap
                        const int slot = npend + __popc(lead & ((1u << lane) - 1u));
                        pend[slot].k = kl;
                        pend[slot].mask = grp;
                        pend[slot].depth = depth;
                        pend[slot].snap = nsnap;
                    }
                    bool fresh = false;
                    if (p > 0) {   // snapshot the inverse once for all children (exact: no rounding drift on the way back)
                        flush();             // the snapshot holds the CURRENT inverse
                        __syncthreads();
                        double* sg = snap_g + (size_t)nsnap * npk;
                        for (int q = tid; q < npk; q += DK_THREADS) sg[q] = Bp[q];
                        nsnap++;
                        npend += p;
                        n_fork++;
                        fresh = true;
                    }
                    __syncthreads();         // pend[] written; snapshot reads of B done
                    if (npend == 0) {        // tile finished
                        if (tid == 0) {
                            atomicAdd(a.tile_counter + 2, (unsigned long long)n_apply);
                            atomicAdd(a.tile_counter + 3, (unsigned long long)n_restore);
                            atomicAdd(a.tile_counter + 4, (unsigned long long)n_fork);
                        }
                        break;
                    }
                    npend--;
                    k = pend[npend].k;
                    mask = pend[npend].mask;
                    const int sdepth = pend[npend].depth, ssnap = pend[npend].snap;
                    const bool last_child = (npend == 0) || (pend[npend - 1].snap != ssnap);
                    if (!fresh) {   // come back to the fork point: restore the snapshot and the active flags
                        const double* sg = snap_g + (size_t)ssnap * npk;
                        for (int q = tid; q < npk; q += DK_THREADS) Bp[q] = sg[q];
                        np = 0;              // the pending downdates belonged to the branch just left
                        for (int i = tid; i < NNP; i += DK_THREADS) act[i] = (i < NN) ? 1 : 0;
                        __syncthreads();
                        for (int d = tid; d < sdepth; d += DK_THREADS) act[kstack[d]] = 0;
                        depth = sdepth;
                        n_restore++;
                    }
                    if (last_child) nsnap = ssnap;   // the snapshot level is free again
                    __syncthreads();                 // B / act restored; pend[npend] read by everyone
                }
            }
            // EXTRACT column k of the current inverse into pending slot np
            n_apply++;
            double* bcol = pcol + np * NNP;
            double* bsm = psm + np * NNP;
            const double beta = current(k, k);
            const double rbeta = dk_rcp(beta);
            for (int i = tid; i < NNP; i += DK_THREADS) {
                double v = 0.0;
                if (i < NN && i != k && act[i]) v = current(i, k);
                bcol[i] = v;
                bsm[i] = v * rbeta;
            }
            if (tid == 0) kstack[depth] = k;
            __syncthreads();
            // UPDATE the weights of the cells in mask, pick their next station, downdate the inverse
#pragma unroll
            for (int h = 0; h < DK_CPW; h++) {
                const int ci = warp + DK_WARPS * h;
                if ((mask >> ci) & 1u) {
                    double* wc = w + ci * NN;
                    const double f = wc[k] * rbeta;
                    for (int j = lane; j < NN; j += 32) wc[j] = fma(-f, bcol[j], wc[j]);
                    if (lane == (k & 31)) wc[k] = 0.0;   // same lane that just updated it
                    __syncwarp();
                }
            }
            select(mask, k, depth + 1);
            np++;
            if (np == DK_PEND) flush();
            if (tid == 0) act[k] = 0;
            depth++;
        }
    }
}

// ------------------------------------------------------------------------------------------
// dk_chain_kernel -- the elimination walk, second design ("one chain per tile, shared by active SET").
//
// Measured on the C4 geometry (tools/dk_path_sim.py, tools/dk_main_chain_sim.py): neighbouring cells drop the SAME
// stations, only in slightly different ORDER -- a 32 x 32 block of cells visits 0.66 distinct active sets per cell
// but 16 distinct path prefixes per cell in 8 x 4 tiles, and a walk keyed by path (dk_build_kernel: fork, snapshot,
// restore) pays one rank-1 downdate of the 162 KB inverse for each of them.  The weights of a cell depend on its
// active SET only.  Here a tile keeps ONE stored inverse S, that of the set A_chain = all stations minus those that
// EVERY live cell of the tile has dropped (and that have been folded in); a cell is ahead of it by its LEAD Q (the
// stations it has dropped since, |Q| <= DK2_QM).  Column k of the inverse of the cell's own set is the Schur complement
//         col = S[:, k] - S[:, Q] (S[Q, Q])^-1 S[Q, k]
// -- what |Q| rank-1 downdates would give -- and it is the SAME for all cells with the same (lead set, k): the
// "combos" of a step.  One step, in lock step for the tile: every cell (one warp each, weights in registers) picks its
// station; one warp per combo solves the small lead system; the CTA builds the combo columns; every cell applies its
// combo's column (7 FMAs per lane).  Cells that swap the order of two drops are in different combos for a step or two
// and then share again; nothing forks.  When the leads have >= DK2_FLMIN stations in common (or a lead is full) those
// stations are folded into S: one pass over the packed triangle for up to DK2_FL of them, once per station and TILE.
// Within ~30 cells of a station the orders differ by more than DK2_QM: when every live cell is full and nothing is
// common, the cells that have dropped the most popular station go on (progress guaranteed), the others wait on a job
// stack with a snapshot of S.
// ------------------------------------------------------------------------------------------
constexpr int DK2_QM = 12;              // largest lead of a cell over the chain
constexpr int DK2_LD = DK2_QM + 1;      // row stride of the lead system [S_QQ | S_Qk] (odd: conflict-free columns)
constexpr int DK2_PEND = 8;             // stations folded into the chain at once at most
constexpr int DK2_FLMIN = 4;            // fold when the leads have at least this many stations in common
constexpr int DK2_FL = 4;               // stations per pass over the packed inverse
constexpr int DK2_CB = 8;               // combo columns built per pass
constexpr int DK2_SCR = DK2_QM * DK2_LD;   // doubles of per-warp scratch: the lead system

constexpr int DK2_NNP_MAX = (DK_MAXN + 1 + 31) / 32 * 32;
// small arrays first, at compile-time offsets (the kernel runs at the 64-register cap: every pointer derived from n at
// run time was being recomputed at each use)
struct Dk2Small {
    int rowoff[DK2_NNP_MAX];               // idx(i, j >= i) = rowoff[i] + j
    int kstack[DK2_NNP_MAX];               // stations of the chain, in order
    unsigned qmask[DK_TILE * DK_WORDS];    // lead of every cell as a bit set
    int qlist[DK_TILE * DK2_QM];           // the same in drop order
    int qcnt[DK_TILE];                     // lead length
    int kc[DK_TILE];                       // station the cell wants to drop in this step
    int clive[DK_TILE];                    // cell still walking
    int clist[DK2_PEND];                   // stations being folded in
    unsigned taken[DK_WORDS];              // the same as a bit set
    int ctrl[8];                           // [0] tile, [1] some lead is full
    int jobs[2 * DK_TILE];                 // waiting groups: cell mask, chain length
    int kcb[DK2_CB], mcb[DK2_CB], lcb[DK2_CB];   // station, lead length, leader cell of the pass's combos
    unsigned char mact[DK2_NNP_MAX];       // station still in the chain's set
    double prmu[DK2_PEND];                 // 1 / pivot of the folded columns
    double cbeta[DK2_CB];                  // pivot of the combo columns
    double ycb[DK2_CB * DK2_QM];           // solutions of the lead systems
};
__host__ __device__ inline size_t dk2_union_doubles(int NN, int NNP) {
    const size_t stage = (size_t)DK_TILE * NN;                                       // right-hand sides at tile start
    const size_t work = (size_t)DK2_PEND * NNP + (size_t)DK_WARPS * DK2_SCR;         // columns, lead systems
    return stage > work ? stage : work;
}
__host__ __device__ inline size_t dk2_smem_bytes(int NN, int NNP) {
    const size_t npk = (size_t)NN * (NN + 1) / 2;
    return sizeof(Dk2Small) + (npk + dk2_union_doubles(NN, NNP)) * sizeof(double);
}

__global__ void __launch_bounds__(DK_THREADS, 1) dk_chain_kernel(DkBuildArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n, NN = a.NN, NNP = a.NNP;
    const int npk = NN * (NN + 1) / 2;
    Dk2Small& sm = *reinterpret_cast<Dk2Small*>(smem_raw);
    double* Sp = reinterpret_cast<double*>(smem_raw + sizeof(Dk2Small));   // [npk] packed upper triangle of the chain's inverse
    double* un = Sp + npk;                                       // union: rhs staging | columns, lead systems
    double* stage = un;                                          // [32][NN]
    double* pcol = un;                                           // [DK2_PEND][NNP] columns being folded in / combo columns
    double* scr = pcol + DK2_PEND * NNP;                         // [warps][DK2_SCR]
    int* const rowoff = sm.rowoff;
    int* const kstack = sm.kstack;
    unsigned* const qmask = sm.qmask;
    int* const qlist = sm.qlist;
    int* const qcnt = sm.qcnt;
    int* const kc = sm.kc;
    int* const clive = sm.clive;
    int* const clist = sm.clist;
    int* const ctrl = sm.ctrl;
    int* const jobs = sm.jobs;
    int* const kcb = sm.kcb;
    int* const mcb = sm.mcb;
    int* const lcb = sm.lcb;
    unsigned char* const mact = sm.mact;
    double* const prmu = sm.prmu;
    double* const cbeta = sm.cbeta;
    double* const ycb = sm.ycb;

    constexpr unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* snap_g = a.snap + (size_t)blockIdx.x * DK_MAXSNAP * npk;
    const DkTiling tiling = dk_tiling(a.g, a.ncell, a.dgrid != nullptr);
    const long long ntiles = tiling.ntiles;

    for (int i = tid; i < NNP; i += DK_THREADS) rowoff[i] = i * NN - (i * (i + 1)) / 2;   // idx(i, j >= i) = rowoff[i] + j
    __syncthreads();
    auto Sat = [&](int i, int j) { return Sp[rowoff[min(i, j)] + max(i, j)]; };
    auto pick = [&](const double (&v)[DK_QMAX], int slot) {      // v[slot] for a warp-uniform slot, no local memory
        double r = v[0];
#pragma unroll
        for (int s = 1; s < DK_QMAX; s++)
            if (slot == s) r = v[s];
        return r;
    };
    unsigned n_step = 0, n_pop = 0, n_split = 0, n_moved = 0, n_combo = 0;

    for (;;) {
        if (tid == 0) ctrl[0] = (int)atomicAdd(a.tile_counter, 1ULL);
        __syncthreads();
        const long long tile = (unsigned)ctrl[0];
        if (tile >= ntiles) break;
        for (int p = tid; p < npk; p += DK_THREADS) Sp[p] = a.B0p[p];
        const long long cell = dk_tile_cell(tiling, a.g, a.ncell, tile, warp);   // this warp's cell
        {   // right-hand side [dist to every station ; 1] in rank order
            double X = 0.0, Y = 0.0;
            if (cell >= 0 && !a.dgrid) cell_xy(a.g, cell, X, Y);
            for (int j = lane; j < NN; j += 32) {
                double v = 0.0;
                if (cell >= 0) {
                    if (j == n) v = 1.0;
                    else if (a.dgrid) v = a.dgrid[(size_t)cell * n + a.perm[j]];
                    else v = dist_rn(X, Y, a.sx[j], a.sy[j]);
                }
                stage[warp * NN + j] = v;
            }
        }
        for (int i = tid; i < NNP; i += DK_THREADS) mact[i] = (i < NN) ? 1 : 0;
        __syncthreads();
        double w[DK_QMAX];                                       // weights of this warp's cell on ITS active set: station lane + 32 s in slot s
#pragma unroll
        for (int s = 0; s < DK_QMAX; s++) w[s] = 0.0;
        for (int i = 0; i < NN; i++) {                           // w = S rhs (S symmetric, packed)
            const double r = stage[warp * NN + i];
            const int ro = rowoff[i];
#pragma unroll
            for (int s = 0; s < DK_QMAX; s++) {
                const int j = lane + 32 * s;
                if (j < NN) w[s] = fma((j >= i) ? Sp[ro + j] : Sp[rowoff[j] + i], r, w[s]);
            }
        }
        unsigned actbits = 0;                                    // bit s: station lane + 32 s is in this cell's active set
#pragma unroll
        for (int s = 0; s < DK_QMAX; s++)
            if (lane + 32 * s < n) actbits |= 1u << s;
        bool mylive = cell >= 0;
        if (lane < DK_WORDS) qmask[warp * DK_WORDS + lane] = 0u;
        if (lane == 0) {
            clive[warp] = mylive ? 1 : 0;
            qcnt[warp] = 0;
        }
        unsigned jobmask;
        {
            const long long cl = dk_tile_cell(tiling, a.g, a.ncell, tile, lane);
            jobmask = __ballot_sync(FULL, cl >= 0);
        }
        int nmain = 0, nstack = 0, since = 0;
        __syncthreads();                                         // staging is dead: the union now holds columns / lead systems

        for (;;) {   // ---- steps
            // ---- A: this warp's cell picks the first (= highest) active station with a negative weight
            {
                int k = -1;
                __syncwarp();                                                // lane 0's lead bookkeeping of the last step
                if (mylive && ((jobmask >> warp) & 1u)) {
#pragma unroll
                    for (int s = 0; s < DK_QMAX; s++) {
                        if (k < 0) {
                            const int j = lane + 32 * s;
                            const unsigned bb = __ballot_sync(FULL, ((actbits >> s) & 1u) && (j < a.nsel) && (w[s] < 0.0));
                            if (bb) k = 32 * s + __ffs(bb) - 1;
                        }
                    }
                    if (k < 0 || n - nmain - qcnt[warp] <= a.handoff) {      // finished, or few enough left for the exact solver
                        unsigned myword = 0;
#pragma unroll
                        for (int s = 0; s < DK_WORDS; s++) {
                            const unsigned bb = __ballot_sync(FULL, s < DK_QMAX && ((actbits >> s) & 1u));
                            if (lane == s) myword = bb;
                        }
                        if (k >= 0 && lane == DK_WORDS - 1) myword |= 0x80000000u;   // bit 255: handed off undecided
                        if (lane < DK_WORDS) a.final_act[(size_t)cell * DK_WORDS + lane] = myword;
                        mylive = false;
                        k = -1;
                        if (lane == 0) clive[warp] = 0;
                    }
                }
                if (lane == 0) kc[warp] = k;
            }
            __syncthreads();   // (1) choices published
            n_step++;

            // ---- B0 (every warp, identically): who is live, what the leads have in common, fold / split / next group
            bool lv = clive[lane] && ((jobmask >> lane) & 1u);               // lane = cell from here on
            unsigned livem = __ballot_sync(FULL, lv);
            if (!livem) {
                if (nstack == 0) break;                                      // tile finished
                nstack--;                                                    // a waiting group: back to the inverse and chain it was split off from
                jobmask = (unsigned)jobs[2 * nstack];
                nmain = jobs[2 * nstack + 1];
                const double* sg = snap_g + (size_t)nstack * npk;
                for (int q = tid; q < npk; q += DK_THREADS) Sp[q] = sg[q];
                for (int i = tid; i < NNP; i += DK_THREADS) mact[i] = (i < NN) ? 1 : 0;
                __syncthreads();
                for (int d = tid; d < nmain; d += DK_THREADS) mact[kstack[d]] = 0;
                n_pop++;
                __syncthreads();
                continue;
            }
            const int kl = lv ? kc[lane] : -1;
            const bool anyfull = __ballot_sync(FULL, lv && qcnt[lane] >= DK2_QM) != 0u;
            const bool canstep = __ballot_sync(FULL, lv && kl >= 0 && qcnt[lane] < DK2_QM) != 0u;
            unsigned cw[DK_WORDS];
            int nco = 0;
            since++;
            if (since >= DK2_FLMIN || anyfull || !canstep) {   // DK2_FLMIN common stations take at least as many steps to appear
#pragma unroll
                for (int q = 0; q < DK_WORDS; q++) {
                    cw[q] = __reduce_and_sync(FULL, lv ? qmask[lane * DK_WORDS + q] : FULL);
                    nco += __popc(cw[q]);
                }
            }
            bool fold = nco >= DK2_FLMIN || (anyfull && nco > 0);
            if (!fold && !canstep) {   // stuck: split by the station most cells have dropped
                int best = -1, bestc = 0;
#pragma unroll
                for (int q = 0; q < DK_WORDS; q++) {
                    const unsigned mine = lv ? qmask[lane * DK_WORDS + q] : 0u;
                    unsigned bits = __reduce_or_sync(FULL, mine);
                    while (bits) {
                        const int b = __ffs(bits) - 1;
                        bits &= bits - 1;
                        const int c = __popc(__ballot_sync(FULL, (mine >> b) & 1u));
                        if (c > bestc) {
                            bestc = c;
                            best = 32 * q + b;
                        }
                    }
                }
                const unsigned g1 = __ballot_sync(FULL, lv && ((qmask[lane * DK_WORDS + (best >> 5)] >> (best & 31)) & 1u));
                if (tid == 0) {
                    jobs[2 * nstack] = (int)(livem & ~g1);
                    jobs[2 * nstack + 1] = nmain;
                }
                double* sg = snap_g + (size_t)nstack * npk;                  // the group left behind restarts from this inverse
                for (int q = tid; q < npk; q += DK_THREADS) sg[q] = Sp[q];
                nstack++;
                n_split++;
                jobmask = g1;
                lv = lv && ((g1 >> lane) & 1u);
                livem = g1;
                nco = 0;
#pragma unroll
                for (int q = 0; q < DK_WORDS; q++) {
                    cw[q] = __reduce_and_sync(FULL, lv ? qmask[lane * DK_WORDS + q] : FULL);
                    nco += __popc(cw[q]);
                }
                fold = true;
                __syncthreads();                                             // snapshot read of S done before the fold writes it
            }
            if (fold) {
                // stations folded in: the first min(nco, DK2_PEND) common ones
                int nc = 0;
#pragma unroll
                for (int q = 0; q < DK_WORDS; q++) {
                    unsigned bits = cw[q], tk = 0;
                    while (bits && nc < DK2_PEND) {
                        const int b = __ffs(bits) - 1;
                        bits &= bits - 1;
                        tk |= 1u << b;
                        if (tid == 0) clist[nc] = 32 * q + b;
                        nc++;
                    }
                    if (warp == 1 && lane == q) sm.taken[q] = tk;            // the stations taken, as a bit set
                }
                if (warp == 0) {   // current column q of the inverse = stored column corrected for the columns folded before it
                    __syncwarp();
                    unsigned dropbits = 0;
                    for (int d = 0; d < nc; d++) {
                        const int q = clist[d];
                        double v[DK_QMAX];
#pragma unroll
                        for (int s = 0; s < DK_QMAX; s++) {
                            const int j = lane + 32 * s;
                            const bool on = (j < NN) && mact[j] && !((dropbits >> s) & 1u);
                            double x = 0.0;
                            if (on) {
                                x = Sat(j, q);
                                for (int e = 0; e < d; e++) x = fma(-(pcol[e * NNP + j] * prmu[e]), pcol[e * NNP + q], x);
                            }
                            v[s] = x;
                        }
                        const double mu = __shfl_sync(FULL, pick(v, q >> 5), q & 31);
                        if (lane == (q & 31)) dropbits |= 1u << (q >> 5);
#pragma unroll
                        for (int s = 0; s < DK_QMAX; s++) {
                            const int j = lane + 32 * s;
                            if (j < NNP) pcol[d * NNP + j] = (j == q) ? 0.0 : v[s];
                        }
                        if (lane == 0) prmu[d] = dk_rcp(mu);
                        __syncwarp();
                    }
                }
                __syncthreads();   // (F1) columns published
                if (mylive && ((jobmask >> warp) & 1u)) {   // this warp's cell: the lead loses the stations folded in (its weights do not change)
                    const int m = qcnt[warp];
                    const int qi = lane < m ? qlist[warp * DK2_QM + lane] : 0;
                    const bool keep = lane < m && !((sm.taken[qi >> 5] >> (qi & 31)) & 1u);
                    const unsigned km = __ballot_sync(FULL, keep);
                    __syncwarp();
                    if (keep) qlist[warp * DK2_QM + __popc(km & ((1u << lane) - 1u))] = qi;
                    if (lane == 0) qcnt[warp] = __popc(km);
                    if (lane < DK_WORDS) qmask[warp * DK_WORDS + lane] &= ~sm.taken[lane];
                }
                // S -= sum_d col_d col_d' / mu_d on the packed triangle, DK2_FL stations per pass; rows dealt round-robin
                // to the warps, zero entries (inactive stations) are no-ops, whole zero rows skipped
                for (int d0 = 0; d0 < nc; d0 += DK2_FL) {
                    const int nd = min(DK2_FL, nc - d0);
                    const double* pc = pcol + d0 * NNP;
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        constexpr int QH = (DK_QMAX + 1) / 2;
                        const int qa = half * QH, qb = half ? DK_QMAX : QH;
                        double bj[DK2_FL][QH];                         // this lane's columns: col_d[j] / mu_d
#pragma unroll
                        for (int d = 0; d < DK2_FL; d++)
#pragma unroll
                            for (int q = 0; q < QH; q++) {
                                const int j = lane + 32 * (qa + q);
                                bj[d][q] = (d < nd && qa + q < qb && j < NN) ? pc[d * NNP + j] * prmu[d0 + d] : 0.0;
                            }
                        const int iend = min(NN, 32 * qb);             // rows further down start right of this half
                        for (int i = warp; i < iend; i += DK_WARPS) {
                            if (!mact[i]) continue;                    // a station that left the chain earlier
                            double bi[DK2_FL];
#pragma unroll
                            for (int d = 0; d < DK2_FL; d++) bi[d] = d < nd ? pc[d * NNP + i] : 0.0;
                            double* row = Sp + rowoff[i];
                            const int q0 = i >> 5;
#pragma unroll
                            for (int q = 0; q < QH; q++) {
                                const int j = lane + 32 * (qa + q);
                                if (qa + q < qb && qa + q >= q0 && j >= i && j < NN) {
                                    double v = row[j];
#pragma unroll
                                    for (int d = 0; d < DK2_FL; d++) v = fma(-bi[d], bj[d][q], v);
                                    row[j] = v;
                                }
                            }
                        }
                    }
                    // (no barrier between passes: every element is handled by the same thread in each of them)
                }
                if (tid < nc) {
                    kstack[nmain + tid] = clist[tid];
                    mact[clist[tid]] = 0;
                }
                nmain += nc;
                n_moved += nc;
                since = nco > nc ? DK2_FLMIN : 0;                            // more common stations are waiting
                __syncthreads();   // (F2) S, mact, leads in place
            }

            // ---- B: combos = distinct (lead set, station) pairs among the cells that step now (every warp, identically)
            const int ml = qcnt[lane];
            const bool stepping = lv && kl >= 0 && ml < DK2_QM;
            unsigned grp = __ballot_sync(FULL, stepping) & __match_any_sync(FULL, kl);
#pragma unroll
            for (int q = 0; q < DK_WORDS; q++) grp &= __match_any_sync(FULL, qmask[lane * DK_WORDS + q]);
            const int lead_lane = __ffs(grp) - 1;                            // defined for stepping cells
            const unsigned leadm = __ballot_sync(FULL, stepping && lane == lead_lane);
            const int ncombo = __popc(leadm);
            const int combo_of_lane = stepping ? __popc(leadm & ((1u << lead_lane) - 1u)) : -1;
            const int mycombo = __shfl_sync(FULL, combo_of_lane, warp);      // this warp's cell
            const int myk = __shfl_sync(FULL, kl, warp);
            n_combo += ncombo;
            for (int g0 = 0; g0 < ncombo; g0 += DK2_CB) {
                const int ncb = min(DK2_CB, ncombo - g0);
                if (warp < ncb) {   // one warp per combo: y = (S[Q, Q])^-1 S[Q, k] by Gauss-Jordan in drop order, no pivoting
                    const int L = __fns(leadm, 0, g0 + warp + 1);            // leader cell of the combo
                    const int k = kc[L], m = qcnt[L];
                    double* G = scr + warp * DK2_SCR;
                    const int myq = lane < m ? qlist[L * DK2_QM + lane] : 0;
                    for (int e0 = 0; e0 < m * (m + 1); e0 += 32) {
                        const int e = e0 + lane;
                        const bool ok = e < m * (m + 1);
                        const int i = ok ? e / (m + 1) : 0, c = ok ? e - i * (m + 1) : 0;
                        const int qi = __shfl_sync(FULL, myq, i), qc = __shfl_sync(FULL, myq, c < m ? c : 0);
                        if (ok) G[i * DK2_LD + c] = Sat(qi, c < m ? qc : k);
                    }
                    __syncwarp();
                    for (int p = 0; p < m; p++) {
                        const double r = dk_rcp(G[p * DK2_LD + p]);
                        const int wd = m - p;                                // columns p + 1 .. m
                        for (int e = lane; e < m * wd; e += 32) {
                            const int i = e / wd, c = p + 1 + (e - i * wd);
                            if (i != p) G[i * DK2_LD + c] = fma(-(G[i * DK2_LD + p] * r), G[p * DK2_LD + c], G[i * DK2_LD + c]);
                        }
                        __syncwarp();
                    }
                    if (lane < m) ycb[warp * DK2_QM + lane] = G[lane * DK2_LD + m] * dk_rcp(G[lane * DK2_LD + lane]);
                    if (lane == 0) {
                        kcb[warp] = k;
                        mcb[warp] = m;
                        lcb[warp] = L;
                    }
                }
                __syncthreads();   // (2) lead systems solved
                for (int item = warp; item < ncb * DK_QMAX; item += DK_WARPS) {   // combo columns, 32 elements per item
                    const int gw = item / DK_QMAX, s = item - gw * DK_QMAX;
                    const int j = lane + 32 * s;
                    const int k = kcb[gw], m = mcb[gw], L = lcb[gw];
                    double x = 0.0;
                    if (j < NN && mact[j] && !((qmask[L * DK_WORDS + (j >> 5)] >> (j & 31)) & 1u)) {
                        x = Sat(j, k);
                        for (int i = 0; i < m; i++) x = fma(-Sat(j, qlist[L * DK2_QM + i]), ycb[gw * DK2_QM + i], x);
                    }
                    if (j == k) {
                        cbeta[gw] = x;
                        x = 0.0;
                    }
                    if (j < NNP) pcol[gw * NNP + j] = x;
                }
                __syncthreads();   // (3) columns built
                if (mycombo >= g0 && mycombo < g0 + ncb) {   // this warp's cell drops myk: w -= (w_k / beta) col
                    const int gw = mycombo - g0;
                    const double f = __shfl_sync(FULL, pick(w, myk >> 5), myk & 31) * dk_rcp(cbeta[gw]);
#pragma unroll
                    for (int s = 0; s < DK_QMAX; s++)
                        if (lane + 32 * s < NN) w[s] = fma(-f, pcol[gw * NNP + lane + 32 * s], w[s]);
                    if (lane == (myk & 31)) actbits &= ~(1u << (myk >> 5));
                    if (lane == 0) {
                        const int m = qcnt[warp];
                        qlist[warp * DK2_QM + m] = myk;
                        qcnt[warp] = m + 1;
                        qmask[warp * DK_WORDS + (myk >> 5)] |= 1u << (myk & 31);
                    }
                }
                if (g0 + DK2_CB < ncombo) __syncthreads();   // the next pass reuses the column buffer
            }
        }
    }
    if (tid == 0) {
        atomicAdd(a.tile_counter + 2, (unsigned long long)n_step);
        atomicAdd(a.tile_counter + 3, (unsigned long long)n_pop);
        atomicAdd(a.tile_counter + 4, (unsigned long long)n_split);
        atomicAdd(a.tile_counter + 5, (unsigned long long)n_moved);
        atomicAdd(a.tile_counter + 6, (unsigned long long)n_combo);
    }
}

// ------------------------------------------------------------------------------------------
// dk_sets_kernel -- the same walk as dk_chain_kernel (one stored inverse per tile, leads, combos, folds, splits) with the
// cell weights TRANSPOSED: warp w (of 16) owns the stations 14 w .. 14 w + 13 and lane l the cells l, l + 32 (a tile is 8 x 8
// cells), so a thread holds 14 x 2 weights (512 threads: at 1024 the 64-register cap spilled them to local memory).  In dk_chain_kernel a warp owned one cell and every per-cell operation was a
// warp-wide instruction sequence (ncu: 5.3 M warp instructions per 32-cell tile, issue-bound); here picking the first
// negative weight is 7 compares and one shared-memory atomicMin per thread, applying a combo column is 7 FMAs per cell
// with the column read as a (mostly broadcast) gather, and twice as many cells share every fold of the inverse.
// The per-cell bookkeeping (leads, live flags, combos) is done by warp 0 with lane = cell.
// ------------------------------------------------------------------------------------------
constexpr int DK3_NC = 2;                   // cell columns per lane: a tile is 32 * DK3_NC cells
constexpr int DK3_FOLD = 4;                 // stations folded into the chain at once
constexpr int DK3_C = 32 * DK3_NC;
constexpr int DK3_TX = 8, DK3_TY = 8;       // tile shape on a mesh
constexpr int DK3_THREADS = 512;
constexpr int DK3_WARPS = DK3_THREADS / 32;
constexpr int DK3_NMAX = 255;               // valid stations per mask: bit 255 of a cell's set flags a hand-off, station indices are uint8
constexpr int DK3_NNP_MAX = 256;
constexpr int DK3_MAXSNAP = DK3_C;          // a split leaves at least one cell behind
constexpr int DK3_BIG = 0x7fffffff;

struct Dk3Small {
    int rowoff[DK3_NNP_MAX];
    int kstack[DK3_NNP_MAX];
    unsigned qmask[DK3_C * DK_WORDS];
    int qlist[DK3_C * DK2_QM];
    int qcnt[DK3_C];
    int kc[2][DK3_C];                       // station picked by the cell in this step (atomicMin over the owner warps), double buffered
    int clive[DK3_C];
    int cid[DK3_C];                         // combo of the cell in this step, -1 = does not step
    int clist[DK2_PEND];
    unsigned taken[DK_WORDS];
    unsigned chainw[DK_WORDS];              // stations still in the chain's set, as a bit set
    int ctrl[16];
    unsigned stats[8];                      // thread 0: steps, job pops, splits, stations folded, combos
    int jobs[(DK3_NC + 1) * DK3_MAXSNAP];
    int kcb[DK3_C], mcb[DK3_C], lcb[DK3_C]; // station, lead length, leader cell of the step's combos
    long long cellidx[DK3_C];               // flat grid index of the tile's cells (-1 outside)
    unsigned char mact[DK3_NNP_MAX];
    double wk[DK3_C];                       // weight of the picked station
    double prmu[DK2_PEND];
    double piv[DK3_FOLD * DK3_FOLD];        // piv[e][d]: entry of folded column e at the station of column d > e
    double rbeta[DK2_CB];                   // 1 / pivot of the pass's combo columns
    double ycb[DK2_CB * DK2_QM];
    int qcb[DK2_CB * DK2_QM];               // the lead of the pass's combos
    double cx[DK3_C], cy[DK3_C];
};
__host__ __device__ inline int dk3_stage_rows(int NN, int NNP) {   // rows of the right-hand-side stage that fit beside the work buffers
    const size_t work = (size_t)2 * DK2_PEND * NNP + (size_t)DK2_CB * DK2_SCR;
    const size_t rows = work / DK3_C;
    return (int)(rows < 8 ? 8 : (rows > (size_t)NN ? (size_t)NN : rows));
}
__host__ __device__ inline size_t dk3_union_doubles(int NN, int NNP) {
    const size_t work = (size_t)2 * DK2_PEND * NNP + (size_t)DK2_CB * DK2_SCR;
    const size_t stage = (size_t)dk3_stage_rows(NN, NNP) * DK3_C;
    return stage > work ? stage : work;
}
__host__ __device__ inline size_t dk3_smem_bytes(int NN, int NNP, bool sglob) {
    const size_t npk = sglob ? 0 : (size_t)NN * (NN + 1) / 2;     // more than 200 stations: the packed inverse lives in global memory (L2)
    return sizeof(Dk3Small) + (npk + dk3_union_doubles(NN, NNP)) * sizeof(double);
}
struct Dk3Tiling {
    int blocked, tiles_x;
    long long ntiles;
};
__host__ __device__ inline Dk3Tiling dk3_tiling(const Geom& g, long long ncell, bool has_dgrid) {
    Dk3Tiling t;
    t.blocked = (!has_dgrid && g.kind == 0 && g.nx > 0 && (long long)g.ny * g.nx == ncell) ? 1 : 0;
    t.tiles_x = t.blocked ? (g.nx + DK3_TX - 1) / DK3_TX : 0;
    t.ntiles = t.blocked ? (long long)t.tiles_x * ((g.ny + DK3_TY - 1) / DK3_TY) : (ncell + DK3_C - 1) / DK3_C;
    return t;
}
__device__ __forceinline__ long long dk3_tile_cell(const Dk3Tiling& tl, const Geom& g, long long ncell, long long tile, int ci) {
    if (tl.blocked) {
        const long long ty = tile / tl.tiles_x;
        const int tx = (int)(tile - ty * tl.tiles_x);
        const long long i = ty * DK3_TY + (ci / DK3_TX);
        const int j = tx * DK3_TX + (ci % DK3_TX);
        return (i < g.ny && j < g.nx) ? i * g.nx + j : -1;
    }
    const long long c = tile * DK3_C + ci;
    return c < ncell ? c : -1;
}

// SPW: stations per warp (32 x SPW >= n + 1); SGLOB: the packed inverse does not fit shared memory (n > 200) and is kept in
// a per-CTA global buffer instead -- every access then costs an L2 round trip: correct, several times slower
template <int SPW, int QS, bool SGLOB>
__global__ void __launch_bounds__(DK3_THREADS, 1) dk_sets_kernel(DkBuildArgs a) {
    constexpr int DK3_SPW = SPW;       // stations per warp; QS: 32-wide slots that cover the n + 1 rows of a column
    static_assert(DK3_WARPS * SPW >= 32 * QS, "every station needs an owner warp");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n, NN = a.NN, NNP = a.NNP;
    const int npk = NN * (NN + 1) / 2;
    Dk3Small& sm = *reinterpret_cast<Dk3Small*>(smem_raw);
    double* Sp = SGLOB ? a.sglob + (size_t)blockIdx.x * npk
                       : reinterpret_cast<double*>(smem_raw + sizeof(Dk3Small));   // [npk] packed upper triangle of the chain's inverse
    double* un = reinterpret_cast<double*>(smem_raw + sizeof(Dk3Small)) + (SGLOB ? 0 : npk);
    double* stage = un;                                          // [SR][64] right-hand sides at tile start
    double* pcol = un;                                           // [DK2_PEND][NNP] columns being folded in / combo columns
    double* psm = pcol + DK2_PEND * NNP;                         // [DK2_PEND][NNP] the folded columns / pivot
    double* scr = psm + DK2_PEND * NNP;                          // [DK2_CB][DK2_SCR] lead systems
    constexpr unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j0 = warp * DK3_SPW;                               // this warp's stations j0 .. j0 + 6
    double* snap_g = a.snap + (size_t)blockIdx.x * DK3_MAXSNAP * npk;
    const Dk3Tiling tiling = dk3_tiling(a.g, a.ncell, a.dgrid != nullptr);
    const long long ntiles = tiling.ntiles;
    const int SR = dk3_stage_rows(NN, NNP);

    for (int i = tid; i < DK3_NNP_MAX; i += DK3_THREADS) sm.rowoff[i] = i * NN - (i * (i + 1)) / 2;
    __syncthreads();
    auto Sat = [&](int i, int j) { return Sp[sm.rowoff[min(i, j)] + max(i, j)]; };
    if (tid < 8) sm.stats[tid] = 0u;
#ifdef DK3_PROFILE
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pc0 = 0, qt[6] = {0, 0, 0, 0, 0, 0}, qc0 = 0;   // thread 0: clocks per phase (init, pick, book, fold, solve, apply)
#define DK3_T0() do { if (tid == 0) pc0 = clock64(); } while (0)
#define DK3_T(i) do { if (tid == 0) { const long long c_ = clock64(); pt[i] += c_ - pc0; pc0 = c_; } } while (0)
#define DK3_Q0() do { if (tid == 0) qc0 = clock64(); } while (0)
#define DK3_Q(i) do { if (tid == 0) { const long long c_ = clock64(); qt[i] += c_ - qc0; qc0 = c_; } } while (0)
#else
#define DK3_Q0() do {} while (0)
#define DK3_Q(i) do {} while (0)
#define DK3_T0() do {} while (0)
#define DK3_T(i) do {} while (0)
#endif

    for (;;) {
        if (tid == 0) sm.ctrl[0] = (int)atomicAdd(a.tile_counter, 1ULL);
        __syncthreads();
        const long long tile = (unsigned)sm.ctrl[0];
        if (tile >= ntiles) break;
        DK3_T0();
        for (int p = tid; p < npk; p += DK3_THREADS) Sp[p] = a.B0p[p];
        if (tid < DK3_C) {
            const long long cell = dk3_tile_cell(tiling, a.g, a.ncell, tile, tid);
            sm.cellidx[tid] = cell;
            double X = 0.0, Y = 0.0;
            if (cell >= 0 && !a.dgrid) cell_xy(a.g, cell, X, Y);
            sm.cx[tid] = X;
            sm.cy[tid] = Y;
            sm.clive[tid] = cell >= 0 ? 1 : 0;
            sm.qcnt[tid] = 0;
            sm.kc[0][tid] = DK3_BIG;
            sm.kc[1][tid] = DK3_BIG;
            sm.cid[tid] = -1;
        }
        for (int i = tid; i < DK3_C * DK_WORDS; i += DK3_THREADS) sm.qmask[i] = 0u;
        for (int i = tid; i < DK3_NNP_MAX; i += DK3_THREADS) sm.mact[i] = (i < NN) ? 1 : 0;
        if (tid < DK_WORDS) {
            const int lo = 32 * tid;
            sm.chainw[tid] = n > lo ? (n - lo >= 32 ? FULL : ((1u << (n - lo)) - 1u)) : 0u;
        }
        __syncthreads();
        // ---- weights on the full set: W = S rhs, the right-hand sides staged SR rows at a time
        double W[DK3_NC][DK3_SPW];
#pragma unroll
        for (int c = 0; c < DK3_NC; c++)
#pragma unroll
            for (int s = 0; s < DK3_SPW; s++) W[c][s] = 0.0;
        for (int i0 = 0; i0 < NN; i0 += SR) {
            const int ni = min(SR, NN - i0);
            for (int e = tid; e < ni * DK3_C; e += DK3_THREADS) {
                const int i = i0 + e / DK3_C, ci = e % DK3_C;
                const long long cell = sm.cellidx[ci];
                double v = 0.0;
                if (cell >= 0) {
                    if (i == n) v = 1.0;
                    else if (a.dgrid) v = a.dgrid[(size_t)cell * n + a.perm[i]];
                    else v = dist_rn(sm.cx[ci], sm.cy[ci], a.sx[i], a.sy[i]);
                }
                stage[e] = v;
            }
            __syncthreads();
            for (int ii = 0; ii < ni; ii++) {
                const int i = i0 + ii;
                double r[DK3_NC];
#pragma unroll
                for (int c = 0; c < DK3_NC; c++) r[c] = stage[ii * DK3_C + lane + 32 * c];
#pragma unroll
                for (int s = 0; s < DK3_SPW; s++) {
                    const int j = j0 + s;
                    if (j < NN) {
                        const double b = Sat(j, i);
#pragma unroll
                        for (int c = 0; c < DK3_NC; c++) W[c][s] = fma(b, r[c], W[c][s]);
                    }
                }
            }
            __syncthreads();
        }
        unsigned act[DK3_NC];                                    // bit s: station j0 + s is in the cell's active set and may be dropped
#pragma unroll
        for (int c = 0; c < DK3_NC; c++) {
            act[c] = 0u;
#pragma unroll
            for (int s = 0; s < DK3_SPW; s++)
                if (j0 + s < a.nsel) act[c] |= 1u << s;                // stations at elevation <= 0 (rank >= nsel) are never dropped
        }
        unsigned jobmask[DK3_NC];
#pragma unroll
        for (int c = 0; c < DK3_NC; c++) jobmask[c] = __ballot_sync(FULL, sm.cellidx[lane + 32 * c] >= 0);
        int nmain = 0, nstack = 0, since = 0, par = 0;
        DK3_T(0);

        for (;;) {   // ---- steps
            // ---- A: every cell's first (= highest) active station with a negative weight: the minimum over the owner warps
#pragma unroll
            for (int c = 0; c < DK3_NC; c++) {
                const int ci = lane + 32 * c;
                if (((jobmask[c] >> lane) & 1u) && sm.clive[ci]) {
                    int best = DK3_BIG;
#pragma unroll
                    for (int s = DK3_SPW - 1; s >= 0; s--)
                        if (((act[c] >> s) & 1u) && (W[c][s] < 0.0)) best = j0 + s;
                    if (best != DK3_BIG) atomicMin(&sm.kc[par][ci], best);
                }
            }
            __syncthreads();   // (1) choices in place
            if (tid == 0) sm.stats[0]++;
            DK3_T(1);

            // ---- B (warp 0, lane = cell): finished cells, folds, splits, the step's combos
            if (warp == 0) {
                bool lv[DK3_NC];
                int kl[DK3_NC];
                unsigned livem[DK3_NC];
                bool anyfull = false, canstep = false, anylive = false;
#pragma unroll
                for (int c = 0; c < DK3_NC; c++) {
                    const int ci = lane + 32 * c;
                    lv[c] = sm.clive[ci] && ((jobmask[c] >> lane) & 1u);
                    const int k = sm.kc[par][ci], m = sm.qcnt[ci];
                    sm.kc[par ^ 1][ci] = DK3_BIG;                            // the next step's slot
                    if (lv[c] && (k == DK3_BIG || n - nmain - m <= a.handoff)) {   // finished, or few enough left for the exact solver
                        const long long cell = sm.cellidx[ci];
#pragma unroll
                        for (int q = 0; q < DK_WORDS; q++) {
                            unsigned word = sm.chainw[q] & ~sm.qmask[ci * DK_WORDS + q];
                            if (q == DK_WORDS - 1 && k != DK3_BIG) word |= 0x80000000u;   // bit 255: handed off undecided
                            a.final_act[(size_t)cell * DK_WORDS + q] = word;
                        }
                        sm.clive[ci] = 0;
                        lv[c] = false;
                    }
                    kl[c] = lv[c] ? k : -1;
                    livem[c] = __ballot_sync(FULL, lv[c]);
                    anylive |= livem[c] != 0u;
                    anyfull |= __ballot_sync(FULL, lv[c] && m >= DK2_QM) != 0u;
                    canstep |= __ballot_sync(FULL, lv[c] && m < DK2_QM) != 0u;
                    sm.cid[ci] = -1;
                }
                int action = 0, nc = 0, snap = 0, ncombo = 0;
                if (!anylive) {
                    action = nstack == 0 ? 2 : 1;
                } else {
                    unsigned cw[DK_WORDS];
                    int nco = 0;
                    since++;
                    auto common = [&]() {
                        nco = 0;
#pragma unroll
                        for (int q = 0; q < DK_WORDS; q++) {
                            unsigned x = FULL;
#pragma unroll
                            for (int c = 0; c < DK3_NC; c++) x &= lv[c] ? sm.qmask[(lane + 32 * c) * DK_WORDS + q] : FULL;
                            cw[q] = __reduce_and_sync(FULL, x);
                            nco += __popc(cw[q]);
                        }
                    };
#pragma unroll
                    for (int q = 0; q < DK_WORDS; q++) cw[q] = 0u;
                    if (since >= DK2_FLMIN || anyfull || !canstep) common();
                    bool fold = nco >= DK2_FLMIN || (anyfull && nco > 0);
                    if (!fold && !canstep) {   // stuck: split by the station most cells have dropped
                        int best = -1, bestc = 0;
#pragma unroll
                        for (int q = 0; q < DK_WORDS; q++) {
                            unsigned mine[DK3_NC], u = 0u;
#pragma unroll
                            for (int c = 0; c < DK3_NC; c++) {
                                mine[c] = lv[c] ? sm.qmask[(lane + 32 * c) * DK_WORDS + q] : 0u;
                                u |= mine[c];
                            }
                            unsigned bits = __reduce_or_sync(FULL, u);
                            while (bits) {
                                const int b = __ffs(bits) - 1;
                                bits &= bits - 1;
                                int cnt = 0;
#pragma unroll
                                for (int c = 0; c < DK3_NC; c++) cnt += __popc(__ballot_sync(FULL, (mine[c] >> b) & 1u));
                                if (cnt > bestc) {
                                    bestc = cnt;
                                    best = 32 * q + b;
                                }
                            }
                        }
#pragma unroll
                        for (int c = 0; c < DK3_NC; c++) {
                            const unsigned g1 = __ballot_sync(FULL, lv[c] && ((sm.qmask[(lane + 32 * c) * DK_WORDS + (best >> 5)] >> (best & 31)) & 1u));
                            if (lane == 0) sm.jobs[(DK3_NC + 1) * nstack + c] = (int)(livem[c] & ~g1);
                            jobmask[c] = g1;
                            lv[c] = lv[c] && ((g1 >> lane) & 1u);
                            livem[c] = g1;
                        }
                        if (lane == 0) sm.jobs[(DK3_NC + 1) * nstack + DK3_NC] = nmain;
                        nstack++;
                        snap = 1;
                        common();
                        fold = true;
                    }
                    if (fold) {
#pragma unroll
                        for (int q = 0; q < DK_WORDS; q++) {
                            unsigned bits = cw[q], tk = 0;
                            while (bits && nc < DK3_FOLD) {
                                const int b = __ffs(bits) - 1;
                                bits &= bits - 1;
                                tk |= 1u << b;
                                if (lane == 0) sm.clist[nc] = 32 * q + b;
                                nc++;
                            }
                            if (lane == q) sm.taken[q] = tk;
                        }
                        since = nco > nc ? DK2_FLMIN : 0;
                    }
                    action = fold ? 3 : 0;
                    if (lane == 0) {
#pragma unroll
                        for (int c = 0; c < DK3_NC; c++) sm.ctrl[8 + c] = (int)jobmask[c];
                    }
                    if (!fold) {   // combos right away: distinct (lead set, station) pairs among the stepping cells, per cell column
#pragma unroll
                        for (int c = 0; c < DK3_NC; c++) {
                            const int ci = lane + 32 * c;
                            const int m = sm.qcnt[ci];
                            const bool stepping = lv[c] && m < DK2_QM;
                            unsigned grp = __ballot_sync(FULL, stepping) & __match_any_sync(FULL, kl[c]);
#pragma unroll
                            for (int q = 0; q < DK_WORDS; q++) grp &= __match_any_sync(FULL, sm.qmask[ci * DK_WORDS + q]);
                            const int lead_lane = __ffs(grp) - 1;
                            const unsigned leadm = __ballot_sync(FULL, stepping && lane == lead_lane);
                            if (stepping) {
                                const int g = ncombo + __popc(leadm & ((1u << lead_lane) - 1u));
                                sm.cid[ci] = g;
                                if (lane == lead_lane) {
                                    sm.kcb[g] = kl[c];
                                    sm.mcb[g] = m;
                                    sm.lcb[g] = ci;
                                }
                            }
                            ncombo += __popc(leadm);
                        }
                    }
                }
                if (lane == 0) {
                    sm.ctrl[1] = action;
                    sm.ctrl[2] = nc;
                    sm.ctrl[3] = ncombo;
                    sm.ctrl[4] = nstack;
                    sm.ctrl[5] = snap;
                    sm.ctrl[6] = since;
                }
            }
            __syncthreads();   // (2) decisions published
            DK3_T(2);
            const int action = sm.ctrl[1];
            nstack = sm.ctrl[4];
            if (action == 2) break;                                          // tile finished
            if (action == 1) {   // a waiting group: back to the inverse and chain it was split off from
                nstack--;
#pragma unroll
                for (int c = 0; c < DK3_NC; c++) jobmask[c] = (unsigned)sm.jobs[(DK3_NC + 1) * nstack + c];
                nmain = sm.jobs[(DK3_NC + 1) * nstack + DK3_NC];
                const double* sg = snap_g + (size_t)nstack * npk;
                for (int q = tid; q < npk; q += DK3_THREADS) Sp[q] = sg[q];
                for (int i = tid; i < DK3_NNP_MAX; i += DK3_THREADS) sm.mact[i] = (i < NN) ? 1 : 0;
                if (tid < DK_WORDS) {
                    const int lo = 32 * tid;
                    sm.chainw[tid] = n > lo ? (n - lo >= 32 ? FULL : ((1u << (n - lo)) - 1u)) : 0u;
                }
                if (tid < DK3_C) sm.kc[par][tid] = DK3_BIG;
                __syncthreads();
                for (int d = tid; d < nmain; d += DK3_THREADS) {
                    const int q = sm.kstack[d];
                    sm.mact[q] = 0;
                    atomicAnd(&sm.chainw[q >> 5], ~(1u << (q & 31)));
                }
                since = 0;
                if (tid == 0) sm.stats[1]++;
                __syncthreads();
                continue;
            }
#pragma unroll
            for (int c = 0; c < DK3_NC; c++) jobmask[c] = (unsigned)sm.ctrl[8 + c];
            since = sm.ctrl[6];
            if (sm.ctrl[5]) {   // split: the group left behind restarts from this inverse
                double* sg = snap_g + (size_t)(nstack - 1) * npk;
                for (int q = tid; q < npk; q += DK3_THREADS) sg[q] = Sp[q];
                if (tid == 0) sm.stats[2]++;
                __syncthreads();
            }
            if (action == 3) {   // ---- fold the common stations into the chain; the cells pick again afterwards (no drop in this step)
                const int nc = sm.ctrl[2];
                if (warp < QS) {   // current column q_d of the inverse = stored column corrected for the columns folded before it:
                    // x_d[j] = S(j, q_d) - sum_{e < d} (c_e[j] / mu_e) c_e[q_d].  One warp per 32 rows, the row's corrected values
                    // stay in registers; what the later columns need of column d -- its entries at their stations -- goes
                    // through shared memory (piv), one named barrier among these warps per column.  (One warp doing all rows
                    // column by column took 12 k clocks per fold; all columns in one warp's registers spilled the weights.)
                    const int j = lane + 32 * warp;
                    const int roj = j * NN - ((j * (j + 1)) >> 1);
                    const bool rowon = j < NN && sm.mact[j];
                    double c[DK3_FOLD];
#pragma unroll
                    for (int d = 0; d < DK3_FOLD; d++) {
                        c[d] = 0.0;
                        if (d < nc) {                                          // uniform
                            const int q = sm.clist[d];
                            bool on = rowon;
#pragma unroll
                            for (int e = 0; e < d; e++) on = on && (j != sm.clist[e]);
                            double x = 0.0;
                            if (on) {
                                x = (j <= q) ? Sp[roj + q] : Sp[sm.rowoff[q] + j];
#pragma unroll
                                for (int e = 0; e < d; e++) x = fma(-(c[e] * sm.prmu[e]), sm.piv[e * DK3_FOLD + d], x);
                            }
                            if (j == q) {
                                sm.prmu[d] = dk_rcp(x);
                                x = 0.0;
                            } else {
#pragma unroll
                                for (int d2 = d + 1; d2 < DK3_FOLD; d2++)
                                    if (d2 < nc && j == sm.clist[d2]) sm.piv[d * DK3_FOLD + d2] = x;
                            }
                            c[d] = x;
                            named_bar_sync(1, 32 * QS);
                        }
                    }
#pragma unroll
                    for (int d = 0; d < DK3_FOLD; d++)
                        if (d < nc && j < NNP) {
                            pcol[d * NNP + j] = c[d];
                            psm[d * NNP + j] = c[d] * sm.prmu[d];
                        }
                }
                __syncthreads();   // (F1) columns published
                DK3_T(3);
                if (warp == 1) {   // leads lose the stations folded in (lane = cell)
#pragma unroll
                    for (int c = 0; c < DK3_NC; c++) {
                        const int ci = lane + 32 * c;
                        if (((jobmask[c] >> lane) & 1u) && sm.clive[ci]) {
                            const int m = sm.qcnt[ci];
                            int keep = 0;
                            for (int i = 0; i < m; i++) {
                                const int qi = sm.qlist[ci * DK2_QM + i];
                                if (!((sm.taken[qi >> 5] >> (qi & 31)) & 1u)) sm.qlist[ci * DK2_QM + keep++] = qi;
                            }
                            sm.qcnt[ci] = keep;
#pragma unroll
                            for (int q = 0; q < DK_WORDS; q++) sm.qmask[ci * DK_WORDS + q] &= ~sm.taken[q];
                        }
                    }
                }
                // S -= sum_d col_d col_d' / mu_d on the packed triangle, DK2_FL stations per pass
                for (int d0 = 0; d0 < nc; d0 += DK2_FL) {
                    const int nd = min(DK2_FL, nc - d0);
                    const double* pc = pcol + d0 * NNP;
                    const double* ps = psm + d0 * NNP;
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        constexpr int QH = (QS + 1) / 2;
                        const int qa = half * QH, qb = half ? QS : QH;
                        double bj[DK2_FL][QH];                         // this lane's columns: col_d[j] / mu_d (zero beyond the set)
#pragma unroll
                        for (int d = 0; d < DK2_FL; d++)
#pragma unroll
                            for (int q = 0; q < QH; q++) bj[d][q] = (d < nd && qa + q < qb) ? ps[d * NNP + lane + 32 * (qa + q)] : 0.0;
                        const int iend = min(NN, 32 * qb);
                        for (int i = warp; i < iend; i += DK3_WARPS) {
                            if (!sm.mact[i]) continue;
                            double bi[DK2_FL];
#pragma unroll
                            for (int d = 0; d < DK2_FL; d++) bi[d] = d < nd ? pc[d * NNP + i] : 0.0;
                            double* row = Sp + sm.rowoff[i];
                            const int q0 = i >> 5;
#pragma unroll
                            for (int q = 0; q < QH; q++) {
                                const int j = lane + 32 * (qa + q);
                                if (qa + q < qb && qa + q >= q0 && j >= i && j < NN) {
                                    double v = row[j];
#pragma unroll
                                    for (int d = 0; d < DK2_FL; d++) v = fma(-bi[d], bj[d][q], v);
                                    row[j] = v;
                                }
                            }
                        }
                    }
                }
                __syncthreads();   // every reader of mact in the pass is done
                if (tid < nc) {
                    const int q = sm.clist[tid];
                    sm.kstack[nmain + tid] = q;
                    sm.mact[q] = 0;
                    atomicAnd(&sm.chainw[q >> 5], ~(1u << (q & 31)));
                }
                if (tid < DK3_C) sm.kc[par][tid] = DK3_BIG;                  // the cells pick again
                nmain += nc;
                if (tid == 0) sm.stats[3] += nc;
                __syncthreads();   // (F2) S, chain set, leads in place
                DK3_T(4);
                continue;
            }
            // ---- C: one warp per combo: y = (S[Q, Q])^-1 S[Q, k] (Gauss-Jordan in drop order, no pivoting), then the column
            //         col = S[:, k] - S[:, Q] y over the combo's active set; meanwhile every owner publishes w_k of its cells
            const int ncombo = sm.ctrl[3];
            DK3_Q0();
            if (tid == 0) sm.stats[4] += ncombo;
#pragma unroll
            for (int c = 0; c < DK3_NC; c++) {
                const int ci = lane + 32 * c;
                const int k = sm.cid[ci] >= 0 ? sm.kc[par][ci] : -1;
                const int s = k - j0;
                if (s >= 0 && s < DK3_SPW) {
                    double v = W[c][0];
#pragma unroll
                    for (int t = 1; t < DK3_SPW; t++)
                        if (s == t) v = W[c][t];
                    sm.wk[ci] = v;
                    act[c] &= ~(1u << s);
                }
            }
            DK3_Q(0);
            for (int g0 = 0; g0 < ncombo; g0 += DK2_CB) {
                const int ncb = min(DK2_CB, ncombo - g0);
                DK3_Q0();
                if (warp < ncb) {
                    const int L = sm.lcb[g0 + warp], k = sm.kcb[g0 + warp], m = sm.mcb[g0 + warp];
                    double* G = scr + warp * DK2_SCR;
                    const int myq = lane < m ? sm.qlist[L * DK2_QM + lane] : 0;
                    for (int e0 = 0; e0 < m * (m + 1); e0 += 32) {
                        const int e = e0 + lane;
                        const bool ok = e < m * (m + 1);
                        const int i = ok ? e / (m + 1) : 0, cc = ok ? e - i * (m + 1) : 0;
                        const int qi = __shfl_sync(FULL, myq, i), qc = __shfl_sync(FULL, myq, cc < m ? cc : 0);
                        if (ok) G[i * DK2_LD + cc] = Sat(qi, cc < m ? qc : k);
                    }
                    __syncwarp();
                    DK3_Q(1);
                    for (int p = 0; p < m; p++) {
                        const double r = dk_rcp(G[p * DK2_LD + p]);
                        const int wd = m - p;
                        for (int e = lane; e < m * wd; e += 32) {
                            const int i = e / wd, cc = p + 1 + (e - i * wd);
                            if (i != p) G[i * DK2_LD + cc] = fma(-(G[i * DK2_LD + p] * r), G[p * DK2_LD + cc], G[i * DK2_LD + cc]);
                        }
                        __syncwarp();
                    }
                    DK3_Q(2);
                    if (lane < m) {
                        sm.ycb[warp * DK2_QM + lane] = G[lane * DK2_LD + m] * dk_rcp(G[lane * DK2_LD + lane]);
                        sm.qcb[warp * DK2_QM + lane] = myq;
                    }
                    DK3_Q(3);
                }
                __syncthreads();   // (3a) lead systems solved
                for (int item = warp; item < ncb * QS; item += DK3_WARPS) {   // col = S[:, k] - S[:, Q] y, 32 rows per item
                    const int g = item / QS, sl = item - g * QS;
                    const int j = lane + 32 * sl;
                    const int L = sm.lcb[g0 + g], kk = sm.kcb[g0 + g], m = sm.mcb[g0 + g];
                    const int roj = j * NN - ((j * (j + 1)) >> 1);
                    double x = 0.0;
                    if (j < NN && sm.mact[j] && !((sm.qmask[L * DK_WORDS + (j >> 5)] >> (j & 31)) & 1u)) {
                        x = (j <= kk) ? Sp[roj + kk] : Sp[sm.rowoff[kk] + j];
                        for (int i = 0; i < m; i++) {
                            const int q = sm.qcb[g * DK2_QM + i];
                            x = fma(-((j <= q) ? Sp[roj + q] : Sp[sm.rowoff[q] + j]), sm.ycb[g * DK2_QM + i], x);
                        }
                    }
                    if (j == kk) {
                        sm.rbeta[g] = dk_rcp(x);
                        x = 0.0;
                    }
                    if (j < NNP) pcol[g * NNP + j] = x;
                }
                DK3_Q(4);
                __syncthreads();   // (3) columns built, w_k published
                DK3_Q(5);
                DK3_T(5);
                // ---- D: w -= (w_k / beta) col for every cell of the pass's combos
#pragma unroll
                for (int c = 0; c < DK3_NC; c++) {
                    const int ci = lane + 32 * c;
                    const int g = sm.cid[ci] - g0;
                    if (g >= 0 && g < ncb) {
                        const double f = sm.wk[ci] * sm.rbeta[g];
                        const double* col = pcol + g * NNP + j0;
#pragma unroll
                        for (int s = 0; s < DK3_SPW; s++)
                            if (j0 + s < NN) W[c][s] = fma(-f, col[s], W[c][s]);
                    }
                }
                if (warp == 0) {   // the lead grows by the station dropped (lane = cell)
#pragma unroll
                    for (int c = 0; c < DK3_NC; c++) {
                        const int ci = lane + 32 * c;
                        const int g = sm.cid[ci] - g0;
                        if (g >= 0 && g < ncb) {
                            const int k = sm.kc[par][ci], m = sm.qcnt[ci];
                            sm.qlist[ci * DK2_QM + m] = k;
                            sm.qcnt[ci] = m + 1;
                            sm.qmask[ci * DK_WORDS + (k >> 5)] |= 1u << (k & 31);
                        }
                    }
                }
                if (g0 + DK2_CB < ncombo) __syncthreads();   // the next pass reuses the column buffer
            }
            par ^= 1;
            DK3_T(6);
        }
    }
#ifdef DK3_PROFILE
    if (tid == 0)
        for (int i = 0; i < 8; i++) atomicAdd(a.tile_counter + 8 + i, (unsigned long long)pt[i]);
    if (tid == 0)
        for (int i = 0; i < 6; i++) atomicAdd(a.tile_counter + 16 + i, (unsigned long long)qt[i]);
#endif
    if (tid == 0) {
        for (int i = 0; i < 5; i++) atomicAdd(a.tile_counter + 2 + i, (unsigned long long)sm.stats[i]);
    }
}

// ------------------------------------------------------------------------------------------
// Exact reference solve on the final active set (krige.c:140-197 + lusolv.c), one thread per cell.
// ------------------------------------------------------------------------------------------
struct DkFinalArgs {
    int n, K;
    long long ncell;
    Geom g;
    const double* cx;        // [n] station x, COMPRESSED (station) order
    const double* cy;
    const double* elev;      // [n]
    const double* ad;        // [n][n]
    const double* dgrid;     // optional [ncell][n]
    const int* perm;         // rank -> compressed
    const int* rankof;       // compressed -> rank
    const int* orig;         // compressed -> original station index (column of the data rows)
    int mode;                // 0: finish the elimination exactly, write the final set back; 1: solve + write the ELL table
    uint32_t* final_act;
    uint8_t* widx;           // [K][ncell]
    double* wval;            // [K][ncell]
    int* status;             // bit0 singular, bit1 some cell has more than DK_NSMAX active stations (BIG pass needed)
    double* big_scratch;     // BIG variant: per thread (n+1)(n+2) + (n+1) doubles
    int* big_iscratch;       //              per thread 2(n+1) ints
    int big_threshold;       // DK_NSMAX (tests lower it through SMRFB_DK_BIG_THRESHOLD to exercise the BIG variant)
    unsigned long long* fallback_cells;
    int* maxcount_final;
};

// Crout LU with implicit scaling on a[(ns1) x (ns1+1)] (row stride lda), solving in place; no FMA.
__device__ int crout_solve_exact(double* a, int nn, int lda, int* indx, double* vv) {
    int imax = 0;
    for (int i = 0; i < nn; i++) {
        double big = 0.0;
        for (int j = 0; j < nn; j++) {
            double t = fabs(a[i * lda + j]);
            if (t > big) big = t;
        }
        if (big == 0.0) return 1;
        vv[i] = __ddiv_rn(1.0, big);
    }
    for (int j = 0; j < nn; j++) {
        for (int i = 0; i < j; i++) {
            double sum = a[i * lda + j];
            for (int k = 0; k < i; k++) sum = __dsub_rn(sum, __dmul_rn(a[i * lda + k], a[k * lda + j]));
            a[i * lda + j] = sum;
        }
        double big = 0.0;
        for (int i = j; i < nn; i++) {
            double sum = a[i * lda + j];
            for (int k = 0; k < j; k++) sum = __dsub_rn(sum, __dmul_rn(a[i * lda + k], a[k * lda + j]));
            a[i * lda + j] = sum;
            double dum = __dmul_rn(vv[i], fabs(sum));
            if (dum >= big) {
                big = dum;
                imax = i;
            }
        }
        if (j != imax) {
            for (int k = 0; k < nn; k++) {
                double t = a[imax * lda + k];
                a[imax * lda + k] = a[j * lda + k];
                a[j * lda + k] = t;
            }
            vv[imax] = vv[j];
        }
        indx[j] = imax;
        if (a[j * lda + j] == 0.0) return 1;
        if (j != nn - 1) {
            double inv = __ddiv_rn(1.0, a[j * lda + j]);
            for (int i = j + 1; i < nn; i++) a[i * lda + j] = __dmul_rn(a[i * lda + j], inv);
        }
    }
    int ii = -1;
    for (int i = 0; i < nn; i++) {
        int ip = indx[i];
        double sum = a[ip * lda + nn];
        a[ip * lda + nn] = a[i * lda + nn];
        if (ii >= 0) {
            for (int j = ii; j < i; j++) sum = __dsub_rn(sum, __dmul_rn(a[i * lda + j], a[j * lda + nn]));
        } else if (sum > 0.0) {
            ii = i;
        }
        a[i * lda + nn] = sum;
    }
    for (int i = nn - 1; i >= 0; i--) {
        double sum = a[i * lda + nn];
        for (int j = i + 1; j < nn; j++) sum = __dsub_rn(sum, __dmul_rn(a[i * lda + j], a[j * lda + nn]));
        a[i * lda + nn] = __ddiv_rn(sum, a[i * lda + i]);
    }
    return 0;
}

// BIG = false: thread-local workspace, active sets up to DK_NSMAX (every cell in practice: 8-11 stations survive at 200).
// BIG = true : the cells the first variant skipped (more than DK_NSMAX active stations when they leave the walk; the
//              reference has no such limit), workspace in global memory, a few thousand threads striding over the grid.
template <bool BIG>
__global__ void __launch_bounds__(128) dk_final_kernel(DkFinalArgs a) {
    constexpr int LDA_S = DK_NSMAX + 2;
    double A_s[BIG ? 1 : (DK_NSMAX + 1) * LDA_S];
    double vv_s[BIG ? 1 : DK_NSMAX + 1];
    int indx_s[BIG ? 1 : DK_NSMAX + 1];
    int ids_s[BIG ? 1 : DK_NSMAX + 1];
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gstride = BIG ? (long long)gridDim.x * blockDim.x : a.ncell;
    const int LDA = BIG ? a.n + 2 : LDA_S;
    const int cap = BIG ? a.n : DK_NSMAX;
    double* A = A_s;
    double* vv = vv_s;
    int* indx = indx_s;
    int* ids = ids_s;
    if (BIG) {
        const size_t per = (size_t)(a.n + 1) * (a.n + 2) + (a.n + 1);
        A = a.big_scratch + (size_t)gtid * per;
        vv = A + (size_t)(a.n + 1) * (a.n + 2);
        indx = a.big_iscratch + (size_t)gtid * 2 * (a.n + 1);
        ids = indx + (a.n + 1);
    }
    for (long long cell = gtid; cell < a.ncell; cell += gstride) {
    int ns = 0;
    bool handed = false;
    {
        int cnt = 0;
        for (int q = 0; q < DK_WORDS; q++) {
            unsigned b = a.final_act[(size_t)cell * DK_WORDS + q];
            if (q == DK_WORDS - 1) b &= 0x7fffffffu;
            cnt += __popc(b);
        }
        const bool big = cnt > a.big_threshold;
        if (big != BIG) {
            if (!BIG && a.mode == 0) atomicOr(a.status, 2);      // a pass of the BIG variant is needed
            continue;
        }
    }
    for (int q = 0; q < DK_WORDS; q++) {
        unsigned b = a.final_act[(size_t)cell * DK_WORDS + q];
        if (q == DK_WORDS - 1) {
            handed = (b & 0x80000000u) != 0;
            b &= 0x7fffffffu;
        }
        while (b) {
            int bit = __ffs(b) - 1;
            b &= b - 1;
            if (ns < cap) ids[ns] = a.perm[32 * q + bit];
            ns++;
        }
    }
    for (int i = 1; i < ns; i++) {   // station (compressed index) order, as the reference loops m = 0..nsta-1
        int v = ids[i], j = i - 1;
        while (j >= 0 && ids[j] > v) {
            ids[j + 1] = ids[j];
            j--;
        }
        ids[j + 1] = v;
    }
    double X = 0.0, Y = 0.0;
    if (!a.dgrid) cell_xy(a.g, cell, X, Y);
    bool fell_back = false, singular = false;
    for (;;) {
        for (int m = 0; m < ns; m++) {
            const int sm = ids[m];
            for (int q = 0; q < ns; q++) A[m * LDA + q] = a.ad[(size_t)sm * a.n + ids[q]];
            A[m * LDA + ns] = 1.0;
            A[ns * LDA + m] = 1.0;
            A[m * LDA + ns + 1] = a.dgrid ? a.dgrid[(size_t)cell * a.n + sm] : dist_rn(X, Y, a.cx[sm], a.cy[sm]);
        }
        A[ns * LDA + ns] = 0.0;
        A[ns * LDA + ns + 1] = 1.0;
        if (crout_solve_exact(A, ns + 1, LDA, indx, vv)) {
            atomicOr(a.status, 1);
            singular = true;
            break;
        }
        double esave = 0.0;   // krige.c:169-185
        int msave = -1;
        for (int m = 0; m < ns; m++) {
            if (A[m * LDA + ns + 1] < 0.0 && a.elev[ids[m]] > esave) {
                msave = m;
                esave = a.elev[ids[m]];
            }
        }
        if (msave < 0) break;
        fell_back = true;      // the downdated path stopped one step early: continue the reference loop exactly
        for (int m = msave; m + 1 < ns; m++) ids[m] = ids[m + 1];
        ns--;
    }
    if (singular) continue;
    if (a.mode == 0) {   // write the final active set back (rank-indexed bits) for the second pass
        if (fell_back && !handed) atomicAdd(a.fallback_cells, 1ULL);
        atomicMax(a.maxcount_final, ns);
        unsigned words[DK_WORDS];
        for (int q = 0; q < DK_WORDS; q++) words[q] = 0u;
        for (int j = 0; j < ns; j++) {
            const int r = a.rankof[ids[j]];
            words[r >> 5] |= 1u << (r & 31);
        }
        for (int q = 0; q < DK_WORDS; q++) a.final_act[(size_t)cell * DK_WORDS + q] = words[q];
        continue;
    }
    for (int j = 0; j < a.K; j++) {
        const bool on = j < ns;
        a.widx[(size_t)j * a.ncell + cell] = on ? (uint8_t)a.orig[ids[j]] : 0;
        a.wval[(size_t)j * a.ncell + cell] = on ? A[j * LDA + ns + 1] : 0.0;
    }
    }
}

// n <= DK_NSMAX: no shared walk at all, every cell starts the exact solver from the full station set
__global__ void dk_fill_all_kernel(long long ncell, int n, uint32_t* final_act) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncell * DK_WORDS) return;
    const int q = (int)(i % DK_WORDS);
    const int lo = 32 * q;
    unsigned w = 0u;
    if (n > lo) w = (n - lo >= 32) ? 0xffffffffu : ((1u << (n - lo)) - 1u);
    if (q == DK_WORDS - 1) w |= 0x80000000u;
    final_act[i] = w;
}

// dense expansion for smrfb_dk_weights / smrfb_krige_grid: out[cell][col], col = compressed index
__global__ void dk_expand_kernel(int K, long long ncell, int n, const uint8_t* widx, const double* wval,
                                 const int* orig2comp, double* out) {
    const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= ncell) return;
    for (int j = 0; j < K; j++) {
        double v = wval[(size_t)j * ncell + cell];
        if (v != 0.0) out[(size_t)cell * n + orig2comp[widx[(size_t)j * ncell + cell]]] = v;
    }
}

// ------------------------------------------------------------------------------------------
// Steady state (dk.py:86-96): out = sum_k w_k * r[station_k] + p0*Z + p1
// ------------------------------------------------------------------------------------------
// Step-pair records for one mask group: recp[p] = { r[t0][s], r[t1][s] interleaved over s | p0,p1 of t0 | p0,p1 of t1 }
// with (t0, t1) = (steps[2p], steps[2p+1]).  Interleaving lets the distribute kernel fetch the residuals of one
// station for two timesteps with a single 128-bit shared-memory read.
__global__ void dk_pair_kernel(const double* __restrict__ rec, RecLayout L, const int* __restrict__ steps, int Tm,
                               double* __restrict__ recp) {
    const int p = blockIdx.x;
    const int PR = 2 * L.S_pad + 4;
    const int t0 = steps[2 * p];
    const int t1 = (2 * p + 1 < Tm) ? steps[2 * p + 1] : -1;
    const double* r0 = rec + (size_t)t0 * L.RS;
    const double* r1 = t1 >= 0 ? rec + (size_t)t1 * L.RS : nullptr;
    double* o = recp + (size_t)p * PR;
    for (int s = threadIdx.x; s < L.S_pad; s += blockDim.x) {
        o[2 * s] = r0[s];
        o[2 * s + 1] = r1 ? r1[s] : 0.0;
    }
    if (threadIdx.x < 4) {
        const int h = threadIdx.x >> 1, q = threadIdx.x & 1;
        const double* r = h ? r1 : r0;
        o[2 * L.S_pad + threadIdx.x] = r ? r[L.meta() + q] : 0.0;
    }
}

// One CTA = 256 consecutive cells of the window, one thread per cell.  The thread keeps its cell's
// (station, weight) pairs in registers; the step-pair records of DK_TCH timesteps are staged in shared memory by
// one 1-D TMA bulk copy per block (completion on an mbarrier), double buffered so the copy of the next block
// overlaps the arithmetic of this one.  The per-cell gathers are 128-bit shared-memory reads (two timesteps each):
// cells of a warp mostly share their station set (broadcast), and different stations fall into different banks.
constexpr int DK_TCH = 16;
constexpr int DK_DT = 256;     // threads (= cells) per CTA of the distribute kernel

struct DkDistArgs {
    long long ncell;     // cells of the whole grid (ELL plane stride)
    long long cbase;     // first cell of the window
    long long ncw;       // cells in the window (out row stride)
    const double* dem;
    int K;
    const uint8_t* widx;
    const double* wval;
    const double* rec;   // prepared step records [T][RS] (generic kernel)
    const double* recp;  // step-pair records of this mask group [ceil(Tm/2)][2*S_pad + 4]
    RecLayout L;
    const int* steps;    // [Tm] step indices of this mask group
    int Tm;
    double vmin, vmax;
    double* out;         // [T][ncw]
    int out_f32;         // 1: out is float32
};

template <int KR>
__global__ void __launch_bounds__(DK_DT) dk_distribute_kernel(DkDistArgs a) {
    extern __shared__ __align__(16) unsigned char dk_smem[];
    const int SP2 = 2 * a.L.S_pad;
    const int PR = SP2 + 4;                                   // doubles per step-pair record (16-byte multiple)
    constexpr int NPR = DK_TCH / 2;                           // pair records per block
    double* res = reinterpret_cast<double*>(dk_smem);        // [2][NPR][PR]
    uint64_t* bars = reinterpret_cast<uint64_t*>(res + 2 * NPR * PR);   // [2]
    const int tid = threadIdx.x;
    const long long c = (long long)blockIdx.x * DK_DT + tid; // cell within the window
    const bool live = c < a.ncw;
    const size_t gc = (size_t)(a.cbase + (live ? c : 0));    // cell within the grid
    const int npair = (a.Tm + 1) / 2;
    const int nchunk = (npair + NPR - 1) / NPR;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int chn) {     // one thread: bulk-copy the pair records of block chn into buffer chn & 1
        const int b = chn & 1;
        const int nrow = min(NPR, npair - chn * NPR);
        const uint32_t bytes = (uint32_t)(nrow * PR * sizeof(double));
        mbar_arrive_expect_tx(&bars[b], bytes);
        bulk_g2s(res + (size_t)b * NPR * PR, a.recp + (size_t)chn * NPR * PR, bytes, &bars[b]);
    };
    if (tid == 0) {
        issue(0);
        if (nchunk > 1) issue(1);
    }
    double wv[KR];
    int wo[KR];
#pragma unroll
    for (int j = 0; j < KR; j++) {
        const bool on = live && j < a.K;
        wv[j] = on ? __ldg(a.wval + (size_t)j * a.ncell + gc) : 0.0;
        wo[j] = on ? 2 * (int)__ldg(a.widx + (size_t)j * a.ncell + gc) : 0;
    }
    const double Z = live ? __ldg(a.dem + gc) : 0.0;
    for (int chn = 0; chn < nchunk; chn++) {
        const int b = chn & 1;
        mbar_wait(&bars[b], (chn >> 1) & 1);
        const double* rb = res + (size_t)b * NPR * PR;
        if (live) {
            const int nrow = min(NPR, npair - chn * NPR);
#pragma unroll 4      // eight independent FMA chains per thread in flight: 415 -> 453 G cell-steps/s (unroll 2: four chains)
            for (int pr = 0; pr < nrow; pr++) {
                const double* rp = rb + (size_t)pr * PR;
                double r0 = 0.0, r1 = 0.0;
#pragma unroll
                for (int j = 0; j < KR; j++) {                // dk.py:91 (entries past K have weight 0)
                    const double2 x = *reinterpret_cast<const double2*>(rp + wo[j]);
                    r0 = fma(wv[j], x.x, r0);
                    r1 = fma(wv[j], x.y, r1);
                }
                const double2 pa = *reinterpret_cast<const double2*>(rp + SP2);
                const double2 pb = *reinterpret_cast<const double2*>(rp + SP2 + 2);
                const int tl = 2 * (chn * NPR + pr);
                const double v0 = __dadd_rn(__dadd_rn(r0, __dmul_rn(pa.x, Z)), pa.y);   // dk.py:168, left to right
                store_out(a.out, (size_t)__ldg(a.steps + tl) * a.ncw + c, clamp_nan_keep(v0, a.vmin, a.vmax), a.out_f32);
                if (tl + 1 < a.Tm) {
                    const double v1 = __dadd_rn(__dadd_rn(r1, __dmul_rn(pb.x, Z)), pb.y);
                    store_out(a.out, (size_t)__ldg(a.steps + tl + 1) * a.ncw + c, clamp_nan_keep(v1, a.vmin, a.vmax), a.out_f32);
                }
            }
        }
        __syncthreads();                                      // everyone is done with buffer b
        if (tid == 0 && chn + 2 < nchunk) issue(chn + 2);
    }
}

// ------------------------------------------------------------------------------------------
// dk_distribute_mma_kernel -- the steady state as a small dense product per warp.
// The 32 consecutive cells of a warp share almost all of their stations (union of the active sets over 32 cells of a
// C4 row: 6.6 on average, 8 at most away from the stations), so  out[t][c] = sum_u R[t][u] W[u][c]  with u over the
// union, plus two more columns for the trend (p0[t] * Z[c] + p1[t] * 1): an [8 steps x 4k] x [4k x 8 cells] product on
// the fp64 tensor cores (DMMA.8x8x4).  W stays in registers as B fragments for the whole launch.  The ELL kernel
// staged every step's WHOLE record (1.6 KB at 200 stations) in shared memory for 256 cells -- as many bytes from L2 as
// it wrote to HBM -- and gathered 16 bytes per station, cell and step pair from there (ncu: L1 81 % busy, 0.55 of the
// HBM peak).  Here the records of a mask group are first transposed to station-major rows (dk_transpose_kernel), and
// a lane reads its A-fragment element straight from them: 8 consecutive steps of one union column are one 64-byte
// run, shared through L1 by the warps around it; no staging, no block barrier in the step loop.  Warps whose union does
// not fit DKM_KS k-steps (near stations, set boundaries) walk their ELL lists against the same rows.
// ------------------------------------------------------------------------------------------
constexpr int DKM_KS = 4;                // k-steps of four columns kept in registers: union + 2 trend columns <= 16
constexpr int DKM_THREADS = 256;
constexpr int DKM_WARPS = DKM_THREADS / 32;
constexpr int DKM_WLD = 33;              // row stride of the per-warp dense weight scratch

// recT[u][i] = record column u (station residual, or p0 / p1 at S_pad, S_pad + 1) of the group's i-th step; TP >= Tm, zero padded
__global__ void dk_transpose_kernel(const double* __restrict__ rec, int RS, int ncol, const int* __restrict__ steps, int Tm, int TP,
                                    double* __restrict__ recT) {
    __shared__ double tile[32][33];
    const int i0 = blockIdx.x * 32, u0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, u = u0 + threadIdx.x;
        tile[r][threadIdx.x] = (i < Tm && u < ncol) ? rec[(size_t)steps[i] * RS + u] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int u = u0 + r, i = i0 + threadIdx.x;
        if (u < ncol && i < TP) recT[(size_t)u * TP + i] = tile[threadIdx.x][r];
    }
}

template <typename OutT>
__global__ void __launch_bounds__(DKM_THREADS, 3) dk_distribute_mma_kernel(DkDistArgs a, const double* __restrict__ recT, int TP) {
    __shared__ double Wd[DKM_WARPS * 4 * DKM_KS * DKM_WLD];                    // setup only: dense weights [warp][16][33]
    __shared__ unsigned bm[DKM_WARPS * 8];                                     // stations of the warp's union
    __shared__ int ulist[DKM_WARPS * 4 * DKM_KS];                              // the same as a list
    const int SP = a.L.S_pad;
    constexpr unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long c = (long long)blockIdx.x * DKM_THREADS + tid;             // cell within the window
    const bool live = c < a.ncw;
    const size_t gc = (size_t)(a.cbase + (live ? c : 0));

    // ---- setup: union of the warp's stations, dense weights, B fragments
    if (lane < 8) bm[warp * 8 + lane] = 0u;
    for (int i = lane; i < 4 * DKM_KS * DKM_WLD; i += 32) Wd[warp * (4 * DKM_KS) * DKM_WLD + i] = 0.0;
    __syncwarp();
    for (int j = 0; j < a.K; j++) {
        const double wv = live ? __ldg(a.wval + (size_t)j * a.ncell + gc) : 0.0;
        if (wv != 0.0) {
            const int s = __ldg(a.widx + (size_t)j * a.ncell + gc);
            atomicOr(&bm[warp * 8 + (s >> 5)], 1u << (s & 31));
        }
    }
    __syncwarp();
    int nu = 0;
    const unsigned mybm = lane < 8 ? bm[warp * 8 + lane] : 0u;
    int mypre = 0;                                                             // lane w < 8: stations of the union below word w
    {
        const int cnt = __popc(mybm);
        int run = cnt;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const int v = __shfl_up_sync(FULL, run, o);
            if (lane >= o) run += v;
        }
        mypre = run - cnt;
        nu = __shfl_sync(FULL, run, 7);
    }
    const bool fast = nu + 2 <= 4 * DKM_KS;                                    // warp-uniform
    const long long cw0 = (long long)blockIdx.x * DKM_THREADS + warp * 32;     // first cell of the warp
    if (cw0 >= a.ncw) return;
    if (fast) {
        double Bf[DKM_KS][4];
        const double* colA[DKM_KS];
        double* wd = Wd + warp * (4 * DKM_KS) * DKM_WLD;
        for (int j = 0; j < a.K; j++) {
            const double wv = live ? __ldg(a.wval + (size_t)j * a.ncell + gc) : 0.0;
            const int s = live ? __ldg(a.widx + (size_t)j * a.ncell + gc) : 0;
            const unsigned word = __shfl_sync(FULL, mybm, s >> 5);
            const int pre = __shfl_sync(FULL, mypre, s >> 5);
            if (wv != 0.0) {
                const int pos = pre + __popc(word & ((1u << (s & 31)) - 1u));
                wd[pos * DKM_WLD + lane] = wv;
                ulist[warp * 4 * DKM_KS + pos] = s;                            // same value from every cell that has the station
            }
        }
        if (live) {
            wd[nu * DKM_WLD + lane] = __ldg(a.dem + gc);                       // trend columns: p0[t] * Z + p1[t] * 1  (dk.py:168)
            wd[(nu + 1) * DKM_WLD + lane] = 1.0;
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < DKM_KS; s++) {
            const int u = 4 * s + (lane & 3);
            const int col = u < nu ? ulist[warp * 4 * DKM_KS + u] : (u == nu ? SP : (u == nu + 1 ? SP + 1 : 0));
            colA[s] = recT + (size_t)col * TP + (lane >> 2);                   // padding columns: weight 0 against a finite value
#pragma unroll
            for (int b = 0; b < 4; b++) Bf[s][b] = wd[u * DKM_WLD + 8 * b + (lane >> 2)];
        }
        const int ks = (nu + 2 + 3) >> 2;
        // a lane's outputs of one row block: step tr, cells cw0 + 8 nb + 2 (lane & 3) + {0, 1}, nb = 0 .. 3
        const bool interior = (cw0 + 32 <= a.ncw) && ((a.ncw & 1) == 0);      // every cell exists, rows stay 16-byte aligned
        OutT* const obase = reinterpret_cast<OutT*>(a.out) + cw0 + 2 * (lane & 3);
        const bool lo_on = a.vmin > -INFINITY, hi_on = a.vmax < INFINITY;
        double af[DKM_KS], an[DKM_KS];
#pragma unroll
        for (int s = 0; s < DKM_KS; s++) af[s] = s < ks ? __ldg(colA[s]) : 0.0;
        for (int t0 = 0; t0 < a.Tm; t0 += 8) {
            const int tr = t0 + (lane >> 2);                                   // this lane's step (position within the group)
            const bool rowok = tr < a.Tm;
#pragma unroll
            for (int s = 0; s < DKM_KS; s++) an[s] = (s < ks && t0 + 8 < TP) ? __ldg(colA[s] + t0 + 8) : 0.0;   // next row block
            OutT* const orow = obase + (rowok ? (size_t)__ldg(a.steps + tr) * a.ncw : 0);
#pragma unroll
            for (int nb = 0; nb < 4; nb++) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int s = 0; s < DKM_KS; s++)
                    if (s < ks) dmma884(c0, c1, af[s], Bf[s][nb]);
                if (lo_on) {                                                   // utils.py:137-161 (NaN fails both tests and stays)
                    c0 = (c0 <= a.vmin) ? a.vmin : c0;
                    c1 = (c1 <= a.vmin) ? a.vmin : c1;
                }
                if (hi_on) {
                    c0 = (c0 >= a.vmax) ? a.vmax : c0;
                    c1 = (c1 >= a.vmax) ? a.vmax : c1;
                }
                // (trading column blocks between the lanes of rows 0-3 and 4-7 so that every store instruction writes whole
                // 128-byte lines was measured neutral: 3.53 against 3.56 ms)
                if (interior) {
                    if (rowok) {
                        if constexpr (sizeof(OutT) == 4) __stcs(reinterpret_cast<float2*>(orow + 8 * nb), make_float2((float)c0, (float)c1));
                        else __stcs(reinterpret_cast<double2*>(orow + 8 * nb), make_double2(c0, c1));
                    }
                } else if (rowok) {
                    const long long cc = cw0 + 8 * nb + 2 * (lane & 3);
                    if (cc < a.ncw) __stcs(orow + 8 * nb, (OutT)c0);
                    if (cc + 1 < a.ncw) __stcs(orow + 8 * nb + 1, (OutT)c1);
                }
            }
#pragma unroll
            for (int s = 0; s < DKM_KS; s++) af[s] = an[s];
        }
    } else if (live) {   // large union: this cell's ELL list against the same rows
        const double Z = __ldg(a.dem + gc);
        for (int t = 0; t < a.Tm; t++) {
            double acc = 0.0;
            for (int j = 0; j < a.K; j++)
                acc = fma(__ldg(a.wval + (size_t)j * a.ncell + gc), __ldg(recT + (size_t)__ldg(a.widx + (size_t)j * a.ncell + gc) * TP + t), acc);
            const double v = __dadd_rn(__dadd_rn(acc, __dmul_rn(__ldg(recT + (size_t)SP * TP + t), Z)), __ldg(recT + (size_t)(SP + 1) * TP + t));
            store_out(a.out, (size_t)__ldg(a.steps + t) * a.ncw + c, clamp_nan_keep(v, a.vmin, a.vmax), a.out_f32);
        }
    }
}

// generic K (> DK_KREG): the ELL entries are re-read per step (L2)
__global__ void __launch_bounds__(256) dk_distribute_generic_kernel(DkDistArgs a) {
    const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
    if (c >= a.ncw) return;
    const size_t gc = (size_t)(a.cbase + c);
    const double Z = __ldg(a.dem + gc);
    for (int tl = 0; tl < a.Tm; tl++) {
        const int st = a.steps[tl];
        const double* rr = a.rec + (size_t)st * a.L.RS;
        double r = 0.0;
        for (int j = 0; j < a.K; j++)
            r = fma(__ldg(a.wval + (size_t)j * a.ncell + gc), __ldg(rr + __ldg(a.widx + (size_t)j * a.ncell + gc)), r);
        const double v = __dadd_rn(__dadd_rn(r, __dmul_rn(rr[a.L.meta()], Z)), rr[a.L.meta() + 1]);
        store_out(a.out, (size_t)st * a.ncw + c, clamp_nan_keep(v, a.vmin, a.vmax), a.out_f32);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct DkWeights {   // cached per NaN mask
    int n = 0, K = 0;
    uint8_t* widx = nullptr;
    double* wval = nullptr;
    std::vector<int> orig;   // compressed -> original station index
    size_t bytes = 0;
    ~DkWeights() {
        cudaFree(widx);
        cudaFree(wval);
    }
};

struct DkHandle : HandleBase {
    int slope_flag = 0;
    std::vector<double> h_mx, h_my, h_mz;
    RecLayout L;
    std::list<std::pair<std::vector<uint8_t>, DkWeights*>> cache;   // LRU, front = most recent
    size_t cache_bytes = 0, cache_budget = (size_t)48 << 30;
    long long masks_built = 0, fallback_cells = 0;
    int max_active = 0;
    int* d_steps = nullptr;
    size_t steps_cap = 0;
    double* d_recp = nullptr;   // step-pair records of the current mask group
    size_t recp_cap = 0;
    DkHandle() : HandleBase(HK_DK) {}
    ~DkHandle() override {
        cudaSetDevice(device);
        for (auto& e : cache) delete e.second;
        cudaFree(d_steps);
        cudaFree(d_recp);
    }
};

// Inverse of [[ad,1],[1',0]] (n+1 x n+1) by Gauss-Jordan with partial pivoting in long double on the
// equilibrated system (distances / dmax), returned in the ORIGINAL scaling.  false = singular.
static bool invert_kriging_system(int n, const std::vector<double>& ad, std::vector<double>& inv) {
    const int NN = n + 1;
    double dmax = 0.0;
    for (double v : ad) dmax = std::max(dmax, std::fabs(v));
    if (dmax == 0.0) dmax = 1.0;
    std::vector<long double> M((size_t)NN * 2 * NN, 0.0L);
    auto at = [&](int i, int j) -> long double& { return M[(size_t)i * 2 * NN + j]; };
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) at(i, j) = (long double)ad[(size_t)i * n + j] / dmax;
        at(i, n) = 1.0L;
        at(n, i) = 1.0L;
    }
    for (int i = 0; i < NN; i++) at(i, NN + i) = 1.0L;
    for (int c = 0; c < NN; c++) {
        int piv = c;
        long double best = fabsl(at(c, c));
        for (int r = c + 1; r < NN; r++)
            if (fabsl(at(r, c)) > best) {
                best = fabsl(at(r, c));
                piv = r;
            }
        if (best < 1e-14L) return false;
        if (piv != c)
            for (int j = 0; j < 2 * NN; j++) std::swap(at(c, j), at(piv, j));
        long double ip = 1.0L / at(c, c);
        for (int j = 0; j < 2 * NN; j++) at(c, j) *= ip;
        for (int r = 0; r < NN; r++) {
            if (r == c) continue;
            long double f = at(r, c);
            if (f == 0.0L) continue;
            for (int j = 0; j < 2 * NN; j++) at(r, j) -= f * at(c, j);
        }
    }
    // A = D A' D with D = diag(sqrt(dmax) x n, 1/sqrt(dmax))  =>  A^-1 = D^-1 A'^-1 D^-1
    inv.assign((size_t)NN * NN, 0.0);
    for (int i = 0; i < NN; i++)
        for (int j = 0; j < NN; j++) {
            long double v = at(i, NN + j);
            if (i < n && j < n) v /= dmax;
            else if (i == n && j == n) v *= dmax;
            inv[(size_t)i * NN + j] = (double)v;
        }
    return true;
}

struct DkBuildInput {
    int n;                         // valid stations
    const double* cx;              // [n] compressed-order coordinates (may be null in dgrid mode)
    const double* cy;
    const double* elev;            // [n]
    const double* ad;              // [n][n] or null (computed from cx, cy like dk.py:108-114)
    const double* d_dgrid;         // device [ncell][n] or null
    const int* orig;               // [n] compressed -> original column
    long long ncell;
    Geom g;
};

template <typename T>
static int dev_upload(T** dst, const T* src, size_t n) {
    SMRFB_CUDA(cudaMalloc((void**)dst, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) SMRFB_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return SMRFB_OK;
}

// Build the ELL weights for one station set. Synchronous.
static int dk_build(const DkBuildInput& in, cudaStream_t st, DkWeights* W, long long* fallback_cells) {
    const int n = in.n, NN = n + 1, NNP = (NN + 31) / 32 * 32;
    SMRFB_CHECK(n >= 1, SMRFB_ERR_NODATA, "DK: no valid station");
    SMRFB_CHECK(n <= DK3_NMAX, SMRFB_ERR_LIMIT, "DK supports at most %d valid stations (got %d)", DK3_NMAX, n);
    std::vector<double> ad((size_t)n * n, 0.0);
    if (in.ad) {
        std::copy(in.ad, in.ad + (size_t)n * n, ad.begin());
    } else {
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++) {
                double dx = in.cx[i] - in.cx[j], dy = in.cy[i] - in.cy[j];
                double d = std::sqrt(dx * dx + dy * dy);   // dk.py:113 (separate roundings; host code is built without FMA contraction)
                ad[(size_t)i * n + j] = ad[(size_t)j * n + i] = d;
            }
    }
    std::vector<double> inv;
    SMRFB_CHECK(invert_kriging_system(n, ad, inv), SMRFB_ERR_SINGULAR,
                "DK: singular kriging system (co-located stations?)");
    // rank order: elevation descending, index ascending (krige.c:169-185 picks the first such station)
    std::vector<int> perm(n);
    for (int i = 0; i < n; i++) perm[i] = i;
    std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) { return in.elev[x] > in.elev[y]; });
    int nsel = 0;
    for (int r = 0; r < n; r++)
        if (in.elev[perm[r]] > 0.0) nsel = r + 1;   // elevations <= 0 sort last and can never be dropped
    std::vector<int> pfull(NN);
    for (int r = 0; r < n; r++) pfull[r] = perm[r];
    pfull[n] = n;
    const int npk = NN * (NN + 1) / 2;
    std::vector<double> B0p(npk);
    for (int i = 0; i < NN; i++)
        for (int j = i; j < NN; j++) {
            // symmetrise: the exact inverse is symmetric
            double v = 0.5 * (inv[(size_t)pfull[i] * NN + pfull[j]] + inv[(size_t)pfull[j] * NN + pfull[i]]);
            B0p[(size_t)i * NN - (size_t)i * (i + 1) / 2 + j] = v;
        }
    std::vector<double> sx(n), sy(n);
    if (in.cx)
        for (int r = 0; r < n; r++) {
            sx[r] = in.cx[perm[r]];
            sy[r] = in.cy[perm[r]];
        }
    int dev = 0, nsm = 0, smem_max = 0;
    SMRFB_CUDA(cudaGetDevice(&dev));
    SMRFB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    SMRFB_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    // SMRFB_DK_WALK: sets (default; transposed weights, 64-cell tiles) | chain (one warp per cell, 32-cell tiles) | tree (round 1:
    // path-keyed walk with forks) -- the older designs are kept for comparison
    const char* walk_env = getenv("SMRFB_DK_WALK");
    const int walk = (walk_env && strcmp(walk_env, "tree") == 0) ? 0 : (walk_env && strcmp(walk_env, "chain") == 0) ? 1 : 2;
    SMRFB_CHECK(walk == 2 || n <= DK_MAXN, SMRFB_ERR_LIMIT, "SMRFB_DK_WALK=%s supports at most %d valid stations (got %d)", walk_env, DK_MAXN, n);
    const bool sglob = walk == 2 && n > DK_MAXN;   // the packed inverse (n + 1)(n + 2) / 2 doubles no longer fits shared memory
    const bool chain = walk != 0;
    const long long ntiles = walk == 2 ? dk3_tiling(in.g, in.ncell, in.d_dgrid != nullptr).ntiles
                                       : dk_tiling(in.g, in.ncell, in.d_dgrid != nullptr).ntiles;
    const int grid = (int)std::min<long long>(ntiles, nsm);
    const int maxsnap = walk == 2 ? DK3_MAXSNAP : DK_MAXSNAP;
    size_t smem = walk == 2 ? dk3_smem_bytes(NN, NNP, sglob)
                  : walk == 1 ? dk2_smem_bytes(NN, NNP)
                              : ((size_t)npk + (size_t)DK_TILE * NN + 2 * DK_PEND * NNP) * sizeof(double) +
                                    ((size_t)NNP + NN + DK_TILE + 8) * sizeof(int) + 64 * sizeof(DkPending) + NNP;
    SMRFB_CHECK(smem <= (size_t)smem_max, SMRFB_ERR_LIMIT, "DK: %d stations need %zu B shared memory (max %d)", n, smem, smem_max);

    double *d_sx = nullptr, *d_sy = nullptr, *d_cx = nullptr, *d_cy = nullptr, *d_elev = nullptr, *d_ad = nullptr,
           *d_B0p = nullptr, *d_undo = nullptr, *d_big = nullptr, *d_sglob = nullptr;
    int *d_perm = nullptr, *d_orig = nullptr, *d_small = nullptr, *d_rankof = nullptr, *d_ibig = nullptr;
    uint32_t* d_final = nullptr;
    unsigned long long* d_cnt = nullptr;
    int rc = SMRFB_OK;
    auto cleanup = [&]() {
        cudaFree(d_sx); cudaFree(d_sy); cudaFree(d_cx); cudaFree(d_cy); cudaFree(d_elev); cudaFree(d_ad);
        cudaFree(d_B0p); cudaFree(d_undo); cudaFree(d_perm); cudaFree(d_rankof); cudaFree(d_orig); cudaFree(d_small); cudaFree(d_final);
        cudaFree(d_cnt); cudaFree(d_big); cudaFree(d_ibig); cudaFree(d_sglob);
    };
#define DK_TRY(x)                 \
    do {                          \
        rc = (x);                 \
        if (rc != SMRFB_OK) {     \
            cleanup();            \
            return rc;            \
        }                         \
    } while (0)
#define DK_CUDA(x)                                                                   \
    do {                                                                             \
        cudaError_t e__ = (x);                                                       \
        if (e__ != cudaSuccess) {                                                    \
            set_error("%s: %s (%s:%d)", #x, cudaGetErrorString(e__), __FILE__, __LINE__); \
            cleanup();                                                               \
            return SMRFB_ERR_CUDA;                                                   \
        }                                                                            \
    } while (0)
    DK_TRY(dev_upload(&d_sx, sx.data(), n));
    DK_TRY(dev_upload(&d_sy, sy.data(), n));
    if (in.cx) {
        DK_TRY(dev_upload(&d_cx, in.cx, n));
        DK_TRY(dev_upload(&d_cy, in.cy, n));
    }
    DK_TRY(dev_upload(&d_elev, in.elev, n));
    DK_TRY(dev_upload(&d_ad, ad.data(), (size_t)n * n));
    DK_TRY(dev_upload(&d_B0p, B0p.data(), npk));
    DK_TRY(dev_upload(&d_perm, perm.data(), n));
    {
        std::vector<int> rankof(n);
        for (int r = 0; r < n; r++) rankof[perm[r]] = r;
        DK_TRY(dev_upload(&d_rankof, rankof.data(), n));
    }
    DK_TRY(dev_upload(&d_orig, in.orig, n));
    if (n > DK_NSMAX) DK_CUDA(cudaMalloc((void**)&d_undo, (size_t)grid * maxsnap * npk * sizeof(double)));
    DK_CUDA(cudaMalloc((void**)&d_final, (size_t)in.ncell * DK_WORDS * sizeof(uint32_t)));
    DK_CUDA(cudaMalloc((void**)&d_cnt, 24 * sizeof(unsigned long long)));
    DK_CUDA(cudaMalloc((void**)&d_small, 4 * sizeof(int)));
    DK_CUDA(cudaMemsetAsync(d_cnt, 0, 24 * sizeof(unsigned long long), st));
    DK_CUDA(cudaMemsetAsync(d_small, 0, 4 * sizeof(int), st));

    DkBuildArgs ba;
    ba.n = n; ba.NN = NN; ba.NNP = NNP; ba.nsel = nsel;
    ba.ncell = in.ncell; ba.g = in.g;
    ba.sx = d_sx; ba.sy = d_sy; ba.dgrid = in.d_dgrid; ba.perm = d_perm; ba.B0p = d_B0p; ba.snap = d_undo;
    ba.final_act = d_final; ba.tile_counter = d_cnt; ba.maxcount = d_small; ba.sglob = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    const bool verbose = getenv("SMRFB_DK_VERBOSE") != nullptr;
    if (verbose) {
        cudaEventCreate(&ev0); cudaEventCreate(&ev1); cudaEventCreate(&ev2);
        cudaEventRecord(ev0, st);
    }
    ba.handoff = 10;   // below ~10 active stations a per-thread exact solve is cheaper than a CTA-wide rank-1 step
    if (n > DK_NSMAX) {
        if (walk == 2 && sglob) {
            DK_CUDA(cudaMalloc((void**)&d_sglob, (size_t)grid * npk * sizeof(double)));
            ba.sglob = d_sglob;
            DK_CUDA(cudaFuncSetAttribute(dk_sets_kernel<16, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dk_sets_kernel<16, 8, true><<<grid, DK3_THREADS, smem, st>>>(ba);
        } else if (walk == 2) {
            DK_CUDA(cudaFuncSetAttribute(dk_sets_kernel<14, 7, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dk_sets_kernel<14, 7, false><<<grid, DK3_THREADS, smem, st>>>(ba);
        } else if (walk == 1) {
            DK_CUDA(cudaFuncSetAttribute(dk_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dk_chain_kernel<<<grid, DK_THREADS, smem, st>>>(ba);
        } else {
            DK_CUDA(cudaFuncSetAttribute(dk_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dk_build_kernel<<<grid, DK_THREADS, smem, st>>>(ba);
        }
    } else {   // small station sets: the exact per-cell solver does the whole elimination (pure reference arithmetic)
        const long long nw = in.ncell * DK_WORDS;
        dk_fill_all_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(in.ncell, n, d_final);
    }
    count_launch();
    DK_CUDA(cudaGetLastError());
    if (verbose) cudaEventRecord(ev1, st);
    DkFinalArgs fa;
    fa.n = n; fa.K = 0; fa.ncell = in.ncell; fa.g = in.g;
    fa.cx = d_cx; fa.cy = d_cy; fa.elev = d_elev; fa.ad = d_ad; fa.dgrid = in.d_dgrid; fa.perm = d_perm; fa.rankof = d_rankof;
    fa.orig = d_orig; fa.final_act = d_final; fa.widx = nullptr; fa.wval = nullptr; fa.status = d_small + 1;
    fa.fallback_cells = d_cnt + 1; fa.maxcount_final = d_small + 2;
    fa.mode = 0;
    fa.big_scratch = nullptr;
    fa.big_iscratch = nullptr;
    fa.big_threshold = DK_NSMAX;
    if (const char* e = getenv("SMRFB_DK_BIG_THRESHOLD")) fa.big_threshold = std::min(DK_NSMAX, std::max(1, atoi(e)));
    dk_final_kernel<false><<<(unsigned)((in.ncell + 127) / 128), 128, 0, st>>>(fa);
    count_launch();
    DK_CUDA(cudaGetLastError());
    int h_small[4] = {0, 0, 0, 0};
    DK_CUDA(cudaMemcpyAsync(h_small, d_small, sizeof(h_small), cudaMemcpyDeviceToHost, st));
    DK_CUDA(cudaStreamSynchronize(st));
    constexpr int BIG_BLOCKS = 16;       // x 128 threads, each with its own (n+1)(n+2) workspace in global memory
    if (h_small[1] & 2) {                // rare: cells that left the walk with more than DK_NSMAX active stations
        const size_t per = (size_t)(n + 1) * (n + 2) + (n + 1);
        DK_CUDA(cudaMalloc((void**)&d_big, (size_t)BIG_BLOCKS * 128 * per * sizeof(double)));
        DK_CUDA(cudaMalloc((void**)&d_ibig, (size_t)BIG_BLOCKS * 128 * 2 * (n + 1) * sizeof(int)));
        fa.big_scratch = d_big;
        fa.big_iscratch = d_ibig;
        dk_final_kernel<true><<<BIG_BLOCKS, 128, 0, st>>>(fa);
        count_launch();
        DK_CUDA(cudaGetLastError());
        DK_CUDA(cudaMemcpyAsync(h_small, d_small, sizeof(h_small), cudaMemcpyDeviceToHost, st));
        DK_CUDA(cudaStreamSynchronize(st));
    }
    if (verbose) cudaEventRecord(ev2, st);
    if (verbose) cudaEventSynchronize(ev2);
    float ms_walk = 0.f, ms_final = 0.f;
    if (verbose) {
        cudaEventElapsedTime(&ms_walk, ev0, ev1);
        cudaEventElapsedTime(&ms_final, ev1, ev2);
        cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
    }
    if (h_small[1] & 1) {
        set_error("DK: singular kriging system (co-located stations?)");
        cleanup();
        return SMRFB_ERR_SINGULAR;
    }
    const int K = std::max(1, h_small[2]);
    W->n = n;
    W->K = K;
    W->orig.assign(in.orig, in.orig + n);
    W->bytes = (size_t)K * in.ncell * (sizeof(double) + 1);
    DK_CUDA(cudaMalloc((void**)&W->widx, (size_t)K * in.ncell));
    DK_CUDA(cudaMalloc((void**)&W->wval, (size_t)K * in.ncell * sizeof(double)));
    fa.K = K; fa.widx = W->widx; fa.wval = W->wval;
    fa.mode = 1;
    dk_final_kernel<false><<<(unsigned)((in.ncell + 127) / 128), 128, 0, st>>>(fa);
    count_launch();
    DK_CUDA(cudaGetLastError());
    if (K > fa.big_threshold) {
        dk_final_kernel<true><<<BIG_BLOCKS, 128, 0, st>>>(fa);
        count_launch();
        DK_CUDA(cudaGetLastError());
    }
    unsigned long long h_cnt[24] = {0};
    DK_CUDA(cudaMemcpyAsync(h_small, d_small, sizeof(h_small), cudaMemcpyDeviceToHost, st));
    DK_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
    DK_CUDA(cudaStreamSynchronize(st));
    cleanup();
#undef DK_TRY
#undef DK_CUDA
    SMRFB_CHECK(!(h_small[1] & 1), SMRFB_ERR_SINGULAR, "DK: singular kriging system in the final solve");
    if (fallback_cells) *fallback_cells += (long long)h_cnt[1];
#ifdef DK3_PROFILE
    if (verbose && walk == 2) {
        double tot = 0;
        for (int i = 0; i < 8; i++) tot += (double)h_cnt[8 + i];
        fprintf(stderr, "[smrfb] dk_sets phases (share of thread 0's clocks): init %.1f%%  pick %.1f%%  book %.1f%%  fold-extract %.1f%%  "
                "fold-flush %.1f%%  solve+column %.1f%%  apply %.1f%%;  clocks per step %.0f\n", 100 * h_cnt[8] / tot, 100 * h_cnt[9] / tot,
                100 * h_cnt[10] / tot, 100 * h_cnt[11] / tot, 100 * h_cnt[12] / tot, 100 * h_cnt[13] / tot, 100 * h_cnt[14] / tot,
                (tot - h_cnt[8]) / std::max(1.0, (double)h_cnt[2]));
        double q = 0;
        for (int i = 0; i < 6; i++) q += (double)h_cnt[16 + i];
        fprintf(stderr, "[smrfb] dk_sets solve phase: publish w_k %.1f%%  fill %.1f%%  Gauss-Jordan %.1f%%  y %.1f%%  column %.1f%%  barrier %.1f%%\n",
                100 * h_cnt[16] / q, 100 * h_cnt[17] / q, 100 * h_cnt[18] / q, 100 * h_cnt[19] / q, 100 * h_cnt[20] / q, 100 * h_cnt[21] / q);
    }
#endif
    if (verbose)
        fprintf(stderr, "[smrfb] dk_build (%s): n=%d cells=%lld tiles=%lld %s=%llu restores=%llu %s=%llu folded=%llu combos=%llu K=%d "
                "fallback=%llu walk %.1f ms, exact finish %.1f ms\n", walk == 2 ? "sets" : chain ? "chain" : "tree", n, in.ncell, ntiles, chain ? "steps" : "applies",
                h_cnt[2], h_cnt[3], chain ? "splits" : "forks", h_cnt[4], h_cnt[5], h_cnt[6], K, h_cnt[1], ms_walk, ms_final);

    return SMRFB_OK;
}

static int dk_get_weights(DkHandle* h, const std::vector<uint8_t>& nanmask, DkWeights** out) {
    for (auto it = h->cache.begin(); it != h->cache.end(); ++it)
        if (it->first == nanmask) {
            h->cache.splice(h->cache.begin(), h->cache, it);
            *out = h->cache.front().second;
            return SMRFB_OK;
        }
    std::vector<double> cx, cy, cz;
    std::vector<int> orig;
    for (int s = 0; s < h->S; s++)
        if (!nanmask[s]) {
            cx.push_back(h->h_mx[s]);
            cy.push_back(h->h_my[s]);
            cz.push_back(h->h_mz[s]);
            orig.push_back(s);
        }
    DkBuildInput in;
    in.n = (int)cx.size();
    in.cx = cx.data(); in.cy = cy.data(); in.elev = cz.data(); in.ad = nullptr; in.d_dgrid = nullptr;
    in.orig = orig.data(); in.ncell = h->g.ncell; in.g = h->g;
    auto evict_lru = [&]() {
        h->cache_bytes -= h->cache.back().second->bytes;
        delete h->cache.back().second;
        h->cache.pop_back();
    };
    // make room BEFORE building (one mask is 9-10 GB at 1e8 cells): estimate the new table with the largest active
    // set seen so far (12 until a mask has been built), then evict least-recently-used masks and retry if an
    // allocation still fails
    const size_t est = (size_t)std::max(h->max_active, 12) * (size_t)h->g.ncell * (sizeof(double) + 1);
    while (!h->cache.empty() && h->cache_bytes + est > h->cache_budget) evict_lru();
    DkWeights* W = nullptr;
    int rc;
    for (;;) {
        W = new DkWeights();
        rc = dk_build(in, h->stream, W, &h->fallback_cells);
        if (rc == SMRFB_OK) break;
        delete W;
        if (rc != SMRFB_ERR_CUDA || h->cache.empty()) return rc;
        cudaGetLastError();   // clear the (non-sticky) allocation failure; a sticky error fails again and ends the loop
        evict_lru();
    }
    h->masks_built++;
    h->max_active = std::max(h->max_active, W->K);
    while (!h->cache.empty() && h->cache_bytes + W->bytes > h->cache_budget) evict_lru();
    h->cache.emplace_front(nanmask, W);
    h->cache_bytes += W->bytes;
    *out = W;
    return SMRFB_OK;
}

static int dk_run(DkHandle* h, const double* d_data, const double* h_data, int T, double vmin, double vmax,
                  const double* d_pv_in, double* d_pv_out, double* d_out, cudaStream_t st, long long cbase = 0,
                  long long ncw = -1) {
    if (ncw < 0) ncw = h->g.ncell;
    SMRFB_CHECK(T > 0, SMRFB_ERR_ARG, "T must be > 0");
    SMRFB_CHECK(cbase >= 0 && ncw > 0 && cbase + ncw <= h->g.ncell, SMRFB_ERR_ARG, "cell window [%lld, +%lld) outside the grid",
                cbase, ncw);
    const int S = h->S;
    // group the steps by NaN mask, in order of first appearance (dk.py:75-85 recomputes on every change)
    std::vector<std::vector<uint8_t>> masks;
    std::vector<std::vector<int>> groups;
    {
        std::map<std::vector<uint8_t>, int> idx;
        std::vector<uint8_t> m(S);
        for (int t = 0; t < T; t++) {
            int nv = 0;
            for (int s = 0; s < S; s++) {
                m[s] = std::isnan(h_data[(size_t)t * S + s]) ? 1 : 0;
                nv += !m[s];
            }
            SMRFB_CHECK(nv > 0, SMRFB_ERR_NODATA, "timestep %d has no valid station (all values NaN)", t);
            auto it = idx.find(m);
            if (it == idx.end()) {
                idx[m] = (int)masks.size();
                masks.push_back(m);
                groups.emplace_back(1, t);
            } else {
                groups[it->second].push_back(t);
            }
        }
    }
    const int T_pad = T;
    SMRFB_PROPAGATE(ensure_cap((void**)&h->d_rec, &h->rec_cap, (size_t)T_pad * h->L.RS * sizeof(double)));
    SMRFB_PROPAGATE(launch_prep(st, d_data, T, T_pad, S, h->L, h->d_mz, nullptr, 1, h->slope_flag, d_pv_in, d_pv_out,
                                h->d_rec, nullptr, h->d_flags));
    for (size_t gi = 0; gi < groups.size(); gi++) {
        DkWeights* W = nullptr;
        SMRFB_PROPAGATE(dk_get_weights(h, masks[gi], &W));
        const int Tm = (int)groups[gi].size();
        // the step-index buffer is reused by the next group: wait for the previous group's kernel
        SMRFB_CUDA(cudaStreamSynchronize(st));
        SMRFB_PROPAGATE(ensure_cap((void**)&h->d_steps, &h->steps_cap, (size_t)Tm * sizeof(int)));
        SMRFB_CUDA(cudaMemcpyAsync(h->d_steps, groups[gi].data(), (size_t)Tm * sizeof(int), cudaMemcpyHostToDevice, st));
        DkDistArgs a;
        a.ncell = h->g.ncell; a.cbase = cbase; a.ncw = ncw; a.dem = h->g.dem; a.K = W->K; a.widx = W->widx; a.wval = W->wval;
        a.rec = h->d_rec; a.recp = nullptr; a.L = h->L; a.steps = h->d_steps; a.Tm = Tm;
        a.vmin = vmin; a.vmax = vmax; a.out = d_out; a.out_f32 = h->out_f32;
        // meshes: neighbouring cells share their stations -> the tensor-core kernel; SMRFB_DK_STEADY=ell|mma overrides
        const char* steady_env = getenv("SMRFB_DK_STEADY");
        const bool use_mma = steady_env ? strcmp(steady_env, "mma") == 0 : h->g.kind == 0;
        if (use_mma) {
            const int ncol = h->L.S_pad + 2, TP = (Tm + 15) / 16 * 16;
            SMRFB_PROPAGATE(ensure_cap((void**)&h->d_recp, &h->recp_cap, (size_t)ncol * TP * sizeof(double)));
            dk_transpose_kernel<<<dim3((TP + 31) / 32, (ncol + 31) / 32), dim3(32, 8), 0, st>>>(h->d_rec, h->L.RS, ncol, h->d_steps, Tm, TP,
                                                                                             h->d_recp);
            if (h->out_f32) dk_distribute_mma_kernel<float><<<(unsigned)((ncw + DKM_THREADS - 1) / DKM_THREADS), DKM_THREADS, 0, st>>>(a, h->d_recp, TP);
            else dk_distribute_mma_kernel<double><<<(unsigned)((ncw + DKM_THREADS - 1) / DKM_THREADS), DKM_THREADS, 0, st>>>(a, h->d_recp, TP);
            count_launch(2);
            SMRFB_CUDA(cudaGetLastError());
            continue;
        }
        const int npair = (Tm + 1) / 2;
        const int PR = 2 * h->L.S_pad + 4;
        SMRFB_PROPAGATE(ensure_cap((void**)&h->d_recp, &h->recp_cap, (size_t)npair * PR * sizeof(double)));
        dk_pair_kernel<<<npair, 128, 0, st>>>(h->d_rec, h->L, h->d_steps, Tm, h->d_recp);
        count_launch();
        SMRFB_CUDA(cudaGetLastError());
        a.recp = h->d_recp;
        const unsigned nblk = (unsigned)((ncw + 255) / 256);
        const unsigned nblk_d = (unsigned)((ncw + DK_DT - 1) / DK_DT);
        const size_t dsm = (size_t)2 * (DK_TCH / 2) * PR * sizeof(double) + 2 * sizeof(uint64_t);
#define DK_LAUNCH(KRV)                                                                                           \
    do {                                                                                                         \
        SMRFB_CUDA(cudaFuncSetAttribute(dk_distribute_kernel<KRV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm)); \
        dk_distribute_kernel<KRV><<<nblk_d, DK_DT, dsm, st>>>(a);                                                 \
    } while (0)
        switch ((W->K + 1) / 2) {   // register-resident (station, weight) pairs, padded to an even count
            case 1: DK_LAUNCH(2); break;
            case 2: DK_LAUNCH(4); break;
            case 3: DK_LAUNCH(6); break;
            case 4: DK_LAUNCH(8); break;
            case 5: DK_LAUNCH(10); break;
            case 6: DK_LAUNCH(12); break;
            case 7: DK_LAUNCH(14); break;
            case 8: DK_LAUNCH(16); break;
            default: dk_distribute_generic_kernel<<<nblk, 256, 0, st>>>(a);
        }
#undef DK_LAUNCH
        count_launch();
        SMRFB_CUDA(cudaGetLastError());
    }
    return SMRFB_OK;
}

}  // namespace smrfb

using namespace smrfb;

extern "C" {

int smrfb_dk_create(int nsta, const double* mx, const double* my, const double* mz, int ny, int nx, int grid_kind,
                    const double* gx, const double* gy, const double* dem, int slope_flag, int device,
                    smrfb_handle_t* out) {
    SMRFB_CHECK(out, SMRFB_ERR_ARG, "out handle pointer is NULL");
    *out = nullptr;
    SMRFB_CHECK(mz && dem, SMRFB_ERR_ARG, "DK needs station elevations and the DEM");
    SMRFB_CHECK(nsta <= 255, SMRFB_ERR_LIMIT, "DK supports at most 255 stations per handle (got %d)", nsta);
    DkHandle* h = new DkHandle();
    h->device = device;
    h->slope_flag = slope_flag;
    int rc = upload_geom(h, nsta, mx, my, mz, ny, nx, grid_kind, gx, gy, dem);
    if (rc != SMRFB_OK) {
        delete h;
        return rc;
    }
    h->h_mx.assign(mx, mx + nsta);
    h->h_my.assign(my, my + nsta);
    h->h_mz.assign(mz, mz + nsta);
    h->L = make_layout(nsta, 2, 0, false);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) h->cache_budget = free_b / 2;
    *out = reinterpret_cast<smrfb_handle_t>(h);
    return SMRFB_OK;
}

int smrfb_dk_distribute_dev(smrfb_handle_t hh, const double* d_data, const double* h_data, int T, double vmin,
                            double vmax, const double* d_pv_in, double* d_pv_out, double* d_out, void* stream) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    DkHandle* h = reinterpret_cast<DkHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_DK, SMRFB_ERR_ARG, "not a DK handle");
    SMRFB_CHECK(d_data && h_data && d_out, SMRFB_ERR_ARG, "NULL data/out");
    SMRFB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    return dk_run(h, d_data, h_data, T, vmin, vmax, d_pv_in, d_pv_out, d_out, st);
}

int smrfb_dk_distribute_window_dev(smrfb_handle_t hh, const double* d_data, const double* h_data, int T, double vmin,
                                   double vmax, const double* d_pv_in, double* d_pv_out, int64_t cell_begin,
                                   int64_t cell_count, double* d_out, void* stream) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    DkHandle* h = reinterpret_cast<DkHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_DK, SMRFB_ERR_ARG, "not a DK handle");
    SMRFB_CHECK(d_data && h_data && d_out, SMRFB_ERR_ARG, "NULL data/out");
    SMRFB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    return dk_run(h, d_data, h_data, T, vmin, vmax, d_pv_in, d_pv_out, d_out, st, cell_begin, cell_count);
}

int smrfb_dk_distribute(smrfb_handle_t hh, const double* data, int T, double vmin, double vmax, const double* pv_in,
                        double* pv_out, double* out) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    DkHandle* h = reinterpret_cast<DkHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_DK, SMRFB_ERR_ARG, "not a DK handle");
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
        return dk_run(h, h->d_data + (size_t)t0 * S, data + (size_t)t0 * S, tn, vmin, vmax,
                      d_pvin ? d_pvin + (size_t)2 * t0 : nullptr, h->d_pv + (size_t)2 * t0, d_out, st);
    });
    if (rc == SMRFB_OK && pv_out &&
        cudaMemcpy(pv_out, h->d_pv, (size_t)T * 2 * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) {
        set_error("DK: reading back coefficients failed");
        rc = SMRFB_ERR_CUDA;
    }
    return rc;
}

int smrfb_dk_weights(smrfb_handle_t hh, const uint8_t* nan_mask, double* weights_out) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    DkHandle* h = reinterpret_cast<DkHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_DK, SMRFB_ERR_ARG, "not a DK handle");
    SMRFB_CHECK(nan_mask && weights_out, SMRFB_ERR_ARG, "NULL mask/out");
    SMRFB_CUDA(cudaSetDevice(h->device));
    std::vector<uint8_t> m(nan_mask, nan_mask + h->S);
    for (auto& v : m) v = v ? 1 : 0;
    DkWeights* W = nullptr;
    SMRFB_PROPAGATE(dk_get_weights(h, m, &W));
    const size_t ncell = (size_t)h->g.ncell;
    std::vector<int> o2c(256, 0);
    for (int i = 0; i < W->n; i++) o2c[W->orig[i]] = i;
    int* d_o2c = nullptr;
    double* d_dense = nullptr;
    SMRFB_CUDA(cudaMalloc((void**)&d_o2c, 256 * sizeof(int)));
    cudaError_t e = cudaMalloc((void**)&d_dense, ncell * W->n * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_o2c, o2c.data(), 256 * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_dense, 0, ncell * W->n * sizeof(double), h->stream);
    if (e == cudaSuccess) {
        dk_expand_kernel<<<(unsigned)((ncell + 255) / 256), 256, 0, h->stream>>>(W->K, (long long)ncell, W->n, W->widx,
                                                                                  W->wval, d_o2c, d_dense);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(weights_out, d_dense, ncell * W->n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_o2c);
    cudaFree(d_dense);
    if (e != cudaSuccess) {
        set_error("smrfb_dk_weights: %s", cudaGetErrorString(e));
        return SMRFB_ERR_CUDA;
    }
    return SMRFB_OK;
}

int smrfb_dk_stats(smrfb_handle_t hh, int64_t* masks_built, int64_t* fallback_cells, int* max_active) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    DkHandle* h = reinterpret_cast<DkHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_DK, SMRFB_ERR_ARG, "not a DK handle");
    if (masks_built) *masks_built = h->masks_built;
    if (fallback_cells) *fallback_cells = h->fallback_cells;
    if (max_active) *max_active = h->max_active;
    return SMRFB_OK;
}

int smrfb_krige_grid(int nsta, int ngrid, const double* ad, const double* dgrid, const double* elevations, int nthreads,
                     double* weights) {
    (void)nthreads;
    SMRFB_CHECK(nsta > 0 && ngrid > 0 && ad && dgrid && elevations && weights, SMRFB_ERR_ARG, "bad krige_grid arguments");
    SMRFB_CHECK(nsta <= DK3_NMAX, SMRFB_ERR_LIMIT, "krige_grid supports at most %d stations (got %d)", DK3_NMAX, nsta);
    int dev = 0;
    if (const char* e = getenv("SMRFB_DEVICE")) dev = atoi(e);
    SMRFB_CUDA(cudaSetDevice(dev));
    cudaStream_t st;
    SMRFB_CUDA(cudaStreamCreate(&st));
    double *d_dgrid = nullptr, *d_dense = nullptr;
    int* d_o2c = nullptr;
    DkWeights W;
    std::vector<int> orig(nsta);
    for (int i = 0; i < nsta; i++) orig[i] = i;
    int rc = SMRFB_OK;
    cudaError_t e = cudaMalloc((void**)&d_dgrid, (size_t)ngrid * nsta * sizeof(double));
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_dgrid, dgrid, (size_t)ngrid * nsta * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        DkBuildInput in;
        in.n = nsta; in.cx = nullptr; in.cy = nullptr; in.elev = elevations; in.ad = ad; in.d_dgrid = d_dgrid;
        in.orig = orig.data(); in.ncell = ngrid;
        in.g = Geom();
        in.g.ncell = ngrid;
        long long fb = 0;
        rc = dk_build(in, st, &W, &fb);
    }
    if (e == cudaSuccess && rc == SMRFB_OK) {
        e = cudaMalloc((void**)&d_dense, (size_t)ngrid * nsta * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_o2c, 256 * sizeof(int));
        std::vector<int> o2c(256, 0);
        for (int i = 0; i < nsta; i++) o2c[i] = i;
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_o2c, o2c.data(), 256 * sizeof(int), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_dense, 0, (size_t)ngrid * nsta * sizeof(double), st);
        if (e == cudaSuccess) {
            dk_expand_kernel<<<(unsigned)((ngrid + 255) / 256), 256, 0, st>>>(W.K, ngrid, nsta, W.widx, W.wval, d_o2c, d_dense);
            count_launch();
            e = cudaGetLastError();
        }
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(weights, d_dense, (size_t)ngrid * nsta * sizeof(double), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    cudaFree(d_dgrid);
    cudaFree(d_dense);
    cudaFree(d_o2c);
    cudaStreamDestroy(st);
    if (rc == SMRFB_OK && e != cudaSuccess) {
        set_error("smrfb_krige_grid: %s", cudaGetErrorString(e));
        rc = SMRFB_ERR_CUDA;
    }
    return rc;
}

}  // extern "C"
