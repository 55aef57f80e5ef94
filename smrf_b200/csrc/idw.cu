// idw.cu -- inverse-distance weighting (smrf/spatial/idw.py) as one fused sm_100a kernel.
//
// out[t][c] = ( sum_{s valid at t} w[c][s] * r[t][s] ) / ( sum_{s valid at t} w[c][s] ) + p0[t]*Z[c] + p1[t]
//   w[c][s] = 1 / dist(c, s)^power                                   (idw.py:53-77)
//   r[t][s] = d[t][s] - (mz[s]*p0[t] + p1[t])  (0 for missing)       (idw.py:112-136, from prep_steps_kernel)
//
// One CTA (512 threads) owns CT=64 consecutive cells for the whole batch of T timesteps:
//   * the [S x 64] weight tile is generated ONCE from the cell / station coordinates into shared
//     memory (the reference's [ny,nx,S] cube, 160 GB at 1e8 x 200, is never materialised),
//     together with the sum D[c] of all S weights;
//   * step records (residuals + p0, p1, missing-station list) stream in blocks of TC=16 steps
//     through a 3-deep shared-memory ring filled by 1-D TMA bulk copies (cp.async.bulk + mbarrier);
//   * the numerator is an fp64 tensor-core contraction  C[16 steps x 64 cells] += R[16 x S] * W[S x 64]
//     on DMMA m8n8k4.  A warp owns a 16 x 16 output tile (2x2 MMA tiles) over HALF of the stations;
//     the partial tiles of the two K-halves go through a shared-memory transpose buffer, after which
//     every warp finishes two whole steps x 64 cells: all lanes of a warp then share the step (uniform
//     missing-station loop, broadcast metadata reads, conflict-free weight reads, 512 B contiguous stores);
//   * the CTA's 16 warps form two groups of 8 that alternate step blocks: while one group issues
//     DMMAs (tensor pipe, two warps per scheduler -- one warp alone reaches only ~55 % of the DMMA
//     rate) the other runs its epilogue (fp64 pipe + LSU); the tensor pipe is handed over through
//     named barriers, so it never waits for an epilogue;
//   * the masked denominator is D[c] minus the weights of that step's missing stations; if that
//     subtraction cancels more than ~2 digits (a missing station dominates D) the
//     valid weights are re-summed directly, so the result keeps >= 13 digits;
//   * epilogue: reciprocal (rcp.approx + 2 Newton steps), retrend against the DEM, NaN-preserving
//     clamp, 128-bit streaming stores.
#include "common.cuh"
#include "idw_far.cuh"
#include "idw_i8.cuh"

namespace smrfb {

constexpr int IDW_CT = 64;             // cells per CTA
constexpr int IDW_TC = 16;             // timesteps per staged block
constexpr int IDW_NBUF = 3;            // record ring depth
constexpr int IDW_WSTR = IDW_CT + 4;   // weight-tile row stride (doubles), = 4 (mod 16): conflict-free B frags
constexpr int IDW_THREADS = 512;
constexpr int IDW_GROUP = IDW_THREADS / 2;
constexpr int IDW_MAXM = 44;           // missing-station list length carried in a record (11 quads after the count quad)
constexpr int IDW_WU = 5;              // independent weight chains per thread and iteration
constexpr int IDW_XSTR = 72;           // transpose buffer row stride (doubles): = 16 (mod 32) words, conflict-free tile writes

struct IdwHandle : HandleBase {
    double power = 2.0;
    RecLayout L;
    size_t smem_bytes = 0;
    bool far_only = false;      // more stations than the tensor-core kernels hold: only idw_far.cu applies
    bool use_small = false;     // register-weight kernel (few stations); decided at create time
    int i8_mode = 0;            // int8-sliced tcgen05 kernel: 0 = never (SMRFB_IDW_KERNEL=dmma or > 224 stations),
                                // 1 = batches of >= 32 steps (default), 2 = always (SMRFB_IDW_KERNEL=i8)
    IdwI8Scratch i8;
    IdwFarPlan far;             // uniform meshes: near field + interpolated far field (idw_far.cu), the default there
    IdwHandle() : HandleBase(HK_IDW) {}
    ~IdwHandle() override {
        cudaSetDevice(device);
        if (i8.d_blob) cudaFree(i8.d_blob);
        far.release();
    }
};

struct IdwArgs {
    Geom g;
    const double* mx;
    const double* my;
    RecLayout L;
    double power;
    const double* rec;
    const uint8_t* valid;
    int T, T_pad;
    int detrend;
    int clamp;            // 0 when both bounds are infinite: the compare / select instructions are skipped
    double vmin, vmax;
    double* out;          // [T][ncw]
    int out_f32;          // 1: out is float32
    long long cbase;      // first cell of the window this launch distributes
    long long ncw;        // cells in the window
};

template <bool DETREND, bool CLAMP>
__global__ void __launch_bounds__(IDW_THREADS, 1) idw_distribute_kernel(IdwArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.L.S, S_pad = a.L.S_pad, RS = a.L.RS;
    double* Wt = reinterpret_cast<double*>(smem_raw);                  // [S_pad + 1][WSTR]
    double* recbuf = Wt + (size_t)(S_pad + 1) * IDW_WSTR;              // [NBUF][TC][RS]   (row S_pad of Wt is all zero)
    double* Dall = recbuf + (size_t)IDW_NBUF * IDW_TC * RS;            // [CT] sum of all S weights per cell
    double* Zc = Dall + IDW_CT;                                        // [CT]
    double* xbuf = Zc + IDW_CT;                                        // [2 K-halves][TC][XSTR] partial numerator tiles
    double* scratch = xbuf + 2 * IDW_TC * IDW_XSTR;                    // [2*S_pad] station coordinates
    double* scratch2 = scratch + 2 * S_pad;                            // [8*CT] partial weight sums
    uint64_t* bars = reinterpret_cast<uint64_t*>(scratch2 + 8 * IDW_CT);  // [NBUF]
    int* zflag = reinterpret_cast<int*>(bars + IDW_NBUF + 1);          // [CT]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long cell0 = (long long)blockIdx.x * IDW_CT;
    const int nchunk = a.T_pad / IDW_TC;
    const uint32_t chunk_bytes = (uint32_t)(IDW_TC * RS * sizeof(double));
    const size_t chunk_elems = (size_t)IDW_TC * RS;

    if (tid == 0) {
        for (int b = 0; b < IDW_NBUF; b++) mbar_init(&bars[b], 1);
        fence_mbar_init();
    }
    if (tid < IDW_CT) zflag[tid] = 0;
    if (tid < IDW_WSTR) Wt[(size_t)S_pad * IDW_WSTR + tid] = 0.0;     // the zero row the padded missing-list slots point at
    for (int s = tid; s < S_pad; s += IDW_THREADS) {   // station coordinates -> shared memory
        scratch[s] = s < S ? __ldg(a.mx + s) : 0.0;
        scratch[S_pad + s] = s < S ? __ldg(a.my + s) : 0.0;
    }
    __syncthreads();
    if (tid == 0) {  // prefetch the first record blocks while the weight tile is generated
        for (int b = 0; b < IDW_NBUF && b < nchunk; b++) {
            mbar_arrive_expect_tx(&bars[b], chunk_bytes);
            bulk_g2s(recbuf + b * chunk_elems, a.rec + b * chunk_elems, chunk_bytes, &bars[b]);
        }
    }

    // ---- weight tile: thread (cell = tid&63, station group = tid>>6); branch-free so that the
    //      IDW_WU independent chains of one iteration interleave (the loop is latency-bound otherwise)
    {
        const int cl = tid & (IDW_CT - 1), sg = tid >> 6;   // 8 station groups
        const long long c = a.cbase + cell0 + cl;          // global cell index
        const bool live = cell0 + cl < a.ncw;
        double X = 0.0, Y = 0.0;
        if (live) cell_xy(a.g, c, X, Y);
        const double* smx = scratch;
        const double* smy = scratch + S_pad;
        double dsum[IDW_WU];
#pragma unroll
        for (int u = 0; u < IDW_WU; u++) dsum[u] = 0.0;
        bool coincident = false;
        const bool p2 = (a.power == 2.0);
        for (int s0 = sg; s0 < S_pad; s0 += 8 * IDW_WU) {
#pragma unroll
            for (int u = 0; u < IDW_WU; u++) {
                const int s = min(s0 + 8 * u, S_pad - 1);
                const bool in = (s0 + 8 * u < S_pad);
                const double dx = __dsub_rn(X, smx[s]), dy = __dsub_rn(Y, smy[s]);
                const double q = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                const bool ok = live && in && (s < S);
                const bool zero = (q == 0.0);
                coincident |= (ok && zero);   // reference weight is inf; handled in the epilogue
                double w = p2 ? fast_rcp<2>(zero ? 1.0 : q) : 1.0 / pow(sqrt(zero ? 1.0 : q), a.power);   // idw.py:77
                w = (ok && !zero) ? w : 0.0;
                dsum[u] += w;
                if (in) Wt[s * IDW_WSTR + cl] = w;
            }
        }
        if (coincident) zflag[cl] = 1;
        double d = 0.0;
#pragma unroll
        for (int u = 0; u < IDW_WU; u++) d += dsum[u];
        scratch2[sg * IDW_CT + cl] = d;
        if (sg == 0) Zc[cl] = (live && a.g.dem) ? __ldg(a.g.dem + c) : 0.0;
    }
    __syncthreads();
    if (tid < IDW_CT) {
        double d = 0.0;
#pragma unroll
        for (int g2 = 0; g2 < 8; g2++) d += scratch2[g2 * IDW_CT + tid];
        Dall[tid] = d;
    }
    __syncthreads();

    // ---- main loop: two warp groups alternate step blocks ----------------------------------------
    const int grp = warp >> 3;                        // group 0/1
    const int kh = (warp >> 2) & 1;                   // K half: stations [kh*S_pad/2, (kh+1)*S_pad/2)
    const int wn = warp & 3;                          // cell quarter
    const int fr = lane >> 2, fc = lane & 3;          // fragment row / col
    const bool even_pairs = ((a.ncw & 1) == 0);
    const int khalf = S_pad >> 1;
    // after the transpose this thread finishes steps (wg, wg + 8) of its group's block for cells 2*lane, 2*lane + 1
    const int wg = warp & 7;
    double dall[2], zc[2];
    bool anyz = false;
#pragma unroll
    for (int e = 0; e < 2; e++) {
        dall[e] = Dall[2 * lane + e];
        zc[e] = Zc[2 * lane + e];
        anyz |= (zflag[2 * lane + e] != 0);
    }
    const double* wcol = Wt + 2 * lane;

    for (int ch = grp; ch < nchunk; ch += 2) {
        const int buf = ch % IDW_NBUF;
        const double* rb = recbuf + buf * chunk_elems;
        mbar_wait(&bars[buf], (ch / IDW_NBUF) & 1);
#ifndef IDW_FREERUN
        if (ch > 0) named_bar_sync(1 + grp, IDW_THREADS);    // my group's turn on the tensor pipe
#endif

        double acc[2][2][2];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
        {
            const double* ap = rb + (size_t)fr * RS + fc + kh * khalf;
            const double* bp = Wt + (size_t)(fc + kh * khalf) * IDW_WSTR + wn * 16 + fr;
            double a0 = ap[0], a1 = ap[8 * RS], b0 = bp[0], b1 = bp[8];
            for (int k0 = 4; k0 < khalf; k0 += 4) {   // software pipelined: next fragments load under the DMMAs
                const double na0 = ap[k0], na1 = ap[k0 + 8 * RS];
                const double nb0 = bp[k0 * IDW_WSTR], nb1 = bp[k0 * IDW_WSTR + 8];
                dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
                dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
                dmma884(acc[1][0][0], acc[1][0][1], a1, b0);
                dmma884(acc[1][1][0], acc[1][1][1], a1, b1);
                a0 = na0; a1 = na1; b0 = nb0; b1 = nb1;
            }
            dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
            dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
            dmma884(acc[1][0][0], acc[1][0][1], a1, b0);
            dmma884(acc[1][1][0], acc[1][1][1], a1, b1);
        }
#ifndef IDW_FREERUN
        if (ch + 1 < nchunk) named_bar_arrive(2 - grp, IDW_THREADS);   // hand the tensor pipe to the other group
#endif

        // ---- transpose the partial tiles through shared memory (one buffer, used by the two groups in turn)
        if (ch > 0) named_bar_sync(5 + grp, IDW_THREADS);              // the other group has read its tile out
        {
            double* xo = xbuf + (size_t)kh * IDW_TC * IDW_XSTR + (size_t)fr * IDW_XSTR + wn * 16 + fc * 2;
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 2; nt++)
                    *reinterpret_cast<double2*>(xo + mt * 8 * IDW_XSTR + nt * 8) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
        }
        named_bar_sync(3 + grp, IDW_GROUP);
        double num[2][2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const double* xi = xbuf + (size_t)(wg + 8 * h) * IDW_XSTR + 2 * lane;
            const double2 p0v = *reinterpret_cast<const double2*>(xi);
            const double2 p1v = *reinterpret_cast<const double2*>(xi + IDW_TC * IDW_XSTR);
            num[h][0] = p0v.x + p1v.x;
            num[h][1] = p0v.y + p1v.y;
        }
        if (ch + 1 < nchunk) named_bar_arrive(6 - grp, IDW_THREADS);   // buffer free for the other group

        // ---- epilogue: steps tl = wg + 8h (h = 0, 1), cells 2*lane + {0,1}; the whole warp works on one step.
        // Written in phases so that the four outputs of a thread advance together: on an SM sub-partition that
        // another group keeps saturated with DMMAs every dependent instruction waits ~16 cycles for an issue slot,
        // so the length of the dependent chain, not the instruction count, decides when this group is ready again.
        // Masked denominator = (sum of all weights) - (weights of this step's missing stations); when that
        // subtraction cancels more than ~2 digits (a missing station almost on the cell) or the
        // list overflowed, the valid weights are summed directly instead.
        double den[2][2];
        double2 pp[2];
        int nmiss2[2];
        {
            const int4* mi0 = reinterpret_cast<const int4*>(rb + (size_t)wg * RS + S_pad + 2);
            const int4* mi1 = reinterpret_cast<const int4*>(rb + (size_t)(wg + 8) * RS + S_pad + 2);
            pp[0] = *reinterpret_cast<const double2*>(rb + (size_t)wg * RS + S_pad);
            pp[1] = *reinterpret_cast<const double2*>(rb + (size_t)(wg + 8) * RS + S_pad);
            nmiss2[0] = mi0[0].x;
            nmiss2[1] = mi1[0].x;
            den[0][0] = den[1][0] = dall[0];
            den[0][1] = den[1][1] = dall[1];
            // the lists are stored as element offsets (station * WSTR) in whole quads; unused slots point at the zero row
            const int g0 = (min(nmiss2[0], a.L.maxm) + 3) >> 2, g1 = (min(nmiss2[1], a.L.maxm) + 3) >> 2;
            const int gmax = max(g0, g1);
            for (int g = 1; g <= gmax; g++) {
                if (g <= g0) {
                    const int4 q = mi0[g];
                    const double2 wa = *reinterpret_cast<const double2*>(wcol + q.x);
                    const double2 wb = *reinterpret_cast<const double2*>(wcol + q.y);
                    const double2 wc2 = *reinterpret_cast<const double2*>(wcol + q.z);
                    const double2 wd = *reinterpret_cast<const double2*>(wcol + q.w);
                    den[0][0] -= (wa.x + wb.x) + (wc2.x + wd.x);
                    den[0][1] -= (wa.y + wb.y) + (wc2.y + wd.y);
                }
                if (g <= g1) {
                    const int4 q = mi1[g];
                    const double2 wa = *reinterpret_cast<const double2*>(wcol + q.x);
                    const double2 wb = *reinterpret_cast<const double2*>(wcol + q.y);
                    const double2 wc2 = *reinterpret_cast<const double2*>(wcol + q.z);
                    const double2 wd = *reinterpret_cast<const double2*>(wcol + q.w);
                    den[1][0] -= (wa.x + wb.x) + (wc2.x + wd.x);
                    den[1][1] -= (wa.y + wb.y) + (wc2.y + wd.y);
                }
            }
        }
        // per-cell decision, so a cell's value never depends on which other cells share its thread / tile:
        // den < ~D * 2^-7 (or den <= 0), compared on the high words with integer instructions (every fp64
        // instruction issued here costs tensor-pipe time, profiles/r01_fp64_mix_probe.txt)
        bool need[2][2], redo = false;
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                need[h][e] = (nmiss2[h] > a.L.maxm) || (__double2hiint(den[h][e]) < __double2hiint(dall[e]) - (7 << 20));
                redo |= need[h][e];
            }
        if (redo) {   // rare: exact masked sum over the valid stations
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int t = ch * IDW_TC + wg + 8 * h;
                if ((need[h][0] || need[h][1]) && t < a.T) {
                    const uint8_t* vt = a.valid + (size_t)t * S;
                    double ds0 = 0.0, ds1 = 0.0;
                    for (int s = 0; s < S; s++)
                        if (vt[s]) {
                            const double2 w0 = *reinterpret_cast<const double2*>(wcol + (size_t)s * IDW_WSTR);
                            ds0 += w0.x;
                            ds1 += w0.y;
                        }
                    if (need[h][0]) den[h][0] = ds0;
                    if (need[h][1]) den[h][1] = ds1;
                }
            }
        }
        // num / den: reciprocal seed (MUFU.RCP64H, ~2^-10) + two Newton steps, then one residual correction of
        // the quotient (~1 ulp; den is a positive normal number for every live cell); four independent chains
        double val[2][2];
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const double d = den[h][e];
                const double r = fast_rcp<2>(d);
                const double q = num[h][e] * r;
                val[h][e] = fma(fma(-d, q, num[h][e]), r, q);
            }
        if (anyz) {   // rare: a station sits exactly on one of this thread's cells
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int tl = wg + 8 * h;
                    const int t = ch * IDW_TC + tl;
                    const long long c = cell0 + 2 * lane + e;
                    if (zflag[2 * lane + e] && c < a.ncw && t < a.T) {
                        double X, Y;
                        cell_xy(a.g, a.cbase + c, X, Y);
                        const uint8_t* vt = a.valid + (size_t)t * S;
                        bool hit = false, bad = false;
                        for (int s = 0; s < S; s++) {
                            double dx = __dsub_rn(X, a.mx[s]), dy = __dsub_rn(Y, a.my[s]);
                            if (__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) == 0.0 && vt[s]) {
                                hit = true;
                                if (rb[(size_t)tl * RS + s] != 0.0) bad = true;   // inf * r: +-inf / inf
                            }
                        }
                        if (hit) val[h][e] = bad ? nan("") : 0.0;                // inf * 0 dropped by nansum: finite / inf
                    }
                }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int t = ch * IDW_TC + wg + 8 * h;
            double v[2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
                double x = val[h][e];
                if (DETREND) x = __dadd_rn(__dadd_rn(x, __dmul_rn(pp[h].x, zc[e])), pp[h].y);   // idw.py:144, left to right
                v[e] = CLAMP ? clamp_nan_keep(x, a.vmin, a.vmax) : x;
            }
            if (t < a.T) {
                const long long c = cell0 + 2 * lane;
                const size_t oi = (size_t)t * (size_t)a.ncw + c;
                if (c + 1 < a.ncw && even_pairs) {
                    store_out2(a.out, oi, v[0], v[1], a.out_f32);
                } else {
                    if (c < a.ncw) store_out(a.out, oi, v[0], a.out_f32);
                    if (c + 1 < a.ncw) store_out(a.out, oi + 1, v[1], a.out_f32);
                }
            }
        }

        named_bar_sync(3 + grp, IDW_GROUP);  // the whole group is done with rb before the TMA engine overwrites it
        if ((tid & (IDW_GROUP - 1)) == 0 && ch + IDW_NBUF < nchunk) {
            mbar_arrive_expect_tx(&bars[buf], chunk_bytes);
            bulk_g2s(recbuf + buf * chunk_elems, a.rec + (size_t)(ch + IDW_NBUF) * chunk_elems, chunk_bytes, &bars[buf]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Few stations (S_pad <= 16; real basins often have a handful, RME has 2): 2*S flop per 8-byte output
// is left of the fp64 ridge, so the tensor-core tile machinery above only costs instructions.
// idw_small_kernel: one thread per cell keeps its S weights in REGISTERS (plus a copy in shared
// memory for the per-step lookup of missing stations' weights), step records stream through a
// double-buffered shared-memory ring of 1-D TMA bulk copies, and every output is S FMAs with
// broadcast shared-memory reads of the residuals, the masked denominator, a reciprocal, retrend,
// clamp and a coalesced streaming store.  Bound by the output stream once S is small.
// ------------------------------------------------------------------------------------------
constexpr int IDWS_THREADS = 256;
constexpr int IDWS_TCH = 16;

template <int SMAX>
__global__ void __launch_bounds__(IDWS_THREADS) idw_small_kernel(IdwArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.L.S, S_pad = a.L.S_pad, RS = a.L.RS;
    double* recbuf = reinterpret_cast<double*>(smem_raw);                 // [2][TCH][RS]
    double* wsm = recbuf + 2 * IDWS_TCH * RS;                             // [S_pad][THREADS] weights, for missing-station lookups
    uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + (size_t)S_pad * IDWS_THREADS);   // [2]
    const int tid = threadIdx.x;
    const long long c = (long long)blockIdx.x * IDWS_THREADS + tid;      // cell within the window
    const bool live = c < a.ncw;
    const long long gcell = a.cbase + (live ? c : 0);
    const int nchunk = a.T_pad / IDWS_TCH;
    const uint32_t chunk_bytes = (uint32_t)(IDWS_TCH * RS * sizeof(double));
    const size_t chunk_elems = (size_t)IDWS_TCH * RS;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int b = 0; b < 2 && b < nchunk; b++) {
            mbar_arrive_expect_tx(&bars[b], chunk_bytes);
            bulk_g2s(recbuf + b * chunk_elems, a.rec + b * chunk_elems, chunk_bytes, &bars[b]);
        }
    }
    // weights of this cell (registers + shared copy), their sum, coincident-station flag
    double w[SMAX];
    double dall = 0.0;
    bool coincident = false;
    {
        double X = 0.0, Y = 0.0;
        if (live) cell_xy(a.g, gcell, X, Y);
        const bool p2 = (a.power == 2.0);
#pragma unroll
        for (int s = 0; s < SMAX; s++) {
            double ws = 0.0;
            if (s < S) {
                const double dx = __dsub_rn(X, __ldg(a.mx + s)), dy = __dsub_rn(Y, __ldg(a.my + s));
                const double q = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                const bool zero = (q == 0.0);
                coincident |= (live && zero);
                ws = p2 ? fast_rcp<2>(zero ? 1.0 : q) : 1.0 / pow(sqrt(zero ? 1.0 : q), a.power);   // idw.py:77
                ws = (live && !zero) ? ws : 0.0;
            }
            w[s] = ws;
            dall += ws;
            if (s < S_pad) wsm[(size_t)s * IDWS_THREADS + tid] = ws;
        }
    }
    const double Z = (live && a.g.dem) ? __ldg(a.g.dem + gcell) : 0.0;
    const double* wmine = wsm + tid;

    for (int ch = 0; ch < nchunk; ch++) {
        const int buf = ch & 1;
        const double* rb = recbuf + buf * chunk_elems;
        mbar_wait(&bars[buf], (ch >> 1) & 1);
        if (live) {
#pragma unroll 2
            for (int tl = 0; tl < IDWS_TCH; tl++) {
                const int t = ch * IDWS_TCH + tl;
                if (t >= a.T) break;
                const double* rr = rb + (size_t)tl * RS;          // warp-uniform address: broadcast reads
                double num = 0.0;
#pragma unroll
                for (int s = 0; s < SMAX; s += 2) {
                    const double2 r2 = *reinterpret_cast<const double2*>(rr + s);   // S_pad is a multiple of 8; pad residuals are 0
                    num = fma(w[s], r2.x, num);
                    num = fma(w[s + 1], r2.y, num);
                }
                const double2 pp = *reinterpret_cast<const double2*>(rr + S_pad);
                const int* mi = reinterpret_cast<const int*>(rr + S_pad + 2);
                const int nmiss = mi[0];
                double den = dall;
                const int nlist = min(nmiss, a.L.maxm);
                for (int j = 0; j < nlist; j++) den -= wmine[mi[4 + j]];          // offsets = station * IDWS_THREADS
                if (nmiss > a.L.maxm || __double2hiint(den) < __double2hiint(dall) - (7 << 20)) {   // rare: exact masked sum
                    const uint8_t* vt = a.valid + (size_t)t * S;
                    double ds = 0.0;
                    for (int s = 0; s < S; s++)
                        if (vt[s]) ds += wmine[(size_t)s * IDWS_THREADS];
                    den = ds;
                }
                const double r = fast_rcp<2>(den);
                const double q = num * r;
                double val = fma(fma(-den, q, num), r, q);
                if (coincident) {   // rare: a station sits exactly on this cell
                    double X, Y;
                    cell_xy(a.g, gcell, X, Y);
                    const uint8_t* vt = a.valid + (size_t)t * S;
                    bool hit = false, bad = false;
                    for (int s = 0; s < S; s++) {
                        double dx = __dsub_rn(X, a.mx[s]), dy = __dsub_rn(Y, a.my[s]);
                        if (__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) == 0.0 && vt[s]) {
                            hit = true;
                            if (rr[s] != 0.0) bad = true;
                        }
                    }
                    if (hit) val = bad ? nan("") : 0.0;
                }
                if (a.detrend) val = __dadd_rn(__dadd_rn(val, __dmul_rn(pp.x, Z)), pp.y);   // idw.py:144, left to right
                if (a.clamp) val = clamp_nan_keep(val, a.vmin, a.vmax);
                store_out(a.out, (size_t)t * (size_t)a.ncw + c, val, a.out_f32);
            }
        }
        __syncthreads();   // everyone is done with rb before the TMA engine overwrites it
        if (tid == 0 && ch + 2 < nchunk) {
            mbar_arrive_expect_tx(&bars[buf], chunk_bytes);
            bulk_g2s(recbuf + buf * chunk_elems, a.rec + (size_t)(ch + 2) * chunk_elems, chunk_bytes, &bars[buf]);
        }
    }
}

static size_t idw_small_smem_bytes(const RecLayout& L) {
    return ((size_t)2 * IDWS_TCH * L.RS + (size_t)L.S_pad * IDWS_THREADS) * sizeof(double) + 2 * sizeof(uint64_t);
}

static size_t idw_smem_bytes(const RecLayout& L) {
    size_t d = (size_t)(L.S_pad + 1) * IDW_WSTR + (size_t)IDW_NBUF * IDW_TC * L.RS + 2 * IDW_CT + 2 * IDW_TC * IDW_XSTR +
               2 * (size_t)L.S_pad + 8 * IDW_CT;
    return d * sizeof(double) + (IDW_NBUF + 1) * sizeof(uint64_t) + IDW_CT * sizeof(int);
}

static int idw_run(IdwHandle* h, const double* d_data, int T, int detrend, int slope_flag, double vmin, double vmax,
                   const double* d_pv_in, double* d_pv_out, double* d_out, cudaStream_t st, long long cbase = 0,
                   long long ncw = -1) {
    if (ncw < 0) ncw = h->g.ncell;
    SMRFB_CHECK(T > 0, SMRFB_ERR_ARG, "T must be > 0");
    SMRFB_CHECK(cbase >= 0 && ncw > 0 && cbase + ncw <= h->g.ncell, SMRFB_ERR_ARG, "cell window [%lld, +%lld) outside the grid",
                cbase, ncw);
    SMRFB_CHECK(!detrend || (h->g.dem && h->d_mz), SMRFB_ERR_ARG, "detrended IDW needs mz and GridZ");
    const int T_pad = (T + IDW_TC - 1) / IDW_TC * IDW_TC;
    SMRFB_PROPAGATE(ensure_cap((void**)&h->d_rec, &h->rec_cap, (size_t)T_pad * h->L.RS * sizeof(double)));
    SMRFB_PROPAGATE(ensure_cap((void**)&h->d_valid, &h->valid_cap, (size_t)T * h->S));
    SMRFB_PROPAGATE(launch_prep(st, d_data, T, T_pad, h->S, h->L, h->d_mz, nullptr, detrend, slope_flag, d_pv_in,
                                d_pv_out, h->d_rec, h->d_valid, h->d_flags));
    IdwArgs a;
    a.g = h->g;
    a.mx = h->d_mx;
    a.my = h->d_my;
    a.L = h->L;
    a.power = h->power;
    a.rec = h->d_rec;
    a.valid = h->d_valid;
    a.T = T;
    a.T_pad = T_pad;
    a.detrend = detrend;
    a.vmin = vmin;
    a.vmax = vmax;
    a.clamp = !(std::isinf(vmin) && vmin < 0 && std::isinf(vmax) && vmax > 0);
    a.out = d_out;
    a.out_f32 = h->out_f32;
    a.cbase = cbase;
    a.ncw = ncw;
    // uniform mesh, whole rows: near field + interpolated far field (idw_far.cu), any number of stations
    if (h->far.enabled && cbase % h->g.nx == 0 && ncw % h->g.nx == 0)
        return idw_far_run(st, &h->far, h->g, h->d_mx, h->d_my, h->S, h->power, h->d_rec, h->L.RS, h->L.S_pad, h->d_valid, T,
                           detrend, a.clamp, vmin, vmax, d_out, h->out_f32, (int)(cbase / h->g.nx), (int)(ncw / h->g.nx), h->device);
    SMRFB_CHECK(!h->far_only, SMRFB_ERR_LIMIT, "IDW: %d stations are only supported on whole-row windows of a uniform mesh", h->S);
    // few stations: register-weight kernel.  Measured on B200 (4000 x 4000 cells, T = 512): S = 2: 348 G cell-steps/s;
    // S = 20: 178 G with either kernel; S = 32: 148 G here vs 163 G on the tensor-core kernel => switch at 16.
    if (h->use_small) {
        const long long nb = (ncw + IDWS_THREADS - 1) / IDWS_THREADS;
        SMRFB_CHECK(nb < 2147483647LL, SMRFB_ERR_LIMIT, "too many cells for one launch");
        const size_t sm = idw_small_smem_bytes(h->L);
#define IDWS_LAUNCH(N)                                                                                              \
    do {                                                                                                            \
        SMRFB_CUDA(cudaFuncSetAttribute(idw_small_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        idw_small_kernel<N><<<(unsigned)nb, IDWS_THREADS, sm, st>>>(a);                                              \
    } while (0)
        if (h->L.S_pad <= 8) IDWS_LAUNCH(8);
        else IDWS_LAUNCH(16);
#undef IDWS_LAUNCH
        count_launch();
        SMRFB_CUDA(cudaGetLastError());
        return SMRFB_OK;
    }
    // long batches: the contraction runs on the int8 tensor cores (idw_i8.cu), ~4x the DMMA rate; a chunk there is 80
    // steps, so short batches stay on the fp64 tensor-core kernel below
    if (h->i8_mode == 2 || (h->i8_mode == 1 && T >= 32))
        return idw_i8_run(st, &h->i8, h->g, h->d_mx, h->d_my, h->S, h->power, h->d_rec, h->L.RS, h->L.S_pad, h->d_valid, T,
                          detrend, vmin, vmax, d_out, h->out_f32, cbase, ncw, h->device);
    long long nblk = (ncw + IDW_CT - 1) / IDW_CT;
    SMRFB_CHECK(nblk < 2147483647LL, SMRFB_ERR_LIMIT, "too many cells for one launch");
    if (detrend) {
        if (a.clamp) idw_distribute_kernel<true, true><<<(unsigned)nblk, IDW_THREADS, h->smem_bytes, st>>>(a);
        else idw_distribute_kernel<true, false><<<(unsigned)nblk, IDW_THREADS, h->smem_bytes, st>>>(a);
    } else {
        if (a.clamp) idw_distribute_kernel<false, true><<<(unsigned)nblk, IDW_THREADS, h->smem_bytes, st>>>(a);
        else idw_distribute_kernel<false, false><<<(unsigned)nblk, IDW_THREADS, h->smem_bytes, st>>>(a);
    }
    count_launch();
    SMRFB_CUDA(cudaGetLastError());
    return SMRFB_OK;
}

}  // namespace smrfb

using namespace smrfb;

extern "C" {

int smrfb_idw_create(int nsta, const double* mx, const double* my, const double* mz, int ny, int nx, int grid_kind,
                     const double* gx, const double* gy, const double* dem, double power, int device,
                     smrfb_handle_t* out) {
    SMRFB_CHECK(out, SMRFB_ERR_ARG, "out handle pointer is NULL");
    *out = nullptr;
    IdwHandle* h = new IdwHandle();
    h->device = device;
    h->power = power;
    int smem_max = 0;
    int rc = upload_geom(h, nsta, mx, my, mz, ny, nx, grid_kind, gx, gy, dem);
    if (rc == SMRFB_OK) rc = idw_far_plan(&h->far, nsta, mx, my, ny, nx, grid_kind, gx, gy, power, device);
    if (rc == SMRFB_OK) {
        h->L = make_layout(nsta, 8, IDW_MAXM, true);
        h->use_small = (h->L.S_pad <= 16) && !getenv("SMRFB_IDW_FORCE_MMA");
        if (h->use_small) {                // register-weight kernel: offsets into wsm[station][thread]
            h->L.miss_scale = IDWS_THREADS;
            h->L.miss_fill = 0;
        } else {                           // tensor-core kernel: offsets into Wt[station][WSTR], fill = the zero row
            h->L.miss_scale = IDW_WSTR;
            h->L.miss_fill = h->L.S_pad * IDW_WSTR;
        }
        if (!h->use_small && nsta <= IDW_I8_MAX_STATIONS) {
            const char* k = getenv("SMRFB_IDW_KERNEL");
            h->i8_mode = (k && !strcmp(k, "dmma")) ? 0 : (k && !strcmp(k, "i8")) ? 2 : 1;
            if (h->i8_mode) rc = idw_i8_prepare(device);
        }
        h->smem_bytes = idw_smem_bytes(h->L);
        cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
        if (h->smem_bytes > (size_t)smem_max) {   // S_pad*68*8 B weight tile + 2*32*RS*8 B record ring <= 227 KB: S <= 200
            if (rc == SMRFB_OK && h->i8_mode) {   // 201..224 stations: only the int8 tensor-core kernel holds them
                h->i8_mode = 2;
                h->smem_bytes = 0;
            } else if (rc == SMRFB_OK && h->far.enabled) {   // beyond: near / far field only
                h->far_only = true;
                h->smem_bytes = 0;
            } else {
                set_error("IDW: %d stations need %zu B of shared memory per CTA, device allows %d", nsta, h->smem_bytes, smem_max);
                rc = SMRFB_ERR_LIMIT;
            }
        }
    }
    if (rc == SMRFB_OK && h->smem_bytes) {
        // the attribute is per function and device, not per handle: opt in to the device maximum so that a later
        // handle with fewer stations cannot lower the limit under an earlier one
        cudaError_t e = cudaFuncSetAttribute(idw_distribute_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(idw_distribute_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(idw_distribute_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(idw_distribute_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(%d B smem): %s", smem_max, cudaGetErrorString(e));
            rc = SMRFB_ERR_CUDA;
        }
    }
    if (rc != SMRFB_OK) {
        delete h;
        return rc;
    }
    *out = reinterpret_cast<smrfb_handle_t>(h);
    return SMRFB_OK;
}

int smrfb_idw_distribute_dev(smrfb_handle_t hh, const double* d_data, int T, int detrend, int slope_flag, double vmin,
                             double vmax, const double* d_pv_in, double* d_pv_out, double* d_out, void* stream) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    IdwHandle* h = reinterpret_cast<IdwHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_IDW, SMRFB_ERR_ARG, "not an IDW handle");
    SMRFB_CHECK(d_data && d_out, SMRFB_ERR_ARG, "NULL data/out");
    SMRFB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    return idw_run(h, d_data, T, detrend, slope_flag, vmin, vmax, d_pv_in, d_pv_out, d_out, st);
}

int smrfb_idw_distribute_window_dev(smrfb_handle_t hh, const double* d_data, int T, int detrend, int slope_flag,
                                    double vmin, double vmax, const double* d_pv_in, double* d_pv_out, int64_t cell_begin,
                                    int64_t cell_count, double* d_out, void* stream) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    IdwHandle* h = reinterpret_cast<IdwHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_IDW, SMRFB_ERR_ARG, "not an IDW handle");
    SMRFB_CHECK(d_data && d_out, SMRFB_ERR_ARG, "NULL data/out");
    SMRFB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    return idw_run(h, d_data, T, detrend, slope_flag, vmin, vmax, d_pv_in, d_pv_out, d_out, st, cell_begin, cell_count);
}

int smrfb_idw_info(smrfb_handle_t hh, int info[8]) {
    SMRFB_CHECK(hh && info, SMRFB_ERR_ARG, "NULL handle / info");
    IdwHandle* h = reinterpret_cast<IdwHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_IDW, SMRFB_ERR_ARG, "not an IDW handle");
    info[0] = h->far.enabled ? 1 : 0;
    info[1] = h->far.nn;
    info[2] = h->far.max_near;
    info[3] = h->far.npx * h->far.npy;
    info[4] = (int)(h->far.theta * 1000.0);
    info[5] = h->i8_mode ? 1 : 0;
    info[6] = h->use_small ? 1 : 0;
    info[7] = h->far.enabled && h->far.npx * h->far.npy > 0 ? (int)(1000.0 * h->far.near_off.back() / (h->far.npx * h->far.npy)) : 0;
    return SMRFB_OK;
}

int smrfb_idw_far_timing(smrfb_handle_t hh, int enable, double ms_out[3]) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    IdwHandle* h = reinterpret_cast<IdwHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_IDW, SMRFB_ERR_ARG, "not an IDW handle");
    SMRFB_CUDA(cudaSetDevice(h->device));
    if (ms_out) {
        ms_out[0] = ms_out[1] = ms_out[2] = 0.0;
        std::vector<cudaEvent_t>& ev = h->far.tev;
        for (size_t i = 0; i + 2 < ev.size(); i += 3) {
            float a = 0.f, b = 0.f;
            SMRFB_CUDA(cudaEventSynchronize(ev[i + 2]));
            SMRFB_CUDA(cudaEventElapsedTime(&a, ev[i], ev[i + 1]));
            SMRFB_CUDA(cudaEventElapsedTime(&b, ev[i + 1], ev[i + 2]));
            ms_out[0] += a;
            ms_out[1] += b;
            ms_out[2] += 1.0;
        }
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
        ev.clear();
    }
    h->far.timing = enable != 0;
    return SMRFB_OK;
}

// Host-buffer entry: the batch is cut into sub-batches of steps so that a double-buffered device
// output ring stays within a fixed budget; the D2H copy of one sub-batch overlaps the kernel of the next.
int smrfb_idw_distribute(smrfb_handle_t hh, const double* data, int T, int detrend, int slope_flag, double vmin,
                         double vmax, const double* pv_in, double* pv_out, double* out) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    IdwHandle* h = reinterpret_cast<IdwHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_IDW, SMRFB_ERR_ARG, "not an IDW handle");
    SMRFB_CHECK(data && out && T > 0, SMRFB_ERR_ARG, "NULL data/out or T <= 0");
    SMRFB_CUDA(cudaSetDevice(h->device));
    const size_t ncell = (size_t)h->g.ncell;
    const size_t S = (size_t)h->S;
    SMRFB_PROPAGATE(ensure_cap((void**)&h->d_data, &h->data_cap, (size_t)T * S * sizeof(double)));
    SMRFB_PROPAGATE(ensure_cap((void**)&h->d_pv, &h->pv_cap, (size_t)T * 4 * sizeof(double)));
    SMRFB_CUDA(cudaMemsetAsync(h->d_flags, 0, sizeof(int), h->stream));
    SMRFB_CUDA(cudaMemcpyAsync(h->d_data, data, (size_t)T * S * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    double* d_pvin = nullptr;
    if (pv_in) {
        d_pvin = h->d_pv + (size_t)2 * T;
        SMRFB_CUDA(cudaMemcpyAsync(d_pvin, pv_in, (size_t)T * 2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    int rc = run_host_batches(h, T, IDW_TC, out, [&](int t0, int tn, double* d_out, cudaStream_t st) {
        return idw_run(h, h->d_data + (size_t)t0 * S, tn, detrend, slope_flag, vmin, vmax,
                       d_pvin ? d_pvin + (size_t)2 * t0 : nullptr, h->d_pv + (size_t)2 * t0, d_out, st);
    });
    int flags = 0;
    if (rc == SMRFB_OK) {
        cudaError_t e = cudaSuccess;
        if (pv_out) e = cudaMemcpy(pv_out, h->d_pv, (size_t)T * 2 * sizeof(double), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(&flags, h->d_flags, sizeof(int), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            set_error("IDW: reading back coefficients failed: %s", cudaGetErrorString(e));
            rc = SMRFB_ERR_CUDA;
        }
    }
    if (rc == SMRFB_OK && (flags & 1)) {
        set_error("a timestep has no valid station (all values NaN)");
        rc = SMRFB_ERR_NODATA;
    }
    return rc;
}

}  // extern "C"
