// idw.cu -- inverse-distance weighting (smrf/spatial/idw.py) as one fused sm_100a kernel.
//
// out[t][c] = ( sum_{s valid at t} w[c][s] * r[t][s] ) / ( sum_{s valid at t} w[c][s] ) + p0[t]*Z[c] + p1[t]
//   w[c][s] = 1 / dist(c, s)^power                                   (idw.py:53-77)
//   r[t][s] = d[t][s] - (mz[s]*p0[t] + p1[t])  (0 for missing)       (idw.py:112-136, from prep_steps_kernel)
//
// One CTA owns CT=64 consecutive cells for the whole batch of T timesteps:
//   * the [S x 64] weight tile is generated ONCE from the cell / station coordinates into shared
//     memory (the reference's [ny,nx,S] cube, 160 GB at 1e8 x 200, is never materialised),
//     together with a double-double sum D[c] of all S weights;
//   * step records (residuals + p0, p1, missing-station list) stream in blocks of TC=32 steps
//     through a 2-deep shared-memory ring filled by 1-D TMA bulk copies (cp.async.bulk + mbarrier);
//   * the numerator is an fp64 tensor-core contraction  C[32 steps x 64 cells] += R[32 x S] * W[S x 64]
//     on DMMA m8n8k4 (8 warps, each 16 steps x 16 cells = 2x2 MMA tiles);
//   * the masked denominator is D[c] minus the weights of that step's missing stations, carried in
//     double-double so no digits are lost when a missing station dominates D;
//   * epilogue: divide, retrend against the DEM, NaN-preserving clamp, 128-bit streaming stores.
#include "common.cuh"

namespace smrfb {

constexpr int IDW_CT = 64;             // cells per CTA
constexpr int IDW_TC = 32;             // timesteps per staged block
constexpr int IDW_WSTR = IDW_CT + 4;   // weight-tile row stride (doubles), = 4 (mod 16): conflict-free B frags
constexpr int IDW_THREADS = 256;
constexpr int IDW_MAXM = 13;           // missing-station list length carried in a record

struct IdwHandle : HandleBase {
    double power = 2.0;
    RecLayout L;
    size_t smem_bytes = 0;
    IdwHandle() : HandleBase(HK_IDW) {}
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
    double vmin, vmax;
    double* out;
};

__device__ __forceinline__ void two_sum(double a, double b, double& s, double& e) {
    s = __dadd_rn(a, b);
    double bb = __dsub_rn(s, a);
    e = __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb));
}

__device__ __forceinline__ double idw_weight(double q, double power) {
    if (power == 2.0) return 1.0 / q;
    return 1.0 / pow(sqrt(q), power);          // idw.py:77 as written
}

__global__ void __launch_bounds__(IDW_THREADS, 1) idw_distribute_kernel(IdwArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.L.S, S_pad = a.L.S_pad, RS = a.L.RS;
    double* Wt = reinterpret_cast<double*>(smem_raw);                  // [S_pad][WSTR]
    double* recbuf = Wt + (size_t)S_pad * IDW_WSTR;                    // [2][TC][RS]
    double* Dhi = recbuf + (size_t)2 * IDW_TC * RS;                    // [CT]
    double* Dlo = Dhi + IDW_CT;                                        // [CT]
    double* Zc = Dlo + IDW_CT;                                         // [CT]
    double* part = Zc + IDW_CT;                                        // [4][CT][2] partial dd sums
    uint64_t* bars = reinterpret_cast<uint64_t*>(part + 4 * IDW_CT * 2);  // [2]
    int* zflag = reinterpret_cast<int*>(bars + 2);                     // [CT]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long cell0 = (long long)blockIdx.x * IDW_CT;
    const int nchunk = a.T_pad / IDW_TC;
    const uint32_t chunk_bytes = (uint32_t)(IDW_TC * RS * sizeof(double));

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    if (tid < IDW_CT) zflag[tid] = 0;
    __syncthreads();
    if (tid == 0) {  // prefetch the first two record blocks while the weight tile is generated
        mbar_arrive_expect_tx(&bars[0], chunk_bytes);
        bulk_g2s(recbuf, a.rec, chunk_bytes, &bars[0]);
        if (nchunk > 1) {
            mbar_arrive_expect_tx(&bars[1], chunk_bytes);
            bulk_g2s(recbuf + (size_t)IDW_TC * RS, a.rec + (size_t)IDW_TC * RS, chunk_bytes, &bars[1]);
        }
    }

    // ---- weight tile: thread (cell = tid&63, station group = tid>>6) ------------------------------
    {
        const int cl = tid & (IDW_CT - 1), sg = tid >> 6;
        const long long c = cell0 + cl;
        const bool live = c < a.g.ncell;
        double X = 0.0, Y = 0.0;
        if (live) cell_xy(a.g, c, X, Y);
        double hi = 0.0, lo = 0.0;
        for (int s = sg; s < S_pad; s += IDW_THREADS / IDW_CT) {
            double w = 0.0;
            if (live && s < S) {
                double dx = __dsub_rn(X, __ldg(a.mx + s)), dy = __dsub_rn(Y, __ldg(a.my + s));
                double q = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                if (q == 0.0) {
                    zflag[cl] = 1;  // coincident station: reference weight is inf; handled in the epilogue
                } else {
                    w = idw_weight(q, a.power);
                    double s1, e1;
                    two_sum(hi, w, s1, e1);
                    hi = s1;
                    lo += e1;
                }
            }
            Wt[s * IDW_WSTR + cl] = w;
        }
        part[(sg * IDW_CT + cl) * 2] = hi;
        part[(sg * IDW_CT + cl) * 2 + 1] = lo;
        if (sg == 0) Zc[cl] = (live && a.g.dem) ? __ldg(a.g.dem + c) : 0.0;
    }
    __syncthreads();
    if (tid < IDW_CT) {
        double hi = part[tid * 2], lo = part[tid * 2 + 1];
        for (int g2 = 1; g2 < 4; g2++) {
            double s1, e1;
            two_sum(hi, part[(g2 * IDW_CT + tid) * 2], s1, e1);
            hi = s1;
            lo += e1 + part[(g2 * IDW_CT + tid) * 2 + 1];
        }
        double s1, e1;
        two_sum(hi, lo, s1, e1);
        Dhi[tid] = s1;
        Dlo[tid] = e1;
    }
    __syncthreads();

    // ---- main loop over staged step blocks -------------------------------------------------------
    const int wm = warp >> 2, wn = warp & 3;          // 2 (time) x 4 (cells) warps
    const int fr = lane >> 2, fc = lane & 3;          // fragment row / col
    const bool even_pairs = ((a.g.ncell & 1) == 0);

    for (int ch = 0; ch < nchunk; ch++) {
        const int buf = ch & 1;
        const double* rb = recbuf + (size_t)buf * IDW_TC * RS;
        mbar_wait(&bars[buf], (ch >> 1) & 1);

        double acc[2][2][2];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

        const double* ap = rb + (size_t)(wm * 16 + fr) * RS + fc;
        const double* bp = Wt + (size_t)fc * IDW_WSTR + wn * 16 + fr;
#pragma unroll 2
        for (int k0 = 0; k0 < S_pad; k0 += 4) {
            double a0 = ap[k0], a1 = ap[k0 + 8 * RS];
            double b0 = bp[k0 * IDW_WSTR], b1 = bp[k0 * IDW_WSTR + 8];
            dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
            dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
            dmma884(acc[1][0][0], acc[1][0][1], a1, b0);
            dmma884(acc[1][1][0], acc[1][1][1], a1, b1);
        }

        // ---- epilogue: thread holds steps {wm*16 + mt*8 + fr} x cells {wn*16 + nt*8 + fc*2 + {0,1}}
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            const int tl = wm * 16 + mt * 8 + fr;
            const int t = ch * IDW_TC + tl;
            if (t >= a.T) continue;
            const double* meta = rb + (size_t)tl * RS + S_pad;
            const double p0 = meta[0], p1 = meta[1];
            const int nmiss = (int)meta[2];
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
                const int cl = wn * 16 + nt * 8 + fc * 2;
                double v[2];
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int c1 = cl + e;
                    double num = acc[mt][nt][e];
                    double den;
                    if (nmiss <= a.L.maxm) {
                        double s = Dhi[c1], er = Dlo[c1];
                        for (int j = 0; j < nmiss; j++) {
                            int m = (int)meta[3 + j];
                            double w = Wt[m * IDW_WSTR + c1];
                            double ns = __dsub_rn(s, w);
                            er = __dadd_rn(er, __dsub_rn(__dsub_rn(s, ns), w));   // fast two-sum error (s >= w)
                            s = ns;
                        }
                        den = __dadd_rn(s, er);
                    } else {  // many stations missing: sum the valid weights directly
                        const uint8_t* vt = a.valid + (size_t)t * S;
                        den = 0.0;
                        for (int s = 0; s < S; s++)
                            if (vt[s]) den += Wt[s * IDW_WSTR + c1];
                    }
                    double val = num / den;
                    if (zflag[c1]) {  // rare: a station sits exactly on this cell
                        const long long c = cell0 + c1;
                        if (c < a.g.ncell) {
                            double X, Y;
                            cell_xy(a.g, c, X, Y);
                            const uint8_t* vt = a.valid + (size_t)t * S;
                            bool hit = false, bad = false;
                            for (int s = 0; s < S; s++) {
                                double dx = __dsub_rn(X, a.mx[s]), dy = __dsub_rn(Y, a.my[s]);
                                if (__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) == 0.0 && vt[s]) {
                                    hit = true;
                                    if (rb[(size_t)tl * RS + s] != 0.0) bad = true;   // inf * r: +-inf / inf
                                }
                            }
                            if (hit) val = bad ? nan("") : 0.0;                      // inf * 0 dropped by nansum: finite / inf
                        }
                    }
                    if (a.detrend) val = __dadd_rn(__dadd_rn(val, __dmul_rn(p0, Zc[c1])), p1);   // idw.py:144, left to right
                    v[e] = clamp_nan_keep(val, a.vmin, a.vmax);
                }
                const long long c = cell0 + cl;
                double* op = a.out + (size_t)t * (size_t)a.g.ncell + c;
                if (c + 1 < a.g.ncell && even_pairs) {
                    __stcs(reinterpret_cast<double2*>(op), make_double2(v[0], v[1]));
                } else {
                    if (c < a.g.ncell) __stcs(op, v[0]);
                    if (c + 1 < a.g.ncell) __stcs(op + 1, v[1]);
                }
            }
        }

        __syncthreads();  // every warp is done with rb before the TMA engine overwrites it
        if (tid == 0 && ch + 2 < nchunk) {
            mbar_arrive_expect_tx(&bars[buf], chunk_bytes);
            bulk_g2s(recbuf + (size_t)buf * IDW_TC * RS, a.rec + (size_t)(ch + 2) * IDW_TC * RS, chunk_bytes, &bars[buf]);
        }
    }
}

static size_t idw_smem_bytes(const RecLayout& L) {
    size_t d = (size_t)L.S_pad * IDW_WSTR + (size_t)2 * IDW_TC * L.RS + 3 * IDW_CT + 4 * IDW_CT * 2;
    return d * sizeof(double) + 2 * sizeof(uint64_t) + IDW_CT * sizeof(int);
}

static int idw_run(IdwHandle* h, const double* d_data, int T, int detrend, int slope_flag, double vmin, double vmax,
                   const double* d_pv_in, double* d_pv_out, double* d_out, cudaStream_t st) {
    SMRFB_CHECK(T > 0, SMRFB_ERR_ARG, "T must be > 0");
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
    a.out = d_out;
    long long nblk = (h->g.ncell + IDW_CT - 1) / IDW_CT;
    SMRFB_CHECK(nblk < 2147483647LL, SMRFB_ERR_LIMIT, "too many cells for one launch");
    idw_distribute_kernel<<<(unsigned)nblk, IDW_THREADS, h->smem_bytes, st>>>(a);
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
    int rc = upload_geom(h, nsta, mx, my, mz, ny, nx, grid_kind, gx, gy, dem);
    if (rc == SMRFB_OK) {
        h->L = make_layout(nsta, 4, IDW_MAXM, true);
        h->smem_bytes = idw_smem_bytes(h->L);
        int smem_max = 0;
        cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
        if (h->smem_bytes > (size_t)smem_max) {   // S_pad*68*8 B weight tile + 2*32*RS*8 B record ring <= 227 KB: S <= 200
            set_error("IDW: %d stations need %zu B of shared memory per CTA, device allows %d", nsta, h->smem_bytes, smem_max);
            rc = SMRFB_ERR_LIMIT;
        }
    }
    if (rc == SMRFB_OK) {
        cudaError_t e = cudaFuncSetAttribute(idw_distribute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)h->smem_bytes);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(%zu B smem): %s", h->smem_bytes, cudaGetErrorString(e));
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
