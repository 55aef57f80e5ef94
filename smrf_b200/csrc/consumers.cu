// consumers.cu -- the immediate per-cell consumers of a freshly distributed grid (SURVEY.md 8f N-a) and the WindNinja-style
// precomputed-weights interpolation (8a G5):
//   * dew point from vapor pressure: smrf/envphys/core/dewpt.c:14-229 (Brent root of sati(T) = e, float-truncated result,
//     K -> C) with the saturation curves of topotherm.c:109-176 -- either as a consumer hooked onto a distribute call
//     (the vapor-pressure block is still in the handle's device ring: no second trip through host memory) or stand-alone;
//   * wet / ice bulb temperature from air and dew point temperature: smrf/envphys/core/iwbt.c:14-134 (Newton iteration on the
//     psychrometer equation, at most 10 steps), what precipitation.py uses as the precipitation temperature;
//   * wind azimuth from the two distributed direction components: smrf/distribute/wind/wind.py:175-178;
//   * precipitation phase and new-snow density from the precipitation temperature and the distributed precipitation:
//     smrf/envphys/snow.py:37-75 with the three models of smrf/envphys/nasde_model/ (susong1999, piecewise_susong1999, marks2017);
//   * grid_interpolate with the vertices / weights of interp_weights: smrf/utils/utils.py:377-427.
#include "common.cuh"

namespace smrfb {

constexpr double ENV_FREEZE = 2.7316e2;      // envphys.h:31
constexpr double ENV_BOIL = 3.7315e2;        // envphys.h:32
constexpr double ENV_SEA_LEVEL = 1.013246e5; // envphys.h:74

// topotherm.c:109-137
__device__ __forceinline__ double env_satw(double tk) {
    const double l10 = log(1.e1);
    double x = -7.90298 * (ENV_BOIL / tk - 1.) + 5.02808 * log(ENV_BOIL / tk) / l10 -
               1.3816e-7 * (pow(1.e1, 1.1344e1 * (1. - tk / ENV_BOIL)) - 1.) +
               8.1328e-3 * (pow(1.e1, -3.49149 * (ENV_BOIL / tk - 1.)) - 1.) + log(ENV_SEA_LEVEL) / l10;
    return pow(1.e1, x);
}
// topotherm.c:143-176
__device__ __forceinline__ double env_sati(double tk) {
    if (tk > ENV_FREEZE) return env_satw(tk);
    const double l10 = log(1.e1);
    const double x = pow(1.e1, -9.09718 * ((ENV_FREEZE / tk) - 1.) - 3.56654 * log(ENV_FREEZE / tk) / l10 +
                                   8.76793e-1 * (1. - (tk / ENV_FREEZE)) + log(6.1071) / l10);
    return x * 1.e2;
}

// dewpt.c:95-229 (zerobr: Brent's method on f(t) = e - sati(t)); returns 0 where the reference gives up
__device__ double env_zerobr(double a, double b, double t, double e_pass) {
    const double meps = 2.2204460492503131e-16;
    double c = 0., d = 0., e = 0., fa, fb, fc, m, p, q, r, s, tol;
    if (a == b) return a;
    fb = (fabs(b) >= fabs(a)) ? fabs(b) : fabs(a);
    tol = 5.e-1 * t + 2 * meps * fb;
    s = log(fabs(b - a) / tol) / log(2.);
    int maxfun = (int)(s * s + 1);
    fa = e_pass - env_sati(a);
    fc = fb = e_pass - env_sati(b);
    if (fabs(fb) <= tol) return b;
    if (fabs(fa) <= tol) return a;
    if (fb * fa > 0) return 0.;
    while (maxfun--) {
        if ((fb > 0 && fc > 0) || (fb <= 0 && fc <= 0)) {
            c = a;
            fc = fa;
            d = e = b - a;
        }
        if (fabs(fc) < fabs(fb)) {
            a = b;
            b = c;
            c = a;
            fa = fb;
            fb = fc;
            fc = fa;
        }
        tol = meps * fabs(b) + t;
        m = (c - b) / 2;
        if (fabs(m) < tol || fb == 0) return b;
        if (fabs(e) < tol || fabs(fa) <= fabs(fb)) {
            d = e = m;
        } else {
            s = fb / fa;
            if (a == c) {
                p = 2 * m * s;
                q = 1 - s;
            } else {
                q = fa / fc;
                r = fb / fc;
                p = s * (2 * m * q * (q - r) - (b - a) * (r - 1));
                q -= 1;
                q *= (r - 1) * (s - 1);
            }
            if (p > 0) q = -q;
            else p = -p;
            s = e;
            e = d;
            if (2 * p < 3 * m * q - fabs(tol * q) && p < fabs(s * q / 2)) d = p / q;
            else d = e = m;
        }
        a = b;
        fa = fb;
        if (fabs(d) > tol) b += d;
        else if (m > 0) b += tol;
        else b -= tol;
        fb = e_pass - env_sati(b);
    }
    return 0.;
}

// dewpt.c:63-92 + :14-47: spanning guesses, root, float truncation, K -> C.  NaN in -> NaN out (the reference would loop).
__device__ double env_dew_point_c(double ea, double tol) {
    if (!(ea == ea)) return ea;
    double a = ENV_FREEZE;
    int guard = 0;
    while (ea < env_sati(a) && guard++ < 64) a *= .75;
    double b = ENV_FREEZE + 15;
    guard = 0;
    while (ea > env_sati(b) && guard++ < 64) b *= 1.25;
    float dpt = (float)env_zerobr(a, b, tol, ea);
    dpt -= (float)ENV_FREEZE;                 // `dpt_p -= FREEZE` on a float: the subtraction rounds to float
    return (double)dpt;
}

__global__ void dew_point_kernel(const double* __restrict__ ea, const float* __restrict__ ea32, long long n, double tol,
                                 double* __restrict__ out, float* __restrict__ out32) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double e = ea ? ea[i] : (double)ea32[i];
        const double v = env_dew_point_c(e, tol);
        if (out) __stcs(out + i, v);
        else __stcs(out32 + i, (float)v);
    }
}

int launch_dew_point(cudaStream_t st, const void* d_ea, int in_f32, long long n, double tol, void* d_out, int out_f32) {
    if (n <= 0) return SMRFB_OK;
    const int threads = 256;
    const long long blocks = std::min<long long>((n + threads - 1) / threads, 148LL * 32);
    dew_point_kernel<<<(unsigned)blocks, threads, 0, st>>>(in_f32 ? nullptr : (const double*)d_ea, in_f32 ? (const float*)d_ea : nullptr, n, tol,
                                                           out_f32 ? nullptr : (double*)d_out, out_f32 ? (float*)d_out : nullptr);
    count_launch();
    SMRFB_CUDA(cudaGetLastError());
    return SMRFB_OK;
}

// iwbt.c:14-81 wetbulb(ta, dpt, press, tol), temperatures in K; *bad is set where the reference would exit(-1)
__device__ double env_wetbulb(double ta, double dpt, double press, double tol, int* bad) {
    const double RH2O = 461.5, EPSW = 18.0153 / 28.9644, CP_AIR = 1.005e3;      // iwbt.c:11-12, envphys.h:16,21,37
    auto lh_vap = [](double t) { return 2.5e6 - 2.95573e3 * (t - ENV_FREEZE); };   // envphys.h:281
    auto lh_fus = [](double t) { return 3.336e5 + 1.6667e2 * (ENV_FREEZE - t); };   // envphys.h:288
    double xlh;
    if (ta <= ENV_FREEZE) {
        xlh = lh_vap((ta + dpt) / 2.0) + lh_fus((ta + dpt) / 2.0);
    } else if (dpt <= ENV_FREEZE) {
        const double xlhv = lh_vap((ta + dpt) / 2.0), xlhf = lh_fus((ENV_FREEZE + dpt) / 2.0);
        const double fu_fac = (ENV_FREEZE - dpt) / (ta - dpt);
        xlh = xlhv + (fu_fac * xlhf);
    } else {
        xlh = lh_vap((ta + dpt) / 2);
    }
    const double ea = env_sati(dpt);
    double esat = env_sati(ta);
    const double psyc = EPSW * (xlh / (CP_AIR * press));
    double dti = 1.0, ti = ta;
    int i = 0;
    while (dti > tol) {
        const double ti0 = ti;
        if (ti != ta) esat = env_sati(ti);
        const double dedt = xlh * (esat / (RH2O * (ti * ti)));
        const double pf = (ti - ta) + (psyc * (esat - ea));
        const double dpdt = 1.0 + (psyc * dedt);
        ti = ti - (pf / dpdt);
        dti = ti0 - ti;
        i++;
        if (i > 10) {          // iwbt.c:74-77: "failure to converge in 10 iterations", exit(-1)
            *bad = 1;
            break;
        }
    }
    return ti;
}

// iwbt.c:84-134: per cell, pressure from the elevation (hydrostatic, standard atmosphere), C -> K, wetbulb, K -> C
__global__ void wet_bulb_kernel(const double* __restrict__ ta, const double* __restrict__ td, const double* __restrict__ z, long long n,
                                double tol, double* __restrict__ tw, int* __restrict__ flags) {
    const double RGAS = 8.31432e3, STD_AIRTMP = 2.88e2, STD_LAPSE = -6.5, GRAVITY = 9.80665, MOL_AIR = 28.9644;   // envphys.h
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double zp = z[i];
        double pa = ENV_SEA_LEVEL;
        if (zp != 0.0)   // HYSTAT(SEA_LEVEL, STD_AIRTMP, STD_LAPSE, z / 1000, GRAVITY, MOL_AIR), envphys.h:205-207
            pa = ENV_SEA_LEVEL * pow(STD_AIRTMP / (STD_AIRTMP + STD_LAPSE * (zp / 1000.0)), GRAVITY * MOL_AIR / (RGAS * STD_LAPSE * 1.e-3));
        const double tak = ta[i] + ENV_FREEZE, tdk = td[i] + ENV_FREEZE;
        int bad = 0;
        double v;
        if (tak < 0 || tdk < 0) {          // iwbt.c:123-126: "ta or td < 0", exit(-1)
            bad = 2;
            v = nan("");
        } else {
            v = env_wetbulb(tak, tdk, pa, tol, &bad) - ENV_FREEZE;
        }
        if (bad) atomicOr(flags, bad);
        __stcs(tw + i, v);
    }
}

// wind.py:175-178: az = arctan2(u, v) * 180 / pi; az[az < 0] += 360
__global__ void wind_direction_kernel(const double* __restrict__ u, const double* __restrict__ v, long long n, double* __restrict__ out) {
    const double pi = 3.141592653589793;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double az = __ddiv_rn(__dmul_rn(atan2(u[i], v[i]), 180.0), pi);
        if (az < 0) az = __dadd_rn(az, 360.0);
        __stcs(out + i, az);
    }
}

// ---- precipitation phase / new-snow density (envphys/snow.py:37-75).  NumPy evaluates every expression below element by element
// in fp64 without contraction: the _rn intrinsics keep that order; pow / exp are the device's (<= 2 ulp from glibc's).
// nasde_model/utils.py:53-71 (check_temperature, t_min = -10, t_max = 0), :12-37 (calc_percent_snow), and
// piecewise_suosong_1999.py:74-83 (density of uncompacted snow)
__device__ __forceinline__ void snow_piecewise(double t, double* pcs, double* rho_ns, double* tsnow) {
    const double tpp = (t < -10.0) ? -10.0 : t;                    // NaN stays NaN
    const double ts = (tpp > 0.0) ? 0.0 : tpp;
    double p = 0.0;
    if (tpp <= -0.5) p = 1.0;
    else if (tpp <= 0.0) p = __dadd_rn(__dmul_rn(-tpp / 0.5, 0.25), 0.75);          // (-0.5, 0]
    else if (tpp <= 1.0) p = __dadd_rn(__dmul_rn(-tpp / 1.0, 0.75), 0.75);          // (0, t_max + 1]
    double ex = __dadd_rn(1.0, __dmul_rn(__dadd_rn(10.0, __dsub_rn(ts, 0.0)) / 10.0, 0.75));
    if (ex > 1.75) ex = 1.75;
    *pcs = p;
    *rho_ns = __dadd_rn(50.0, __dmul_rn(1.7, pow(__dadd_rn(__dsub_rn(tpp, 0.0), 15.0), ex)));
    *tsnow = ts;
}

// model 0 susong1999 (susong_1999.py:29-116), 1 piecewise_susong1999, 2 marks2017 (marks_2017.py:73-207; any_pp: np.any(pp > 0))
__global__ void snow_phase_kernel(const double* __restrict__ tpp, const double* __restrict__ pp, long long n, int model,
                                  const int* __restrict__ any_pp, double* __restrict__ pcs_out, double* __restrict__ rho_out) {
    const int any = *any_pp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double t = tpp[i], p = pp[i];
        double pcs = 0.0, rho = 0.0;
        if (model == 0) {
            // rows [lo, hi): a NaN temperature matches none; cells without precipitation report nothing.  The reference's early
            // return on np.sum(precipitation) == 0 gives the same zeros (precipitation is clamped to >= 0 upstream)
            if (p != 0.0) {
                if (t < -5.0) { pcs = 1.0; rho = 75.0; }
                else if (t < -3.0) { pcs = 1.0; rho = 100.0; }
                else if (t < -1.5) { pcs = 1.0; rho = 150.0; }
                else if (t < -0.5) { pcs = 1.0; rho = 175.0; }
                else if (t < 0.0) { pcs = 0.75; rho = 200.0; }
                else if (t < 0.5) { pcs = 0.25; rho = 250.0; }
            }
        } else if (model == 1) {
            double ts;
            snow_piecewise(t, &pcs, &rho, &ts);
        } else if (any) {
            double ts, rho_orig;
            snow_piecewise(t, &pcs, &rho_orig, &ts);
            const double swe = __dmul_rn(p, pcs);
            if (swe > 0.0) {
                const double rns = rho_orig / 1000.0;
                const double drc = __dmul_rn(__dmul_rn(__dmul_rn(0.026, exp(__dmul_rn(-0.08, __dsub_rn(0.0, ts)))), swe),
                                             exp(__dmul_rn(-21.0, rns)));                                     // :149-150
                const double r1000 = __dmul_rn(rns, 1000.0);
                const double c11 = (r1000 >= 100.0) ? exp(__dmul_rn(-0.046, __dsub_rn(r1000, 100.0))) : 1.0;   // :152-155
                const double drm = __dmul_rn(__dmul_rn(0.01, c11), exp(__dmul_rn(-0.04, __dsub_rn(0.0, ts))));  // :157-159
                rho = __dmul_rn(__dadd_rn(rns, __dmul_rn(__dadd_rn(drc, drm), rns)), 1000.0);                  // :162, :195
            }
        }
        __stcs(pcs_out + i, pcs);
        __stcs(rho_out + i, rho);
    }
}

__global__ void snow_any_kernel(const double* __restrict__ pp, long long n, int* __restrict__ any_pp) {
    bool a = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a |= pp[i] > 0.0;
    if (__any_sync(0xffffffffu, a) && (threadIdx.x & 31) == 0) atomicOr(any_pp, 1);
}

static int launch_snow_phase(cudaStream_t st, const double* d_t, const double* d_pp, long long n, int model, int* d_any, bool scan,
                             double* d_pcs, double* d_rho) {
    if (n <= 0) return SMRFB_OK;
    const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, 148LL * 32);
    if (scan) {
        snow_any_kernel<<<blocks, 256, 0, st>>>(d_pp, n, d_any);
        count_launch();
    }
    snow_phase_kernel<<<blocks, 256, 0, st>>>(d_t, d_pp, n, model, d_any, d_pcs, d_rho);
    count_launch();
    SMRFB_CUDA(cudaGetLastError());
    return SMRFB_OK;
}

// utils.py:409-427: ret = einsum('nj,nj->n', take(values, vtx), wts); ret[any(wts < 0, axis=1)] = fill
__global__ void interp_apply_kernel(const int32_t* __restrict__ vtx, const double* __restrict__ wts, long long n, const double* __restrict__ values,
                                    int npts, int T, double fill, int fill_negative, double* __restrict__ out) {
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        const int v0 = vtx[3 * c], v1 = vtx[3 * c + 1], v2 = vtx[3 * c + 2];
        const double w0 = wts[3 * c], w1 = wts[3 * c + 1], w2 = wts[3 * c + 2];
        const bool outside = fill_negative && ((w0 < 0) || (w1 < 0) || (w2 < 0));
        for (int t = 0; t < T; t++) {
            const double* val = values + (size_t)t * npts;
            double r = __dadd_rn(__dadd_rn(__dmul_rn(__ldg(val + v0), w0), __dmul_rn(__ldg(val + v1), w1)), __dmul_rn(__ldg(val + v2), w2));
            if (outside) r = fill;
            __stcs(out + (size_t)t * n + c, r);
        }
    }
}

struct InterpHandle : HandleBase {
    int32_t* d_vtx = nullptr;
    double* d_wts = nullptr;
    int npts = 0;
    int fill_negative = 1;   // utils.py:424: cells with a negative weight take the fill value
    InterpHandle() : HandleBase(HK_INTERP) {}
    ~InterpHandle() override {
        cudaSetDevice(device);
        cudaFree(d_vtx);
        cudaFree(d_wts);
    }
};

}  // namespace smrfb

using namespace smrfb;

extern "C" {

int smrfb_dew_point(const double* vp, int64_t n, double tol, double* dpt, int device) {
    SMRFB_CHECK(vp && dpt && n >= 0, SMRFB_ERR_ARG, "NULL vapor pressure / output");
    SMRFB_CUDA(cudaSetDevice(device));
    double *d_in = nullptr, *d_out = nullptr;
    const size_t chunk = (size_t)1 << 26;        // 64 Mi cells per piece: 1 GiB of device scratch
    const size_t cap = std::min<size_t>((size_t)n, chunk);
    if (n == 0) return SMRFB_OK;
    SMRFB_CUDA(cudaMalloc((void**)&d_in, cap * sizeof(double)));
    if (cudaMalloc((void**)&d_out, cap * sizeof(double)) != cudaSuccess) {
        cudaFree(d_in);
        set_error("cudaMalloc failed");
        return SMRFB_ERR_CUDA;
    }
    int rc = SMRFB_OK;
    for (size_t o = 0; o < (size_t)n && rc == SMRFB_OK; o += cap) {
        const size_t m = std::min(cap, (size_t)n - o);
        if (cudaMemcpy(d_in, vp + o, m * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) rc = SMRFB_ERR_CUDA;
        if (rc == SMRFB_OK) rc = launch_dew_point(nullptr, d_in, 0, (long long)m, tol, d_out, 0);
        if (rc == SMRFB_OK && cudaMemcpy(dpt + o, d_out, m * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) rc = SMRFB_ERR_CUDA;
    }
    if (rc == SMRFB_ERR_CUDA) set_error("dew point: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_in);
    cudaFree(d_out);
    return rc;
}

int smrfb_dew_point_dev(const double* d_vp, int64_t n, double tol, double* d_dpt, void* stream) {
    SMRFB_CHECK(d_vp && d_dpt && n >= 0, SMRFB_ERR_ARG, "NULL vapor pressure / output");
    return launch_dew_point((cudaStream_t)stream, d_vp, 0, (long long)n, tol, d_dpt, 0);
}

static int launch_wet_bulb(cudaStream_t st, const double* d_ta, const double* d_td, const double* d_z, long long n, double tol, double* d_tw,
                           int* d_flags) {
    if (n <= 0) return SMRFB_OK;
    const int threads = 256;
    const long long blocks = std::min<long long>((n + threads - 1) / threads, 148LL * 32);
    wet_bulb_kernel<<<(unsigned)blocks, threads, 0, st>>>(d_ta, d_td, d_z, n, tol, d_tw, d_flags);
    count_launch();
    SMRFB_CUDA(cudaGetLastError());
    return SMRFB_OK;
}

int smrfb_wet_bulb(const double* ta, const double* td, const double* z, int64_t n, double tol, double* tw, int device) {
    SMRFB_CHECK(ta && td && z && tw && n >= 0, SMRFB_ERR_ARG, "NULL air temperature / dew point / elevation / output");
    SMRFB_CHECK(tol > 0, SMRFB_ERR_ARG, "wet bulb: the tolerance must be positive (the reference iterates while the step exceeds it)");
    SMRFB_CUDA(cudaSetDevice(device));
    if (n == 0) return SMRFB_OK;
    const size_t cap = std::min<size_t>((size_t)n, (size_t)1 << 25);
    double* d = nullptr;
    int* d_flags = nullptr;
    SMRFB_CUDA(cudaMalloc((void**)&d, 4 * cap * sizeof(double)));
    if (cudaMalloc((void**)&d_flags, sizeof(int)) != cudaSuccess) {
        cudaFree(d);
        set_error("cudaMalloc failed");
        return SMRFB_ERR_CUDA;
    }
    cudaMemset(d_flags, 0, sizeof(int));
    int rc = SMRFB_OK, flags = 0;
    for (size_t o = 0; o < (size_t)n && rc == SMRFB_OK; o += cap) {
        const size_t m = std::min(cap, (size_t)n - o);
        if (cudaMemcpy(d, ta + o, m * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(d + cap, td + o, m * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(d + 2 * cap, z + o, m * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
            rc = SMRFB_ERR_CUDA;
        if (rc == SMRFB_OK) rc = launch_wet_bulb(nullptr, d, d + cap, d + 2 * cap, (long long)m, tol, d + 3 * cap, d_flags);
        if (rc == SMRFB_OK && cudaMemcpy(tw + o, d + 3 * cap, m * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) rc = SMRFB_ERR_CUDA;
    }
    if (rc == SMRFB_OK && cudaMemcpy(&flags, d_flags, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) rc = SMRFB_ERR_CUDA;
    if (rc == SMRFB_ERR_CUDA) set_error("wet bulb: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
    cudaFree(d_flags);
    // where the reference prints and calls exit(-1) (iwbt.c:74-77, 123-126) this returns an error; tw is filled regardless
    SMRFB_CHECK(rc != SMRFB_OK || !(flags & 2), SMRFB_ERR_ARG, "wet bulb: ta or td below 0 K at some cell");
    SMRFB_CHECK(rc != SMRFB_OK || !(flags & 1), SMRFB_ERR_NODATA, "wet bulb: failure to converge in 10 iterations at some cell");
    return rc;
}

int smrfb_wet_bulb_dev(const double* d_ta, const double* d_td, const double* d_z, int64_t n, double tol, double* d_tw, int* d_flags,
                       void* stream) {
    SMRFB_CHECK(d_ta && d_td && d_z && d_tw && d_flags && n >= 0, SMRFB_ERR_ARG, "NULL air temperature / dew point / elevation / output / flags");
    SMRFB_CHECK(tol > 0, SMRFB_ERR_ARG, "wet bulb: the tolerance must be positive");
    return launch_wet_bulb((cudaStream_t)stream, d_ta, d_td, d_z, (long long)n, tol, d_tw, d_flags);
}

int smrfb_snow_phase(const double* tpp, const double* pp, int64_t n, int model, double* pcs, double* rho_s, int device) {
    SMRFB_CHECK(tpp && pp && pcs && rho_s && n >= 0, SMRFB_ERR_ARG, "NULL temperature / precipitation / outputs");
    SMRFB_CHECK(model >= 0 && model <= 2, SMRFB_ERR_ARG, "model must be 0 (susong1999), 1 (piecewise_susong1999) or 2 (marks2017)");
    if (n == 0) return SMRFB_OK;
    SMRFB_CUDA(cudaSetDevice(device));
    const size_t cap = std::min<size_t>((size_t)n, (size_t)1 << 25);
    double* d = nullptr;
    int* d_any = nullptr;
    SMRFB_CUDA(cudaMalloc((void**)&d, 4 * cap * sizeof(double)));
    if (cudaMalloc((void**)&d_any, sizeof(int)) != cudaSuccess) {
        cudaFree(d);
        set_error("cudaMalloc failed");
        return SMRFB_ERR_CUDA;
    }
    int rc = SMRFB_OK;
    // marks2017 looks at the WHOLE precipitation array first (np.any(pp > 0), marks_2017.py:118): decided on the host here
    int any = 0;
    if (model == 2)
        for (int64_t i = 0; i < n && !any; i++) any = pp[i] > 0.0;
    if (cudaMemcpy(d_any, &any, sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) rc = SMRFB_ERR_CUDA;
    for (size_t o = 0; o < (size_t)n && rc == SMRFB_OK; o += cap) {
        const size_t m = std::min(cap, (size_t)n - o);
        if (cudaMemcpy(d, tpp + o, m * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(d + cap, pp + o, m * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
            rc = SMRFB_ERR_CUDA;
        if (rc == SMRFB_OK) rc = launch_snow_phase(nullptr, d, d + cap, (long long)m, model, d_any, false, d + 2 * cap, d + 3 * cap);
        if (rc == SMRFB_OK && (cudaMemcpy(pcs + o, d + 2 * cap, m * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
                               cudaMemcpy(rho_s + o, d + 3 * cap, m * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess))
            rc = SMRFB_ERR_CUDA;
    }
    if (rc == SMRFB_ERR_CUDA) set_error("snow phase: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
    cudaFree(d_any);
    return rc;
}

int smrfb_snow_phase_dev(const double* d_tpp, const double* d_pp, int64_t n, int model, double* d_pcs, double* d_rho_s, int* d_any,
                         void* stream) {
    SMRFB_CHECK(d_tpp && d_pp && d_pcs && d_rho_s && d_any && n >= 0, SMRFB_ERR_ARG, "NULL temperature / precipitation / outputs / scratch");
    SMRFB_CHECK(model >= 0 && model <= 2, SMRFB_ERR_ARG, "model must be 0 (susong1999), 1 (piecewise_susong1999) or 2 (marks2017)");
    return launch_snow_phase((cudaStream_t)stream, d_tpp, d_pp, (long long)n, model, d_any, model == 2, d_pcs, d_rho_s);
}

int smrfb_set_consumer(smrfb_handle_t hh, int kind, double tol, void* out2) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    SMRFB_CHECK(kind == 0 || kind == 1, SMRFB_ERR_ARG, "consumer kind must be 0 (none) or 1 (dew point)");
    SMRFB_CHECK(kind == 0 || out2, SMRFB_ERR_ARG, "the consumer needs an output buffer");
    HandleBase* h = reinterpret_cast<HandleBase*>(hh);
    h->consumer = kind;
    h->consumer_tol = tol;
    h->consumer_out = out2;
    return SMRFB_OK;
}

int smrfb_wind_direction(const double* u, const double* v, int64_t n, double* az, int device) {
    SMRFB_CHECK(u && v && az && n >= 0, SMRFB_ERR_ARG, "NULL u / v / output");
    if (n == 0) return SMRFB_OK;
    SMRFB_CUDA(cudaSetDevice(device));
    double *d_u = nullptr, *d_v = nullptr;
    const size_t cap = std::min<size_t>((size_t)n, (size_t)1 << 26);
    SMRFB_CUDA(cudaMalloc((void**)&d_u, cap * sizeof(double)));
    if (cudaMalloc((void**)&d_v, cap * sizeof(double)) != cudaSuccess) {
        cudaFree(d_u);
        set_error("cudaMalloc failed");
        return SMRFB_ERR_CUDA;
    }
    int rc = SMRFB_OK;
    for (size_t o = 0; o < (size_t)n && rc == SMRFB_OK; o += cap) {
        const size_t m = std::min(cap, (size_t)n - o);
        if (cudaMemcpy(d_u, u + o, m * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(d_v, v + o, m * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
            rc = SMRFB_ERR_CUDA;
        if (rc == SMRFB_OK) {
            wind_direction_kernel<<<(unsigned)std::min<size_t>((m + 255) / 256, 148 * 32), 256>>>(d_u, d_v, (long long)m, d_u);
            count_launch();
            if (cudaMemcpy(az + o, d_u, m * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) rc = SMRFB_ERR_CUDA;
        }
    }
    if (rc == SMRFB_ERR_CUDA) set_error("wind direction: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_u);
    cudaFree(d_v);
    return rc;
}

int smrfb_wind_direction_dev(const double* d_u, const double* d_v, int64_t n, double* d_az, void* stream) {
    SMRFB_CHECK(d_u && d_v && d_az && n >= 0, SMRFB_ERR_ARG, "NULL u / v / output");
    if (n == 0) return SMRFB_OK;
    wind_direction_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(d_u, d_v, n, d_az);
    count_launch();
    SMRFB_CUDA(cudaGetLastError());
    return SMRFB_OK;
}

int smrfb_interp_create(int64_t n, const int32_t* vtx, const double* wts, int npts, int fill_negative, int device, smrfb_handle_t* out) {
    SMRFB_CHECK(out, SMRFB_ERR_ARG, "out handle pointer is NULL");
    *out = nullptr;
    SMRFB_CHECK(n > 0 && vtx && wts && npts > 0, SMRFB_ERR_ARG, "need n > 0 cells, vertices, weights and npts > 0");
    for (int64_t i = 0; i < 3 * n; i++) SMRFB_CHECK(vtx[i] >= 0 && vtx[i] < npts, SMRFB_ERR_ARG, "vertex %d outside [0, %d)", vtx[i], npts);
    int ndev = 0;
    SMRFB_CUDA(cudaGetDeviceCount(&ndev));
    SMRFB_CHECK(device >= 0 && device < ndev, SMRFB_ERR_ARG, "device %d out of range (%d visible)", device, ndev);
    InterpHandle* h = new InterpHandle();
    h->device = device;
    h->npts = npts;
    h->fill_negative = fill_negative;
    h->g.ncell = n;
    h->g.ny = 1;
    h->g.nx = (int)std::min<int64_t>(n, 2147483647);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_vtx, (size_t)3 * n * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_wts, (size_t)3 * n * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_vtx, vtx, (size_t)3 * n * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_wts, wts, (size_t)3 * n * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("interp create: %s", cudaGetErrorString(e));
        delete h;
        return SMRFB_ERR_CUDA;
    }
    *out = reinterpret_cast<smrfb_handle_t>(h);
    return SMRFB_OK;
}

static int interp_run(InterpHandle* h, const double* d_values, int T, double fill, double* d_out, cudaStream_t st) {
    const long long n = h->g.ncell;
    interp_apply_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 32), 256, 0, st>>>(h->d_vtx, h->d_wts, n, d_values, h->npts, T, fill, h->fill_negative, d_out);
    count_launch();
    SMRFB_CUDA(cudaGetLastError());
    return SMRFB_OK;
}

int smrfb_interp_apply_dev(smrfb_handle_t hh, const double* d_values, int T, double fill, double* d_out, void* stream) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    InterpHandle* h = reinterpret_cast<InterpHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_INTERP, SMRFB_ERR_ARG, "not an interpolation-weights handle");
    SMRFB_CHECK(d_values && d_out && T > 0, SMRFB_ERR_ARG, "NULL values/out or T <= 0");
    SMRFB_CUDA(cudaSetDevice(h->device));
    return interp_run(h, d_values, T, fill, d_out, stream ? (cudaStream_t)stream : h->stream);
}

int smrfb_interp_apply(smrfb_handle_t hh, const double* values, int T, double fill, double* out) {
    SMRFB_CHECK(hh, SMRFB_ERR_ARG, "NULL handle");
    InterpHandle* h = reinterpret_cast<InterpHandle*>(hh);
    SMRFB_CHECK(h->kind == HK_INTERP, SMRFB_ERR_ARG, "not an interpolation-weights handle");
    SMRFB_CHECK(values && out && T > 0, SMRFB_ERR_ARG, "NULL values/out or T <= 0");
    SMRFB_CUDA(cudaSetDevice(h->device));
    SMRFB_PROPAGATE(ensure_cap((void**)&h->d_data, &h->data_cap, (size_t)T * h->npts * sizeof(double)));
    SMRFB_CUDA(cudaMemcpyAsync(h->d_data, values, (size_t)T * h->npts * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return run_host_batches(h, T, 1, out, [&](int t0, int tn, double* d_out, cudaStream_t st) {
        return interp_run(h, h->d_data + (size_t)t0 * h->npts, tn, fill, d_out, st);
    });
}

}  // extern "C"
