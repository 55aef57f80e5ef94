// common.cu -- handles, errors, geometry upload, the per-timestep station-prep kernel, clamp.
#include <cstdarg>

#include "common.cuh"

namespace smrfb {

static thread_local std::string t_err;
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_err = buf;
}

HandleBase::~HandleBase() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    if (copy_stream) cudaStreamSynchronize(copy_stream);
    cudaFree(g.gx);
    cudaFree(g.gy);
    cudaFree(g.dem);
    cudaFree(d_mx);
    cudaFree(d_my);
    cudaFree(d_mz);
    cudaFree(d_fitmask);
    cudaFree(d_rec);
    cudaFree(d_valid);
    cudaFree(d_data);
    cudaFree(d_pv);
    cudaFree(d_flags);
    cudaFree(d_ring);
    cudaFree(d_ring2);
    if (stream) cudaStreamDestroy(stream);
    if (copy_stream) cudaStreamDestroy(copy_stream);
}

int ensure_cap(void** p, size_t* cap, size_t need_bytes) {
    if (*cap >= need_bytes && *p) return SMRFB_OK;
    if (*p) SMRFB_CUDA(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    SMRFB_CUDA(cudaMalloc(p, need_bytes));
    *cap = need_bytes;
    return SMRFB_OK;
}

static int upload(double** dst, const double* src, size_t n) {
    SMRFB_CUDA(cudaMalloc((void**)dst, n * sizeof(double)));
    SMRFB_CUDA(cudaMemcpy(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice));
    return SMRFB_OK;
}

int upload_geom(HandleBase* h, int nsta, const double* mx, const double* my, const double* mz, int ny, int nx,
                int grid_kind, const double* gx, const double* gy, const double* dem) {
    SMRFB_CHECK(nsta > 0 && mx && my, SMRFB_ERR_ARG, "need nsta > 0 and station coordinates");
    SMRFB_CHECK(ny > 0 && nx > 0 && gx && gy, SMRFB_ERR_ARG, "need ny, nx > 0 and grid coordinates");
    SMRFB_CHECK(grid_kind == 0 || grid_kind == 1, SMRFB_ERR_ARG, "grid_kind must be 0 (mesh) or 1 (points)");
    int ndev = 0;
    SMRFB_CUDA(cudaGetDeviceCount(&ndev));
    SMRFB_CHECK(h->device >= 0 && h->device < ndev, SMRFB_ERR_ARG, "device %d out of range (%d visible)", h->device, ndev);
    SMRFB_CUDA(cudaSetDevice(h->device));
    SMRFB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    SMRFB_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    h->S = nsta;
    h->g.ny = ny;
    h->g.nx = nx;
    h->g.kind = grid_kind;
    h->g.ncell = (long long)ny * nx;
    SMRFB_PROPAGATE(upload(&h->d_mx, mx, nsta));
    SMRFB_PROPAGATE(upload(&h->d_my, my, nsta));
    if (mz) SMRFB_PROPAGATE(upload(&h->d_mz, mz, nsta));
    size_t ngx = grid_kind == 0 ? (size_t)nx : (size_t)h->g.ncell;
    size_t ngy = grid_kind == 0 ? (size_t)ny : (size_t)h->g.ncell;
    SMRFB_PROPAGATE(upload(&h->g.gx, gx, ngx));
    SMRFB_PROPAGATE(upload(&h->g.gy, gy, ngy));
    if (dem) SMRFB_PROPAGATE(upload(&h->g.dem, dem, (size_t)h->g.ncell));
    SMRFB_CUDA(cudaMalloc((void**)&h->d_flags, sizeof(int)));
    SMRFB_CUDA(cudaMemset(h->d_flags, 0, sizeof(int)));
    return SMRFB_OK;
}

RecLayout make_layout(int S, int k_align, int maxm, bool mma_stride) {
    RecLayout L;
    L.S = S;
    L.S_pad = (S + k_align - 1) / k_align * k_align;
    L.maxm = maxm;
    int rs = L.S_pad + 2 + (maxm > 0 ? (4 + maxm + 1) / 2 : 1);
    if (rs & 1) rs++;
    if (mma_stride)
        while ((rs & 15) != 4 && (rs & 15) != 12) rs += 2;
    L.RS = rs;
    return L;
}

// ------------------------------------------------------------------------------------------
// Station prep: one CTA per timestep.
//   trend (A.1 of SURVEY.md; idw.py:119-136, dk.py:146-160, grid.py:188-201):
//     least-squares line through (mz_s, d_s) over the fit set, centred two-pass closed form
//     (agrees with np.polyfit's scaled-Vandermonde SVD to ~3e-14 relative);
//     zero-variance elevations (one valid station) -> polyfit's minimum-norm solution
//     p0 = mean(d)/(2 z), p1 = mean(d)/2;  sign constraint zeroes BOTH coefficients.
// ------------------------------------------------------------------------------------------
constexpr int PREP_THREADS = 128;

__device__ __forceinline__ double block_sum(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < PREP_THREADS / 32; i++) r += sh[i];
    return r;
}

__global__ void __launch_bounds__(PREP_THREADS)
prep_steps_kernel(const double* __restrict__ data, int T, int S, RecLayout L, const double* __restrict__ mz,
                  const uint8_t* __restrict__ fitmask, int detrend, int slope_flag,
                  const double* __restrict__ pv_in, double* __restrict__ pv_out, double* __restrict__ rec,
                  uint8_t* __restrict__ valid, int* __restrict__ flags) {
    __shared__ double sh[PREP_THREADS / 32];
    const int t = blockIdx.x;
    double* r = rec + (size_t)t * L.RS;
    if (t >= T) {
        for (int s = threadIdx.x; s < L.RS; s += PREP_THREADS) r[s] = 0.0;
        return;
    }
    const double* row = data + (size_t)t * S;
    double p0 = 0.0, p1 = 0.0;
    double cnt_valid = 0.0;
    {
        double c = 0.0;
        for (int s = threadIdx.x; s < S; s += PREP_THREADS) c += isnan(row[s]) ? 0.0 : 1.0;
        cnt_valid = block_sum(c, sh);
    }
    if (detrend) {
        if (pv_in) {
            p0 = pv_in[2 * t];
            p1 = pv_in[2 * t + 1];
        } else {
            double n = 0.0, sz = 0.0, sd = 0.0;
            for (int s = threadIdx.x; s < S; s += PREP_THREADS) {
                double d = row[s];
                bool fit = !isnan(d) && (!fitmask || fitmask[s]);
                if (fit) {
                    n += 1.0;
                    sz += mz[s];
                    sd += d;
                }
            }
            n = block_sum(n, sh);
            sz = block_sum(sz, sh);
            sd = block_sum(sd, sh);
            if (n > 0.0) {
                double zbar = sz / n, dbar = sd / n;
                double szz = 0.0, szd = 0.0;
                for (int s = threadIdx.x; s < S; s += PREP_THREADS) {
                    double d = row[s];
                    bool fit = !isnan(d) && (!fitmask || fitmask[s]);
                    if (fit) {
                        double dz = mz[s] - zbar;
                        szz += dz * dz;
                        szd += dz * (d - dbar);
                    }
                }
                szz = block_sum(szz, sh);
                szd = block_sum(szd, sh);
                // np.polyfit solves the column-scaled Vandermonde system [z/|z|, 1/sqrt(n)] with rcond = n * eps: its
                // singular values are sqrt(1 -+ |c|), c = sum(z) / (|z| sqrt(n)), and 1 - c^2 = szz / sum(z^2), so the
                // fit is rank-deficient when szz < (rcond (1 + |c|))^2 sum(z^2) (all fitted stations at one elevation,
                // exactly or up to rounding); lstsq then returns the minimum-norm solution along (1, sign c).
                double szsq = 0.0, szdd = 0.0;
                for (int s = threadIdx.x; s < S; s += PREP_THREADS) {
                    double d = row[s];
                    bool fit = !isnan(d) && (!fitmask || fitmask[s]);
                    if (fit) {
                        szsq += mz[s] * mz[s];
                        szdd += mz[s] * d;
                    }
                }
                szsq = block_sum(szsq, sh);
                szdd = block_sum(szdd, sh);
                const double znorm = sqrt(szsq), rn = sqrt(n);
                const double cc = znorm > 0.0 ? sz / (znorm * rn) : 0.0;
                const double rc = n * 2.220446049250313e-16 * (1.0 + fabs(cc));
                if (szz > rc * rc * szsq) {
                    p0 = szd / szz;
                    p1 = dbar - p0 * zbar;
                } else if (znorm > 0.0) {
                    const double sg = cc < 0.0 ? -1.0 : 1.0;
                    const double xs = (szdd / znorm + sg * sd / rn) / (2.0 * (1.0 + fabs(cc)));
                    p0 = xs / znorm;
                    p1 = sg * xs / rn;
                } else {  // every fitted elevation is exactly 0: the slope column vanishes
                    p0 = 0.0;
                    p1 = dbar;
                }
                if ((slope_flag == 1 && p0 < 0.0) || (slope_flag == -1 && p0 > 0.0)) {
                    p0 = 0.0;
                    p1 = 0.0;
                }
            } else {
                p0 = p1 = nan("");
                if (threadIdx.x == 0) atomicOr(flags, 2);
            }
        }
    }
    if (cnt_valid == 0.0 && threadIdx.x == 0) atomicOr(flags, 1);
    for (int s = threadIdx.x; s < L.S_pad; s += PREP_THREADS) {
        double res = 0.0;
        if (s < S) {
            double d = row[s];
            bool ok = !isnan(d);
            if (valid) valid[(size_t)t * S + s] = ok ? 1 : 0;
            if (ok) res = detrend ? __dsub_rn(d, __dadd_rn(__dmul_rn(mz[s], p0), p1)) : d;
            else if (L.keep_nan) res = d;
        }
        r[s] = res;
    }
    if (threadIdx.x == 0) {
        r[L.meta()] = p0;
        r[L.meta() + 1] = p1;
        if (pv_out) {
            pv_out[2 * t] = p0;
            pv_out[2 * t + 1] = p1;
        }
        int* mi = reinterpret_cast<int*>(r + L.meta() + 2);
        const int nints = 2 * (L.RS - L.meta() - 2);
        int nm = 0;
        if (L.maxm > 0) {
            for (int s = 0; s < S; s++) {
                if (isnan(row[s])) {
                    if (nm < L.maxm) mi[4 + nm] = s * L.miss_scale;
                    nm++;
                }
            }
            for (int j = 1; j < 4; j++) mi[j] = 0;
            for (int j = (nm < L.maxm ? nm : L.maxm) + 4; j < nints; j++) mi[j] = L.miss_fill;
        } else {
            nm = S - (int)cnt_valid;
            for (int j = 1; j < nints; j++) mi[j] = 0;
        }
        mi[0] = nm;
    }
}

int launch_prep(cudaStream_t st, const double* d_data, int T, int T_pad, int S, const RecLayout& L, const double* d_mz,
                const uint8_t* d_fitmask, int detrend, int slope_flag, const double* d_pv_in, double* d_pv_out,
                double* d_rec, uint8_t* d_valid, int* d_flags) {
    SMRFB_CHECK(!detrend || d_mz || d_pv_in, SMRFB_ERR_ARG, "detrend needs station elevations (mz)");
    prep_steps_kernel<<<T_pad, PREP_THREADS, 0, st>>>(d_data, T, S, L, d_mz, d_fitmask, detrend, slope_flag, d_pv_in,
                                                      d_pv_out, d_rec, d_valid, d_flags);
    count_launch();
    SMRFB_CUDA(cudaGetLastError());
    return SMRFB_OK;
}

// ------------------------------------------------------------------------------------------
__global__ void clamp_kernel(double* x, long long n, double vmin, double vmax) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) x[i] = clamp_nan_keep(x[i], vmin, vmax);
}

}  // namespace smrfb

using namespace smrfb;

extern "C" {

const char* smrfb_last_error(void) { return t_err.c_str(); }
int smrfb_version(void) { return 100; }
int64_t smrfb_launch_count(void) { return (int64_t)g_launches.load(); }

int smrfb_device_count(int* count) {
    SMRFB_CHECK(count, SMRFB_ERR_ARG, "count is NULL");
    SMRFB_CUDA(cudaGetDeviceCount(count));
    return SMRFB_OK;
}

int smrfb_destroy(smrfb_handle_t h) {
    if (!h) return SMRFB_OK;
    delete reinterpret_cast<HandleBase*>(h);
    return SMRFB_OK;
}

int smrfb_synchronize(smrfb_handle_t h) {
    SMRFB_CHECK(h, SMRFB_ERR_ARG, "NULL handle");
    HandleBase* b = reinterpret_cast<HandleBase*>(h);
    SMRFB_CUDA(cudaSetDevice(b->device));
    SMRFB_CUDA(cudaStreamSynchronize(b->stream));
    SMRFB_CUDA(cudaStreamSynchronize(b->copy_stream));
    return SMRFB_OK;
}

int smrfb_set_output_dtype(smrfb_handle_t h, int dtype) {
    SMRFB_CHECK(h, SMRFB_ERR_ARG, "NULL handle");
    SMRFB_CHECK(dtype == 0 || dtype == 1, SMRFB_ERR_ARG, "dtype must be 0 (float64) or 1 (float32)");
    reinterpret_cast<HandleBase*>(h)->out_f32 = dtype;
    return SMRFB_OK;
}

int smrfb_host_alloc(size_t bytes, void** out) {
    SMRFB_CHECK(out, SMRFB_ERR_ARG, "out is NULL");
    *out = nullptr;
    // portable (pinned for every device) and mapped (kernels may address it directly: smrfb_set_min_max)
    SMRFB_CUDA(cudaHostAlloc(out, std::max<size_t>(bytes, 1), cudaHostAllocPortable | cudaHostAllocMapped));
    return SMRFB_OK;
}

int smrfb_host_free(void* p) {
    if (p) SMRFB_CUDA(cudaFreeHost(p));
    return SMRFB_OK;
}

int smrfb_set_min_max(double* data, int64_t n, double vmin, double vmax, int device) {
    SMRFB_CHECK(data || n == 0, SMRFB_ERR_ARG, "data is NULL");
    if (n == 0) return SMRFB_OK;
    SMRFB_CUDA(cudaSetDevice(device));
    const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
    // Arrays that came out of this library's pinned pool (smrfb_host_alloc; the drop-in classes return those) are
    // mapped into the device's address space: the kernel clamps them in place over PCIe -- no device allocation,
    // no staging copies.  Anything else (ordinary pageable NumPy memory) is staged through a device buffer.
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, data) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
        clamp_kernel<<<blocks, 256>>>(reinterpret_cast<double*>(at.devicePointer), n, vmin, vmax);
        count_launch();
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            set_error("smrfb_set_min_max: %s", cudaGetErrorString(e));
            return SMRFB_ERR_CUDA;
        }
        return SMRFB_OK;
    }
    cudaGetLastError();
    double* d = nullptr;
    SMRFB_CUDA(cudaMalloc((void**)&d, (size_t)n * sizeof(double)));
    cudaError_t e = cudaMemcpy(d, data, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        clamp_kernel<<<blocks, 256>>>(d, n, vmin, vmax);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(data, d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) {
        set_error("smrfb_set_min_max: %s", cudaGetErrorString(e));
        return SMRFB_ERR_CUDA;
    }
    return SMRFB_OK;
}

}  // extern "C"
