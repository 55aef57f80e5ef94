// common.cuh -- shared host/device plumbing of libsmrfb (handles, errors, step records, PTX helpers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/smrfb.h"

namespace smrfb {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define SMRFB_CUDA(call)                                                                          \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            smrfb::set_error("%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return SMRFB_ERR_CUDA;                                                                \
        }                                                                                         \
    } while (0)

#define SMRFB_CHECK(cond, code, ...)       \
    do {                                   \
        if (!(cond)) {                     \
            smrfb::set_error(__VA_ARGS__); \
            return (code);                 \
        }                                  \
    } while (0)

#define SMRFB_PROPAGATE(call)          \
    do {                               \
        int rc__ = (call);             \
        if (rc__ != SMRFB_OK) return rc__; \
    } while (0)

// ---------------------------------------------------------------- geometry (device pointers)
struct Geom {
    int ny = 0, nx = 0, kind = 0;
    long long ncell = 0;
    double* gx = nullptr;   // kind 0: [nx]; kind 1: [ncell]
    double* gy = nullptr;   // kind 0: [ny]; kind 1: [ncell]
    double* dem = nullptr;  // [ncell] or null
};

__device__ __forceinline__ void cell_xy(const Geom& g, long long c, double& X, double& Y) {
    if (g.kind == 0) {
        long long i = c / g.nx;
        int j = (int)(c - i * g.nx);
        X = __ldg(g.gx + j);
        Y = __ldg(g.gy + i);
    } else {
        X = __ldg(g.gx + c);
        Y = __ldg(g.gy + c);
    }
}

// ---------------------------------------------------------------- per-timestep station record
// One record per timestep, written by prep_steps_kernel, consumed by the distribute kernels:
//   rec[t][0 .. S_pad)            residuals d_s - (mz_s*p0 + p1); 0 for missing stations and padding
//   rec[t][S_pad + 0]             p0 (slope)            rec[t][S_pad + 1]   p1 (intercept)
//   rec[t][S_pad + 2 ...]         int32 array: [0] = number of missing stations, [1..3] = 0, [4 + j] = (index of
//                                 the j-th missing station) * miss_scale for j < min(nmiss, maxm); the unused
//                                 slots hold miss_fill (a consumer reads whole quads with 128-bit loads and
//                                 points the fill at a zero weight row); maxm is a multiple of 4
// RS (record stride, doubles) is chosen = 4 or 12 (mod 16) so that DMMA A-fragment loads from a
// staged record block are shared-memory bank-conflict free.
// keep_nan: missing stations keep NaN in the residual slot (GRID: NaN propagates like griddata's).
struct RecLayout {
    int S = 0, S_pad = 0, RS = 0, maxm = 0, keep_nan = 0, miss_scale = 1, miss_fill = 0;
    __host__ __device__ int meta() const { return S_pad; }
};
RecLayout make_layout(int S, int k_align, int maxm, bool mma_stride);

enum HandleKind { HK_IDW = 0x1d0, HK_DK = 0xd4, HK_GRID = 0x621d, HK_INTERP = 0x17e4 };

struct HandleBase {
    int kind;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    Geom g;
    int S = 0;
    double *d_mx = nullptr, *d_my = nullptr, *d_mz = nullptr;
    uint8_t* d_fitmask = nullptr;  // optional [S]
    // scratch for prepared step records
    double* d_rec = nullptr;
    size_t rec_cap = 0;          // doubles
    uint8_t* d_valid = nullptr;  // [T][S]
    size_t valid_cap = 0;
    double* d_data = nullptr;    // staged station rows (host API)
    size_t data_cap = 0;
    double* d_pv = nullptr;      // [T][2]
    size_t pv_cap = 0;
    int* d_flags = nullptr;      // [1] device-side status flags (bit0: a step had no valid station)
    double* d_ring = nullptr;    // host-API output ring (up to 2 slots), kept across calls
    size_t ring_cap = 0;
    int out_f32 = 0;             // smrfb_set_output_dtype: 1 = outputs are written as float32 (the reference's NetCDF type)
    // smrfb_set_consumer: a second output computed from every distributed sub-batch while it is still in the device ring
    int consumer = 0;            // 0 none, 1 dew point (dewpt.c) of the distributed vapor pressure
    double consumer_tol = 0.01;
    void* consumer_out = nullptr;   // host [T][ncell], element type as out
    double* d_ring2 = nullptr;
    size_t ring2_cap = 0;
    explicit HandleBase(int k) : kind(k) {}
    virtual ~HandleBase();
};

int upload_geom(HandleBase* h, int nsta, const double* mx, const double* my, const double* mz, int ny, int nx,
                int grid_kind, const double* gx, const double* gy, const double* dem);
int ensure_cap(void** p, size_t* cap, size_t need_bytes);
// consumers.cu: dew point (C) of n vapor pressures (Pa); element types float64 / float32 per flag
int launch_dew_point(cudaStream_t st, const void* d_ea, int in_f32, long long n, double tol, void* d_out, int out_f32);

// Launch the station-prep kernel: regression, residuals, missing lists for T steps.
//   fit_mode 0: residuals for every non-NaN station (IDW, GRID);   fit over (valid & fitmask)
//   want_miss: fill the missing-station list
int launch_prep(cudaStream_t st, const double* d_data, int T, int T_pad, int S, const RecLayout& L, const double* d_mz,
                const uint8_t* d_fitmask, int detrend, int slope_flag, const double* d_pv_in, double* d_pv_out,
                double* d_rec, uint8_t* d_valid, int* d_flags);

// Host-buffer driver shared by the three methods: the T steps are cut into sub-batches whose output
// fits one slot of a 2-slot device ring (SMRFB_RING_BYTES per slot, default 8 GiB); the D2H copy of
// one sub-batch (copy_stream) overlaps the kernels of the next (stream).  run(t0, tn, d_out, stream)
// enqueues the kernels for steps [t0, t0+tn).  Synchronises before returning.
template <class RunFn>
int run_host_batches(HandleBase* h, int T, int step_align, double* out_, RunFn run) {
    const size_t esz = h->out_f32 ? sizeof(float) : sizeof(double);      // bytes per output element
    const size_t ncell = ((size_t)h->g.ncell * esz + sizeof(double) - 1) / sizeof(double);   // doubles per step in the ring
    char* out = reinterpret_cast<char*>(out_);
    size_t budget = (size_t)8 << 30;
    {   // at least one full staged block of steps per slot when memory allows (large grids)
        size_t free_b = 0, total_b = 0;
        const size_t want = ncell * sizeof(double) * (size_t)step_align;
        if (want > budget && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && want <= (free_b + h->ring_cap) / 3)
            budget = want;
    }
    if (const char* e = getenv("SMRFB_RING_BYTES")) budget = (size_t)strtoull(e, nullptr, 10);
    int tb = (int)std::max<size_t>(1, std::min<size_t>((size_t)T, budget / (ncell * sizeof(double))));
    if (tb > step_align) tb = tb / step_align * step_align;
    // a batch that fits one slot whole would run kernel, then copy: cut it in two (>= 32 steps each, so long batches
    // stay on the tensor-core IDW kernel) and the copy of the first half overlaps the kernels of the second
    if (tb >= T && T >= 64) tb = ((T + 1) / 2 + step_align - 1) / step_align * step_align;
    const int nslots = tb >= T ? 1 : 2;
    const size_t slot_bytes = (size_t)tb * ncell * sizeof(double);
    if (h->ring_cap < slot_bytes * nslots) {
        if (h->d_ring) cudaFree(h->d_ring);
        h->d_ring = nullptr;
        h->ring_cap = 0;
        if (cudaMalloc((void**)&h->d_ring, slot_bytes * nslots) != cudaSuccess) {
            set_error("cudaMalloc of the %zu-byte output ring failed", slot_bytes * nslots);
            return SMRFB_ERR_CUDA;
        }
        h->ring_cap = slot_bytes * nslots;
    }
    double* ring[2] = {h->d_ring, nslots > 1 ? h->d_ring + (size_t)tb * ncell : nullptr};
    double* ring2[2] = {nullptr, nullptr};
    char* out2 = reinterpret_cast<char*>(h->consumer_out);
    if (h->consumer) {
        if (h->ring2_cap < slot_bytes * nslots) {
            if (h->d_ring2) cudaFree(h->d_ring2);
            h->d_ring2 = nullptr;
            h->ring2_cap = 0;
            if (cudaMalloc((void**)&h->d_ring2, slot_bytes * nslots) != cudaSuccess) {
                set_error("cudaMalloc of the %zu-byte consumer ring failed", slot_bytes * nslots);
                return SMRFB_ERR_CUDA;
            }
            h->ring2_cap = slot_bytes * nslots;
        }
        ring2[0] = h->d_ring2;
        ring2[1] = nslots > 1 ? h->d_ring2 + (size_t)tb * ncell : nullptr;
    }
    cudaEvent_t done[2], copied[2];
    for (int i = 0; i < 2; i++) {
        SMRFB_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        SMRFB_CUDA(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
    }
    int rc = SMRFB_OK, slot = 0;
    for (int t0 = 0; t0 < T && rc == SMRFB_OK; t0 += tb, slot ^= 1) {
        const int tn = std::min(tb, T - t0);
        if (t0 >= 2 * tb) cudaStreamWaitEvent(h->stream, copied[slot], 0);   // ring slot free again
        rc = run(t0, tn, ring[slot], h->stream);
        if (rc == SMRFB_OK && h->consumer == 1)
            rc = launch_dew_point(h->stream, ring[slot], h->out_f32, (long long)tn * h->g.ncell, h->consumer_tol, ring2[slot], h->out_f32);
        if (rc != SMRFB_OK) break;
        cudaEventRecord(done[slot], h->stream);
        cudaStreamWaitEvent(h->copy_stream, done[slot], 0);
        if (cudaMemcpyAsync(out + (size_t)t0 * (size_t)h->g.ncell * esz, ring[slot], (size_t)tn * (size_t)h->g.ncell * esz,
                            cudaMemcpyDeviceToHost, h->copy_stream) != cudaSuccess) {
            set_error("D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = SMRFB_ERR_CUDA;
            break;
        }
        if (h->consumer && cudaMemcpyAsync(out2 + (size_t)t0 * (size_t)h->g.ncell * esz, ring2[slot], (size_t)tn * (size_t)h->g.ncell * esz,
                                           cudaMemcpyDeviceToHost, h->copy_stream) != cudaSuccess) {
            set_error("D2H copy of the consumer output failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = SMRFB_ERR_CUDA;
            break;
        }
        cudaEventRecord(copied[slot], h->copy_stream);
    }
    cudaError_t e1 = cudaStreamSynchronize(h->stream);
    cudaError_t e2 = cudaStreamSynchronize(h->copy_stream);
    for (int i = 0; i < 2; i++) {
        cudaEventDestroy(done[i]);
        cudaEventDestroy(copied[i]);
    }
    if (rc == SMRFB_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) {
        set_error("distribute failed on the device: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        rc = SMRFB_ERR_CUDA;
    }
    return rc;
}

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine (SASS: UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}

// fp64 tensor-core MMA, D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4.
// lane holds a = A[lane>>2][lane&3], b = B[lane&3][lane>>2], c0/c1 = C[lane>>2][(lane&3)*2 + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Output element store: fp64, or float32 when the handle was switched with smrfb_set_output_dtype (the flag is
// uniform across the grid).  Streaming (evict-first) stores: the output is written once and never read back.
__device__ __forceinline__ void store_out(double* base, size_t idx, double v, int f32) {
    if (f32) __stcs(reinterpret_cast<float*>(base) + idx, (float)v);
    else __stcs(base + idx, v);
}
// two adjacent elements, idx even and the row 16-byte (f32: 8-byte) aligned
__device__ __forceinline__ void store_out2(double* base, size_t idx, double v0, double v1, int f32) {
    if (f32) __stcs(reinterpret_cast<float2*>(reinterpret_cast<float*>(base) + idx), make_float2((float)v0, (float)v1));
    else __stcs(reinterpret_cast<double2*>(base + idx), make_double2(v0, v1));
}

__device__ __forceinline__ double clamp_nan_keep(double v, double vmin, double vmax) {
    // utils.py:137-161: x<=min -> min, then x>=max -> max; NaN fails both tests and stays NaN
    if (v <= vmin) v = vmin;
    if (v >= vmax) v = vmax;
    return v;
}

// 1/x for positive normal x: MUFU.RCP64H seed (measured ~2^-10 relative) + Newton steps; no special cases.
// Two steps give <= ~4e-13 relative (weights: far inside the 1e-9 contract, and numerator and denominator
// use the same weight values); the quotient in the epilogue adds a residual correction to reach ~1 ulp.
template <int STEPS>
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#pragma unroll
    for (int i = 0; i < STEPS; i++) r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

}  // namespace smrfb
