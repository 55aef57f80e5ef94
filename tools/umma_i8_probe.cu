// umma_i8_probe.cu -- tcgen05.mma kind::i8 on sm_100a: operand layouts, mixed u8 x s8, A from TMEM, issue rate.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_i8_probe tools/umma_i8_probe.cu && /tmp/umma_i8_probe
//
// D[128 x N] (s32, TMEM: lane = row, column = n) = A[128 x K] * B[N x K]^T, both K-major, no swizzle.
// Shared-memory operands use the canonical interleaved layout: 8 rows x 16 bytes core matrices;
//   offset(row, k) = (k/16)*LBO + (row/8)*SBO + (row%8)*16 + k%16.
// TMEM A operand (hypothesis under test): row -> lane, k -> column k/4, byte k%4.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    return d;                 // base offset 0, layout type 0 = no swizzle
}

__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_signed, int b_signed) {
    return (2u << 4)                       // D format s32
           | ((uint32_t)a_signed << 7)     // A: 0 = u8, 1 = s8
           | ((uint32_t)b_signed << 10)    // B
           | (0u << 15) | (0u << 16)       // K-major A and B
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(n));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

constexpr int M = 128, N = 48, K = 64;       // two K = 32 steps
constexpr int LBO_A = (M / 8) * 128, LBO_B = (N / 8) * 128, SBO = 128;

// mode 0: A from shared memory; mode 1: A from TMEM.  a_signed / b_signed pick the operand formats.
__global__ void __launch_bounds__(128) probe(const uint8_t* A, const uint8_t* B, int32_t* D, int mode, int a_signed, int b_signed,
                                           int perf_iters, int perf_n, long long* clk, int nacc = 1) {
    __shared__ __align__(128) uint8_t sA[M * K];
    __shared__ __align__(128) uint8_t sB[256 * K];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    // operands -> canonical shared layout
    for (int i = tid; i < M * K; i += 128) {
        int row = i / K, k = i % K;
        sA[(k / 16) * LBO_A + (row / 8) * SBO + (row % 8) * 16 + k % 16] = A[i];
    }
    for (int i = tid; i < 256 * K; i += 128) sB[i] = 0;
    __syncthreads();
    for (int i = tid; i < N * K; i += 128) {
        int row = i / K, k = i % K;
        sB[(k / 16) * LBO_B + (row / 8) * SBO + (row % 8) * 16 + k % 16] = B[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tbase_s;
    const uint32_t acol = 256;     // A operand columns in TMEM: [256, 256 + K/4)
    if (mode == 1) {               // row = thread, k -> word k/4, byte k%4
        for (int kb = 0; kb < K / 32; kb++) {
            uint32_t v[8];
            for (int j = 0; j < 8; j++) {
                uint32_t w = 0;
                for (int b = 0; b < 4; b++) w |= (uint32_t)A[tid * K + kb * 32 + j * 4 + b] << (8 * b);
                v[j] = w;
            }
            st8(tbase + ((uint32_t)(warp * 32) << 16) + acol + kb * 8, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t idesc = make_idesc(M, N, a_signed, b_signed);
    if (tid == 0) {
        for (int kb = 0; kb < K / 32; kb++) {
            const uint64_t bd = make_desc(smem_u32(sB) + kb * 2 * LBO_B, LBO_B, SBO);
            if (mode == 0) mma_ss(tbase, make_desc(smem_u32(sA) + kb * 2 * LBO_A, LBO_A, SBO), bd, idesc, kb > 0);
            else mma_ts(tbase, tbase + acol + kb * 8, bd, idesc, kb > 0);
        }
        commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        ld8(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int j = 0; j < 8; j++) D[tid * N + c0 + j] = (int32_t)v[j];
    }
    // issue-rate test: perf_iters MMAs of N = perf_n back to back on one accumulator
    if (perf_iters > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t idp = make_idesc(M, perf_n, 0, 0);
        const int lbo = (perf_n / 8) * 128;
        long long t0 = clock64();
        if (warp == 0) {
            // the whole warp runs the loop (descriptors stay in uniform registers); one elected lane issues
            uint32_t leader;
            asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(leader));
            uint64_t bd[2], ad[2];
            uint32_t at[2], dc[4];
            for (int i = 0; i < 2; i++) {
                bd[i] = make_desc(smem_u32(sB) + i * 2 * lbo, lbo, SBO);
                ad[i] = make_desc(smem_u32(sA) + i * 2 * LBO_A, LBO_A, SBO);
                at[i] = tbase + acol + i * 8;
            }
            for (int i = 0; i < 4; i++) dc[i] = tbase + (uint32_t)((i % nacc) * perf_n);
            if (mode == 0) {
                for (int it = 0; it < perf_iters; it += 4) {
                    if (leader) {
                        mma_ss(dc[0], ad[0], bd[0], idp, 1);
                        mma_ss(dc[1], ad[1], bd[1], idp, 1);
                        mma_ss(dc[2], ad[0], bd[0], idp, 1);
                        mma_ss(dc[3], ad[1], bd[1], idp, 1);
                    }
                }
            } else {
                for (int it = 0; it < perf_iters; it += 4) {
                    if (leader) {
                        mma_ts(dc[0], at[0], bd[0], idp, 1);
                        mma_ts(dc[1], at[1], bd[1], idp, 1);
                        mma_ts(dc[2], at[0], bd[0], idp, 1);
                        mma_ts(dc[3], at[1], bd[1], idp, 1);
                    }
                }
            }
            if (leader) commit(&bar);
        }
        mbar_wait(&bar, 1);
        long long t1 = clock64();
        if (tid == 0) clk[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

// Issue pattern of the IDW kernel: 7 x 7 A column blocks in TMEM, 8 B planes x 7 K blocks in 86 KB of shared memory,
// passes of (g + 1) x 7 MMAs alternating between two accumulators.  Measures clk per MMA without any consumer.
__global__ void __launch_bounds__(128) probe_cycle(int n, int iters, long long* clk, int commit_each, int signed_top) {
    extern __shared__ __align__(128) uint8_t dynB[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t bar2[2];
    __shared__ uint32_t tbase_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_init(&bar2[0], 1);
        mbar_init(&bar2[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    const int lbo = (n / 8) * 128, kstep = 2 * lbo, plane = 7 * kstep;
    for (int i = tid; i < 8 * plane; i += 128) dynB[i] = (uint8_t)(i * 7);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tbase_s;
    const uint32_t idp = make_idesc(128, n, 0, 0), ids = make_idesc(128, n, 0, signed_top);
    long long t0 = clock64();
    int count = 0;
    if (warp == 0) {
        uint32_t leader;
        asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(leader));
        const uint64_t bdesc = make_desc(smem_u32(dynB), lbo, SBO);
        const uint32_t acol = tbase + 2 * n;
        for (int it = 0; it < iters; it++) {
            for (int gg = 6; gg >= 0; gg--) {
                const uint32_t dcol = tbase + (gg & 1) * n;
                if (leader) {
                    for (int i = 0; i <= gg; i++) {
                        const int j = gg - i;
#pragma unroll
                        for (int kk = 0; kk < 7; kk++)
                            mma_ts(dcol, acol + (i * 7 + kk) * 8, bdesc + (uint64_t)((j * plane + kk * kstep) >> 4), j == 0 ? ids : idp, (i | kk) != 0);
                    }
                    if (commit_each) commit(&bar2[gg & 1]);
                }
                __syncwarp();
                count += (gg + 1) * 7;
            }
        }
        if (leader) commit(&bar);
    }
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (tid == 0) { clk[2 * blockIdx.x] = t1 - t0; clk[2 * blockIdx.x + 1] = count; }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

int main() {
    std::vector<uint8_t> A(M * K), B(N * K);
    srand(1);
    for (auto& x : A) x = rand() & 255;
    for (auto& x : B) x = rand() & 255;
    uint8_t *dA, *dB;
    int32_t* dD;
    long long* dclk;
    cudaMalloc(&dA, A.size());
    cudaMalloc(&dB, B.size());
    cudaMalloc(&dD, M * N * 4);
    cudaMalloc(&dclk, 8 * 256);
    cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
    std::vector<int32_t> D(M * N);
    for (int mode = 0; mode < 2; mode++)
        for (int as = 0; as < 2; as++)
            for (int bs = 0; bs < 2; bs++) {
                cudaMemset(dD, 0xff, M * N * 4);
                probe<<<1, 128>>>(dA, dB, dD, mode, as, bs, 0, 48, dclk);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d a_s %d b_s %d: CUDA error %s\n", mode, as, bs, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
                int bad = 0, first = -1;
                for (int m = 0; m < M; m++)
                    for (int n = 0; n < N; n++) {
                        long long ref = 0;
                        for (int k = 0; k < K; k++) {
                            int a = as ? (int)(int8_t)A[m * K + k] : (int)A[m * K + k];
                            int b = bs ? (int)(int8_t)B[n * K + k] : (int)B[n * K + k];
                            ref += a * b;
                        }
                        if ((int32_t)ref != D[m * N + n]) { if (first < 0) first = m * N + n; bad++; }
                    }
                printf("A from %s, A %s, B %s: %s (%d of %d wrong%s)\n", mode ? "TMEM" : "smem", as ? "s8" : "u8", bs ? "s8" : "u8",
                       bad ? "FAIL" : "PASS", bad, M * N, bad ? "" : "");
                if (bad) printf("   first mismatch at m=%d n=%d got %d\n", first / N, first % N, D[first]);
            }
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int mode = 0; mode < 2; mode++)
        for (int nacc : {1, 2})
            for (int pn : {16, 32, 48, 64, 80, 96, 112, 128, 192, 256}) {
                if (nacc * pn > 256) continue;          // accumulators live in columns [0, 256), the TMEM A operand above
                const int iters = 4000;
                probe<<<sms, 128>>>(dA, dB, dD, mode, 0, 0, iters, pn, dclk, nacc);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("perf: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
                std::vector<long long> c(sms);
                cudaMemcpy(c.data(), dclk, sms * 8, cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (auto x : c) mx = x > mx ? x : mx;
                double per = (double)mx / iters;
                printf("A from %s, %d accumulators, N=%3d: %.1f clk per MMA (128 x %d x 32) = %.0f MAC/clk/SM\n", mode ? "TMEM" : "smem", nacc, pn, per, pn,
                       128.0 * pn * 32 / per);
            }
    for (int n : {32, 48}) {
        const int smem = 8 * 7 * 2 * (n / 8) * 128;
        cudaFuncSetAttribute(probe_cycle, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int variant = 0; variant < 4; variant++) {
            probe_cycle<<<sms, 128, smem>>>(n, 200, dclk, variant & 1, variant >> 1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("cycle: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
            long long c[2];
            cudaMemcpy(c, dclk, 16, cudaMemcpyDeviceToHost);
            printf("IDW issue pattern, N=%d, commit per pass %d, signed top digit %d: %.1f clk per MMA over %lld MMAs\n", n, variant & 1,
                   variant >> 1, (double)c[0] / c[1], c[1]);
        }
    }
    return 0;
}
