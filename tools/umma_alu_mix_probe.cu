// umma_alu_mix_probe.cu -- do ordinary ALU instructions and tcgen05.mma kind::i8 slow each other down on sm_100a?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/umma_alu_mix_probe tools/umma_alu_mix_probe.cu
//
// One CTA per SM: warp 0 issues back-to-back 128 x 48 x 32 int8 MMAs (A in tensor memory, B in shared memory,
// the shape of the IDW kernel), warps 1..12 run a loop of one instruction type (8 independent chains per thread).
// Three launches per type: MMAs alone, ALU loop alone, both together; each role reports its own clock64 span.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}" ::"r"(d),
                 "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}

constexpr int N = 48, LBO = (N / 8) * 128, SBO = 128;
enum { OP_DFMA, OP_DADD, OP_DMUL, OP_DSETP, OP_FFMA, OP_IMAD, OP_IMADWIDE, OP_IADD3, OP_LDS, OP_LDS32, OP_STS, OP_SHFL, OP_MUFU, OP_I2F64, OP_F2F, OP_I2F32, OP_STG, OP_LDG, OP_COUNT };
static const char* op_name[] = {"DFMA", "DADD", "DMUL", "DSETP+FSEL", "FFMA", "IMAD", "IMAD.WIDE", "IADD3", "LDS.64", "LDS.32", "STS.64", "SHFL", "MUFU.RCP", "I2F.F64.S64", "F2F.F64.F32", "I2F.F32.S32", "STG.64", "LDG.64"};

template <int op>
__global__ void __launch_bounds__(13 * 32) mix(int n_mma, int n_alu, long long* clk, double* sink, double* gbuf) {
    __shared__ __align__(128) uint8_t sB[2 * 2 * LBO];
    __shared__ __align__(16) double sD[13 * 32];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    for (int i = tid; i < (int)sizeof(sB); i += blockDim.x) sB[i] = (uint8_t)(i * 7);
    sD[tid] = 1.0 + tid * 1e-9;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tbase_s;
    const long long t0 = clock64();
    if (warp == 0) {
        if (n_mma > 0) {
            uint32_t leader;
            asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(leader));
            const uint32_t idesc = make_idesc(128, N);
            const uint64_t bd0 = make_desc(smem_u32(sB), LBO, SBO), bd1 = make_desc(smem_u32(sB) + 2 * LBO, LBO, SBO);
            const uint32_t a0 = tbase + 256, a1 = tbase + 264;
            for (int it = 0; it < n_mma; it += 4) {
                if (leader) {
                    mma_ts(tbase, a0, bd0, idesc, 1);
                    mma_ts(tbase + N, a1, bd1, idesc, 1);
                    mma_ts(tbase, a1, bd0, idesc, 1);
                    mma_ts(tbase + N, a0, bd1, idesc, 1);
                }
            }
            if (leader) commit(&bar);
            __syncwarp();
            mbar_wait(&bar, 0);
            if (tid == 0) clk[2 * blockIdx.x] = clock64() - t0;
        }
    } else if (n_alu > 0) {
        double d[8];
        float f[8];
        int x[8];
        long long w[8];
        for (int i = 0; i < 8; i++) {
            d[i] = sD[(tid + i) % (13 * 32)];
            f[i] = (float)d[i];
            x[i] = tid + i;
            w[i] = tid * 3 + i;
        }
        const double c1 = sD[1], c2 = sD[2];
        const float g1 = (float)c1, g2 = (float)c2;
        for (int it = 0; it < n_alu; it += 8) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                switch (op) {
                    case OP_DFMA: d[i] = fma(d[i], c1, c2); break;
                    case OP_DADD: d[i] = d[i] + c1; break;
                    case OP_DMUL: d[i] = d[i] * c1; break;
                    case OP_DSETP: d[i] = d[i] <= c2 ? c1 : d[i]; break;
                    case OP_FFMA: f[i] = fmaf(f[i], g1, g2); break;
                    case OP_IMAD: x[i] = x[i] * 3 + it; break;
                    case OP_IMADWIDE: w[i] += (long long)x[i] * 256LL; break;
                    case OP_IADD3: x[i] = x[i] + it + i; break;
                    case OP_LDS: d[i] = sD[(__double2loint(d[i]) + x[i] + it) & 255]; break;
                    case OP_LDS32: x[i] = reinterpret_cast<const int*>(sD)[(x[i] + it) & 511]; break;
                    case OP_STS: sD[(tid + 32 * i) % (13 * 32)] = __longlong_as_double(w[i] + it); break;
                    case OP_SHFL: x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1 + (i & 3)); break;
                    case OP_MUFU: f[i] = __fdividef(1.0f, f[i]) ; asm volatile("" : "+f"(f[i])); break;
                    case OP_I2F64: d[i] = (double)(w[i] + it); w[i] ^= __double2loint(d[i]); break;
                    case OP_F2F: d[i] = (double)f[i]; f[i] += __double2loint(d[i]) ; break;
                    case OP_I2F32: f[i] = (float)(x[i] + it); x[i] ^= __float_as_int(f[i]); break;
                    case OP_STG: gbuf[(size_t)(tid + blockDim.x * (i + 8 * ((it >> 3) & 63))) + (size_t)blockIdx.x * 13 * 32 * 512] = __longlong_as_double(w[i] + it); break;
                    case OP_LDG: d[i] = __ldg(gbuf + (size_t)((__double2loint(d[i]) & 0) + tid + blockDim.x * (i + 8 * ((it >> 3) & 63))) + (size_t)blockIdx.x * 13 * 32 * 512); break;
                }
            }
        }
        double acc = 0;
        for (int i = 0; i < 8; i++) acc += d[i] + f[i] + x[i] + (double)w[i];
        if (acc == 1.2345e-300) sink[tid] = acc;
        if (tid == 32) clk[2 * blockIdx.x + 1] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* dclk;
    double* dsink;
    cudaMalloc(&dclk, 2 * sms * 8);
    cudaMalloc(&dsink, 13 * 32 * 8);
    double* dgbuf;
    cudaMalloc(&dgbuf, (size_t)sms * 13 * 32 * 512 * 8);
    cudaMemset(dgbuf, 0, (size_t)sms * 13 * 32 * 512 * 8);
    const int n_mma = 8000;
    auto run = [&](int nm, int na, int op, double* mma_clk, double* alu_clk) {
        cudaMemset(dclk, 0, 2 * sms * 8);
        switch (op) {
#define CASE(o) case o: mix<o><<<sms, 13 * 32>>>(nm, na, dclk, dsink, dgbuf); break;
            CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16) CASE(17)
#undef CASE
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
        long long c[2];
        cudaMemcpy(c, dclk, 16, cudaMemcpyDeviceToHost);
        *mma_clk = (double)c[0];
        *alu_clk = (double)c[1];
    };
    double m0, a0, m1, a1, dummy;
    run(n_mma, 0, 0, &m0, &dummy);
    run(n_mma, 0, 0, &m0, &dummy);
    printf("%d MMAs (128 x %d x 32, A in TMEM) alone: %.0f clk = %.1f clk per MMA\n", n_mma, N, m0, m0 / n_mma);
    for (int op = 0; op < OP_COUNT; op++) {
        // size the ALU loop so that alone it takes about as long as the MMAs
        int na = 4096;
        run(0, na, op, &dummy, &a0);
        na = (int)(na * (m0 / a0)) / 8 * 8;
        run(0, na, op, &dummy, &a0);
        run(n_mma, na, op, &m1, &a1);
        printf("%-10s x %7d per thread, 12 warps: alone %.0f clk (%.2f clk per warp-instr and scheduler) | together: MMAs %.0f clk (x%.2f), ALU %.0f clk (x%.2f)\n",
               op_name[op], na, a0, a0 / (na * 3.0), m1, m1 / m0, a1, a1 / a0);
    }
    return 0;
}
