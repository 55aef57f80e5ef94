// fp64_mix_probe.cu -- do scalar fp64 instructions (DADD/DFMA/DSETP) steal DMMA time on B200?
// One CTA per SM, 8 warps: warps 0-3 (one per scheduler) stream DMMAs, warps 4-7 stream scalar fp64 ops.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, int n_mma, int n_dadd, int mode) {
    const int warp = threadIdx.x >> 5;
    double s = 0;
    if (warp < 4) {
        double c[4][2]; double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
#pragma unroll
        for (int i = 0; i < 4; i++) { c[i][0] = i; c[i][1] = -i; }
        for (int it = 0; it < n_mma; it++) {
#pragma unroll
            for (int i = 0; i < 4; i++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
#pragma unroll
        for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1];
    } else {
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
        if (mode == 0) {
            for (int it = 0; it < n_dadd; it++) {
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = __dadd_rn(x[i], 1.25);
            }
        } else if (mode == 1) {
            for (int it = 0; it < n_dadd; it++) {
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = fma(x[i], 1.0000001, 1e-9);
            }
        } else {   // fp32 FMA for comparison
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; i++) y[i] = threadIdx.x + i;
            for (int it = 0; it < n_dadd; it++) {
#pragma unroll
                for (int i = 0; i < 8; i++) y[i] = fmaf(y[i], 1.0000001f, 1e-9f);
            }
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = y[i];
        }
#pragma unroll
        for (int i = 0; i < 8; i++) s += x[i];
    }
    if (s == 12345.678) out[0] = s;
}
float run(double* d, int sms, int nm, int nd, int mode) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<sms, 256>>>(d, nm, nd, mode); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<sms, 256>>>(d, nm, nd, mode); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* d; cudaMalloc(&d, 64);
    const int NM = 20000;   // 80000 DMMAs per MMA warp = 1.28e6 clks at 16 clk each
    float tm = run(d, p.multiProcessorCount, NM, 0, 0);
    printf("DMMA only: %.3f ms (%.1f clk per DMMA at 1.965 GHz)\n", tm, tm * 1e-3 * 1.965e9 / (NM * 4));
    for (int mode = 0; mode < 3; mode++) {
        const char* nm[] = {"DADD", "DFMA", "FFMA"};
        for (int nd : {2500, 5000, 10000, 20000, 40000}) {
            float td = run(d, p.multiProcessorCount, 0, nd, mode);
            float tb = run(d, p.multiProcessorCount, NM, nd, mode);
            printf("%s x %6d (8 per iter per warp): alone %.3f ms, with DMMA %.3f ms, DMMA-only %.3f ms -> extra %.3f ms = %.2f clk per scalar warp-instr\n",
                   nm[mode], nd * 8, td, tb, tm, tb - tm, (tb - tm) * 1e-3 * 1.965e9 / (nd * 8.0));
        }
    }
    return 0;
}
