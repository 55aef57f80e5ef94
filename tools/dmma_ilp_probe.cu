// dmma_ilp_probe.cu -- how many independent DMMA chains does one SM sub-partition need?
// For W warps per scheduler and A independent accumulators per warp, measure DMMA throughput as % of peak.
#include <cstdio>
#include <cuda_runtime.h>
template <int A>
__global__ void k(double* out, int iters) {
    double c[A][2];
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
#pragma unroll
    for (int i = 0; i < A; i++) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < A; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < A; i++) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}
template <int A> void run(int warps_per_sched, double* d, int sms) {
    int threads = warps_per_sched * 4 * 32, iters = 40000 / A;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<A><<<sms, threads>>>(d, iters); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<A><<<sms, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 256 * A * (double)iters * sms * (threads / 32);
    printf("warps/scheduler %d  accumulators/warp %d  chains/scheduler %2d : %6.2f TFLOP/s\n", warps_per_sched, A, warps_per_sched * A, flops / ms / 1e9);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* d; cudaMalloc(&d, 64);
    for (int w : {1, 2, 4}) { run<1>(w, d, p.multiProcessorCount); run<2>(w, d, p.multiProcessorCount); run<4>(w, d, p.multiProcessorCount); run<8>(w, d, p.multiProcessorCount); run<16>(w, d, p.multiProcessorCount); }
    return 0;
}
