// peak_probe.cu -- measures the roofline denominators MEASURED_PEAKS.json does not hold:
// fp64 FMA rate, fp64 DMMA rate (mma.sync m8n8k4 / m16n8k8 / m16n8k16), write-only HBM bandwidth.
// Build+run on the B200 box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/peak_probe.cu -o /tmp/pp && /tmp/pp
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void dmma884_kernel(double* out, int iters) {
    double c[8][2];
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

__global__ void dmma1688_kernel(double* out, int iters) {
    double c[4][4];
    double a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = threadIdx.x * 1e-3 + i;
    b[0] = 1.0 + threadIdx.x * 1e-6; b[1] = 0.5;
#pragma unroll
    for (int i = 0; i < 4; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 2 * i; c[i][3] = 1; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

__global__ void dmma16816_kernel(double* out, int iters) {
    double c[4][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = 1.0 + threadIdx.x * 1e-6 + i;
#pragma unroll
    for (int i = 0; i < 4; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 2 * i; c[i][3] = 1; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

__global__ void write_kernel(double2* p, size_t n2, double v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    double2 val = make_double2(v, v + 1);
    for (; i < n2; i += stride) p[i] = val;
}
__global__ void write_cs_kernel(double2* p, size_t n2, double v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    double2 val = make_double2(v, v + 1);
    for (; i < n2; i += stride) __stcs(&p[i], val);
}
__global__ void copy_kernel(const double2* __restrict__ a, double2* __restrict__ b, size_t n2) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) b[i] = a[i];
}

template <typename F> float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* d; CK(cudaMalloc(&d, 1024));
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    const int iters = 20000;
    {   // DFMA: blocks = 8/SM, 256 threads
        int blocks = sms * 8, threads = 256;
        float ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9); }, 5);
        double flops = 2.0 * 16 * (double)iters * blocks * threads;
        printf(", \"fp64_dfma_tflops\": %.2f", flops / ms / 1e9);
    }
    {
        int blocks = sms * 8, threads = 256;
        float ms = time_ms([&] { dmma884_kernel<<<blocks, threads>>>(d, iters); }, 5);
        double flops = 2.0 * 8 * 8 * 4 * 8 * (double)iters * blocks * (threads / 32);
        printf(", \"fp64_dmma_m8n8k4_tflops\": %.2f", flops / ms / 1e9);
    }
    {
        int blocks = sms * 8, threads = 256;
        float ms = time_ms([&] { dmma1688_kernel<<<blocks, threads>>>(d, iters); }, 5);
        double flops = 2.0 * 16 * 8 * 8 * 4 * (double)iters * blocks * (threads / 32);
        printf(", \"fp64_dmma_m16n8k8_tflops\": %.2f", flops / ms / 1e9);
    }
    {
        int blocks = sms * 8, threads = 256;
        float ms = time_ms([&] { dmma16816_kernel<<<blocks, threads>>>(d, iters); }, 5);
        double flops = 2.0 * 16 * 8 * 16 * 4 * (double)iters * blocks * (threads / 32);
        printf(", \"fp64_dmma_m16n8k16_tflops\": %.2f", flops / ms / 1e9);
    }
    {   // HBM: 8 GiB buffers
        size_t bytes = (size_t)8 << 30, n2 = bytes / 16;
        double2 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
        CK(cudaMemset(a, 0, bytes));
        float ms = time_ms([&] { write_kernel<<<sms * 16, 512>>>(b, n2, 1.0); }, 5);
        printf(", \"hbm_write_gbs\": %.1f", bytes / ms / 1e6);
        ms = time_ms([&] { write_cs_kernel<<<sms * 16, 512>>>(b, n2, 1.0); }, 5);
        printf(", \"hbm_write_stcs_gbs\": %.1f", bytes / ms / 1e6);
        ms = time_ms([&] { CK(cudaMemsetAsync(b, 1, bytes)); }, 5);
        printf(", \"hbm_memset_gbs\": %.1f", bytes / ms / 1e6);
        ms = time_ms([&] { copy_kernel<<<sms * 16, 512>>>(a, b, n2); }, 5);
        printf(", \"hbm_copy_rw_gbs\": %.1f", 2.0 * bytes / ms / 1e6);
        cudaFree(a); cudaFree(b);
    }
    printf("}\n");
    return 0;
}
