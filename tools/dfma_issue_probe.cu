// Does a warp-wide DFMA hold the SM sub-partition's issue port for both of its two fp64-pipe cycles, or can integer / shared-memory
// instructions issue under it?  Per thread 8 independent DFMA chains; per DFMA, K independent integer instructions (IMAD) or shared
// loads.  One CTA of W warps per SM, 148 CTAs.  Prints clocks per loop iteration per SM sub-partition-resident warp set.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dfma_issue_probe tools/dfma_issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int K, bool LDS>
__global__ void probe(double* out, int* iout, int iters, double a, int m) {
    __shared__ int sh[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh[i] = i;
    __syncthreads();
    const volatile int* shp = sh + (threadIdx.x & 31);
    double x[8];
    int q[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3 + i, q[i] = threadIdx.x + i;
    int ls = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            x[i] = fma(x[i], a, 1e-9);
            if (LDS) {
#pragma unroll
                for (int k = 0; k < K; k++) ls ^= shp[(i * K + k) * 32];
            } else {
#pragma unroll
                for (int k = 0; k < K; k++) q[(i + k) & 7] = q[(i + k) & 7] * m + it;
            }
        }
    }
    double s = 0.0;
    int qs = ls;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i], qs += q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = qs;
}

template <int K, bool LDS>
void run(int warps, double* out, int* iout) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<K, LDS><<<148, warps * 32>>>(out, iout, 100, 0.999999, 3);
    cudaEventRecord(e0);
    probe<K, LDS><<<148, warps * 32>>>(out, iout, iters, 0.999999, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cyc = ms * 1e-3 * clk * 1e3;
    // per sub-partition: warps / 4 warps, each 8 DFMA + 8 K others per iteration
    const double wpp = warps / 4.0;
    printf("%s K=%d warps/SM=%2d: %.3f ms  %.2f clk per DFMA-slot per sub-partition (1 DFMA + %d %s)\n", LDS ? "LDS " : "IMAD", K, warps, ms,
           cyc / (iters * 8.0 * wpp), K, LDS ? "LDS" : "IMAD");
}


// DFMAs with three distinct register operands (acc = y * z + acc with y, z, acc all in registers), as in the far-field kernel's
// interpolation sums, against the two-register form above: does operand delivery slow the pipe down?
template <int NCH>
__global__ void probe3(double* out, int iters, double a) {
    double x[NCH], y[NCH], z[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) x[i] = threadIdx.x * 1e-3 + i, y[i] = 1.0 - 1e-7 * (threadIdx.x + i), z[i] = a + 1e-9 * i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) x[i] = fma(y[i], z[(i + 1) % NCH], x[i]);
#pragma unroll
        for (int i = 0; i < NCH; i++) y[i] = fma(x[i], z[i], y[(i + 3) % NCH]);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NCH>
void run3(int warps, double* out) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe3<NCH><<<148, warps * 32>>>(out, 100, 1e-9);
    cudaEventRecord(e0);
    probe3<NCH><<<148, warps * 32>>>(out, iters, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cyc = ms * 1e-3 * clk * 1e3;
    printf("3-register DFMA, %2d chains, warps/SM=%2d: %.3f ms  %.2f clk per DFMA per sub-partition\n", NCH, warps, ms,
           cyc / (iters * 2.0 * NCH * (warps / 4.0)));
}

int main() {
    double* out;
    int* iout;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&iout, 148 * 1024 * sizeof(int));
    for (int warps : {4, 16, 32}) {
        run<0, false>(warps, out, iout);
        run<1, false>(warps, out, iout);
        run<2, false>(warps, out, iout);
        run<3, false>(warps, out, iout);
        run<1, true>(warps, out, iout);
        run<2, true>(warps, out, iout);
    }
    for (int warps : {4, 16, 32}) {
        run3<8>(warps, out);
        run3<16>(warps, out);
    }
    return 0;
}
