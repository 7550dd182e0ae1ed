// Micro-benchmark: what the FP64 pipe of one B200 SM sustains under different operand patterns.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/dfma_probe tools/exp/dfma_probe.cu && /tmp/dfma_probe
// Prints warp-level DFMA per clock per SM (nominal: 2 = 64 lanes / 32) for
//   shared   x_i = fma(x_i, a, b)        one register operand per instruction that is not shared
//   half     x_i = fma(y_i, z, x_i)      two
//   distinct x_i = fma(y_i, z_i, x_i)    three distinct 64-bit register operands per instruction
//   mixed    distinct + one integer instruction per DFMA (issue-slot pressure)
#include <cstdio>
#include <cuda_runtime.h>

template <int PATTERN, int CHAINS>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b, long long* cycles) {
    double x[CHAINS], y[CHAINS], z[CHAINS];
    unsigned m = threadIdx.x;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { x[i] = threadIdx.x * 1e-3 + i; y[i] = 1.0 + 1e-9 * (threadIdx.x + i); z[i] = 1e-9 * (i + 1); }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) {
                if (PATTERN == 0) x[i] = fma(x[i], a, b);
                if (PATTERN == 1) x[i] = fma(y[i], a, x[i]);
                if (PATTERN >= 2) x[i] = fma(y[i], z[i], x[i]);
                if (PATTERN == 3) m = m * 1664525u + 1013904223u;
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += x[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + m;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int PATTERN, int CHAINS>
void run(const char* name, int ctas_per_sm, int sms) {
    const int iters = 2000, grid = sms * ctas_per_sm;
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * grid * 256);
    cudaMalloc(&cyc, sizeof(long long) * grid);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<PATTERN, CHAINS><<<grid, 256>>>(out, 10, 1.0000001, 1e-9, cyc);
    cudaEventRecord(e0);
    k<PATTERN, CHAINS><<<grid, 256>>>(out, iters, 1.0000001, 1e-9, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[4096]; cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
    const double warp_instr_per_sm = double(iters) * 8 * CHAINS * 8 * ctas_per_sm;     // 8 warps per CTA
    printf("%-9s chains %d  warps/SM %2d : %.3f DFMA warp-instr/clk/SM (SM clock), %.3f ms\n", name, CHAINS, 8 * ctas_per_sm,
           warp_instr_per_sm / mean, ms);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    for (int c = 1; c <= 4; c *= 2) {
        run<0, 8>("shared", c, sms);
        run<1, 8>("half", c, sms);
        run<2, 8>("distinct", c, sms);
        run<3, 8>("mixed", c, sms);
        run<2, 2>("distinct", c, sms);
        run<2, 1>("distinct", c, sms);
    }
    return 0;
}
