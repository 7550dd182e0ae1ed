// Micro-benchmark: issue rate of the FP64 instruction kinds the solver kernels are made of (B200, sm_100a).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/exp/fp64_ops_probe.bin tools/exp/fp64_ops_probe.cu
// 8 independent chains per thread, 32 warps per SM; rate from the difference of two run lengths (CUDA events).
#include <cstdio>
#include <cuda_runtime.h>

enum { FMA2 = 0, FMA3, MUL2, MUL1C, ADD2, ADD1C, SETP, RCP, FMA_RCP_MIX, MINMAX, F2F };

template <int OP>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b) {
    double x[8], y[8];
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = 1.0 + threadIdx.x * 1e-6 + i * 1e-3; y[i] = 1.0 + 1e-9 * (threadIdx.x + i + 1); }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == FMA2) x[i] = fma(y[i], a, x[i]);
                if (OP == FMA3) x[i] = fma(y[i], y[(i + 3) & 7], x[i]);
                if (OP == MUL2) x[i] = x[i] * y[i];
                if (OP == MUL1C) x[i] = x[i] * a;
                if (OP == ADD2) x[i] = x[i] + y[i];
                if (OP == ADD1C) x[i] = x[i] + b;
                if (OP == SETP) cnt += (x[i] + 0.0 * cnt < y[(i + u) & 7]) ? 1 : 0;
                if (OP == RCP) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[i])); x[i] = r; }
                if (OP == FMA_RCP_MIX) { if (u == 0) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[i])); x[i] = r; } else x[i] = fma(y[i], a, x[i]); }
                if (OP == MINMAX) x[i] = fmax(x[i], y[i]) ;
                if (OP == F2F) x[i] = (double)((float)x[i]) ;
            }
        }
    }
    double s = cnt;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int sms) {
    const int grid = sms * 4, iters = 2000;
    double* out; cudaMalloc(&out, sizeof(double) * grid * 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<grid, 256>>>(out, 50, 1.0000001, 1e-9);
    float m1, m2;
    cudaEventRecord(e0); k<OP><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&m1, e0, e1);
    cudaEventRecord(e0); k<OP><<<grid, 256>>>(out, 3 * iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&m2, e0, e1);
    const double instr = 2.0 * iters * 64 * 8 * 4;     // warp instructions per SM in the difference
    const double cycles = (m2 - m1) * 1e-3 * 1.965e9;
    printf("%-12s %.3f warp instructions / clock / SM   (%.2f cycles per instruction per sub-partition)\n", name, instr / cycles,
           4.0 * cycles / instr);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, cycles at 1.965 GHz\n", p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    run<FMA2>("DFMA r,c,r", sms); run<FMA3>("DFMA r,r,r", sms); run<MUL2>("DMUL r,r", sms); run<MUL1C>("DMUL r,c", sms);
    run<ADD2>("DADD r,r", sms); run<ADD1C>("DADD r,c", sms); run<SETP>("DSETP+", sms); run<RCP>("MUFU.RCP64H", sms);
    run<FMA_RCP_MIX>("7 DFMA+1 RCP", sms); run<MINMAX>("DMNMX?", sms); run<F2F>("F2F pair", sms);
    return 0;
}
