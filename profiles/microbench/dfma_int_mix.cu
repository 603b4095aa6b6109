#include <cstdio>
#include <cuda_runtime.h>
// NI integer ops per DFMA, 4 independent chains each
template <int NI>
__global__ void mix(double* out, int* iout, int iters, long long* cyc) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    int b0 = threadIdx.x, b1 = b0 + 1, b2 = b0 + 2, b3 = b0 + 3;
    const double m = 1.0000001, c = 1e-7;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                b0 = (b0 ^ 0x55aa) + 12345; b1 = (b1 ^ 0x55aa) + 12345; b2 = (b2 ^ 0x55aa) + 12345; b3 = (b3 ^ 0x55aa) + 12345;
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = b0 + b1 + b2 + b3;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NI> void run(int warps) {
    double* out; int* iout; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&iout, 1 << 24); cudaMalloc(&cyc, 8);
    int iters = 1000;
    mix<NI><<<148, 32 * warps>>>(out, iout, iters, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double dfma = iters * 8.0 * 4 * (warps / 4.0);           // per SMSP
    double total = dfma * (1 + 2 * NI);                       // xor+add = 2 int instr (maybe fused to 1 LOP3+IADD)
    printf("NI %d warps/SMSP %.0f: cycles %lld  DFMA/cycle/SMSP %.3f  (dfma+2*NI int)/cycle %.3f\n", NI, warps / 4.0, h, dfma / h, total / h);
}
int main() { run<0>(16); run<1>(16); run<2>(16); run<3>(16); run<4>(16); run<1>(4); run<2>(4); return 0; }
