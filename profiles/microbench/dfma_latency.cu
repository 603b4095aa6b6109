#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double* out, int iters, long long* cyc) {
    double a[ILP];
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-9 + i;
    const double m = 1.0000001, c = 1e-7;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run(int warps_per_sm) {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    int iters = 2000;
    // one block per SM with warps_per_sm warps
    chain<ILP><<<148, 32 * warps_per_sm>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / (iters * 8.0 * ILP);
    printf("ILP %d warps/SM %2d (per SMSP %.2f): %.2f cycles per DFMA per warp; SMSP throughput %.3f DFMA/cycle\n", ILP, warps_per_sm, warps_per_sm / 4.0, per, (warps_per_sm / 4.0) / per);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1>(4); run<2>(4); run<4>(4); run<8>(4);
    run<1>(8); run<1>(16); run<2>(16); run<4>(16); run<1>(32); run<2>(32);
    return 0;
}
