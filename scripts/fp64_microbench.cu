// FP64 pipe characterisation on the GPU box (development aid):
//   cycles per instruction for C independent DFMA chains per thread with W warps per SM sub-partition.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/fp64_microbench scripts/fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void chains(double* out, long long* cyc, int iters, double a, double b) {
    double x[C];
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = threadIdx.x + c;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int c = 0; c < C; ++c) x[c] = fma(x[c], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += x[c];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int C>
void run(int warps_per_block) {
    double* out; long long* cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    chains<C><<<148, 32 * warps_per_block>>>(out, cyc, iters, 0.999999, 1e-6);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / ((double)iters * 16 * C);
    printf("chains=%d warps/SM=%d (%.1f per scheduler): %.2f cycles per DFMA per warp -> %.2f DFMA warp-instr/cycle/scheduler\n", C, warps_per_block,
           warps_per_block / 4.0, per, (warps_per_block / 4.0) / per);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16}) {
        run<1>(w); run<2>(w); run<4>(w); run<8>(w);
    }
    return 0;
}
