// operand_probe.cu -- how fast can a thread-per-particle FP64 mat-vec be fed with its (warp-uniform) matrix operands on
// sm_100a?  The mutation kernel multiplies two 20 x 20 triangular factors per MH step into per-thread vectors; the
// factor entries are the same for every thread.  Three delivery paths are timed at the kernel's occupancy (20 warps/SM):
//   reg   : both operands in vector registers (no delivery cost; the FP64 pipe's own ceiling)
//   ldcu  : __constant__ memory -> uniform registers (LDCU.128, two entries per load) -> DFMA R, R, UR   [what nvcc emits]
//   lds   : shared memory broadcast (LDS.128, two entries per load) -> DFMA R, R, R
// Pattern per "mat-vec": 210 DFMA over 20 accumulators, column by column like the proposal increment s = (cL) z.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o operand_probe operand_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int D = 20, E = D * (D + 1) / 2;
__constant__ double c_L[2 * E];

template <int MODE>
__global__ void __launch_bounds__(128, 5) k_probe(double* out, int reps, const double* gL)
{
    __shared__ __align__(16) double sL[2 * E];
    for (int k = threadIdx.x; k < 2 * E; k += 128) sL[k] = gL[k];
    __syncthreads();
    double z[D], s[D];
#pragma unroll
    for (int k = 0; k < D; ++k) { z[k] = 1.0 + 1e-3 * (threadIdx.x + k); s[k] = 0.0; }
    double rl[E > 0 ? 1 : 1];
    (void)rl;
#pragma unroll 1
    for (int it = 0; it < reps; ++it) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            int e = half * E;
#pragma unroll
            for (int j = 0; j < D; ++j) {
#pragma unroll
                for (int r = j; r < D; ++r) {
                    double m;
                    if (MODE == 0) m = z[(r + j) % D];
                    else if (MODE == 1) m = c_L[e];
                    else m = sL[e];
                    s[r] = fma(m, z[j], s[r]);
                    ++e;
                }
            }
#pragma unroll
            for (int k = 0; k < D; ++k) z[k] = z[k] * 0.999 + 1e-9 * s[k];      // keeps the chain live (40 more FP64 ops)
        }
    }
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) acc += s[k] + z[k];
    out[(size_t)blockIdx.x * 128 + threadIdx.x] = acc;
}

// MODE 3: the same two mat-vecs as straight-line code in a real (non-inlined) function -- outside a loop ptxas rotates the
// uniform registers of the LDCU.128 stream several loads ahead; vectors travel through shared memory
__device__ __noinline__ void matvec_call(double* zs)
{
    double z[D], s[D];
#pragma unroll
    for (int k = 0; k < D; ++k) { z[k] = zs[k * 128]; s[k] = zs[(D + k) * 128]; }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        int e = half * E;
#pragma unroll
        for (int j = 0; j < D; ++j) {
#pragma unroll
            for (int r = j; r < D; ++r) { s[r] = fma(c_L[e], z[j], s[r]); ++e; }
        }
#pragma unroll
        for (int k = 0; k < D; ++k) z[k] = z[k] * 0.999 + 1e-9 * s[k];
    }
#pragma unroll
    for (int k = 0; k < D; ++k) { zs[k * 128] = z[k]; zs[(D + k) * 128] = s[k]; }
}
__global__ void __launch_bounds__(128, 5) k_probe_call(double* out, int reps)
{
    extern __shared__ double sm[];
    double* zs = sm + threadIdx.x;
#pragma unroll
    for (int k = 0; k < D; ++k) { zs[k * 128] = 1.0 + 1e-3 * (threadIdx.x + k); zs[(D + k) * 128] = 0.0; }
#pragma unroll 1
    for (int it = 0; it < reps; ++it) matvec_call(zs);
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 2 * D; ++k) acc += zs[k * 128];
    out[(size_t)blockIdx.x * 128 + threadIdx.x] = acc;
}
double run_call(double* out, int blocks, int reps)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int smem = 2 * D * 128 * sizeof(double);
    cudaFuncSetAttribute(k_probe_call, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_probe_call<<<blocks, 128, smem>>>(out, 4);
    float best = 1e30f;
    for (int t = 0; t < 5; ++t) {
        cudaEventRecord(a);
        k_probe_call<<<blocks, 128, smem>>>(out, reps);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double fma_count = (double)blocks * 128 * reps * 2.0 * (E + 2.0 * D);
    return 2.0 * fma_count / (best * 1e-3) / 1e12;
}

template <int MODE>
double run(double* out, const double* gL, int blocks, int reps)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_probe<MODE><<<blocks, 128>>>(out, 4, gL);
    float best = 1e30f;
    for (int t = 0; t < 5; ++t) {
        cudaEventRecord(a);
        k_probe<MODE><<<blocks, 128>>>(out, reps, gL);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double fma_count = (double)blocks * 128 * reps * 2.0 * (E + 2.0 * D);
    return 2.0 * fma_count / (best * 1e-3) / 1e12;
}

int main()
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double h[2 * E];
    for (int i = 0; i < 2 * E; ++i) h[i] = 1e-3 * (i % 17) - 5e-3;
    cudaMemcpyToSymbol(c_L, h, sizeof(h));
    double *gL, *out;
    cudaMalloc(&gL, sizeof(h)); cudaMemcpy(gL, h, sizeof(h), cudaMemcpyHostToDevice);
    const int blocks = sms * 5;
    cudaMalloc(&out, sizeof(double) * blocks * 128);
    const int reps = 2000;
    printf("SMs %d, %d blocks x 128 threads (20 warps/SM), %d reps of two 210-DFMA triangular mat-vecs\n", sms, blocks, reps);
    printf("reg  operands: %.2f TFLOP/s\n", run<0>(out, gL, blocks, reps));
    printf("ldcu operands: %.2f TFLOP/s\n", run<1>(out, gL, blocks, reps));
    printf("lds  operands: %.2f TFLOP/s\n", run<2>(out, gL, blocks, reps));
    printf("ldcu operands, mat-vecs in a called function (rotating uniform registers): %.2f TFLOP/s\n", run_call(out, blocks, reps));
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
