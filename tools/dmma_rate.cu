// dmma_rate.cu -- raw throughput of mma.sync.m8n8k4.f64 (DMMA) against DFMA on sm_100a, per occupancy.
// Decides whether the one-pass moments (a weighted SYRK over the particles) can go through the FP64 tensor path.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate dmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void __launch_bounds__(128) k_dmma(double* out, int reps)
{
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + 1e-6 * threadIdx.x, b = 1.0 - 1e-6 * threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < reps; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[(size_t)blockIdx.x * 128 + threadIdx.x] = s;
}
template <int NACC>
__global__ void __launch_bounds__(128) k_dfma(double* out, int reps)
{
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = 1e-3 * i;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
#pragma unroll 1
    for (int it = 0; it < reps; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[(size_t)blockIdx.x * 128 + threadIdx.x] = s;
}
template <class K>
double timeit(K k, double* out, int blocks, int reps, double flops_per_thread_rep)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<blocks, 128>>>(out, 8);
    float best = 1e30f;
    for (int t = 0; t < 5; ++t) {
        cudaEventRecord(a);
        k<<<blocks, 128>>>(out, reps);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return flops_per_thread_rep * blocks * 128.0 * reps / (best * 1e-3) / 1e12;
}
int main()
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 16 * 128);
    const int reps = 20000;
    for (int bps : {1, 2, 4, 8}) {
        const int blocks = sms * bps;
        // one DMMA = 8x8x4 FMA per warp = 8 FMA per thread = 16 flops per thread
        printf("%2d warps/SM: DMMA x6 acc %.2f TFLOP/s, DMMA x12 acc %.2f, DFMA x8 chains %.2f, DFMA x16 chains %.2f\n", bps * 4,
               timeit(k_dmma<6>, out, blocks, reps, 6 * 16.0), timeit(k_dmma<12>, out, blocks, reps, 12 * 16.0),
               timeit(k_dfma<8>, out, blocks, reps, 8 * 2.0), timeit(k_dfma<16>, out, blocks, reps, 16 * 2.0));
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
