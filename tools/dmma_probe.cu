// dmma_probe.cu -- which arithmetic does mma.sync.m8n8k4.f64 perform on sm_100a?
//
// Round-2 groundwork ("next levers"): the mutation kernel is instruction-issue bound and its two
// triangular mat-vecs could run as FP64 tensor-core tiles over the 32 particles of a warp -- but only if the CPU
// oracle can reproduce the hardware's accumulation order bit for bit.  This probe compares the instruction's result
// with candidate orders evaluated with explicit fma on the host:
//   seq     d = fma(a3,b3, fma(a2,b2, fma(a1,b1, fma(a0,b0, c))))      (k ascending, one rounding per step)
//   rev     the same with k descending
//   pair    (fma(a1,b1, a0*b0) + fma(a3,b3, a2*b2)) + c
//   exact   round(c + sum a_k b_k) with a single rounding (long double / __float128 accumulation)
// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/dmma_probe tools/dmma_probe.cu ; run on the GPU box.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

__global__ void k_dmma(const double* __restrict__ A, const double* __restrict__ B, const double* __restrict__ C, double* __restrict__ D, int ntile)
{
    const int lane = threadIdx.x & 31;
    for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
        const double* a = A + (size_t)t * 32;     // A[8][4] row-major
        const double* b = B + (size_t)t * 32;     // B[4][8] row-major
        const double* c = C + (size_t)t * 64;     // C[8][8] row-major
        const int row = lane >> 2, q = lane & 3;
        const double av = a[row * 4 + q];          // A fragment: (row = lane/4, k = lane%4)
        const double bv = b[q * 8 + row];          // B fragment: (k = lane%4, col = lane/4)
        double c0 = c[row * 8 + 2 * q], c1 = c[row * 8 + 2 * q + 1];
        double d0, d1;
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                     : "=d"(d0), "=d"(d1) : "d"(av), "d"(bv), "d"(c0), "d"(c1));
        D[(size_t)t * 64 + row * 8 + 2 * q] = d0;
        D[(size_t)t * 64 + row * 8 + 2 * q + 1] = d1;
    }
}

static uint64_t rs = 0x9E3779B97F4A7C15ull;
static double rnd()
{
    rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17;
    const double u = (double)(rs >> 11) * 0x1p-53 - 0.5;
    rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17;
    const int e = (int)(rs % 41) - 20;             // wide dynamic range: cancellations expose the order
    return ldexp(u, e);
}

int main()
{
    const int ntile = 4096;
    double *hA = (double*)malloc(sizeof(double) * 32 * ntile), *hB = (double*)malloc(sizeof(double) * 32 * ntile);
    double *hC = (double*)malloc(sizeof(double) * 64 * ntile), *hD = (double*)malloc(sizeof(double) * 64 * ntile);
    for (int i = 0; i < 32 * ntile; ++i) { hA[i] = rnd(); hB[i] = rnd(); }
    for (int i = 0; i < 64 * ntile; ++i) hC[i] = rnd();
    double *A, *B, *C, *D;
    cudaMalloc(&A, sizeof(double) * 32 * ntile); cudaMalloc(&B, sizeof(double) * 32 * ntile);
    cudaMalloc(&C, sizeof(double) * 64 * ntile); cudaMalloc(&D, sizeof(double) * 64 * ntile);
    cudaMemcpy(A, hA, sizeof(double) * 32 * ntile, cudaMemcpyHostToDevice);
    cudaMemcpy(B, hB, sizeof(double) * 32 * ntile, cudaMemcpyHostToDevice);
    cudaMemcpy(C, hC, sizeof(double) * 64 * ntile, cudaMemcpyHostToDevice);
    k_dmma<<<64, 32>>>(A, B, C, D, ntile);
    if (cudaMemcpy(hD, D, sizeof(double) * 64 * ntile, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("cuda error\n"); return 1; }
    long n = 0, m_seq = 0, m_rev = 0, m_pair = 0, m_exact = 0, distinct = 0;
    for (int t = 0; t < ntile; ++t)
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) {
                const double* a = hA + (size_t)t * 32 + i * 4;
                const double* b = hB + (size_t)t * 32;
                const double c = hC[(size_t)t * 64 + i * 8 + j];
                const double bk[4] = {b[0 * 8 + j], b[1 * 8 + j], b[2 * 8 + j], b[3 * 8 + j]};
                const double seq = fma(a[3], bk[3], fma(a[2], bk[2], fma(a[1], bk[1], fma(a[0], bk[0], c))));
                const double rev = fma(a[0], bk[0], fma(a[1], bk[1], fma(a[2], bk[2], fma(a[3], bk[3], c))));
                const double pair = (fma(a[1], bk[1], a[0] * bk[0]) + fma(a[3], bk[3], a[2] * bk[2])) + c;
                __float128 ex = (__float128)c;
                for (int k = 0; k < 4; ++k) ex += (__float128)a[k] * (__float128)bk[k];
                const double exact = (double)ex;
                const double d = hD[(size_t)t * 64 + i * 8 + j];
                ++n;
                if (seq != rev || seq != pair || seq != exact) ++distinct;
                m_seq += (memcmp(&d, &seq, 8) == 0); m_rev += (memcmp(&d, &rev, 8) == 0);
                m_pair += (memcmp(&d, &pair, 8) == 0); m_exact += (memcmp(&d, &exact, 8) == 0);
            }
    printf("mma.sync.m8n8k4.f64 on this GPU: %ld outputs (%ld where the candidate orders differ)\n", n, distinct);
    printf("  bit-identical to  seq (k ascending fma chain): %ld\n", m_seq);
    printf("  bit-identical to  rev (k descending fma chain): %ld\n", m_rev);
    printf("  bit-identical to  pair ((p0+p1)+(p2+p3))+c    : %ld\n", m_pair);
    printf("  bit-identical to  exact (single rounding)     : %ld\n", m_exact);
    return 0;
}
