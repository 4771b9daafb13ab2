// Exploratory (gated): can the Hamming cost matrix popc(c1[a] ^ c2[b]) be built on the tensor cores?
// mma.sync m16n8k128 .b1 .xor.popc IS a Hamming-distance tile: D[m][n] += popc(A[m][0..127] ^ B[n][0..127]); a 64-bit census code
// fills half of K. This measures the rate of that instruction on B200 (legacy mma.sync path; tcgen05 has no .b1 kind) against
// the 1.72 G Hamming evaluations of a frame (0.74 ms on the POPC pipe today).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o b1_mma b1_mma.cu && ./b1_mma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(int iters, int *out, unsigned seed)
{
    unsigned a0 = threadIdx.x * 2654435761u + seed, a1 = a0 ^ 0x9E3779B9u, b0 = a0 * 31u + 7u;
    int c[4][4] = {};
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) // four independent accumulator tiles per warp
            asm volatile("mma.sync.aligned.m16n8k128.row.col.s32.b1.b1.s32.xor.popc {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                         : "+r"(c[u][0]), "+r"(c[u][1]), "+r"(c[u][2]), "+r"(c[u][3]) : "r"(a0 + u), "r"(a1), "r"(b0 + i));
    }
    int s = 0;
    for (int u = 0; u < 4; u++) for (int j = 0; j < 4; j++) s += c[u][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the int8 route: popc(a ^ b) = (64 - <a', b'>) / 2 with a', b' the codes expanded to +-1 bytes; K = 64 -> two m16n8k32 per tile
__global__ void k8(int iters, int *out, unsigned seed)
{
    unsigned a[4], b[2];
    for (int j = 0; j < 4; j++) a[j] = (threadIdx.x + j) * 2654435761u + seed;
    b[0] = a[0] * 31u + 7u; b[1] = a[1] * 17u + 3u;
    int c[4][4] = {};
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+r"(c[u][0]), "+r"(c[u][1]), "+r"(c[u][2]), "+r"(c[u][3]) : "r"(a[0] + u), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0] + i), "r"(b[1]));
    }
    int s = 0;
    for (int u = 0; u < 4; u++) for (int j = 0; j < 4; j++) s += c[u][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    int *d;
    const int blocks = 148 * 8, threads = 256, iters = 20000;
    cudaMalloc(&d, blocks * threads * 4);
    k<<<blocks, threads>>>(100, d, 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<blocks, threads>>>(iters, d, 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)blocks * (threads / 32) * iters * 4;
    const double pairs = mmas * 16 * 8; // Hamming evaluations of 64-bit codes if K is half filled: one per (m, n)
    printf("%s: %.3f ms, %.2f G mma.m16n8k128.b1/s, %.1f G code pairs/s -> 1.72 G evals of a frame in %.3f ms (POPC pipe today: 0.74 ms floor)\n",
           cudaGetErrorString(cudaGetLastError()), ms, mmas / ms / 1e6, pairs / ms / 1e6, 1.72e9 / (pairs / ms));
    k8<<<blocks, threads>>>(100, d, 1);
    cudaEventRecord(e0);
    k8<<<blocks, threads>>>(iters, d, 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double pairs8 = mmas * 16 * 8 / 2; // two k32 steps per 64-element code pair
    printf("%s: %.3f ms, %.2f G mma.m16n8k32.s8/s (%.1f dense int8 TOP/s), %.1f G code pairs/s -> 1.72 G evals in %.3f ms before the band / epilogue costs\n",
           cudaGetErrorString(cudaGetLastError()), ms, mmas / ms / 1e6, mmas * 16 * 8 * 32 * 2 / ms / 1e9, pairs8 / ms / 1e6, 1.72e9 / (pairs8 / ms));
    return 0;
}
