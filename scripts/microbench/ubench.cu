// Instruction-throughput microbenchmarks for the integer ops the matching / SGM kernels lean on.
// Each kernel runs a long dependent-per-thread but ILP-4 chain; grid = 148 * 8 blocks of 256 threads.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
template <int OP> __global__ void k(uint32_t *out, uint32_t seed)
{
    uint32_t a = threadIdx.x * 2654435761u + seed, b = a ^ 0x9E3779B9u, c = a + 0x7F4A7C15u, d = b * 3u + 1u;
    uint32_t x = seed | 1u;
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (OP == 0) { a = __popc(a) + x; b = __popc(b) + x; c = __popc(c) + x; d = __popc(d) + x; }
            if (OP == 1) { a = __vminu2(a, x) + 1; b = __vminu2(b, x) + 1; c = __vminu2(c, x) + 1; d = __vminu2(d, x) + 1; }
            if (OP == 2) { a = __viaddmin_u16x2(a, x, b); b = __viaddmin_u16x2(b, x, c); c = __viaddmin_u16x2(c, x, d); d = __viaddmin_u16x2(d, x, a); }
            if (OP == 3) { a = (a ^ x) + b; b = (b ^ x) + c; c = (c ^ x) + d; d = (d ^ x) + a; }
            if (OP == 4) { a = __shfl_up_sync(0xffffffffu, a, 1); b = __shfl_up_sync(0xffffffffu, b, 1); c = __shfl_up_sync(0xffffffffu, c, 1); d = __shfl_up_sync(0xffffffffu, d, 1); }
            if (OP == 5) { a = __reduce_min_sync(0xffffffffu, a) + threadIdx.x; b = __reduce_min_sync(0xffffffffu, b) + threadIdx.x; c = __reduce_min_sync(0xffffffffu, c) + threadIdx.x; d = __reduce_min_sync(0xffffffffu, d) + threadIdx.x; }
            if (OP == 6) { a = __byte_perm(a, b, 0x5432); b = __byte_perm(b, c, 0x5432); c = __byte_perm(c, d, 0x5432); d = __byte_perm(d, x, 0x5432); }
            if (OP == 7) { a = __vimin3_u16x2(a, b, x); b = __vimin3_u16x2(b, c, x); c = __vimin3_u16x2(c, d, x); d = __vimin3_u16x2(d, a, x); }
            if (OP == 8) { a = __popc(a ^ x) + 1; b = __vminu2(b, x) + 1; c = __popc(c ^ x) + 1; d = __vminu2(d, x) + 1; } // popc + alu mix
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
}
template <int OP> void run(const char *name, int ops_per_iter, uint32_t *out)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256;
    k<OP><<<blocks, threads>>>(out, 12345u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 12345u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * ITERS * 4.0 * ops_per_iter;
    printf("%-28s %8.3f ms  %8.2f Gop/s/lane-total  = %6.2f lane-ops/clk/SM @1.9GHz\n", name, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / 148.0 / 1.9e9);
}
int main()
{
    uint32_t *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s sm_%d%d SMs %d clock %d kHz L2 %d MB\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate, p.l2CacheSize >> 20);
    run<0>("POPC (+IADD)", 1, out);
    run<1>("VIMNMX.U16x2 (+IADD)", 1, out);
    run<2>("VIADDMNMX.U16x2", 1, out);
    run<3>("LOP3+IADD", 1, out);
    run<4>("SHFL.UP", 1, out);
    run<5>("REDUX.MIN (+IADD)", 1, out);
    run<6>("PRMT", 1, out);
    run<7>("VIMNMX3.U16x2", 1, out);
    run<8>("mix popc^ / vmin", 1, out);
    return 0;
}
