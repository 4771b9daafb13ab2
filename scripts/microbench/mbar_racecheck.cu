// Tool check: do compute-sanitizer racecheck / synccheck follow mbarrier arrive (release) -> try_wait (acquire) pairs written
// as inline PTX the way sgm.cu writes them? A chain of warps hands a vector from warp to warp through rings of four entries
// guarded by full / empty mbarriers (32 arrivals each), exactly the protocol of the sweep kernel's exchange rings.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o mbar_racecheck mbar_racecheck.cu
//   compute-sanitizer --tool racecheck ./mbar_racecheck ; compute-sanitizer --tool synccheck ./mbar_racecheck
#include <cstdio>
#include <cstdint>
#ifndef VARIANT
#define VARIANT 0
#endif
__device__ __forceinline__ void mbar_init(unsigned a, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(a) : "memory"); }
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t)); return t; }
__device__ __forceinline__ bool mbar_wait(unsigned bar_s, unsigned parity)
{
#if VARIANT == 0
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar_s), "r"(parity) : "memory");
    if (ok) return true;
    const unsigned long long t0 = global_ns();
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar_s), "r"(parity) : "memory");
        if (ok) return true;
        if (global_ns() - t0 > 4000000000ull) return false;
    }
#else
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}\n" ::"r"(bar_s), "r"(parity) : "memory");
    return true;
#endif
}
constexpr int NW = 6;
__global__ void k(int steps, uint32_t *out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bars = (unsigned)__cvta_generic_to_shared(smem);      // boundary b: full[4], empty[4]
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem + 1024);          // boundary b, entry e: 32 words
    if (warp == 0) {
        for (int i = lane; i < NW * 8; i += 32) mbar_init(bars + 8 * i, 32);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (warp >= NW) return;
    const int pos = NW - 1 - warp; // reversed, as in sgm.cu
    const unsigned in_full = bars + 64 * pos, in_empty = in_full + 32, out_full = bars + 64 * (pos + 1), out_empty = out_full + 32;
    uint32_t acc = 0;
    unsigned lap = 0;
    for (int s = 0; s < steps; s++) {
        const int u = s & 3, e_in = (u + 3) & 3;
        uint32_t v = lane;
        if (pos > 0) {
            if (s == 0) { for (int e = 0; e < 4; e++) mbar_arrive(in_empty + 8 * e); }
            else {
                mbar_wait(in_full + 8 * e_in, u == 0 ? lap ^ 1u : lap);
                __syncwarp();
                v = ring[(pos * 4 + e_in) * 32 + (lane ^ 5)];
                mbar_arrive(in_empty + 8 * e_in);
            }
        }
        acc += v;
        v = v * 3 + s;
        if (pos + 1 < NW) {
            mbar_wait(out_empty + 8 * u, lap);
            ring[((pos + 1) * 4 + u) * 32 + lane] = v;
            mbar_arrive(out_full + 8 * u);
        }
        if (u == 3) lap ^= 1u;
    }
    out[warp * 32 + lane] = acc;
}
int main()
{
    uint32_t *d, h[NW * 32];
    cudaMalloc(&d, sizeof h);
    k<<<1, 256, 1024 + NW * 4 * 32 * 4>>>(1000, d);
    cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("%s last warp lane 0 acc %u\n", cudaGetErrorString(e), h[0]);
    return 0;
}
