// icache_probe.cu -- how much straight-line code can one SM loop over before instruction fetch limits it? (development tool)
// A kernel body of N independent-ish FFMA instructions (16 B each) is executed in a loop by W warps per SM; reports cycles per
// instruction per warp.  Below the instruction-cache capacity the cost is the issue/dependency latency; above it, the fetch path.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o icache_probe icache_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#define R4(x) x x x x
#define R16(x) R4(R4(x))
#define R64(x) R4(R16(x))
#define R256(x) R4(R64(x))
#define R1024(x) R4(R256(x))
#define BODY "fma.rn.f32 %0, %0, %4, %5; fma.rn.f32 %1, %1, %4, %5; fma.rn.f32 %2, %2, %4, %5; fma.rn.f32 %3, %3, %4, %5;\n"
template <int KB>   // KB of code = KB * 64 instructions
__global__ void probe(float *out, int iters, long long *cyc) {
    float a = threadIdx.x, b = a + 1, c = a + 2, d = a + 3;
    const float m = 1.0000001f, k = 1e-9f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < KB / 4; ++r)      // 256 instructions (4 KB) per block
            asm volatile(R64(BODY) : "+f"(a), "+f"(b), "+f"(c), "+f"(d) : "f"(m), "f"(k));
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
}
template <int KB> void run(int warps, float *out, long long *dcyc) {
    const int iters = 4096 * 16 / KB;
    probe<KB><<<148, 32 * warps>>>(out, 8, dcyc);
    cudaDeviceSynchronize();
    probe<KB><<<148, 32 * warps>>>(out, iters, dcyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost);
    printf("code %4d KB  warps/SM %2d : %.2f cycles per instruction per warp\n", KB, warps, (double)c / ((double)iters * KB * 64));
}
int main() {
    float *out; long long *dcyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&dcyc, 8);
    for (int w : {1, 4, 16}) {
        run<4>(w, out, dcyc); run<8>(w, out, dcyc); run<16>(w, out, dcyc); run<24>(w, out, dcyc); run<32>(w, out, dcyc); run<48>(w, out, dcyc);
        run<64>(w, out, dcyc); run<96>(w, out, dcyc); run<128>(w, out, dcyc); run<192>(w, out, dcyc); run<256>(w, out, dcyc); run<512>(w, out, dcyc);
    }
    return 0;
}
