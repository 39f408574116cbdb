// pairbench.cu -- microbenchmark of pair-energy inner-loop variants on sm_100a (development tool,
// not part of the library).  Each CTA repeats the receptor x ligand Coulomb sum REPS times with the
// ligand in shared memory, exactly like one MC step of the chain kernel without the proposal.
// Reports pairs/s, pairs/clk/SM (at the SM clock measured inside the kernel) and % of FP32 peak
// at 12 flop/pair.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o pairbench pairbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>
#include <cmath>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float rsq(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

struct Timing { long long clk0, clk1; unsigned long long ns0, ns1; };
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// ---------------------------------------------------------------- V0: scalar, direct form, A atoms per lane
template <int THREADS, int A, int UNROLL>
__global__ void __launch_bounds__(THREADS) k_scalar(const float4 *__restrict__ rec, int n_pad, const float4 *__restrict__ lig,
                                                    int nl, int reps, float rinv_max, double *out, Timing *tm) {
    extern __shared__ float4 lf[];
    for (int i = threadIdx.x; i < nl; i += THREADS) lf[i] = lig[(size_t)blockIdx.x % 64 * nl + i];
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) { tm->clk0 = clock64(); tm->ns0 = gtime(); }
    double e = 0.0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int base = 0; base + THREADS * A <= n_pad; base += THREADS * A) {
            float4 P[A]; float s[A];
#pragma unroll
            for (int a = 0; a < A; ++a) { P[a] = __ldg(rec + base + a * THREADS + threadIdx.x); s[a] = 0.f; }
#pragma unroll UNROLL
            for (int l = 0; l < nl; ++l) {
                const float4 L = lf[l];
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    const float dx = P[a].x - L.x, dy = P[a].y - L.y, dz = P[a].z - L.z;
                    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    const float r = fminf(rsq(d2), rinv_max);
                    s[a] = fmaf(L.w, r, s[a]);
                }
            }
#pragma unroll
            for (int a = 0; a < A; ++a) e = fma((double)P[a].w, (double)s[a], e);
        }
        if (rep + 1 < reps) { __syncthreads(); if (threadIdx.x < nl) lf[threadIdx.x].x += 1e-3f; __syncthreads(); }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { tm->clk1 = clock64(); tm->ns1 = gtime(); }
    out[(size_t)blockIdx.x * THREADS + threadIdx.x] = e;
}

// ---------------------------------------------------------------- V1: packed f32x2, direct form, A2 atom PAIRS per lane
// ligand in smem as two float4: (-x,-x,-y,-y), (-z,-z,q,q)
template <int THREADS, int A2, int UNROLL>
__global__ void __launch_bounds__(THREADS) k_packed(const float4 *__restrict__ rec, int n_pad, const float4 *__restrict__ lig,
                                                    int nl, int reps, float rinv_max, double *out, Timing *tm) {
    extern __shared__ float4 lf[];
    for (int i = threadIdx.x; i < nl; i += THREADS) {
        const float4 L = lig[(size_t)blockIdx.x % 64 * nl + i];
        lf[2 * i] = make_float4(-L.x, -L.x, -L.y, -L.y);
        lf[2 * i + 1] = make_float4(-L.z, -L.z, L.w, L.w);
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) { tm->clk0 = clock64(); tm->ns0 = gtime(); }
    double e = 0.0;
    const ulonglong2 *lf2 = reinterpret_cast<const ulonglong2 *>(lf);
    for (int rep = 0; rep < reps; ++rep) {
        for (int base = 0; base + THREADS * A2 * 2 <= n_pad; base += THREADS * A2 * 2) {
            u64 px[A2], py[A2], pz[A2], s[A2]; float q0[A2], q1[A2];
#pragma unroll
            for (int a = 0; a < A2; ++a) {
                const float4 P0 = __ldg(rec + base + (2 * a) * THREADS + threadIdx.x);
                const float4 P1 = __ldg(rec + base + (2 * a + 1) * THREADS + threadIdx.x);
                px[a] = pack2(P0.x, P1.x); py[a] = pack2(P0.y, P1.y); pz[a] = pack2(P0.z, P1.z);
                q0[a] = P0.w; q1[a] = P1.w; s[a] = pack2(0.f, 0.f);
            }
#pragma unroll UNROLL
            for (int l = 0; l < nl; ++l) {
                const ulonglong2 Lxy = lf2[2 * l], Lzq = lf2[2 * l + 1];
#pragma unroll
                for (int a = 0; a < A2; ++a) {
                    const u64 dx = add2(px[a], Lxy.x), dy = add2(py[a], Lxy.y), dz = add2(pz[a], Lzq.x);
                    const u64 d2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
                    float d2a, d2b; unpack2(d2, d2a, d2b);
                    const float ra = fminf(rsq(d2a), rinv_max), rb = fminf(rsq(d2b), rinv_max);
                    s[a] = fma2(Lzq.y, pack2(ra, rb), s[a]);
                }
            }
#pragma unroll
            for (int a = 0; a < A2; ++a) { float sa, sb; unpack2(s[a], sa, sb); e = fma((double)q0[a], (double)sa, e); e = fma((double)q1[a], (double)sb, e); }
        }
        if (rep + 1 < reps) { __syncthreads(); if (threadIdx.x < nl) lf[2 * threadIdx.x].x += 1e-3f; __syncthreads(); }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { tm->clk1 = clock64(); tm->ns1 = gtime(); }
    out[(size_t)blockIdx.x * THREADS + threadIdx.x] = e;
}

// ---------------------------------------------------------------- V2: scalar, expanded form d2 = |p|^2 + |l|^2 - 2 p.l
// receptor float4 (x,y,z,q) + separate |p|^2 array; ligand smem float4 (-2x,-2y,-2z,|l|^2) + q array
template <int THREADS, int A, int UNROLL>
__global__ void __launch_bounds__(THREADS) k_expanded(const float4 *__restrict__ rec, const float *__restrict__ rec_pp, int n_pad,
                                                      const float4 *__restrict__ lig, int nl, int reps, float rinv_max, double *out, Timing *tm) {
    extern __shared__ float4 lf[];
    float *lq = reinterpret_cast<float *>(lf + nl);
    for (int i = threadIdx.x; i < nl; i += THREADS) {
        const float4 L = lig[(size_t)blockIdx.x % 64 * nl + i];
        lf[i] = make_float4(-2.f * L.x, -2.f * L.y, -2.f * L.z, L.x * L.x + L.y * L.y + L.z * L.z);
        lq[i] = L.w;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) { tm->clk0 = clock64(); tm->ns0 = gtime(); }
    double e = 0.0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int base = 0; base + THREADS * A <= n_pad; base += THREADS * A) {
            float4 P[A]; float pp[A], s[A];
#pragma unroll
            for (int a = 0; a < A; ++a) { P[a] = __ldg(rec + base + a * THREADS + threadIdx.x); pp[a] = __ldg(rec_pp + base + a * THREADS + threadIdx.x); s[a] = 0.f; }
#pragma unroll UNROLL
            for (int l = 0; l < nl; ++l) {
                const float4 L = lf[l]; const float q = lq[l];
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    const float d2 = fmaf(P[a].z, L.z, fmaf(P[a].y, L.y, fmaf(P[a].x, L.x, pp[a] + L.w)));
                    const float r = fminf(rsq(d2), rinv_max);
                    s[a] = fmaf(q, r, s[a]);
                }
            }
#pragma unroll
            for (int a = 0; a < A; ++a) e = fma((double)P[a].w, (double)s[a], e);
        }
        if (rep + 1 < reps) { __syncthreads(); if (threadIdx.x < nl) lf[threadIdx.x].x += 1e-3f; __syncthreads(); }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { tm->clk1 = clock64(); tm->ns1 = gtime(); }
    out[(size_t)blockIdx.x * THREADS + threadIdx.x] = e;
}

// ---------------------------------------------------------------- V3: packed, ligand-pair packing: lane holds A atoms, each f32x2 op
// covers TWO LIGAND atoms (l, l+1) against one receptor atom.  Receptor operand (px,px) is built once per tile.
// ligand smem: per ligand pair two float4: (-x0,-x1,-y0,-y1), (-z0,-z1,q0,q1)
template <int THREADS, int A, int UNROLL, int MODE = 0, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB) k_packed_lig(const float4 *__restrict__ rec, int n_pad, const float4 *__restrict__ lig,
                                                        int nl, int reps, float rinv_max, double *out, Timing *tm) {
    extern __shared__ float4 lf[];
    const int nl2 = nl / 2;      // nl even in this benchmark
    for (int i = threadIdx.x; i < nl2; i += THREADS) {
        const float4 L0 = lig[(size_t)blockIdx.x % 64 * nl + 2 * i], L1 = lig[(size_t)blockIdx.x % 64 * nl + 2 * i + 1];
        lf[2 * i] = make_float4(-L0.x, -L1.x, -L0.y, -L1.y);
        lf[2 * i + 1] = make_float4(-L0.z, -L1.z, L0.w, L1.w);
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) { tm->clk0 = clock64(); tm->ns0 = gtime(); }
    double e = 0.0;
    const ulonglong2 *lf2 = reinterpret_cast<const ulonglong2 *>(lf);
    for (int rep = 0; rep < reps; ++rep) {
        for (int base = 0; base + THREADS * A <= n_pad; base += THREADS * A) {
            u64 px[A], py[A], pz[A], s[A]; float q[A];
#pragma unroll
            for (int a = 0; a < A; ++a) {
                const float4 P = __ldg(rec + base + a * THREADS + threadIdx.x);
                px[a] = pack2(P.x, P.x); py[a] = pack2(P.y, P.y); pz[a] = pack2(P.z, P.z); q[a] = P.w; s[a] = pack2(0.f, 0.f);
            }
#pragma unroll UNROLL
            for (int l = 0; l < nl2; ++l) {
                const ulonglong2 Lxy = lf2[2 * l], Lzq = lf2[2 * l + 1];
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    const u64 dx = add2(px[a], Lxy.x), dy = add2(py[a], Lxy.y), dz = add2(pz[a], Lzq.x);
                    const u64 d2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
                    float d2a, d2b; unpack2(d2, d2a, d2b);
                    float ra, rb;
                    if (MODE == 1) { ra = d2a; rb = d2b; }
                    else if (MODE == 2) { ra = rsq(d2a); rb = rsq(d2b); }
                    else { ra = fminf(rsq(d2a), rinv_max); rb = fminf(rsq(d2b), rinv_max); }
                    s[a] = fma2(Lzq.y, pack2(ra, rb), s[a]);
                }
            }
#pragma unroll
            for (int a = 0; a < A; ++a) { float sa, sb; unpack2(s[a], sa, sb); e = fma((double)q[a], (double)(sa + sb), e); }
        }
        if (rep + 1 < reps) { __syncthreads(); if (threadIdx.x < nl2) lf[2 * threadIdx.x].x += 1e-3f; __syncthreads(); }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { tm->clk1 = clock64(); tm->ns1 = gtime(); }
    out[(size_t)blockIdx.x * THREADS + threadIdx.x] = e;
}


// ---------------------------------------------------------------- V4: packed, ligand-pair packing, EXPANDED form in the ligand-centred frame:
// d2 = |p'|^2 + |l'|^2 - 2 p'.l' with p' = p - c, l' = l - c (c = ligand centroid).  Per receptor atom and ligand PAIR:
// 3 FFMA2 + FADD2 (d2), 2 MUFU.RSQ, FFMA2 = 7 issue slots, 5 FMA-pipe ops (vs 9 / 7 in the direct clamp-free form).
// ligand smem per pair: (-2x0,-2x1,-2y0,-2y1), (-2z0,-2z1,l2_0,l2_1), (q0,q1)
template <int THREADS, int A, int UNROLL, int MODE = 0>
__global__ void __launch_bounds__(THREADS, (A > 8 && THREADS == 128) ? 4 : 1) k_plig_exp(const float4 *__restrict__ rec, int n_pad, const float4 *__restrict__ lig,
                                                        int nl, int reps, float rinv_max, double *out, Timing *tm) {
    extern __shared__ float4 lf[];
    const int nl2 = nl / 2;
    float2 *lq = reinterpret_cast<float2 *>(lf + 2 * nl2);
    __shared__ float cen[3];
    if (threadIdx.x == 0) {
        float cx = 0, cy = 0, cz = 0;
        for (int i = 0; i < nl; ++i) { const float4 L = lig[(size_t)blockIdx.x % 64 * nl + i]; cx += L.x; cy += L.y; cz += L.z; }
        cen[0] = cx / nl; cen[1] = cy / nl; cen[2] = cz / nl;
    }
    __syncthreads();
    const float cx = cen[0], cy = cen[1], cz = cen[2];
    for (int i = threadIdx.x; i < nl2; i += THREADS) {
        float4 L0 = lig[(size_t)blockIdx.x % 64 * nl + 2 * i], L1 = lig[(size_t)blockIdx.x % 64 * nl + 2 * i + 1];
        L0.x -= cx; L0.y -= cy; L0.z -= cz; L1.x -= cx; L1.y -= cy; L1.z -= cz;
        lf[2 * i] = make_float4(-2.f * L0.x, -2.f * L1.x, -2.f * L0.y, -2.f * L1.y);
        lf[2 * i + 1] = make_float4(-2.f * L0.z, -2.f * L1.z, L0.x * L0.x + L0.y * L0.y + L0.z * L0.z, L1.x * L1.x + L1.y * L1.y + L1.z * L1.z);
        lq[i] = make_float2(L0.w, L1.w);
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) { tm->clk0 = clock64(); tm->ns0 = gtime(); }
    double e = 0.0;
    const ulonglong2 *lf2 = reinterpret_cast<const ulonglong2 *>(lf);
    const u64 *lq2 = reinterpret_cast<const u64 *>(lq);
    for (int rep = 0; rep < reps; ++rep) {
        for (int base = 0; base + THREADS * A <= n_pad; base += THREADS * A) {
            u64 px[A], py[A], pz[A], pp[A], s[A]; float q[A];
#pragma unroll
            for (int a = 0; a < A; ++a) {
                const float4 P = __ldg(rec + base + a * THREADS + threadIdx.x);
                const float x = P.x - cx, y = P.y - cy, z = P.z - cz;
                const float p2 = fmaf(z, z, fmaf(y, y, x * x));
                px[a] = pack2(x, x); py[a] = pack2(y, y); pz[a] = pack2(z, z); pp[a] = pack2(p2, p2); q[a] = P.w; s[a] = pack2(0.f, 0.f);
            }
            ulonglong2 Lxy = lf2[0], Lzl = lf2[1];
            u64 Lq = lq2[0];
#pragma unroll UNROLL
            for (int l = 0; l < nl2; ++l) {
                if (MODE != 2) { Lxy = lf2[2 * l]; Lzl = lf2[2 * l + 1]; Lq = lq2[l]; }
                else { Lxy.x = add2(Lxy.x, Lq); }     // diagnostic: no LDS inside the loop (one extra FADD2 keeps the iterations distinct)
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    u64 d2;
                    if (MODE == 1) d2 = fma2(px[a], Lxy.x, fma2(py[a], Lxy.y, fma2(pz[a], Lzl.x, add2(pp[a], Lzl.y))));
                    else d2 = add2(fma2(px[a], Lxy.x, fma2(py[a], Lxy.y, fma2(pz[a], Lzl.x, Lzl.y))), pp[a]);
                    float d2a, d2b; unpack2(d2, d2a, d2b);
                    const float ra = rsq(d2a), rb = rsq(d2b);
                    s[a] = fma2(Lq, pack2(ra, rb), s[a]);
                }
            }
#pragma unroll
            for (int a = 0; a < A; ++a) { float sa, sb; unpack2(s[a], sa, sb); e = fma((double)q[a], (double)(sa + sb), e); }
        }
        if (rep + 1 < reps) { __syncthreads(); if (threadIdx.x < nl2) lf[2 * threadIdx.x].x += 1e-3f; __syncthreads(); }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { tm->clk1 = clock64(); tm->ns1 = gtime(); }
    out[(size_t)blockIdx.x * THREADS + threadIdx.x] = e;
}


// ---------------------------------------------------------------- V5: V4 with the rsqrt of the last NN of the A atoms per lane done on the FMA pipe
// (integer seed + tuned Newton step + one plain Newton step, max rel. error 7.6e-7): MUFU.RSQ is the binding pipe, the
// expanded form leaves the FMA pipe half idle, so a fraction of the rsqrt's moves over.
__device__ __forceinline__ u64 rsq2_fma(u64 d2) {
    float a, b; unpack2(d2, a, b);
    const float ya = __uint_as_float(0x5F1FFFF9u - (__float_as_uint(a) >> 1)), yb = __uint_as_float(0x5F1FFFF9u - (__float_as_uint(b) >> 1));
    u64 y = pack2(ya, yb);
    const u64 hk = mul2(d2, pack2(-0.703952253f, -0.703952253f)), h = mul2(d2, pack2(-0.5f, -0.5f));
    u64 t = mul2(y, y);
    t = fma2(hk, t, pack2(1.68191391f, 1.68191391f));
    y = mul2(y, t);
    t = mul2(y, y);
    t = fma2(h, t, pack2(1.5f, 1.5f));
    return mul2(y, t);
}
template <int THREADS, int A, int NN, int UNROLL>
__global__ void __launch_bounds__(THREADS, 1) k_plig_mix(const float4 *__restrict__ rec, int n_pad, const float4 *__restrict__ lig,
                                                        int nl, int reps, float rinv_max, double *out, Timing *tm) {
    extern __shared__ float4 lf[];
    const int nl2 = nl / 2;
    float2 *lq = reinterpret_cast<float2 *>(lf + 2 * nl2);
    __shared__ float cen[3];
    if (threadIdx.x == 0) {
        float cx = 0, cy = 0, cz = 0;
        for (int i = 0; i < nl; ++i) { const float4 L = lig[(size_t)blockIdx.x % 64 * nl + i]; cx += L.x; cy += L.y; cz += L.z; }
        cen[0] = cx / nl; cen[1] = cy / nl; cen[2] = cz / nl;
    }
    __syncthreads();
    const float cx = cen[0], cy = cen[1], cz = cen[2];
    for (int i = threadIdx.x; i < nl2; i += THREADS) {
        float4 L0 = lig[(size_t)blockIdx.x % 64 * nl + 2 * i], L1 = lig[(size_t)blockIdx.x % 64 * nl + 2 * i + 1];
        L0.x -= cx; L0.y -= cy; L0.z -= cz; L1.x -= cx; L1.y -= cy; L1.z -= cz;
        lf[2 * i] = make_float4(-2.f * L0.x, -2.f * L1.x, -2.f * L0.y, -2.f * L1.y);
        lf[2 * i + 1] = make_float4(-2.f * L0.z, -2.f * L1.z, L0.x * L0.x + L0.y * L0.y + L0.z * L0.z, L1.x * L1.x + L1.y * L1.y + L1.z * L1.z);
        lq[i] = make_float2(L0.w, L1.w);
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) { tm->clk0 = clock64(); tm->ns0 = gtime(); }
    double e = 0.0;
    const ulonglong2 *lf2 = reinterpret_cast<const ulonglong2 *>(lf);
    const u64 *lq2 = reinterpret_cast<const u64 *>(lq);
    for (int rep = 0; rep < reps; ++rep) {
        for (int base = 0; base + THREADS * A <= n_pad; base += THREADS * A) {
            u64 px[A], py[A], pz[A], pp[A], s[A]; float q[A];
#pragma unroll
            for (int a = 0; a < A; ++a) {
                const float4 P = __ldg(rec + base + a * THREADS + threadIdx.x);
                const float x = P.x - cx, y = P.y - cy, z = P.z - cz;
                const float p2 = fmaf(z, z, fmaf(y, y, x * x));
                px[a] = pack2(x, x); py[a] = pack2(y, y); pz[a] = pack2(z, z); pp[a] = pack2(p2, p2); q[a] = P.w; s[a] = pack2(0.f, 0.f);
            }
#pragma unroll UNROLL
            for (int l = 0; l < nl2; ++l) {
                const ulonglong2 Lxy = lf2[2 * l], Lzl = lf2[2 * l + 1];
                const u64 Lq = lq2[l];
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    const u64 d2 = add2(fma2(px[a], Lxy.x, fma2(py[a], Lxy.y, fma2(pz[a], Lzl.x, Lzl.y))), pp[a]);
                    if (a < A - NN) {
                        float d2a, d2b; unpack2(d2, d2a, d2b);
                        const float ra = rsq(d2a), rb = rsq(d2b);
                        s[a] = fma2(Lq, pack2(ra, rb), s[a]);
                    } else s[a] = fma2(Lq, rsq2_fma(d2), s[a]);
                }
            }
#pragma unroll
            for (int a = 0; a < A; ++a) { float sa, sb; unpack2(s[a], sa, sb); e = fma((double)q[a], (double)(sa + sb), e); }
        }
        if (rep + 1 < reps) { __syncthreads(); if (threadIdx.x < nl2) lf[2 * threadIdx.x].x += 1e-3f; __syncthreads(); }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { tm->clk1 = clock64(); tm->ns1 = gtime(); }
    out[(size_t)blockIdx.x * THREADS + threadIdx.x] = e;
}

// pure MUFU.RSQ issue rate: 8 independent chains per thread
__global__ void __launch_bounds__(128) k_mufu(int iters, float *out, Timing *tm) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 1.0f + threadIdx.x * 0.01f + i;
    if (blockIdx.x == 0 && threadIdx.x == 0) { tm->clk0 = clock64(); tm->ns0 = gtime(); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = rsq(v[i]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { tm->clk1 = clock64(); tm->ns1 = gtime(); }
    float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * 128 + threadIdx.x] = s;
}

struct Ctx {
    float4 *rec; float *pp; float4 *lig; double *out; Timing *tm; int n_pad, nl, sms; double max_mhz;
};

static const char *g_filter = "";
template <typename F>
void run(const char *name, F launch, const Ctx &c, int threads, int ctas_per_sm, int reps, size_t smem, int tile = 0) {
    if (g_filter[0] && !strstr(name, g_filter)) return;
    const int grid = c.sms * ctas_per_sm;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(grid, threads, smem, reps / 4 + 1);          // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f; Timing tm{};
    for (int it = 0; it < 3; ++it) {
        CK(cudaEventRecord(e0));
        launch(grid, threads, smem, reps);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) { printf("%-34s launch failed: %s\n", name, cudaGetErrorString(err)); return; }
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) { best = ms; CK(cudaMemcpy(&tm, c.tm, sizeof tm, cudaMemcpyDeviceToHost)); }
    }
    const double covered = tile > 0 ? (double)(c.n_pad / tile * tile) : (double)c.n_pad;     // atoms a pass really visits
    const double pairs = (double)grid * reps * covered * c.nl;
    const double mhz = (double)(tm.clk1 - tm.clk0) / (double)(tm.ns1 - tm.ns0) * 1e3;
    const double pps = pairs / (best * 1e-3);
    const double ppc = pps / (mhz * 1e6) / c.sms;
    const double peak = c.sms * 128.0 * 2.0 * c.max_mhz * 1e6;
    double h = 0; std::vector<double> o(16); CK(cudaMemcpy(o.data(), c.out, 16 * sizeof(double), cudaMemcpyDeviceToHost)); for (double v : o) h += v;
    printf("%-34s thr=%3d cta/sm=%2d  %8.3f ms  %7.3f Tpairs/s  clk=%6.0f MHz  %6.2f pairs/clk/SM  %5.1f%% of FP32 peak@%.0fMHz  (chk %.6e)\n",
           name, threads, ctas_per_sm, best, pps * 1e-12, mhz, ppc, 100.0 * 12.0 * pps / peak, c.max_mhz, h);
    fflush(stdout);
}

int main(int argc, char **argv) {
    const int n_atoms = argc > 1 ? atoi(argv[1]) : 5000;
    const int nl = argc > 2 ? atoi(argv[2]) : 40;
    const int reps = argc > 3 ? atoi(argv[3]) : 200;
    const char *filter = argc > 4 ? argv[4] : "";
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    Ctx c; c.sms = prop.multiProcessorCount; c.max_mhz = khz / 1000.0; c.nl = nl;
    c.n_pad = (n_atoms + 1023) / 1024 * 1024;           // multiple of every tile used below
    printf("# %s, %d SMs, max clock %.0f MHz, receptor %d (padded %d) x ligand %d, reps %d\n", prop.name, c.sms, c.max_mhz, n_atoms, c.n_pad, nl, reps);
    std::vector<float4> hrec(c.n_pad); std::vector<float> hpp(c.n_pad); std::vector<float4> hlig(64 * nl);
    srand(1);
    auto U = []() { return (float)rand() / RAND_MAX; };
    for (int i = 0; i < c.n_pad; ++i) {
        float x = (U() - .5f) * 64, y = (U() - .5f) * 64, z = (U() - .5f) * 64;
        hrec[i] = make_float4(x, y, z, i < n_atoms ? (U() - .5f) * .4f : 0.f); hpp[i] = x * x + y * y + z * z;
    }
    for (int i = 0; i < 64 * nl; ++i) hlig[i] = make_float4(36 + U() * 6, U() * 6, U() * 6, (U() - .5f) * .4f);
    CK(cudaMalloc(&c.rec, sizeof(float4) * c.n_pad)); CK(cudaMalloc(&c.pp, sizeof(float) * c.n_pad));
    CK(cudaMalloc(&c.lig, sizeof(float4) * 64 * nl)); CK(cudaMalloc(&c.out, sizeof(double) * c.sms * 32 * 256)); CK(cudaMalloc(&c.tm, sizeof(Timing)));
    CK(cudaMemcpy(c.rec, hrec.data(), sizeof(float4) * c.n_pad, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.pp, hpp.data(), sizeof(float) * c.n_pad, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.lig, hlig.data(), sizeof(float4) * 64 * nl, cudaMemcpyHostToDevice));
    g_filter = filter;
    const float rmax = 1e6f;
    const size_t sm1 = sizeof(float4) * nl, sm2 = sizeof(float4) * 2 * nl, sm3 = sizeof(float4) * nl + sizeof(float) * nl;

#define RUN_SCALAR(T, A, U_, OCC) run("scalar T" #T " A" #A " U" #U_, [&](int g, int t, size_t s, int r) { k_scalar<T, A, U_><<<g, t, s>>>(c.rec, c.n_pad, c.lig, nl, r, rmax, c.out, c.tm); }, c, T, OCC, reps, sm1)
#define RUN_PACKED(T, A, U_, OCC) run("packed-rec T" #T " A2=" #A " U" #U_, [&](int g, int t, size_t s, int r) { k_packed<T, A, U_><<<g, t, s>>>(c.rec, c.n_pad, c.lig, nl, r, rmax, c.out, c.tm); }, c, T, OCC, reps, sm2)
#define RUN_PLIG(T, A, U_, OCC) run("packed-lig T" #T " A" #A " U" #U_, [&](int g, int t, size_t s, int r) { k_packed_lig<T, A, U_><<<g, t, s>>>(c.rec, c.n_pad, c.lig, nl, r, rmax, c.out, c.tm); }, c, T, OCC, reps, sm2)
#define RUN_EXP(T, A, U_, OCC) run("expanded T" #T " A" #A " U" #U_, [&](int g, int t, size_t s, int r) { k_expanded<T, A, U_><<<g, t, s>>>(c.rec, c.pp, c.n_pad, c.lig, nl, r, rmax, c.out, c.tm); }, c, T, OCC, reps, sm3)

    RUN_SCALAR(128, 4, 2, 4); RUN_SCALAR(128, 4, 2, 7); RUN_SCALAR(128, 4, 2, 8);
    RUN_SCALAR(128, 2, 2, 8); RUN_SCALAR(128, 8, 1, 4); RUN_SCALAR(128, 8, 1, 7);
    RUN_SCALAR(256, 4, 2, 4); RUN_SCALAR(128, 4, 4, 7); RUN_SCALAR(128, 4, 1, 7);
    RUN_PACKED(128, 2, 2, 4); RUN_PACKED(128, 2, 2, 7); RUN_PACKED(128, 4, 1, 4); RUN_PACKED(128, 4, 1, 7); RUN_PACKED(128, 2, 4, 7);
    RUN_PLIG(128, 4, 2, 4); RUN_PLIG(128, 4, 2, 7); RUN_PLIG(128, 8, 1, 4); RUN_PLIG(128, 2, 2, 8);
    RUN_EXP(128, 4, 2, 4); RUN_EXP(128, 4, 2, 7); RUN_EXP(128, 8, 1, 4);
#define RUN_PLIGM(T, A, U_, M, MB, OCC) run("plig-mode" #M " T" #T " A" #A " U" #U_ " minb" #MB, [&](int g, int t, size_t s, int r) { k_packed_lig<T, A, U_, M, MB><<<g, t, s>>>(c.rec, c.n_pad, c.lig, nl, r, rmax, c.out, c.tm); }, c, T, OCC, reps, sm2, T * A)
    RUN_PLIGM(128, 4, 2, 1, 1, 7); RUN_PLIGM(128, 4, 2, 2, 1, 7);
    RUN_PLIGM(128, 4, 2, 0, 1, 8); RUN_PLIGM(128, 4, 2, 0, 1, 9); RUN_PLIGM(128, 4, 4, 0, 1, 8);
    RUN_PLIGM(128, 2, 4, 0, 1, 12); RUN_PLIGM(128, 2, 2, 0, 1, 12); RUN_PLIGM(128, 2, 2, 0, 1, 16);
    RUN_PLIGM(256, 4, 2, 0, 1, 4); RUN_PLIGM(64, 4, 2, 0, 1, 16); RUN_PLIGM(64, 8, 1, 0, 1, 12); RUN_PLIGM(32, 8, 1, 0, 1, 28); RUN_PLIGM(32, 4, 2, 0, 1, 32);
#define RUN_PEXP(T, A, U_, M, OCC) run("plig-exp" #M " T" #T " A" #A " U" #U_, [&](int g, int t, size_t s, int r) { k_plig_exp<T, A, U_, M><<<g, t, s>>>(c.rec, c.n_pad, c.lig, nl, r, rmax, c.out, c.tm); }, c, T, OCC, reps, sm2 + sizeof(float2) * nl, T * A)
    RUN_PEXP(128, 4, 2, 0, 4); RUN_PEXP(128, 8, 1, 0, 4); RUN_PEXP(128, 4, 2, 0, 3); RUN_PEXP(128, 8, 1, 0, 3); RUN_PEXP(128, 8, 2, 0, 4); RUN_PEXP(128, 4, 4, 0, 4);
    RUN_PEXP(128, 4, 2, 1, 4); RUN_PEXP(128, 8, 1, 1, 4); RUN_PEXP(128, 6, 1, 0, 4); RUN_PEXP(128, 2, 4, 0, 6); RUN_PEXP(256, 4, 2, 0, 2); RUN_PEXP(512, 4, 2, 0, 1); RUN_PEXP(512, 8, 1, 0, 1);
    RUN_PLIGM(512, 8, 1, 2, 1, 1); RUN_PLIGM(512, 8, 1, 0, 1, 1);
    RUN_PEXP(128, 8, 1, 2, 4); RUN_PEXP(128, 4, 2, 2, 4); RUN_PEXP(128, 10, 1, 0, 4); RUN_PEXP(128, 20, 1, 0, 4); RUN_PEXP(128, 20, 1, 0, 2); RUN_PEXP(128, 10, 1, 0, 5); RUN_PEXP(128, 8, 1, 0, 5); RUN_PEXP(128, 8, 1, 0, 6); RUN_PEXP(128, 8, 1, 0, 8);
    RUN_PEXP(256, 10, 1, 0, 2); RUN_PEXP(512, 10, 1, 0, 1); RUN_PEXP(512, 5, 1, 0, 1); RUN_PEXP(512, 5, 2, 0, 1); RUN_PEXP(256, 20, 1, 0, 1); RUN_PEXP(256, 10, 1, 0, 1);
#define RUN_PMIX(T, A, NN, U_, OCC) run("plig-mix T" #T " A" #A " N" #NN " U" #U_, [&](int g, int t, size_t s, int r) { k_plig_mix<T, A, NN, U_><<<g, t, s>>>(c.rec, c.n_pad, c.lig, nl, r, rmax, c.out, c.tm); }, c, T, OCC, reps, sm2 + sizeof(float2) * nl, T * A)
    RUN_PMIX(128, 8, 0, 1, 4); RUN_PMIX(128, 8, 1, 1, 4); RUN_PMIX(128, 8, 2, 1, 4); RUN_PMIX(128, 8, 3, 1, 4); RUN_PMIX(128, 8, 8, 1, 4);
    RUN_PMIX(128, 10, 1, 1, 4); RUN_PMIX(128, 10, 2, 1, 4); RUN_PMIX(128, 5, 1, 2, 4); RUN_PMIX(128, 4, 1, 2, 4); RUN_PMIX(128, 8, 1, 1, 8); RUN_PMIX(128, 8, 2, 1, 8);
    RUN_PMIX(512, 10, 1, 1, 1); RUN_PMIX(512, 10, 2, 1, 1); RUN_PMIX(512, 5, 1, 1, 1); RUN_PMIX(512, 5, 1, 2, 1);
    if (!filter[0] || strstr("mufu-only", filter)) {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float *o; CK(cudaMalloc(&o, sizeof(float) * c.sms * 16 * 128));
        k_mufu<<<c.sms * 16, 128>>>(1000, o, c.tm); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); k_mufu<<<c.sms * 16, 128>>>(20000, o, c.tm); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        Timing tm; CK(cudaMemcpy(&tm, c.tm, sizeof tm, cudaMemcpyDeviceToHost));
        const double mhz = (double)(tm.clk1 - tm.clk0) / (double)(tm.ns1 - tm.ns0) * 1e3;
        const double n = (double)c.sms * 16 * 128 * 8 * 20000;
        printf("mufu-only: %.3f ms, clk %.0f MHz, %.2f MUFU.RSQ/clk/SM\n", ms, mhz, n / (ms * 1e-3) / (mhz * 1e6) / c.sms);
    }
    return 0;
}
