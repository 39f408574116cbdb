// ddd_kernels.cuh -- sm_100a device code of the Metropolis Monte-Carlo hot path.
//
// One CTA (32..256 threads, chosen from the chain count) runs one independent ligand chain of
// SimulateEnergyMinimizationOneProc (reference metropolisMethod/metropolis.go:116-136)
// from start to finish on device: proposal (:154-208), collision test (:229-242),
// protein x ligand Coulomb sum (:247-262) and AcceptMove (:141-149), with no host round trip.
//
// Compiled with -fmad=false: plain fp64 expressions below are NOT contracted, so the
// per-term arithmetic of the F64 paths is bit-identical to Go/amd64 (which never fuses);
// every fused multiply-add in this file is an explicit fmaf()/fma().
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdio.h>

namespace ddd {

constexpr int kThreads = 128;   // threads per CTA of the batched helper kernels (energy / closest / randomize)
constexpr int kTileA   = 4;     // receptor atoms held in registers per lane per tile (fp32 path)
constexpr int kRecPad  = 256;   // receptor fp32 array is padded to a multiple of the widest chain CTA
constexpr float kPadCoord = 1.0e4f;   // padding atoms: q = 0, parked 1e4 A away (receptor +, ligand -)

// receptor atom, fp64 master copy: absolute coordinates, Go Atom layout (datatypes.go:18-25)
struct __align__(32) Atom64 { double x, y, z, q; };

struct DevParams {
    double k_coulomb, min_distance, max_angle, jitter_width, clamp_distance;
    double rigid_translation, rigid_angle, temperature;
    long long max_retries;
    float rinv_max;              // 1 / clamp_distance: clamp d >= 1e-6  <=>  1/d <= 1e6 (:255-257)
    float apart_gap;             // receptor-group / pose boxes farther apart than this skip the clamp: max(1e-3, 2.5 clamp_distance)
    int move_mode, rotate;
    double cutoff;               // EXTENSION: > 0 = shifted-Coulomb cutoff radius (0: the reference's untruncated sum)
    float rc_inv;                // 1 / cutoff (0 when off)
    double lj_scale, lj_cap;     // EXTENSION: Lennard-Jones 12-6 on top of the Coulomb term (lj_scale == 0: off)
    float lj_c32, lj_cap32;      // 4 lj_scale / K (the fp32 sums are multiplied by K at the end), the cap on (sigma/d)^2
};

struct RecView {
    const float4 *f32;           // (x-cx, y-cy, z-cz, q), padded to a multiple of kRecPad
    const Atom64 *f64;           // absolute, n atoms
    int n, n_pad;
    double cx, cy, cz;           // receptor centroid: origin of the fp32 frame
    // cell list for the cutoff extension: the same fp32 atoms sorted along a Morton curve of 6 A grid cells, cut into
    // units of 32 consecutive atoms (one per lane); ubox[2u], ubox[2u+1] = (min xyz, max xyz) of unit u's real atoms
    const float4 *f32s;
    const Atom64 *f64s;          // the fp64 atoms in the same Morton order (absolute coordinates)
    const float4 *ubox;
    const float4 *gbox;          // (min xyz, max xyz) of every group of 4 consecutive units (128 atoms)
    int n_units_s;
    // LJ extension: (sigma/2, sqrt(eps)) per atom; fp64 in input order and in Morton order, fp32 in Morton order
    const double2 *lj64, *lj64s;
    const float2 *lj32s;
};

// ------------------------------------------------------------------ Philox4x32-10
// Salmon et al. SC'11 (Random123).  Same counter/key convention as the host side:
// counter = (index/2 lo, index/2 hi, stream lo, stream hi), key = (seed lo, seed hi);
// words (0,1) -> draw 2k, words (2,3) -> draw 2k+1, 53-bit mantissa, [0,1).
// UNROLL: the ten rounds are kept rolled where the call is off the hot path (accept draw, axis batch, per-atom draws of
// oversized ligands): the chain kernels are large and a warp in its serial phase is bound by instruction fetch
// (ncu: no_instruction), not by issue slots.  The staged jitter draws of the proposal use the unrolled form.
template <int UNROLL>
__host__ __device__ __forceinline__ void philox4x32_10_t(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll UNROLL
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
#ifndef DDD_PHILOX_UNROLL
#define DDD_PHILOX_UNROLL 1
#endif
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
    philox4x32_10_t<DDD_PHILOX_UNROLL>(c0, c1, c2, c3, k0, k1, out);
}

__host__ __device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}

__host__ __device__ __forceinline__ double philox_u01(uint64_t seed, uint64_t stream, uint64_t index) {
    const uint64_t blk = index >> 1;
    uint32_t o[4];
    philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)stream, (uint32_t)(stream >> 32),
                  (uint32_t)seed, (uint32_t)(seed >> 32), o);
    return (index & 1) ? u53(o[2], o[3]) : u53(o[0], o[1]);
}

// Uniform stream of one chain: Philox, or a recorded array (replay of a Go math/rand stream).
struct Stream {
    const double *rec;           // nullptr -> Philox
    long long rec_len;
    uint64_t seed, id;
    __device__ __forceinline__ double draw(uint64_t index, int &status) const {
        if (rec) {
            if ((long long)index >= rec_len) { status |= 2; return 0.5; }
            return rec[index];
        }
        return philox_u01(seed, id, index);
    }
    // three consecutive draws (one atom's X,Y,Z jitter) from two Philox blocks
    __device__ __forceinline__ void draw3(uint64_t index, double u[3], int &status) const {
        if (rec) {
            u[0] = draw(index, status); u[1] = draw(index + 1, status); u[2] = draw(index + 2, status);
            return;
        }
        const uint64_t blk = index >> 1;
        uint32_t a[4], b[4];
        philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)id, (uint32_t)(id >> 32),
                      (uint32_t)seed, (uint32_t)(seed >> 32), a);
        philox4x32_10((uint32_t)(blk + 1), (uint32_t)((blk + 1) >> 32), (uint32_t)id, (uint32_t)(id >> 32),
                      (uint32_t)seed, (uint32_t)(seed >> 32), b);
        const double v0 = u53(a[0], a[1]), v1 = u53(a[2], a[3]), v2 = u53(b[0], b[1]), v3 = u53(b[2], b[3]);
        if (index & 1) { u[0] = v1; u[1] = v2; u[2] = v3; } else { u[0] = v0; u[1] = v1; u[2] = v2; }
    }
    // four consecutive draws (axis X,Y,Z and theta): two Philox blocks when index is even, three when odd
    __device__ __forceinline__ void draw4(uint64_t index, double u[4], int &status) const {
        if (rec) {
            for (int k = 0; k < 4; ++k) u[k] = draw(index + k, status);
            return;
        }
        const uint64_t blk = index >> 1;
        uint32_t a[4], b[4];
        philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)id, (uint32_t)(id >> 32),
                      (uint32_t)seed, (uint32_t)(seed >> 32), a);
        philox4x32_10((uint32_t)(blk + 1), (uint32_t)((blk + 1) >> 32), (uint32_t)id, (uint32_t)(id >> 32),
                      (uint32_t)seed, (uint32_t)(seed >> 32), b);
        if (index & 1) {
            uint32_t c[4];
            philox4x32_10((uint32_t)(blk + 2), (uint32_t)((blk + 2) >> 32), (uint32_t)id, (uint32_t)(id >> 32),
                          (uint32_t)seed, (uint32_t)(seed >> 32), c);
            u[0] = u53(a[2], a[3]); u[1] = u53(b[0], b[1]); u[2] = u53(b[2], b[3]); u[3] = u53(c[0], c[1]);
        } else {
            u[0] = u53(a[0], a[1]); u[1] = u53(a[2], a[3]); u[2] = u53(b[0], b[1]); u[3] = u53(b[2], b[3]);
        }
    }
};

// ------------------------------------------------------------------ reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block sum returned to every thread.  `red` holds 2 x kMaxWarps doubles used
// alternately (parity) so that back-to-back calls need only the one barrier inside.
constexpr int kMaxWarps = 8;
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double *red, int &parity) {
    v = warp_sum(v);
    if (THREADS == 32) return v;
    double *buf = red + parity * kMaxWarps;
    parity ^= 1;
    if ((threadIdx.x & 31) == 0) buf[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = buf[0];
#pragma unroll
    for (int w = 1; w < THREADS / 32; ++w) s += buf[w];
    return s;
}

// packed fp32x2 arithmetic (sm_100 FADD2 / FMUL2 / FFMA2): one instruction, two lanes of a register pair
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

__device__ __forceinline__ float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ------------------------------------------------------------------ pair sums (CalculateEnergy :247-262)
// fp32 path.  A lane holds A receptor atoms in registers; the ligand sits in shared memory packed
// by PAIRS of atoms (2k, 2k+1) in the receptor-centred frame:
//     lf[2k] = (-x0, -x1, -y0, -y1)     lf[2k+1] = (-z0, -z1, q0, q1)
// so that every FADD2 / FMUL2 / FFMA2 evaluates two ligand atoms against one receptor atom (the
// receptor coordinate enters as the broadcast-scalar operand).  Per receptor atom and ligand PAIR:
// 3 FADD2, FMUL2 + 2 FFMA2 (d^2), 2 MUFU.RSQ, 2 FMNMX (clamp d >= 1e-6 <=> 1/d <= 1e6), FFMA2
// (s += q_l / d): 11 issue slots per 2 pair terms.  MUFU (16/clk/SM) is the binding pipe.
// q_p is applied once per tile and K once per call.  An odd ligand is padded with a q = 0 atom.
// CUT (the cutoff extension): every term is the shifted Coulomb term q_l (max(min(1/d, 1/clamp), 1/cutoff) - 1/cutoff), which is
// exactly 0 from the cutoff outwards, so no distance compare is needed.  The loop sums q_l * max(.., 1/cutoff) (2 FMNMX per
// ligand pair on top of the plain loop) and the constant part, (1/cutoff) * sum_l q_l = `rc_q`, is subtracted once per
// receptor atom at the end.
// `ubase[a]` is the first atom of the a-th unit of 32 receptor atoms this call covers (units need not be adjacent).
// CLAMP = false drops the two FMNMX of the d >= 1e-6 clamp (:255-257): the caller has established that no atom of these
// units can be within the clamp distance of the ligand (their boxes are apart), so min(1/d, 1/clamp) == 1/d bit for bit.
template <int A, bool CUT, bool CLAMP = true>
__device__ __forceinline__ float warp_atoms_sum_f32(const float4 (&P)[A], const ulonglong2 *__restrict__ lf2, int n_pairs,
                                                    float rinv_max, float rc_inv, float *t_lo = nullptr, float rc_q = 0.f) {
    f32x2 px[A], py[A], pz[A], s[A];
#pragma unroll
    for (int a = 0; a < A; ++a) {
        px[a] = pack2(P[a].x, P[a].x); py[a] = pack2(P[a].y, P[a].y); pz[a] = pack2(P[a].z, P[a].z);
        s[a] = pack2(0.f, 0.f);
    }
    constexpr int kPairUnroll = A >= 8 ? 1 : 2;          // measured: 8 atoms per lane already fill the pipes
#pragma unroll kPairUnroll
    for (int k = 0; k < n_pairs; ++k) {
        const ulonglong2 Lxy = lf2[2 * k], Lzq = lf2[2 * k + 1];
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const f32x2 dx = add2(px[a], Lxy.x), dy = add2(py[a], Lxy.y), dz = add2(pz[a], Lzq.x);
            const f32x2 d2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
            float d2a, d2b;
            unpack2(d2, d2a, d2b);
            float ra = rsqrt_approx(d2a), rb = rsqrt_approx(d2b);
            if (CLAMP) { ra = fminf(ra, rinv_max); rb = fminf(rb, rinv_max); }
            if (CUT) { ra = fmaxf(ra, rc_inv); rb = fmaxf(rb, rc_inv); }
            s[a] = fma2(Lzq.y, pack2(ra, rb), s[a]);
        }
    }
    float t = 0.f;                       // fp32 within the lane's A * nl terms, fp64 beyond
#pragma unroll
    for (int a = 0; a < A; ++a) {
        if (A == 8 && a == 4 && t_lo) { *t_lo = t; t = 0.f; }           // 8 atoms per lane = two groups of 4, summed separately
        float sa, sb; unpack2(s[a], sa, sb);
        t = fmaf(P[a].w, CUT ? (sa + sb) - rc_q : sa + sb, t);
    }
    return t;
}

// LJ extension: Coulomb + Lennard-Jones 12-6 for A receptor atoms per lane.  Per ligand pair and receptor atom, on top of
// the Coulomb ops: s = sigma_p/2 + sigma_l/2, x = min((s / d)^2, cap), x3 = x^3, lj += sqrt(eps_l) (x3^2 - x3):
// FADD2, 4 FMUL2, 2 FFMA2, 2 FMNMX.  Returns  sum_p [ q_p sum_l q_l/d  +  c_lj sqrt(eps_p) sum_l sqrt(eps_l)(x^6 - x^3) ].
// lj2[k] = (sigma0/2, sigma1/2, sqrt(eps0), sqrt(eps1)) of ligand pair k.
template <int A, bool CUT>
__device__ __forceinline__ float warp_units_sum_lj_f32(const float4 *__restrict__ rec, const float2 *__restrict__ rec_lj, const int (&ubase)[A],
                                                       const ulonglong2 *__restrict__ lf2, const ulonglong2 *__restrict__ lj2, int n_pairs,
                                                       float rinv_max, float rc_inv, float c_lj, float cap) {
    f32x2 px[A], py[A], pz[A], hs[A], s[A], l[A];
    float q[A], se[A];
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int a = 0; a < A; ++a) {
        const float4 P = __ldg(rec + ubase[a] + lane);
        const float2 J = __ldg(rec_lj + ubase[a] + lane);
        px[a] = pack2(P.x, P.x); py[a] = pack2(P.y, P.y); pz[a] = pack2(P.z, P.z); hs[a] = pack2(J.x, J.x);
        q[a] = P.w; se[a] = J.y; s[a] = pack2(0.f, 0.f); l[a] = pack2(0.f, 0.f);
    }
    const f32x2 m_rc = pack2(-rc_inv, -rc_inv);
    for (int k = 0; k < n_pairs; ++k) {
        const ulonglong2 Lxy = lf2[2 * k], Lzq = lf2[2 * k + 1], Lj = lj2[k];
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const f32x2 dx = add2(px[a], Lxy.x), dy = add2(py[a], Lxy.y), dz = add2(pz[a], Lzq.x);
            const f32x2 d2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
            float d2a, d2b;
            unpack2(d2, d2a, d2b);
            const float ua = rsqrt_approx(d2a), ub = rsqrt_approx(d2b);          // 1/d, unclamped: the LJ core has its own cap
            float ra = fminf(ua, rinv_max), rb = fminf(ub, rinv_max);
            const f32x2 t = mul2(add2(hs[a], Lj.x), pack2(ua, ub));              // sigma_pl / d
            float xa, xb;
            unpack2(mul2(t, t), xa, xb);
            xa = fminf(xa, cap); xb = fminf(xb, cap);
            if (CUT) {
                unpack2(add2(pack2(ra, rb), m_rc), ra, rb);
                if (!(ra > 0.f)) xa = 0.f;                                       // beyond the cutoff: no LJ either
                if (!(rb > 0.f)) xb = 0.f;
                ra = fmaxf(ra, 0.f); rb = fmaxf(rb, 0.f);
            }
            const f32x2 x = pack2(xa, xb);
            const f32x2 x3 = mul2(mul2(x, x), x);
            l[a] = fma2(Lj.y, fma2(x3, x3, mul2(x3, pack2(-1.f, -1.f))), l[a]);
            s[a] = fma2(Lzq.y, pack2(ra, rb), s[a]);
        }
    }
    float t = 0.f;
#pragma unroll
    for (int a = 0; a < A; ++a) {
        float sa, sb, la, lb;
        unpack2(s[a], sa, sb); unpack2(l[a], la, lb);
        t = fmaf(q[a], sa + sb, t);
        t = fmaf(c_lj * se[a], la + lb, t);
    }
    return t;
}

template <int A, bool CUT, bool CLAMP = true>
__device__ __forceinline__ float warp_units_sum_f32(const float4 *__restrict__ rec, const int (&ubase)[A],
                                                    const ulonglong2 *__restrict__ lf2, int n_pairs, float rinv_max, float rc_inv,
                                                    float *t_lo = nullptr, float rc_q = 0.f) {
    float4 P[A];
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int a = 0; a < A; ++a) P[a] = __ldg(rec + ubase[a] + lane);
    return warp_atoms_sum_f32<A, CUT, CLAMP>(P, lf2, n_pairs, rinv_max, rc_inv, t_lo, rc_q);
}

template <int A, bool CUT = false>
__device__ __forceinline__ float warp_tile_sum_f32(const float4 *__restrict__ rec, int base,
                                                   const ulonglong2 *__restrict__ lf2, int n_pairs, float rinv_max, float rc_inv = 0.f, float rc_q = 0.f) {
    int ub[A];
#pragma unroll
    for (int a = 0; a < A; ++a) ub[a] = base + a * 32;
    return warp_units_sum_f32<A, CUT>(rec, ub, lf2, n_pairs, rinv_max, rc_inv, nullptr, rc_q);
}

// sum_p sum_l q_p q_l / max(d, clamp) for the whole CTA, returned to every thread (K applied by the caller).
// The receptor is cut into work items of `chunk` warp tiles (32 lanes x kTileA atoms); the warps of the CTA
// pull items from a shared-memory counter, so a warp that the scheduler favours simply takes more items
// instead of idling at the barrier.  Each item's sum is stored under its item id and the items are added
// in id order afterwards: the result does not depend on which warp computed what (bit-reproducible).
// `ctl[0]` must be 0 on entry (reset by one thread before the barrier that precedes the call).
constexpr int kWarpTile = 32 * kTileA;
// work items a CTA of `threads` threads stages in shared memory (12 per warp; a single-warp CTA needs none)
__host__ __device__ constexpr int max_items(int threads) { return threads > 32 ? 12 * (threads / 32) : 0; }
#ifndef DDD_DYNAMIC_TILES
#define DDD_DYNAMIC_TILES 1
#endif
template <int THREADS, bool CUT = false>
__device__ __forceinline__ double cta_pair_sum_f32(const RecView &rv, const float4 *lf, int nl, float rinv_max,
                                                   int *ctl, double *item_out, double *red, int &parity, float rc_inv = 0.f, float rc_q = 0.f) {
    const ulonglong2 *lf2 = reinterpret_cast<const ulonglong2 *>(lf);
    const int n_pairs = (nl + 1) >> 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_wt = (rv.n + kWarpTile - 1) / kWarpTile;                 // n_pad is a multiple of kRecPad >= kWarpTile
    if (THREADS == 32) {                                                  // one warp: tiles in order, no staging
        double e1 = 0.0;
        for (int t = 0; t < n_wt; ++t) e1 += (double)warp_tile_sum_f32<kTileA, CUT>(rv.f32, t * kWarpTile, lf2, n_pairs, rinv_max, rc_inv, rc_q);
        return warp_sum(e1);
    }
    constexpr int kMaxItems = max_items(THREADS) > 0 ? max_items(THREADS) : 1;
    const int chunk = (n_wt + kMaxItems - 1) / kMaxItems;
    const int n_items = chunk ? (n_wt + chunk - 1) / chunk : 0;
    int next_static = warp;
    for (;;) {
        int it = 0;
        if (DDD_DYNAMIC_TILES) {
            if (lane == 0) it = atomicAdd(ctl, 1);
            it = __shfl_sync(0xffffffffu, it, 0);
        } else { it = next_static; next_static += THREADS / 32; }
        if (it >= n_items) break;
        const int t_end = min((it + 1) * chunk, n_wt);
        double acc = 0.0;
        for (int t = it * chunk; t < t_end; ++t)
            acc += (double)warp_tile_sum_f32<kTileA, CUT>(rv.f32, t * kWarpTile, lf2, n_pairs, rinv_max, rc_inv, rc_q);
        item_out[it * 32 + lane] = acc;                                   // per item AND per lane: no reduction here
    }
    __syncthreads();
    // fixed summation order, independent of which warp computed which item: thread (warp, lane) adds lane `lane`
    // of items warp, warp + W, ... ; then the deterministic block sum
    double e = 0.0;
    for (int it = warp; it < n_items; it += THREADS / 32) e += item_out[it * 32 + lane];
    return block_sum<THREADS>(e, red, parity);
}

// sum of the ligand's fp32 charges as stored in the pair-packed pose, in atom order (every thread computes the same value):
// the constant part of the shifted-Coulomb cutoff terms
__device__ __forceinline__ float ligand_charge_sum_f32(const float4 *lf, int nl) {
    const float *f = reinterpret_cast<const float *>(lf);
    float q = 0.f;
    for (int i = 0; i < nl; ++i) q += f[8 * (i >> 1) + 6 + (i & 1)];
    return q;
}

// write atom i of the fp32 pose in the pair-packed layout
__device__ __forceinline__ void store_lig_f32(float4 *lf, int i, float x, float y, float z, float q) {
    float *f = reinterpret_cast<float *>(lf + 2 * (i >> 1)) + (i & 1);
    f[0] = -x; f[2] = -y; f[4] = -z; f[6] = q;
}

// fp64 path: the reference's per-term arithmetic, term by term (no contraction: -fmad=false).
template <int THREADS, bool WITH_ABS>
__device__ __forceinline__ double pair_sum_f64(const RecView &rv, const double *lx, const double *ly,
                                               const double *lz, const double *lq, int nl,
                                               double K, double clampd, double *abs_out, double cutoff = 0.0) {
    double e = 0.0, ab = 0.0;
    for (int p = threadIdx.x; p < rv.n; p += THREADS) {
        const Atom64 P = rv.f64[p];
        for (int l = 0; l < nl; ++l) {
            const double ddx = P.x - lx[l], ddy = P.y - ly[l], ddz = P.z - lz[l];
            double d = sqrt(ddx * ddx + ddy * ddy + ddz * ddz);          // :253 (Pow(x,2) == x*x)
            if (cutoff > 0 && !(d < cutoff)) continue;                    // EXTENSION: shifted-Coulomb cutoff
            if (d < clampd) d = clampd;                                   // :255-257
            double term = K * (P.q * lq[l]) / d;                          // :258
            if (cutoff > 0) term -= K * (P.q * lq[l]) / cutoff;
            e += term;
            if (WITH_ABS) ab += fabs(term);
        }
    }
    if (WITH_ABS) *abs_out = ab;
    return e;
}

// ------------------------------------------------------------------ chain kernel
struct ChainArgs {
    RecView rec;
    DevParams prm;
    const long long *offsets;     // global CSR, n_lig+1
    const double *lig_xyzq;       // atoms [atom_base, ...) of the start poses, 4 doubles each
    long long atom_base;          // first atom index present in lig_xyzq
    const int *order;             // nullable: CTA -> local chain (longest ligands first)
    long long chain_begin;        // first chain (index within the call's ligand batch) of this launch
    uint64_t chain_id_base;       // Philox stream id of chain 0 of the batch
    int replicas;
    long long steps;
    uint64_t seed;
    const double *recorded;       // nullable
    const long long *rec_offsets; // global per-chain offsets into `recorded` (n_chains+1)
    long long rec_base;           // element index of recorded[0]
    double *out_xyzq;             // chain-major final poses of this launch
    long long out_base;           // atom index of out_xyzq[0] in the global chain-major layout
    void *out_stats;              // ddd_chain_stats[n_chains_local]
    uint8_t *out_trace;           // nullable: n_chains_local * steps
    int nl_cap;                   // even; shared-memory arrays are sized for nl_cap atoms
    // SM-persistent engine only (nullable):
    const double *ref_xyzq;       // reference poses for the fused CalculateRMSD (validateResults.go:50-65), indexed like lig_xyzq
    double *out_rmsd;             // [n_chains_local] RMSD(final pose, reference pose) written by the chain's epilogue
    double *out_best_xyzq;        // best-ever pose per chain, laid out like out_xyzq (EXTENSION: the reference returns the final pose)
    double *out_best_e;           // [n_chains_local] lowest chain energy seen (start pose included)
};

// mirrors ddd_chain_stats (include/ddd_mc.h)
struct ChainStats {
    long long steps, accepted, downhill, retries, draws, status;
    double final_energy, start_energy, chain_energy;
};

__host__ __device__ inline size_t chain_smem_bytes(int nl_cap, int threads) {
    // control words + counters + reduction scratch + per-lane work-item sums, then the packed fp32 pose and the
    // 7 fp64 arrays (cur xyz, prev xyz, q); see ChainSmem
    return 48 + (2 * kMaxWarps + max_items(threads) * 32) * sizeof(double) + (size_t)nl_cap * (7 * sizeof(double) + sizeof(float4));
}

// Views into the CTA's dynamic shared memory.  Layout (bytes): [ctl 16][cnt 32][red 2*kMaxWarps*8][item_out
// max_items(THREADS)*32*8][lf 16*cap][cx cy cz px py pz q: 7*8*cap].  Everything the pair loop touches sits at a
// compile-time constant offset; only `cap` is kept in a register (ten live 64-bit pointers would cost 20 of the 64
// registers the chain kernel is allowed).
template <int THREADS>
struct ChainSmem {
    int cap;
    static constexpr int kCtlOff = 0, kCntOff = 16, kRedOff = 48, kItemOff = kRedOff + 2 * kMaxWarps * 8,
                         kLfOff = kItemOff + max_items(THREADS) * 32 * 8;
    __device__ __forceinline__ ChainSmem(unsigned char *, int cap_) : cap(cap_) {}
    __device__ __forceinline__ unsigned char *raw() const {
        extern __shared__ __align__(16) unsigned char ddd_smem_raw[];
        return ddd_smem_raw;
    }
    __device__ __forceinline__ int *ctl() const { return reinterpret_cast<int *>(raw() + kCtlOff); }
    __device__ __forceinline__ long long *cnt() const { return reinterpret_cast<long long *>(raw() + kCntOff); }
    __device__ __forceinline__ double *red() const { return reinterpret_cast<double *>(raw() + kRedOff); }
    __device__ __forceinline__ double *item_out() const { return reinterpret_cast<double *>(raw() + kItemOff); }
    __device__ __forceinline__ float4 *lf() const { return reinterpret_cast<float4 *>(raw() + kLfOff); }
    __device__ __forceinline__ double *base() const { return reinterpret_cast<double *>(raw() + kLfOff) + 2 * cap; }
    __device__ __forceinline__ double *cx() const { return base(); }
    __device__ __forceinline__ double *cy() const { return base() + cap; }
    __device__ __forceinline__ double *cz() const { return base() + 2 * cap; }
    __device__ __forceinline__ double *px() const { return base() + 3 * cap; }
    __device__ __forceinline__ double *py() const { return base() + 4 * cap; }
    __device__ __forceinline__ double *pz() const { return base() + 5 * cap; }
    __device__ __forceinline__ double *q() const { return base() + 6 * cap; }
};

// IsCollisionFree (:229-242) over all i<j pairs, spread over the CTA; returns "this thread saw
// a pair closer than min_distance".  Pairs are enumerated as (i, i+m mod nl), m = 1..nl/2.
// sqrt(d2) < min_distance is decided from d2 alone unless d2 lies within 1e-9 (relative) of
// min_distance^2, where the reference's own sqrt comparison (:235-236) is evaluated.
template <int THREADS>
__device__ __forceinline__ bool collision_partial(const ChainSmem<THREADS> &s, int nl, double min_distance) {
    const int full_rows = (nl - 1) / 2;
    const int n_full = full_rows * nl;
    const int total = n_full + ((nl & 1) ? 0 : nl / 2);
    const double m2 = min_distance * min_distance, m2_lo = m2 * (1.0 - 1e-9), m2_hi = m2 * (1.0 + 1e-9);
    bool hit = false;
    for (int w = threadIdx.x; w < total; w += THREADS) {
        int i, m;
        if (w < n_full) { m = w / nl + 1; i = w - (m - 1) * nl; } else { i = w - n_full; m = nl / 2; }
        int j = i + m; if (j >= nl) j -= nl;
        const double dx = s.cx()[i] - s.cx()[j], dy = s.cy()[i] - s.cy()[j], dz = s.cz()[i] - s.cz()[j];
        const double d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < m2_lo) hit = true;
        else if (d2 < m2_hi) hit |= (sqrt(d2) < min_distance);              // :235-236
    }
    return hit;
}

// fp32 screen in front of IsCollisionFree, on the pair-packed fp32 pose (same FADD2/FMUL2/FFMA2 form as
// the energy loop): atom i against ligand pairs (2k, 2k+1), both orders of every pair, self excluded.
// Returns true when some pair lies within 2 % (+ the fp32 coordinate rounding bound) of min_distance^2,
// i.e. "fp32 cannot rule a collision out"; only then does the CTA run the exact fp64 test above.
struct CollisionPlan { int i0, i_step, k0, k_step; };      // fixed per chain: no division inside the step loop

template <int THREADS>
__device__ __forceinline__ CollisionPlan collision_plan(int nl) {
    CollisionPlan p;
    const int tid = threadIdx.x;
    if (nl <= 0) { p.i0 = 0; p.i_step = 1; p.k0 = 0; p.k_step = 1; return p; }
    const int groups = THREADS / nl;
    if (groups >= 1) {                       // every atom gets `groups` threads, each taking every groups-th pair
        p.i0 = (tid < groups * nl) ? tid % nl : nl; p.i_step = nl; p.k0 = tid / nl; p.k_step = groups;
    } else { p.i0 = tid; p.i_step = THREADS; p.k0 = 0; p.k_step = 1; }
    return p;
}

// returns bit 1: some pair is certainly closer than min_distance; bit 0: some pair is inside the uncertainty band
template <class SM>
__device__ __forceinline__ int collision_screen_f32(const SM &s, int nl, const CollisionPlan &cp, float m2) {
    const int n_pairs = (nl + 1) >> 1;
    const ulonglong2 *lf2 = reinterpret_cast<const ulonglong2 *>(s.lf());
    const float *lff = reinterpret_cast<const float *>(s.lf());
    float dmin = 3.0e38f, slack = 0.f;
    for (int i = cp.i0; i < nl; i += cp.i_step) {
        const float *mine = lff + 8 * (i >> 1) + (i & 1);
        const float xi = -mine[0], yi = -mine[2], zi = -mine[4];
        // |error of the fp32 d2| of a pair near min_distance <= ~4e-7 * (|x|+|y|+|z|); 1e-6 * (...) is a safe bound
        slack = fmaxf(slack, 1e-6f * (fabsf(xi) + fabsf(yi) + fabsf(zi) + 1.0f));
        const f32x2 ax = pack2(xi, xi), ay = pack2(yi, yi), az = pack2(zi, zi);
#pragma unroll 4
        for (int k = cp.k0; k < n_pairs; k += cp.k_step) {
            const ulonglong2 Lxy = lf2[2 * k], Lzq = lf2[2 * k + 1];
            const f32x2 dx = add2(ax, Lxy.x), dy = add2(ay, Lxy.y), dz = add2(az, Lzq.x);
            const f32x2 d2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
            float d2a, d2b;
            unpack2(d2, d2a, d2b);
            if (k == (i >> 1)) { if (i & 1) d2b = 3.0e38f; else d2a = 3.0e38f; }       // j == i
            dmin = fminf(dmin, fminf(d2a, d2b));
        }
    }
    const int hit = dmin < m2 * 0.98f - slack;
    const int near = dmin < m2 * 1.02f + slack;
    return (hit << 1) | (near & ~hit);
}

// RotateAtom (:213-224): Rodrigues about an axis through the origin, reference expression order
__device__ __forceinline__ void rotate_atom(double &x, double &y, double &z, double ax, double ay, double az,
                                            double cos_t, double sin_t) {
    const double dot = x * ax + y * ay + z * az;
    const double nx = x * cos_t + sin_t * (ay * z - az * y) + ax * dot * (1 - cos_t);
    const double ny = y * cos_t + sin_t * (az * x - ax * z) + ay * dot * (1 - cos_t);
    const double nz = z * cos_t + sin_t * (ax * y - ay * x) + az * dot * (1 - cos_t);
    x = nx; y = ny; z = nz;
}

// IEEE fp64 divide / square root as ONE out-of-line copy each: the inline expansions are ~30 instructions per use, and the
// serial side of the chain kernels is bound by instruction fetch.  Same operations, same bits.
__device__ __noinline__ double cold_div(double a, double b) { return a / b; }
__device__ __noinline__ double cold_sqrt(double a) { return sqrt(a); }

// axis draws X,Y,Z (U*2-1), Normalize (no-op at zero magnitude, datatypes.go:33-40), theta draw (:192-200)
__device__ __forceinline__ void draw_axis_angle(const Stream &st, uint64_t pos, double max_angle, int &status,
                                                double &ax, double &ay, double &az, double &theta) {
    double u[4];
    st.draw4(pos, u, status);
    ax = u[0] * 2 - 1;
    ay = u[1] * 2 - 1;
    az = u[2] * 2 - 1;
    const double mag = cold_sqrt(ax * ax + ay * ay + az * az);
    if (mag > 0) { ax = cold_div(ax, mag); ay = cold_div(ay, mag); az = cold_div(az, mag); }
    theta = (u[3] * 2 - 1) * max_angle;
}

template <int THREADS, bool PRECISE>
__device__ __forceinline__ double chain_energy(const ChainArgs &a, const ChainSmem<THREADS> &s, int nl, int &parity) {
    if (!PRECISE) return a.prm.k_coulomb * cta_pair_sum_f32<THREADS>(a.rec, s.lf(), nl, a.prm.rinv_max, s.ctl(), s.item_out(), s.red(), parity);
    const double part = pair_sum_f64<THREADS, false>(a.rec, s.cx(), s.cy(), s.cz(), s.q(), nl, a.prm.k_coulomb, a.prm.clamp_distance, nullptr);
    return block_sum<THREADS>(part, s.red(), parity);
}

template <class SM>
__device__ __forceinline__ void store_pose(const ChainArgs &a, const SM &s, int i, double x, double y, double z) {
    s.cx()[i] = x; s.cy()[i] = y; s.cz()[i] = z;
    store_lig_f32(s.lf(), i, (float)(x - a.rec.cx), (float)(y - a.rec.cy), (float)(z - a.rec.cz), (float)s.q()[i]);
}

// One CTA of THREADS threads = one chain.  THREADS is chosen by the host from the number of chains
// (few chains: wide CTAs for latency; many chains: narrow CTAs so the per-step proposal / collision /
// accept work, which every warp of the CTA repeats, is amortised over more pair terms per warp).
#ifndef DDD_CHAIN_WARPS_PER_SM
#define DDD_CHAIN_WARPS_PER_SM 32
#endif
template <int THREADS, bool PRECISE>
__global__ void __launch_bounds__(THREADS, (PRECISE ? 1 : DDD_CHAIN_WARPS_PER_SM * 32 / THREADS)) mc_chain_kernel(const ChainArgs a) {
    const ChainSmem<THREADS> s(nullptr, a.nl_cap);
    const int tid = threadIdx.x;
    const long long local = a.order ? (long long)a.order[blockIdx.x] : (long long)blockIdx.x;
    const long long chain = a.chain_begin + local;
    const long long lig = chain / a.replicas;
    const int rep = (int)(chain - lig * a.replicas);
    const long long off = a.offsets[lig];
    const int nl = (int)(a.offsets[lig + 1] - off);
    const DevParams &P = a.prm;
    const bool warp_has_atoms = (tid & ~31) < nl;

    Stream st;
    st.seed = a.seed; st.id = a.chain_id_base + (uint64_t)chain; st.rec = nullptr; st.rec_len = 0;
    if (a.recorded) {
        st.rec = a.recorded + (a.rec_offsets[chain] - a.rec_base);
        st.rec_len = a.rec_offsets[chain + 1] - a.rec_offsets[chain];
    }

    // private copy of the start pose (each chain owns its atoms)
    const double *src = a.lig_xyzq + (off - a.atom_base) * 4;
    for (int i = tid; i < nl; i += THREADS) {
        s.q()[i] = src[4 * i + 3];
        store_pose(a, s, i, src[4 * i], src[4 * i + 1], src[4 * i + 2]);
    }
    if (tid == 0) {
        if (nl & 1) store_lig_f32(s.lf(), nl, -kPadCoord, -kPadCoord, -kPadCoord, 0.f);   // q = 0 pad atom
        s.cnt()[0] = 0; s.cnt()[1] = 0; s.cnt()[2] = 0; s.cnt()[3] = 0;
        s.ctl()[0] = 0; s.ctl()[1] = 0; s.ctl()[2] = 0;
    }
    __syncthreads();

    int parity = 0, status = 0;
    const double start_energy = chain_energy<THREADS, true>(a, s, nl, parity);       // :118 in fp64
    double cur_e = PRECISE ? start_energy : chain_energy<THREADS, false>(a, s, nl, parity);

    uint64_t pos = 0;
    const int draws_per_attempt = (P.move_mode == 1) ? 7 : (P.rotate ? 4 : 0) + 3 * nl;
    const CollisionPlan cplan = collision_plan<THREADS>(nl);
    const float m2_f32 = (float)(P.min_distance * P.min_distance);

    for (long long step = 0; step < a.steps; ++step) {
        for (int i = tid; i < nl; i += THREADS) { s.px()[i] = s.cx()[i]; s.py()[i] = s.cy()[i]; s.pz()[i] = s.cz()[i]; }   // :121
        if (P.move_mode == 1) __syncthreads();     // RIGID reads every atom of prev for the centroid

        long long rej = 0;
        bool bail = false;
        for (;;) {
            if (P.move_mode == 1) {
                // EXTENSION: rigid-body move, draws tx,ty,tz,ax,ay,az,theta (no reference row)
                if (warp_has_atoms) {
                    const double tx = (st.draw(pos, status) - 0.5) * P.rigid_translation;
                    const double ty = (st.draw(pos + 1, status) - 0.5) * P.rigid_translation;
                    const double tz = (st.draw(pos + 2, status) - 0.5) * P.rigid_translation;
                    double ax, ay, az, theta;
                    draw_axis_angle(st, pos + 3, P.rigid_angle, status, ax, ay, az, theta);
                    double gx = 0, gy = 0, gz = 0;
                    for (int i = 0; i < nl; ++i) { gx += s.px()[i]; gy += s.py()[i]; gz += s.pz()[i]; }
                    if (nl > 0) { gx /= (double)nl; gy /= (double)nl; gz /= (double)nl; }
                    const double h = 0.5 * theta;
                    double sh, w; sincos(h, &sh, &w);
                    const double qx = sh * ax, qy = sh * ay, qz = sh * az;
                    for (int i = tid; i < nl; i += THREADS) {
                        const double vx = s.px()[i] - gx, vy = s.py()[i] - gy, vz = s.pz()[i] - gz;
                        const double c1x = qy * vz - qz * vy, c1y = qz * vx - qx * vz, c1z = qx * vy - qy * vx;
                        const double c2x = qy * c1z - qz * c1y, c2y = qz * c1x - qx * c1z, c2z = qx * c1y - qy * c1x;
                        store_pose(a, s, i, gx + (vx + 2.0 * (w * c1x + c2x)) + tx,
                                   gy + (vy + 2.0 * (w * c1y + c2y)) + ty, gz + (vz + 2.0 * (w * c1z + c2z)) + tz);
                    }
                }
                pos += 7;
                if (tid == 0) s.ctl()[0] = 0;
                __syncthreads();
                break;                          // rigid: intra-ligand distances unchanged, no collision test
            }
            // REFERENCE move: JitterAndRotateLigand (:172-184) / JitterLigand (:154-167), in place
            if (warp_has_atoms) {               // warps that own no atom skip the draws and the sincos
                double ax = 0, ay = 0, az = 0, cos_t = 1, sin_t = 0;
                uint64_t jpos = pos;
                if (P.rotate) {                                                   // RotateLigand :189-208
                    double theta;
                    draw_axis_angle(st, pos, P.max_angle, status, ax, ay, az, theta);
                    sincos(theta, &sin_t, &cos_t);                               // :214-215
                    jpos += 4;
                }
                for (int i = tid; i < nl; i += THREADS) {
                    double x = s.cx()[i], y = s.cy()[i], z = s.cz()[i];
                    if (P.rotate) rotate_atom(x, y, z, ax, ay, az, cos_t, sin_t);
                    double u[3];
                    st.draw3(jpos + 3ull * (uint64_t)i, u, status);
                    x += (u[0] - 0.5) * P.jitter_width;                          // :176-178 / :158-160
                    y += (u[1] - 0.5) * P.jitter_width;
                    z += (u[2] - 0.5) * P.jitter_width;
                    store_pose(a, s, i, x, y, z);
                }
            }
            pos += (uint64_t)draws_per_attempt;
            if (tid == 0) s.ctl()[0] = 0;
            __syncthreads();
            // IsCollisionFree (:180 / :162).  Retries are frequent late in a chain (the per-atom jitter lets atoms
            // pile up until only MINDISTANCE separates them), so the test must be cheap: an fp32 screen decides
            // "certainly colliding" / "certainly free"; only a pair inside the 2 % band goes to the exact fp64 test.
            const int code = collision_screen_f32(s, nl, cplan, m2_f32);
            bool collided = __syncthreads_or(code & 2) != 0;
            if (!collided && __syncthreads_or(code & 1)) {
                const bool hit = collision_partial<THREADS>(s, nl, P.min_distance);
                collided = __syncthreads_or(hit) != 0;
            }
            if (!collided) break;
            ++rej;                                  // retry on the already-moved atoms
            if (rej >= P.max_retries) { bail = true; break; }
        }
        if (tid == 0) s.cnt()[2] += rej;
        if (bail) {                                  // the reference would spin forever here
            status |= 1;
            for (int i = tid; i < nl; i += THREADS) { s.cx()[i] = s.px()[i]; s.cy()[i] = s.py()[i]; s.cz()[i] = s.pz()[i]; }
            break;
        }

        const double new_e = chain_energy<THREADS, PRECISE>(a, s, nl, parity);    // :127
        bool acc;                                                             // AcceptMove :141-149
        const bool dh = new_e < cur_e;
        if (dh) acc = true;
        else {
            const double delta = new_e - cur_e;
            const double x = -delta / P.temperature;
            // exp(x) is exactly +0 below -746, and no uniform is < 0: reject without evaluating either.  The draw
            // is still consumed (:148).  (With K = 9e9 this is nearly every uphill move)
            if (x < -746.0 && !st.rec) acc = false;
            else acc = st.draw(pos, status) < exp(x);
            pos += 1;
        }
        if (acc) cur_e = new_e;                                               // :129-130
        else for (int i = tid; i < nl; i += THREADS) { s.cx()[i] = s.px()[i]; s.cy()[i] = s.py()[i]; s.cz()[i] = s.pz()[i]; }  // :132
        if (tid == 0) {
            s.cnt()[0] += acc ? 1 : 0; s.cnt()[1] += dh ? 1 : 0; s.cnt()[3] += 1;
            if (a.out_trace) a.out_trace[local * a.steps + step] = acc ? 1 : 0;
        }
    }
    __syncthreads();

    // CalculateEnergy of the returned pose in fp64 (:61 / :104)
    const double final_energy = PRECISE ? cur_e : chain_energy<THREADS, true>(a, s, nl, parity);

    const long long out_atom = (long long)a.replicas * off + (long long)rep * nl - a.out_base;
    double *dst = a.out_xyzq + out_atom * 4;
    for (int i = tid; i < nl; i += THREADS) {
        dst[4 * i] = s.cx()[i]; dst[4 * i + 1] = s.cy()[i]; dst[4 * i + 2] = s.cz()[i]; dst[4 * i + 3] = s.q()[i];
    }
    // status bit 2 (stream exhausted) can be raised by any thread that drew: OR it over the CTA
    status = __syncthreads_or(status & 2) ? (status | 2) : status;
    if (tid == 0) {
        ChainStats cs;
        cs.accepted = s.cnt()[0]; cs.downhill = s.cnt()[1]; cs.retries = s.cnt()[2]; cs.steps = s.cnt()[3];
        cs.draws = (long long)pos; cs.status = status;
        cs.final_energy = final_energy; cs.start_energy = start_energy; cs.chain_energy = cur_e;
        reinterpret_cast<ChainStats *>(a.out_stats)[local] = cs;
    }
}

// ------------------------------------------------------------------ CalculateEnergy batch (:247-262)
struct EnergyArgs {
    RecView rec;
    DevParams prm;
    const long long *offsets;   // CSR of this launch's ligands (n+1 entries of the caller's array)
    long long atom_base;        // atom index of lig_xyzq[0] (a device may hold only a slice of the batch)
    const double *lig_xyzq;
    double *out_energy;
    double *out_abs;            // nullable
    int nl_cap;
};

template <bool PRECISE, bool WITH_ABS>
__global__ void __launch_bounds__(kThreads) energy_batch_kernel(const EnergyArgs a) {
    const ChainSmem<kThreads> s(nullptr, a.nl_cap);
    const int tid = threadIdx.x;
    const long long lig = blockIdx.x;
    const long long off = a.offsets[lig];
    const int nl = (int)(a.offsets[lig + 1] - off);
    const double *src = a.lig_xyzq + (off - a.atom_base) * 4;
    for (int i = tid; i < nl; i += kThreads) {
        const double x = src[4 * i], y = src[4 * i + 1], z = src[4 * i + 2], q = src[4 * i + 3];
        s.cx()[i] = x; s.cy()[i] = y; s.cz()[i] = z; s.q()[i] = q;
        store_lig_f32(s.lf(), i, (float)(x - a.rec.cx), (float)(y - a.rec.cy), (float)(z - a.rec.cz), (float)q);
    }
    if (tid == 0) {
        if (nl & 1) store_lig_f32(s.lf(), nl, -kPadCoord, -kPadCoord, -kPadCoord, 0.f);
        s.ctl()[0] = 0;
    }
    __syncthreads();
    int parity = 0;
    double part = 0.0, ab = 0.0, e;
    if (PRECISE || WITH_ABS) {
        part = pair_sum_f64<kThreads, WITH_ABS>(a.rec, s.cx(), s.cy(), s.cz(), s.q(), nl, a.prm.k_coulomb, a.prm.clamp_distance, &ab, a.prm.cutoff);
    }
    if (PRECISE) e = block_sum<kThreads>(part, s.red(), parity);
    else if (a.prm.cutoff > 0) {                        // every unit, no cell list: this entry point is not the hot path
        const float rc_q = a.prm.rc_inv * ligand_charge_sum_f32(s.lf(), nl);
        e = a.prm.k_coulomb * cta_pair_sum_f32<kThreads, true>(a.rec, s.lf(), nl, a.prm.rinv_max, s.ctl(), s.item_out(), s.red(), parity, a.prm.rc_inv, rc_q);
    }
    else e = a.prm.k_coulomb * cta_pair_sum_f32<kThreads>(a.rec, s.lf(), nl, a.prm.rinv_max, s.ctl(), s.item_out(), s.red(), parity);
    if (WITH_ABS) ab = block_sum<kThreads>(ab, s.red(), parity);
    if (tid == 0) {
        a.out_energy[lig] = e;
        if (WITH_ABS) a.out_abs[lig] = ab;
    }
}

// ------------------------------------------------------------------ FindClosestAtomDistance (:290-304) + shift (:267-285)
struct ClosestArgs {
    RecView rec;
    const long long *offsets;   // points at the first ligand of this launch
    long long atom_base;        // atom index of lig_xyzq[0]
    double *lig_xyzq;           // in/out when do_shift
    double *out_dist;           // nullable
    long long *out_lig_atom, *out_prot_atom;   // nullable
    double threshold;
    int do_shift;
};

// first minimum in ligand-outer / protein-inner order: smallest (distance, i*np + j)
__device__ __forceinline__ void argmin_merge(double &d, long long &k, double od, long long ok) {
    if (od < d || (od == d && ok < k)) { d = od; k = ok; }
}

__global__ void __launch_bounds__(kThreads) closest_atom_kernel(const ClosestArgs a) {
    __shared__ double sd[kThreads / 32];
    __shared__ long long sk[kThreads / 32];
    __shared__ double shift[3];
    const int tid = threadIdx.x;
    const long long lig = blockIdx.x;
    const long long off = a.offsets[lig];
    const int nl = (int)(a.offsets[lig + 1] - off);
    double *atoms = a.lig_xyzq + (off - a.atom_base) * 4;
    const int np = a.rec.n;
    double best = 1.7976931348623157e308;        // math.MaxFloat64 (:291)
    long long key = 0x7fffffffffffffffll;
    for (int j = tid; j < np; j += kThreads) {
        const Atom64 Pj = a.rec.f64[j];
        for (int i = 0; i < nl; ++i) {
            const double dx = atoms[4 * i] - Pj.x, dy = atoms[4 * i + 1] - Pj.y, dz = atoms[4 * i + 2] - Pj.z;
            const double dist = sqrt(dx * dx + dy * dy + dz * dz);          // Distance :309-314
            argmin_merge(best, key, dist, (long long)i * np + j);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, best, o);
        const long long ok = __shfl_xor_sync(0xffffffffu, key, o);
        argmin_merge(best, key, od, ok);
    }
    if ((tid & 31) == 0) { sd[tid >> 5] = best; sk[tid >> 5] = key; }
    __syncthreads();
    best = sd[0]; key = sk[0];
    for (int w = 1; w < kThreads / 32; ++w) argmin_merge(best, key, sd[w], sk[w]);
    const bool found = key != 0x7fffffffffffffffll;
    const long long li = found ? key / np : -1, pj = found ? key % np : -1;
    if (tid == 0) {
        if (a.out_dist) a.out_dist[lig] = best;
        if (a.out_lig_atom) a.out_lig_atom[lig] = li;
        if (a.out_prot_atom) a.out_prot_atom[lig] = pj;
    }
    if (!a.do_shift) return;
    if (tid == 0) {
        double vx = 0, vy = 0, vz = 0;
        if (best > a.threshold) {                                            // :269
            if (found) {
                const Atom64 Pp = a.rec.f64[pj];
                vx = Pp.x - atoms[4 * li]; vy = Pp.y - atoms[4 * li + 1]; vz = Pp.z - atoms[4 * li + 2];   // :271-275
            }
            const double mag = sqrt(vx * vx + vy * vy + vz * vz);           // Normalize :277
            if (mag > 0) { vx /= mag; vy /= mag; vz /= mag; }
            const double sc = best - a.threshold;                             // :278
            vx *= sc; vy *= sc; vz *= sc;
        }
        shift[0] = vx; shift[1] = vy; shift[2] = vz;
    }
    __syncthreads();
    if (best > a.threshold)
        for (int i = tid; i < nl; i += kThreads) {                           // :280-282
            atoms[4 * i] += shift[0]; atoms[4 * i + 1] += shift[1]; atoms[4 * i + 2] += shift[2];
        }
}

// ------------------------------------------------------------------ CalculateRMSD batch (validateResults.go:50-65)
// sum_i |sim_i - ref_i|^2 over one warp's lane-strided atoms (32-byte Go atoms).  The loads of the first 96 atoms are issued
// before the first is used (a warp has only 1-3 trips for 20..80-atom ligands, so memory parallelism has to come from here);
// the additions stay in atom order, so the result does not depend on the batching.
__device__ __forceinline__ double warp_rmsd_partial(const double4 *__restrict__ s4, const double4 *__restrict__ r4, int n) {
    const int lane = threadIdx.x & 31;
    double4 sa[3], ra[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int i = lane + 32 * t;
        if (i < n) { sa[t] = s4[i]; ra[t] = r4[i]; }
    }
    double acc = 0.0;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        if (lane + 32 * t < n) {
            const double dx = sa[t].x - ra[t].x, dy = sa[t].y - ra[t].y, dz = sa[t].z - ra[t].z;
            acc += dx * dx + dy * dy + dz * dz;                               // :58-62
        }
    }
    for (int i = lane + 96; i < n; i += 32) {
        const double4 a = s4[i], b = r4[i];
        const double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
        acc += dx * dx + dy * dy + dz * dz;
    }
    return acc;
}

// One warp per (simulated, reference) pair at a time; lane-strided atoms, coalesced; HBM-bound.  PERSISTENT: the grid is sized to
// the resident warps of the device and every warp walks the pairs with the grid's warp count as its stride, fetching the NEXT
// pair's offsets before it sums the current one -- a 20..80-atom pair is one or two memory round trips, so the dependent
// offsets -> atoms latency and the CTA turnover of a one-pair-per-warp launch cost as much as the traffic itself.
__global__ void __launch_bounds__(256) rmsd_batch_kernel(const long long *__restrict__ offsets, long long atom_base,
                                                         const double *__restrict__ sim,
                                                         const double *__restrict__ ref, int stride,
                                                         long long n_pairs, double *__restrict__ out) {
    const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
    long long pair = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pair >= n_pairs) return;
    const int lane = threadIdx.x & 31;
    long long o0 = offsets[pair], o1 = offsets[pair + 1];
    for (;;) {
        const long long next = pair + n_warps;
        long long n0 = 0, n1 = 0;
        if (next < n_pairs) { n0 = offsets[next]; n1 = offsets[next + 1]; }
        const long long off = o0 - atom_base;
        const int n = (int)(o1 - o0);
        double acc = 0.0;
        if (stride == 4) {
            acc = warp_rmsd_partial(reinterpret_cast<const double4 *>(sim) + off, reinterpret_cast<const double4 *>(ref) + off, n);
        } else {
            const double *sp = sim + off * 3, *rp = ref + off * 3;
            for (int i = lane; i < n; i += 32) {
                const double dx = sp[3 * i] - rp[3 * i], dy = sp[3 * i + 1] - rp[3 * i + 1], dz = sp[3 * i + 2] - rp[3 * i + 2];
                acc += dx * dx + dy * dy + dz * dz;
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) out[pair] = sqrt(acc / (double)n);                     // :64 (n == 0 -> NaN)
        if (next >= n_pairs) break;
        pair = next; o0 = n0; o1 = n1;
    }
}

// CalculateRMSD (validateResults.go:50-65) of every chain of a resident screen: final pose (chain-major output) against
// the ligand's start pose, without leaving the device.  One warp per chain at a time (persistent, like rmsd_batch_kernel),
// 32-byte atoms (Go layout) read as double4.
__global__ void __launch_bounds__(256) rmsd_chains_kernel(const long long *__restrict__ offsets, const double *__restrict__ start_xyzq,
                                                          long long atom_base, const double *__restrict__ final_xyzq, long long out_base,
                                                          long long chain_begin, long long n_local, int replicas, double *__restrict__ out) {
    const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
    long long local = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (local >= n_local) return;
    const int lane = threadIdx.x & 31;
    long long lig = (chain_begin + local) / replicas;
    long long o0 = offsets[lig], o1 = offsets[lig + 1];
    for (;;) {
        const long long next = local + n_warps;
        long long nlig = 0, n0 = 0, n1 = 0;
        if (next < n_local) { nlig = (chain_begin + next) / replicas; n0 = offsets[nlig]; n1 = offsets[nlig + 1]; }
        const int rep = (int)(chain_begin + local - lig * replicas);
        const int n = (int)(o1 - o0);
        const double4 *r4 = reinterpret_cast<const double4 *>(start_xyzq) + (o0 - atom_base);
        const double4 *s4 = reinterpret_cast<const double4 *>(final_xyzq) + ((long long)replicas * o0 + (long long)rep * n - out_base);
        double acc = warp_rmsd_partial(s4, r4, n);
        acc = warp_sum(acc);
        if (lane == 0) out[local] = sqrt(acc / (double)n);                // :64 (n == 0 -> NaN)
        if (next >= n_local) break;
        local = next; lig = nlig; o0 = n0; o1 = n1;
    }
}

// ------------------------------------------------------------------ replica pick (metropolis.go:102-109) on device
// One thread per ligand: currentEnergy = E(start) (:96); for every replica in replica-index order (the reference's order is
// channel arrival) newEnergy = E(result) (:104), AcceptMove (:105, :141-149) with the ligand's parent stream.
// out_picked = the replica kept (-1: the start pose), out_energy = its energy = CalculateEnergy(minLigand) (:61).
__global__ void __launch_bounds__(128) replica_pick_kernel(const ChainStats *__restrict__ stats, long long n_lig_local, long long lig0,
                                                           long long chain_begin, int replicas, double temperature, uint64_t seed,
                                                           uint64_t parent_base, double *__restrict__ out_energy, long long *__restrict__ out_picked) {
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lig_local) return;
    const ChainStats *cs = stats + ((lig0 + l) * replicas - chain_begin);
    double cur_e = cs[0].start_energy;                                      // :96
    long long picked = -1;
    uint64_t pos = 0;
    const uint64_t parent = parent_base + (uint64_t)(lig0 + l);
    for (int r = 0; r < replicas; ++r) {
        const double new_e = cs[r].final_energy;                            // :104
        bool acc = new_e < cur_e;                                           // :142
        if (!acc) {
            const double probability = exp(-(new_e - cur_e) / temperature); // :147
            acc = philox_u01(seed, parent, pos++) < probability;            // :148
        }
        if (acc) { cur_e = new_e; picked = r; }                             // :106-107
    }
    out_energy[l] = cur_e;
    out_picked[l] = picked;
}

// One warp per ligand: copy the picked replica's final pose (or the start pose when no replica was kept) into the
// per-ligand CSR output, and CalculateRMSD (validateResults.go:50-65) of that pose against the reference pose.
__global__ void __launch_bounds__(256) pick_gather_kernel(const long long *__restrict__ offsets, const double *__restrict__ start_xyzq,
                                                          const double *__restrict__ ref_xyzq, long long atom_base,
                                                          const double *__restrict__ chain_xyzq, long long out_base, int replicas,
                                                          const long long *__restrict__ picked, long long n_lig_local, long long lig0,
                                                          double *__restrict__ out_xyzq, double *__restrict__ out_rmsd) {
    const long long l = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (l >= n_lig_local) return;
    const int lane = threadIdx.x & 31;
    const long long lig = lig0 + l;
    const long long off = offsets[lig];
    const int n = (int)(offsets[lig + 1] - off);
    const long long pk = picked[l];
    const double4 *src = pk >= 0 ? reinterpret_cast<const double4 *>(chain_xyzq) + ((long long)replicas * off + pk * n - out_base)
                                 : reinterpret_cast<const double4 *>(start_xyzq) + (off - atom_base);
    const double4 *r4 = reinterpret_cast<const double4 *>(ref_xyzq) + (off - atom_base);
    double4 *dst = reinterpret_cast<double4 *>(out_xyzq) + (off - atom_base);
    double acc = 0.0;
    for (int i = lane; i < n; i += 32) {
        const double4 v = src[i], ra = r4[i];
        dst[i] = v;
        const double dx = v.x - ra.x, dy = v.y - ra.y, dz = v.z - ra.z;
        acc += dx * dx + dy * dy + dz * dz;
    }
    acc = warp_sum(acc);
    if (lane == 0 && out_rmsd) out_rmsd[l] = sqrt(acc / (double)n);
}

// ------------------------------------------------------------------ SaveMinimumEnergyLigand's argmin (plotting.go:109-121)
// The reference scans the energies in order with `if energy < minEnergy` from (minIndex 0, minEnergy init): the result is the
// FIRST index that attains the global minimum if that minimum is below `init`, else (0, init).  One CTA: (value, index) pairs
// reduced with "smaller value, then smaller index"; NaN never satisfies `<` and is skipped like in Go.
__global__ void __launch_bounds__(1024) argmin_first_kernel(const double *__restrict__ e, long long n, long long index_base,
                                                            double *__restrict__ out_val, long long *__restrict__ out_idx) {
    __shared__ double sv[32];
    __shared__ long long si[32];
    double best = 1.7976931348623157e308 * 2.0;                             // +inf
    long long bi = 0x7fffffffffffffffll;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = e[i];
        if (v < best) { best = v; bi = i; }                                 // ascending i per thread: first minimum kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov < best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (sv[w] < best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
        *out_val = best;
        *out_idx = bi == 0x7fffffffffffffffll ? -1 : bi + index_base;
    }
}

// ------------------------------------------------------------------ RandomizeLigandPose (validateResults.go:70-79)
// `offsets` points at the first ligand of this launch, `lig_xyzq` at atom `atom_base` (a device may hold only a slice)
__global__ void __launch_bounds__(kThreads) randomize_pose_kernel(const long long *offsets, double *lig_xyzq, long long atom_base,
                                                                  uint64_t seed, uint64_t first_index) {
    const long long lig = blockIdx.x;
    const long long off = offsets[lig];
    const int nl = (int)(offsets[lig + 1] - off);
    double *atoms = lig_xyzq + (off - atom_base) * 4;
    Stream st; st.rec = nullptr; st.rec_len = 0; st.seed = seed;
    st.id = 0x2000000000000000ull + first_index + (uint64_t)lig;
    int status = 0;
    double ax, ay, az, theta, sin_t, cos_t;
    draw_axis_angle(st, 3ull * (uint64_t)nl, 3.14159265358979323846, status, ax, ay, az, theta);   // :78
    sincos(theta, &sin_t, &cos_t);
    for (int i = threadIdx.x; i < nl; i += kThreads) {
        double u[3];
        st.draw3(3ull * (uint64_t)i, u, status);
        double x = atoms[4 * i] + (u[0] - 0.5) * 10.0;                         // :73-75
        double y = atoms[4 * i + 1] + (u[1] - 0.5) * 10.0;
        double z = atoms[4 * i + 2] + (u[2] - 0.5) * 10.0;
        rotate_atom(x, y, z, ax, ay, az, cos_t, sin_t);
        atoms[4 * i] = x; atoms[4 * i + 1] = y; atoms[4 * i + 2] = z;
    }
}

}  // namespace ddd
