// ddd_kernels.cuh -- sm_100a device code of the Metropolis Monte-Carlo hot path.
//
// One CTA (kThreads = 128) runs one independent ligand chain of
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

namespace ddd {

constexpr int kThreads = 128;   // threads per chain CTA
constexpr int kTileA   = 4;     // receptor atoms held in registers per lane per tile (fp32 path)
constexpr int kTile    = kThreads * kTileA;
constexpr float kPadCoord = 1.0e4f;   // padding receptor atoms: q = 0, parked 1e4 A away

// receptor atom, fp64 master copy: absolute coordinates, Go Atom layout (datatypes.go:18-25)
struct __align__(32) Atom64 { double x, y, z, q; };

struct DevParams {
    double k_coulomb, min_distance, max_angle, jitter_width, clamp_distance;
    double rigid_translation, rigid_angle, temperature;
    long long max_retries;
    float rinv_max;              // 1 / clamp_distance: clamp d >= 1e-6  <=>  1/d <= 1e6 (:255-257)
    int move_mode, rotate;
};

struct RecView {
    const float4 *f32;           // (x-cx, y-cy, z-cz, q), padded to a multiple of kThreads
    const Atom64 *f64;           // absolute, n atoms
    int n, n_pad;
    double cx, cy, cz;           // receptor centroid: origin of the fp32 frame
};

// ------------------------------------------------------------------ Philox4x32-10
// Salmon et al. SC'11 (Random123).  Same counter/key convention as the host side:
// counter = (index/2 lo, index/2 hi, stream lo, stream hi), key = (seed lo, seed hi);
// words (0,1) -> draw 2k, words (2,3) -> draw 2k+1, 53-bit mantissa, [0,1).
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
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
};

// ------------------------------------------------------------------ reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block sum returned to every thread.  `red` holds 2 x (kThreads/32) doubles used
// alternately (parity) so that back-to-back calls need only the one barrier inside.
__device__ __forceinline__ double block_sum(double v, double *red, int &parity) {
    v = warp_sum(v);
    double *buf = red + parity * (kThreads / 32);
    parity ^= 1;
    if ((threadIdx.x & 31) == 0) buf[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = buf[0];
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w) s += buf[w];
    return s;
}

__device__ __forceinline__ float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ------------------------------------------------------------------ pair sums (CalculateEnergy :247-262)
// fp32 path.  Lane holds A receptor atoms in registers, ligand atoms are broadcast from shared
// memory as float4 (x, y, z, q) in the receptor-centred frame.  Per pair: 3 FADD, FMUL + 2 FFMA
// (d^2), MUFU.RSQ, FMNMX (clamp), FFMA (s += q_l * 1/d).  q_p and K are applied per tile / per call.
template <int A>
__device__ __forceinline__ double tile_sum_f32(const float4 *__restrict__ rec, int base,
                                               const float4 *__restrict__ lf, int nl, float rinv_max) {
    float4 P[A];
    float s[A];
#pragma unroll
    for (int a = 0; a < A; ++a) { P[a] = __ldg(rec + base + a * kThreads + threadIdx.x); s[a] = 0.f; }
#pragma unroll 2
    for (int l = 0; l < nl; ++l) {
        const float4 L = lf[l];
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const float dx = P[a].x - L.x, dy = P[a].y - L.y, dz = P[a].z - L.z;
            const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const float r = fminf(rsqrt_approx(d2), rinv_max);
            s[a] = fmaf(L.w, r, s[a]);
        }
    }
    double e = 0.0;
#pragma unroll
    for (int a = 0; a < A; ++a) e = fma((double)P[a].w, (double)s[a], e);
    return e;
}

// returns this thread's share of sum_p sum_l q_p q_l / max(d, clamp)   (K applied by the caller)
__device__ __forceinline__ double pair_sum_f32(const RecView &rv, const float4 *lf, int nl, float rinv_max) {
    double e = 0.0;
    int base = 0;
    for (; base + kTile <= rv.n_pad; base += kTile) e += tile_sum_f32<kTileA>(rv.f32, base, lf, nl, rinv_max);
    for (; base < rv.n_pad; base += kThreads) e += tile_sum_f32<1>(rv.f32, base, lf, nl, rinv_max);
    return e;
}

// fp64 path: the reference's per-term arithmetic, term by term (no contraction: -fmad=false).
template <bool WITH_ABS>
__device__ __forceinline__ double pair_sum_f64(const RecView &rv, const double *lx, const double *ly,
                                               const double *lz, const double *lq, int nl,
                                               double K, double clampd, double *abs_out) {
    double e = 0.0, ab = 0.0;
    for (int p = threadIdx.x; p < rv.n; p += kThreads) {
        const Atom64 P = rv.f64[p];
        for (int l = 0; l < nl; ++l) {
            const double ddx = P.x - lx[l], ddy = P.y - ly[l], ddz = P.z - lz[l];
            double d = sqrt(ddx * ddx + ddy * ddy + ddz * ddz);          // :253 (Pow(x,2) == x*x)
            if (d < clampd) d = clampd;                                   // :255-257
            const double term = K * (P.q * lq[l]) / d;                    // :258
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
};

// mirrors ddd_chain_stats (include/ddd_mc.h)
struct ChainStats {
    long long steps, accepted, downhill, retries, draws, status;
    double final_energy, start_energy, chain_energy;
};

__host__ __device__ inline size_t chain_smem_bytes(int nl_cap) {
    // cur xyz, prev xyz, q : 7 double arrays; float4 pose; 2 x 4 reduction doubles
    return (size_t)nl_cap * (7 * sizeof(double) + sizeof(float4)) + 2 * (kThreads / 32) * sizeof(double);
}

struct ChainSmem {
    double *cx, *cy, *cz, *px, *py, *pz, *q;
    float4 *lf;
    double *red;
    __device__ __forceinline__ ChainSmem(unsigned char *base, int cap) {
        cx = reinterpret_cast<double *>(base);
        cy = cx + cap; cz = cy + cap; px = cz + cap; py = px + cap; pz = py + cap; q = pz + cap;
        lf = reinterpret_cast<float4 *>(q + cap);
        red = reinterpret_cast<double *>(lf + cap);
    }
};

// IsCollisionFree (:229-242) over all i<j pairs, spread over the CTA; returns "this thread saw
// a pair closer than min_distance".  Pairs are enumerated as (i, i+m mod nl), m = 1..nl/2.
__device__ __forceinline__ bool collision_partial(const ChainSmem &s, int nl, double min_distance) {
    const int full_rows = (nl - 1) / 2;
    const int n_full = full_rows * nl;
    const int total = n_full + ((nl & 1) ? 0 : nl / 2);
    bool hit = false;
    for (int w = threadIdx.x; w < total; w += kThreads) {
        int i, m;
        if (w < n_full) { m = w / nl + 1; i = w - (m - 1) * nl; } else { i = w - n_full; m = nl / 2; }
        int j = i + m; if (j >= nl) j -= nl;
        const double dx = s.cx[i] - s.cx[j], dy = s.cy[i] - s.cy[j], dz = s.cz[i] - s.cz[j];
        const double distance = sqrt(dx * dx + dy * dy + dz * dz);       // :235
        hit |= (distance < min_distance);                                 // :236
    }
    return hit;
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

// axis draws X,Y,Z (U*2-1), Normalize (no-op at zero magnitude, datatypes.go:33-40), theta draw (:192-200)
__device__ __forceinline__ void draw_axis_angle(const Stream &st, uint64_t pos, double max_angle, int &status,
                                                double &ax, double &ay, double &az, double &theta) {
    ax = st.draw(pos, status) * 2 - 1;
    ay = st.draw(pos + 1, status) * 2 - 1;
    az = st.draw(pos + 2, status) * 2 - 1;
    const double mag = sqrt(ax * ax + ay * ay + az * az);
    if (mag > 0) { ax /= mag; ay /= mag; az /= mag; }
    theta = (st.draw(pos + 3, status) * 2 - 1) * max_angle;
}

template <bool PRECISE>
__device__ __forceinline__ double chain_energy(const ChainArgs &a, const ChainSmem &s, int nl, int &parity) {
    double part;
    if (PRECISE) part = pair_sum_f64<false>(a.rec, s.cx, s.cy, s.cz, s.q, nl, a.prm.k_coulomb, a.prm.clamp_distance, nullptr);
    else part = pair_sum_f32(a.rec, s.lf, nl, a.prm.rinv_max);
    double e = block_sum(part, s.red, parity);
    if (!PRECISE) e *= a.prm.k_coulomb;
    return e;
}

template <bool PRECISE>
__global__ void __launch_bounds__(kThreads) mc_chain_kernel(const ChainArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ChainSmem s(smem_raw, a.nl_cap);
    const int tid = threadIdx.x;
    const long long local = a.order ? (long long)a.order[blockIdx.x] : (long long)blockIdx.x;
    const long long chain = a.chain_begin + local;
    const long long lig = chain / a.replicas;
    const int rep = (int)(chain - lig * a.replicas);
    const long long off = a.offsets[lig];
    const int nl = (int)(a.offsets[lig + 1] - off);
    const DevParams &P = a.prm;

    Stream st;
    st.seed = a.seed; st.id = a.chain_id_base + (uint64_t)chain; st.rec = nullptr; st.rec_len = 0;
    if (a.recorded) {
        st.rec = a.recorded + (a.rec_offsets[chain] - a.rec_base);
        st.rec_len = a.rec_offsets[chain + 1] - a.rec_offsets[chain];
    }

    // private copy of the start pose (each chain owns its atoms)
    const double *src = a.lig_xyzq + (off - a.atom_base) * 4;
    for (int i = tid; i < nl; i += kThreads) {
        const double x = src[4 * i], y = src[4 * i + 1], z = src[4 * i + 2], q = src[4 * i + 3];
        s.cx[i] = x; s.cy[i] = y; s.cz[i] = z; s.q[i] = q;
        s.lf[i] = make_float4((float)(x - a.rec.cx), (float)(y - a.rec.cy), (float)(z - a.rec.cz), (float)q);
    }
    __syncthreads();

    int parity = 0, status = 0;
    const double start_energy = chain_energy<true>(a, s, nl, parity);            // :118 in fp64
    double cur_e = PRECISE ? start_energy : chain_energy<false>(a, s, nl, parity);

    uint64_t pos = 0;
    long long accepted = 0, downhill = 0, retries = 0, steps_done = 0;
    const int draws_per_attempt = (P.move_mode == 1) ? 7 : (P.rotate ? 4 : 0) + 3 * nl;

    for (long long step = 0; step < a.steps; ++step) {
        for (int i = tid; i < nl; i += kThreads) { s.px[i] = s.cx[i]; s.py[i] = s.cy[i]; s.pz[i] = s.cz[i]; }   // :121
        if (P.move_mode == 1) __syncthreads();     // RIGID reads every atom of prev for the centroid

        long long rej = 0;
        bool bail = false;
        for (;;) {
            if (P.move_mode == 1) {
                // EXTENSION: rigid-body move, draws tx,ty,tz,ax,ay,az,theta (oracle: orc_rigid_move)
                const double tx = (st.draw(pos, status) - 0.5) * P.rigid_translation;
                const double ty = (st.draw(pos + 1, status) - 0.5) * P.rigid_translation;
                const double tz = (st.draw(pos + 2, status) - 0.5) * P.rigid_translation;
                double ax, ay, az, theta;
                draw_axis_angle(st, pos + 3, P.rigid_angle, status, ax, ay, az, theta);
                double gx = 0, gy = 0, gz = 0;
                for (int i = 0; i < nl; ++i) { gx += s.px[i]; gy += s.py[i]; gz += s.pz[i]; }
                if (nl > 0) { gx /= (double)nl; gy /= (double)nl; gz /= (double)nl; }
                const double h = 0.5 * theta;
                double sh, w; sincos(h, &sh, &w);
                const double qx = sh * ax, qy = sh * ay, qz = sh * az;
                for (int i = tid; i < nl; i += kThreads) {
                    const double vx = s.px[i] - gx, vy = s.py[i] - gy, vz = s.pz[i] - gz;
                    const double c1x = qy * vz - qz * vy, c1y = qz * vx - qx * vz, c1z = qx * vy - qy * vx;
                    const double c2x = qy * c1z - qz * c1y, c2y = qz * c1x - qx * c1z, c2z = qx * c1y - qy * c1x;
                    const double x = gx + (vx + 2.0 * (w * c1x + c2x)) + tx;
                    const double y = gy + (vy + 2.0 * (w * c1y + c2y)) + ty;
                    const double z = gz + (vz + 2.0 * (w * c1z + c2z)) + tz;
                    s.cx[i] = x; s.cy[i] = y; s.cz[i] = z;
                    s.lf[i] = make_float4((float)(x - a.rec.cx), (float)(y - a.rec.cy), (float)(z - a.rec.cz), (float)s.q[i]);
                }
                pos += 7;
                __syncthreads();
                break;                          // rigid: intra-ligand distances unchanged, no collision test
            }
            // REFERENCE move: JitterAndRotateLigand (:172-184) / JitterLigand (:154-167), in place
            double ax = 0, ay = 0, az = 0, cos_t = 1, sin_t = 0;
            uint64_t jpos = pos;
            if (P.rotate) {                                                   // RotateLigand :189-208
                double theta;
                draw_axis_angle(st, pos, P.max_angle, status, ax, ay, az, theta);
                sincos(theta, &sin_t, &cos_t);                               // :214-215
                jpos += 4;
            }
            for (int i = tid; i < nl; i += kThreads) {
                double x = s.cx[i], y = s.cy[i], z = s.cz[i];
                if (P.rotate) rotate_atom(x, y, z, ax, ay, az, cos_t, sin_t);
                double u[3];
                st.draw3(jpos + 3ull * (uint64_t)i, u, status);
                x += (u[0] - 0.5) * P.jitter_width;                          // :176-178 / :158-160
                y += (u[1] - 0.5) * P.jitter_width;
                z += (u[2] - 0.5) * P.jitter_width;
                s.cx[i] = x; s.cy[i] = y; s.cz[i] = z;
                s.lf[i] = make_float4((float)(x - a.rec.cx), (float)(y - a.rec.cy), (float)(z - a.rec.cz), (float)s.q[i]);
            }
            pos += (uint64_t)draws_per_attempt;
            __syncthreads();
            const bool hit = collision_partial(s, nl, P.min_distance);       // :180 / :162
            if (!__syncthreads_or(hit)) break;
            ++rej;                                  // retry on the already-moved atoms
            if (rej >= P.max_retries) { bail = true; break; }
        }
        retries += rej;
        if (bail) {                                  // the reference would spin forever here
            status |= 1;
            for (int i = tid; i < nl; i += kThreads) { s.cx[i] = s.px[i]; s.cy[i] = s.py[i]; s.cz[i] = s.pz[i]; }
            break;
        }

        const double new_e = chain_energy<PRECISE>(a, s, nl, parity);         // :127
        bool acc;                                                             // AcceptMove :141-149
        const bool dh = new_e < cur_e;
        if (dh) acc = true;
        else {
            const double delta = new_e - cur_e;
            const double probability = exp(-delta / P.temperature);
            acc = st.draw(pos, status) < probability;
            pos += 1;
        }
        if (acc) { cur_e = new_e; ++accepted; downhill += dh ? 1 : 0; }      // :129-130
        else for (int i = tid; i < nl; i += kThreads) { s.cx[i] = s.px[i]; s.cy[i] = s.py[i]; s.cz[i] = s.pz[i]; }  // :132
        if (a.out_trace && tid == 0) a.out_trace[local * a.steps + step] = acc ? 1 : 0;
        ++steps_done;
    }
    __syncthreads();

    // CalculateEnergy of the returned pose in fp64 (:61 / :104)
    const double final_energy = (PRECISE && !(status & 1)) ? cur_e : chain_energy<true>(a, s, nl, parity);

    const long long out_atom = (long long)a.replicas * off + (long long)rep * nl - a.out_base;
    double *dst = a.out_xyzq + out_atom * 4;
    for (int i = tid; i < nl; i += kThreads) {
        dst[4 * i] = s.cx[i]; dst[4 * i + 1] = s.cy[i]; dst[4 * i + 2] = s.cz[i]; dst[4 * i + 3] = s.q[i];
    }
    if (tid == 0) {
        ChainStats cs;
        cs.steps = steps_done; cs.accepted = accepted; cs.downhill = downhill; cs.retries = retries;
        cs.draws = (long long)pos; cs.status = status;
        cs.final_energy = final_energy; cs.start_energy = start_energy; cs.chain_energy = cur_e;
        reinterpret_cast<ChainStats *>(a.out_stats)[local] = cs;
    }
}

// ------------------------------------------------------------------ CalculateEnergy batch (:247-262)
struct EnergyArgs {
    RecView rec;
    DevParams prm;
    const long long *offsets;   // local CSR (n+1), offsets[0] == 0
    const double *lig_xyzq;
    double *out_energy;
    double *out_abs;            // nullable
    int nl_cap;
};

template <bool PRECISE, bool WITH_ABS>
__global__ void __launch_bounds__(kThreads) energy_batch_kernel(const EnergyArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ChainSmem s(smem_raw, a.nl_cap);
    const int tid = threadIdx.x;
    const long long lig = blockIdx.x;
    const long long off = a.offsets[lig];
    const int nl = (int)(a.offsets[lig + 1] - off);
    const double *src = a.lig_xyzq + off * 4;
    for (int i = tid; i < nl; i += kThreads) {
        const double x = src[4 * i], y = src[4 * i + 1], z = src[4 * i + 2], q = src[4 * i + 3];
        s.cx[i] = x; s.cy[i] = y; s.cz[i] = z; s.q[i] = q;
        s.lf[i] = make_float4((float)(x - a.rec.cx), (float)(y - a.rec.cy), (float)(z - a.rec.cz), (float)q);
    }
    __syncthreads();
    int parity = 0;
    double part, ab = 0.0;
    if (PRECISE || WITH_ABS) {
        part = pair_sum_f64<WITH_ABS>(a.rec, s.cx, s.cy, s.cz, s.q, nl, a.prm.k_coulomb, a.prm.clamp_distance, &ab);
    }
    if (!PRECISE) part = pair_sum_f32(a.rec, s.lf, nl, a.prm.rinv_max);
    double e = block_sum(part, s.red, parity);
    if (!PRECISE) e *= a.prm.k_coulomb;
    if (WITH_ABS) ab = block_sum(ab, s.red, parity);
    if (tid == 0) {
        a.out_energy[lig] = e;
        if (WITH_ABS) a.out_abs[lig] = ab;
    }
}

// ------------------------------------------------------------------ FindClosestAtomDistance (:290-304) + shift (:267-285)
struct ClosestArgs {
    RecView rec;
    const long long *offsets;
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
    double *atoms = a.lig_xyzq + off * 4;
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
// One warp per (simulated, reference) pair; lane-strided atoms, coalesced; memory-bound.
__global__ void __launch_bounds__(256) rmsd_batch_kernel(const long long *__restrict__ offsets,
                                                         const double *__restrict__ sim,
                                                         const double *__restrict__ ref, int stride,
                                                         long long n_pairs, double *__restrict__ out) {
    const long long pair = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pair >= n_pairs) return;
    const int lane = threadIdx.x & 31;
    const long long off = offsets[pair];
    const int n = (int)(offsets[pair + 1] - off);
    double acc = 0.0;
    if (stride == 4) {
        const double4 *s4 = reinterpret_cast<const double4 *>(sim) + off;
        const double4 *r4 = reinterpret_cast<const double4 *>(ref) + off;
        for (int i = lane; i < n; i += 32) {
            const double4 sa = s4[i], ra = r4[i];
            const double dx = sa.x - ra.x, dy = sa.y - ra.y, dz = sa.z - ra.z;
            acc += dx * dx + dy * dy + dz * dz;                               // :58-62
        }
    } else {
        const double *sp = sim + off * 3, *rp = ref + off * 3;
        for (int i = lane; i < n; i += 32) {
            const double dx = sp[3 * i] - rp[3 * i], dy = sp[3 * i + 1] - rp[3 * i + 1], dz = sp[3 * i + 2] - rp[3 * i + 2];
            acc += dx * dx + dy * dy + dz * dz;
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) out[pair] = sqrt(acc / (double)n);                         // :64 (n == 0 -> NaN)
}

// ------------------------------------------------------------------ RandomizeLigandPose (validateResults.go:70-79)
__global__ void __launch_bounds__(kThreads) randomize_pose_kernel(const long long *offsets, double *lig_xyzq,
                                                                  uint64_t seed, uint64_t first_index) {
    const long long lig = blockIdx.x;
    const long long off = offsets[lig];
    const int nl = (int)(offsets[lig + 1] - off);
    double *atoms = lig_xyzq + off * 4;
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
