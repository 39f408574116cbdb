/*
 * ddd_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE).
 * See ddd_oracle.h for the scope, the parity-pinning statement and who may call this.
 *
 * Every function restates the Go reference and cites metropolisMethod/<file>:<line>.
 * float64 throughout; compile with -ffp-contract=off.
 *
 * -DORC_SQUARE_VIA_POW routes the three squares of CalculateEnergy through an
 * out-of-line pow() (build with -fno-builtin-pow) to mimic the COST of Go's
 * out-of-line math.Pow (metropolis.go:253).  That build is for timing only:
 * glibc's pow is not always correctly rounded, so it may differ from x*x in
 * the last bit, whereas Go's math.Pow(x, 2) is exactly x*x (frexp, one
 * rounded multiply of the mantissas, ldexp) -- the default build is the oracle.
 */
#include "ddd_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <float.h>

/* ------------------------------------------------------------------ params */

void orc_params_default(orc_params *p) {
    p->k_coulomb = 9e9;          /* datatypes.go:6 */
    p->threshold = 5.0;          /* datatypes.go:9 */
    p->min_distance = 0.5;       /* datatypes.go:10 */
    p->max_angle = 0.79;         /* datatypes.go:11 */
    p->jitter_width = 0.1;       /* metropolis.go:158 */
    p->clamp_distance = 1e-6;    /* metropolis.go:255 */
    p->move_mode = 0;
    p->rigid_translation = 0.5;
    p->rigid_angle = 0.79;
    p->max_retries = 1024;
    p->cutoff = 0.0;             /* extension, off: the reference sums every pair */
    p->lj_scale = 0.0;           /* extension, off: the reference has no Lennard-Jones term */
    p->lj_cap = 4.0;
    p->prot_lj = NULL; p->lig_lj = NULL;
}

/* ------------------------------------------------------------------ philox */
/* Philox4x32-10 (Salmon et al., SC'11; Random123).  Published constants. */

static inline void mulhilo32(uint32_t a, uint32_t b, uint32_t *hi, uint32_t *lo) {
    uint64_t p = (uint64_t)a * (uint64_t)b;
    *hi = (uint32_t)(p >> 32);
    *lo = (uint32_t)p;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(0xD2511F53u, c0, &hi0, &lo0);
        mulhilo32(0xCD9E8D57u, c2, &hi1, &lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n1 = lo1;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        uint32_t n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Draw `index` of stream (seed, stream_id): counter = (index/2, stream_id), key = seed;
 * the 4 output words give two 53-bit uniforms in [0,1). */
double orc_philox_u01(uint64_t seed, uint64_t stream_id, uint64_t index) {
    uint64_t blk = index >> 1;
    uint32_t ctr[4] = { (uint32_t)blk, (uint32_t)(blk >> 32),
                        (uint32_t)stream_id, (uint32_t)(stream_id >> 32) };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t o[4];
    orc_philox4x32_10(ctr, key, o);
    uint64_t w = (index & 1u) ? (((uint64_t)o[2] << 32) | o[3]) : (((uint64_t)o[0] << 32) | o[1]);
    return (double)(w >> 11) * (1.0 / 9007199254740992.0); /* 2^-53 */
}

void orc_stream_philox(orc_stream *s, uint64_t seed, uint64_t stream_id) {
    memset(s, 0, sizeof *s);
    s->kind = 0; s->seed = seed; s->stream_id = stream_id;
}

void orc_stream_recorded(orc_stream *s, const double *u, uint64_t n) {
    memset(s, 0, sizeof *s);
    s->kind = 1; s->rec = u; s->rec_len = n;
}

/* stands in for Go's rand.Float64() (math/rand; the reference never seeds it) */
double orc_stream_next(orc_stream *s) {
    uint64_t i = s->pos++;
    if (s->kind == 1) {
        if (i >= s->rec_len) { s->exhausted = 1; return 0.5; }
        return s->rec[i];
    }
    return orc_philox_u01(s->seed, s->stream_id, i);
}

/* ------------------------------------------------------------ datatypes.go */

/* datatypes.go:43-45 */
double orc_magnitude(const double v[3]) {
    return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
}

/* datatypes.go:33-40: no-op when the magnitude is 0 */
void orc_normalize(double v[3]) {
    double mag = orc_magnitude(v);
    if (mag > 0) { v[0] /= mag; v[1] /= mag; v[2] /= mag; }
}

/* datatypes.go:48-50 */
double orc_dot(const double a[3], const double b[3]) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

/* ----------------------------------------------------------- metropolis.go */

/* metropolis.go:309-314 */
double orc_distance(const double a[3], const double b[3]) {
    double dx = a[0] - b[0];
    double dy = a[1] - b[1];
    double dz = a[2] - b[2];
    return sqrt(dx * dx + dy * dy + dz * dz);
}

/* metropolis.go:213-224 (Rodrigues about an axis through the ORIGIN) */
void orc_rotate_atom(const double pos[3], const double axis[3], double theta, double out[3]) {
    double cos_t = cos(theta);
    double sin_t = sin(theta);
    double dot = orc_dot(pos, axis);
    double x = pos[0] * cos_t + sin_t * (axis[1] * pos[2] - axis[2] * pos[1]) + axis[0] * dot * (1 - cos_t);
    double y = pos[1] * cos_t + sin_t * (axis[2] * pos[0] - axis[0] * pos[2]) + axis[1] * dot * (1 - cos_t);
    double z = pos[2] * cos_t + sin_t * (axis[0] * pos[1] - axis[1] * pos[0]) + axis[2] * dot * (1 - cos_t);
    out[0] = x; out[1] = y; out[2] = z;
}

/* metropolis.go:229-242: i<j, early-out, strict '<' on the sqrt */
int orc_is_collision_free(const orc_atom *lig, int64_t nl, double min_distance) {
    for (int64_t i = 0; i < nl; ++i) {
        for (int64_t j = i + 1; j < nl; ++j) {
            double dx = lig[i].x - lig[j].x;
            double dy = lig[i].y - lig[j].y;
            double dz = lig[i].z - lig[j].z;
            double distance = sqrt(dx * dx + dy * dy + dz * dz);
            if (distance < min_distance) return 0;
        }
    }
    return 1;
}

/* ------------------------------------------------------------------ Go math.Pow (third-party to the reference)
 * metropolis.go:253 squares through math.Pow(x, 2).  The Go standard library is not under /root/reference and
 * its version is unpinned (go.mod is git-ignored; go1.23.0 inferred from the one shipped binary);
 * on amd64 math.Pow, math.Frexp, math.Ldexp and math.Modf are the portable Go functions of src/math/pow.go,
 * frexp.go, ldexp.go, modf.go (no assembly stubs), none of them inlinable.  Their published algorithm is restated
 * here so that (i) the claim "math.Pow(x, 2) == x*x bit for bit" is tested instead of assumed
 * (tests/test_oracle_reference_kats.py) and (ii) the CPU baseline can be timed with the arithmetic the Go
 * program really executes (-DORC_SQUARE_VIA_GOPOW).  noinline mirrors Go's call structure. */
#define GO_SHIFT 52
#define GO_MASK  0x7FFull
#define GO_BIAS  1023
static inline uint64_t go_bits(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }
static inline double go_frombits(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }

static inline double go_normalize(double x, int *e) {                 /* bits.go normalize */
    if (fabs(x) < 2.2250738585072014e-308) { *e = -52; return x * 4503599627370496.0; }
    *e = 0;
    return x;
}

double orc_go_frexp(double f, int *exp_out) {  /* frexp.go */
    if (f == 0 || isinf(f) || isnan(f)) { *exp_out = 0; return f; }
    int e;
    f = go_normalize(f, &e);
    uint64_t x = go_bits(f);
    e += (int)((x >> GO_SHIFT) & GO_MASK) - GO_BIAS + 1;
    x &= ~(GO_MASK << GO_SHIFT);
    x |= (uint64_t)(-1 + GO_BIAS) << GO_SHIFT;
    *exp_out = e;
    return go_frombits(x);
}

double orc_go_ldexp(double frac, int exp_) {   /* ldexp.go */
    if (frac == 0 || isinf(frac) || isnan(frac)) return frac;
    int e;
    frac = go_normalize(frac, &e);
    exp_ += e;
    uint64_t x = go_bits(frac);
    exp_ += (int)((x >> GO_SHIFT) & GO_MASK) - GO_BIAS;
    if (exp_ < -1075) return copysign(0.0, frac);                     /* underflow */
    if (exp_ > 1023) return frac < 0 ? -INFINITY : INFINITY;          /* overflow */
    double m = 1;
    if (exp_ < -1022) { exp_ += 53; m = 1.0 / 9007199254740992.0; }   /* denormal */
    x &= ~(GO_MASK << GO_SHIFT);
    x |= (uint64_t)(exp_ + GO_BIAS) << GO_SHIFT;
    return m * go_frombits(x);
}

double orc_go_modf(double f, double *frac) {   /* modf.go */
    if (f < 1) {
        if (f < 0) { double fr; const double ip = orc_go_modf(-f, &fr); *frac = -fr; return -ip; }
        if (f == 0) { *frac = f; return f; }
        *frac = f;
        return 0;
    }
    uint64_t x = go_bits(f);
    const unsigned e = (unsigned)((x >> GO_SHIFT) & GO_MASK) - GO_BIAS;
    if (e < 64 - 12) x &= ~((1ull << (64 - 12 - e)) - 1);             /* keep the top 12+e bits: the integer part */
    const double ip = go_frombits(x);
    *frac = f - ip;
    return ip;
}

static int go_is_odd_int(double x) {
    double fr;
    const double xi = orc_go_modf(x, &fr);
    return fr == 0 && ((int64_t)xi & 1) == 1;
}

__attribute__((noinline)) double orc_go_pow(double x, double y) {         /* pow.go */
    if (y == 0 || x == 1) return 1;
    if (y == 1) return x;
    if (isnan(x) || isnan(y)) return NAN;
    if (x == 0) {
        if (y < 0) return (signbit(x) && go_is_odd_int(y)) ? copysign(INFINITY, x) : INFINITY;
        return (signbit(x) && go_is_odd_int(y)) ? x : 0;
    }
    if (isinf(y)) {
        if (x == -1) return 1;
        return ((fabs(x) < 1) == (y > 0)) ? 0 : INFINITY;
    }
    if (isinf(x)) {
        if (x < 0) return orc_go_pow(1 / x, -y);
        return y < 0 ? 0 : INFINITY;
    }
    if (y == 0.5) return sqrt(x);
    if (y == -0.5) return 1 / sqrt(x);
    double yf;
    double yi = orc_go_modf(fabs(y), &yf);
    if (yf != 0 && x < 0) return NAN;
    if (yi >= 9.223372036854775808e18) {
        if (x == -1) return 1;
        return ((fabs(x) < 1) == (y > 0)) ? 0 : INFINITY;
    }
    double a1 = 1.0;                                                   /* ans = a1 * 2**ae */
    int ae = 0;
    if (yf != 0) {
        if (yf > 0.5) { yf--; yi++; }
        a1 = exp(yf * log(x));
    }
    int xe;
    double x1 = orc_go_frexp(x, &xe);
    for (int64_t i = (int64_t)yi; i != 0; i >>= 1) {
        if (xe < -(1 << 12) || (1 << 12) < xe) { ae += xe; break; }   /* catastrophic overflow */
        if (i & 1) { a1 *= x1; ae += xe; }
        x1 *= x1;
        xe <<= 1;
        if (x1 < .5) { x1 += x1; xe--; }
    }
    if (y < 0) { a1 = 1 / a1; ae = -ae; }
    return orc_go_ldexp(a1, ae);
}

#if defined(ORC_SQUARE_VIA_GOPOW)
#define ORC_SQ(x) orc_go_pow((x), 2)
#elif defined(ORC_SQUARE_VIA_POW)
#define ORC_SQ(x) pow((x), 2)
#else
#define ORC_SQ(x) ((x) * (x))     /* math.Pow(x,2) == x*x bit for bit */
#endif

/* metropolis.go:247-262: protein-outer, ligand-inner, sequential accumulation */
double orc_calculate_energy_abs(const orc_atom *prot, int64_t np, const orc_atom *lig, int64_t nl,
                                const orc_params *p, double *abs_sum) {
    double energy = 0.0, asum = 0.0;
    const double K = p->k_coulomb, clampd = p->clamp_distance, cutoff = p->cutoff;
    const int lj = p->lj_scale != 0.0 && p->prot_lj && p->lig_lj;
    for (int64_t i = 0; i < np; ++i) {
        for (int64_t j = 0; j < nl; ++j) {
            double ddx = prot[i].x - lig[j].x;
            double ddy = prot[i].y - lig[j].y;
            double ddz = prot[i].z - lig[j].z;
            double distance = sqrt(ORC_SQ(ddx) + ORC_SQ(ddy) + ORC_SQ(ddz));      /* :253 */
            if (cutoff > 0 && !(distance < cutoff)) continue;                     /* EXTENSION (no reference row) */
            if (distance < clampd) distance = clampd;                             /* :255-257 */
            double term = K * (prot[i].q * lig[j].q) / distance;                  /* :258 */
            if (cutoff > 0) term -= K * (prot[i].q * lig[j].q) / cutoff;          /* shifted: continuous at the cutoff */
            if (lj) {                                                             /* EXTENSION: Lennard-Jones 12-6 */
                const double sg = 0.5 * (p->prot_lj[2 * i] + p->lig_lj[2 * j]);
                const double d2 = ORC_SQ(ddx) + ORC_SQ(ddy) + ORC_SQ(ddz);
                double x = d2 > 0 ? sg * sg / d2 : p->lj_cap;
                if (!(x < p->lj_cap)) x = p->lj_cap;
                const double x3 = x * x * x;
                term += p->lj_scale * 4.0 * (sqrt(p->prot_lj[2 * i + 1]) * sqrt(p->lig_lj[2 * j + 1])) * (x3 * x3 - x3);
            }
            energy += term;
            asum += fabs(term);
        }
    }
    if (abs_sum) *abs_sum = asum;
    return energy;
}

double orc_calculate_energy(const orc_atom *prot, int64_t np, const orc_atom *lig, int64_t nl,
                            const orc_params *p) {
    if (p->cutoff > 0 || p->lj_scale != 0.0) return orc_calculate_energy_abs(prot, np, lig, nl, p, NULL);   /* EXTENSIONS */
    double energy = 0.0;
    const double K = p->k_coulomb, clampd = p->clamp_distance;
    for (int64_t i = 0; i < np; ++i) {
        const double px = prot[i].x, py = prot[i].y, pz = prot[i].z, pq = prot[i].q;
        for (int64_t j = 0; j < nl; ++j) {
            double ddx = px - lig[j].x;
            double ddy = py - lig[j].y;
            double ddz = pz - lig[j].z;
            double distance = sqrt(ORC_SQ(ddx) + ORC_SQ(ddy) + ORC_SQ(ddz));
            if (distance < clampd) distance = clampd;
            energy += K * (pq * lig[j].q) / distance;
        }
    }
    return energy;
}

/* metropolis.go:141-149: strict '<' short-circuit WITHOUT a draw; else one draw */
int orc_accept_move(double cur_e, double new_e, double temperature, orc_stream *s) {
    if (new_e < cur_e) return 1;
    double delta = new_e - cur_e;
    double probability = exp(-delta / temperature);
    return orc_stream_next(s) < probability;
}

/* metropolis.go:189-208: axis draws X,Y,Z (U*2-1), Normalize, theta draw, rotate in place */
void orc_rotate_ligand(orc_atom *lig, int64_t nl, double max_angle, orc_stream *s) {
    double axis[3];
    axis[0] = orc_stream_next(s) * 2 - 1;
    axis[1] = orc_stream_next(s) * 2 - 1;
    axis[2] = orc_stream_next(s) * 2 - 1;
    orc_normalize(axis);
    double theta = (orc_stream_next(s) * 2 - 1) * max_angle;
    for (int64_t i = 0; i < nl; ++i) {
        double pos[3] = { lig[i].x, lig[i].y, lig[i].z }, out[3];
        orc_rotate_atom(pos, axis, theta, out);
        lig[i].x = out[0]; lig[i].y = out[1]; lig[i].z = out[2];
    }
}

static void jitter_atoms(orc_atom *lig, int64_t nl, double width, orc_stream *s) {
    for (int64_t i = 0; i < nl; ++i) {            /* atom-major, X then Y then Z */
        lig[i].x += (orc_stream_next(s) - 0.5) * width;
        lig[i].y += (orc_stream_next(s) - 0.5) * width;
        lig[i].z += (orc_stream_next(s) - 0.5) * width;
    }
}

/* metropolis.go:154-167.  `newLigand := ligand` aliases the backing array: a
 * collision-rejected attempt is NOT undone, the next attempt jitters the moved atoms.
 * Returns the number of rejected attempts, or -(rejected) - 1 when max_retries was hit. */
int64_t orc_jitter_ligand(orc_atom *lig, int64_t nl, const orc_params *p, orc_stream *s) {
    int64_t rejected = 0;
    for (;;) {
        jitter_atoms(lig, nl, p->jitter_width, s);
        if (orc_is_collision_free(lig, nl, p->min_distance)) return rejected;
        ++rejected;
        if (p->max_retries > 0 && rejected >= p->max_retries) return -rejected - 1;
    }
}

/* metropolis.go:172-184: rotate THEN jitter, cumulative on retry */
int64_t orc_jitter_and_rotate_ligand(orc_atom *lig, int64_t nl, const orc_params *p, orc_stream *s) {
    int64_t rejected = 0;
    for (;;) {
        orc_rotate_ligand(lig, nl, p->max_angle, s);
        jitter_atoms(lig, nl, p->jitter_width, s);
        if (orc_is_collision_free(lig, nl, p->min_distance)) return rejected;
        ++rejected;
        if (p->max_retries > 0 && rejected >= p->max_retries) return -rejected - 1;
    }
}

/* EXTENSION (no reference counterpart; north_star's rigid-body move): 7 draws
 * tx,ty,tz,ax,ay,az,theta; rotate about the centroid with the unit quaternion
 * (cos t/2, sin t/2 * axis), then translate.  The ligand stays rigid, so no collision test. */
void orc_rigid_move(orc_atom *lig, int64_t nl, const orc_params *p, orc_stream *s) {
    double t[3], axis[3];
    t[0] = (orc_stream_next(s) - 0.5) * p->rigid_translation;
    t[1] = (orc_stream_next(s) - 0.5) * p->rigid_translation;
    t[2] = (orc_stream_next(s) - 0.5) * p->rigid_translation;
    axis[0] = orc_stream_next(s) * 2 - 1;
    axis[1] = orc_stream_next(s) * 2 - 1;
    axis[2] = orc_stream_next(s) * 2 - 1;
    orc_normalize(axis);
    double theta = (orc_stream_next(s) * 2 - 1) * p->rigid_angle;
    double c[3] = { 0, 0, 0 };
    for (int64_t i = 0; i < nl; ++i) { c[0] += lig[i].x; c[1] += lig[i].y; c[2] += lig[i].z; }
    if (nl > 0) { c[0] /= (double)nl; c[1] /= (double)nl; c[2] /= (double)nl; }
    double h = 0.5 * theta, w = cos(h), sh = sin(h);
    double qx = sh * axis[0], qy = sh * axis[1], qz = sh * axis[2];
    for (int64_t i = 0; i < nl; ++i) {
        double vx = lig[i].x - c[0], vy = lig[i].y - c[1], vz = lig[i].z - c[2];
        /* v' = v + 2w(q x v) + 2 q x (q x v) */
        double cx = qy * vz - qz * vy, cy = qz * vx - qx * vz, cz = qx * vy - qy * vx;
        double ccx = qy * cz - qz * cy, ccy = qz * cx - qx * cz, ccz = qx * cy - qy * cx;
        lig[i].x = c[0] + (vx + 2.0 * (w * cx + ccx)) + t[0];
        lig[i].y = c[1] + (vy + 2.0 * (w * cy + ccy)) + t[1];
        lig[i].z = c[2] + (vz + 2.0 * (w * cz + ccz)) + t[2];
    }
}

/* metropolis.go:290-304: ligand-outer, protein-inner, strict '<', init MaxFloat64 */
double orc_find_closest_atom_distance(const orc_atom *lig, int64_t nl, const orc_atom *prot, int64_t np,
                                      int64_t *lig_idx, int64_t *prot_idx) {
    double min_distance = DBL_MAX;
    int64_t li = -1, pi = -1;
    for (int64_t i = 0; i < nl; ++i) {
        for (int64_t j = 0; j < np; ++j) {
            double a[3] = { lig[i].x, lig[i].y, lig[i].z };
            double b[3] = { prot[j].x, prot[j].y, prot[j].z };
            double dist = orc_distance(a, b);
            if (dist < min_distance) { min_distance = dist; li = i; pi = j; }
        }
    }
    if (lig_idx) *lig_idx = li;
    if (prot_idx) *prot_idx = pi;
    return min_distance;
}

/* metropolis.go:267-285.  With an empty ligand or protein Go's zero-valued positions give a
 * zero shift vector (Normalize no-op) scaled by MaxFloat64-threshold = (0,0,0): no movement. */
void orc_shift_ligand_closer(orc_atom *lig, int64_t nl, const orc_atom *prot, int64_t np, double threshold) {
    int64_t li, pi;
    double closest = orc_find_closest_atom_distance(lig, nl, prot, np, &li, &pi);
    if (closest > threshold) {
        double v[3] = { 0, 0, 0 };
        if (li >= 0 && pi >= 0) {
            v[0] = prot[pi].x - lig[li].x;
            v[1] = prot[pi].y - lig[li].y;
            v[2] = prot[pi].z - lig[li].z;
        }
        orc_normalize(v);
        double sc = closest - threshold;
        v[0] *= sc; v[1] *= sc; v[2] *= sc;
        for (int64_t i = 0; i < nl; ++i) { lig[i].x += v[0]; lig[i].y += v[1]; lig[i].z += v[2]; }
    }
}

/* metropolis.go:116-136.  The reference's currentLigand aliases the caller's slice until the
 * first rejection; run on a PRIVATE copy this is exactly: propose in place, keep on accept,
 * restore prevLigand (:121 CopyLigand, :132) on reject. */
void orc_chain_one_proc(const orc_atom *prot, int64_t np, orc_atom *lig, int64_t nl,
                        int64_t iterations, int rotate, double temperature,
                        const orc_params *p, orc_stream *s, orc_chain_stats *stats,
                        uint8_t *accept_trace, double *energy_trace) {
    orc_chain_stats st; memset(&st, 0, sizeof st);
    orc_atom *prev = (orc_atom *)malloc(sizeof(orc_atom) * (size_t)(nl > 0 ? nl : 1));
    double current_energy = orc_calculate_energy(prot, np, lig, nl, p);            /* :118 */
    st.start_energy = current_energy;
    for (int64_t i = 0; i < iterations; ++i) {
        memcpy(prev, lig, sizeof(orc_atom) * (size_t)nl);                          /* :121 */
        int64_t rej;
        if (p->move_mode == 1) { orc_rigid_move(lig, nl, p, s); rej = 0; }
        else if (rotate) rej = orc_jitter_and_rotate_ligand(lig, nl, p, s);        /* :123 */
        else rej = orc_jitter_ligand(lig, nl, p, s);                               /* :125 */
        if (rej < 0) {                       /* retry bound hit: undo, flag, stop the chain */
            st.retries += -rej - 1; st.status |= 1;
            memcpy(lig, prev, sizeof(orc_atom) * (size_t)nl);
            break;
        }
        st.retries += rej;
        double new_energy = orc_calculate_energy(prot, np, lig, nl, p);            /* :127 */
        if (energy_trace) energy_trace[i] = new_energy;
        int downhill = new_energy < current_energy;
        int acc = orc_accept_move(current_energy, new_energy, temperature, s);     /* :128 */
        if (acc) { current_energy = new_energy; st.accepted++; st.downhill += downhill; } /* :129-130 */
        else memcpy(lig, prev, sizeof(orc_atom) * (size_t)nl);                     /* :132 */
        if (accept_trace) accept_trace[i] = (uint8_t)acc;
        st.steps++;
    }
    if (s->exhausted) st.status |= 2;
    st.draws = (int64_t)s->pos;
    st.final_energy = current_energy;
    free(prev);
    if (stats) *stats = st;
}

/* EXTENSION (no reference row): the same chain, additionally tracking the lowest-energy pose it visits (start pose
 * included; a later pose replaces it only on a strictly lower energy).  best_pose has nl atoms. */
void orc_chain_one_proc_best(const orc_atom *prot, int64_t np, orc_atom *lig, int64_t nl,
                             int64_t iterations, int rotate, double temperature,
                             const orc_params *p, orc_stream *s, orc_chain_stats *stats,
                             orc_atom *best_pose, double *best_energy) {
    orc_chain_stats st; memset(&st, 0, sizeof st);
    orc_atom *prev = (orc_atom *)malloc(sizeof(orc_atom) * (size_t)(nl > 0 ? nl : 1));
    double current_energy = orc_calculate_energy(prot, np, lig, nl, p);
    double best = current_energy;
    memcpy(best_pose, lig, sizeof(orc_atom) * (size_t)nl);
    st.start_energy = current_energy;
    for (int64_t i = 0; i < iterations; ++i) {
        memcpy(prev, lig, sizeof(orc_atom) * (size_t)nl);
        int64_t rej;
        if (p->move_mode == 1) { orc_rigid_move(lig, nl, p, s); rej = 0; }
        else if (rotate) rej = orc_jitter_and_rotate_ligand(lig, nl, p, s);
        else rej = orc_jitter_ligand(lig, nl, p, s);
        if (rej < 0) { st.retries += -rej - 1; st.status |= 1; memcpy(lig, prev, sizeof(orc_atom) * (size_t)nl); break; }
        st.retries += rej;
        double new_energy = orc_calculate_energy(prot, np, lig, nl, p);
        int downhill = new_energy < current_energy;
        int acc = orc_accept_move(current_energy, new_energy, temperature, s);
        if (acc) {
            current_energy = new_energy; st.accepted++; st.downhill += downhill;
            if (new_energy < best) { best = new_energy; memcpy(best_pose, lig, sizeof(orc_atom) * (size_t)nl); }
        } else memcpy(lig, prev, sizeof(orc_atom) * (size_t)nl);
        st.steps++;
    }
    if (s->exhausted) st.status |= 2;
    st.draws = (int64_t)s->pos;
    st.final_energy = current_energy;
    free(prev);
    if (stats) *stats = st;
    if (best_energy) *best_energy = best;
}

/* plotting.go:109-121 SaveMinimumEnergyLigand's choice (rule 0: minIndex 0, minEnergy 0.0) and RShinyAppMain's
 * (main.go:35-36, 51-54; rule 1: minEnergy 1e20): a sequential scan with a strict `<`. */
int64_t orc_min_energy_index(const double *energy, int64_t n, int rule, double *min_energy_out) {
    int64_t min_index = 0;
    double min_energy = rule == 1 ? 1e20 : 0.0;
    for (int64_t i = 0; i < n; ++i)
        if (energy[i] < min_energy) { min_index = i; min_energy = energy[i]; }
    if (min_energy_out) *min_energy_out = min_energy;
    return min_index;
}

/* metropolis.go:94-111 */
int orc_minimize_parallel(const orc_atom *prot, int64_t np, orc_atom *lig, int64_t nl,
                          int64_t iterations, int rotate, double temperature, int num_procs,
                          const orc_params *p, uint64_t seed, uint64_t lig_index,
                          orc_chain_stats *chain_stats, orc_atom *chain_poses,
                          int64_t *picked_replica) {
    if (num_procs <= 0) return -1;                                                 /* :97 div by zero */
    double current_energy = orc_calculate_energy(prot, np, lig, nl, p);            /* :96 */
    int64_t width = iterations / num_procs;                                        /* :97 truncating */
    size_t bytes = sizeof(orc_atom) * (size_t)(nl > 0 ? nl : 1);
    orc_atom *work = (orc_atom *)malloc(bytes * (size_t)num_procs);
    orc_stream parent; orc_stream_philox(&parent, seed, ORC_STREAM_PARENT_BASE + lig_index);
    for (int r = 0; r < num_procs; ++r) {                                          /* :99-101 */
        orc_atom *mine = work + (size_t)r * (size_t)nl;
        memcpy(mine, lig, sizeof(orc_atom) * (size_t)nl);   /* private copy of the start pose (F5) */
        orc_stream s; orc_stream_philox(&s, seed, lig_index * (uint64_t)num_procs + (uint64_t)r);
        orc_chain_stats st;
        orc_chain_one_proc(prot, np, mine, nl, width, rotate, temperature, p, &s, &st, NULL, NULL);
        if (chain_stats) chain_stats[r] = st;
    }
    if (chain_poses) memcpy(chain_poses, work, sizeof(orc_atom) * (size_t)nl * (size_t)num_procs);
    int64_t picked = -1;
    for (int r = 0; r < num_procs; ++r) {                                          /* :102-109 */
        const orc_atom *cand = work + (size_t)r * (size_t)nl;
        double new_energy = orc_calculate_energy(prot, np, cand, nl, p);           /* :104 */
        if (orc_accept_move(current_energy, new_energy, temperature, &parent)) {   /* :105 */
            memcpy(lig, cand, sizeof(orc_atom) * (size_t)nl);
            current_energy = new_energy;
            picked = r;
        }
    }
    if (picked_replica) *picked_replica = picked;
    free(work);
    return 0;
}

/* metropolis.go:34-42 */
void orc_ligand_block(int64_t n_ligands, int num_procs, int worker, int64_t *begin, int64_t *end) {
    int64_t width = n_ligands / num_procs;
    int64_t b = (int64_t)worker * width;
    int64_t e = (worker != num_procs - 1) ? b + width : n_ligands;
    *begin = b; *end = e;
}

typedef struct {
    const orc_atom *prot; int64_t np; orc_atom *lig; const int64_t *offsets; int64_t n_lig;
    int64_t iterations; int rotate; double temperature; int num_procs; const orc_params *p;
    uint64_t seed; double *energies; orc_chain_stats *chain_stats;
    int tid, n_threads; int rc;
} ml_job;

static void *ml_worker(void *arg) {
    ml_job *j = (ml_job *)arg;
    for (int64_t i = j->tid; i < j->n_lig; i += j->n_threads) {
        orc_atom *l = j->lig + j->offsets[i];
        int64_t nl = j->offsets[i + 1] - j->offsets[i];
        int rc = orc_minimize_parallel(j->prot, j->np, l, nl, j->iterations, j->rotate, j->temperature,   /* :60 */
                                       j->num_procs, j->p, j->seed, (uint64_t)i,
                                       j->chain_stats ? j->chain_stats + i * j->num_procs : NULL, NULL, NULL);
        if (rc) { j->rc = rc; return NULL; }
        j->energies[i] = orc_calculate_energy(j->prot, j->np, l, nl, j->p);        /* :61 / :19 */
    }
    return NULL;
}

/* metropolis.go:27-67 (parallel) and :11-22 (serial, shift_first=1) */
int orc_simulate_multiple_ligands(const orc_atom *prot, int64_t np, orc_atom *lig,
                                  const int64_t *offsets, int64_t n_lig, int64_t iterations,
                                  int rotate, double temperature, int num_procs,
                                  const orc_params *p, uint64_t seed, int shift_first,
                                  int n_threads, double *energies, orc_chain_stats *chain_stats) {
    if (num_procs <= 0) return -1;
    if (shift_first)                                                                /* :14-16 */
        for (int64_t i = 0; i < n_lig; ++i)
            orc_shift_ligand_closer(lig + offsets[i], offsets[i + 1] - offsets[i], prot, np, p->threshold);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256]; ml_job jobs[256];
    for (int t = 0; t < n_threads; ++t) {
        ml_job j = { prot, np, lig, offsets, n_lig, iterations, rotate, temperature, num_procs, p,
                     seed, energies, chain_stats, t, n_threads, 0 };
        jobs[t] = j;
        pthread_create(&th[t], NULL, ml_worker, &jobs[t]);
    }
    int rc = 0;
    for (int t = 0; t < n_threads; ++t) { pthread_join(th[t], NULL); if (jobs[t].rc) rc = jobs[t].rc; }
    return rc;
}

/* ------------------------------------------------------ validateResults.go */

/* validateResults.go:50-65: no superposition; divides by len(simulated) */
double orc_calculate_rmsd(const orc_atom *sim, int64_t n_sim, const orc_atom *ref, int64_t n_ref) {
    if (n_ref < n_sim) return NAN;            /* Go: index out of range panic */
    double sum_sq = 0.0;
    double num_atoms = (double)n_sim;
    for (int64_t i = 0; i < n_sim; ++i) {
        double dx = sim[i].x - ref[i].x;
        double dy = sim[i].y - ref[i].y;
        double dz = sim[i].z - ref[i].z;
        sum_sq += dx * dx + dy * dy + dz * dz;
    }
    return sqrt(sum_sq / num_atoms);          /* n_sim == 0 -> 0/0 = NaN, as in Go */
}

/* validateResults.go:70-79: per-atom +-5 A jitter (3 draws/atom), then RotateLigand(pi) */
void orc_randomize_ligand_pose(orc_atom *lig, int64_t nl, orc_stream *s) {
    for (int64_t i = 0; i < nl; ++i) {
        lig[i].x += (orc_stream_next(s) - 0.5) * 10.0;
        lig[i].y += (orc_stream_next(s) - 0.5) * 10.0;
        lig[i].z += (orc_stream_next(s) - 0.5) * 10.0;
    }
    orc_rotate_ligand(lig, nl, 3.14159265358979323846, s);   /* math.Pi */
}

/* validateResults.go:40-45.  The randomisation draws come from stream
 * (seed, ORC_STREAM_RANDOMIZE_BASE + lig_index) so they do not overlap chain or parent streams. */
double orc_compare_rmsd(const orc_atom *prot, int64_t np, orc_atom *lig, int64_t nl,
                        int64_t iterations, int rotate, double temperature, int num_procs,
                        const orc_params *p, uint64_t seed, uint64_t lig_index) {
    orc_atom *reference = (orc_atom *)malloc(sizeof(orc_atom) * (size_t)(nl > 0 ? nl : 1));
    memcpy(reference, lig, sizeof(orc_atom) * (size_t)nl);                         /* :41 */
    orc_stream rs; orc_stream_philox(&rs, seed, ORC_STREAM_RANDOMIZE_BASE + lig_index);
    orc_randomize_ligand_pose(lig, nl, &rs);                                       /* :42 */
    int rc = orc_minimize_parallel(prot, np, lig, nl, iterations, rotate, temperature, num_procs,
                                   p, seed, lig_index, NULL, NULL, NULL);           /* :43 */
    double r = rc ? NAN : orc_calculate_rmsd(lig, nl, reference, nl);              /* :44 */
    free(reference);
    return r;
}

/* -------------------------------------------------------------- parsing.go */

/* parsing.go:50-106.  strings.Fields split, >= 9 fields, x/y/z = fields 2..4, charge = last
 * field; strconv.ParseFloat errors are discarded (value 0).  Section ends at the next
 * "@<TRIPOS>" line after "@<TRIPOS>ATOM". */
static double parse_float_or_zero(const char *tok) {
    char *end = NULL;
    double v = strtod(tok, &end);
    if (end == tok || *end != '\0') return 0.0;   /* Go returns 0 with err on junk */
    return v;
}

int64_t orc_parse_mol2(const char *path, orc_atom *out, int64_t cap) {
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char *line = NULL; size_t lcap = 0; ssize_t n;
    int in_atoms = 0; int64_t count = 0;
    while ((n = getline(&line, &lcap, f)) >= 0) {
        while (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = '\0';
        if (strncmp(line, "@<TRIPOS>ATOM", 13) == 0) { in_atoms = 1; continue; }    /* :65 */
        if (in_atoms && strncmp(line, "@<TRIPOS>", 9) == 0) break;                  /* :71 */
        if (!in_atoms) continue;
        char *fields[64]; int nf = 0; char *save = NULL;
        for (char *t = strtok_r(line, " \t\v\f\r\n", &save); t; t = strtok_r(NULL, " \t\v\f\r\n", &save)) {
            if (nf < 64) fields[nf] = t;
            ++nf;
        }
        if (nf < 9) continue;                                                       /* :79 */
        int last = nf <= 64 ? nf - 1 : 63;
        if (count < cap && out) {
            out[count].x = parse_float_or_zero(fields[2]);
            out[count].y = parse_float_or_zero(fields[3]);
            out[count].z = parse_float_or_zero(fields[4]);
            out[count].q = parse_float_or_zero(fields[last]);
        }
        ++count;
    }
    free(line);
    fclose(f);
    return count;
}

/* ------------------------------------------------------ CPU baseline driver */

typedef struct {
    const orc_atom *prot; int64_t np; const orc_atom *lig; const int64_t *offsets; int64_t n_lig;
    int replicas; int64_t steps; int rotate; double temperature; const orc_params *p; uint64_t seed;
    int tid, n_threads; int64_t steps_done, pairs_done;
} bench_job;

static void *bench_worker(void *arg) {
    bench_job *j = (bench_job *)arg;
    int64_t n_chains = j->n_lig * j->replicas;
    orc_atom *mine = NULL; int64_t cap = 0;
    for (int64_t c = j->tid; c < n_chains; c += j->n_threads) {
        int64_t li = c / j->replicas;
        int64_t nl = j->offsets[li + 1] - j->offsets[li];
        if (nl > cap) { free(mine); mine = (orc_atom *)malloc(sizeof(orc_atom) * (size_t)nl); cap = nl; }
        memcpy(mine, j->lig + j->offsets[li], sizeof(orc_atom) * (size_t)nl);
        orc_stream s; orc_stream_philox(&s, j->seed, (uint64_t)c);
        orc_chain_stats st;
        orc_chain_one_proc(j->prot, j->np, mine, nl, j->steps, j->rotate, j->temperature, j->p, &s, &st, NULL, NULL);
        j->steps_done += st.steps;
        j->pairs_done += st.steps * j->np * nl;
    }
    free(mine);
    return NULL;
}

double orc_bench_chains(const orc_atom *prot, int64_t np, const orc_atom *lig,
                        const int64_t *offsets, int64_t n_lig, int replicas,
                        int64_t steps_per_chain, int rotate, double temperature,
                        const orc_params *p, uint64_t seed, int n_threads,
                        int64_t *total_steps, int64_t *total_pairs) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256]; bench_job jobs[256];
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < n_threads; ++t) {
        bench_job j = { prot, np, lig, offsets, n_lig, replicas, steps_per_chain, rotate, temperature,
                        p, seed, t, n_threads, 0, 0 };
        jobs[t] = j;
        pthread_create(&th[t], NULL, bench_worker, &jobs[t]);
    }
    int64_t ts = 0, tp = 0;
    for (int t = 0; t < n_threads; ++t) { pthread_join(th[t], NULL); ts += jobs[t].steps_done; tp += jobs[t].pairs_done; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (total_steps) *total_steps = ts;
    if (total_pairs) *total_pairs = tp;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

const char *orc_build_info(void) {
#if defined(ORC_SQUARE_VIA_GOPOW)
    return "ddd_oracle C float64 restatement (NOT Go); squares via the restated Go math.Pow (pow.go/frexp.go/ldexp.go)";
#elif defined(ORC_SQUARE_VIA_POW)
    return "ddd_oracle C float64 restatement (NOT Go); squares via out-of-line pow()";
#else
    return "ddd_oracle C float64 restatement (NOT Go); squares as x*x";
#endif
}
