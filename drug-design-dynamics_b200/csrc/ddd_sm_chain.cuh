// ddd_sm_chain.cuh -- SM-persistent engine of the Metropolis chain kernel (sm_100a).
//
// One CTA of 16 warps (512 threads, 128 registers) owns one SM and keeps up to K chains ("slots") of
// SimulateEnergyMinimizationOneProc (reference metropolisMethod/metropolis.go:116-136) resident in
// shared memory.  A chain alternates between
//   * an ENERGY phase: CalculateEnergy (:247-262) cut into work items (runs of receptor warp tiles)
//     that ANY warp of the CTA may take from the slot's counter, and
//   * a SERIAL phase run by the one warp that finished the last item: item sums added in item order,
//     AcceptMove (:141-149), the next proposal (:154-208) with its collision retries (:229-242).
// There is no CTA-wide barrier inside the step loop: a chain that is proposing (latency-bound fp64 +
// Philox work of one warp) overlaps with the pair sums of the other chains, and when chains finish the
// remaining ones inherit all 16 warps -- the SM stays busy until its last chain ends, which the
// one-CTA-per-chain engine (ddd_kernels.cuh) cannot do when a launch has only ~7 chains per SM.
// Chains beyond grid*K are pulled from a global counter as slots free up; when a launch has more chains than slots (or
// chains of unequal cost per step) a chain is cut into SEGMENTS: every few hundred steps it is parked in its own output
// records and queues behind the launch's other chains, so that any SM resumes it and all SMs run dry together
// (sm_park_chain / sm_load_chain).
//
// SPECULATIVE PROPOSALS.  At T = 310 K with K = 9e9 nearly every step is rejected, and after a reject the
// next proposal is fully determined: it starts from the unchanged current pose and from stream position pos + 1 (the
// accept draw is consumed on every uphill move, :148).  So while the pair sums of step t are being evaluated, a free
// warp already builds the proposal of step t+1 -- including all its collision retries -- into a third pose buffer
// (one more work item, handed out first).  On a reject the evaluation of step t+1 starts at once; on the rare accept the
// speculation is discarded and the proposal is redone from the new pose.  The latency-bound proposal work leaves the
// chain's critical path: a step costs max(pair sums, proposal) instead of their sum.
//
// Results do not depend on which warp computed which item or on K: item sums are stored under their
// item id and added in id order, and a chain's uniform stream is indexed by (chain id, draw index).
#pragma once

#include "ddd_kernels.cuh"

namespace ddd {

#ifndef DDD_SM_THREADS
#define DDD_SM_THREADS 512
#endif
constexpr int kSmThreads  = DDD_SM_THREADS;
constexpr int kSmWarps    = kSmThreads / 32;
constexpr int kSmMaxSlots = 16;
constexpr int kSmMaxItems = 128;            // work items per energy evaluation (item sums kept per slot)
constexpr int kSlotIdle   = 0x3fffffff;     // next_item of a slot that has nothing to hand out
constexpr int kUbufMaxAtoms = 512;         // ligands up to this size stage an attempt's jitter uniforms in shared memory (Rshiny/MCdata reaches 140)
// uniforms staged per proposal attempt for ligands of up to `cap` atoms: 3 per atom + 2 (block alignment), even
__host__ __device__ constexpr int ubuf_draws(int cap) { return ((3 * (cap < kUbufMaxAtoms ? cap : kUbufMaxAtoms) + 2) + 1) & ~1; }
constexpr int kNearCap    = 128;            // near-pair list entries per slot
constexpr int kUnitCap    = 512;            // cutoff extension: receptor units (32 atoms) within reach of one pose
#ifndef DDD_NEAR_MARGIN
#define DDD_NEAR_MARGIN 1.4
#endif
constexpr int kGroupsPerItemBusy = 10;     // 128-atom groups per item while >= 4 chains share the SM (measured: 4/8/10/14/20 -> 54.9/55.8/56.1/56.3/52.8 %)
#ifndef DDD_SM_IDLE_NS
#define DDD_SM_IDLE_NS 100
#endif
#ifndef DDD_SM_IDLE_MAX_NS
#define DDD_SM_IDLE_MAX_NS 400             // idle polling backs off from DDD_SM_IDLE_NS to this (measured: caps of 100 / 400 / 800 / 1600 ns -> lone worst chain 530 / 518 / 513 / 514 ms, lone chain without retries 129 / 130 / 133 / 143 ms)
#endif
#ifndef DDD_CUT_UPI
#define DDD_CUT_UPI 16                     // cutoff extension: receptor units per pair-sum item (a multiple of 8)
#endif
#ifndef DDD_SM_WAIT_NS
#define DDD_SM_WAIT_NS 500                 // between two looks at a ring entry that is not published yet
#endif
constexpr double kNearMargin = DDD_NEAR_MARGIN;         // a near-pair list holds every ligand pair closer than MINDISTANCE + this

enum : int { ST_START64 = 0, ST_START32 = 1, ST_STEP = 2, ST_FINAL64 = 3, ST_NEXT = 4, ST_LOAD = 5, ST_WAIT = 6 };

struct __align__(16) SmSlotCtl {
    int next_item, done_items, n_total, mode;         // n_total = n_items + has_spec; mode: 0 = fp32 items, 1 = fp64 items
    int stage, nl, status, rep;
    int i_cur, i_trial, i_spec, has_spec;             // pose buffers (0..2): accepted pose, pose being evaluated, speculation
    long long local, step;
    unsigned long long pos;                           // uniforms consumed up to and including the trial proposal
    long long accepted;
    long long downhill, retries;
    double cur_e, start_e;
    long long off;                                    // first atom of the ligand in the global CSR
    unsigned long long stream_id;
    const double *rec_ptr;                            // recorded stream (nullable)
    long long rec_len;
    int base_n;                                       // near-pair list of the current pose: -1 = not built, > kNearCap = overflow
    int list_cnt;                                     // append counter while a list is being built
    int spec_status, bailed;
    unsigned long long spec_pos;                      // results of the speculative proposal: merged only if it is adopted
    long long spec_rej;
    int n_e, upi_e;                                   // pair-sum items of the evaluation handed out, receptor units per item
    int n_surv[3];                                    // cutoff extension: units within reach of pose buffer b (-1: list overflow)
    int best_is_cur; float qsum; int pad2;            // best-ever tracker: the current pose is the best one; qsum = sum of the fp32 charges (cutoff extension)
    long long units_done;                             // cutoff extension: receptor units evaluated by the chain's fp32 sums
    double best_e;                                    // lowest chain energy seen (start pose included)
    float lbox[3][8];                                 // fp32 bounding box of pose buffer b: (min xyz, -, max xyz, -)
    unsigned apart[3][2];                             // bit g: receptor group g's box is apart from pose buffer b's box (clamp-free)
    long long wait_pos;                               // ST_WAIT: the queue position (a ring entry) this slot is waiting for
};
static_assert(sizeof(SmSlotCtl) == 352, "SmSlotCtl is laid out as twenty-two 16-byte words");
constexpr int kSlotFixedBytes = (int)sizeof(SmSlotCtl) + kSmMaxItems * 8 + 2 * kNearCap * 4;      // + ubuf_draws(cap) * 8, see slot_fixed_bytes
__host__ __device__ constexpr int slot_fixed_bytes(int cap) { return kSlotFixedBytes + ubuf_draws(cap) * 8; }
constexpr int kAtomRecBytes = 88;                     // fp64 record per atom: 3 poses x (x,y,z), q, pad
constexpr int kAtomBytes = kAtomRecBytes + 3 * 16;    // + three pair-packed fp32 poses

constexpr int kCtaHeaderBytes = 16 + 4 * kSmMaxSlots;   // live-slot count, then one scheduling key per slot
constexpr int kKeyNone = 0x7fffffff;                    // key of a slot that has nothing to hand out
constexpr int kKeyWait = 0x7ffffffe;                    // key of a slot that waits for a parked chain: served when no pair-sum item is on offer
constexpr long long kStatusBestIsCur = 1ll << 30;       // parked chain records only: SmSlotCtl::best_is_cur

struct SmArgs {
    ChainArgs c;
    int *queue;              // device counter, zero before the launch: next dynamic chain = first_dynamic + atomicAdd
    long long n_local;       // chains of this launch
    long long first_dynamic; // gridDim.x * slots: chains below are assigned statically (slot s of CTA b: s*grid + b)
    int slots;               // K
    int n_units;             // receptor units of 32 atoms (one per lane)
    int upi, n_items;        // units per item; items per energy evaluation (<= kSmMaxItems)
    int cut;                 // cutoff extension on: fp32 sums run over the per-pose unit lists of the sorted receptor
    int unit_list_bytes;     // per slot: 3 * kUnitCap * 2 when cut, else 0
    int gsum_bytes;          // per slot: 128 B per group of 4 units when the fp32 sums are kept per group and lane (see
                             // sm_item_f32_groups), else 0
    int n_groups4;           // groups of 4 units (128 receptor atoms)
    int lj_bytes;            // LJ extension: per slot 24 B per ligand atom ((sigma/2, sqrt eps) fp64 + pair-packed fp32), else 0
    const double2 *lig_lj;   // LJ extension: (sigma/2, sqrt eps) of the ligand atoms of this launch (atom_base-relative)
    long long *out_units;    // nullable: per-chain count of receptor units evaluated (ddd_screen_pair_units)
    // Chain SEGMENTS (seg_steps > 0): a chain leaves its slot every seg_steps steps -- pose, energies and counters parked in
    // its own output records -- and re-enters the launch's queue behind every other chain; any SM resumes it.  The launch's
    // work unit shrinks from a whole chain to a segment, so the SMs run out of work together (see sm_park_chain).
    long long seg_steps;     // a power of two; 0: chains run to completion in the slot that loaded them
    long long n_positions;   // queue positions of the launch: n_local fresh chains + one per park (ring entry)
    int *ring;               // [n_positions - n_local], zero before the launch: entry e = local chain + 1 of the e-th park (-1: void)
    int *push_ctr;           // zero before the launch: parks so far
    int n_seg, seg_shift;    // segments per chain; seg_steps == 1 << seg_shift
    // cutoff extension: the cell list's box tables (group boxes, then unit boxes) staged in shared memory behind the slots when
    // they fit -- every proposal that passes its collision test scans them, one warp chasing loads (0: they stay in global memory)
    int box_off, box_bytes, gbox_count;      // byte offset in the CTA's dynamic shared memory; total bytes; float4 entries of the group table
};

constexpr int kMaxLaneSumGroups = 64;       // receptors up to 8 192 atoms keep per-group, per-lane fp32 sums
__host__ __device__ inline size_t sm_smem_bytes(int nl_cap, int slots, int extra_bytes = 0) {
    return kCtaHeaderBytes + (size_t)slots * (slot_fixed_bytes(nl_cap) + (size_t)nl_cap * kAtomBytes + extra_bytes);
}

// Views into the CTA's shared memory: [live 16][keys 64] then per slot [ctl][item sums][base list][dyn list][uniform staging]
// [fp64 atom records (cap x 88 B)][fp32 pair-packed poses (3 x cap x 16 B)]
struct SlotView {
    unsigned char *blk;      // this slot's block
    unsigned char *pose;     // its pose part
    int cap;
    __device__ __forceinline__ ushort2 *base_list() const { return reinterpret_cast<ushort2 *>(blk + sizeof(SmSlotCtl) + kSmMaxItems * 8); }
    __device__ __forceinline__ double *ubuf() const { return reinterpret_cast<double *>(blk + kSlotFixedBytes); }
    __device__ __forceinline__ ushort2 *dyn_list() const { return base_list() + kNearCap; }
    // One 88-byte record per atom {x0 y0 z0 x1 y1 z1 x2 y2 z2 q pad}: every field sits at a compile-time offset from
    // pose + 88 * i whatever the capacity, and half-warp 8-byte accesses of stride 88 B are bank-conflict free.
    __device__ __forceinline__ double *xyz(int buf, int i) const { return reinterpret_cast<double *>(pose + (size_t)i * kAtomRecBytes) + 3 * buf; }
    __device__ __forceinline__ double &q(int i) const { return *(reinterpret_cast<double *>(pose + (size_t)i * kAtomRecBytes) + 9); }
    __device__ __forceinline__ float4 *lf(int buf) const { return reinterpret_cast<float4 *>(pose + (size_t)cap * kAtomRecBytes) + (size_t)buf * cap; }
    // cutoff extension: ids of the receptor units within reach of pose buffer `buf`, ascending
    __device__ __forceinline__ unsigned short *units(int buf) const { return reinterpret_cast<unsigned short *>(pose + (size_t)cap * kAtomBytes) + buf * kUnitCap; }
    // per-group, per-lane fp32 partial sums of the evaluation in flight (follows the unit lists)
    __device__ __forceinline__ float *gsum(int ulb) const { return reinterpret_cast<float *>(pose + (size_t)cap * kAtomBytes + ulb); }
    // LJ extension (follows the group sums): (sigma/2, sqrt eps) per ligand atom in fp64, then pair-packed fp32
    // (sigma0/2, sigma1/2, sqrt eps0, sqrt eps1) per ligand pair
    __device__ __forceinline__ double2 *lj64(int off) const { return reinterpret_cast<double2 *>(pose + (size_t)cap * kAtomBytes + off); }
    __device__ __forceinline__ float4 *ljf(int off) const { return reinterpret_cast<float4 *>(pose + (size_t)cap * kAtomBytes + off + (size_t)cap * 16); }
};

struct SmView {
    unsigned char *raw;
    int slots, cap, ulb;
    int stride;              // bytes per slot: slot_fixed_bytes(cap) + cap * kAtomBytes + ulb (computed once, not at every use)
    __device__ __forceinline__ int *live() const { return reinterpret_cast<int *>(raw); }
    // scheduling key of a slot: the chain's step count while it has work to hand out, kKeyNone otherwise
    __device__ __forceinline__ int *keys() const { return reinterpret_cast<int *>(raw + 16); }
    __device__ __forceinline__ unsigned char *block(int s) const { return raw + kCtaHeaderBytes + s * stride; }   // ulb: all per-slot extras
    __device__ __forceinline__ SmSlotCtl *ctl(int s) const { return reinterpret_cast<SmSlotCtl *>(block(s)); }
    __device__ __forceinline__ double *items(int s) const { return reinterpret_cast<double *>(block(s) + sizeof(SmSlotCtl)); }
    __device__ __forceinline__ SlotView slot(int s) const {
        SlotView v;
        v.cap = cap;
        v.blk = block(s);
        v.pose = v.blk + slot_fixed_bytes(cap);
        return v;
    }
};

// Idle warps poll the scheduling keys with a short nanosleep in between.  (Blocking them on an mbarrier that flips on
// every hand-out was measured slower on B200: the wake-up of try_wait costs more than the poll.)
// CTA-scope acquire/release fence (__threadfence_block() would be the sequentially-consistent MEMBAR.SC.CTA)
__device__ __forceinline__ void cta_fence() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }

__device__ __forceinline__ int ld_volatile(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
__device__ __forceinline__ void st_volatile(int *p, int v) { *reinterpret_cast<volatile int *>(p) = v; }

// `gap` = DevParams::apart_gap: max(1e-3 A, 2.5 x clamp_distance) -- boxes farther apart than this cannot hold a pair inside
// the clamp distance, whatever the caller set it to (fp32 box / pose rounding is orders of magnitude below the margin)
__device__ __forceinline__ bool boxes_apart(const float4 gmin, const float4 gmax, const float *lbox, float gap) {
    return gmin.x - lbox[4] > gap || lbox[0] - gmax.x > gap || gmin.y - lbox[5] > gap ||
           lbox[1] - gmax.y > gap || gmin.z - lbox[6] > gap || lbox[2] - gmax.z > gap;
}

// Development instrumentation (-DDDD_SM_PROF; never in the shipped library): warp-cycles per phase, summed per CTA in shared
// memory and printed by CTA 0..3 at exit.  Phases: 0 pair items (fp32), 1 fp64 items, 2 speculative proposals, 3 serial phase
// (incl. inline proposals), 4 idle polling, 5 inline proposals inside the serial phase, 6 cell-list query; counts: 8 items,
// 9 spec tasks, 10 serial phases, 11 proposal attempts, 12 idle polls.
#ifdef DDD_SM_PROF
__shared__ unsigned long long ddd_prof[16];
#define PROF_T0() const long long prof_t0_ = clock64()
#define PROF_ADD(k) do { if ((threadIdx.x & 31) == 0) atomicAdd(&ddd_prof[k], (unsigned long long)(clock64() - prof_t0_)); } while (0)
#define PROF_CNT(k, n) do { if ((threadIdx.x & 31) == 0) atomicAdd(&ddd_prof[k], (unsigned long long)(n)); } while (0)
#else
#define PROF_T0() do {} while (0)
#define PROF_ADD(k) do {} while (0)
#define PROF_CNT(k, n) do {} while (0)
#endif

// MODE of a kernel instantiation: the extensions live in their own instantiations so that the default (reference) kernel
// does not carry their code -- the chain kernels are bound by instruction fetch wherever a warp is not inside the pair loop
// (ncu: no_instruction), and 16 warps in different phases share one instruction cache.
enum : int { MODE_PLAIN = 0, MODE_CUT = 1, MODE_LJ = 2 };      // MODE_LJ: Lennard-Jones, with or without the cutoff
template <int MODE> __device__ __forceinline__ bool mode_cut(const SmArgs &A) { return MODE == MODE_CUT ? true : MODE == MODE_LJ ? A.cut != 0 : false; }
template <int MODE> __device__ __forceinline__ int mode_lj_bytes(const SmArgs &A) { return MODE == MODE_LJ ? A.lj_bytes : 0; }

// ------------------------------------------------------------------ energy work items
// Receptors up to 8 192 atoms: the fp32 sum is kept PER GROUP (4 units = 128 atoms, 4 per lane) AND PER LANE in shared
// memory -- gsum[g][lane] = the lane's fp32 sum over its 4 atoms x nL ligand atoms -- and added in (group, lane) order by
// the serial phase.  No reduction on the item path, and the result does not depend on how the groups were cut into
// items, so the item size is free to follow the load: long items (8 groups) while several chains share the SM, single
// groups when one chain has it to itself.  Two adjacent groups go through the loop together (8 atoms per lane share every
// ligand load) but keep separate sums.  The clamp d >= 1e-6 (:255-257) can only matter where a group's box touches the
// pose's box; everywhere else min(1/d, 1e6) == 1/d exactly and the loop without the two FMNMX gives the same bits.
template <int MODE>
__device__ __forceinline__ void sm_item_f32_groups(const SmArgs &A, const float4 *lf, const float4 *ljf, const unsigned *apart_mask, float *gsum, int nl, int g, int g1) {
    const ulonglong2 *lf2 = reinterpret_cast<const ulonglong2 *>(lf);
    const ulonglong2 *ljp = reinterpret_cast<const ulonglong2 *>(ljf);
    const int n_pairs = (nl + 1) >> 1;
    const float rmax = A.c.prm.rinv_max;
    const float4 *rec = A.c.rec.f32s;                 // Morton order: a group is spatially compact
    const int lane = threadIdx.x & 31;
    if (mode_lj_bytes<MODE>(A)) {                       // LJ extension: one group at a time, Coulomb + 12-6
        for (; g < g1; ++g) {
            int ub[4] = {4 * g * 32, 4 * g * 32 + 32, 4 * g * 32 + 64, 4 * g * 32 + 96};
            gsum[g * 32 + lane] = warp_units_sum_lj_f32<4, false>(rec, A.c.rec.lj32s, ub, lf2, ljp, n_pairs, rmax, 0.f, A.c.prm.lj_c32, A.c.prm.lj_cap32);
        }
        return;
    }
    const unsigned long long am = (unsigned long long)apart_mask[0] | (unsigned long long)apart_mask[1] << 32;   // computed with the pose
    for (; g + 2 <= g1; g += 2) {
        const bool apart = ((am >> g) & 3ull) == 3ull;
        int ub[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) ub[a] = (4 * g + a) * 32;
        float lo, hi;
        if (apart) hi = warp_units_sum_f32<8, false, false>(rec, ub, lf2, n_pairs, rmax, 0.f, &lo);
        else hi = warp_units_sum_f32<8, false, true>(rec, ub, lf2, n_pairs, rmax, 0.f, &lo);
        gsum[g * 32 + lane] = lo;
        gsum[(g + 1) * 32 + lane] = hi;
    }
    if (g < g1) {
        const bool apart = (am >> g) & 1ull;
        int ub[4] = {4 * g * 32, 4 * g * 32 + 32, 4 * g * 32 + 64, 4 * g * 32 + 96};
        float t;
        if (apart) t = warp_units_sum_f32<4, false, false>(rec, ub, lf2, n_pairs, rmax, 0.f);
        else t = warp_units_sum_f32<4, false, true>(rec, ub, lf2, n_pairs, rmax, 0.f);
        gsum[g * 32 + lane] = t;
    }
}

// larger receptors: static items (runs of A.upi units), one fp64 sum per item
template <int MODE>
__device__ __noinline__ double sm_item_f32(const SmArgs &A, const float4 *lf, const float4 *ljf, const float *lbox, int nl, int it) {
    const ulonglong2 *lf2 = reinterpret_cast<const ulonglong2 *>(lf);
    const int n_pairs = (nl + 1) >> 1;
    int u = it * A.upi;
    const int u1 = min(u + A.upi, A.n_units);
    const float rmax = A.c.prm.rinv_max;
    const float4 *rec = A.c.rec.f32s;
    double acc = 0.0;
    if (mode_lj_bytes<MODE>(A)) {                       // LJ extension
        const ulonglong2 *ljp = reinterpret_cast<const ulonglong2 *>(ljf);
        for (; u + 4 <= u1; u += 4) {
            int ub[4] = {u * 32, u * 32 + 32, u * 32 + 64, u * 32 + 96};
            acc += (double)warp_units_sum_lj_f32<4, false>(rec, A.c.rec.lj32s, ub, lf2, ljp, n_pairs, rmax, 0.f, A.c.prm.lj_c32, A.c.prm.lj_cap32);
        }
        for (; u < u1; ++u) {
            int ub[1] = {u * 32};
            acc += (double)warp_units_sum_lj_f32<1, false>(rec, A.c.rec.lj32s, ub, lf2, ljp, n_pairs, rmax, 0.f, A.c.prm.lj_c32, A.c.prm.lj_cap32);
        }
        return warp_sum(acc);
    }
    if ((A.upi & 3) == 0) {
        for (; u + 4 <= u1; u += 4) {
            const float4 gmin = __ldg(A.c.rec.gbox + (u >> 1)), gmax = __ldg(A.c.rec.gbox + (u >> 1) + 1);
            int ub[4] = {u * 32, u * 32 + 32, u * 32 + 64, u * 32 + 96};
            if (boxes_apart(gmin, gmax, lbox, A.c.prm.apart_gap)) acc += (double)warp_units_sum_f32<4, false, false>(rec, ub, lf2, n_pairs, rmax, 0.f);
            else acc += (double)warp_units_sum_f32<4, false, true>(rec, ub, lf2, n_pairs, rmax, 0.f);
        }
    } else {
        for (; u + 4 <= u1; u += 4) acc += (double)warp_tile_sum_f32<4>(rec, u * 32, lf2, n_pairs, rmax);
    }
    if (u + 2 <= u1) { acc += (double)warp_tile_sum_f32<2>(rec, u * 32, lf2, n_pairs, rmax); u += 2; }
    if (u < u1) acc += (double)warp_tile_sum_f32<1>(rec, u * 32, lf2, n_pairs, rmax);
    return warp_sum(acc);
}

// cutoff extension: item `it` covers entries [it*upi, (it+1)*upi) of the pose's unit list (or of all units after a list
// overflow); the units are whatever the Morton order put within reach, so each call gathers its own bases.  Eight units
// (8 receptor atoms per lane) go through the loop together.
// n_surv[buf] of a listed pose: (units really within reach) << 16 | (list length: the same rounded up to a multiple of 8 with
// the receptor's FILLER unit -- 32 padding atoms, q = 0 -- so that the fp32 sums only ever run the 8-atoms-per-lane loop);
// -1: list overflow, the sums run over every unit (the sorted receptor is padded to whole groups of 8 as well)
__device__ __forceinline__ int surv_len(const SmArgs &A, int ns) { return ns >= 0 ? (ns & 0xffff) : ((A.c.rec.n_units_s + 7) & ~7); }
__device__ __forceinline__ int surv_real(const SmArgs &A, int ns) { return ns >= 0 ? (ns >> 16) : A.c.rec.n_units_s; }
template <int MODE>
__device__ __forceinline__ double sm_item_f32_cut(const SmArgs &A, const SlotView &s, const SmSlotCtl *c, int buf, int nl, int it) {
    const ulonglong2 *lf2 = reinterpret_cast<const ulonglong2 *>(s.lf(buf));
    const int n_pairs = (nl + 1) >> 1;
    const int ns = c->n_surv[buf];
    const int nu = surv_len(A, ns);
    const unsigned short *ul = s.units(buf);
    const int upi = c->upi_e;
    int j = it * upi;
    const int j1 = min(j + upi, nu);
    const float rmax = A.c.prm.rinv_max, rci = A.c.prm.rc_inv;
    const float4 *rec = A.c.rec.f32s;
    double acc = 0.0;
    if (mode_lj_bytes<MODE>(A)) {                       // LJ extension
        const ulonglong2 *ljp = reinterpret_cast<const ulonglong2 *>(s.ljf(A.unit_list_bytes + A.gsum_bytes));
        for (; j + 4 <= j1; j += 4) {
            int ub[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) ub[a] = 32 * (ns >= 0 ? (int)(ul[j + a]) : j + a);
            acc += (double)warp_units_sum_lj_f32<4, true>(rec, A.c.rec.lj32s, ub, lf2, ljp, n_pairs, rmax, rci, A.c.prm.lj_c32, A.c.prm.lj_cap32);
        }
        for (; j < j1; ++j) {
            int ub[1] = {32 * (ns >= 0 ? (int)(ul[j]) : j)};
            acc += (double)warp_units_sum_lj_f32<1, true>(rec, A.c.rec.lj32s, ub, lf2, ljp, n_pairs, rmax, rci, A.c.prm.lj_c32, A.c.prm.lj_cap32);
        }
        return warp_sum(acc);
    }
    const float rc_q = rci * c->qsum;                   // (1/cutoff) * sum_l q_l: the constant part of the shifted terms
    for (; j + 8 <= j1; j += 8) {
        int ub[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) ub[a] = 32 * (ns >= 0 ? (int)ul[j + a] : j + a);
        // (no L1 prefetch of the next call's atoms: measured, same box, 314 -> 310 ms without it -- 32 more instructions per call)
        // ONE loop variant, the clamped one: with ~60 units per evaluation this kernel is bound by instruction fetch on its
        // serial side (ncu: half of those warps' stall samples are no_instruction; the step's hot code is ~45 KB against a
        // 32 KB instruction cache), and a second copy of the loop without the clamp's two FMNMX cost more there than it saved
        // here (measured on config 4: 314 -> 307 ms)
        float lo;
        const float hi = warp_units_sum_f32<8, true, true>(rec, ub, lf2, n_pairs, rmax, rci, &lo, rc_q);
        acc += (double)lo + (double)hi;
    }
    return warp_sum(acc);                               // lists and items are whole groups of 8 units: no remainder
}

// LJ extension, one pair in fp64, in the CPU restatement's expression order: pj / lj = (sigma/2, sqrt eps)
__device__ __forceinline__ double lj_term(const DevParams &P, const double2 pj, const double2 lj, double d2) {
    const double sg = pj.x + lj.x;
    double x = d2 > 0 ? sg * sg / d2 : P.lj_cap;
    if (!(x < P.lj_cap)) x = P.lj_cap;
    const double x3 = x * x * x;
    return P.lj_scale * 4.0 * (pj.y * lj.y) * (x3 * x3 - x3);
}

// the reference's per-term fp64 arithmetic (:253-258) for the receptor atoms of one item
template <int MODE>
__device__ __noinline__ double sm_item_f64(const SmArgs &A, const SlotView &s, const SmSlotCtl *c, int buf, int nl, int it) {
    const double K = A.c.prm.k_coulomb, clampd = A.c.prm.clamp_distance, cutoff = A.c.prm.cutoff;
    const bool lj = mode_lj_bytes<MODE>(A) != 0;
    const double2 *ljl = s.lj64(A.unit_list_bytes + A.gsum_bytes);
    double e = 0.0;
    if (mode_cut<MODE>(A)) {                                                          // the pose's cell-list units, exact per-pair test
        const int ns = c->n_surv[buf];
        const int nu = surv_len(A, ns);
        const unsigned short *ul = s.units(buf);
        const int j1 = min((it + 1) * c->upi_e, nu);
        for (int j = it * c->upi_e; j < j1; ++j) {
            const int p = 32 * (ns >= 0 ? (int)(ul[j]) : j) + (int)(threadIdx.x & 31);
            if (p >= A.c.rec.n) continue;
            const Atom64 P = A.c.rec.f64s[p];
            const double2 PJ = lj ? A.c.rec.lj64s[p] : make_double2(0.0, 0.0);
            for (int l = 0; l < nl; ++l) {
                const double *L = s.xyz(buf, l);
                const double ddx = P.x - L[0], ddy = P.y - L[1], ddz = P.z - L[2];
                const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
                double d = sqrt(d2);
                if (!(d < cutoff)) continue;
                if (d < clampd) d = clampd;
                double term = K * (P.q * s.q(l)) / d - K * (P.q * s.q(l)) / cutoff;
                if (lj) term += lj_term(A.c.prm, PJ, ljl[l], d2);
                e += term;
            }
        }
        return warp_sum(e);
    }
    const int p0 = it * A.upi * 32, p1 = min((it + 1) * A.upi * 32, A.c.rec.n);
    for (int p = p0 + (int)(threadIdx.x & 31); p < p1; p += 32) {
        const Atom64 P = A.c.rec.f64[p];
        const double2 PJ = lj ? A.c.rec.lj64[p] : make_double2(0.0, 0.0);
        for (int l = 0; l < nl; ++l) {
            const double *L = s.xyz(buf, l);
            const double ddx = P.x - L[0], ddy = P.y - L[1], ddz = P.z - L[2];
            const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
            double d = sqrt(d2);                                          // :253 (Pow(x,2) == x*x)
            if (MODE != MODE_PLAIN && cutoff > 0 && !(d < cutoff)) continue;      // EXTENSION: shifted-Coulomb cutoff
            if (d < clampd) d = clampd;                                   // :255-257
            double term = K * (P.q * s.q(l)) / d;                         // :258
            if (MODE != MODE_PLAIN && cutoff > 0) term -= K * (P.q * s.q(l)) / cutoff;
            if (lj) term += lj_term(A.c.prm, PJ, ljl[l], d2);
            e += term;
        }
    }
    return warp_sum(e);
}

// ------------------------------------------------------------------ one-warp proposal helpers
// Every loop on the serial side is kept ROLLED (#pragma unroll 1): this code runs once per step per chain on one warp, it is
// bound by instruction fetch, not by issue slots -- ncu shows the SM's instruction cache missing (sm__icc_request_hit_rate 69 % with
// the cutoff extension, 93 % without) and the GPC-level cache behind it at 60-80 % of its bandwidth, so bytes of code are what counts.
// fp32 screen in front of IsCollisionFree, one warp: atoms in rows of 32 lanes; a short last row gives every atom
// 32/rem lanes that split the ligand pairs.  Same classification as collision_screen_f32 (ddd_kernels.cuh).
__device__ __noinline__ int warp_collision_screen(const float4 *lf, int nl, float m2) {
    const int lane = threadIdx.x & 31;
    const int n_pairs = (nl + 1) >> 1;
    const ulonglong2 *lf2 = reinterpret_cast<const ulonglong2 *>(lf);
    const float *lff = reinterpret_cast<const float *>(lf);
    float dmin = 3.0e38f, slack = 0.f;
    #pragma unroll 1
    for (int b = 0; b < nl; b += 32) {
        const int rem = nl - b;
        int i, k0, ks;
        if (rem >= 32) { i = b + lane; k0 = 0; ks = 1; }
        else { const int g = 32 / rem; const int grp = lane / rem; i = grp < g ? b + (lane - grp * rem) : nl; k0 = grp; ks = g; }
        if (i >= nl) continue;
        const float *mine = lff + 8 * (i >> 1) + (i & 1);
        const float xi = -mine[0], yi = -mine[2], zi = -mine[4];
        slack = fmaxf(slack, 1e-6f * (fabsf(xi) + fabsf(yi) + fabsf(zi) + 1.0f));
        const f32x2 ax = pack2(xi, xi), ay = pack2(yi, yi), az = pack2(zi, zi);
        #pragma unroll 1
        for (int k = k0; k < n_pairs; k += ks) {
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

// exact IsCollisionFree (:229-242) by one warp: pairs (i, i+m mod nl), m = 1..nl/2, as in collision_partial
__device__ __noinline__ bool warp_collision_exact(const SlotView &s, int buf, int nl, double min_distance) {
    const int full_rows = (nl - 1) / 2;
    const int n_full = full_rows * nl;
    const int total = n_full + ((nl & 1) ? 0 : nl / 2);
    const double m2 = min_distance * min_distance, m2_lo = m2 * (1.0 - 1e-9), m2_hi = m2 * (1.0 + 1e-9);
    bool hit = false;
    #pragma unroll 1
    for (int w = threadIdx.x & 31; w < total; w += 32) {
        int i, m;
        if (w < n_full) { m = w / nl + 1; i = w - (m - 1) * nl; } else { i = w - n_full; m = nl / 2; }
        int j = i + m; if (j >= nl) j -= nl;
        const double *a = s.xyz(buf, i), *b = s.xyz(buf, j);
        const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
        const double d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < m2_lo) hit = true;
        else if (d2 < m2_hi) hit |= (sqrt(d2) < min_distance);              // :235-236
    }
    return hit;
}

// fp32 pose of every atom from the fp64 master copy (the pair-packed layout the pair loop and the list scan read)
// the same without the bounding box: enough for the near-pair scan of a pose that has just failed its collision test
__device__ __noinline__ void warp_write_lf_only(const SmArgs &A, const SlotView &s, int buf, int nl) {
    float4 *lf = s.lf(buf);
    #pragma unroll 1
    for (int i = threadIdx.x & 31; i < nl; i += 32) {
        const double *a = s.xyz(buf, i);
        store_lig_f32(lf, i, (float)(a[0] - A.c.rec.cx), (float)(a[1] - A.c.rec.cy), (float)(a[2] - A.c.rec.cz), (float)s.q(i));
    }
    if ((nl & 1) && (threadIdx.x & 31) == 0) store_lig_f32(lf, nl, -kPadCoord, -kPadCoord, -kPadCoord, 0.f);   // q = 0 pad atom
    __syncwarp();
}

__device__ __noinline__ void warp_write_lf(const SmArgs &A, const SlotView &s, SmSlotCtl *c, int buf, int nl) {
    float4 *lf = s.lf(buf);
    const int lane = threadIdx.x & 31;
    float lox = 3.0e38f, loy = 3.0e38f, loz = 3.0e38f, hix = -3.0e38f, hiy = -3.0e38f, hiz = -3.0e38f;
#pragma unroll 1
    for (int i = lane; i < nl; i += 32) {
        const double *a = s.xyz(buf, i);
        const float x = (float)(a[0] - A.c.rec.cx), y = (float)(a[1] - A.c.rec.cy), z = (float)(a[2] - A.c.rec.cz);
        store_lig_f32(lf, i, x, y, z, (float)s.q(i));
        lox = fminf(lox, x); hix = fmaxf(hix, x); loy = fminf(loy, y); hiy = fmaxf(hiy, y);
        loz = fminf(loz, z); hiz = fmaxf(hiz, z);
    }
    if ((nl & 1) && lane == 0) store_lig_f32(lf, nl, -kPadCoord, -kPadCoord, -kPadCoord, 0.f);   // q = 0 pad atom
    // the pose's fp32 bounding box: lets the pair loop skip the clamp (and the cutoff extension skip whole units)
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, o)); hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, o));
        loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, o)); hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
        loz = fminf(loz, __shfl_xor_sync(0xffffffffu, loz, o)); hiz = fmaxf(hiz, __shfl_xor_sync(0xffffffffu, hiz, o));
    }
    if (lane == 0) {
        float *lb = c->lbox[buf];
        lb[0] = lox; lb[1] = loy; lb[2] = loz; lb[4] = hix; lb[5] = hiy; lb[6] = hiz;
    }
    if (A.gsum_bytes) {                                 // which 128-atom groups can skip the clamp for this pose (<= 64 groups)
        const float lb[8] = {lox, loy, loz, 0.f, hix, hiy, hiz, 0.f};
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const int g = 32 * w + lane;
            const bool ap = g < A.n_groups4 && boxes_apart(__ldg(A.c.rec.gbox + 2 * g), __ldg(A.c.rec.gbox + 2 * g + 1), lb, A.c.prm.apart_gap);
            const unsigned m = __ballot_sync(0xffffffffu, ap);
            if (lane == 0) c->apart[buf][w] = m;
        }
    }
    __syncwarp();
}

// Cutoff extension, the cell-list query of one pose: the ligand's fp32 bounding box against the boxes of the receptor's units
// (32 Morton-ordered atoms); a unit whose box is farther than the cutoff from the ligand's box cannot hold a pair inside the
// cutoff and is dropped.  Three levels of boxes -- super-groups of 32 units, groups of 4, units -- each filtered by the same
// rolled loop (warp_filter_boxes): this query runs once per step on one warp that shares its scheduler with three pair-loop
// warps, so what it costs is the number of instructions it executes (two levels: 1 100 per step for 1 563 units, ncu).  The
// survivors' ids go to the pose buffer's list in ascending order (deterministic), padded to a multiple of 8 with the filler
// unit; more than kUnitCap survivors: -1, the sums then run over every unit.
//
// One level: candidate e of [0, n_in << fan_shift) is box (in[e >> fan_shift] << fan_shift) + (e & (fan - 1)) -- every child of
// every survivor of the level above -- or simply box e when `in` is null (the top level).  Kept when it exists (< n_box) and
// lies within the cutoff of the pose's box; kept ids are appended to out[0 .. cap) in ascending order.  Returns how many were
// kept (may exceed cap: the caller treats that as an overflow).  One warp, 32 boxes per trip.
__device__ __noinline__ int warp_filter_boxes(const float4 *box, int n_box, const unsigned short *in, int n_in, int fan_shift,
                                              unsigned short *out, int cap, float lox, float loy, float loz, float hix, float hiy, float hiz, float rc2) {
    const int lane = threadIdx.x & 31;
    const int n_cand = in ? (n_in << fan_shift) : n_box;
    int count = 0;
#pragma unroll 1
    for (int b = 0; b < n_cand; b += 32) {
        const int e = min(b + lane, n_cand - 1);
        const int id = in ? ((int)in[e >> fan_shift] << fan_shift) + (e & ((1 << fan_shift) - 1)) : e;
        const int idc = min(id, n_box - 1);
        const float4 bmin = box[2 * idc], bmax = box[2 * idc + 1];
        const float gx = fmaxf(0.f, fmaxf(bmin.x - hix, lox - bmax.x));
        const float gy = fmaxf(0.f, fmaxf(bmin.y - hiy, loy - bmax.y));
        const float gz = fmaxf(0.f, fmaxf(bmin.z - hiz, loz - bmax.z));
        const bool keep = b + lane < n_cand && id < n_box && gx * gx + gy * gy + gz * gz < rc2;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int idx = count + __popc(m & ((1u << lane) - 1));
            if (idx < cap) out[idx] = (unsigned short)id;
        }
        count += __popc(m);
    }
    __syncwarp();
    return count;
}

__device__ __noinline__ void warp_build_unit_list(const SmArgs &A, const SlotView &s, SmSlotCtl *c, int buf, int nl) {
    const int lane = threadIdx.x & 31;
    const float lox = c->lbox[buf][0], loy = c->lbox[buf][1], loz = c->lbox[buf][2], hix = c->lbox[buf][4], hiy = c->lbox[buf][5], hiz = c->lbox[buf][6];
    const float rc = (float)A.c.prm.cutoff;
    const float rc2 = rc * rc * 1.0001f + 1e-3f;                      // conservative: fp32 boxes, fp32 pose
    unsigned short *ul = s.units(buf);
    const int n_units = A.c.rec.n_units_s;
    const int n_groups = (n_units + 3) >> 2, n_super = (n_groups + 7) >> 3;
    extern __shared__ __align__(16) unsigned char ddd_sm_smem[];
    // group boxes (padded to an even count), then super-group boxes, then unit boxes: staged behind the slots when they fit
    const float4 *gbox = A.box_bytes ? reinterpret_cast<const float4 *>(ddd_sm_smem + A.box_off) : A.c.rec.gbox;
    const float4 *sbox = gbox + 2 * ((n_groups + 1) & ~1);
    const float4 *ubox = A.box_bytes ? gbox + A.gbox_count : A.c.rec.ubox;
    // the survivors of the two upper levels are staged in the (idle) dynamic near-pair list
    unsigned short *sup = reinterpret_cast<unsigned short *>(s.dyn_list());
    constexpr int kSuperCap = 64, kGroupCap = kNearCap * 2 - kSuperCap;   // ushort entries; more survivors than that are > kUnitCap units anyway
    unsigned short *grp = sup + kSuperCap;
    int count = 0;
    if (nl > 0 && n_units > 0) {
        const int n_s = warp_filter_boxes(sbox, n_super, nullptr, 0, 0, sup, kSuperCap, lox, loy, loz, hix, hiy, hiz, rc2);
        count = kUnitCap + 1;
        if (n_s <= kSuperCap) {
            const int n_g = warp_filter_boxes(gbox, n_groups, sup, n_s, 3, grp, kGroupCap, lox, loy, loz, hix, hiy, hiz, rc2);
            if (n_g <= kGroupCap) count = warp_filter_boxes(ubox, n_units, grp, n_g, 2, ul, kUnitCap, lox, loy, loz, hix, hiy, hiz, rc2);
        }
    }
    if (count <= kUnitCap) {                                                // filler entries up to the next multiple of 8 (kUnitCap is one)
        const int padded = (count + 7) & ~7;
        if (count + lane < padded) ul[count + lane] = (unsigned short)((n_units + 7) & ~7);
        if (lane == 0) c->n_surv[buf] = count << 16 | padded;
    } else if (lane == 0) c->n_surv[buf] = -1;
    __syncwarp();
}

// Near-pair list of the pose in lf: every pair (i < j) that is closer than sqrt(r2) -- conservatively: the fp32 test
// carries the coordinate-rounding slack, so a listed pair may be slightly farther, never the reverse.  A pair that is
// NOT listed cannot come within MINDISTANCE during the next k_max attempts (each attempt moves a pair distance by at
// most jitter_width * sqrt(3); RotateLigand is rigid), so IsCollisionFree (:229-242) for those attempts only has to
// look at the listed pairs -- with the reference's own fp64 test.  Returns the pair count (> kNearCap: overflow).
__device__ __noinline__ int warp_build_near_list(const float4 *lf, SmSlotCtl *c, ushort2 *list, int nl, float r2) {
    const int lane = threadIdx.x & 31;
    const int n_pairs = (nl + 1) >> 1;
    const ulonglong2 *lf2 = reinterpret_cast<const ulonglong2 *>(lf);
    const float *lff = reinterpret_cast<const float *>(lf);
    if (lane == 0) c->list_cnt = 0;
    __syncwarp();
    #pragma unroll 1
    for (int b = 0; b < nl; b += 32) {
        const int rem = nl - b;
        int i, k0, ks;
        if (rem >= 32) { i = b + lane; k0 = 0; ks = 1; }
        else { const int g = 32 / rem; const int grp = lane / rem; i = grp < g ? b + (lane - grp * rem) : nl; k0 = grp; ks = g; }
        if (i >= nl) continue;
        const float *mine = lff + 8 * (i >> 1) + (i & 1);
        const float xi = -mine[0], yi = -mine[2], zi = -mine[4];
        const float lim = r2 * 1.001f + 4e-6f * (fabsf(xi) + fabsf(yi) + fabsf(zi) + 1.0f);
        const f32x2 ax = pack2(xi, xi), ay = pack2(yi, yi), az = pack2(zi, zi);
        #pragma unroll 1
        for (int k = k0; k < n_pairs; k += ks) {
            const ulonglong2 Lxy = lf2[2 * k], Lzq = lf2[2 * k + 1];
            const f32x2 dx = add2(ax, Lxy.x), dy = add2(ay, Lxy.y), dz = add2(az, Lzq.x);
            const f32x2 d2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
            float d2a, d2b;
            unpack2(d2, d2a, d2b);
            const int j0 = 2 * k, j1 = 2 * k + 1;
            if (d2a < lim && j0 > i) { const int idx = atomicAdd(&c->list_cnt, 1); if (idx < kNearCap) list[idx] = make_ushort2((unsigned short)i, (unsigned short)j0); }
            if (d2b < lim && j1 > i && j1 < nl) { const int idx = atomicAdd(&c->list_cnt, 1); if (idx < kNearCap) list[idx] = make_ushort2((unsigned short)i, (unsigned short)j1); }
        }
    }
    __syncwarp();
    return c->list_cnt;
}

// IsCollisionFree (:229-242) restricted to the listed pairs, in the reference's fp64 arithmetic; true = some pair collides
__device__ __noinline__ bool warp_check_near_list(const SlotView &s, int buf, const ushort2 *list, int n, double min_distance) {
    const double m2 = min_distance * min_distance, m2_lo = m2 * (1.0 - 1e-9), m2_hi = m2 * (1.0 + 1e-9);
    bool hit = false;
    #pragma unroll 1
    for (int e = threadIdx.x & 31; e < n; e += 32) {
        const ushort2 pr = list[e];
        const double *a = s.xyz(buf, pr.x), *b = s.xyz(buf, pr.y);
        const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
        const double d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < m2_lo) hit = true;
        else if (d2 < m2_hi) hit |= (sqrt(d2) < min_distance);              // :235-236
    }
    return __any_sync(0xffffffffu, hit) != 0;
}

struct ProposalResult { uint64_t pos; long long rej; int status; };

// EXTENSION (no reference row), out of line: rigid-body move
template <int MODE>
__device__ __noinline__ ProposalResult sm_propose_rigid(const SmArgs &A, const SlotView &s, SmSlotCtl *c, int src, int dst, uint64_t pos0) {
    const int lane = threadIdx.x & 31;
    const DevParams &P = A.c.prm;
    const int nl = c->nl;
    Stream st; st.seed = A.c.seed; st.id = c->stream_id; st.rec = c->rec_ptr; st.rec_len = c->rec_len;
    ProposalResult R; R.pos = pos0; R.rej = 0; R.status = 0;
    int status = 0;
    // EXTENSION: rigid-body move, draws tx,ty,tz,ax,ay,az,theta (no reference row); no collision test
    const double tx = (st.draw(pos0, status) - 0.5) * P.rigid_translation;
    const double ty = (st.draw(pos0 + 1, status) - 0.5) * P.rigid_translation;
    const double tz = (st.draw(pos0 + 2, status) - 0.5) * P.rigid_translation;
    double ax, ay, az, theta;
    draw_axis_angle(st, pos0 + 3, P.rigid_angle, status, ax, ay, az, theta);
    double gx = 0, gy = 0, gz = 0;
    #pragma unroll 1
    for (int i = 0; i < nl; ++i) { const double *p = s.xyz(src, i); gx += p[0]; gy += p[1]; gz += p[2]; }
    if (nl > 0) { gx /= (double)nl; gy /= (double)nl; gz /= (double)nl; }
    const double h = 0.5 * theta;
    double sh, w; sincos(h, &sh, &w);
    const double qx = sh * ax, qy = sh * ay, qz = sh * az;
    #pragma unroll 1
    for (int i = lane; i < nl; i += 32) {
        const double *p = s.xyz(src, i);
        const double vx = p[0] - gx, vy = p[1] - gy, vz = p[2] - gz;
        const double c1x = qy * vz - qz * vy, c1y = qz * vx - qx * vz, c1z = qx * vy - qy * vx;
        const double c2x = qy * c1z - qz * c1y, c2y = qz * c1x - qx * c1z, c2z = qx * c1y - qy * c1x;
        double *o = s.xyz(dst, i);
        o[0] = gx + (vx + 2.0 * (w * c1x + c2x)) + tx;
        o[1] = gy + (vy + 2.0 * (w * c1y + c2y)) + ty;
        o[2] = gz + (vz + 2.0 * (w * c1z + c2z)) + tz;
    }
    __syncwarp();
    warp_write_lf(A, s, c, dst, nl);
    if (mode_cut<MODE>(A)) warp_build_unit_list(A, s, c, dst, nl);
    R.pos = pos0 + 7;
    R.status = __any_sync(0xffffffffu, status & 2) ? 2 : 0;
    return R;
}


// One proposal of the chain in slot `c`: pose buffer `src` (the current pose) -> buffer `dst`, uniforms from stream
// position `pos0`.  REFERENCE move: JitterAndRotateLigand (:172-184) / JitterLigand (:154-167), retried cumulatively on
// `dst` until IsCollisionFree (:180 / :162).  Latency is what matters here (late in a chain the jitter piles atoms up
// at MINDISTANCE and a step needs tens of attempts), so:
//  * axis / angle / sincos of up to 32 FUTURE attempts are computed at once, one attempt per lane (their stream
//    positions are known in advance); an attempt fetches its own by shuffle;
//  * the jitter uniforms of an attempt are produced one Philox block per lane and staged in shared memory;
//  * the collision test looks only at the near-pair list of the current pose (or of a later attempt's pose once the
//    list's distance budget is spent); the fp32 pose is written only for the attempt that passes.
// status bit 1: retry bound hit (the pose in dst is meaningless), bit 2: recorded stream exhausted.
template <int MODE>
__device__ __noinline__ ProposalResult sm_propose(const SmArgs &A, const SlotView &s, SmSlotCtl *c, int src, int dst, uint64_t pos0) {
    const int lane = threadIdx.x & 31;
    const ChainArgs &a = A.c;
    const DevParams &P = a.prm;
    const int nl = c->nl;
    Stream st; st.seed = a.seed; st.id = c->stream_id; st.rec = c->rec_ptr; st.rec_len = c->rec_len;
    ProposalResult R; R.pos = pos0; R.rej = 0; R.status = 0;
    int status = 0;
    if (P.move_mode == 1) return sm_propose_rigid<MODE>(A, s, c, src, dst, pos0);   // EXTENSION, out of line
    const int dpa = (P.rotate ? 4 : 0) + 3 * nl;                         // draws per attempt
    const double pair_step = fabs(P.jitter_width) * 1.7320508075688774;   // max change of a pair distance per attempt
    const int k_max = pair_step > 0.0 ? (int)fmin((kNearMargin - 1e-6) / (pair_step * (1.0 + 1e-9)), 1.0e6) : 1000000;
    const float r_near = (float)(P.min_distance + kNearMargin), r2_near = r_near * r_near;
    const bool use_ubuf = !st.rec && nl <= kUbufMaxAtoms;
    int base_n = c->base_n;
    if (base_n < 0) {                                                     // first proposal after a load or an accepted move
        base_n = warp_build_near_list(s.lf(src), c, s.base_list(), nl, r2_near);
        if (lane == 0) c->base_n = base_n;
    }
    int dyn_n = -1, dyn_kb = 0, kb = 0, bstat = 0;
    bool lists_ok = true, bail = false;
    double bax = 0, bay = 0, baz = 0, bsin = 0, bcos = 1;                 // this lane's attempt of the batch
    int k = 0;                                                            // attempts made
    int from = src;
    for (;;) {
        ++k;
        PROF_CNT(11, 1);
        const uint64_t apos = pos0 + (uint64_t)(k - 1) * (uint64_t)dpa;
        double ax = 0, ay = 0, az = 0, cos_t = 1, sin_t = 0;
        uint64_t jpos = apos;
        if (P.rotate) {                                                   // RotateLigand :189-208
            int j = (k - 1) - kb;
            if (k == 1 || j == 32) {
                kb = k - 1; j = 0; bstat = 0;
                double theta;
                draw_axis_angle(st, pos0 + (uint64_t)(kb + lane) * (uint64_t)dpa, P.max_angle, bstat, bax, bay, baz, theta);
                sincos(theta, &bsin, &bcos);                             // :214-215
            }
            ax = __shfl_sync(0xffffffffu, bax, j); ay = __shfl_sync(0xffffffffu, bay, j); az = __shfl_sync(0xffffffffu, baz, j);
            sin_t = __shfl_sync(0xffffffffu, bsin, j); cos_t = __shfl_sync(0xffffffffu, bcos, j);
            status |= __shfl_sync(0xffffffffu, bstat, j);
            if (ax == 0.0 && ay == 0.0 && az == 0.0) lists_ok = false;    // Normalize no-op (datatypes.go:33-40): not a rotation
            jpos += 4;
        }
        const uint64_t b_first = jpos >> 1;
        if (use_ubuf && nl > 0) {
            const int nb = (int)(((jpos + 3ull * (uint64_t)nl - 1) >> 1) - b_first) + 1;
            double2 *ub = reinterpret_cast<double2 *>(s.ubuf());
            #pragma unroll 1
            for (int b = lane; b < nb; b += 32) {
                const uint64_t blk = b_first + (uint64_t)b;
                uint32_t o[4];
                philox4x32_10_t<10>((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)st.id, (uint32_t)(st.id >> 32),
                                    (uint32_t)st.seed, (uint32_t)(st.seed >> 32), o);
                ub[b] = make_double2(u53(o[0], o[1]), u53(o[2], o[3]));
            }
            __syncwarp();
        }
        const double *uu = s.ubuf() + (jpos - 2 * b_first);
        #pragma unroll 1
        for (int i = lane; i < nl; i += 32) {
            const double *p = s.xyz(from, i);
            double x = p[0], y = p[1], z = p[2];
            if (P.rotate) rotate_atom(x, y, z, ax, ay, az, cos_t, sin_t);
            double u[3];
            if (use_ubuf) { u[0] = uu[3 * i]; u[1] = uu[3 * i + 1]; u[2] = uu[3 * i + 2]; }
            else st.draw3(jpos + 3ull * (uint64_t)i, u, status);
            x += (u[0] - 0.5) * P.jitter_width;                          // :176-178 / :158-160
            y += (u[1] - 0.5) * P.jitter_width;
            z += (u[2] - 0.5) * P.jitter_width;
            double *o = s.xyz(dst, i);
            o[0] = x; o[1] = y; o[2] = z;
        }
        from = dst;                                                       // retries are cumulative
        __syncwarp();
        bool collided;
        if (lists_ok && base_n <= kNearCap && k <= k_max) collided = warp_check_near_list(s, dst, s.base_list(), base_n, P.min_distance);
        else if (lists_ok && dyn_n >= 0 && k - dyn_kb <= k_max) collided = warp_check_near_list(s, dst, s.dyn_list(), dyn_n, P.min_distance);
        else {
            warp_write_lf_only(A, s, dst, nl);
            const int n = warp_build_near_list(s.lf(dst), c, s.dyn_list(), nl, r2_near);
            if (n <= kNearCap) { dyn_n = n; dyn_kb = k; collided = warp_check_near_list(s, dst, s.dyn_list(), n, P.min_distance); }
            else {                                                        // list overflow: screen every pair
                dyn_n = -1;
                const int code = warp_collision_screen(s.lf(dst), nl, (float)(P.min_distance * P.min_distance));
                collided = __any_sync(0xffffffffu, code & 2) != 0;
                if (!collided && __any_sync(0xffffffffu, code & 1))
                    collided = __any_sync(0xffffffffu, warp_collision_exact(s, dst, nl, P.min_distance)) != 0;
            }
        }
        if (!collided) { warp_write_lf(A, s, c, dst, nl); break; }     // fp32 pose + bounding box of the pose that will be evaluated
        ++R.rej;
        if (R.rej >= P.max_retries) { bail = true; break; }                // the reference would spin forever here
    }
    if (mode_cut<MODE>(A) && !bail) { PROF_T0(); warp_build_unit_list(A, s, c, dst, nl); PROF_ADD(6); }         // the cell-list query of the pose that will be evaluated
    R.pos = pos0 + (uint64_t)k * (uint64_t)dpa;
    R.status = (bail ? 1 : 0) | (__any_sync(0xffffffffu, status & 2) ? 2 : 0);
    return R;
}

// Hand the slot out: pair-sum items of the trial pose (mode/stage) and, when `with_spec`, the speculative proposal of the
// following step.  False when the receptor has no atoms (nothing to hand out: the caller goes on inline).
template <int MODE>
__device__ __noinline__ bool sm_publish(const SmArgs &A, const SmView &V, int slot_id, SmSlotCtl *c, int mode, int stage, bool with_spec) {
    // pair-sum items of this evaluation: the static cut of the receptor, or (cutoff extension, fp32 sums) runs of the
    // trial pose's unit list, 4*m units per item with m as small as the item-sum array allows
    int n_e = A.n_items, upi_e = A.upi;
    if (A.gsum_bytes && mode == 0 && !mode_cut<MODE>(A)) {   // per-group sums: items are scheduling granules only
        const int live = ld_volatile(V.live());
        const int gpi = live >= 4 ? kGroupsPerItemBusy : live >= 2 ? 2 : 1;     // groups per item
        upi_e = 4 * gpi;
        n_e = (A.n_groups4 + gpi - 1) / gpi;
    }
    if (mode_cut<MODE>(A)) {
        const int ns = c->n_surv[c->i_trial];
        const int nu = surv_len(A, ns);
        upi_e = DDD_CUT_UPI * max(1, (nu + DDD_CUT_UPI * kSmMaxItems - 1) / (DDD_CUT_UPI * kSmMaxItems));  // runs of 16 units (two 8-atoms-per-lane calls), more only for huge lists
        n_e = (nu + upi_e - 1) / upi_e;
        if ((threadIdx.x & 31) == 0 && mode == 0) c->units_done += surv_real(A, ns);
    }
    if (n_e == 0 && (!with_spec || A.n_items == 0)) {
        if ((threadIdx.x & 31) == 0) { c->mode = mode; c->stage = stage; c->has_spec = 0; c->n_e = 0; }
        __syncwarp();
        return false;
    }
    if ((threadIdx.x & 31) == 0) {
        c->mode = mode; c->stage = stage; c->done_items = 0; c->has_spec = with_spec ? 1 : 0;
        c->n_e = n_e; c->upi_e = upi_e;
        c->n_total = n_e + (with_spec ? 1 : 0);
    }
    cta_fence();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        st_volatile(&c->next_item, 0);
        st_volatile(V.keys() + slot_id, (int)(c->step & 0x3fffffff));
    }
    return true;
}

// one out-of-line, rolled copy of the fp64 warp sum for the serial side (same butterfly, same bits as warp_sum)
__device__ __noinline__ double cold_warp_sum(double v) {
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Cold path, out of line (the step loop's code should stay compact: see MODE above): load the chain at launch position
// `load_pos` into the slot -- private copy of the start pose, counters, stream, fp32 pose, cell-list query.
// `resume_local` >= 0: the chain was parked by sm_park_chain (possibly on another SM): same load, then the parked pose and
// counters replace the start pose (read with ld.cg: this SM's L1 may hold the records of an earlier park of the same chain).
template <int MODE>
__device__ __noinline__ void sm_load_chain(const SmArgs &A, const SlotView &s, SmSlotCtl *c, long long load_pos, long long resume_local) {
    const int lane = threadIdx.x & 31;
    const ChainArgs &a = A.c;
    const long long local = resume_local >= 0 ? resume_local : a.order ? (long long)a.order[load_pos] : load_pos;
    const long long chain = a.chain_begin + local;
    const long long lig = chain / a.replicas;
    const long long off = a.offsets[lig];
    const int nl = (int)(a.offsets[lig + 1] - off);
    const double *src = a.lig_xyzq + (off - a.atom_base) * 4;           // private copy of the start pose
    #pragma unroll 1
    for (int i = lane; i < nl; i += 32) {
        double *o = s.xyz(0, i);
        o[0] = src[4 * i]; o[1] = src[4 * i + 1]; o[2] = src[4 * i + 2];
        s.q(i) = src[4 * i + 3];
    }
    if (mode_lj_bytes<MODE>(A)) {                                       // LJ extension: the ligand's (sigma/2, sqrt eps)
        const int lo = A.unit_list_bytes + A.gsum_bytes;
        const double2 *lsrc = A.lig_lj + (off - a.atom_base);
        float *lf4 = reinterpret_cast<float *>(s.ljf(lo));
        #pragma unroll 1
        for (int i = lane; i < nl + (nl & 1); i += 32) {
            const double2 v = i < nl ? lsrc[i] : make_double2(0.0, 0.0);
            if (i < nl) s.lj64(lo)[i] = v;
            lf4[4 * (i >> 1) + (i & 1)] = (float)v.x;
            lf4[4 * (i >> 1) + 2 + (i & 1)] = (float)v.y;
        }
    }
    if (lane == 0) {
        c->nl = nl; c->status = 0; c->rep = (int)(chain - lig * a.replicas); c->bailed = 0;
        c->i_cur = 0; c->i_trial = 0; c->i_spec = 1; c->has_spec = 0;
        c->local = local; c->step = 0; c->pos = 0; c->accepted = 0; c->downhill = 0; c->retries = 0;
        c->cur_e = 0.0; c->start_e = 0.0; c->off = off; c->base_n = -1; c->list_cnt = 0; c->units_done = 0;
        c->best_e = 0.0; c->best_is_cur = 1;
        c->stream_id = a.chain_id_base + (uint64_t)chain;
        c->rec_ptr = nullptr; c->rec_len = 0;
        if (a.recorded) {
            c->rec_ptr = a.recorded + (a.rec_offsets[chain] - a.rec_base);
            c->rec_len = a.rec_offsets[chain + 1] - a.rec_offsets[chain];
        }
    }
    __syncwarp();
    if (resume_local >= 0) {
        const long long out_atom = (long long)a.replicas * off + (long long)c->rep * nl - a.out_base;
        const double *parked = a.out_xyzq + out_atom * 4;
        #pragma unroll 1
        for (int i = lane; i < nl; i += 32) {
            double *o = s.xyz(0, i);
            o[0] = __ldcg(parked + 4 * i); o[1] = __ldcg(parked + 4 * i + 1); o[2] = __ldcg(parked + 4 * i + 2);
        }
        if (lane == 0) {
            const long long *st = reinterpret_cast<const long long *>(reinterpret_cast<const ChainStats *>(a.out_stats) + local);
            const double *sd = reinterpret_cast<const double *>(st);
            static_assert(sizeof(ChainStats) == 72, "six counters, three energies");
            c->step = __ldcg(st); c->accepted = __ldcg(st + 1); c->downhill = __ldcg(st + 2); c->retries = __ldcg(st + 3);
            c->pos = (unsigned long long)__ldcg(st + 4);
            const long long status = __ldcg(st + 5);
            c->status = (int)(status & ~kStatusBestIsCur); c->best_is_cur = (status & kStatusBestIsCur) ? 1 : 0;
            c->best_e = __ldcg(sd + 6); c->start_e = __ldcg(sd + 7); c->cur_e = __ldcg(sd + 8);
            if (A.out_units) c->units_done = __ldcg(A.out_units + local);
        }
        __syncwarp();
    }
    warp_write_lf(A, s, c, 0, nl);
    if (mode_cut<MODE>(A)) {
        if (lane == 0) c->qsum = ligand_charge_sum_f32(s.lf(0), nl);
        __syncwarp();
        warp_build_unit_list(A, s, c, 0, nl);
    }
}

// Cold path, out of line: end of a chain SEGMENT.  The chain's state between two steps is its current pose, two energies, the
// stream position and a few counters (the near-pair list, the fp32 pose and the speculation are functions of those): it is
// parked in the chain's own output records (final pose, ddd_chain_stats), and the chain's index goes to the tail of the
// launch's ring.  Whichever slot draws that queue position -- on any SM -- resumes the chain exactly where it stopped, so
// results do not change.  Why: a launch lasts as long as its slowest SM; with whole chains as the work unit the SMs run
// dry up to one chain duration apart (config 3 on 8 GPUs: 3.5 waves of chains per SM, 8 % lost), with segments one segment apart.
__device__ __noinline__ void sm_park_chain(const SmArgs &A, const SlotView &s, SmSlotCtl *c) {
    const int lane = threadIdx.x & 31;
    const ChainArgs &a = A.c;
    const int nl = c->nl;
    const long long out_atom = (long long)a.replicas * c->off + (long long)c->rep * nl - a.out_base;
    double *dst = a.out_xyzq + out_atom * 4;
    const int cu = c->i_cur;
    #pragma unroll 1
    for (int i = lane; i < nl; i += 32) {
        const double *p = s.xyz(cu, i);
        __stcg(dst + 4 * i, p[0]); __stcg(dst + 4 * i + 1, p[1]); __stcg(dst + 4 * i + 2, p[2]);
    }
    if (lane == 0) {
        long long *st = reinterpret_cast<long long *>(reinterpret_cast<ChainStats *>(a.out_stats) + c->local);
        double *sd = reinterpret_cast<double *>(st);
        __stcg(st, c->step); __stcg(st + 1, c->accepted); __stcg(st + 2, c->downhill); __stcg(st + 3, c->retries);
        __stcg(st + 4, (long long)c->pos);
        __stcg(st + 5, (long long)c->status | (c->best_is_cur ? kStatusBestIsCur : 0ll));
        __stcg(sd + 6, c->best_e); __stcg(sd + 7, c->start_e); __stcg(sd + 8, c->cur_e);
        if (A.out_units) __stcg(A.out_units + c->local, c->units_done);
    }
    __threadfence();                                    // every lane's stores are visible device-wide before the ring entry is
    __syncwarp();
    if (lane == 0) {
        const int e = atomicAdd(A.push_ctr, 1);
        atomicExch(A.ring + e, (int)c->local + 1);
    }
}

// A chain that stops early (retry bound) will not park again: its remaining ring entries are published void, so that the
// slots that draw those queue positions move on instead of waiting for ever.
__device__ __noinline__ void sm_void_parks(const SmArgs &A, long long step) {
    if ((threadIdx.x & 31) != 0) return;
    #pragma unroll 1
    for (long long k = step >> A.seg_shift; k < A.n_seg - 1; ++k) {
        const int e = atomicAdd(A.push_ctr, 1);
        atomicExch(A.ring + e, -1);
    }
}

// Cold path, out of line: the chain's results -- final pose, fused CalculateRMSD, best-ever pose, counters.
__device__ __noinline__ void sm_store_chain(const SmArgs &A, const SlotView &s, SmSlotCtl *c, double e) {
    const int lane = threadIdx.x & 31;
    const ChainArgs &a = A.c;
    const int nl = c->nl;
    const double final_energy = e;
    const long long out_atom = (long long)a.replicas * c->off + (long long)c->rep * nl - a.out_base;
    double *dst = a.out_xyzq + out_atom * 4;
    const int cu = c->i_cur;
    // fused CalculateRMSD (validateResults.go:50-65) of the final pose against the reference pose: same lane-strided
    // partial sums and butterfly as rmsd_chains_kernel, so both give the same bits
    const double4 *ref4 = a.out_rmsd ? reinterpret_cast<const double4 *>(a.ref_xyzq) + (c->off - a.atom_base) : nullptr;
    double *bdst = (a.out_best_e && c->best_is_cur) ? a.out_best_xyzq + out_atom * 4 : nullptr;
    double rsum = 0.0;
    #pragma unroll 1
    for (int i = lane; i < nl; i += 32) {
        const double *p = s.xyz(cu, i);
        dst[4 * i] = p[0]; dst[4 * i + 1] = p[1]; dst[4 * i + 2] = p[2]; dst[4 * i + 3] = s.q(i);
        if (bdst) { bdst[4 * i] = p[0]; bdst[4 * i + 1] = p[1]; bdst[4 * i + 2] = p[2]; bdst[4 * i + 3] = s.q(i); }
        if (ref4) {
            const double4 r = ref4[i];
            const double dx = p[0] - r.x, dy = p[1] - r.y, dz = p[2] - r.z;
            rsum += dx * dx + dy * dy + dz * dz;                       // :58-62
        }
    }
    if (ref4) {
        rsum = cold_warp_sum(rsum);
        if (lane == 0) a.out_rmsd[c->local] = sqrt(rsum / (double)nl);  // :64 (nl == 0 -> NaN)
    }
    if (lane == 0) {
        if (a.out_best_e) a.out_best_e[c->local] = c->best_e;
        ChainStats cs;
        cs.accepted = c->accepted; cs.downhill = c->downhill; cs.retries = c->retries; cs.steps = c->step;
        cs.draws = (long long)c->pos; cs.status = c->status;
        cs.final_energy = final_energy; cs.start_energy = c->start_e; cs.chain_energy = c->cur_e;
        reinterpret_cast<ChainStats *>(a.out_stats)[c->local] = cs;
        if (A.out_units) A.out_units[c->local] = c->units_done;
    }
}

// Cold path, out of line: the draw of AcceptMove's uphill branch (:147-148) when exp(-dE/T) is not exactly 0
__device__ __noinline__ bool sm_accept_draw(const ChainArgs &a, const SmSlotCtl *c, double x, uint64_t pos, int &status) {
    Stream st; st.seed = a.seed; st.id = c->stream_id; st.rec = c->rec_ptr; st.rec_len = c->rec_len;
    return st.draw(pos, status) < exp(x);
}

// ------------------------------------------------------------------ serial phase (one warp)
// `load_pos` >= 0: (re)start the slot with the chain at launch position load_pos; -1: everything handed out is complete.
template <bool PRECISE, int MODE>
__device__ __noinline__ void sm_serial_phase(const SmArgs &A, const SmView &V, int slot_id, long long load_pos) {
    const int lane = threadIdx.x & 31;
    SmSlotCtl *c = V.ctl(slot_id);
    const SlotView s = V.slot(slot_id);
    const ChainArgs &a = A.c;
    const DevParams &P = a.prm;
    int stage;
    double e = 0.0;
    bool spec_ready = false;
    if (load_pos >= 0) stage = ST_LOAD;
    else {
        stage = c->stage;
        spec_ready = c->has_spec != 0;
        const double *items = V.items(slot_id);
        if (A.gsum_bytes && c->mode == 0 && !mode_cut<MODE>(A)) {           // (group, lane) order: independent of the item cut
            const float *gs = s.gsum(A.unit_list_bytes);
            #pragma unroll 8
            for (int g = 0; g < A.n_groups4; ++g) e += (double)gs[g * 32 + lane];      // loads in flight together; the adds stay in g order
        } else {
            const int n_e = c->n_e;
            #pragma unroll 1
            for (int it = lane; it < n_e; it += 32) e += items[it];         // fixed order: independent of who computed what
        }
        e = cold_warp_sum(e);
    }
    for (;;) {
        if (stage == ST_LOAD) {
            if (load_pos < 0) {                                                 // pull the next chain of the launch
                int t = 0;
                if (lane == 0) t = atomicAdd(A.queue, 1);
                t = __shfl_sync(0xffffffffu, t, 0);
                load_pos = A.first_dynamic + (long long)t;
            }
            if (load_pos >= A.n_positions) {                                    // nothing left: the slot dies
                if (lane == 0) { st_volatile(&c->next_item, kSlotIdle); atomicSub(V.live(), 1); }
                return;
            }
            long long resume_local = -1;
            if (load_pos >= A.n_local) {                                        // a ring entry: the chain some slot parked (sm_park_chain)
                int v = 0;
                if (lane == 0) v = ld_volatile(A.ring + (load_pos - A.n_local));
                v = __shfl_sync(0xffffffffu, v, 0);
                if (v == 0) {                                                   // not parked yet: the slot waits; idle warps look again
                    __nanosleep(DDD_SM_WAIT_NS);
                    if (lane == 0) {
                        c->wait_pos = load_pos; c->mode = 2; c->stage = ST_WAIT; c->has_spec = 0; c->n_e = 0; c->done_items = 0; c->n_total = 1;
                    }
                    cta_fence();
                    __syncwarp();
                    if (lane == 0) { st_volatile(&c->next_item, 0); st_volatile(V.keys() + slot_id, kKeyWait); }
                    return;
                }
                if (v < 0) { load_pos = -1; continue; }                         // void entry (its chain stopped early)
                __threadfence();
                resume_local = v - 1;
            }
            sm_load_chain<MODE>(A, s, c, load_pos, resume_local);
            load_pos = -1;
            spec_ready = false;
            if (resume_local >= 0) stage = ST_NEXT;
            else {
                stage = ST_START64;
                if (sm_publish<MODE>(A, V, slot_id, c, 1, ST_START64, false)) return;             // :118 in fp64
                e = 0.0;
            }
        }
        if (stage == ST_START64) {
            if (lane == 0) { c->start_e = e; c->cur_e = e; c->best_e = e; }
            __syncwarp();
            if (PRECISE) stage = ST_NEXT;
            else {
                stage = ST_START32;                                             // the chain's own currentEnergy, fp32-summed
                if (sm_publish<MODE>(A, V, slot_id, c, 0, ST_START32, false)) return;
                e = 0.0;
            }
        }
        if (stage == ST_START32) {
            if (lane == 0) { c->cur_e = P.k_coulomb * e; c->best_e = c->cur_e; }
            __syncwarp();
            stage = ST_NEXT;
        }
        if (stage == ST_STEP) {                                                 // AcceptMove :141-149
            const double new_e = PRECISE ? e : P.k_coulomb * e;
            const double cur_e = c->cur_e;
            uint64_t pos = c->pos;
            int status = 0;
            bool acc;
            const bool dh = new_e < cur_e;
            if (dh) acc = true;
            else {
                const double x = -(new_e - cur_e) / P.temperature;
                // exp(x) is exactly +0 below -746 and no uniform is < 0: reject without evaluating either; the draw
                // is still consumed (:148).  (With K = 9e9 this is nearly every uphill move)
                if (x < -746.0 && !c->rec_ptr) acc = false;
                else acc = sm_accept_draw(a, c, x, pos, status);
                pos += 1;
            }
            __syncwarp();
            // EXTENSION (no reference row): best-ever pose.  While every accepted move lowers the energy the current pose IS
            // the best one; only when a move that does not improve on the best is accepted (an uphill accept, or a downhill
            // one below an earlier peak) does the pose being left have to be saved -- never at the reference constants (F7).
            bool save_best = false;
            if (acc && a.out_best_e) {
                if (new_e < c->best_e) { if (lane == 0) { c->best_e = new_e; c->best_is_cur = 1; } }
                else if (c->best_is_cur) { save_best = true; if (lane == 0) c->best_is_cur = 0; }
            }
            __syncwarp();
            if (save_best) {
                const long long out_atom = (long long)a.replicas * c->off + (long long)c->rep * c->nl - a.out_base;
                double *bdst = a.out_best_xyzq + out_atom * 4;
                const int cu = c->i_cur;
                #pragma unroll 1
                for (int i = lane; i < c->nl; i += 32) {
                    const double *p = s.xyz(cu, i);
                    bdst[4 * i] = p[0]; bdst[4 * i + 1] = p[1]; bdst[4 * i + 2] = p[2]; bdst[4 * i + 3] = s.q(i);
                }
                __syncwarp();
            }
            if (lane == 0) {
                const int t = c->i_trial, sp = c->i_spec, cu = c->i_cur;
                if (acc) { c->cur_e = new_e; c->base_n = -1; c->i_cur = t; c->i_trial = cu; c->i_spec = sp; }   // :129-130
                else { c->i_trial = sp; c->i_spec = t; }                        // :132: the current pose was never touched
                c->pos = pos; c->status |= status & 2;
                c->accepted += acc ? 1 : 0; c->downhill += dh ? 1 : 0;
                if (a.out_trace) a.out_trace[c->local * a.steps + c->step] = acc ? 1 : 0;
                c->step += 1;
            }
            __syncwarp();
            if (acc) spec_ready = false;                                        // the speculation assumed a reject
            stage = ST_NEXT;
            if (A.seg_steps > 0 && c->step < a.steps && (c->step & (A.seg_steps - 1)) == 0) {   // end of a segment: back into the queue
                sm_park_chain(A, s, c);
                __syncwarp();
                stage = ST_LOAD;
                continue;
            }
        }
        if (stage == ST_NEXT) {
            bool finish = c->step >= a.steps || c->bailed;
            if (!finish) {
                if (c->i_trial == c->i_cur) {                                   // first step: the start stages evaluated the current pose
                    __syncwarp();
                    if (lane == 0) { const int cu = c->i_cur; c->i_trial = (cu + 1) % 3; c->i_spec = (cu + 2) % 3; }
                    __syncwarp();
                }
                // the trial pose of this step: the adopted speculation (after a reject) or a proposal made here
                ProposalResult R;
                if (spec_ready) { R.pos = c->spec_pos; R.rej = c->spec_rej; R.status = c->spec_status; }
                else { PROF_T0(); R = sm_propose<MODE>(A, s, c, c->i_cur, c->i_trial, c->pos); PROF_ADD(5); }
                spec_ready = false;
                __syncwarp();
                if (lane == 0) {
                    c->pos = R.pos; c->retries += R.rej; c->status |= R.status;
                    if (R.status & 1) c->bailed = 1;
                }
                __syncwarp();
                if (R.status & 1) {                                             // retry bound: the chain ends on its current pose
                    finish = true;
                    if (A.seg_steps > 0) sm_void_parks(A, c->step);
                }
                else {
                    stage = ST_STEP;
                    if (sm_publish<MODE>(A, V, slot_id, c, PRECISE ? 1 : 0, ST_STEP, c->step + 1 < a.steps && !(A.seg_steps > 0 && ((c->step + 1) & (A.seg_steps - 1)) == 0))) return;   // :127
                    e = 0.0;
                    continue;
                }
            }
            // CalculateEnergy of the returned pose in fp64 (:61 / :104)
            if (lane == 0) c->i_trial = c->i_cur;
            __syncwarp();
            stage = ST_FINAL64;
            if (PRECISE) e = c->cur_e;
            else {
                if (sm_publish<MODE>(A, V, slot_id, c, 1, ST_FINAL64, false)) return;
                e = 0.0;
            }
        }
        if (stage == ST_FINAL64) {
            sm_store_chain(A, s, c, e);
            __syncwarp();
            stage = ST_LOAD;                                // load_pos == -1: pull from the queue
        }
    }
}

// the speculative proposal of the step after the one being evaluated: assumes a reject (pose unchanged, one accept draw)
template <int MODE>
__device__ __forceinline__ void sm_spec_task(const SmArgs &A, const SmView &V, int slot_id) {
    SmSlotCtl *c = V.ctl(slot_id);
    const SlotView s = V.slot(slot_id);
    const ProposalResult R = sm_propose<MODE>(A, s, c, c->i_cur, c->i_spec, c->pos + 1);
    __syncwarp();
    if ((threadIdx.x & 31) == 0) { c->spec_pos = R.pos; c->spec_rej = R.rej; c->spec_status = R.status; }
}

template <bool PRECISE, int MODE>
__global__ void __launch_bounds__(kSmThreads, 1) mc_sm_kernel(const __grid_constant__ SmArgs A) {
    extern __shared__ __align__(16) unsigned char ddd_sm_smem[];
    SmView V;
    V.raw = ddd_sm_smem; V.slots = A.slots; V.cap = A.c.nl_cap; V.ulb = A.unit_list_bytes + A.gsum_bytes + A.lj_bytes;
    V.stride = slot_fixed_bytes(V.cap) + V.cap * kAtomBytes + V.ulb;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = A.slots;
#ifdef DDD_SM_PROF
    if (threadIdx.x < 16) ddd_prof[threadIdx.x] = 0;
    const long long prof_start = clock64();
#endif
    if (MODE != MODE_PLAIN && A.box_bytes) {              // cutoff extension: the box tables of the cell list, staged once per CTA
        float4 *dst = reinterpret_cast<float4 *>(ddd_sm_smem + A.box_off);
        const int n_u = A.box_bytes / 16 - A.gbox_count;
        for (int i = threadIdx.x; i < A.gbox_count; i += kSmThreads) dst[i] = __ldg(A.c.rec.gbox + i);
        for (int i = threadIdx.x; i < n_u; i += kSmThreads) dst[A.gbox_count + i] = __ldg(A.c.rec.ubox + i);
    }
    if (threadIdx.x == 0) *V.live() = K;
    if (threadIdx.x < kSmMaxSlots) V.keys()[threadIdx.x] = kKeyNone;
    if (threadIdx.x < K) { V.ctl(threadIdx.x)->next_item = kSlotIdle; V.ctl(threadIdx.x)->n_total = 0; V.ctl(threadIdx.x)->step = 0; }
    __syncthreads();
    for (int sl = warp; sl < K; sl += kSmWarps) sm_serial_phase<PRECISE, MODE>(A, V, sl, (long long)sl * gridDim.x + blockIdx.x);
    const unsigned keys_s = (unsigned)__cvta_generic_to_shared(V.keys());
    int idle_n = 0;                                         // consecutive looks that found nothing

    for (;;) {
        // the chain that is furthest behind goes first: chains of an SM advance in step and end together, and one that
        // loses time in collision retries gets its pair sums served ahead of the others
        // (one key per lane, one shared-memory load and a warp-wide minimum: the 16-way unrolled scan this replaces was 75
        // instructions on every look -- 5 % of all instructions executed with the cutoff extension, ncu)
        int key = kKeyNone;
        if (lane < K) asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(key) : "r"(keys_s + 4u * (unsigned)lane) : "memory");
        const int best = __reduce_min_sync(0xffffffffu, key);
        const int sid = best == kKeyNone ? -1 : __ffs(__ballot_sync(0xffffffffu, key == best)) - 1;     // ties: the lowest slot
        if (sid < 0) {
            if (ld_volatile(V.live()) == 0) break;
            // back off while nothing turns up: a warp that polls every 100 ns issues ~0.2 instructions per cycle, and three of
            // them share a scheduler with the one warp that works through a chain's collision retries at the end of a launch
            // (ncu on a lone chain: 680 polls per attempt, 14 000 poll instructions against the attempt's 600)
            { PROF_T0(); __nanosleep(min(DDD_SM_IDLE_NS << idle_n, DDD_SM_IDLE_MAX_NS)); PROF_ADD(4); PROF_CNT(12, 1); }
            if (idle_n < 5) ++idle_n;
            continue;
        }
        idle_n = 0;
        SmSlotCtl *c = V.ctl(sid);
        int it = 0;
        if (lane == 0) it = atomicAdd(&c->next_item, 1);
        it = __shfl_sync(0xffffffffu, it, 0);
        cta_fence();
        const int n_total = ld_volatile(&c->n_total);
        // whoever takes the last piece clears the key AND parks the counter (it still holds that piece, so the slot cannot be
        // re-published before): a warp that read the key long ago and bumps the counter only now gets a value >= kSlotIdle,
        // never an index that the next hand-out (with a larger n_total) would take for one of its own pieces
        if (it == n_total - 1 && lane == 0) { st_volatile(V.keys() + sid, kKeyNone); st_volatile(&c->next_item, kSlotIdle); }
        if (it >= n_total) continue;
        if (c->mode == 2) {                                 // the slot waits for a parked chain: look at its ring entry again
            sm_serial_phase<PRECISE, MODE>(A, V, sid, c->wait_pos);
            continue;
        }
        const int has_spec = c->has_spec;                   // the speculative proposal is handed out first: it is the long pole
        if (has_spec && it == 0) { PROF_T0(); sm_spec_task<MODE>(A, V, sid); PROF_ADD(2); PROF_CNT(9, 1); }
        else {
            PROF_T0();
            const int e_it = it - has_spec;
            const SlotView s = V.slot(sid);
            const int nl = c->nl, buf = c->i_trial;
            if (c->mode == 0 && A.gsum_bytes && !mode_cut<MODE>(A)) {
                const int gpi = c->upi_e >> 2;
                sm_item_f32_groups<MODE>(A, s.lf(buf), s.ljf(A.unit_list_bytes + A.gsum_bytes), c->apart[buf], s.gsum(A.unit_list_bytes), nl,
                                   e_it * gpi, min((e_it + 1) * gpi, A.n_groups4));
            } else {
                const double sum = c->mode ? sm_item_f64<MODE>(A, s, c, buf, nl, e_it)
                                 : mode_cut<MODE>(A) ? sm_item_f32_cut<MODE>(A, s, c, buf, nl, e_it)
                                                     : sm_item_f32<MODE>(A, s.lf(buf), s.ljf(A.unit_list_bytes + A.gsum_bytes), c->lbox[buf], nl, e_it);
                if (lane == 0) V.items(sid)[e_it] = sum;
            }
            PROF_ADD(c->mode ? 1 : 0); PROF_CNT(8, 1);
        }
        cta_fence();
        int d = 0;
        if (lane == 0) d = atomicAdd(&c->done_items, 1);
        d = __shfl_sync(0xffffffffu, d, 0);
        if (d == n_total - 1) {                         // this warp completed the hand-out: it runs the serial phase
            cta_fence();
            PROF_T0();
            sm_serial_phase<PRECISE, MODE>(A, V, sid, -1);
            PROF_ADD(3); PROF_CNT(10, 1);
        }
    }
#ifdef DDD_SM_PROF
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x < 4) {
        const double tot = (double)(clock64() - prof_start) * kSmWarps;
        printf("PROF cta %d: warp-cycles %.3e | items32 %.1f%% items64 %.1f%% spec %.1f%% serial %.1f%% (inline propose %.1f%%) idle %.1f%% celllist %.1f%% | items %llu spec %llu serial %llu attempts %llu polls %llu | cyc/item %.0f cyc/spec %.0f cyc/serial %.0f\n",
               blockIdx.x, tot, 100.0 * ddd_prof[0] / tot, 100.0 * ddd_prof[1] / tot, 100.0 * ddd_prof[2] / tot, 100.0 * ddd_prof[3] / tot, 100.0 * ddd_prof[5] / tot,
               100.0 * ddd_prof[4] / tot, 100.0 * ddd_prof[6] / tot, ddd_prof[8], ddd_prof[9], ddd_prof[10], ddd_prof[11], ddd_prof[12],
               (double)ddd_prof[0] / (double)(ddd_prof[8] ? ddd_prof[8] : 1), (double)ddd_prof[2] / (double)(ddd_prof[9] ? ddd_prof[9] : 1),
               (double)ddd_prof[3] / (double)(ddd_prof[10] ? ddd_prof[10] : 1));
    }
#endif
}

}  // namespace ddd
