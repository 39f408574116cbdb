// ddd_mc.cu -- C ABI (include/ddd_mc.h) over the sm_100a kernels in ddd_kernels.cuh.
//
// Host responsibilities: device selection, receptor replication, CSR staging, contiguous
// cost-balanced sharding of chains over devices (no collective: chains are independent,
// reference metropolis.go:27-51 / :94-111 only gathers), kernel launch + CUDA-event timing,
// and the replica pick of SimulateEnergyMinimizationParallel (:102-109).
// There is deliberately no CPU implementation of any compute entry point in this file.
#include "ddd_kernels.cuh"
#include "ddd_sm_chain.cuh"
#include "../../include/ddd_mc.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <functional>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

using namespace ddd;

static_assert(sizeof(ChainStats) == sizeof(ddd_chain_stats), "ChainStats must mirror ddd_chain_stats");
static_assert(sizeof(Atom64) == 32, "Atom64 must be 4 doubles");

namespace {

thread_local std::string g_err;
thread_local int64_t g_status_bits = 0;    // OR of the chain status bits of this thread's last ddd_minimize_batch / ddd_run_chains
thread_local int64_t g_status_chains = 0;  // how many chains of that call carried DDD_STATUS_RETRY_LIMIT
std::recursive_mutex g_mu;
uint64_t g_epoch = 1;                      // bumped by ddd_shutdown: handles created before it are stale (refused by every entry
                                           // point, freed with plain cudaFree on the device id they recorded)

void purge_cache(bool everything);         // receptor cache (defined with ddd_receptor_acquire)

struct Dev {
    int id = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 0;
    size_t max_smem = 0;                   // opt-in dynamic shared memory per CTA
};
std::vector<Dev> g_devs;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) return fail(DDD_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

#define REQUIRE_INIT()                                                                         \
    do { if (g_devs.empty()) return fail(DDD_ENODEVICE, "no CUDA device in use: call ddd_init() first (there is no CPU fallback)"); } while (0)

// Long-lived allocations (the receptor): plain cudaMalloc / cudaFree.
template <typename T>
int dev_alloc_sync(T **p, size_t count) {
    *p = nullptr;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(p), std::max<size_t>(count, 1) * sizeof(T)));
    return DDD_OK;
}

// Per-call buffers come from the device's stream-ordered memory pool on the library's own stream: cudaMalloc / cudaFree
// per call stalled ddd_minimize_batch by up to 300 ms on a busy host (measured), the pool keeps its memory
// (release threshold = max, ddd_init) and costs microseconds.  Every use of such a buffer is on the same stream.
cudaStream_t current_stream() {
    int id = 0;
    cudaGetDevice(&id);
    for (auto &d : g_devs) if (d.id == id) return d.stream;
    return nullptr;
}

template <typename T>
int dev_alloc(T **p, size_t count) {
    *p = nullptr;
    CUDA_TRY(cudaMallocAsync(reinterpret_cast<void **>(p), std::max<size_t>(count, 1) * sizeof(T), current_stream()));
    return DDD_OK;
}

void dev_free(void *p) {
    if (p) cudaFreeAsync(p, current_stream());
}

DevParams make_dev_params(const ddd_params &p, int rotate, double temperature) {
    DevParams d;
    d.k_coulomb = p.k_coulomb; d.min_distance = p.min_distance; d.max_angle = p.max_angle;
    d.jitter_width = p.jitter_width; d.clamp_distance = p.clamp_distance;
    d.rigid_translation = p.rigid_translation; d.rigid_angle = p.rigid_angle;
    d.temperature = temperature;
    d.max_retries = p.max_retries > 0 ? p.max_retries : (1ll << 62);
    d.rinv_max = p.clamp_distance > 0 ? (float)(1.0 / p.clamp_distance) : 3.402823466e38f;
    d.apart_gap = (float)std::max(1e-3, 2.5 * p.clamp_distance);
    d.move_mode = p.move_mode; d.rotate = rotate ? 1 : 0;
    d.cutoff = p.cutoff > 0 ? p.cutoff : 0.0;
    d.rc_inv = p.cutoff > 0 ? (float)(1.0 / p.cutoff) : 0.f;
    d.lj_scale = p.lj_scale; d.lj_cap = p.lj_cap;
    d.lj_c32 = p.k_coulomb != 0.0 ? (float)(4.0 * p.lj_scale / p.k_coulomb) : 0.f;   // the fp32 sums are multiplied by K at the end
    d.lj_cap32 = (float)p.lj_cap;
    return d;
}

int check_params(const ddd_params *p) {
    if (!p) return fail(DDD_EINVAL, "params is NULL (use ddd_params_default)");
    if (p->precision != DDD_PRECISION_F32 && p->precision != DDD_PRECISION_F64)
        return fail(DDD_EINVAL, "params.precision %d unknown", p->precision);
    if (p->move_mode != DDD_MOVE_REFERENCE && p->move_mode != DDD_MOVE_RIGID)
        return fail(DDD_EINVAL, "params.move_mode %d unknown", p->move_mode);
    if (!(p->cutoff >= 0.0) || std::isinf(p->cutoff)) return fail(DDD_EINVAL, "params.cutoff must be 0 (off) or a finite radius");
    if (p->lj_scale != 0.0 && !(p->lj_cap > 0.0)) return fail(DDD_EINVAL, "params.lj_cap must be positive");
    if (p->lj_scale != 0.0 && p->k_coulomb == 0.0) return fail(DDD_EINVAL, "params.lj_scale needs a non-zero k_coulomb");
    return DDD_OK;
}

int check_csr(const int64_t *offsets, int64_t n, const void *atoms, int64_t *max_nl) {
    if (n < 0) return fail(DDD_EINVAL, "negative ligand count");
    if (n > 0 && !offsets) return fail(DDD_EINVAL, "offsets is NULL");
    int64_t mx = 0;
    if (n > 0) {
        if (offsets[0] != 0) return fail(DDD_EINVAL, "offsets[0] must be 0");
        for (int64_t i = 0; i < n; ++i) {
            const int64_t nl = offsets[i + 1] - offsets[i];
            if (nl < 0) return fail(DDD_EINVAL, "offsets not monotone at ligand %lld", (long long)i);
            mx = std::max(mx, nl);
        }
        if (offsets[n] > 0 && !atoms) return fail(DDD_EINVAL, "atom array is NULL");
    }
    if (mx > DDD_MAX_LIGAND_ATOMS) return fail(DDD_ELIMIT, "ligand with %lld atoms exceeds DDD_MAX_LIGAND_ATOMS=%d", (long long)mx, DDD_MAX_LIGAND_ATOMS);
    if (max_nl) *max_nl = mx;
    return DDD_OK;
}

int even_cap(int64_t nl) { return (int)std::max<int64_t>(2, (nl + 1) & ~1ll); }

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return DDD_OK;
}

}  // namespace

// ------------------------------------------------------------------ opaque handles
struct ddd_receptor {
    uint64_t epoch = 0;                      // g_epoch at creation: a handle that outlived ddd_shutdown is stale
    std::vector<int> dev_ids;                // CUDA ordinal of every per-device copy (for frees after a shutdown)
    int64_t refs = 0;                        // receptor cache (ddd_receptor_acquire): users of this entry
    uint64_t hash = 0;                       // receptor cache: content hash of the xyzq array
    bool cached = false;
    int64_t n = 0, n_pad = 0;
    double cx = 0, cy = 0, cz = 0;
    std::vector<float4 *> f32;     // per device
    std::vector<Atom64 *> f64;
    std::vector<float4 *> f32s, ubox, gbox;   // Morton-sorted atoms, unit boxes, boxes of groups of 4 units
    std::vector<Atom64 *> f64s;
    std::vector<int64_t> perm;               // sorted position -> input index (host)
    // LJ extension: (sigma/2, sqrt(eps)) per atom, fp64 in input order and in Morton order, fp32 in Morton order
    std::vector<double2 *> lj64, lj64s;
    std::vector<float2 *> lj32s;
    bool has_lj = false;
    int64_t n_units_s = 0;
    RecView view(size_t di) const {
        RecView v; v.f32 = f32[di]; v.f64 = f64[di]; v.n = (int)n; v.n_pad = (int)n_pad; v.cx = cx; v.cy = cy; v.cz = cz;
        v.f32s = f32s[di]; v.f64s = f64s[di]; v.ubox = ubox[di]; v.gbox = gbox[di]; v.n_units_s = (int)n_units_s;
        v.lj64 = has_lj ? lj64[di] : nullptr; v.lj64s = has_lj ? lj64s[di] : nullptr; v.lj32s = has_lj ? lj32s[di] : nullptr;
        return v;
    }
};

struct ScreenDev {
    size_t di = 0;                         // index into g_devs
    int dev_id = 0;                        // CUDA ordinal (for frees after a shutdown)
    int64_t c0 = 0, c1 = 0;                // global chain range
    int64_t lig0 = 0, lig1 = 0;            // ligands touched [lig0, lig1)
    int64_t atom_base = 0, n_atoms_in = 0; // slice of the start poses
    int64_t out_base = 0, out_atoms = 0;   // slice of the chain-major output
    int nl_cap = 2;
    long long *d_offsets = nullptr;
    double *d_lig = nullptr;
    int *d_order = nullptr;
    double *d_out = nullptr;
    ChainStats *d_stats = nullptr;
    uint8_t *d_trace = nullptr;
    double *d_rec = nullptr;
    long long *d_rec_off = nullptr;
    long long rec_base = 0;
    float last_ms = 0.f;
    int threads = 0;                       // CTA width used by the last launch
    int slots = 0;                         // chains resident per SM in the last launch (0: one-CTA-per-chain engine)
    int *d_queue = nullptr;                // SM-persistent engine: [0] dynamic chain counter, [32] park counter (own 128-byte line)
    int *d_ring = nullptr;                 // chain segments: one entry per park of the launch (SmArgs::ring)
    size_t ring_cap = 0;
    long long seg_steps = 0;               // segment length of the last launch (0: whole chains)
    double *d_rmsd = nullptr;              // per-chain RMSD (ddd_screen_rmsd)
    long long *d_units = nullptr;          // per-chain receptor units evaluated (cutoff extension)
    double2 *d_lig_lj = nullptr;           // LJ extension: (sigma/2, sqrt(eps)) of this device's slice of the ligand atoms
    float rmsd_ms = 0.f;
    // resident pipeline (randomise -> shift -> chains -> pick -> RMSD -> argmin, validateResults.go:14-45 / plotting.go:109-121)
    double *d_ref = nullptr;               // reference poses: the start poses as uploaded, kept once randomise / shift touch d_lig
    double *d_best = nullptr;              // best-ever pose per chain (chain-major like d_out), EXTENSION
    double *d_best_e = nullptr;            // lowest chain energy seen, per chain
    double *d_min = nullptr;               // per-ligand picked pose (ligand CSR slice like d_lig)
    double *d_min_e = nullptr;             // per-ligand picked energy
    long long *d_picked = nullptr;         // per-ligand replica kept (-1: start pose)
    double *d_min_rmsd = nullptr;          // per-ligand RMSD(picked pose, reference)
    double *d_arg_val = nullptr;           // argmin scratch
    long long *d_arg_idx = nullptr;
    bool fused_rmsd = false;               // the last run's epilogue filled d_rmsd
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // around this screen's chain kernel: its own timing and its own completion
};

struct ddd_screen {
    uint64_t epoch = 0;
    bool track_best = false;               // ddd_screen_set_track_best
    bool picked = false;                   // ddd_screen_pick has run on the last run's results
    const ddd_receptor *rec = nullptr;
    int64_t n_lig = 0, n_chains = 0, steps = 0;
    int replicas = 1;
    uint64_t chain_id_base = 0;            // Philox stream id of local chain 0
    std::vector<int64_t> offsets;
    std::vector<ScreenDev> devs;
    bool launched = false;
    int threads_override = 0;              // > 0: one-CTA-per-chain engine with this CTA width
    int slots_override = 0;                // > 0: SM-persistent engine with this many chains per SM; both 0: automatic
    int64_t segment_override = -1;         // ddd_screen_set_segment_steps: -1 automatic, 0 whole chains, > 0 steps per segment (rounded up to a power of two)
    bool last_cut = false;                 // the last run used the cutoff extension
    bool has_lj = false;                   // ddd_screen_set_lj was called
};

namespace {

// CTA width of the chain kernel.  Every warp of a chain's CTA repeats the per-step proposal /
// collision / accept work, so the narrowest CTA that still gives each SM ~24+ warps wins.
int pick_chain_threads(int64_t n_chains, int sm_count) {
    // measured on B200 (5k x 40): 32768 chains 64 thr 55.0 % / 128 thr 54.1 % / 32 thr 51.5 % of FP32 peak;
    // 4096 chains 64 thr 51.0 % / 128 thr 51.4 %; 1024 chains 128 thr 46 % / 256 thr 48 % (but two waves)
    if (n_chains >= (int64_t)sm_count * 48) return 64;
    if (n_chains >= (int64_t)sm_count * 4) return 128;
    return 256;
}

// Chains resident per SM in the SM-persistent engine.  A launch whose chains all fit (<= grid * kSlotsAuto) is
// dealt statically, ceil(n/grid) per SM; beyond that slots refill from the launch's queue as chains finish.
constexpr int kSlotsAuto = 8;     // same-box A/B on config 3 in full (32 768 chains): 8 -> 18 244 ms, 10 -> 18 293 ms; on a 4096-chain shard 6 / 10 / 12 / 14 give 2 327 / 2 299 / 2 301 / 2 303 ms against 2 314
int pick_chain_slots(int64_t n_chains, int grid) {
    const int64_t per_sm = (n_chains + grid - 1) / std::max(grid, 1);
    return (int)std::max<int64_t>(1, std::min<int64_t>(kSlotsAuto, per_sm));
}

// chains [begin[k], begin[k+1]) go to shard k; contiguous, balanced by sum(nL + 1).  A ligand's replicas stay together on one
// shard (the cut nearest to the balanced position that falls between two ligands), so the replica pick (:102-109) and the
// per-ligand results need no exchange between devices.
void plan_shards(const int64_t *offsets, int64_t n_lig, int replicas, int n_shards, std::vector<int64_t> &begin) {
    begin.assign((size_t)n_shards + 1, 0);
    const int64_t n_chains = n_lig * replicas;
    begin[(size_t)n_shards] = n_chains;
    if (n_shards <= 1 || n_chains == 0) return;
    const double total = (double)(offsets[n_lig] + n_lig);
    int64_t lig = 0;
    for (int k = 1; k < n_shards; ++k) {
        const double target = total * (double)k / (double)n_shards;
        while (lig < n_lig && (double)(offsets[lig + 1] + lig + 1) <= target) ++lig;
        int64_t cut = lig;                                                  // ligands [0, cut) weigh <= target
        if (lig < n_lig) {
            const double before = (double)(offsets[lig] + lig), after = (double)(offsets[lig + 1] + lig + 1);
            if (after - target < target - before) cut = lig + 1;           // the nearer ligand boundary
        }
        begin[(size_t)k] = std::max(begin[(size_t)k - 1], std::min(cut * replicas, n_chains));
    }
}

// `live`: the screen belongs to the current initialisation (frees go to the device's pool on the library's stream);
// otherwise it outlived a ddd_shutdown and its buffers are released with plain cudaFree on the device they live on
void free_screen_dev(ScreenDev &sd, bool live) {
    void *bufs[] = {sd.d_offsets, sd.d_lig, sd.d_order, sd.d_out, sd.d_stats, sd.d_trace, sd.d_rec, sd.d_rec_off, sd.d_queue, sd.d_ring, sd.d_rmsd,
                    sd.d_units, sd.d_lig_lj, sd.d_ref, sd.d_best, sd.d_best_e, sd.d_min, sd.d_min_e, sd.d_picked, sd.d_min_rmsd,
                    sd.d_arg_val, sd.d_arg_idx};
    cudaSetDevice(sd.dev_id);
    for (void *b : bufs) { if (!b) continue; if (live) dev_free(b); else cudaFree(b); }
    if (sd.ev0) cudaEventDestroy(sd.ev0);
    if (sd.ev1) cudaEventDestroy(sd.ev1);
    sd = ScreenDev();
}

// grid of the persistent RMSD kernels: every resident warp of the device (256-thread CTAs), or one warp per pair when there are fewer
unsigned rmsd_grid(const Dev &dev, int64_t n_pairs) {
    static int per_sm = 0;                                                   // same kernel image on every device
    if (per_sm == 0) {
        int a = 0, b = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, rmsd_batch_kernel, 256, 0) != cudaSuccess) a = 4;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, rmsd_chains_kernel, 256, 0) != cudaSuccess) b = 4;
        per_sm = std::max(1, std::min(a, b));
    }
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n_pairs + 7) / 8, (int64_t)dev.sm_count * per_sm));
}

// Morton-sorted receptor arrays: the units rounded up to whole groups of 8 (padding atoms: q = 0, parked far away), then one
// more group of padding only -- its first unit is the FILLER that pads a pose's unit list to a multiple of 8 (ddd_sm_chain.cuh)
int64_t sorted_pad_atoms(int64_t n_units_s) { return ((n_units_s + 7) / 8 * 8 + 8) * 32; }

int check_receptor(const ddd_receptor *r) {
    if (!r) return fail(DDD_EINVAL, "receptor is NULL");
    if (r->epoch != g_epoch) return fail(DDD_EINVAL, "receptor handle is stale: it was created before the last ddd_shutdown / ddd_init");
    return DDD_OK;
}

int check_screen(const ddd_screen *s) {
    if (!s) return fail(DDD_EINVAL, "screen is NULL");
    if (s->epoch != g_epoch) return fail(DDD_EINVAL, "screen handle is stale: it was created before the last ddd_shutdown / ddd_init");
    return DDD_OK;
}

int screen_create_impl(const ddd_receptor *r, const int64_t *offsets, const double *lig_xyzq, int64_t n_lig,
                       int32_t replicas, uint64_t first_chain_id, ddd_screen **out) {
    if (!out) return fail(DDD_EINVAL, "out is NULL");
    if (int rc = check_receptor(r)) return rc;
    if (replicas <= 0) return fail(DDD_EINVAL, "replicas must be >= 1 (got %d)", replicas);
    int64_t max_nl = 0;
    if (int rc = check_csr(offsets, n_lig, lig_xyzq, &max_nl)) return rc;
    ddd_screen *s = new ddd_screen();
    s->epoch = g_epoch;
    s->rec = r; s->n_lig = n_lig; s->replicas = replicas; s->n_chains = n_lig * replicas; s->chain_id_base = first_chain_id;
    // testing knob: the default segment length of every screen this process creates, the one-shot entry points included
    // (ddd_screen_set_segment_steps overrides it); results never depend on it
    if (const char *e = std::getenv("DDD_SEGMENT_STEPS")) { const long long v = std::atoll(e); if (v >= -1) s->segment_override = v; }
    s->offsets.assign(offsets ? offsets : nullptr, offsets ? offsets + n_lig + 1 : nullptr);
    if (s->offsets.empty()) s->offsets.push_back(0);
    std::vector<int64_t> begin;
    plan_shards(s->offsets.data(), n_lig, replicas, (int)g_devs.size(), begin);
    for (size_t di = 0; di < g_devs.size(); ++di) {
        ScreenDev sd; sd.di = di; sd.dev_id = g_devs[di].id; sd.c0 = begin[di]; sd.c1 = begin[di + 1];
        if (sd.c1 <= sd.c0) continue;
        sd.lig0 = sd.c0 / replicas; sd.lig1 = (sd.c1 - 1) / replicas + 1;
        sd.atom_base = s->offsets[(size_t)sd.lig0];
        sd.n_atoms_in = s->offsets[(size_t)sd.lig1] - sd.atom_base;
        auto out_index = [&](int64_t c) {
            if (c >= s->n_chains) return (int64_t)replicas * s->offsets[(size_t)n_lig];
            const int64_t l = c / replicas, rep = c % replicas;
            return (int64_t)replicas * s->offsets[(size_t)l] + rep * (s->offsets[(size_t)l + 1] - s->offsets[(size_t)l]);
        };
        sd.out_base = out_index(sd.c0);
        sd.out_atoms = out_index(sd.c1) - sd.out_base;
        int64_t mx = 0; bool uniform = true;
        const int64_t nl_first = s->offsets[(size_t)sd.lig0 + 1] - s->offsets[(size_t)sd.lig0];
        for (int64_t l = sd.lig0; l < sd.lig1; ++l) {
            const int64_t nl = s->offsets[(size_t)l + 1] - s->offsets[(size_t)l];
            mx = std::max(mx, nl); uniform &= (nl == nl_first);
        }
        sd.nl_cap = even_cap(mx);
        s->devs.push_back(sd);
        ScreenDev &d = s->devs.back();
        int rc = DDD_OK;
        auto guard = [&](int code) { if (code && !rc) rc = code; return code; };
        cudaError_t ce = cudaSetDevice(g_devs[di].id);
        if (ce != cudaSuccess) guard(fail(DDD_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(ce)));
        const int64_t n_local = d.c1 - d.c0;
        if (!rc && (cudaEventCreate(&d.ev0) != cudaSuccess || cudaEventCreate(&d.ev1) != cudaSuccess)) guard(fail(DDD_ECUDA, "cudaEventCreate failed"));
        if (!rc) guard(dev_alloc(&d.d_offsets, (size_t)n_lig + 1));
        if (!rc) guard(dev_alloc(&d.d_lig, (size_t)d.n_atoms_in * 4));
        if (!rc) guard(dev_alloc(&d.d_out, (size_t)d.out_atoms * 4));
        if (!rc) guard(dev_alloc(&d.d_stats, (size_t)n_local));
        if (!rc) guard(dev_alloc(&d.d_queue, 64));
        if (!rc) {
            static_assert(sizeof(long long) == sizeof(int64_t), "");
            ce = cudaMemcpyAsync(d.d_offsets, s->offsets.data(), sizeof(int64_t) * ((size_t)n_lig + 1), cudaMemcpyHostToDevice, g_devs[di].stream);
            if (ce == cudaSuccess && d.n_atoms_in > 0)
                ce = cudaMemcpyAsync(d.d_lig, lig_xyzq + d.atom_base * 4, sizeof(double) * 4 * (size_t)d.n_atoms_in, cudaMemcpyHostToDevice, g_devs[di].stream);
            if (ce != cudaSuccess) guard(fail(DDD_ECUDA, "H2D copy of start poses: %s", cudaGetErrorString(ce)));
        }
        if (!rc && !uniform) {
            // longest ligands first: the hardware CTA scheduler then fills the tail with short chains
            std::vector<int> order((size_t)n_local);
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
                const int64_t lx = (d.c0 + x) / replicas, ly = (d.c0 + y) / replicas;
                return s->offsets[(size_t)lx + 1] - s->offsets[(size_t)lx] > s->offsets[(size_t)ly + 1] - s->offsets[(size_t)ly];
            });
            guard(dev_alloc(&d.d_order, order.size()));
            if (!rc) {
                ce = cudaMemcpyAsync(d.d_order, order.data(), sizeof(int) * order.size(), cudaMemcpyHostToDevice, g_devs[di].stream);
                if (ce == cudaSuccess) ce = cudaStreamSynchronize(g_devs[di].stream);      // `order` goes out of scope
                if (ce != cudaSuccess) guard(fail(DDD_ECUDA, "H2D copy of chain order: %s", cudaGetErrorString(ce)));
            }
        }
        if (!rc) {
            ce = cudaStreamSynchronize(g_devs[di].stream);   // s->offsets / caller buffers may go away
            if (ce != cudaSuccess) guard(fail(DDD_ECUDA, "sync after upload: %s", cudaGetErrorString(ce)));
        }
        if (rc) {
            for (auto &x : s->devs) free_screen_dev(x, true);
            delete s;
            return rc;
        }
    }
    *out = s;
    return DDD_OK;
}

int screen_run_impl(ddd_screen *s, int64_t steps, int32_t rotate, double temperature, const ddd_params *p,
                    uint64_t seed, const double *recorded, const int64_t *recorded_offsets, bool want_trace) {
    if (int rc = check_screen(s)) return rc;
    if (steps < 0) return fail(DDD_EINVAL, "negative step count");
    if (int rc = check_params(p)) return rc;
    if ((recorded == nullptr) != (recorded_offsets == nullptr) && s->n_chains > 0)
        return fail(DDD_EINVAL, "recorded and recorded_offsets must be given together");
    s->steps = steps;
    s->last_cut = p->cutoff > 0;
    const DevParams prm = make_dev_params(*p, rotate, temperature);
    for (auto &d : s->devs) {
        const Dev &dev = g_devs[d.di];
        CUDA_TRY(cudaSetDevice(dev.id));
        const int64_t n_local = d.c1 - d.c0;
        dev_free(d.d_trace); d.d_trace = nullptr;
        dev_free(d.d_rec); d.d_rec = nullptr;
        dev_free(d.d_rec_off); d.d_rec_off = nullptr;
        if (want_trace) {
            if (int rc = dev_alloc(&d.d_trace, (size_t)(n_local * steps))) return rc;
            CUDA_TRY(cudaMemsetAsync(d.d_trace, 0, std::max<size_t>(1, (size_t)(n_local * steps)), dev.stream));
        }
        if (recorded) {
            d.rec_base = recorded_offsets[d.c0];
            const int64_t n_rec = recorded_offsets[d.c1] - d.rec_base;
            if (n_rec < 0) return fail(DDD_EINVAL, "recorded_offsets not monotone");
            if (int rc = dev_alloc(&d.d_rec, (size_t)n_rec)) return rc;
            if (int rc = dev_alloc(&d.d_rec_off, (size_t)s->n_chains + 1)) return rc;
            CUDA_TRY(cudaMemcpyAsync(d.d_rec_off, recorded_offsets, sizeof(int64_t) * ((size_t)s->n_chains + 1), cudaMemcpyHostToDevice, dev.stream));
            if (n_rec > 0)
                CUDA_TRY(cudaMemcpyAsync(d.d_rec, recorded + d.rec_base, sizeof(double) * (size_t)n_rec, cudaMemcpyHostToDevice, dev.stream));
        }
        ChainArgs a;
        a.rec = s->rec->view(d.di);
        a.prm = prm;
        a.offsets = d.d_offsets; a.lig_xyzq = d.d_lig; a.atom_base = d.atom_base; a.order = d.d_order;
        a.chain_begin = d.c0; a.chain_id_base = s->chain_id_base; a.replicas = s->replicas; a.steps = steps; a.seed = seed;
        a.recorded = d.d_rec; a.rec_offsets = d.d_rec_off; a.rec_base = d.rec_base;
        a.out_xyzq = d.d_out; a.out_base = d.out_base; a.out_stats = d.d_stats; a.out_trace = d.d_trace;
        a.nl_cap = d.nl_cap;
        a.ref_xyzq = nullptr; a.out_rmsd = nullptr; a.out_best_xyzq = nullptr; a.out_best_e = nullptr;
        d.fused_rmsd = false;
        s->picked = false;
        const bool precise = p->precision == DDD_PRECISION_F64;
        const bool cut = p->cutoff > 0;
        const bool lj = p->lj_scale != 0.0;
        if (lj && (!s->has_lj || !s->rec->has_lj))
            return fail(DDD_EINVAL, "params.lj_scale != 0 needs ddd_receptor_set_lj and ddd_screen_set_lj");
        if (lj && s->threads_override > 0) return fail(DDD_EINVAL, "params.lj_scale needs the SM-persistent engine");
        const int n_groups4 = (int)((s->rec->n_units_s + 3) / 4);
        const int gsum_bytes = (!cut && n_groups4 >= 1 && n_groups4 <= kMaxLaneSumGroups) ? n_groups4 * 128 : 0;
        const int lj_bytes = lj ? d.nl_cap * 24 : 0;                          // per ligand atom: (sigma/2, sqrt eps) fp64 + packed fp32
        const int ulb = (cut ? 3 * kUnitCap * (int)sizeof(unsigned short) : 0) + gsum_bytes + lj_bytes;
        const bool fits_sm_engine = sm_smem_bytes(d.nl_cap, 1, ulb) <= dev.max_smem;      // very large ligands: one CTA per chain
        if (lj && !fits_sm_engine) return fail(DDD_ELIMIT, "params.lj_scale: the ligand does not fit one SM's shared memory");
        if (cut && (s->threads_override > 0 || !fits_sm_engine))
            return fail(DDD_EINVAL, "params.cutoff needs the SM-persistent engine (chain_threads 0 and a ligand that fits one SM's shared memory)");
        if (cut && s->rec->n_units_s > 32752) return fail(DDD_ELIMIT, "params.cutoff supports receptors up to 1048064 atoms");
        if (s->track_best && n_local > 0 && (s->threads_override > 0 || !fits_sm_engine))
            return fail(DDD_EINVAL, "best-ever tracking needs the SM-persistent engine (chain_threads 0 and a ligand that fits one SM's shared memory)");
        if (s->threads_override <= 0 && n_local > 0 && fits_sm_engine) {
            // SM-persistent engine: one 16-warp CTA per SM, K resident chains each (ddd_sm_chain.cuh)
            // its epilogue also leaves CalculateRMSD(final pose, reference pose) of every chain in d_rmsd (CompareRMSD :40-45)
            if (!d.d_rmsd) { if (int rc = dev_alloc(&d.d_rmsd, (size_t)n_local)) return rc; }
            a.ref_xyzq = d.d_ref ? d.d_ref : d.d_lig; a.out_rmsd = d.d_rmsd; d.fused_rmsd = true;
            if (s->track_best) {
                if (!d.d_best) { if (int rc = dev_alloc(&d.d_best, (size_t)d.out_atoms * 4)) return rc; }
                if (!d.d_best_e) { if (int rc = dev_alloc(&d.d_best_e, (size_t)n_local)) return rc; }
                a.out_best_xyzq = d.d_best; a.out_best_e = d.d_best_e;
            }
            SmArgs A;
            A.c = a;
            const int grid = (int)std::min<int64_t>(dev.sm_count, n_local);
            const int k_fit = (int)std::min<size_t>(kSmMaxSlots, (dev.max_smem - kCtaHeaderBytes) / (sm_smem_bytes(d.nl_cap, 1, ulb) - kCtaHeaderBytes));
            const int k_want = s->slots_override > 0 ? s->slots_override : pick_chain_slots(n_local, grid);
            A.slots = std::max(1, std::min(k_want, k_fit));
            A.queue = d.d_queue; A.n_local = n_local; A.first_dynamic = (long long)grid * A.slots;
            const int n = (int)s->rec->n;
            A.n_units = (n + 31) / 32;
            // work items: runs of 4*m units (128*m receptor atoms: whole A=4 warp tiles), m as small as the item-sum
            // array allows; receptors too small to give every warp such an item are cut into kSmWarps short items
            if (A.n_units >= 4 * kSmWarps) {
                const int quads = (A.n_units + 3) / 4;
#ifndef DDD_SM_ITEM_QUADS
#define DDD_SM_ITEM_QUADS 3
#endif
                A.upi = 4 * std::max(DDD_SM_ITEM_QUADS, (quads + kSmMaxItems - 1) / kSmMaxItems);
            } else A.upi = std::max(1, (A.n_units + kSmWarps - 1) / kSmWarps);
            A.n_items = (A.n_units + A.upi - 1) / A.upi;
            A.cut = cut ? 1 : 0; A.unit_list_bytes = ulb - gsum_bytes - lj_bytes; A.gsum_bytes = gsum_bytes; A.lj_bytes = lj_bytes; A.lig_lj = d.d_lig_lj; A.n_groups4 = n_groups4; A.out_units = nullptr;
            if (cut) {
                if (!d.d_units) { if (int rc = dev_alloc(&d.d_units, (size_t)n_local)) return rc; }
                A.out_units = d.d_units;
            }
            // chain segments (SmArgs::seg_steps): when a launch has more chains than slots the SMs would run dry up to one chain
            // duration apart; cut into ~32 segments per chain they run dry one segment apart.  A launch whose chains are all
            // resident at once has nothing to rebalance (a chain cannot run faster than its own serial path) -- unless the
            // chains differ in cost per step: with the cutoff extension a ligand buried in the receptor has several times the
            // units in reach of one at its surface (measured on config 4: busiest SM 1.34 x the lightest), and segments rotate
            // the chains through the SMs.
            long long seg = 0;
            if (s->segment_override > 0) seg = s->segment_override;
            else if (s->segment_override < 0 && (n_local > A.first_dynamic || (cut && n_local > grid))) seg = std::max<int64_t>(64, steps / 32);
            if (seg > 0) { long long p2 = 1; while (p2 < seg) p2 <<= 1; seg = p2; }
            if (seg >= steps || (double)n_local * (double)((steps + std::max<long long>(seg, 1) - 1) / std::max<long long>(seg, 1)) > 2.0e9) seg = 0;
            const long long n_seg = seg > 0 ? (steps + seg - 1) / seg : 1;
            A.seg_steps = seg; A.n_positions = n_local * n_seg; A.ring = nullptr; A.push_ctr = d.d_queue + 32;
            A.n_seg = (int)n_seg; A.seg_shift = 0;
            while ((1ll << A.seg_shift) < seg) ++A.seg_shift;
            d.seg_steps = seg;
            if (seg > 0) {
                const size_t need = (size_t)(n_local * (n_seg - 1));
                if (need > d.ring_cap) {
                    dev_free(d.d_ring); d.d_ring = nullptr; d.ring_cap = 0;
                    if (int rc = dev_alloc(&d.d_ring, need)) return rc;
                    d.ring_cap = need;
                }
                CUDA_TRY(cudaMemsetAsync(d.d_ring, 0, need * sizeof(int), dev.stream));
                A.ring = d.d_ring;
            }
            size_t smem = sm_smem_bytes(d.nl_cap, A.slots, ulb);
            A.box_off = 0; A.box_bytes = 0; A.gbox_count = 0;
            if (cut) {                                                       // the cell list's box tables ride in shared memory when they fit
                const int64_t n_groups = (s->rec->n_units_s + 3) / 4;
                const int64_t n_super = (n_groups + 7) / 8;
                const int64_t g_entries = 2 * ((n_groups + 1) / 2 * 2) + std::max<int64_t>(2 * n_super, 2), u_entries = 2 * s->rec->n_units_s;   // group + super-group boxes
                const size_t off = (smem + 15) & ~(size_t)15, bytes = (size_t)(g_entries + u_entries) * sizeof(float4);
                if (s->rec->n_units_s > 0 && off + bytes <= dev.max_smem) {
                    A.box_off = (int)off; A.box_bytes = (int)bytes; A.gbox_count = (int)g_entries;
                    smem = off + bytes;
                }
            }
            d.threads = kSmThreads; d.slots = A.slots;
            CUDA_TRY(cudaMemsetAsync(d.d_queue, 0, 64 * sizeof(int), dev.stream));
            CUDA_TRY(cudaEventRecord(d.ev0, dev.stream));
            // one instantiation per (precision, extension): the default kernel carries no cutoff / Lennard-Jones code
#define LAUNCH_SM(PR, MD)                                                                     \
            do { if (int rc = set_smem(mc_sm_kernel<PR, MD>, smem)) return rc;                 \
                 mc_sm_kernel<PR, MD><<<(unsigned)grid, kSmThreads, smem, dev.stream>>>(A); } while (0)
            const int mode = lj ? MODE_LJ : cut ? MODE_CUT : MODE_PLAIN;
            if (precise) { if (mode == MODE_LJ) LAUNCH_SM(true, MODE_LJ); else if (mode == MODE_CUT) LAUNCH_SM(true, MODE_CUT); else LAUNCH_SM(true, MODE_PLAIN); }
            else { if (mode == MODE_LJ) LAUNCH_SM(false, MODE_LJ); else if (mode == MODE_CUT) LAUNCH_SM(false, MODE_CUT); else LAUNCH_SM(false, MODE_PLAIN); }
#undef LAUNCH_SM
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaEventRecord(d.ev1, dev.stream));
            continue;
        }
        const int threads = s->threads_override > 0 ? s->threads_override : pick_chain_threads(n_local, dev.sm_count);
        const size_t smem = chain_smem_bytes(d.nl_cap, threads);
        d.threads = threads; d.slots = 0; d.seg_steps = 0;
        CUDA_TRY(cudaEventRecord(d.ev0, dev.stream));
#define LAUNCH_CHAIN(T, PR)                                                                   \
        do { if (int rc = set_smem(mc_chain_kernel<T, PR>, smem)) return rc;                   \
             mc_chain_kernel<T, PR><<<(unsigned)n_local, T, smem, dev.stream>>>(a); } while (0)
        if (precise) {
            switch (threads) { case 32: LAUNCH_CHAIN(32, true); break; case 64: LAUNCH_CHAIN(64, true); break;
                               case 256: LAUNCH_CHAIN(256, true); break; default: LAUNCH_CHAIN(128, true); }
        } else {
            switch (threads) { case 32: LAUNCH_CHAIN(32, false); break; case 64: LAUNCH_CHAIN(64, false); break;
                               case 256: LAUNCH_CHAIN(256, false); break; default: LAUNCH_CHAIN(128, false); }
        }
#undef LAUNCH_CHAIN
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(d.ev1, dev.stream));
    }
    s->launched = true;
    return DDD_OK;
}

// Waits for the screen's own chain kernels.  `lk` (nullable): the library lock of the calling entry point -- it is released
// while this thread blocks on the device, so that other threads can stage and launch their own batches meanwhile (they
// queue behind this kernel on the device's stream; a goroutine per ligand block no longer serialises on the host side).
int screen_sync_impl(ddd_screen *s, std::unique_lock<std::recursive_mutex> *lk = nullptr) {
    if (int rc0 = check_screen(s)) return rc0;
    for (auto &d : s->devs) {
        const Dev &dev = g_devs[d.di];
        CUDA_TRY(cudaSetDevice(dev.id));
        if (s->launched && d.ev1) {
            cudaEvent_t ev = d.ev1;
            if (lk) lk->unlock();
            const cudaError_t ce = cudaEventSynchronize(ev);
            if (lk) lk->lock();
            if (ce != cudaSuccess) return fail(DDD_ECUDA, "cudaEventSynchronize failed: %s", cudaGetErrorString(ce));
            if (int rc0 = check_screen(s)) return rc0;                    // the library was shut down while we waited
            CUDA_TRY(cudaEventElapsedTime(&d.last_ms, d.ev0, d.ev1));
        } else CUDA_TRY(cudaStreamSynchronize(dev.stream));
    }
    return DDD_OK;
}

int screen_download_impl(ddd_screen *s, double *out_xyzq, ddd_chain_stats *out_stats, uint8_t *out_trace,
                         std::unique_lock<std::recursive_mutex> *lk = nullptr) {
    if (int rc0 = check_screen(s)) return rc0;
    if (!s->launched) return fail(DDD_EINVAL, "screen has not been run");
    if (int rc = screen_sync_impl(s, lk)) return rc;
    for (auto &d : s->devs) {
        const Dev &dev = g_devs[d.di];
        CUDA_TRY(cudaSetDevice(dev.id));
        const int64_t n_local = d.c1 - d.c0;
        if (out_xyzq && d.out_atoms > 0)
            CUDA_TRY(cudaMemcpyAsync(out_xyzq + d.out_base * 4, d.d_out, sizeof(double) * 4 * (size_t)d.out_atoms, cudaMemcpyDeviceToHost, dev.stream));
        if (out_stats)
            CUDA_TRY(cudaMemcpyAsync(out_stats + d.c0, d.d_stats, sizeof(ChainStats) * (size_t)n_local, cudaMemcpyDeviceToHost, dev.stream));
        if (out_trace && d.d_trace && s->steps > 0)
            CUDA_TRY(cudaMemcpyAsync(out_trace + d.c0 * s->steps, d.d_trace, (size_t)(n_local * s->steps), cudaMemcpyDeviceToHost, dev.stream));
    }
    for (auto &d : s->devs) {
        CUDA_TRY(cudaSetDevice(g_devs[d.di].id));
        CUDA_TRY(cudaStreamSynchronize(g_devs[d.di].stream));
    }
    return DDD_OK;
}

void screen_destroy_impl(ddd_screen *s) {
    if (!s) return;
    for (auto &d : s->devs) free_screen_dev(d, s->epoch == g_epoch);
    delete s;
}

// frees every per-device copy on the device it was created on (valid after a ddd_shutdown too: the CUDA context persists)
void destroy_receptor(ddd_receptor *r) {
    for (size_t di = 0; di < r->f32.size() && di < r->dev_ids.size(); ++di) {
        cudaSetDevice(r->dev_ids[di]);
        cudaFree(r->f32[di]); cudaFree(r->f64[di]);
        if (di < r->f32s.size()) { cudaFree(r->f32s[di]); cudaFree(r->ubox[di]); cudaFree(r->f64s[di]); cudaFree(r->gbox[di]);
                                    cudaFree(r->lj64[di]); cudaFree(r->lj64s[di]); cudaFree(r->lj32s[di]); }
    }
    delete r;
}

// The resident pipeline keeps the start poses as uploaded as the RMSD reference (CompareRMSD :41 takes its CopyLigand before
// RandomizeLigandPose); the copy is made the first time randomise / shift is about to change d_lig.
int ensure_reference(ddd_screen *s) {
    for (auto &d : s->devs) {
        if (d.d_ref || d.n_atoms_in == 0) continue;
        const Dev &dev = g_devs[d.di];
        CUDA_TRY(cudaSetDevice(dev.id));
        if (int rc = dev_alloc(&d.d_ref, (size_t)d.n_atoms_in * 4)) return rc;
        CUDA_TRY(cudaMemcpyAsync(d.d_ref, d.d_lig, sizeof(double) * 4 * (size_t)d.n_atoms_in, cudaMemcpyDeviceToDevice, dev.stream));
    }
    return DDD_OK;
}

// Replica pick of SimulateEnergyMinimizationParallel (:102-109) + gather of the picked poses + their RMSD against the
// reference poses, per device, nothing leaves HBM.  Shards are ligand-aligned (plan_shards), so every ligand's replicas
// are on one device.
int screen_pick_impl(ddd_screen *s, double temperature, uint64_t seed, uint64_t first_ligand_index) {
    if (!s->launched) return fail(DDD_EINVAL, "screen has not been run");
    for (auto &d : s->devs) {
        const Dev &dev = g_devs[d.di];
        CUDA_TRY(cudaSetDevice(dev.id));
        const int64_t n_ll = d.lig1 - d.lig0;
        if (!d.d_min) { if (int rc = dev_alloc(&d.d_min, (size_t)d.n_atoms_in * 4)) return rc; }
        if (!d.d_min_e) { if (int rc = dev_alloc(&d.d_min_e, (size_t)n_ll)) return rc; }
        if (!d.d_picked) { if (int rc = dev_alloc(&d.d_picked, (size_t)n_ll)) return rc; }
        if (!d.d_min_rmsd) { if (int rc = dev_alloc(&d.d_min_rmsd, (size_t)n_ll)) return rc; }
        if (n_ll == 0) continue;
        replica_pick_kernel<<<(unsigned)((n_ll + 127) / 128), 128, 0, dev.stream>>>(d.d_stats, n_ll, d.lig0, d.c0, s->replicas, temperature, seed,
                                                                                 DDD_STREAM_PARENT_BASE + first_ligand_index, d.d_min_e, d.d_picked);
        CUDA_TRY(cudaGetLastError());
        pick_gather_kernel<<<(unsigned)((n_ll + 7) / 8), 256, 0, dev.stream>>>(d.d_offsets, d.d_lig, d.d_ref ? d.d_ref : d.d_lig, d.atom_base, d.d_out,
                                                                            d.out_base, s->replicas, d.d_picked, n_ll, d.lig0, d.d_min, d.d_min_rmsd);
        CUDA_TRY(cudaGetLastError());
    }
    s->picked = true;
    return DDD_OK;
}

int screen_download_pick(ddd_screen *s, double *out_xyzq, double *out_energy, int64_t *out_picked, double *out_rmsd) {
    for (auto &d : s->devs) {
        const Dev &dev = g_devs[d.di];
        CUDA_TRY(cudaSetDevice(dev.id));
        const int64_t n_ll = d.lig1 - d.lig0;
        if (out_xyzq && d.n_atoms_in > 0)
            CUDA_TRY(cudaMemcpyAsync(out_xyzq + d.atom_base * 4, d.d_min, sizeof(double) * 4 * (size_t)d.n_atoms_in, cudaMemcpyDeviceToHost, dev.stream));
        if (out_energy && n_ll > 0) CUDA_TRY(cudaMemcpyAsync(out_energy + d.lig0, d.d_min_e, sizeof(double) * (size_t)n_ll, cudaMemcpyDeviceToHost, dev.stream));
        if (out_picked && n_ll > 0) CUDA_TRY(cudaMemcpyAsync(out_picked + d.lig0, d.d_picked, sizeof(int64_t) * (size_t)n_ll, cudaMemcpyDeviceToHost, dev.stream));
        if (out_rmsd && n_ll > 0) CUDA_TRY(cudaMemcpyAsync(out_rmsd + d.lig0, d.d_min_rmsd, sizeof(double) * (size_t)n_ll, cudaMemcpyDeviceToHost, dev.stream));
    }
    for (auto &d : s->devs) {
        CUDA_TRY(cudaSetDevice(g_devs[d.di].id));
        CUDA_TRY(cudaStreamSynchronize(g_devs[d.di].stream));
    }
    return DDD_OK;
}

// contiguous ligand blocks for the one-shot batch entry points (energy / RMSD / closest / shift / randomise), one per
// device in use, balanced by atoms; block k = ligands [b[k], b[k+1])
std::vector<int64_t> split_ligands(const int64_t *offsets, int64_t n_lig) {
    const int nd = (int)g_devs.size();
    std::vector<int64_t> b((size_t)nd + 1, 0);
    b[(size_t)nd] = n_lig;
    // small batches stay on the first device: a second device costs more in launch + sync than it saves
    if (nd <= 1 || n_lig < 64 * nd) { for (int k = 1; k < nd; ++k) b[(size_t)k] = n_lig; return b; }
    int64_t lig = 0;
    const double total = (double)(offsets[n_lig] + n_lig);
    for (int k = 1; k < nd; ++k) {
        const double target = total * (double)k / (double)nd;
        while (lig < n_lig && (double)(offsets[lig + 1] + lig + 1) <= target) ++lig;
        b[(size_t)k] = lig;
    }
    return b;
}

}  // namespace

// ------------------------------------------------------------------ library
extern "C" {

int ddd_init(const int *devices, int n_devices) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count <= 0) {
        (void)cudaGetLastError();
        return fail(DDD_ENODEVICE, "no CUDA device available (%s); libddd_mc has no CPU fallback",
                    ce == cudaSuccess ? "device count is 0" : cudaGetErrorString(ce));
    }
    std::vector<int> ids;
    if (!devices || n_devices <= 0) { for (int i = 0; i < count; ++i) ids.push_back(i); }
    else {
        for (int i = 0; i < n_devices; ++i) {
            if (devices[i] < 0 || devices[i] >= count) return fail(DDD_EINVAL, "device ordinal %d out of range [0,%d)", devices[i], count);
            if (std::find(ids.begin(), ids.end(), devices[i]) != ids.end()) return fail(DDD_EINVAL, "device ordinal %d listed twice", devices[i]);
            ids.push_back(devices[i]);
        }
    }
    // Re-initialising with the device list already in use is a no-op (callers may initialise lazily from many entry
    // points).  A different list shuts the library down first: handles of the old list turn stale (g_epoch).
    bool same = ids.size() == g_devs.size();
    for (size_t i = 0; same && i < ids.size(); ++i) same = g_devs[i].id == ids[i];
    if (same && !g_devs.empty()) return DDD_OK;
    ddd_shutdown();                                                       // receptors / screens of the old list become stale handles
    for (int id : ids) {
        Dev d; d.id = id;
        cudaDeviceProp prop;
        auto bail = [&](int rc) { g_devs.push_back(d); const std::string msg = g_err; ddd_shutdown(); g_err = msg; return rc; };   // destroys what was created so far
        if (cudaGetDeviceProperties(&prop, id) != cudaSuccess) return bail(fail(DDD_ECUDA, "cudaGetDeviceProperties(%d) failed", id));
        // the library ships sm_100a SASS only; architecture-specific ("a") code runs on exactly that compute capability
        if (prop.major != 10 || prop.minor != 0)
            return bail(fail(DDD_ENODEVICE, "device %d (%s, sm_%d%d) is not sm_100 (B200): this library ships sm_100a code only", id, prop.name, prop.major, prop.minor));
        d.sm_count = prop.multiProcessorCount;
        d.max_smem = prop.sharedMemPerBlockOptin;
        cudaMemPool_t pool = nullptr;
        unsigned long long keep = ~0ull;                                  // never hand pool memory back between calls
        if ((ce = cudaSetDevice(id)) != cudaSuccess || (ce = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking)) != cudaSuccess ||
            (ce = cudaDeviceGetDefaultMemPool(&pool, id)) != cudaSuccess ||
            (ce = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep)) != cudaSuccess ||
            (ce = cudaEventCreate(&d.ev0)) != cudaSuccess || (ce = cudaEventCreate(&d.ev1)) != cudaSuccess)
            return bail(fail(DDD_ECUDA, "device %d setup failed: %s", id, cudaGetErrorString(ce)));
        g_devs.push_back(d);
    }
    return DDD_OK;
}

int ddd_shutdown(void) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!g_devs.empty()) ++g_epoch;
    purge_cache(true);
    for (auto &d : g_devs) {
        cudaSetDevice(d.id);
        if (d.ev0) cudaEventDestroy(d.ev0);
        if (d.ev1) cudaEventDestroy(d.ev1);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    g_devs.clear();
    return DDD_OK;
}

int ddd_num_devices(void) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    return (int)g_devs.size();
}

const char *ddd_last_error(void) { return g_err.c_str(); }

const char *ddd_version(void) { return "ddd_mc 0.4 (sm_100a; SM-persistent chain engine: 16 warps per SM, speculative proposals, per-group pair sums, chain segments; resident randomise/shift/pick/RMSD/argmin pipeline)"; }

void ddd_params_default(ddd_params *p) {
    if (!p) return;
    p->k_coulomb = 9e9;            // datatypes.go:6
    p->threshold = 5.0;            // datatypes.go:9
    p->min_distance = 0.5;         // datatypes.go:10
    p->max_angle = 0.79;           // datatypes.go:11
    p->jitter_width = 0.1;         // metropolis.go:158
    p->clamp_distance = 1e-6;      // metropolis.go:255
    p->rigid_translation = 0.5;
    p->rigid_angle = 0.79;
    p->max_retries = 1024;
    p->precision = DDD_PRECISION_F32;
    p->move_mode = DDD_MOVE_REFERENCE;
    p->cutoff = 0.0;               // extension, off: every pair is summed (metropolis.go:249-259)
    p->lj_scale = 0.0;             // extension, off: Coulomb only (metropolis.go:258)
    p->lj_cap = 4.0;
}

// pure host logic, usable without a device: chains [out_begin[k], out_begin[k+1]) -> shard k
int ddd_shard_plan(const int64_t *offsets, int64_t n_lig, int32_t replicas, int32_t n_shards, int64_t *out_begin) {
    if (n_shards <= 0 || replicas <= 0 || !out_begin) return fail(DDD_EINVAL, "bad shard plan arguments");
    if (int rc = check_csr(offsets, n_lig, (const void *)1, nullptr)) return rc;
    std::vector<int64_t> b;
    const int64_t zero = 0;
    plan_shards(n_lig > 0 ? offsets : &zero, n_lig, replicas, n_shards, b);
    std::copy(b.begin(), b.end(), out_begin);
    return DDD_OK;
}

// ------------------------------------------------------------------ receptor
int ddd_receptor_create(const double *xyzq, int64_t n_atoms, ddd_receptor **out) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (!out) return fail(DDD_EINVAL, "out is NULL");
    if (n_atoms < 0 || (n_atoms > 0 && !xyzq)) return fail(DDD_EINVAL, "bad receptor array");
    if (n_atoms > 0x7fffff00ll) return fail(DDD_ELIMIT, "receptor too large");
    ddd_receptor *r = new ddd_receptor();
    r->epoch = g_epoch;
    for (auto &d : g_devs) r->dev_ids.push_back(d.id);
    r->n = n_atoms;
    r->n_pad = (n_atoms + kRecPad - 1) / kRecPad * kRecPad;
    double sx = 0, sy = 0, sz = 0;
    for (int64_t i = 0; i < n_atoms; ++i) { sx += xyzq[4 * i]; sy += xyzq[4 * i + 1]; sz += xyzq[4 * i + 2]; }
    if (n_atoms > 0) { r->cx = sx / (double)n_atoms; r->cy = sy / (double)n_atoms; r->cz = sz / (double)n_atoms; }
    std::vector<float4> h32((size_t)r->n_pad);
    for (int64_t i = 0; i < r->n_pad; ++i) {
        if (i < n_atoms)
            h32[(size_t)i] = make_float4((float)(xyzq[4 * i] - r->cx), (float)(xyzq[4 * i + 1] - r->cy),
                                         (float)(xyzq[4 * i + 2] - r->cz), (float)xyzq[4 * i + 3]);
        else h32[(size_t)i] = make_float4(kPadCoord, kPadCoord, kPadCoord, 0.f);
    }
    // Cell list for the cutoff extension: atoms sorted along the Morton curve of a 6 A grid (ties: input order), cut
    // into units of 32 consecutive atoms, each with the box of its real atoms (padding atoms are parked far away, q = 0)
    std::vector<float4> h32s;
    std::vector<float4> hbox, hgbox;
    std::vector<Atom64> h64s;
    {
        const float cell = 6.0f;
        float mn[3] = {0, 0, 0};
        for (int64_t i = 0; i < n_atoms; ++i) {
            const float v[3] = {h32[(size_t)i].x, h32[(size_t)i].y, h32[(size_t)i].z};
            for (int d = 0; d < 3; ++d) mn[d] = i == 0 ? v[d] : std::min(mn[d], v[d]);
        }
        auto spread = [](uint64_t x) {       // 21 bits -> every third bit
            x &= 0x1fffff; x = (x | x << 32) & 0x1f00000000ffffull; x = (x | x << 16) & 0x1f0000ff0000ffull;
            x = (x | x << 8) & 0x100f00f00f00f00full; x = (x | x << 4) & 0x10c30c30c30c30c3ull; x = (x | x << 2) & 0x1249249249249249ull;
            return x;
        };
        std::vector<std::pair<uint64_t, int64_t>> key((size_t)n_atoms);
        for (int64_t i = 0; i < n_atoms; ++i) {
            const float4 a = h32[(size_t)i];
            const uint64_t ix = (uint64_t)std::max(0.f, (a.x - mn[0]) / cell), iy = (uint64_t)std::max(0.f, (a.y - mn[1]) / cell),
                           iz = (uint64_t)std::max(0.f, (a.z - mn[2]) / cell);
            key[(size_t)i] = {spread(ix) | spread(iy) << 1 | spread(iz) << 2, i};
        }
        std::sort(key.begin(), key.end());
        r->n_units_s = (n_atoms + 31) / 32;
        const int64_t n_pad_s = sorted_pad_atoms(r->n_units_s);             // whole groups of 8 units may be read, + the filler unit
        h32s.assign((size_t)std::max<int64_t>(n_pad_s, 1), make_float4(kPadCoord, kPadCoord, kPadCoord, 0.f));
        hbox.assign((size_t)std::max<int64_t>(2 * r->n_units_s, 1), make_float4(0, 0, 0, 0));
        h64s.resize((size_t)std::max<int64_t>(n_atoms, 1));
        r->perm.resize((size_t)n_atoms);
        for (int64_t i = 0; i < n_atoms; ++i) {
            const int64_t src = key[(size_t)i].second;
            r->perm[(size_t)i] = src;
            h32s[(size_t)i] = h32[(size_t)src];
            h64s[(size_t)i] = Atom64{xyzq[4 * src], xyzq[4 * src + 1], xyzq[4 * src + 2], xyzq[4 * src + 3]};
        }
        for (int64_t u = 0; u < r->n_units_s; ++u) {
            float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
            for (int64_t i = 32 * u; i < std::min<int64_t>(32 * u + 32, n_atoms); ++i) {
                const float v[3] = {h32s[(size_t)i].x, h32s[(size_t)i].y, h32s[(size_t)i].z};
                for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], v[d]); hi[d] = std::max(hi[d], v[d]); }
            }
            hbox[(size_t)(2 * u)] = make_float4(lo[0], lo[1], lo[2], 0.f);
            hbox[(size_t)(2 * u + 1)] = make_float4(hi[0], hi[1], hi[2], 0.f);
        }
        const int64_t n_groups = (r->n_units_s + 3) / 4;
        // boxes of 4-unit groups, padded to an even count with far-away boxes (a group of padding atoms only)
        hgbox.assign((size_t)std::max<int64_t>(2 * ((n_groups + 1) / 2 * 2), 2), make_float4(kPadCoord, kPadCoord, kPadCoord, 0.f));
        for (int64_t g = 0; g < n_groups; ++g) {
            float4 lo = make_float4(3e38f, 3e38f, 3e38f, 0.f), hi = make_float4(-3e38f, -3e38f, -3e38f, 0.f);
            for (int64_t u = 4 * g; u < std::min<int64_t>(4 * g + 4, r->n_units_s); ++u) {
                const float4 a = hbox[(size_t)(2 * u)], b = hbox[(size_t)(2 * u + 1)];
                lo.x = std::min(lo.x, a.x); lo.y = std::min(lo.y, a.y); lo.z = std::min(lo.z, a.z);
                hi.x = std::max(hi.x, b.x); hi.y = std::max(hi.y, b.y); hi.z = std::max(hi.z, b.z);
            }
            hgbox[(size_t)(2 * g)] = lo; hgbox[(size_t)(2 * g + 1)] = hi;
        }
        // behind them, the boxes of the super-groups of 8 groups (32 units): the top level of the cell-list query
        const int64_t n_super = (n_groups + 7) / 8;
        const size_t g_pad = hgbox.size();
        hgbox.resize(g_pad + (size_t)std::max<int64_t>(2 * n_super, 2), make_float4(kPadCoord, kPadCoord, kPadCoord, 0.f));
        for (int64_t sg = 0; sg < n_super; ++sg) {
            float4 lo = make_float4(3e38f, 3e38f, 3e38f, 0.f), hi = make_float4(-3e38f, -3e38f, -3e38f, 0.f);
            for (int64_t g = 8 * sg; g < std::min<int64_t>(8 * sg + 8, n_groups); ++g) {
                const float4 a = hgbox[(size_t)(2 * g)], b = hgbox[(size_t)(2 * g + 1)];
                lo.x = std::min(lo.x, a.x); lo.y = std::min(lo.y, a.y); lo.z = std::min(lo.z, a.z);
                hi.x = std::max(hi.x, b.x); hi.y = std::max(hi.y, b.y); hi.z = std::max(hi.z, b.z);
            }
            hgbox[g_pad + (size_t)(2 * sg)] = lo; hgbox[g_pad + (size_t)(2 * sg + 1)] = hi;
        }
    }
    r->f32.assign(g_devs.size(), nullptr);
    r->f64.assign(g_devs.size(), nullptr);
    r->f32s.assign(g_devs.size(), nullptr);
    r->ubox.assign(g_devs.size(), nullptr);
    r->f64s.assign(g_devs.size(), nullptr);
    r->gbox.assign(g_devs.size(), nullptr);
    r->lj64.assign(g_devs.size(), nullptr); r->lj64s.assign(g_devs.size(), nullptr); r->lj32s.assign(g_devs.size(), nullptr);
    int rc = DDD_OK;
    for (size_t di = 0; di < g_devs.size() && !rc; ++di) {
        cudaError_t ce = cudaSetDevice(g_devs[di].id);
        if (ce == cudaSuccess) { rc = dev_alloc_sync(&r->f32[di], (size_t)r->n_pad); if (rc) break; rc = dev_alloc_sync(&r->f64[di], (size_t)n_atoms); if (rc) break; }
        if (ce == cudaSuccess && r->n_pad > 0) ce = cudaMemcpy(r->f32[di], h32.data(), sizeof(float4) * (size_t)r->n_pad, cudaMemcpyHostToDevice);
        if (ce == cudaSuccess && n_atoms > 0) ce = cudaMemcpy(r->f64[di], xyzq, sizeof(Atom64) * (size_t)n_atoms, cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) { rc = dev_alloc_sync(&r->f32s[di], h32s.size()); if (rc) break; rc = dev_alloc_sync(&r->ubox[di], hbox.size()); if (rc) break; }
        if (ce == cudaSuccess) ce = cudaMemcpy(r->f32s[di], h32s.data(), sizeof(float4) * h32s.size(), cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) ce = cudaMemcpy(r->ubox[di], hbox.data(), sizeof(float4) * hbox.size(), cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) { rc = dev_alloc_sync(&r->gbox[di], hgbox.size()); if (rc) break; }
        if (ce == cudaSuccess) ce = cudaMemcpy(r->gbox[di], hgbox.data(), sizeof(float4) * hgbox.size(), cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) { rc = dev_alloc_sync(&r->f64s[di], h64s.size()); if (rc) break; }
        if (ce == cudaSuccess) ce = cudaMemcpy(r->f64s[di], h64s.data(), sizeof(Atom64) * h64s.size(), cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) rc = fail(DDD_ECUDA, "receptor upload on device %d: %s", g_devs[di].id, cudaGetErrorString(ce));
    }
    if (rc) { destroy_receptor(r); return rc; }
    *out = r;
    return DDD_OK;
}

int ddd_receptor_destroy(ddd_receptor *r) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!r) return DDD_OK;
    if (r->cached) return fail(DDD_EINVAL, "this receptor belongs to the cache: use ddd_receptor_release");
    destroy_receptor(r);
    return DDD_OK;
}

// Receptor cache for the scalar entry points of the host mirrors (CalculateEnergy, FindClosestAtomDistance, Shift...,
// SimulateEnergyMinimizationParallel called per ligand by RShinyAppMain, main.go:42-48): the same protein arrives on every
// call, and building a receptor costs a host sort + 9 allocations + synchronous uploads.  Entries are keyed by the content
// (length + 64-bit FNV-1a of the bytes, compared in full on a hit), reference-counted, and evicted least-recently-used.
extern "C++" {
namespace {
constexpr size_t kReceptorCacheEntries = 8;
struct CacheEntry { ddd_receptor *r; std::vector<double> copy; uint64_t last_use; };
std::vector<CacheEntry> g_cache;
uint64_t g_cache_clock = 0;
uint64_t fnv1a(const void *p, size_t n) {
    const unsigned char *b = (const unsigned char *)p;
    uint64_t h = 1469598103934665603ull;
    // 8 bytes at a time: this runs on every scalar call
    size_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t w; std::memcpy(&w, b + i, 8); h = (h ^ w) * 1099511628211ull; }
    for (; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
    return h;
}
void purge_cache(bool everything) {
    for (size_t i = 0; i < g_cache.size();) {
        ddd_receptor *r = g_cache[i].r;
        if (everything || r->epoch != g_epoch) {
            if (r->refs <= 0) destroy_receptor(r);
            else r->cached = false;                     // still referenced: the last ddd_receptor_release frees it
            g_cache.erase(g_cache.begin() + (long)i);
        } else ++i;
    }
}
}  // namespace
}  // extern "C++"

int ddd_receptor_acquire(const double *xyzq, int64_t n_atoms, ddd_receptor **out) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (!out) return fail(DDD_EINVAL, "out is NULL");
    if (n_atoms < 0 || (n_atoms > 0 && !xyzq)) return fail(DDD_EINVAL, "bad receptor array");
    purge_cache(false);
    const size_t bytes = (size_t)n_atoms * 32;
    const uint64_t h = fnv1a(xyzq, bytes);
    for (auto &e : g_cache) {
        if (e.r->n == n_atoms && e.r->hash == h && (bytes == 0 || std::memcmp(e.copy.data(), xyzq, bytes) == 0)) {
            e.last_use = ++g_cache_clock; e.r->refs += 1; *out = e.r;
            return DDD_OK;
        }
    }
    ddd_receptor *r = nullptr;
    if (int rc = ddd_receptor_create(xyzq, n_atoms, &r)) return rc;
    r->hash = h; r->cached = true; r->refs = 1;
    if (g_cache.size() >= kReceptorCacheEntries) {      // evict the least recently used entry nobody holds
        size_t victim = g_cache.size();
        for (size_t i = 0; i < g_cache.size(); ++i)
            if (g_cache[i].r->refs <= 0 && (victim == g_cache.size() || g_cache[i].last_use < g_cache[victim].last_use)) victim = i;
        if (victim < g_cache.size()) { destroy_receptor(g_cache[victim].r); g_cache.erase(g_cache.begin() + (long)victim); }
    }
    CacheEntry e; e.r = r; e.copy.assign(xyzq, xyzq + n_atoms * 4); e.last_use = ++g_cache_clock;
    g_cache.push_back(std::move(e));
    *out = r;
    return DDD_OK;
}

int ddd_receptor_release(ddd_receptor *r) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!r) return DDD_OK;
    r->refs -= 1;
    if (!r->cached && r->refs <= 0) destroy_receptor(r);   // dropped from the cache (shutdown) while in use
    return DDD_OK;
}

int64_t ddd_receptor_num_atoms(const ddd_receptor *r) { return r ? r->n : -1; }

int ddd_receptor_set_lj(ddd_receptor *r, const double *sigma, const double *epsilon) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (int rc0 = check_receptor(r)) return rc0;
    if ((sigma == nullptr) != (epsilon == nullptr)) return fail(DDD_EINVAL, "sigma and epsilon must be given together");
    if (!sigma) { r->has_lj = false; return DDD_OK; }
    const int64_t n = r->n;
    const size_t n_pad = (size_t)sorted_pad_atoms(r->n_units_s);
    std::vector<double2> a((size_t)std::max<int64_t>(n, 1)), as((size_t)std::max<int64_t>(n, 1));
    std::vector<float2> fs(std::max<size_t>(n_pad, 1), make_float2(0.f, 0.f));          // padding atoms: epsilon = 0
    for (int64_t i = 0; i < n; ++i) {
        if (!(sigma[i] >= 0.0) || !(epsilon[i] >= 0.0)) return fail(DDD_EINVAL, "sigma/epsilon of receptor atom %lld is negative or NaN", (long long)i);
        a[(size_t)i] = make_double2(0.5 * sigma[i], std::sqrt(epsilon[i]));
    }
    for (int64_t i = 0; i < n; ++i) {
        as[(size_t)i] = a[(size_t)r->perm[(size_t)i]];
        fs[(size_t)i] = make_float2((float)as[(size_t)i].x, (float)as[(size_t)i].y);
    }
    for (size_t di = 0; di < g_devs.size(); ++di) {
        CUDA_TRY(cudaSetDevice(g_devs[di].id));
        if (!r->lj64[di]) {
            if (int rc = dev_alloc_sync(&r->lj64[di], a.size())) return rc;
            if (int rc = dev_alloc_sync(&r->lj64s[di], as.size())) return rc;
            if (int rc = dev_alloc_sync(&r->lj32s[di], fs.size())) return rc;
        }
        CUDA_TRY(cudaMemcpy(r->lj64[di], a.data(), sizeof(double2) * a.size(), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(r->lj64s[di], as.data(), sizeof(double2) * as.size(), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(r->lj32s[di], fs.data(), sizeof(float2) * fs.size(), cudaMemcpyHostToDevice));
    }
    r->has_lj = true;
    return DDD_OK;
}

// ------------------------------------------------------------------ CalculateEnergy batch
// One-shot batch entry points: the ligand list is cut into one contiguous block per device in use (split_ligands); every
// block is uploaded, processed and downloaded asynchronously on its device's stream, then all streams are joined.
extern "C++" {
namespace {
struct Scratch {                                 // per-call device buffers, freed (stream-ordered) on the device they live on
    std::vector<std::pair<int, void *>> bufs;
    template <typename T> int get(int dev_id, T **p, size_t count) {
        if (int rc = dev_alloc(p, count)) return rc;
        bufs.emplace_back(dev_id, (void *)*p);
        return DDD_OK;
    }
    ~Scratch() { for (auto &b : bufs) { cudaSetDevice(b.first); dev_free(b.second); } }
};
// the downloads run after EVERY device has its work queued (a D2H copy into pageable memory blocks the host until the
// device gets there), then all streams are joined
int join_devices(std::vector<std::function<int()>> &downloads) {
    for (auto &f : downloads) if (int rc = f()) return rc;
    for (auto &d : g_devs) { CUDA_TRY(cudaSetDevice(d.id)); CUDA_TRY(cudaStreamSynchronize(d.stream)); }
    return DDD_OK;
}
}  // namespace
}  // extern "C++"

int ddd_energy_batch(const ddd_receptor *r, const int64_t *offsets, const double *lig_xyzq, int64_t n_lig,
                     const ddd_params *p, double *out_energy, double *out_abs_sum) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (int rc0 = check_receptor(r)) return rc0;
    if (int rc = check_params(p)) return rc;
    if (p->lj_scale != 0.0) return fail(DDD_EINVAL, "params.lj_scale: ligand LJ parameters travel with a screen (run it with 0 steps for energies)");
    int64_t max_nl = 0;
    if (int rc = check_csr(offsets, n_lig, lig_xyzq, &max_nl)) return rc;
    if (n_lig == 0) return DDD_OK;
    if (!out_energy) return fail(DDD_EINVAL, "out_energy is NULL");
    Scratch sc;
    std::vector<std::function<int()>> later;
    const std::vector<int64_t> blk = split_ligands(offsets, n_lig);
    for (size_t di = 0; di < g_devs.size(); ++di) {
        const int64_t l0 = blk[di], l1 = blk[di + 1], nl_blk = l1 - l0;
        if (nl_blk <= 0) continue;
        const Dev &dev = g_devs[di];
        CUDA_TRY(cudaSetDevice(dev.id));
        const int64_t a0 = offsets[l0], na = offsets[l1] - a0;
        long long *d_off = nullptr; double *d_lig = nullptr, *d_e = nullptr, *d_abs = nullptr;
        if (int rc = sc.get(dev.id, &d_off, (size_t)nl_blk + 1)) return rc;
        if (int rc = sc.get(dev.id, &d_lig, (size_t)na * 4)) return rc;
        if (int rc = sc.get(dev.id, &d_e, (size_t)nl_blk)) return rc;
        if (out_abs_sum) { if (int rc = sc.get(dev.id, &d_abs, (size_t)nl_blk)) return rc; }
        CUDA_TRY(cudaMemcpyAsync(d_off, offsets + l0, sizeof(int64_t) * ((size_t)nl_blk + 1), cudaMemcpyHostToDevice, dev.stream));
        if (na > 0) CUDA_TRY(cudaMemcpyAsync(d_lig, lig_xyzq + a0 * 4, sizeof(double) * 4 * (size_t)na, cudaMemcpyHostToDevice, dev.stream));
        EnergyArgs a;
        a.rec = r->view(di); a.prm = make_dev_params(*p, 0, 1.0);
        a.offsets = d_off; a.atom_base = a0; a.lig_xyzq = d_lig; a.out_energy = d_e; a.out_abs = d_abs; a.nl_cap = even_cap(max_nl);
        const size_t smem = chain_smem_bytes(a.nl_cap, kThreads);
        const bool precise = p->precision == DDD_PRECISION_F64, with_abs = out_abs_sum != nullptr;
        const unsigned grid = (unsigned)nl_blk;
#define LAUNCH_E(PR, AB)                                                                   \
        do { if (int rc2 = set_smem(energy_batch_kernel<PR, AB>, smem)) return rc2;        \
             energy_batch_kernel<PR, AB><<<grid, kThreads, smem, dev.stream>>>(a); } while (0)
        if (precise && with_abs) LAUNCH_E(true, true);
        else if (precise) LAUNCH_E(true, false);
        else if (with_abs) LAUNCH_E(false, true);
        else LAUNCH_E(false, false);
#undef LAUNCH_E
        CUDA_TRY(cudaGetLastError());
        later.push_back([=]() -> int {
            CUDA_TRY(cudaSetDevice(dev.id));
            CUDA_TRY(cudaMemcpyAsync(out_energy + l0, d_e, sizeof(double) * (size_t)nl_blk, cudaMemcpyDeviceToHost, dev.stream));
            if (with_abs) CUDA_TRY(cudaMemcpyAsync(out_abs_sum + l0, d_abs, sizeof(double) * (size_t)nl_blk, cudaMemcpyDeviceToHost, dev.stream));
            return DDD_OK;
        });
    }
    return join_devices(later);
}

// ------------------------------------------------------------------ chains
int ddd_run_chains(const ddd_receptor *r, const int64_t *offsets, const double *lig_xyzq, int64_t n_lig,
                   int32_t replicas, int64_t steps_per_chain, int32_t rotate, double temperature,
                   const ddd_params *p, uint64_t seed, uint64_t first_chain_id,
                   const double *recorded, const int64_t *recorded_offsets,
                   double *out_xyzq, ddd_chain_stats *out_stats, uint8_t *out_accept_trace) {
    std::unique_lock<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (int rc = check_params(p)) return rc;
    if (p->lj_scale != 0.0) return fail(DDD_EINVAL, "params.lj_scale: ligand LJ parameters travel with a screen (ddd_screen_create + ddd_screen_set_lj)");
    ddd_screen *s = nullptr;
    if (int rc = screen_create_impl(r, offsets, lig_xyzq, n_lig, replicas, first_chain_id, &s)) return rc;
    int rc = screen_run_impl(s, steps_per_chain, rotate, temperature, p, seed, recorded, recorded_offsets, out_accept_trace != nullptr);
    if (!rc) rc = screen_download_impl(s, out_xyzq, out_stats, out_accept_trace, &lk);
    screen_destroy_impl(s);
    g_status_bits = 0; g_status_chains = 0;
    if (!rc && out_stats)
        for (int64_t c = 0; c < n_lig * replicas; ++c) { g_status_bits |= out_stats[c].status; g_status_chains += (out_stats[c].status & DDD_STATUS_RETRY_LIMIT) ? 1 : 0; }
    return rc;
}

int ddd_minimize_batch(const ddd_receptor *r, const int64_t *offsets, const double *lig_xyzq, int64_t n_lig,
                       int64_t iterations, int32_t rotate, double temperature, int32_t num_procs,
                       const ddd_params *p, uint64_t seed, uint64_t first_ligand_index,
                       double *out_xyzq, double *out_energy, ddd_chain_stats *out_chain_stats, int64_t *out_picked) {
    std::unique_lock<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    g_status_bits = 0; g_status_chains = 0;
    if (num_procs <= 0) return fail(DDD_EINVAL, "num_procs must be >= 1 (the reference divides iterations by it, metropolis.go:97)");
    if (iterations < 0) return fail(DDD_EINVAL, "negative iteration count");
    if (int rc = check_csr(offsets, n_lig, lig_xyzq, nullptr)) return rc;
    if (n_lig == 0) return DDD_OK;
    if (!out_xyzq && offsets[n_lig] > 0) return fail(DDD_EINVAL, "out_xyzq is NULL");
    if (!out_energy) return fail(DDD_EINVAL, "out_energy is NULL");
    const int64_t width = iterations / num_procs;                                  // :97
    const int64_t n_chains = n_lig * num_procs;
    std::vector<ddd_chain_stats> stats((size_t)n_chains);
    // chain (ligand i, replica r) draws from stream (first_ligand_index + i) * num_procs + r, so a
    // ligand list split over several calls (or ranks) sees the same streams as one call
    ddd_screen *s = nullptr;
    if (int rc = check_params(p)) return rc;
    const bool dbg = std::getenv("DDD_DEBUG_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    if (int rc = screen_create_impl(r, offsets, lig_xyzq, n_lig, num_procs, first_ligand_index * (uint64_t)num_procs, &s)) return rc;
    const double t1 = now();
    int rc = screen_run_impl(s, width, rotate, temperature, p, seed, nullptr, nullptr, false);
    // the replica pick (:102-109) and the gather of the picked poses run on the device behind the chains: only the
    // per-ligand results (and the 72-byte chain records) come back
    if (!rc) rc = screen_pick_impl(s, temperature, seed, first_ligand_index);
    const double t2 = now();
    if (!rc) rc = screen_download_impl(s, nullptr, stats.data(), nullptr, &lk);
    if (!rc) rc = screen_download_pick(s, out_xyzq, out_energy, out_picked, nullptr);
    const double t3 = now();
    const double kms = ddd_screen_last_kernel_ms(s);
    screen_destroy_impl(s);
    const double t4 = now();
    if (dbg) fprintf(stderr, "[ddd_minimize_batch] create %.2f ms, launch %.2f ms, sync+download %.2f ms (kernel %.2f ms), destroy %.2f ms\n",
                     t1 - t0, t2 - t1, t3 - t2, kms, t4 - t3);
    if (rc) return rc;
    for (auto &cs : stats) { g_status_bits |= cs.status; g_status_chains += (cs.status & DDD_STATUS_RETRY_LIMIT) ? 1 : 0; }
    if (out_chain_stats) std::memcpy(out_chain_stats, stats.data(), sizeof(ddd_chain_stats) * stats.size());
    return DDD_OK;
}

int64_t ddd_last_chain_status(int64_t *out_retry_limited_chains) {
    if (out_retry_limited_chains) *out_retry_limited_chains = g_status_chains;
    return g_status_bits;
}

// ------------------------------------------------------------------ RMSD / closest / shift / randomize (one block per device)
int ddd_rmsd_batch(const int64_t *offsets, const double *sim, const double *ref, int32_t atom_stride,
                   int64_t n_pairs, double *out_rmsd) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (atom_stride != 3 && atom_stride != 4) return fail(DDD_EINVAL, "atom_stride must be 3 or 4");
    if (n_pairs < 0) return fail(DDD_EINVAL, "negative pair count");
    if (n_pairs == 0) return DDD_OK;
    if (!offsets || !out_rmsd) return fail(DDD_EINVAL, "offsets/out is NULL");
    if (offsets[0] != 0) return fail(DDD_EINVAL, "offsets[0] must be 0");
    for (int64_t i = 0; i < n_pairs; ++i) if (offsets[i + 1] < offsets[i]) return fail(DDD_EINVAL, "offsets not monotone");
    if (offsets[n_pairs] > 0 && (!sim || !ref)) return fail(DDD_EINVAL, "sim/ref is NULL");
    Scratch sc;
    std::vector<std::function<int()>> later;
    const std::vector<int64_t> blk = split_ligands(offsets, n_pairs);
    for (size_t di = 0; di < g_devs.size(); ++di) {
        const int64_t l0 = blk[di], l1 = blk[di + 1], np_blk = l1 - l0;
        if (np_blk <= 0) continue;
        const Dev &dev = g_devs[di];
        CUDA_TRY(cudaSetDevice(dev.id));
        const int64_t a0 = offsets[l0], na = offsets[l1] - a0;
        long long *d_off = nullptr; double *d_s = nullptr, *d_r = nullptr, *d_o = nullptr;
        if (int rc = sc.get(dev.id, &d_off, (size_t)np_blk + 1)) return rc;
        if (int rc = sc.get(dev.id, &d_s, (size_t)na * atom_stride)) return rc;
        if (int rc = sc.get(dev.id, &d_r, (size_t)na * atom_stride)) return rc;
        if (int rc = sc.get(dev.id, &d_o, (size_t)np_blk)) return rc;
        CUDA_TRY(cudaMemcpyAsync(d_off, offsets + l0, sizeof(int64_t) * ((size_t)np_blk + 1), cudaMemcpyHostToDevice, dev.stream));
        if (na > 0) {
            CUDA_TRY(cudaMemcpyAsync(d_s, sim + a0 * atom_stride, sizeof(double) * (size_t)(na * atom_stride), cudaMemcpyHostToDevice, dev.stream));
            CUDA_TRY(cudaMemcpyAsync(d_r, ref + a0 * atom_stride, sizeof(double) * (size_t)(na * atom_stride), cudaMemcpyHostToDevice, dev.stream));
        }
        rmsd_batch_kernel<<<rmsd_grid(dev, np_blk), 256, 0, dev.stream>>>(d_off, a0, d_s, d_r, atom_stride, np_blk, d_o);
        CUDA_TRY(cudaGetLastError());
        later.push_back([=]() -> int {
            CUDA_TRY(cudaSetDevice(dev.id));
            CUDA_TRY(cudaMemcpyAsync(out_rmsd + l0, d_o, sizeof(double) * (size_t)np_blk, cudaMemcpyDeviceToHost, dev.stream));
            return DDD_OK;
        });
    }
    return join_devices(later);
}

static int closest_impl(const ddd_receptor *r, const int64_t *offsets, double *lig_xyzq, bool copy_back,
                        int64_t n_lig, double *out_distance, int64_t *out_lig_atom, int64_t *out_prot_atom,
                        double threshold, int do_shift) {
    REQUIRE_INIT();
    if (int rc0 = check_receptor(r)) return rc0;
    if (int rc = check_csr(offsets, n_lig, lig_xyzq, nullptr)) return rc;
    if (n_lig == 0) return DDD_OK;
    Scratch sc;
    std::vector<std::function<int()>> later;
    const std::vector<int64_t> blk = split_ligands(offsets, n_lig);
    for (size_t di = 0; di < g_devs.size(); ++di) {
        const int64_t l0 = blk[di], l1 = blk[di + 1], nl_blk = l1 - l0;
        if (nl_blk <= 0) continue;
        const Dev &dev = g_devs[di];
        CUDA_TRY(cudaSetDevice(dev.id));
        const int64_t a0 = offsets[l0], na = offsets[l1] - a0;
        long long *d_off = nullptr, *d_li = nullptr, *d_pj = nullptr; double *d_lig = nullptr, *d_d = nullptr;
        if (int rc = sc.get(dev.id, &d_off, (size_t)nl_blk + 1)) return rc;
        if (int rc = sc.get(dev.id, &d_lig, (size_t)na * 4)) return rc;
        if (int rc = sc.get(dev.id, &d_d, (size_t)nl_blk)) return rc;
        if (int rc = sc.get(dev.id, &d_li, (size_t)nl_blk)) return rc;
        if (int rc = sc.get(dev.id, &d_pj, (size_t)nl_blk)) return rc;
        CUDA_TRY(cudaMemcpyAsync(d_off, offsets + l0, sizeof(int64_t) * ((size_t)nl_blk + 1), cudaMemcpyHostToDevice, dev.stream));
        if (na > 0) CUDA_TRY(cudaMemcpyAsync(d_lig, lig_xyzq + a0 * 4, sizeof(double) * 4 * (size_t)na, cudaMemcpyHostToDevice, dev.stream));
        ClosestArgs a;
        a.rec = r->view(di); a.offsets = d_off; a.atom_base = a0; a.lig_xyzq = d_lig; a.out_dist = d_d; a.out_lig_atom = d_li; a.out_prot_atom = d_pj;
        a.threshold = threshold; a.do_shift = do_shift;
        closest_atom_kernel<<<(unsigned)nl_blk, kThreads, 0, dev.stream>>>(a);
        CUDA_TRY(cudaGetLastError());
        later.push_back([=]() -> int {
            CUDA_TRY(cudaSetDevice(dev.id));
            if (out_distance) CUDA_TRY(cudaMemcpyAsync(out_distance + l0, d_d, sizeof(double) * (size_t)nl_blk, cudaMemcpyDeviceToHost, dev.stream));
            if (out_lig_atom) CUDA_TRY(cudaMemcpyAsync(out_lig_atom + l0, d_li, sizeof(int64_t) * (size_t)nl_blk, cudaMemcpyDeviceToHost, dev.stream));
            if (out_prot_atom) CUDA_TRY(cudaMemcpyAsync(out_prot_atom + l0, d_pj, sizeof(int64_t) * (size_t)nl_blk, cudaMemcpyDeviceToHost, dev.stream));
            if (copy_back && na > 0) CUDA_TRY(cudaMemcpyAsync(lig_xyzq + a0 * 4, d_lig, sizeof(double) * 4 * (size_t)na, cudaMemcpyDeviceToHost, dev.stream));
            return DDD_OK;
        });
    }
    return join_devices(later);
}

int ddd_closest_atom_batch(const ddd_receptor *r, const int64_t *offsets, const double *lig_xyzq, int64_t n_lig,
                           double *out_distance, int64_t *out_lig_atom, int64_t *out_prot_atom) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    return closest_impl(r, offsets, const_cast<double *>(lig_xyzq), false, n_lig, out_distance, out_lig_atom, out_prot_atom, 0.0, 0);
}

int ddd_shift_closer_batch(const ddd_receptor *r, const int64_t *offsets, double *lig_xyzq_inout, int64_t n_lig, double threshold) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    return closest_impl(r, offsets, lig_xyzq_inout, true, n_lig, nullptr, nullptr, nullptr, threshold, 1);
}

int ddd_randomize_pose_batch(const int64_t *offsets, double *lig_xyzq_inout, int64_t n_lig, uint64_t seed, uint64_t first_ligand_index) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (int rc = check_csr(offsets, n_lig, lig_xyzq_inout, nullptr)) return rc;
    if (n_lig == 0 || offsets[n_lig] == 0) return DDD_OK;
    Scratch sc;
    std::vector<std::function<int()>> later;
    const std::vector<int64_t> blk = split_ligands(offsets, n_lig);
    for (size_t di = 0; di < g_devs.size(); ++di) {
        const int64_t l0 = blk[di], l1 = blk[di + 1], nl_blk = l1 - l0;
        if (nl_blk <= 0) continue;
        const Dev &dev = g_devs[di];
        CUDA_TRY(cudaSetDevice(dev.id));
        const int64_t a0 = offsets[l0], na = offsets[l1] - a0;
        long long *d_off = nullptr; double *d_lig = nullptr;
        if (int rc = sc.get(dev.id, &d_off, (size_t)nl_blk + 1)) return rc;
        if (int rc = sc.get(dev.id, &d_lig, (size_t)na * 4)) return rc;
        CUDA_TRY(cudaMemcpyAsync(d_off, offsets + l0, sizeof(int64_t) * ((size_t)nl_blk + 1), cudaMemcpyHostToDevice, dev.stream));
        if (na > 0) CUDA_TRY(cudaMemcpyAsync(d_lig, lig_xyzq_inout + a0 * 4, sizeof(double) * 4 * (size_t)na, cudaMemcpyHostToDevice, dev.stream));
        randomize_pose_kernel<<<(unsigned)nl_blk, kThreads, 0, dev.stream>>>(d_off, d_lig, a0, seed, first_ligand_index + (uint64_t)l0);
        CUDA_TRY(cudaGetLastError());
        later.push_back([=]() -> int {
            CUDA_TRY(cudaSetDevice(dev.id));
            if (na > 0) CUDA_TRY(cudaMemcpyAsync(lig_xyzq_inout + a0 * 4, d_lig, sizeof(double) * 4 * (size_t)na, cudaMemcpyDeviceToHost, dev.stream));
            return DDD_OK;
        });
    }
    return join_devices(later);
}

// ------------------------------------------------------------------ device-resident screen
int ddd_screen_create(const ddd_receptor *r, const int64_t *offsets, const double *lig_xyzq, int64_t n_lig,
                      int32_t replicas, uint64_t first_chain_id, ddd_screen **out) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    return screen_create_impl(r, offsets, lig_xyzq, n_lig, replicas, first_chain_id, out);
}

int ddd_screen_run(ddd_screen *s, int64_t steps_per_chain, int32_t rotate, double temperature, const ddd_params *p, uint64_t seed) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    return screen_run_impl(s, steps_per_chain, rotate, temperature, p, seed, nullptr, nullptr, false);
}

int ddd_screen_set_chain_threads(ddd_screen *s, int32_t threads) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (int rc0 = check_screen(s)) return rc0;
    if (threads != 0 && threads != 32 && threads != 64 && threads != 128 && threads != 256)
        return fail(DDD_EINVAL, "chain CTA width must be 0, 32, 64, 128 or 256 (got %d)", threads);
    s->threads_override = threads;
    return DDD_OK;
}

int32_t ddd_screen_chain_threads(const ddd_screen *s) {
    if (!s || s->devs.empty()) return 0;
    return s->devs[0].threads;
}

int ddd_screen_set_chain_slots(ddd_screen *s, int32_t slots) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (int rc0 = check_screen(s)) return rc0;
    if (slots < 0 || slots > kSmMaxSlots) return fail(DDD_EINVAL, "chains per SM must be 0..%d (got %d)", kSmMaxSlots, slots);
    s->slots_override = slots;
    if (slots > 0) s->threads_override = 0;
    return DDD_OK;
}

int ddd_screen_set_segment_steps(ddd_screen *s, int64_t steps) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (int rc0 = check_screen(s)) return rc0;
    if (steps < -1) return fail(DDD_EINVAL, "segment steps must be -1 (automatic), 0 (whole chains) or > 0");
    s->segment_override = steps;
    return DDD_OK;
}

int64_t ddd_screen_segment_steps(const ddd_screen *s) {
    if (!s || s->devs.empty()) return 0;
    return s->devs[0].seg_steps;
}

int32_t ddd_screen_chain_slots(const ddd_screen *s) {
    if (!s || s->devs.empty()) return 0;
    return s->devs[0].slots;
}

int ddd_screen_sync(ddd_screen *s) {
    std::unique_lock<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    return screen_sync_impl(s, &lk);
}

double ddd_screen_last_kernel_ms(const ddd_screen *s) {
    if (!s) return -1.0;
    float mx = 0.f;
    for (auto &d : s->devs) mx = std::max(mx, d.last_ms);
    return (double)mx;
}

int ddd_screen_download(ddd_screen *s, double *out_xyzq, ddd_chain_stats *out_stats) {
    std::unique_lock<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    return screen_download_impl(s, out_xyzq, out_stats, nullptr, &lk);
}

// CompareRMSD (validateResults.go:40-45) for every chain of the screen, device-resident: CalculateRMSD(final pose of the
// chain, start pose of its ligand).  out_rmsd[n_chains]; the kernel time is returned through out_kernel_ms (nullable).
int ddd_screen_rmsd(ddd_screen *s, double *out_rmsd, double *out_kernel_ms) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (int rc0 = check_screen(s)) return rc0;
    if (!s->launched) return fail(DDD_EINVAL, "screen has not been run");
    if (!out_rmsd && s->n_chains > 0) return fail(DDD_EINVAL, "out_rmsd is NULL");
    for (auto &d : s->devs) {
        const Dev &dev = g_devs[d.di];
        CUDA_TRY(cudaSetDevice(dev.id));
        const int64_t n_local = d.c1 - d.c0;
        if (!d.d_rmsd) { if (int rc = dev_alloc(&d.d_rmsd, (size_t)n_local)) return rc; }
        CUDA_TRY(cudaEventRecord(dev.ev0, dev.stream));
        rmsd_chains_kernel<<<rmsd_grid(dev, n_local), 256, 0, dev.stream>>>(d.d_offsets, d.d_lig, d.atom_base, d.d_out, d.out_base,
                                                                               d.c0, n_local, s->replicas, d.d_rmsd);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(dev.ev1, dev.stream));
        CUDA_TRY(cudaMemcpyAsync(out_rmsd + d.c0, d.d_rmsd, sizeof(double) * (size_t)n_local, cudaMemcpyDeviceToHost, dev.stream));
    }
    double mx = 0.0;
    for (auto &d : s->devs) {
        const Dev &dev = g_devs[d.di];
        CUDA_TRY(cudaSetDevice(dev.id));
        CUDA_TRY(cudaStreamSynchronize(dev.stream));
        CUDA_TRY(cudaEventElapsedTime(&d.rmsd_ms, dev.ev0, dev.ev1));
        mx = std::max(mx, (double)d.rmsd_ms);
    }
    if (out_kernel_ms) *out_kernel_ms = mx;
    return DDD_OK;
}

// The RMSD values the chain kernel's own epilogue left behind (SM-persistent engine): same bits as ddd_screen_rmsd, no extra pass.
int ddd_screen_fused_rmsd(ddd_screen *s, double *out_rmsd) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (int rc0 = check_screen(s)) return rc0;
    if (!s->launched) return fail(DDD_EINVAL, "screen has not been run");
    if (!out_rmsd && s->n_chains > 0) return fail(DDD_EINVAL, "out_rmsd is NULL");
    for (auto &d : s->devs) {
        if (!d.fused_rmsd) return fail(DDD_EINVAL, "the last run used the one-CTA-per-chain engine: call ddd_screen_rmsd");
        CUDA_TRY(cudaSetDevice(g_devs[d.di].id));
        CUDA_TRY(cudaMemcpyAsync(out_rmsd + d.c0, d.d_rmsd, sizeof(double) * (size_t)(d.c1 - d.c0), cudaMemcpyDeviceToHost, g_devs[d.di].stream));
    }
    for (auto &d : s->devs) { CUDA_TRY(cudaSetDevice(g_devs[d.di].id)); CUDA_TRY(cudaStreamSynchronize(g_devs[d.di].stream)); }
    return DDD_OK;
}

// ------------------------------------------------------------------ resident pipeline: randomise -> shift -> chains -> pick -> RMSD -> argmin
// RandomizeLigandPose (validateResults.go:70-79) on the resident start poses, in place.  Ligand i draws from stream
// (seed, DDD_STREAM_RANDOMIZE_BASE + first_ligand_index + i), like ddd_randomize_pose_batch.
int ddd_screen_randomize(ddd_screen *s, uint64_t seed, uint64_t first_ligand_index) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (int rc0 = check_screen(s)) return rc0;
    if (int rc = ensure_reference(s)) return rc;
    for (auto &d : s->devs) {
        const int64_t n_ll = d.lig1 - d.lig0;
        if (n_ll <= 0 || d.n_atoms_in == 0) continue;
        const Dev &dev = g_devs[d.di];
        CUDA_TRY(cudaSetDevice(dev.id));
        randomize_pose_kernel<<<(unsigned)n_ll, kThreads, 0, dev.stream>>>(d.d_offsets + d.lig0, d.d_lig, d.atom_base, seed, first_ligand_index + (uint64_t)d.lig0);
        CUDA_TRY(cudaGetLastError());
    }
    return DDD_OK;
}

// ShiftLigandCloserByThreshold (metropolis.go:267-285; what SimulateMultipleLigands does first, :14-16) on the resident start poses
int ddd_screen_shift_closer(ddd_screen *s, double threshold) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (int rc0 = check_screen(s)) return rc0;
    if (int rc = ensure_reference(s)) return rc;
    for (auto &d : s->devs) {
        const int64_t n_ll = d.lig1 - d.lig0;
        if (n_ll <= 0) continue;
        const Dev &dev = g_devs[d.di];
        CUDA_TRY(cudaSetDevice(dev.id));
        ClosestArgs a;
        a.rec = s->rec->view(d.di); a.offsets = d.d_offsets + d.lig0; a.atom_base = d.atom_base; a.lig_xyzq = d.d_lig;
        a.out_dist = nullptr; a.out_lig_atom = nullptr; a.out_prot_atom = nullptr; a.threshold = threshold; a.do_shift = 1;
        closest_atom_kernel<<<(unsigned)n_ll, kThreads, 0, dev.stream>>>(a);
        CUDA_TRY(cudaGetLastError());
    }
    return DDD_OK;
}

// the start poses as they are resident now (after randomise / shift): for checking the pipeline stage by stage
int ddd_screen_download_start(ddd_screen *s, double *out_xyzq) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (int rc0 = check_screen(s)) return rc0;
    for (auto &d : s->devs) {
        if (d.n_atoms_in == 0) continue;
        if (!out_xyzq) return fail(DDD_EINVAL, "out_xyzq is NULL");
        CUDA_TRY(cudaSetDevice(g_devs[d.di].id));
        CUDA_TRY(cudaMemcpyAsync(out_xyzq + d.atom_base * 4, d.d_lig, sizeof(double) * 4 * (size_t)d.n_atoms_in, cudaMemcpyDeviceToHost, g_devs[d.di].stream));
    }
    for (auto &d : s->devs) { CUDA_TRY(cudaSetDevice(g_devs[d.di].id)); CUDA_TRY(cudaStreamSynchronize(g_devs[d.di].stream)); }
    return DDD_OK;
}

// Replica pick (:102-109) of the last run, per ligand, on the device; the picked pose, its energy (= CalculateEnergy(minLigand),
// :61), the replica index and CalculateRMSD(picked pose, reference pose) (= CompareRMSD's return, validateResults.go:40-45)
// stay resident for ddd_screen_argmin and are copied out where a pointer is given.
int ddd_screen_pick(ddd_screen *s, double temperature, uint64_t seed, uint64_t first_ligand_index,
                    double *out_xyzq, double *out_energy, int64_t *out_picked, double *out_rmsd) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (int rc0 = check_screen(s)) return rc0;
    if (int rc = screen_pick_impl(s, temperature, seed, first_ligand_index)) return rc;
    return screen_download_pick(s, out_xyzq, out_energy, out_picked, out_rmsd);
}

// SaveMinimumEnergyLigand's choice (plotting.go:109-121) over the picked per-ligand energies, as a device reduction:
// rule DDD_ARGMIN_SAVE_MIN_LIGAND = (minIndex 0, minEnergy 0.0, strict <): only a negative energy can win;
// rule DDD_ARGMIN_RSHINY = RShinyAppMain's (main.go:35, 51-54): minEnergy starts at 1e20.
int ddd_screen_argmin(ddd_screen *s, int32_t rule, int64_t *out_index, double *out_energy) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (int rc0 = check_screen(s)) return rc0;
    if (!s->picked) return fail(DDD_EINVAL, "ddd_screen_argmin needs ddd_screen_pick on the last run's results");
    if (rule != DDD_ARGMIN_SAVE_MIN_LIGAND && rule != DDD_ARGMIN_RSHINY) return fail(DDD_EINVAL, "unknown argmin rule %d", rule);
    std::vector<double> hv(s->devs.size(), 0.0);
    std::vector<long long> hi(s->devs.size(), -1);
    for (size_t k = 0; k < s->devs.size(); ++k) {
        auto &d = s->devs[k];
        const Dev &dev = g_devs[d.di];
        CUDA_TRY(cudaSetDevice(dev.id));
        if (!d.d_arg_val) { if (int rc = dev_alloc(&d.d_arg_val, 1)) return rc; }
        if (!d.d_arg_idx) { if (int rc = dev_alloc(&d.d_arg_idx, 1)) return rc; }
        argmin_first_kernel<<<1, 1024, 0, dev.stream>>>(d.d_min_e, d.lig1 - d.lig0, d.lig0, d.d_arg_val, d.d_arg_idx);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(&hv[k], d.d_arg_val, sizeof(double), cudaMemcpyDeviceToHost, dev.stream));
        CUDA_TRY(cudaMemcpyAsync(&hi[k], d.d_arg_idx, sizeof(long long), cudaMemcpyDeviceToHost, dev.stream));
    }
    for (auto &d : s->devs) { CUDA_TRY(cudaSetDevice(g_devs[d.di].id)); CUDA_TRY(cudaStreamSynchronize(g_devs[d.di].stream)); }
    double best = rule == DDD_ARGMIN_RSHINY ? 1e20 : 0.0;                  // main.go:35 / plotting.go:111
    int64_t bi = 0;                                                        // minIndex := 0 (plotting.go:110; main.go:36 likewise)
    for (size_t k = 0; k < s->devs.size(); ++k)                            // devices hold ascending ligand blocks: strict < keeps the first
        if (hi[k] >= 0 && hv[k] < best) { best = hv[k]; bi = hi[k]; }
    if (out_index) *out_index = bi;
    if (out_energy) *out_energy = best;
    return DDD_OK;
}

// EXTENSION (north_star: "per-ligand minimum energies and poses"; the reference returns the FINAL pose, metropolis.go:110,135):
// track the lowest-energy pose every chain visits.  Identical to the final pose while the sampler is greedy.
int ddd_screen_set_track_best(ddd_screen *s, int32_t on) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!s) return fail(DDD_EINVAL, "screen is NULL");
    s->track_best = on != 0;
    return DDD_OK;
}

// out_xyzq: chain-major like ddd_screen_download; out_energy[n_chains] = the chain's own (fp32-summed in F32 mode) energy of that pose
int ddd_screen_download_best(ddd_screen *s, double *out_xyzq, double *out_energy) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (int rc0 = check_screen(s)) return rc0;
    if (!s->launched) return fail(DDD_EINVAL, "screen has not been run");
    for (auto &d : s->devs) {
        if (!d.d_best || !d.d_best_e) return fail(DDD_EINVAL, "the last run did not track best-ever poses (ddd_screen_set_track_best)");
        CUDA_TRY(cudaSetDevice(g_devs[d.di].id));
        if (out_xyzq && d.out_atoms > 0)
            CUDA_TRY(cudaMemcpyAsync(out_xyzq + d.out_base * 4, d.d_best, sizeof(double) * 4 * (size_t)d.out_atoms, cudaMemcpyDeviceToHost, g_devs[d.di].stream));
        if (out_energy) CUDA_TRY(cudaMemcpyAsync(out_energy + d.c0, d.d_best_e, sizeof(double) * (size_t)(d.c1 - d.c0), cudaMemcpyDeviceToHost, g_devs[d.di].stream));
    }
    for (auto &d : s->devs) { CUDA_TRY(cudaSetDevice(g_devs[d.di].id)); CUDA_TRY(cudaStreamSynchronize(g_devs[d.di].stream)); }
    return DDD_OK;
}

// Cutoff extension: receptor atoms actually paired with each ligand atom by chain c's fp32 sums, summed over its
// evaluations (32 x the units its poses' cell-list queries returned).  Zero-filled when the last run had no cutoff.
int ddd_screen_pair_units(ddd_screen *s, int64_t *out_receptor_atoms) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (int rc0 = check_screen(s)) return rc0;
    if (!s->launched) return fail(DDD_EINVAL, "screen has not been run");
    if (!out_receptor_atoms && s->n_chains > 0) return fail(DDD_EINVAL, "out is NULL");
    if (int rc = screen_sync_impl(s)) return rc;
    for (auto &d : s->devs) {
        const int64_t n_local = d.c1 - d.c0;
        if (!d.d_units || !s->last_cut) { std::fill(out_receptor_atoms + d.c0, out_receptor_atoms + d.c1, (int64_t)0); continue; }
        CUDA_TRY(cudaSetDevice(g_devs[d.di].id));
        CUDA_TRY(cudaMemcpyAsync(out_receptor_atoms + d.c0, d.d_units, sizeof(int64_t) * (size_t)n_local, cudaMemcpyDeviceToHost, g_devs[d.di].stream));
        CUDA_TRY(cudaStreamSynchronize(g_devs[d.di].stream));
        for (int64_t i = d.c0; i < d.c1; ++i) out_receptor_atoms[i] *= 32;
    }
    return DDD_OK;
}

int ddd_screen_set_lj(ddd_screen *s, const double *sigma, const double *epsilon) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (int rc0 = check_screen(s)) return rc0;
    if (!sigma || !epsilon) return fail(DDD_EINVAL, "sigma/epsilon is NULL");
    for (auto &d : s->devs) {
        std::vector<double2> h((size_t)std::max<int64_t>(d.n_atoms_in, 1));
        for (int64_t i = 0; i < d.n_atoms_in; ++i) {
            const double sg = sigma[d.atom_base + i], ep = epsilon[d.atom_base + i];
            if (!(sg >= 0.0) || !(ep >= 0.0)) return fail(DDD_EINVAL, "sigma/epsilon of ligand atom %lld is negative or NaN", (long long)(d.atom_base + i));
            h[(size_t)i] = make_double2(0.5 * sg, std::sqrt(ep));
        }
        CUDA_TRY(cudaSetDevice(g_devs[d.di].id));
        if (!d.d_lig_lj) { if (int rc = dev_alloc(&d.d_lig_lj, h.size())) return rc; }
        CUDA_TRY(cudaMemcpyAsync(d.d_lig_lj, h.data(), sizeof(double2) * h.size(), cudaMemcpyHostToDevice, g_devs[d.di].stream));
        CUDA_TRY(cudaStreamSynchronize(g_devs[d.di].stream));
    }
    s->has_lj = true;
    return DDD_OK;
}

int ddd_screen_destroy(ddd_screen *s) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    screen_destroy_impl(s);
    return DDD_OK;
}

// ------------------------------------------------------------------ introspection
int ddd_device_info(int ordinal, int32_t *sm_count, int32_t *clock_khz, int32_t *cc_major, int32_t *cc_minor,
                    char *name, int32_t name_cap) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (ordinal < 0 || ordinal >= (int)g_devs.size()) return fail(DDD_EINVAL, "ordinal out of range");
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, g_devs[(size_t)ordinal].id));
    int khz = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, g_devs[(size_t)ordinal].id));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (clock_khz) *clock_khz = khz;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (name && name_cap > 0) { std::strncpy(name, prop.name, (size_t)name_cap - 1); name[name_cap - 1] = '\0'; }
    return DDD_OK;
}

int ddd_chain_kernel_info(int32_t precision, int32_t *threads, int32_t *regs, int32_t *static_smem, int32_t *max_blocks_per_sm_at_nl40) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    CUDA_TRY(cudaSetDevice(g_devs[0].id));
    cudaFuncAttributes fa;
    int blocks = 0;
    const size_t smem = sm_smem_bytes(40, 7);            // the SM-persistent engine at config 2's shape: 7 chains of 40 atoms per SM
    if (precision == DDD_PRECISION_F64) {
        if (int rc = set_smem(mc_sm_kernel<true, MODE_PLAIN>, smem)) return rc;
        CUDA_TRY(cudaFuncGetAttributes(&fa, mc_sm_kernel<true, MODE_PLAIN>));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, mc_sm_kernel<true, MODE_PLAIN>, kSmThreads, smem));
    } else {
        if (int rc = set_smem(mc_sm_kernel<false, MODE_PLAIN>, smem)) return rc;
        CUDA_TRY(cudaFuncGetAttributes(&fa, mc_sm_kernel<false, MODE_PLAIN>));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, mc_sm_kernel<false, MODE_PLAIN>, kSmThreads, smem));
    }
    if (threads) *threads = kSmThreads;
    if (regs) *regs = fa.numRegs;
    if (static_smem) *static_smem = (int32_t)fa.sharedSizeBytes;
    if (max_blocks_per_sm_at_nl40) *max_blocks_per_sm_at_nl40 = blocks;
    return DDD_OK;
}

}  // extern "C"
