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
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

using namespace ddd;

static_assert(sizeof(ChainStats) == sizeof(ddd_chain_stats), "ChainStats must mirror ddd_chain_stats");
static_assert(sizeof(Atom64) == 32, "Atom64 must be 4 doubles");

namespace {

thread_local std::string g_err;
std::recursive_mutex g_mu;

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
    int *d_queue = nullptr;                // SM-persistent engine: dynamic chain counter
    double *d_rmsd = nullptr;              // per-chain RMSD (ddd_screen_rmsd)
    long long *d_units = nullptr;          // per-chain receptor units evaluated (cutoff extension)
    double2 *d_lig_lj = nullptr;           // LJ extension: (sigma/2, sqrt(eps)) of this device's slice of the ligand atoms
    float rmsd_ms = 0.f;
};

struct ddd_screen {
    const ddd_receptor *rec = nullptr;
    int64_t n_lig = 0, n_chains = 0, steps = 0;
    int replicas = 1;
    uint64_t chain_id_base = 0;            // Philox stream id of local chain 0
    std::vector<int64_t> offsets;
    std::vector<ScreenDev> devs;
    bool launched = false;
    int threads_override = 0;              // > 0: one-CTA-per-chain engine with this CTA width
    int slots_override = 0;                // > 0: SM-persistent engine with this many chains per SM; both 0: automatic
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
constexpr int kSlotsAuto = 8;
int pick_chain_slots(int64_t n_chains, int grid) {
    const int64_t per_sm = (n_chains + grid - 1) / std::max(grid, 1);
    return (int)std::max<int64_t>(1, std::min<int64_t>(kSlotsAuto, per_sm));
}

// chains [begin[k], begin[k+1]) go to shard k; contiguous, balanced by sum(nL + 1)
void plan_shards(const int64_t *offsets, int64_t n_lig, int replicas, int n_shards, std::vector<int64_t> &begin) {
    begin.assign((size_t)n_shards + 1, 0);
    const int64_t n_chains = n_lig * replicas;
    begin[(size_t)n_shards] = n_chains;
    if (n_shards <= 1 || n_chains == 0) return;
    const double total = (double)replicas * (double)(offsets[n_lig] + n_lig);
    int64_t lig = 0;
    for (int k = 1; k < n_shards; ++k) {
        const double target = total * (double)k / (double)n_shards;
        while (lig < n_lig && (double)replicas * (double)(offsets[lig + 1] + lig + 1) <= target) ++lig;
        int64_t c = lig * replicas;
        if (lig < n_lig) {
            const double before = (double)replicas * (double)(offsets[lig] + lig);
            const double per_chain = (double)(offsets[lig + 1] - offsets[lig] + 1);
            int64_t r = (int64_t)std::llround((target - before) / per_chain);
            r = std::max<int64_t>(0, std::min<int64_t>(replicas, r));
            c += r;
        }
        begin[(size_t)k] = std::max(begin[(size_t)k - 1], std::min(c, n_chains));
    }
}

void free_screen_dev(ScreenDev &sd) {
    if (sd.di >= g_devs.size()) { sd = ScreenDev(); return; }      // the library was shut down: the context owns what is left
    cudaSetDevice(g_devs[sd.di].id);
    dev_free(sd.d_offsets); dev_free(sd.d_lig); dev_free(sd.d_order); dev_free(sd.d_out);
    dev_free(sd.d_stats); dev_free(sd.d_trace); dev_free(sd.d_rec); dev_free(sd.d_rec_off); dev_free(sd.d_queue); dev_free(sd.d_rmsd); dev_free(sd.d_units); dev_free(sd.d_lig_lj);
    sd = ScreenDev();
}

int screen_create_impl(const ddd_receptor *r, const int64_t *offsets, const double *lig_xyzq, int64_t n_lig,
                       int32_t replicas, uint64_t first_chain_id, ddd_screen **out) {
    if (!r || !out) return fail(DDD_EINVAL, "receptor/out is NULL");
    if (replicas <= 0) return fail(DDD_EINVAL, "replicas must be >= 1 (got %d)", replicas);
    int64_t max_nl = 0;
    if (int rc = check_csr(offsets, n_lig, lig_xyzq, &max_nl)) return rc;
    ddd_screen *s = new ddd_screen();
    s->rec = r; s->n_lig = n_lig; s->replicas = replicas; s->n_chains = n_lig * replicas; s->chain_id_base = first_chain_id;
    s->offsets.assign(offsets ? offsets : nullptr, offsets ? offsets + n_lig + 1 : nullptr);
    if (s->offsets.empty()) s->offsets.push_back(0);
    std::vector<int64_t> begin;
    plan_shards(s->offsets.data(), n_lig, replicas, (int)g_devs.size(), begin);
    for (size_t di = 0; di < g_devs.size(); ++di) {
        ScreenDev sd; sd.di = di; sd.c0 = begin[di]; sd.c1 = begin[di + 1];
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
        if (!rc) guard(dev_alloc(&d.d_offsets, (size_t)n_lig + 1));
        if (!rc) guard(dev_alloc(&d.d_lig, (size_t)d.n_atoms_in * 4));
        if (!rc) guard(dev_alloc(&d.d_out, (size_t)d.out_atoms * 4));
        if (!rc) guard(dev_alloc(&d.d_stats, (size_t)n_local));
        if (!rc) guard(dev_alloc(&d.d_queue, 1));
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
            for (auto &x : s->devs) free_screen_dev(x);
            delete s;
            return rc;
        }
    }
    *out = s;
    return DDD_OK;
}

int screen_run_impl(ddd_screen *s, int64_t steps, int32_t rotate, double temperature, const ddd_params *p,
                    uint64_t seed, const double *recorded, const int64_t *recorded_offsets, bool want_trace) {
    if (!s) return fail(DDD_EINVAL, "screen is NULL");
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
        if (cut && s->rec->n_units_s > 65535) return fail(DDD_ELIMIT, "params.cutoff supports receptors up to 2097120 atoms");
        if (s->threads_override <= 0 && n_local > 0 && fits_sm_engine) {
            // SM-persistent engine: one 32-warp CTA per SM, K resident chains each (ddd_sm_chain.cuh)
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
            const size_t smem = sm_smem_bytes(d.nl_cap, A.slots, ulb);
            d.threads = kSmThreads; d.slots = A.slots;
            CUDA_TRY(cudaMemsetAsync(d.d_queue, 0, sizeof(int), dev.stream));
            CUDA_TRY(cudaEventRecord(dev.ev0, dev.stream));
            if (precise) {
                if (int rc = set_smem(mc_sm_kernel<true>, smem)) return rc;
                mc_sm_kernel<true><<<(unsigned)grid, kSmThreads, smem, dev.stream>>>(A);
            } else {
                if (int rc = set_smem(mc_sm_kernel<false>, smem)) return rc;
                mc_sm_kernel<false><<<(unsigned)grid, kSmThreads, smem, dev.stream>>>(A);
            }
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaEventRecord(dev.ev1, dev.stream));
            continue;
        }
        const int threads = s->threads_override > 0 ? s->threads_override : pick_chain_threads(n_local, dev.sm_count);
        const size_t smem = chain_smem_bytes(d.nl_cap, threads);
        d.threads = threads; d.slots = 0;
        CUDA_TRY(cudaEventRecord(dev.ev0, dev.stream));
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
        CUDA_TRY(cudaEventRecord(dev.ev1, dev.stream));
    }
    s->launched = true;
    return DDD_OK;
}

int screen_sync_impl(ddd_screen *s) {
    if (!s) return fail(DDD_EINVAL, "screen is NULL");
    for (auto &d : s->devs) {
        const Dev &dev = g_devs[d.di];
        CUDA_TRY(cudaSetDevice(dev.id));
        CUDA_TRY(cudaStreamSynchronize(dev.stream));
        if (s->launched) CUDA_TRY(cudaEventElapsedTime(&d.last_ms, dev.ev0, dev.ev1));
    }
    return DDD_OK;
}

int screen_download_impl(ddd_screen *s, double *out_xyzq, ddd_chain_stats *out_stats, uint8_t *out_trace) {
    if (!s) return fail(DDD_EINVAL, "screen is NULL");
    if (!s->launched) return fail(DDD_EINVAL, "screen has not been run");
    if (int rc = screen_sync_impl(s)) return rc;
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
    for (auto &d : s->devs) free_screen_dev(d);
    delete s;
}

// AcceptMove (:141-149) on the host, for the replica pick only (:105)
bool accept_move_host(double cur_e, double new_e, double temperature, uint64_t seed, uint64_t stream, uint64_t &pos) {
    if (new_e < cur_e) return true;
    const double delta = new_e - cur_e;
    const double probability = std::exp(-delta / temperature);
    return philox_u01(seed, stream, pos++) < probability;
}

}  // namespace

// ------------------------------------------------------------------ library
extern "C" {

int ddd_init(const int *devices, int n_devices) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ddd_shutdown();
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
    for (int id : ids) {
        Dev d; d.id = id;
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, id));
        if (prop.major < 10) { g_devs.clear(); return fail(DDD_ENODEVICE, "device %d (%s, sm_%d%d) is not Blackwell: this library ships sm_100a code only", id, prop.name, prop.major, prop.minor); }
        d.sm_count = prop.multiProcessorCount;
        d.max_smem = prop.sharedMemPerBlockOptin;
        CUDA_TRY(cudaSetDevice(id));
        CUDA_TRY(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
        cudaMemPool_t pool = nullptr;
        CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, id));
        unsigned long long keep = ~0ull;                                  // never hand pool memory back between calls
        CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        CUDA_TRY(cudaEventCreate(&d.ev0));
        CUDA_TRY(cudaEventCreate(&d.ev1));
        g_devs.push_back(d);
    }
    return DDD_OK;
}

int ddd_shutdown(void) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
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

const char *ddd_version(void) { return "ddd_mc 0.2 (sm_100a; SM-persistent chain engine: 16 warps per SM, speculative proposals, per-group pair sums)"; }

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
        const int64_t n_pad_s = (r->n_units_s + 7) / 8 * 8 * 32;            // whole groups of 8 units may be read
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
    if (rc) { ddd_receptor_destroy(r); return rc; }
    *out = r;
    return DDD_OK;
}

int ddd_receptor_destroy(ddd_receptor *r) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!r) return DDD_OK;
    for (size_t di = 0; di < r->f32.size() && di < g_devs.size(); ++di) {
        cudaSetDevice(g_devs[di].id);
        cudaFree(r->f32[di]); cudaFree(r->f64[di]);
        if (di < r->f32s.size()) { cudaFree(r->f32s[di]); cudaFree(r->ubox[di]); cudaFree(r->f64s[di]); cudaFree(r->gbox[di]);
                                    cudaFree(r->lj64[di]); cudaFree(r->lj64s[di]); cudaFree(r->lj32s[di]); }
    }
    delete r;
    return DDD_OK;
}

int64_t ddd_receptor_num_atoms(const ddd_receptor *r) { return r ? r->n : -1; }

int ddd_receptor_set_lj(ddd_receptor *r, const double *sigma, const double *epsilon) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (!r) return fail(DDD_EINVAL, "receptor is NULL");
    if ((sigma == nullptr) != (epsilon == nullptr)) return fail(DDD_EINVAL, "sigma and epsilon must be given together");
    if (!sigma) { r->has_lj = false; return DDD_OK; }
    const int64_t n = r->n;
    const size_t n_pad = (size_t)((r->n_units_s + 7) / 8 * 8 * 32);
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
int ddd_energy_batch(const ddd_receptor *r, const int64_t *offsets, const double *lig_xyzq, int64_t n_lig,
                     const ddd_params *p, double *out_energy, double *out_abs_sum) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (!r) return fail(DDD_EINVAL, "receptor is NULL");
    if (int rc = check_params(p)) return rc;
    if (p->lj_scale != 0.0) return fail(DDD_EINVAL, "params.lj_scale: ligand LJ parameters travel with a screen (run it with 0 steps for energies)");
    int64_t max_nl = 0;
    if (int rc = check_csr(offsets, n_lig, lig_xyzq, &max_nl)) return rc;
    if (n_lig == 0) return DDD_OK;
    if (!out_energy) return fail(DDD_EINVAL, "out_energy is NULL");
    const Dev &dev = g_devs[0];
    CUDA_TRY(cudaSetDevice(dev.id));
    long long *d_off = nullptr; double *d_lig = nullptr, *d_e = nullptr, *d_abs = nullptr;
    int rc = DDD_OK;
    auto cleanup = [&]() { dev_free(d_off); dev_free(d_lig); dev_free(d_e); dev_free(d_abs); };
    if ((rc = dev_alloc(&d_off, (size_t)n_lig + 1)) || (rc = dev_alloc(&d_lig, (size_t)offsets[n_lig] * 4)) ||
        (rc = dev_alloc(&d_e, (size_t)n_lig)) || (out_abs_sum && (rc = dev_alloc(&d_abs, (size_t)n_lig)))) { cleanup(); return rc; }
    auto run = [&]() -> int {
        CUDA_TRY(cudaMemcpyAsync(d_off, offsets, sizeof(int64_t) * ((size_t)n_lig + 1), cudaMemcpyHostToDevice, dev.stream));
        if (offsets[n_lig] > 0)
            CUDA_TRY(cudaMemcpyAsync(d_lig, lig_xyzq, sizeof(double) * 4 * (size_t)offsets[n_lig], cudaMemcpyHostToDevice, dev.stream));
        EnergyArgs a;
        a.rec = r->view(0); a.prm = make_dev_params(*p, 0, 1.0);
        a.offsets = d_off; a.lig_xyzq = d_lig; a.out_energy = d_e; a.out_abs = d_abs; a.nl_cap = even_cap(max_nl);
        const size_t smem = chain_smem_bytes(a.nl_cap, kThreads);
        const bool precise = p->precision == DDD_PRECISION_F64, with_abs = out_abs_sum != nullptr;
        const unsigned grid = (unsigned)n_lig;
#define LAUNCH_E(PR, AB)                                                                   \
        do { if (int rc2 = set_smem(energy_batch_kernel<PR, AB>, smem)) return rc2;        \
             energy_batch_kernel<PR, AB><<<grid, kThreads, smem, dev.stream>>>(a); } while (0)
        if (precise && with_abs) LAUNCH_E(true, true);
        else if (precise) LAUNCH_E(true, false);
        else if (with_abs) LAUNCH_E(false, true);
        else LAUNCH_E(false, false);
#undef LAUNCH_E
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(out_energy, d_e, sizeof(double) * (size_t)n_lig, cudaMemcpyDeviceToHost, dev.stream));
        if (with_abs) CUDA_TRY(cudaMemcpyAsync(out_abs_sum, d_abs, sizeof(double) * (size_t)n_lig, cudaMemcpyDeviceToHost, dev.stream));
        CUDA_TRY(cudaStreamSynchronize(dev.stream));
        return DDD_OK;
    };
    rc = run();
    cleanup();
    return rc;
}

// ------------------------------------------------------------------ chains
int ddd_run_chains(const ddd_receptor *r, const int64_t *offsets, const double *lig_xyzq, int64_t n_lig,
                   int32_t replicas, int64_t steps_per_chain, int32_t rotate, double temperature,
                   const ddd_params *p, uint64_t seed, uint64_t first_chain_id,
                   const double *recorded, const int64_t *recorded_offsets,
                   double *out_xyzq, ddd_chain_stats *out_stats, uint8_t *out_accept_trace) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (int rc = check_params(p)) return rc;
    if (p->lj_scale != 0.0) return fail(DDD_EINVAL, "params.lj_scale: ligand LJ parameters travel with a screen (ddd_screen_create + ddd_screen_set_lj)");
    ddd_screen *s = nullptr;
    if (int rc = screen_create_impl(r, offsets, lig_xyzq, n_lig, replicas, first_chain_id, &s)) return rc;
    int rc = screen_run_impl(s, steps_per_chain, rotate, temperature, p, seed, recorded, recorded_offsets, out_accept_trace != nullptr);
    if (!rc) rc = screen_download_impl(s, out_xyzq, out_stats, out_accept_trace);
    screen_destroy_impl(s);
    return rc;
}

int ddd_minimize_batch(const ddd_receptor *r, const int64_t *offsets, const double *lig_xyzq, int64_t n_lig,
                       int64_t iterations, int32_t rotate, double temperature, int32_t num_procs,
                       const ddd_params *p, uint64_t seed, uint64_t first_ligand_index,
                       double *out_xyzq, double *out_energy, ddd_chain_stats *out_chain_stats, int64_t *out_picked) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (num_procs <= 0) return fail(DDD_EINVAL, "num_procs must be >= 1 (the reference divides iterations by it, metropolis.go:97)");
    if (iterations < 0) return fail(DDD_EINVAL, "negative iteration count");
    if (int rc = check_csr(offsets, n_lig, lig_xyzq, nullptr)) return rc;
    if (n_lig == 0) return DDD_OK;
    if (!out_xyzq && offsets[n_lig] > 0) return fail(DDD_EINVAL, "out_xyzq is NULL");
    if (!out_energy) return fail(DDD_EINVAL, "out_energy is NULL");
    const int64_t width = iterations / num_procs;                                  // :97
    const int64_t n_chains = n_lig * num_procs;
    std::vector<double> chain_pose((size_t)offsets[n_lig] * (size_t)num_procs * 4);
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
    const double t2 = now();
    if (!rc) rc = screen_download_impl(s, chain_pose.data(), stats.data(), nullptr);
    const double t3 = now();
    const double kms = ddd_screen_last_kernel_ms(s);
    screen_destroy_impl(s);
    const double t4 = now();
    if (dbg) fprintf(stderr, "[ddd_minimize_batch] create %.2f ms, launch %.2f ms, sync+download %.2f ms (kernel %.2f ms), destroy %.2f ms\n",
                     t1 - t0, t2 - t1, t3 - t2, kms, t4 - t3);
    if (rc) return rc;
    for (int64_t i = 0; i < n_lig; ++i) {
        const int64_t nl = offsets[i + 1] - offsets[i];
        double cur_e = stats[(size_t)(i * num_procs)].start_energy;                // :96
        int64_t picked = -1;
        uint64_t pos = 0;
        const uint64_t parent = DDD_STREAM_PARENT_BASE + first_ligand_index + (uint64_t)i;
        for (int rp = 0; rp < num_procs; ++rp) {                                    // :102-109, replica-index order
            const double new_e = stats[(size_t)(i * num_procs + rp)].final_energy; // :104
            if (accept_move_host(cur_e, new_e, temperature, seed, parent, pos)) { cur_e = new_e; picked = rp; }
        }
        double *dst = out_xyzq + offsets[i] * 4;
        if (picked >= 0)
            std::memcpy(dst, chain_pose.data() + ((size_t)num_procs * (size_t)offsets[i] + (size_t)picked * (size_t)nl) * 4, sizeof(double) * 4 * (size_t)nl);
        else std::memcpy(dst, lig_xyzq + offsets[i] * 4, sizeof(double) * 4 * (size_t)nl);
        out_energy[i] = cur_e;                                                      // :61 CalculateEnergy(final)
        if (out_picked) out_picked[i] = picked;
    }
    if (out_chain_stats) std::memcpy(out_chain_stats, stats.data(), sizeof(ddd_chain_stats) * stats.size());
    return DDD_OK;
}

// ------------------------------------------------------------------ RMSD / closest / shift / randomize (device 0)
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
    const int64_t n_atoms = offsets[n_pairs];
    if (n_atoms > 0 && (!sim || !ref)) return fail(DDD_EINVAL, "sim/ref is NULL");
    const Dev &dev = g_devs[0];
    CUDA_TRY(cudaSetDevice(dev.id));
    long long *d_off = nullptr; double *d_s = nullptr, *d_r = nullptr, *d_o = nullptr;
    auto cleanup = [&]() { dev_free(d_off); dev_free(d_s); dev_free(d_r); dev_free(d_o); };
    int rc;
    if ((rc = dev_alloc(&d_off, (size_t)n_pairs + 1)) || (rc = dev_alloc(&d_s, (size_t)n_atoms * atom_stride)) ||
        (rc = dev_alloc(&d_r, (size_t)n_atoms * atom_stride)) || (rc = dev_alloc(&d_o, (size_t)n_pairs))) { cleanup(); return rc; }
    auto run = [&]() -> int {
        CUDA_TRY(cudaMemcpyAsync(d_off, offsets, sizeof(int64_t) * ((size_t)n_pairs + 1), cudaMemcpyHostToDevice, dev.stream));
        if (n_atoms > 0) {
            CUDA_TRY(cudaMemcpyAsync(d_s, sim, sizeof(double) * (size_t)(n_atoms * atom_stride), cudaMemcpyHostToDevice, dev.stream));
            CUDA_TRY(cudaMemcpyAsync(d_r, ref, sizeof(double) * (size_t)(n_atoms * atom_stride), cudaMemcpyHostToDevice, dev.stream));
        }
        const unsigned grid = (unsigned)((n_pairs + 7) / 8);
        rmsd_batch_kernel<<<grid, 256, 0, dev.stream>>>(d_off, d_s, d_r, atom_stride, n_pairs, d_o);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(out_rmsd, d_o, sizeof(double) * (size_t)n_pairs, cudaMemcpyDeviceToHost, dev.stream));
        CUDA_TRY(cudaStreamSynchronize(dev.stream));
        return DDD_OK;
    };
    rc = run();
    cleanup();
    return rc;
}

static int closest_impl(const ddd_receptor *r, const int64_t *offsets, double *lig_xyzq, bool copy_back,
                        int64_t n_lig, double *out_distance, int64_t *out_lig_atom, int64_t *out_prot_atom,
                        double threshold, int do_shift) {
    REQUIRE_INIT();
    if (!r) return fail(DDD_EINVAL, "receptor is NULL");
    if (int rc = check_csr(offsets, n_lig, lig_xyzq, nullptr)) return rc;
    if (n_lig == 0) return DDD_OK;
    const Dev &dev = g_devs[0];
    CUDA_TRY(cudaSetDevice(dev.id));
    const int64_t n_atoms = offsets[n_lig];
    long long *d_off = nullptr, *d_li = nullptr, *d_pj = nullptr; double *d_lig = nullptr, *d_d = nullptr;
    auto cleanup = [&]() { dev_free(d_off); dev_free(d_li); dev_free(d_pj); dev_free(d_lig); dev_free(d_d); };
    int rc;
    if ((rc = dev_alloc(&d_off, (size_t)n_lig + 1)) || (rc = dev_alloc(&d_lig, (size_t)n_atoms * 4)) ||
        (rc = dev_alloc(&d_d, (size_t)n_lig)) || (rc = dev_alloc(&d_li, (size_t)n_lig)) || (rc = dev_alloc(&d_pj, (size_t)n_lig))) { cleanup(); return rc; }
    auto run = [&]() -> int {
        CUDA_TRY(cudaMemcpyAsync(d_off, offsets, sizeof(int64_t) * ((size_t)n_lig + 1), cudaMemcpyHostToDevice, dev.stream));
        if (n_atoms > 0) CUDA_TRY(cudaMemcpyAsync(d_lig, lig_xyzq, sizeof(double) * 4 * (size_t)n_atoms, cudaMemcpyHostToDevice, dev.stream));
        ClosestArgs a;
        a.rec = r->view(0); a.offsets = d_off; a.lig_xyzq = d_lig; a.out_dist = d_d; a.out_lig_atom = d_li; a.out_prot_atom = d_pj;
        a.threshold = threshold; a.do_shift = do_shift;
        closest_atom_kernel<<<(unsigned)n_lig, kThreads, 0, dev.stream>>>(a);
        CUDA_TRY(cudaGetLastError());
        if (out_distance) CUDA_TRY(cudaMemcpyAsync(out_distance, d_d, sizeof(double) * (size_t)n_lig, cudaMemcpyDeviceToHost, dev.stream));
        if (out_lig_atom) CUDA_TRY(cudaMemcpyAsync(out_lig_atom, d_li, sizeof(int64_t) * (size_t)n_lig, cudaMemcpyDeviceToHost, dev.stream));
        if (out_prot_atom) CUDA_TRY(cudaMemcpyAsync(out_prot_atom, d_pj, sizeof(int64_t) * (size_t)n_lig, cudaMemcpyDeviceToHost, dev.stream));
        if (copy_back && n_atoms > 0) CUDA_TRY(cudaMemcpyAsync(lig_xyzq, d_lig, sizeof(double) * 4 * (size_t)n_atoms, cudaMemcpyDeviceToHost, dev.stream));
        CUDA_TRY(cudaStreamSynchronize(dev.stream));
        return DDD_OK;
    };
    rc = run();
    cleanup();
    return rc;
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
    const Dev &dev = g_devs[0];
    CUDA_TRY(cudaSetDevice(dev.id));
    const int64_t n_atoms = offsets[n_lig];
    long long *d_off = nullptr; double *d_lig = nullptr;
    auto cleanup = [&]() { dev_free(d_off); dev_free(d_lig); };
    int rc;
    if ((rc = dev_alloc(&d_off, (size_t)n_lig + 1)) || (rc = dev_alloc(&d_lig, (size_t)n_atoms * 4))) { cleanup(); return rc; }
    auto run = [&]() -> int {
        CUDA_TRY(cudaMemcpyAsync(d_off, offsets, sizeof(int64_t) * ((size_t)n_lig + 1), cudaMemcpyHostToDevice, dev.stream));
        CUDA_TRY(cudaMemcpyAsync(d_lig, lig_xyzq_inout, sizeof(double) * 4 * (size_t)n_atoms, cudaMemcpyHostToDevice, dev.stream));
        randomize_pose_kernel<<<(unsigned)n_lig, kThreads, 0, dev.stream>>>(d_off, d_lig, seed, first_ligand_index);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(lig_xyzq_inout, d_lig, sizeof(double) * 4 * (size_t)n_atoms, cudaMemcpyDeviceToHost, dev.stream));
        CUDA_TRY(cudaStreamSynchronize(dev.stream));
        return DDD_OK;
    };
    rc = run();
    cleanup();
    return rc;
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
    if (!s) return fail(DDD_EINVAL, "screen is NULL");
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
    if (!s) return fail(DDD_EINVAL, "screen is NULL");
    if (slots < 0 || slots > kSmMaxSlots) return fail(DDD_EINVAL, "chains per SM must be 0..%d (got %d)", kSmMaxSlots, slots);
    s->slots_override = slots;
    if (slots > 0) s->threads_override = 0;
    return DDD_OK;
}

int32_t ddd_screen_chain_slots(const ddd_screen *s) {
    if (!s || s->devs.empty()) return 0;
    return s->devs[0].slots;
}

int ddd_screen_sync(ddd_screen *s) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    return screen_sync_impl(s);
}

double ddd_screen_last_kernel_ms(const ddd_screen *s) {
    if (!s) return -1.0;
    float mx = 0.f;
    for (auto &d : s->devs) mx = std::max(mx, d.last_ms);
    return (double)mx;
}

int ddd_screen_download(ddd_screen *s, double *out_xyzq, ddd_chain_stats *out_stats) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    return screen_download_impl(s, out_xyzq, out_stats, nullptr);
}

// CompareRMSD (validateResults.go:40-45) for every chain of the screen, device-resident: CalculateRMSD(final pose of the
// chain, start pose of its ligand).  out_rmsd[n_chains]; the kernel time is returned through out_kernel_ms (nullable).
int ddd_screen_rmsd(ddd_screen *s, double *out_rmsd, double *out_kernel_ms) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (!s) return fail(DDD_EINVAL, "screen is NULL");
    if (!s->launched) return fail(DDD_EINVAL, "screen has not been run");
    if (!out_rmsd && s->n_chains > 0) return fail(DDD_EINVAL, "out_rmsd is NULL");
    for (auto &d : s->devs) {
        const Dev &dev = g_devs[d.di];
        CUDA_TRY(cudaSetDevice(dev.id));
        const int64_t n_local = d.c1 - d.c0;
        if (!d.d_rmsd) { if (int rc = dev_alloc(&d.d_rmsd, (size_t)n_local)) return rc; }
        CUDA_TRY(cudaEventRecord(dev.ev0, dev.stream));
        rmsd_chains_kernel<<<(unsigned)((n_local + 7) / 8), 256, 0, dev.stream>>>(d.d_offsets, d.d_lig, d.atom_base, d.d_out, d.out_base,
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

// Cutoff extension: receptor atoms actually paired with each ligand atom by chain c's fp32 sums, summed over its
// evaluations (32 x the units its poses' cell-list queries returned).  Zero-filled when the last run had no cutoff.
int ddd_screen_pair_units(ddd_screen *s, int64_t *out_receptor_atoms) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    REQUIRE_INIT();
    if (!s) return fail(DDD_EINVAL, "screen is NULL");
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
    if (!s) return fail(DDD_EINVAL, "screen is NULL");
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
        if (int rc = set_smem(mc_sm_kernel<true>, smem)) return rc;
        CUDA_TRY(cudaFuncGetAttributes(&fa, mc_sm_kernel<true>));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, mc_sm_kernel<true>, kSmThreads, smem));
    } else {
        if (int rc = set_smem(mc_sm_kernel<false>, smem)) return rc;
        CUDA_TRY(cudaFuncGetAttributes(&fa, mc_sm_kernel<false>));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, mc_sm_kernel<false>, kSmThreads, smem));
    }
    if (threads) *threads = kSmThreads;
    if (regs) *regs = fa.numRegs;
    if (static_smem) *static_smem = (int32_t)fa.sharedSizeBytes;
    if (max_blocks_per_sm_at_nl40) *max_blocks_per_sm_at_nl40 = blocks;
    return DDD_OK;
}

}  // extern "C"
