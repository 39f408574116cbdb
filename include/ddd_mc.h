/*
 * ddd_mc.h -- C ABI of libddd_mc.so: the B200-native (sm_100a CUDA) Metropolis Monte-Carlo
 * binding-energy hot path of Simran-Sodhi/drug-design-dynamics (metropolisMethod/).
 *
 * The reference has no FFI: its boundary is a set of Go functions in `package main`
 * (SURVEY.md 8b).  Each entry point below names the Go function(s) whose body it replaces;
 * the cgo overlay that binds them is go/metropolis_cuda.go and is described in INTEGRATION.md.
 * All citations are relative to /root/reference/metropolisMethod/.
 *
 * Conventions
 *  - Plain pointers and sizes only.  Every function returns 0 (DDD_OK) or a negative DDD_E*
 *    code; ddd_last_error() returns a thread-local message.  The Go shim panics on non-zero,
 *    which is what the reference's Check() does (main.go:179-183).
 *  - Atoms travel as `xyzq`: 4 float64 per atom (X, Y, Z, Charge) -- byte-identical to Go's
 *    `[]Atom` (datatypes.go:18-25), so cgo passes &mol.atoms[0] without repacking.
 *  - Ligand batches are CSR: offsets[n_lig+1] in atoms, offsets[0] == 0.
 *  - Chains: chain c = ligand (c / replicas), replica (c % replicas); it draws from the
 *    Philox4x32-10 stream (seed, first_chain_id + c).  Chain outputs are chain-major: chain c's atoms start at
 *    atom index replicas*offsets[lig] + replica*nL(lig).
 *  - No CPU fallback: without a usable CUDA device every compute entry point fails with
 *    DDD_ENODEVICE.  The caller's buffers are host memory; nothing is retained after return
 *    except inside the opaque handles.
 *  - Thread-safe: calls are serialised on an internal mutex and select their device on entry
 *    (goroutines migrate between OS threads).
 *  - ddd_init with the device list already in use is a no-op; ddd_shutdown (and ddd_init with a different list) makes every
 *    receptor / screen created before it STALE: entry points refuse stale handles, the destroy functions still free them.
 */
#ifndef DDD_MC_H
#define DDD_MC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDD_OK            0
#define DDD_EINVAL       -1   /* bad argument (NULL pointer, negative size, num_procs <= 0 ...) */
#define DDD_ENODEVICE    -2   /* no CUDA device / ddd_init not called or failed */
#define DDD_ECUDA        -3   /* CUDA runtime error (message in ddd_last_error) */
#define DDD_ELIMIT       -4   /* size beyond a kernel limit (ligand larger than DDD_MAX_LIGAND_ATOMS) */

#define DDD_MAX_LIGAND_ATOMS 2048

/* precision of the pair-energy sum inside the chains / ddd_energy_batch */
#define DDD_PRECISION_F32  0  /* fp32 pair terms (rsqrt), fp64 accumulation of per-tile partials */
#define DDD_PRECISION_F64  1  /* fp64 terms with the reference's exact per-term arithmetic */

/* proposal */
#define DDD_MOVE_REFERENCE 0  /* metropolis.go:154-208: per-atom jitter, Rodrigues about the ORIGIN */
#define DDD_MOVE_RIGID     1  /* extension: centroid translation + quaternion rotation (no reference row) */

/* chain status bits */
#define DDD_STATUS_RETRY_LIMIT      1  /* collision retry bound hit (reference would spin, :155,:173) */
#define DDD_STATUS_STREAM_EXHAUSTED 2  /* recorded uniform stream ran dry */

/* ddd_screen_argmin rules */
#define DDD_ARGMIN_SAVE_MIN_LIGAND 0  /* plotting.go:109-121: minIndex 0, minEnergy 0.0, strict < (only a negative energy wins) */
#define DDD_ARGMIN_RSHINY          1  /* main.go:35-36, 51-54: minEnergy 1e20, strict < */

/* Philox stream-id spaces */
#define DDD_STREAM_PARENT_BASE    0x4000000000000000ull  /* + ligand index: replica-pick draws (:105) */
#define DDD_STREAM_RANDOMIZE_BASE 0x2000000000000000ull  /* + ligand index: RandomizeLigandPose draws */

/* datatypes.go:6-12 and the literals in metropolis.go; ddd_params_default() fills the reference values */
typedef struct ddd_params {
    double  k_coulomb;          /* K = 9e9            datatypes.go:6  */
    double  threshold;          /* THRESHOLD = 5      datatypes.go:9  */
    double  min_distance;       /* MINDISTANCE = 0.5  datatypes.go:10 */
    double  max_angle;          /* MAXANGLE = 0.79    datatypes.go:11 */
    double  jitter_width;       /* 0.1                metropolis.go:158-160,176-178 */
    double  clamp_distance;     /* 1e-6               metropolis.go:255-257 */
    double  rigid_translation;  /* RIGID only: full width of the per-axis uniform translation (0.5 A) */
    double  rigid_angle;        /* RIGID only: max |angle| (0.79 rad) */
    int64_t max_retries;        /* collision-retry bound per step (1024); <= 0: 2^62 */
    int32_t precision;          /* DDD_PRECISION_*  (default F32) */
    int32_t move_mode;          /* DDD_MOVE_*       (default REFERENCE) */
    double  cutoff;             /* EXTENSION, default 0 = off (the reference sums every pair, metropolis.go:249-259).
                                   > 0: shifted-Coulomb cutoff radius in A -- a pair with d < cutoff contributes
                                   K q_p q_l (1/max(d, clamp) - 1/cutoff), a pair beyond it nothing.  The chains then
                                   evaluate only the receptor cells within reach of the pose (uniform-grid cell list
                                   built at ddd_receptor_create); needs the SM-persistent engine. */
    double  lj_scale;           /* EXTENSION, default 0 = off (the reference has no Lennard-Jones term).  != 0: every pair
                                   of the Coulomb sum also adds lj_scale * 4 sqrt(eps_p eps_l) (x^6 - x^3),
                                   x = min(((sigma_p + sigma_l)/2)^2 / d^2, lj_cap).  Needs ddd_receptor_set_lj and
                                   ddd_screen_set_lj (the Go data model carries no atom types: datatypes.go:18-21). */
    double  lj_cap;             /* cap on (sigma/d)^2 (4: the 12-6 core is frozen below d = sigma/2) */
} ddd_params;

/* per-chain counters returned by ddd_run_chains (no reference counterpart; SURVEY 5 "metrics") */
typedef struct ddd_chain_stats {
    int64_t steps;         /* iterations of metropolis.go:119-134 executed */
    int64_t accepted;      /* AcceptMove == true */
    int64_t downhill;      /* accepted on newE < curE (no draw, :142-144) */
    int64_t retries;       /* collision-rejected proposal attempts (:162,:180) */
    int64_t draws;         /* uniforms consumed from the chain's stream */
    int64_t status;        /* DDD_STATUS_* bits */
    double  final_energy;  /* CalculateEnergy(protein, final pose) evaluated in fp64 (:61) */
    double  start_energy;  /* CalculateEnergy(protein, start pose) in fp64 (:96,:118) */
    double  chain_energy;  /* the chain's own currentEnergy at exit (fp32-summed in F32 mode) */
} ddd_chain_stats;

typedef struct ddd_receptor ddd_receptor;   /* protein Molecule resident on every initialised device */
typedef struct ddd_screen   ddd_screen;     /* ligand x replica chain batch resident on the devices */

/* ---- library ---- */
int         ddd_init(const int *devices, int n_devices);  /* devices == NULL: every visible device */
int         ddd_shutdown(void);
int         ddd_num_devices(void);                         /* devices in use (0 before ddd_init) */
const char *ddd_last_error(void);
const char *ddd_version(void);
void        ddd_params_default(ddd_params *p);

/* ---- receptor (protein Molecule, datatypes.go:14-16) ---- */
int     ddd_receptor_create(const double *xyzq, int64_t n_atoms, ddd_receptor **out);
int     ddd_receptor_destroy(ddd_receptor *r);
int64_t ddd_receptor_num_atoms(const ddd_receptor *r);
/* Content-keyed receptor cache for the scalar Go functions that receive the protein on every call (CalculateEnergy
 * metropolis.go:247, FindClosestAtomDistance :290, ShiftLigandCloserByThreshold :267, SimulateEnergyMinimizationParallel :94
 * as RShinyAppMain calls it per ligand, main.go:42-48): acquire returns the resident receptor of an identical xyzq array
 * (length + hash, compared in full) or builds one; release drops the caller's reference.  Up to 8 receptors stay cached
 * (least recently used evicted); handles from acquire must not be passed to ddd_receptor_destroy. */
int     ddd_receptor_acquire(const double *xyzq, int64_t n_atoms, ddd_receptor **out);
int     ddd_receptor_release(ddd_receptor *r);
/* LJ extension: per-atom sigma (A) and epsilon of the receptor (n_atoms entries each); NULL, NULL clears them */
int     ddd_receptor_set_lj(ddd_receptor *r, const double *sigma, const double *epsilon);

/* ---- CalculateEnergy (metropolis.go:247-262), batched over ligand poses ----
 * out_energy[n_lig]; out_abs_sum (nullable) receives sum|term| (fp64) for conditioning reports. */
int ddd_energy_batch(const ddd_receptor *r, const int64_t *offsets, const double *lig_xyzq,
                     int64_t n_lig, const ddd_params *p, double *out_energy, double *out_abs_sum);

/* ---- SimulateEnergyMinimizationOneProc (metropolis.go:116-136), batched ----
 * Runs n_lig*replicas independent chains of steps_per_chain iterations, each from a private
 * copy of its ligand's start pose, entirely on device (proposal :154-208, collision test
 * :229-242, energy :247-262, AcceptMove :141-149).
 * recorded / recorded_offsets (both nullable): per-chain recorded uniform streams replacing
 *   Philox (chain c reads recorded[recorded_offsets[c] .. recorded_offsets[c+1])).
 * out_xyzq: (replicas * offsets[n_lig]) atoms x 4 doubles, chain-major final poses.
 * out_stats[n_chains].  out_accept_trace (nullable): n_chains*steps_per_chain bytes, 1 = accepted. */
int ddd_run_chains(const ddd_receptor *r, const int64_t *offsets, const double *lig_xyzq,
                   int64_t n_lig, int32_t replicas, int64_t steps_per_chain, int32_t rotate,
                   double temperature, const ddd_params *p, uint64_t seed, uint64_t first_chain_id,
                   const double *recorded, const int64_t *recorded_offsets,
                   double *out_xyzq, ddd_chain_stats *out_stats, uint8_t *out_accept_trace);

/* ---- SimulateEnergyMinimizationParallel (:94-111) + the final CalculateEnergy (:61, :19, main.go:48),
 *      batched over ligands = SimulateMultipleLigandsParallel (:27-67) ----
 * num_procs replica chains of iterations/num_procs steps per ligand, then the sequential
 * AcceptMove pick over the replicas' end poses in replica-index order (draws from stream
 * (seed, DDD_STREAM_PARENT_BASE + first_ligand_index + i)); chain (i, r) draws from stream
 * (seed, (first_ligand_index + i) * num_procs + r), so a ligand list split over calls or ranks
 * sees the streams of one call.
 * out_xyzq: offsets[n_lig] atoms x 4 doubles (same CSR as the input); out_energy[n_lig];
 * out_chain_stats (nullable) [n_lig*num_procs]; out_picked (nullable) [n_lig]: replica kept, -1 = start pose. */
int ddd_minimize_batch(const ddd_receptor *r, const int64_t *offsets, const double *lig_xyzq,
                       int64_t n_lig, int64_t iterations, int32_t rotate, double temperature,
                       int32_t num_procs, const ddd_params *p, uint64_t seed,
                       uint64_t first_ligand_index,
                       double *out_xyzq, double *out_energy,
                       ddd_chain_stats *out_chain_stats, int64_t *out_picked);

/* Chain status of this THREAD's last ddd_minimize_batch / ddd_run_chains call: returns the OR of the DDD_STATUS_* bits over
 * its chains; *out_retry_limited_chains (nullable) = how many stopped at the collision-retry bound, where the reference
 * would still be retrying (metropolis.go:155, :173).  The Go / C++ / Python mirrors warn when this is non-zero. */
int64_t ddd_last_chain_status(int64_t *out_retry_limited_chains);

/* ---- CalculateRMSD (validateResults.go:50-65), batched over (simulated, reference) pairs ----
 * atom_stride: doubles per atom in sim/ref (4 = Go []Atom layout, 3 = packed xyz).
 * out_rmsd[n_pairs]; a pair with zero atoms yields NaN (0/0, as in Go). */
int ddd_rmsd_batch(const int64_t *offsets, const double *sim, const double *ref,
                   int32_t atom_stride, int64_t n_pairs, double *out_rmsd);

/* ---- FindClosestAtomDistance (:290-304) / ShiftLigandCloserByThreshold (:267-285), batched ----
 * out_lig_atom / out_prot_atom: indices of the first minimum in ligand-outer, protein-inner order. */
int ddd_closest_atom_batch(const ddd_receptor *r, const int64_t *offsets, const double *lig_xyzq,
                           int64_t n_lig, double *out_distance, int64_t *out_lig_atom,
                           int64_t *out_prot_atom);
int ddd_shift_closer_batch(const ddd_receptor *r, const int64_t *offsets, double *lig_xyzq_inout,
                           int64_t n_lig, double threshold);

/* ---- RandomizeLigandPose (validateResults.go:70-79), batched; ligand i draws from stream
 *      (seed, DDD_STREAM_RANDOMIZE_BASE + first_ligand_index + i) ---- */
int ddd_randomize_pose_batch(const int64_t *offsets, double *lig_xyzq_inout, int64_t n_lig,
                             uint64_t seed, uint64_t first_ligand_index);

/* ---- device-resident screen: the same chains with inputs/outputs kept in HBM ----
 * create uploads the start poses and shards the chains over the devices (contiguous blocks,
 * balanced by sum nL); run launches steps_per_chain iterations from the start poses on every
 * device asynchronously; sync waits; last_kernel_ms = max over devices of the CUDA-event time
 * of the chain kernel on its launch stream; download copies final poses + stats to the host. */
int    ddd_screen_create(const ddd_receptor *r, const int64_t *offsets, const double *lig_xyzq,
                         int64_t n_lig, int32_t replicas, uint64_t first_chain_id, ddd_screen **out);
/* Engine of the chain kernel for subsequent runs.  Default (both 0): the SM-persistent engine -- one 16-warp
 * CTA per SM keeps `slots` chains resident and every warp takes pair-sum work items of any of them -- with the
 * chains per SM chosen from the chain count.  set_chain_slots(k > 0) fixes k (1..16; clipped to what fits in
 * shared memory).  set_chain_threads(32|64|128|256) selects the one-CTA-per-chain engine with that CTA width;
 * set_chain_threads(0) returns to the default.  Results do not depend on the engine in F64 precision. */
int    ddd_screen_set_chain_threads(ddd_screen *s, int32_t threads);
int32_t ddd_screen_chain_threads(const ddd_screen *s);   /* CTA width used by the last run (512: SM-persistent engine) */
int    ddd_screen_set_chain_slots(ddd_screen *s, int32_t slots);
int32_t ddd_screen_chain_slots(const ddd_screen *s);     /* chains per SM of the last run (0: one CTA per chain) */
/* Chain segments of the SM-persistent engine: every `steps` steps (rounded up to a power of two) a chain leaves its slot --
 * pose, energies, stream position and counters parked in its own output records -- and queues behind the launch's other
 * chains; any SM resumes it.  The SMs then run out of work one segment apart instead of one chain apart (a launch lasts as
 * long as its slowest SM).  -1 (default): automatic, ~32 segments per chain when the launch has more chains than slots;
 * 0: whole chains.  Results never depend on it (the parked state is exact).  No reference counterpart: goroutines are
 * scheduled by the Go runtime (metropolis.go:43, :100).  The environment variable DDD_SEGMENT_STEPS sets the default of
 * every screen created afterwards, the one-shot entry points' internal screens included (a testing knob). */
int    ddd_screen_set_segment_steps(ddd_screen *s, int64_t steps);
int64_t ddd_screen_segment_steps(const ddd_screen *s);   /* segment length of the last run (0: whole chains) */
int    ddd_screen_run(ddd_screen *s, int64_t steps_per_chain, int32_t rotate, double temperature,
                      const ddd_params *p, uint64_t seed);
int    ddd_screen_sync(ddd_screen *s);
double ddd_screen_last_kernel_ms(const ddd_screen *s);
int    ddd_screen_download(ddd_screen *s, double *out_xyzq, ddd_chain_stats *out_stats);
/* CompareRMSD (validateResults.go:40-45) for every chain of the screen without leaving the device:
 * out_rmsd[c] = CalculateRMSD(final pose of chain c, start pose of its ligand) (validateResults.go:50-65).
 * out_kernel_ms (nullable): CUDA-event time of the RMSD kernel, max over devices. */
int    ddd_screen_rmsd(ddd_screen *s, double *out_rmsd, double *out_kernel_ms);
/* The same values as left behind by the chain kernel's own epilogue (SM-persistent engine): no extra pass, same bits. */
int    ddd_screen_fused_rmsd(ddd_screen *s, double *out_rmsd);

/* ---- resident pipeline: TestMethodRMSD / CompareRMSD (validateResults.go:14-45), SimulateMultipleLigands' shift
 *      (metropolis.go:14-16) and SaveMinimumEnergyLigand's choice (plotting.go:109-121) without leaving HBM ----
 * ddd_screen_randomize: RandomizeLigandPose (validateResults.go:70-79) on the resident start poses, in place; ligand i draws
 *   from stream (seed, DDD_STREAM_RANDOMIZE_BASE + first_ligand_index + i).  The poses as uploaded stay resident as the RMSD
 *   reference (CompareRMSD takes its CopyLigand before randomising, :41).
 * ddd_screen_shift_closer: ShiftLigandCloserByThreshold (metropolis.go:267-285) on the resident start poses.
 * ddd_screen_download_start: the start poses as resident now (out_xyzq: offsets[n_lig] atoms x 4).
 * ddd_screen_pick: the replica pick of SimulateEnergyMinimizationParallel (:102-109) over the last run's chains, replica-index
 *   order, draws from (seed, DDD_STREAM_PARENT_BASE + first_ligand_index + i); per ligand: picked pose (CSR like the input),
 *   energy = CalculateEnergy(picked pose) in fp64 (:61), replica kept (-1 = start pose), CalculateRMSD(picked pose, reference
 *   pose) = CompareRMSD's return value.  Every out pointer is nullable: the results stay resident either way.
 * ddd_screen_argmin: index and energy SaveMinimumEnergyLigand (rule 0) / RShinyAppMain (rule 1) would select among the
 *   picked per-ligand energies. */
int    ddd_screen_randomize(ddd_screen *s, uint64_t seed, uint64_t first_ligand_index);
int    ddd_screen_shift_closer(ddd_screen *s, double threshold);
int    ddd_screen_download_start(ddd_screen *s, double *out_xyzq);
int    ddd_screen_pick(ddd_screen *s, double temperature, uint64_t seed, uint64_t first_ligand_index,
                       double *out_xyzq, double *out_energy, int64_t *out_picked, double *out_rmsd);
int    ddd_screen_argmin(ddd_screen *s, int32_t rule, int64_t *out_index, double *out_energy);
/* EXTENSION (no reference row; north_star asks for per-ligand MINIMUM energies and poses, the reference returns the FINAL pose,
 * metropolis.go:110,135): when on, every chain also keeps the lowest-energy pose it visited.  Equal to the final pose while
 * the sampler is greedy.  download_best: out_xyzq chain-major like ddd_screen_download, out_energy[n_chains] =
 * the chain's own energy of that pose (fp32-summed in F32 mode).  Needs the SM-persistent engine. */
int    ddd_screen_set_track_best(ddd_screen *s, int32_t on);
int    ddd_screen_download_best(ddd_screen *s, double *out_xyzq, double *out_energy);
/* LJ extension: per-atom sigma / epsilon of the screen's ligands, in the CSR order of ddd_screen_create
 * (offsets[n_lig] entries each); used by runs whose params.lj_scale != 0 */
int    ddd_screen_set_lj(ddd_screen *s, const double *sigma, const double *epsilon);
/* Cutoff extension (params.cutoff > 0): out[c] = receptor atoms that chain c's pair sums visited, summed over its
 * evaluations (what the cell-list queries of its poses returned); x nL(c) = pair terms evaluated.  Zeros without cutoff. */
int    ddd_screen_pair_units(ddd_screen *s, int64_t *out_receptor_atoms);
int    ddd_screen_destroy(ddd_screen *s);

/* ---- sharding plan (pure host logic, works without a device): chains
 *      [out_begin[k], out_begin[k+1]) go to shard k; contiguous blocks (the reference's output order,
 *      metropolis.go:34-49) balanced by sum(nL + 1).  out_begin has n_shards + 1 entries. ---- */
int ddd_shard_plan(const int64_t *offsets, int64_t n_lig, int32_t replicas, int32_t n_shards,
                   int64_t *out_begin);

/* ---- introspection for the roofline report ---- */
int ddd_device_info(int ordinal, int32_t *sm_count, int32_t *clock_khz, int32_t *cc_major,
                    int32_t *cc_minor, char *name, int32_t name_cap);
int ddd_chain_kernel_info(int32_t precision, int32_t *threads, int32_t *regs, int32_t *static_smem,
                          int32_t *max_blocks_per_sm_at_nl40);

#ifdef __cplusplus
}
#endif
#endif /* DDD_MC_H */
