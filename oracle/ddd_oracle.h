/*
 * ddd_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE).
 *
 * A plain-C float64 restatement of the Metropolis Monte-Carlo hot path of
 * Simran-Sodhi/drug-design-dynamics, following metropolisMethod/metropolis.go,
 * datatypes.go and validateResults.go line by line (each function cites the
 * file:line it follows).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the CUDA product
 * (drug-design-dynamics_b200/csrc) never links, includes or calls it.
 *
 * PARITY PINNING.  The reference is Go; this image has no Go toolchain, so the
 * reference cannot be executed here (oracle/_ref does not exist).  The oracle
 * is pinned against every known-answer the reference's own tests hold for this
 * path (metropolis_test.go: TestDistance, TestFindClosestAtomDistance exact;
 * TestRotateAtom 1e-6; the energy-sign, AcceptMove, collision-free, rotate,
 * shift, copy and shape tests) -- see tests/test_oracle_reference_kats.py.
 * Energy magnitudes, acceptance statistics, RMSD values and the replica pick
 * are NOT pinned by any reference test or golden file: for those the status
 * is "parity unpinned" and the oracle is cross-checked only against a second,
 * independently written numpy restatement (tests/_np_restatement.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: Go/amd64 never fuses
 * a*b+c, so contraction must stay off for bit-identical per-term arithmetic).
 */
#ifndef DDD_ORACLE_H
#define DDD_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Go: Atom{Position Position3d{X,Y,Z}; Charge}  (datatypes.go:18-25), 4 x f64 AoS. */
typedef struct { double x, y, z, q; } orc_atom;

/* Simulation constants (datatypes.go:6-12) + the literals buried in metropolis.go. */
typedef struct {
    double k_coulomb;        /* K = 9e9                  datatypes.go:6  */
    double threshold;        /* THRESHOLD = 5            datatypes.go:9  */
    double min_distance;     /* MINDISTANCE = 0.5        datatypes.go:10 */
    double max_angle;        /* MAXANGLE = 0.79          datatypes.go:11 */
    double jitter_width;     /* literal 0.1              metropolis.go:158-160,176-178 */
    double clamp_distance;   /* literal 1e-6             metropolis.go:255-257 */
    int    move_mode;        /* 0 = REFERENCE (per-atom jitter + rotation about the origin)
                                1 = RIGID extension (centroid translation + quaternion rotation);
                                    no counterpart in the reference -> parity unpinned */
    double rigid_translation;/* RIGID only: full width of the uniform translation per axis */
    double rigid_angle;      /* RIGID only: max |angle| in rad */
    int64_t max_retries;     /* bound on the collision retry loop (reference: unbounded,
                                metropolis.go:155,173).  <=0 means unbounded. */
    double cutoff;           /* EXTENSION (no reference row, parity unpinned): > 0 = shifted-Coulomb cutoff in A:
                                pairs with d < cutoff add K q q (1/max(d,clamp) - 1/cutoff), the others nothing.
                                0 (default) = the reference's sum over every pair */
    /* EXTENSION (no reference row, parity unpinned): Lennard-Jones 12-6 on top of the Coulomb term, off when lj_scale == 0.
       Every pair that enters the Coulomb sum also adds lj_scale * 4 sqrt(eps_p eps_l) (x^6 - x^3) with
       x = min(((sigma_p + sigma_l) / 2)^2 / d^2, lj_cap) (Lorentz-Berthelot mixing; the cap keeps the core finite).
       prot_lj / lig_lj: (sigma, epsilon) per atom, interleaved; lig_lj belongs to the ligand passed with these params. */
    double lj_scale, lj_cap;
    const double *prot_lj, *lig_lj;
} orc_params;

void orc_params_default(orc_params *p);

/* Uniform [0,1) stream: either counter-based Philox4x32-10 or a recorded array. */
typedef struct {
    int      kind;           /* 0 = philox, 1 = recorded */
    uint64_t seed;           /* philox key */
    uint64_t stream_id;      /* philox counter high half (chain id) */
    uint64_t pos;            /* next draw index */
    const double *rec;       /* recorded uniforms */
    uint64_t rec_len;
    int      exhausted;      /* set when a recorded stream runs dry (draw returns 0.5) */
} orc_stream;

void   orc_stream_philox(orc_stream *s, uint64_t seed, uint64_t stream_id);
void   orc_stream_recorded(orc_stream *s, const double *u, uint64_t n);
double orc_stream_next(orc_stream *s);
void   orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_philox_u01(uint64_t seed, uint64_t stream_id, uint64_t index);

/* Philox stream-id spaces (the product uses the same convention, include/ddd_mc.h). */
#define ORC_STREAM_PARENT_BASE    0x4000000000000000ull  /* + ligand index: replica pick draws */
#define ORC_STREAM_RANDOMIZE_BASE 0x2000000000000000ull  /* + ligand index: RandomizeLigandPose */

/* Per-chain counters (not in the reference; used to compare with the device counters). */
typedef struct {
    int64_t steps;           /* iterations executed */
    int64_t accepted;        /* AcceptMove == true */
    int64_t downhill;        /* accepted through newE < curE (no draw) */
    int64_t retries;         /* collision-rejected proposal attempts */
    int64_t draws;           /* uniforms consumed */
    int64_t status;          /* bit0: retry bound hit, bit1: recorded stream exhausted */
    double  final_energy;    /* currentEnergy after the last step */
    double  start_energy;
} orc_chain_stats;

/* ---- datatypes.go:33-60 ---- */
double orc_magnitude(const double v[3]);
void   orc_normalize(double v[3]);
double orc_dot(const double a[3], const double b[3]);

/* ---- metropolis.go ---- */
double orc_distance(const double a[3], const double b[3]);                        /* :309 */
void   orc_rotate_atom(const double pos[3], const double axis[3], double theta,
                       double out[3]);                                             /* :213 */
int    orc_is_collision_free(const orc_atom *lig, int64_t nl, double min_distance);/* :229 */
double orc_calculate_energy(const orc_atom *prot, int64_t np,
                            const orc_atom *lig, int64_t nl, const orc_params *p); /* :247 */
double orc_calculate_energy_abs(const orc_atom *prot, int64_t np,
                                const orc_atom *lig, int64_t nl, const orc_params *p,
                                double *abs_sum);   /* same + sum|term| for conditioning */
int    orc_accept_move(double cur_e, double new_e, double temperature, orc_stream *s); /* :141 */
void   orc_rotate_ligand(orc_atom *lig, int64_t nl, double max_angle, orc_stream *s);   /* :189 */
int64_t orc_jitter_ligand(orc_atom *lig, int64_t nl, const orc_params *p, orc_stream *s);            /* :154 */
int64_t orc_jitter_and_rotate_ligand(orc_atom *lig, int64_t nl, const orc_params *p, orc_stream *s); /* :172 */
double orc_find_closest_atom_distance(const orc_atom *lig, int64_t nl,
                                      const orc_atom *prot, int64_t np,
                                      int64_t *lig_idx, int64_t *prot_idx);        /* :290 */
void   orc_shift_ligand_closer(orc_atom *lig, int64_t nl, const orc_atom *prot,
                               int64_t np, double threshold);                      /* :267 */
void   orc_rigid_move(orc_atom *lig, int64_t nl, const orc_params *p, orc_stream *s); /* extension */

/* SimulateEnergyMinimizationOneProc (:116-136): one chain, in place on a private copy.
 * accept_trace (nullable, length >= iterations): 1 accepted / 0 rejected per step.
 * energy_trace (nullable): newEnergy proposed at each step. */
void orc_chain_one_proc(const orc_atom *prot, int64_t np, orc_atom *lig, int64_t nl,
                        int64_t iterations, int rotate, double temperature,
                        const orc_params *p, orc_stream *s, orc_chain_stats *stats,
                        uint8_t *accept_trace, double *energy_trace);

/* SimulateEnergyMinimizationParallel (:94-111): num_procs replica chains of
 * iterations/num_procs steps from private copies of the start pose, then the
 * sequential AcceptMove pick in replica-index order (the reference's order is
 * channel-arrival order -- nondeterministic).  Chain r of ligand
 * `lig_index` draws from philox stream (seed, lig_index*num_procs + r); the
 * parent pick draws from stream (seed, ORC_STREAM_PARENT_BASE + lig_index).
 * chain_poses (nullable): num_procs*nl atoms, each chain's final pose.
 * Returns 0, or -1 if num_procs <= 0 (Go: integer divide by zero panic). */
/* EXTENSION: orc_chain_one_proc + best-ever (lowest-energy) pose tracking; no reference row */
void orc_chain_one_proc_best(const orc_atom *prot, int64_t np, orc_atom *lig, int64_t nl,
                             int64_t iterations, int rotate, double temperature,
                             const orc_params *p, orc_stream *s, orc_chain_stats *stats,
                             orc_atom *best_pose, double *best_energy);
/* plotting.go:109-121 (rule 0) / main.go:35-36,51-54 (rule 1): index the reference would select */
int64_t orc_min_energy_index(const double *energy, int64_t n, int rule, double *min_energy_out);
int orc_minimize_parallel(const orc_atom *prot, int64_t np, orc_atom *lig, int64_t nl,
                          int64_t iterations, int rotate, double temperature, int num_procs,
                          const orc_params *p, uint64_t seed, uint64_t lig_index,
                          orc_chain_stats *chain_stats, orc_atom *chain_poses,
                          int64_t *picked_replica);

/* SimulateMultipleLigandsParallel (:27-51) / SimulateMultipleLigands (:11-22) over a CSR batch.
 * offsets[n_lig+1] in atoms; lig holds the start poses and receives the minimised poses;
 * energies[n_lig] = CalculateEnergy(protein, minimised) (:61).  shift_first=1 restates the
 * serial variant's ShiftLigandCloserByThreshold pre-pass (:14-16).  n_threads host threads
 * run ligands concurrently (results do not depend on it). */
int orc_simulate_multiple_ligands(const orc_atom *prot, int64_t np, orc_atom *lig,
                                  const int64_t *offsets, int64_t n_lig, int64_t iterations,
                                  int rotate, double temperature, int num_procs,
                                  const orc_params *p, uint64_t seed, int shift_first,
                                  int n_threads, double *energies, orc_chain_stats *chain_stats);

/* Block split of SimulateMultipleLigandsParallel (:34-42): worker i gets [begin,end). */
void orc_ligand_block(int64_t n_ligands, int num_procs, int worker, int64_t *begin, int64_t *end);

/* ---- validateResults.go ---- */
double orc_calculate_rmsd(const orc_atom *sim, int64_t n_sim, const orc_atom *ref, int64_t n_ref); /* :50; NaN if n_ref < n_sim (Go: index panic) */
void   orc_randomize_ligand_pose(orc_atom *lig, int64_t nl, orc_stream *s);                         /* :70 */
double orc_compare_rmsd(const orc_atom *prot, int64_t np, orc_atom *lig, int64_t nl,
                        int64_t iterations, int rotate, double temperature, int num_procs,
                        const orc_params *p, uint64_t seed, uint64_t lig_index);                    /* :40 */

/* ---- parsing.go:50-106 ---- returns atom count (>=0) or -1 on I/O error; fills up to cap atoms. */
int64_t orc_parse_mol2(const char *path, orc_atom *out, int64_t cap);

/* ---- CPU baseline driver (bench.py cpu_baseline / --impl reference) ----
 * Runs n_chains independent OneProc chains (chain c = ligand c / replicas) on n_threads
 * host threads; returns wall seconds; total_steps receives the steps executed. */
double orc_bench_chains(const orc_atom *prot, int64_t np, const orc_atom *lig,
                        const int64_t *offsets, int64_t n_lig, int replicas,
                        int64_t steps_per_chain, int rotate, double temperature,
                        const orc_params *p, uint64_t seed, int n_threads,
                        int64_t *total_steps, int64_t *total_pairs);

/* Go standard library math.Pow / Frexp / Ldexp / Modf restated (go1.23 src/math; third-party to the reference) */
double orc_go_pow(double x, double y);
double orc_go_frexp(double f, int *exp_out);
double orc_go_ldexp(double frac, int exp_);
double orc_go_modf(double f, double *frac);

const char *orc_build_info(void);

#ifdef __cplusplus
}
#endif
#endif
