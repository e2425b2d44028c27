/*
 * mbgdml_b200 -- C ABI of the B200-native many-body GDML prediction path.
 *
 * The reference (keithgroup/mbGDML) has no FFI: its plugin surface is the Python
 * callable handed to mbePredict (reference mbgdml/mbe.py:521-532, call sites
 * mbe.py:809-811 and 949-951).  This header is the boundary a binding for that
 * callable targets: plain pointers and sizes, no torch / numpy types.
 * INTEGRATION.md shows the ctypes stub that mbgdml/predictors/gdml_predict.py
 * would use.  Each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - every function returns MBG_OK (0) or a negative error code; the message
 *     of the last error on the calling thread is mbg_last_error();
 *   - all floating point is IEEE double; all index arrays are int32 unless noted;
 *   - the caller owns every buffer it passes; nothing is retained after a call
 *     returns except by mbg_model_create (which copies to device memory);
 *   - one in-flight call per context; calls are stream-ordered on the context's
 *     CUDA stream and return after the result buffers are valid (host buffers)
 *     or after the work is enqueued (device buffers, MBG_FLAG_*_ON_DEVICE).
 */
#ifndef MBGDML_B200_H
#define MBGDML_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBG_ABI_VERSION 3

#define MBG_OK 0
#define MBG_ERR_ARG (-1)         /* invalid argument (the Python mirror raises ValueError/AssertionError) */
#define MBG_ERR_CUDA (-2)        /* CUDA runtime failure */
#define MBG_ERR_UNSUPPORTED (-3) /* valid request this build has no kernel for (no CPU fallback exists) */
#define MBG_ERR_NOMEM (-4)

/* Contraction kernel selection (mbg_context_set_contraction). */
#define MBG_CONTRACT_AUTO 0 /* = MBG_CONTRACT_DMMA in this build                                              */
#define MBG_CONTRACT_FMA 1  /* difference form on the FP64 FMA pipe                      (csrc/matern.cuh)     */
#define MBG_CONTRACT_DMMA 2 /* GEMM recast on the FP64 tensor pipe (mma.sync.m8n8k4.f64), persistent per-warp workers that read the
                               fragment count on the device: no host round trip inside a call    (csrc/matern_pw.cuh)  */
#define MBG_CONTRACT_DMMA_WAVE 3 /* the same arithmetic, wave-launched one CTA per query batch   (csrc/matern_mma.cuh) */

/* Candidate generation (mbg_context_set_culling). */
#define MBG_CULL_AUTO 0     /* MBG_CULL_PAIRLIST from 2^21 candidates per job on (where it applies), else MBG_CULL_FULL */
#define MBG_CULL_FULL 1     /* every tuple of utils.gen_combs is unranked and tested, as the reference does (csrc/enumerate.cuh) */
#define MBG_CULL_PAIRLIST 2 /* per-frame pair list first, only tuples of mutually admissible pairs are tested: same accepted
                               set, same order, cost ~ accepted fragments instead of C(n, order)          (csrc/pairlist.cuh) */

#define MBG_MAX_ORDER 8       /* entities per fragment */
#define MBG_MAX_FRAG_ATOMS 24 /* atoms per fragment */
#define MBG_MAX_ALCHEMY 16    /* alchemical scalers per job */

typedef struct mbg_context_s *mbg_context_t;
typedef struct mbg_model_s *mbg_model_t;

/* Acceptance criteria evaluated on device (reference mbgdml/descriptors.py:32-201). */
enum {
  MBG_CRIT_NONE = 0,              /* Criteria is None or cutoff is None: accept everything      */
  MBG_CRIT_COM_DISTANCE_SUM = 1,  /* descriptors.com_distance_sum  (descriptors.py:159-201)       */
  MBG_CRIT_MAX_ATOM_PAIR_DIST = 2 /* descriptors.max_atom_pair_dist (descriptors.py:131-156)      */
};
enum {
  MBG_BOUND_UPPER = 0, /* accept iff v <  hi            (descriptors.py:96-97) */
  MBG_BOUND_LOWER = 1, /* accept iff v >  lo            (descriptors.py:98-99) */
  MBG_BOUND_RANGE = 2  /* accept iff lo < v && v < hi   (descriptors.py:93-94) */
};

/* A GDML model as stored in the reference's .npz (gdml_model.py:60-125). */
typedef struct mbg_model_desc {
  int32_t n_atoms;               /* atoms per fragment = len(model["z"])                               */
  int32_t order;                 /* n-body order = len(comp_ids)                                        */
  int32_t n_train;               /* T = R_desc.shape[1]                                                 */
  int32_t n_perms;               /* P = perms.shape[0]                                                  */
  double sig;                    /* model["sig"]                                                        */
  double c;                      /* model["c"]   (integ_c)                                              */
  double std;                    /* model["std"] (1.0 when absent)                                      */
  const double *R_desc;          /* [D][T] row-major, D = n_atoms(n_atoms-1)/2  (model["R_desc"])      */
  const double *R_d_desc_alpha;  /* [T][D] row-major                         (model["R_d_desc_alpha"]) */
  const int64_t *tril_perms_lin; /* [P*D]                                    (model["tril_perms_lin"]) */
  const double *alphas_E;        /* [T] or NULL                              (model["alphas_E"])       */
  /* Criteria attached to the model by the caller (gdmlModel(model, criteria=...)). */
  int32_t criteria_kind;         /* MBG_CRIT_*                                                          */
  int32_t criteria_bound;        /* MBG_BOUND_*                                                         */
  double criteria_lo, criteria_hi;
  const int32_t *criteria_entity_ids; /* [n_atoms] desc_kwargs["entity_ids"] for com_distance_sum, else NULL */
  /* model["lattice"] (gdml_model.py:97-101), lattice vectors as COLUMNS, or NULL: when present the descriptor and
   * its Jacobian use pair differences clamped to this lattice (_gdml/desc.py:42-76 _pbc_diff).  Runs on the
   * MBG_CONTRACT_DMMA / AUTO kernel. */
  const double *lattice;         /* [9] row-major or NULL                                                  */
} mbg_model_desc;

/* The supersystem: which atoms form which entity, masses, optional periodic cell
 * (reference mbgdml/periodic.py:32-140).  Coordinates are passed per call. */
typedef struct mbg_system_desc {
  int32_t n_atoms;
  int32_t n_entities;
  const int32_t *entity_offsets; /* [n_entities+1] CSR offsets into entity_atoms                           */
  const int32_t *entity_atoms;   /* atom indices of each entity, ascending (np.where(entity_ids == e)[0])  */
  const double *masses;          /* [n_atoms] qcelemental.periodictable.to_mass(Z[i]) (descriptors.py:125) */
  const double *cell;            /* [9] row-major cell vectors (Cell.cell_v) or NULL for no periodicity    */
  double cell_cutoff;            /* Cell.cutoff (periodic.py:82); ignored when cell == NULL                */
  /* One cell per frame (overrides cell / cell_cutoff when non-NULL): NPT trajectories, and the 18 strained boxes
   * of the finite-difference virial (stress.py:75-152) evaluated as ONE 18-frame call.                     */
  const double *frame_cells;     /* [n_frames][9] or NULL                                                  */
  const double *frame_cutoffs;   /* [n_frames] Cell.cutoff of each frame's cell (required with frame_cells) */
  const int32_t *pbc;            /* [3] Cell.pbc (periodic.py:40-59): 0 = not periodic along that cell vector, or NULL =
                                    periodic in all three (find_mic(d, cell_v, pbc), periodic.py:78)            */
} mbg_system_desc;

/* One model applied to the system: the fragments it predicts and its alchemical scalers. */
typedef struct mbg_job {
  mbg_model_t model;
  /* Fragment source A (combs == NULL): utils.gen_combs(avail_entity_ids) (utils.py:449-489,
   * mbe.py:785-787): tuples of the Cartesian product of the per-slot entity sets that are
   * strictly ascending, in itertools.product order.  Enumerated on device. */
  const int32_t *set_offsets;  /* [order+1] */
  const int32_t *set_entities; /* concatenated per-slot entity ids */
  /* Fragment source B (combs != NULL): explicit entity tuples, e.g. a ray chunk (mbe.py:823-828). */
  const int32_t *combs; /* [n_combs][order] */
  int64_t n_combs;
  /* mbeAlchemyScale with linear_switching, already filtered to this model's order
   * (alchemy.py:30-62, switching.py:30-48, mbe.py:789-796). */
  int32_t n_alchemy;
  const int32_t *alchemy_entity; /* [n_alchemy] */
  const double *alchemy_factor;  /* [n_alchemy] */
} mbg_job;

#define MBG_FLAG_R_ON_DEVICE (1u << 0)   /* R is a device pointer                                       */
#define MBG_FLAG_OUT_ON_DEVICE (1u << 1) /* E, F, virial are device pointers (left enqueued on stream)  */
#define MBG_FLAG_VIRIAL (1u << 2)        /* accumulate sum_k r_k (x) f_k per fragment (stress.py:34-56) */
#define MBG_FLAG_ACCUMULATE (1u << 3)    /* add into E/F/virial instead of overwriting                  */

/* Counters returned per job. */
typedef struct mbg_job_stats {
  int64_t n_candidates; /* tuples examined (product space; this shard only) */
  int64_t n_enumerated; /* strictly ascending tuples = what gen_combs yields (this shard only) */
  int64_t n_accepted;   /* fragments that passed MIC cutoff and criteria and were evaluated    */
} mbg_job_stats;

const char *mbg_last_error(void);
int mbg_abi_version(void);
int mbg_device_count(int *count);

int mbg_context_create(int device, mbg_context_t *ctx);
int mbg_context_destroy(mbg_context_t ctx);
/* Use an existing CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); NULL = context's own. */
int mbg_context_set_stream(mbg_context_t ctx, void *cuda_stream);
int mbg_context_synchronize(mbg_context_t ctx);
/* Choose the kernel that evaluates _predict_gdml_wkr (gdml_predict.py:94-142): MBG_CONTRACT_*.  Both give the
 * same results to rounding; the choice exists so that the benchmark can time one against the other. */
int mbg_context_set_contraction(mbg_context_t ctx, int kind);
/* Choose how mbg_predict / mbg_accept_mask generate the candidates of a job given as entity sets (2- and 3-body jobs;
 * 4- to 6-body jobs over one homogeneous set), i.e. the tuples of utils.gen_combs, utils.py:449-489, that the
 * reference feeds to the slice -> r_mic -> criteria loop, gdml_predict.py:210-235: MBG_CULL_*.  The accepted fragments and their order do not depend on the choice.  In a
 * plain call a job of 2^23 candidates or more reads one count back from the device (it sizes the launches that
 * follow); smaller jobs and prepared steps (mbg_step_*, which take the choice in force when they are created) keep the
 * count on the device. */
int mbg_context_set_culling(mbg_context_t ctx, int kind);
/* Launch counters since context creation: number of kernels this library launched. */
int mbg_context_launch_count(mbg_context_t ctx, int64_t *n_launches);
/* Device time (ms, CUDA events on the context's stream) of the Matern contraction launches and of
 * all other launches during the most recent mbg_predict / mbg_predict_decomp call. */
int mbg_context_last_timing(mbg_context_t ctx, double *ms_contract, double *ms_other, int64_t *n_contract_launches);
/* Per-launch records of the Matern contraction kernel in the most recent call: fragment size,
 * fragments processed, padded training rows, permutations and device time (ms).  Arrays hold
 * max_records entries; *n_records receives the number of launches (may exceed max_records). */
int mbg_context_last_launches(mbg_context_t ctx, int32_t max_records, int32_t *n_records, int32_t *n_frag_atoms,
                              int64_t *n_fragments, int32_t *n_train_padded, int32_t *n_perms, double *ms);

/* gdmlModel(model, criteria=...)._unpack (gdml_model.py:90-125): upload + permutation tables. */
int mbg_model_create(mbg_context_t ctx, const mbg_model_desc *desc, mbg_model_t *model);
int mbg_model_destroy(mbg_model_t model);

/* mbePredict.predict (mbe.py:1009-1104) for all frames and all jobs in one call:
 *   for frame, for job: predict_gdml(Z, R[frame], entity_ids, combs, model, periodic_cell, ...)
 *   (gdml_predict.py:170-272) accumulated into E[n_frames], F[n_frames][n_atoms][3] and,
 *   with MBG_FLAG_VIRIAL, virial[n_frames][9] (un-normalised; the caller divides by the volume,
 *   mbe.py:1079).  Fragments are partitioned block-cyclically over `shard_count` callers
 *   (one per GPU); each returns its partial sums, to be all-reduced by the caller.
 *   stats may be NULL, else [n_jobs]. */
int mbg_predict(mbg_context_t ctx, const mbg_system_desc *sys, const double *R, int32_t n_frames,
                const mbg_job *jobs, int32_t n_jobs, uint32_t flags, int32_t shard_rank,
                int32_t shard_count, double *E, double *F, double *virial, mbg_job_stats *stats);

/* predict_gdml_decomp (gdml_predict.py:275-364) for one frame and one job with explicit
 * combs: E[n_combs], F[n_combs][model n_atoms][3]; rows of rejected fragments are NaN. */
int mbg_predict_decomp(mbg_context_t ctx, const mbg_system_desc *sys, const double *R,
                       const mbg_job *job, uint32_t flags, double *E, double *F, mbg_job_stats *stats);

/* ---- prepared steps: the MD caller (mbeCalculator.calculate, interfaces/ase.py:65-108, one mbePredict.predict per
 * integrator step) evaluates the SAME system, models and frame count over and over.  mbg_step_create uploads the
 * tables once and records the whole evaluation -- H2D of the coordinates (and cells), enumeration, culling,
 * contraction, scatter, D2H of the results -- into a CUDA graph; mbg_step_run replays it: one graph launch and one
 * synchronisation per step.  Flags: MBG_FLAG_VIRIAL.  Needs MBG_CONTRACT_DMMA / AUTO (the persistent kernel reads the
 * accepted-fragment counts on the device, so the launch sequence does not depend on them).
 *   cells / cutoffs: NULL = the cell(s) the step was created with, else [n_cells][9] / [n_cells] with n_cells = 1
 *   (sys->cell) or n_frames (sys->frame_cells) -- NPT steps and the strained boxes of virial_finite_diff
 *   (stress.py:75-152) reuse one step. */
typedef struct mbg_step_s *mbg_step_t;
int mbg_step_create(mbg_context_t ctx, const mbg_system_desc *sys, int32_t n_frames, const mbg_job *jobs,
                    int32_t n_jobs, uint32_t flags, int32_t shard_rank, int32_t shard_count, mbg_step_t *step);
int mbg_step_run(mbg_step_t step, const double *R, const double *cells, const double *cutoffs, double *E, double *F,
                 double *virial, mbg_job_stats *stats);
/* The same with R, E, F, virial resident on the device: enqueued on the context's stream, no synchronisation (the
 * caller orders its own work on that stream; multi-GPU callers all-reduce the partial sums of their shard). */
int mbg_step_run_device(mbg_step_t step, const double *dR, const double *cells, const double *cutoffs, double *dE,
                        double *dF, double *dvirial);
/* Counters of the most recent replay (after the stream has been synchronised). */
int mbg_step_stats(mbg_step_t step, mbg_job_stats *stats);
int mbg_step_destroy(mbg_step_t step);

/* The accept decision predict_gdml takes for every candidate fragment before it evaluates the model
 * (gdml_predict.py:214-235: slice -> Cell.r_mic, periodic.py:61-107 -> Criteria.accept, descriptors.py:65-103),
 * for all frames of R[n_frames][n_atoms][3] and one job: mask[n_frames][candidates per frame], candidates in the
 * job's own order (product order of the sets, or the explicit tuples); 0 = tuple not strictly ascending (gen_combs
 * never yields it), 1 = accepted (the fragment is evaluated), 2 = yielded but rejected by the cell cut-off or the
 * criteria.  *n_per_frame receives the candidates per frame; call with mask == NULL to size the buffer.
 * Host buffers (R may be a device pointer with MBG_FLAG_R_ON_DEVICE). */
int mbg_accept_mask(mbg_context_t ctx, const mbg_system_desc *sys, const double *R, int32_t n_frames,
                    const mbg_job *job, uint32_t flags, uint8_t *mask, int64_t *n_per_frame);

/* utils.gen_combs (utils.py:449-489) on device.  Call with combs == NULL to get the count,
 * then with a buffer of [*n_combs][order] int32. */
int mbg_enumerate(mbg_context_t ctx, int32_t order, const int32_t *set_offsets,
                  const int32_t *set_entities, int32_t *combs, int64_t *n_combs);

/* _from_r (_gdml/desc.py:203-233) on device for n_R geometries R[n_R][n_atoms][3]: descriptor x[n_R][D] (1 / pair
 * distance, np.tril_indices(n, -1) order) and its compressed Jacobian J[n_R][D][3] ((r_i - r_j) / |r_i - r_j|^3), with the
 * pair differences clamped to `lattice` ([9], vectors as columns; NULL: none) as _pbc_diff does.  This is what
 * gdmlModel.desc_func evaluates for a predictor that consumes the model object itself. */
int mbg_from_r(mbg_context_t ctx, int32_t n_atoms, const double *lattice, const double *R, int64_t n_R, double *x,
               double *J);

/* Criteria.desc on device for a batch of same-size structures (descriptors.py:131-201):
 * R[n_R][n_atoms][3] -> values[n_R].  entity_ids is used by MBG_CRIT_COM_DISTANCE_SUM only. */
int mbg_criteria_eval(mbg_context_t ctx, int32_t kind, int32_t n_atoms, const double *masses,
                      const int32_t *entity_ids, const double *R, int64_t n_R, double *values);

/* Cell.r_mic (periodic.py:86-107) on device for one fragment: r[n][3] -> out[n][3];
 * *accepted = 0 when any pair distance >= cutoff (the reference returns None). */
int mbg_r_mic(mbg_context_t ctx, const double *cell, double cutoff, const double *r, int32_t n,
              double *out, int32_t *accepted, const int32_t *pbc /* [3] or NULL = all periodic */);

/* Cell.d_mic (periodic.py:61-84) on device: minimum image of n displacement vectors d[n][3] -> out[n][3]; with
 * check_cutoff, *accepted = 0 when any pair distance among the imaged vectors is >= cutoff (the reference returns
 * None); without it *accepted = 1. */
int mbg_d_mic(mbg_context_t ctx, const double *cell, double cutoff, const double *d, int32_t n, int32_t check_cutoff,
              double *out, int32_t *accepted, const int32_t *pbc /* [3] or NULL = all periodic */);

/* mbe_worker + the add/remove step of mbe_contrib (mbe.py:73-197, 381-427), used by decomp_to_total
 * (mbe.py:432-508): for each selected reference structure i (row ref_rows[i] of E[n_total] /
 * Deriv[n_total][n_atoms_ref][3], updated in place) the pairs [pair_offsets[i], pair_offsets[i+1]) name the
 * lower-order structures found for its entity combinations, in itertools.combinations order;
 * slot_entity[p][k] is the entity id of the reference structure that slot k of lower structure lower_idx[p]
 * is added onto.  atom_entity / atom_rank describe the reference atoms (entity id, index within the entity),
 * lower_off / lower_atoms the atoms of each slot of a lower structure.  Sums run in pair order, so the result
 * is bit-identical to the reference's sequential loop.  sign = +1 ("add") or -1 ("remove"). */
int mbg_mbe_accumulate(mbg_context_t ctx, int64_t n_ref, const int64_t *ref_rows, const int64_t *pair_offsets,
                       const int64_t *lower_idx, const int32_t *slot_entity, int32_t n_slots,
                       int32_t n_atoms_ref, const int32_t *atom_entity, const int32_t *atom_rank,
                       int32_t n_atoms_lower, const int32_t *lower_off, const int32_t *lower_atoms,
                       const double *E_lower, const double *Deriv_lower, int64_t n_lower, double sign,
                       double *E, double *Deriv, int64_t n_total);

/* ---- training-side reuse of the distance/exp engine ------------------------------------- */

/* GDMLTrain._assemble_kernel_mat with col_idxs = np.s_[:] (reference mbgdml/_gdml/train.py:1105-1337; worker
 * _assemble_kernel_mat_wkr, train.py:92-292): the Hessian-of-Matern force-field kernel matrix.
 *   R_desc[n_train][D], R_d_desc[n_train][D][3] (compressed Jacobian, desc.py:448-480), D = n_atoms(n_atoms-1)/2,
 *   tril_perms_lin[n_perms * D]  ->  K[rows][rows] row-major, rows = n_train*3*n_atoms (+ n_train with
 *   use_E_cstr).  One CTA per ordered pair of training points.  Flags: MBG_FLAG_OUT_ON_DEVICE (K is a device
 *   pointer; the call returns with the work enqueued). */
int mbg_kernel_mat(mbg_context_t ctx, int32_t n_atoms, int32_t n_train, int32_t n_perms, const double *R_desc,
                   const double *R_d_desc, const int64_t *tril_perms_lin, double sig, int32_t use_E_cstr,
                   uint32_t flags, double *K);

/* GDMLTrain._assemble_kernel_mat with a column subset (col_idxs, train.py:1105-1115, 1187-1245; what the iterative
 * solver's preconditioner asks for): the column BLOCKS col_blocks[n_col_blocks] (training points j, any order, no
 * energy constraints) against all training points i  ->  K[n_train*3*n_atoms][n_col_blocks*3*n_atoms] row-major,
 * block (i, jj) = the (i, col_blocks[jj]) block of the full matrix.  The Python mirror gathers individual columns from
 * it on the device.  Flags: MBG_FLAG_OUT_ON_DEVICE. */
int mbg_kernel_mat_cols(mbg_context_t ctx, int32_t n_atoms, int32_t n_train, int32_t n_perms, const double *R_desc,
                        const double *R_d_desc, const int64_t *tril_perms_lin, double sig, const int32_t *col_blocks,
                        int32_t n_col_blocks, uint32_t flags, double *K);

/* analysis.models.gdml_mat52 (reference mbgdml/analysis/models.py:84-139): Matern-5/2 covariances between
 * n_R structures R[n_R][model n_atoms][3] and the T*P permuted training descriptors of `model`, in the row
 * order of gdmlModel.R_desc_perms (row t*P + p, gdml_model.py:109-113): out[n_R][T*P].
 * Flags: MBG_FLAG_R_ON_DEVICE, MBG_FLAG_OUT_ON_DEVICE. */
int mbg_mat52(mbg_context_t ctx, mbg_model_t model, const double *R, int64_t n_R, uint32_t flags, double *out);

/* analysis.models.gdml_mat52_wrk (analysis/models.py:31-81) for n_q descriptors r_desc[n_q][dim_d] against
 * explicit rows[n_rows][dim_d] (an expanded R_desc_perms): out[n_q][n_rows].  Host buffers. */
int mbg_mat52_rows(mbg_context_t ctx, int32_t dim_d, const double *r_desc, int64_t n_q, double sig,
                   const double *rows, int64_t n_rows, double *out);

/* Measured FP64 FMA-pipe peak of this GPU (dependent-chain-free DFMA loop, CUDA-event timed),
 * in TFLOP/s: the roofline denominator bench.py reports against. */
int mbg_measure_fp64_peak(mbg_context_t ctx, double *tflops, double *sm_clock_mhz_est);
/* Measured peak of the FP64 tensor pipe (back-to-back mma.sync.m8n8k4.f64 on independent accumulators),
 * in TFLOP/s: the roofline denominator for the MBG_CONTRACT_DMMA kernel. */
int mbg_measure_fp64_tensor_peak(mbg_context_t ctx, double *tflops);

/* Measured rate of exp(-c sqrt(x)) evaluated with the CUDA math library (4 independent chains per thread), in
 * evaluations per second: what the transcendental-bound 1-body workload (BASELINE config 2) is reported against. */
int mbg_measure_exp_sqrt_rate(mbg_context_t ctx, double *evals_per_s);

#ifdef __cplusplus
}
#endif
#endif /* MBGDML_B200_H */
