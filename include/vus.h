/* vus.h -- C-ABI of the B200-native batch factor-graph optimizer (libvus.so).
 *
 * Drop-in boundary for the one hot path of hvak/visual-underwater-slam:
 *     gtsam.LevenbergMarquardtOptimizer(graph, initial, gtsam.LevenbergMarquardtParams()).optimize()
 *                                                                   (/root/reference/batch.py:337)
 * over the graph that /root/reference/batch.py:270-305 builds.  The reference has no FFI of its own
 * (it calls gtsam through gtsam's pybind11 wrapper); these entry points are what a binding for this
 * path would bind.  Each one cites the reference / gtsam interface it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; every call returns 0 on success or a negative vus_status;
 *     vus_last_error() returns a message owned by the handle.  No exceptions cross the ABI.
 *   - all numeric tables are FP64 structure-of-arrays, COMPONENT-MAJOR:  table[c * n + i].
 *   - `mem` says where a numeric table lives and how it is laid out: VUS_MEM_HOST (the library copies
 *     host->device) or VUS_MEM_DEVICE (a device pointer, e.g. torch.Tensor.data_ptr(); copied device->device),
 *     both component-major; the *_ROWS variants take the natural row-major [n][dim] array (what numpy / gtsam
 *     callers hold) and transpose it on the device, so the host never re-packs a table.
 *     Key / index / insertion-order arrays are always host memory.
 *   - one host thread per handle (several handles may run concurrently from several threads); all kernels run on
 *     the stream passed in, or on a stream the handle owns when NULL is passed.  The library captures its fixed
 *     launch sequences into CUDA graphs on that stream: do not share one stream between handles that run at the
 *     same time, and do not pass the legacy default stream expecting capture (NULL selects the handle's own).
 */
#ifndef VUS_H_
#define VUS_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vus_handle vus_handle;

enum vus_status { VUS_OK = 0, VUS_ERR_INVALID = -1, VUS_ERR_UNSUPPORTED = -2, VUS_ERR_CUDA = -3, VUS_ERR_STATE = -4 };
enum vus_mem { VUS_MEM_HOST = 0, VUS_MEM_DEVICE = 1,            /* component-major table[c * n + i] */
               VUS_MEM_HOST_ROWS = 2, VUS_MEM_DEVICE_ROWS = 3 }; /* row-major table[i * dim + c]: transposed on the device */

/* Variable kinds (gtsam::Values entries, batch.py:274, :283-288, :297-298).
 *   POSE  12 comps: R row-major (9) then t (3)      X(i)   gtsam::Pose3
 *   VEL    3 comps                                   V(i)   Vector3
 *   BIAS   6 comps: accelerometer (3), gyro (3)      B(0)   imuBias::ConstantBias
 *   LM     3 comps                                   L(id)  Point3                                  */
enum vus_var_kind { VUS_VAR_POSE = 0, VUS_VAR_VEL = 1, VUS_VAR_BIAS = 2, VUS_VAR_LM = 3 };

/* Factor types.  idx = int32 variable indices [slots][n] into the per-kind tables (ascending-key order).
 *   type          replaces (reference)                         slots (kind)            meas comps            sqrt_info comps
 *   PRIOR_POSE    gtsam.PriorFactorPose3   batch.py:281        x                       12 (Pose3)            6  (1/sigma)
 *   PRIOR_VEL     gtsam.PriorFactorVector  batch.py:282        v                       3                     3
 *   BETWEEN       gtsam.BetweenFactorPose3 (north_star)        x1, x2                  12 (Pose3)            6
 *   DVL           gtsam.CustomFactor + velocity_error          v, x   (batch.py:247)   3 (body velocity)     3
 *                 batch.py:196-233, :241-250
 *   STEREO        gtsam.GenericStereoFactor3D batch.py:300-305 x, l                    3 (uL, uR, v)         3
 *   IMU           gtsam.ImuFactor          batch.py:238        xi, vi, xj, vj, b       67 (packed PIM)       45 (upper-tri R, R^T R = Sigma^-1)
 * Packed PIM: dR 9 | dP 3 | dV 3 | dt 1 | bias_hat 6 | dR/dbg 9 | dP/dba 9 | dP/dbg 9 | dV/dba 9 | dV/dbg 9   (manifold build)
 *             theta 3, 6 unused | p 3 | v 3 | dt 1 | bias_hat 6 | dtheta/dbg 9 | dp/dba 9 | dp/dbg 9 | dv/dba 9 | dv/dbg 9   (tangent build,
 *             vus_set_gtsam_build; every IMU factor of a handle is of the handle's build).
 * Whitened Jacobian layouts returned by vus_linearize (node order, element (row,col) at J[(row*cols+col)*n+f]):
 *   PRIOR_POSE 6x6 | PRIOR_VEL 3x3 | BETWEEN 6x12 [H1|H2] | DVL 3x9 [Hx|Hv] | STEREO 3x9 [Hpose|Hlm]
 *   IMU 9x24 [Hxi Hvi | Hxj Hvj | Hbias]                                                              */
enum vus_factor_type { VUS_FACTOR_PRIOR_POSE = 0, VUS_FACTOR_PRIOR_VEL = 1, VUS_FACTOR_BETWEEN = 2,
                       VUS_FACTOR_DVL = 3, VUS_FACTOR_STEREO = 4, VUS_FACTOR_IMU = 5 };

/* gtsam::LevenbergMarquardtParams (defaults = gtsam's, batch.py:337 passes a default-constructed object)
 * plus the knobs of the iterative linear solver that replaces gtsam's multifrontal Cholesky.          */
typedef struct vus_lm_params {
  int32_t max_iterations;        /* 100   */
  double relative_error_tol;     /* 1e-5  */
  double absolute_error_tol;     /* 1e-5  */
  double error_tol;              /* 0     */
  double lambda_initial;         /* 1e-5  */
  double lambda_factor;          /* 10    */
  double lambda_upper_bound;     /* 1e5   */
  double lambda_lower_bound;     /* 0     */
  double min_model_fidelity;     /* 1e-3  */
  int32_t pcg_max_iterations;    /* 500   */
  double pcg_rel_tol;            /* 1e-12: ||r|| <= tol * ||rhs|| (stops earlier if the residual stagnates) */
  int32_t max_supernode;         /* 0 = auto (band width from the graph, capped so k*D <= 96) */
  int32_t verbose;
  int32_t profile_kernels;       /* 1: time every kernel launch with CUDA events, per kernel class (ms_class) */
  int32_t band_chunks;           /* band factorization of stereo-scale supernodes: 0 = auto (block Cholesky inside one chunk per SM,
                                    cyclic reduction across the separators), n > 0 = n chunks, < 0 = cyclic reduction of the whole chain */
} vus_lm_params;

typedef struct vus_lm_result {
  int32_t iterations;            /* accepted LM steps (gtsam's iterations()) */
  int32_t inner_iterations;      /* lambda tries */
  int32_t linearizations;
  int32_t pcg_iterations;        /* total */
  int32_t solve_failures;        /* lambda tries whose damped system could not be solved (non-positive pivot, or the PCG below) */
  int32_t pcg_not_converged;     /* of those: PCG ended on a NaN or stalled with a true relative residual above 1e-6 */
  double initial_error, final_error, final_lambda;
  double ms_total, ms_linearize, ms_assemble, ms_schur, ms_factor, ms_pcg, ms_update;
  int64_t kernel_launches;
  int64_t factors_linearized;    /* sum over linearizations of #factors */
  /* device time per kernel class (CUDA events on the launching stream; only with profile_kernels=1):
   * 0 linearize | 1 error | 2 linearised error | 3 assemble | 4 stereo assemble | 5 damp+Schur | 6 BCR factor |
   * 7 BCR solve | 8 matvec | 9 border | 10 vector/reduce | 11 retract */
  double ms_class[16];
  int64_t launches_class[16];
  double worst_pcg_rel_residual; /* largest TRUE relative residual ||rhs - A x|| / ||rhs|| any damped solve ended with */
} vus_lm_result;

void vus_default_lm_params(vus_lm_params* p);

/* lifetime */
int vus_create(int device, vus_handle** out);
void vus_destroy(vus_handle* h);
const char* vus_last_error(const vus_handle* h);

/* gtsam::Values  (batch.py:81, :274, :283-288, :297-298): keys ascending uint64 Symbol keys. */
int vus_set_variables(vus_handle* h, int kind, int64_t n, const uint64_t* keys, const double* data, int mem);
int vus_get_variables(vus_handle* h, int kind, double* out, int mem);
/* keep / restore a device-side copy of the current values (re-running optimize() from the same initial estimate) */
int vus_save_values(vus_handle* h);
int vus_restore_values(vus_handle* h);

/* gtsam::NonlinearFactorGraph::add / push_back (batch.py:281-282, :291-292, :305), one call per type.
 * orig_index = NonlinearFactorGraph insertion index of each factor (kept for bit-exact factor indexing). */
int vus_add_factors(vus_handle* h, int type, int64_t n, const int32_t* var_idx, const double* meas,
                    const double* sqrt_info, const int64_t* orig_index, int mem);
int vus_set_calibration(vus_handle* h, const double K[6]);       /* Cal3_S2Stereo fx fy s u0 v0 b (batch.py:115) */
int vus_set_gravity(vus_handle* h, const double g[3]);           /* PreintegrationParams n_gravity (batch.py:181) */
int vus_set_lm_params(vus_handle* h, const vus_lm_params* p);
/* The two CMake switches of gtsam that change this path's arithmetic (the reference pins no gtsam version, README.md:18):
 *   tangent_preintegration          GTSAM_TANGENT_PREINTEGRATION (ON in the gtsam 4.0 - 4.2 wheels): ImuFactor::evaluateError
 *                                   corrects the preintegrated 9-vector linearly for the bias and retracts it
 *                                   (TangentPreintegration.cpp); 0 = ManifoldPreintegration.cpp
 *   slow_but_correct_betweenfactor  GTSAM_SLOW_BUT_CORRECT_BETWEENFACTOR (OFF in the wheels): BetweenFactor::evaluateError
 *                                   multiplies its Jacobians by the derivative of Local(); 0 = H1 = -Ad(hx^-1), H2 = I
 * Defaults: 1, 0 (the pip wheel build).  Also selects the variant vus_preintegrate_imu computes. */
int vus_set_gtsam_build(vus_handle* h, int tangent_preintegration, int slow_but_correct_betweenfactor);

/* ---- one graph split across GPUs by contiguous pose range (BASELINE.json config 5) -------------------
 * Each rank holds a LOCAL graph: its owned poses first, then the halo poses (owned by other ranks) that its factors
 * touch; every factor that touches an owned pose is present (cut factors are duplicated on both sides, so each rank
 * assembles complete Hessian rows for its owned poses and nothing has to be reduced).  Per factor type the factors this
 * rank owns come first (n_owned_factors): only those enter the error sums.  Pose graphs only (PRIOR_POSE / BETWEEN).
 * The collectives are supplied by the caller -- torch.distributed over NCCL in the shipped host code, gloo in the CPU
 * tests -- so the library does not link a communication stack:
 *   VUS_COMM_ALLREDUCE_SUM  buf = `count` doubles in device memory, summed over ranks in place
 *   VUS_COMM_HALO           buf = a node vector in device memory, `count` (= D) doubles per local node: fill the halo
 *                           nodes' entries with their owners' entries
 * The callback is invoked after the library synchronised its stream and must return (0 = ok) once the result is visible
 * to the device.  Every rank takes identical accept / reject decisions from the all-reduced scalars.              */
enum vus_comm_op { VUS_COMM_ALLREDUCE_SUM = 0, VUS_COMM_HALO = 1 };
typedef int (*vus_comm_fn)(void* ctx, int op, void* buf, int64_t count);
int vus_set_partition(vus_handle* h, int64_t n_owned_nodes, const int64_t n_owned_factors[6]);
int vus_set_comm(vus_handle* h, vus_comm_fn fn, void* ctx);
/* Where this rank sits in the chain of pose ranges (rank r owns the poses right after rank r - 1's): prev_local[j] / next_local[j],
 * j < nside, are the LOCAL indices of the halo poses that are the (j + 1)-th pose before its first / after its last owned pose
 * globally (-1: not in the halo or beyond the ends of the chain); nside >= the supernode width the analysis will find (16 always
 * is).  With it the per-rank band factorizations are tied together exactly (a partitioned solve over the rank interfaces: one
 * extra all-reduce of 12 k nranks doubles per preconditioner application, csrc/spike.cuh), so the partitioned PCG takes the
 * iterations of the one-rank solve; without it -- or when a rank's owned range is not a whole number of supernodes -- the band
 * preconditioner is block-Jacobi across ranks.  Call before vus_analyze, on every rank.  Replaces nothing in the reference:
 * gtsam runs batch.py:337 in one process. */
int vus_set_partition_chain(vus_handle* h, int32_t nside, const int64_t* prev_local, const int64_t* next_local, int32_t rank, int32_t nranks);
/* stream_ordered = 1: the callback enqueues its collective ON THE STREAM passed to vus_optimize (NCCL through torch.distributed
 * with that stream current) and returns without waiting; the library then neither synchronises before the call nor expects
 * the result on the host -- the stream orders everything.  0 (default): host-synchronous callbacks (gloo, the CPU tests). */
int vus_set_comm_mode(vus_handle* h, int stream_ordered);
/* Collectives issued by the library itself (vus_comm_init(ncclUniqueId, rank, nranks)): NCCL over NVLink on the
 * stream the solve runs on -- ncclAllReduce for the PCG / LM scalars, grouped ncclSend / ncclRecv for the halo nodes -- so the
 * collectives sit INSIDE the captured launch sequences (one PCG iteration = one CUDA graph) and no host code runs per
 * collective.  libnccl.so.2 is loaded at run time (no link-time dependency).  Rank 0 obtains an id with vus_nccl_unique_id
 * and hands it to every rank by any out-of-band means (128 bytes); vus_set_halo gives the exchange lists:
 *   peers[q]                       rank exchanged with
 *   send_idx[send_ptr[q] .. send_ptr[q+1])   LOCAL indices of the owned nodes peer q holds as halo, in the order of its halo
 *   recv_off[q], recv_cnt[q]       where peer q's nodes sit in this rank's halo (halo nodes are sorted by owner: contiguous)
 * Overrides vus_set_comm for the handle. */
int vus_nccl_unique_id(void* out128);
int vus_comm_init(vus_handle* h, const void* unique_id128, int rank, int nranks);
int vus_set_halo(vus_handle* h, int32_t npeers, const int32_t* peers, const int64_t* send_ptr, const int32_t* send_idx,
                 const int64_t* recv_off, const int64_t* recv_cnt);

/* ---- a batch of INDEPENDENT trajectories in one handle (BASELINE.json config 4: 4096 x 500-pose graphs, a block per GPU)
 * The reference runs one gtsam.LevenbergMarquardtOptimizer per graph (batch.py:337); a 500-pose solve cannot fill a B200,
 * so the caller concatenates the trajectories of a shard into one graph -- component c owns the contiguous pose / velocity
 * index range [node_start[c], node_start[c+1]) and, if the graphs have one, the bias with index c -- and the library
 * runs gtsam's LM loop for every component at once: each component keeps its own lambda, error, accept / reject /
 * convergence decisions and PCG scalars, exactly as ncomp separate optimizers would.  No factor may connect two
 * components; stereo factors are not supported in this mode.  Off-band factors (loop closures) of a component are
 * inverted exactly by a capacitance (Woodbury) solve on top of the factored band when no component has more than 8.  Call before vus_analyze; vus_optimize then fills the
 * summed errors / the largest iteration count into vus_lm_result and vus_get_component_results the per-trajectory ones. */
typedef struct vus_component_result {
  int32_t iterations;            /* accepted LM steps of this trajectory */
  int32_t inner_iterations;      /* its lambda tries */
  double initial_error, final_error, final_lambda;
} vus_component_result;
int vus_set_components(vus_handle* h, int64_t ncomp, const int64_t* node_start /* [ncomp + 1] */);
int vus_get_component_results(vus_handle* h, vus_component_result* out /* [ncomp] */);

/* symbolic phase: node ordering, supernode band layout, off-band blocks, Schur destination lists */
int vus_analyze(vus_handle* h);
/* band description after analyze: D (node dof), k (nodes per supernode), Ns (supernodes), nrem (off-band blocks), stereo observations,
 * B = k D, L (reduced dofs), factors, chunks of the band factorization (0 = plain cyclic reduction), supernodes the band
 * preconditioner covers (the owned prefix of a partition), landmarks, reserved */
int vus_get_layout(vus_handle* h, int64_t out[12]);

/* LevenbergMarquardtOptimizer::optimize()  (batch.py:337). Asynchronous work is issued on `stream`
 * (a cudaStream_t, may be NULL); the call returns after the convergence decision. Values are updated in place. */
int vus_optimize(vus_handle* h, void* stream, vus_lm_result* result);

/* Try-by-try history of the last single-graph vus_optimize -- gtsam's LM with verbosityLM = TRYLAMBDA prints the same
 * sequence: the lambda tried, whether the damped system could be solved, the error at the trial point (inf when the step
 * was rejected before evaluation) and whether the step was accepted.  Writes min(count, capacity) entries. */
typedef struct vus_lm_try {
  double lambda, new_error;
  int32_t solved, accepted, pcg_iterations, reserved;
} vus_lm_try;
int vus_get_trace(vus_handle* h, int32_t capacity, vus_lm_try* out, int32_t* count);

/* NonlinearFactorGraph::error(values) = sum 1/2 ||whitened r||^2 at the current values */
int vus_error(vus_handle* h, void* stream, double* out);
/* per-factor 1/2 ||r||^2 in ORIGINAL insertion order (length = total factor count) */
int vus_factor_errors(vus_handle* h, void* stream, double* out);
/* NonlinearFactorGraph::linearize for one factor type at the current values: whitened r [m][n], J [m*cols][n] -> host */
int vus_linearize(vus_handle* h, void* stream, int type, double* r_out, double* J_out);
/* one damped Gauss-Newton step at the current values WITHOUT applying it:
 * solves (J^T J + lambda I) delta = -J^T r; delta_pose [nx][6], delta_vel [nv][3], delta_bias [nb][6], delta_lm [nl][3] (host, AoS) */
int vus_solve_step(vus_handle* h, void* stream, double lambda, double* d_pose, double* d_vel, double* d_bias, double* d_lm,
                   int32_t* pcg_iterations);
/* gtsam::Marginals(graph, values).marginalCovariance(key) / .jointMarginalCovariance(keys)  (what users
 * of batch.py:337's result ask for next).  Linearizes at the CURRENT values (no damping), and returns the joint
 * covariance of nq variables -- query q is variable idx[q] of kind kinds[q] (vus_var_kind) -- as one dense symmetric
 * matrix cov_out [M][M] (host, row-major), M = sum of the queries' tangent dimensions (POSE 6: rotation then
 * translation, gtsam's Pose3 tangent order; VEL 3; BIAS 6; LM 3), blocks in query order.  Each column is one exact
 * solve of the undamped normal equations with a unit right-hand side (landmarks by their Schur complement, camera
 * block by the band factorization + PCG).  Returns VUS_ERR_STATE when the undamped system is not positive definite
 * (gtsam throws IndeterminantLinearSystemException there). */
int vus_marginal_covariance(vus_handle* h, void* stream, int64_t nq, const int32_t* kinds, const int32_t* idx, double* cov_out);

/* ---- front-end loops that feed the graph; tables are row-major (mem = VUS_MEM_*_ROWS) --------
 * gtsam.PreintegratedImuMeasurements.integrateMeasurement x k + resetIntegration per keyframe interval
 * (batch.py:289-293; covariances of batch.py:183-185): acc, gyro [n][k][3], constant dt (batch.py:290 passes 0.005).
 * Emits what vus_add_factors(VUS_FACTOR_IMU) takes: pim_out [n][67] (packed PIM), sqrt_info_out [n][45] (upper R,
 * R^T R = preintMeasCov^-1), in the preintegration variant of vus_set_gtsam_build. */
int vus_preintegrate_imu(vus_handle* h, void* stream, int64_t n, int32_t k, const double* acc, const double* gyro, double dt,
                         const double bias_hat[6], const double acc_cov[9], const double gyro_cov[9], const double int_cov[9],
                         double* pim_out, double* sqrt_info_out, int mem);
/* landmark initial values from stereo measurements (get_landmarks, batch.py:144-176, in gtsam's StereoCamera
 * convention uL > uR): points_out[o] = X(pose_idx[o]) * backproject(uL, uR, v) with the calibration of
 * vus_set_calibration and the CURRENT pose values.  pose_idx is host memory; meas [n][3], points_out [n][3]. */
int vus_backproject_stereo(vus_handle* h, void* stream, int64_t n, const int32_t* pose_idx, const double* meas, double* points_out, int mem);

/* test hook for the band solver (kernel 3b): linearize + assemble + damp/Schur at the current values, copy out the
 * block-tridiagonal band (SD [Ns][B][B], SU [Ns-1][B][B], host, may be NULL), factor it by block cyclic reduction and
 * solve nrhs right-hand sides in place (x_inout [nrhs][Ns*B], host).  Returns 1 if a pivot was not positive. */
int vus_debug_band_solve(vus_handle* h, void* stream, double lambda, double* SD_out, double* SU_out, double* x_inout, int nrhs);
/* time `reps` launches of kernel 1 (linearize, all factor types) with CUDA events; ms per repetition */
int vus_time_linearize(vus_handle* h, void* stream, int reps, double* ms_per_rep);

#ifdef __cplusplus
}
#endif
#endif /* VUS_H_ */
