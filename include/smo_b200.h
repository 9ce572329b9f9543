/* smo_b200.h - C ABI of the B200-native SphereManOpt hot path (libsmo_b200.so).
 *
 * Every entry point replaces one piece of the reference's Python/Dedalus hot path (citations are into the
 * reference repository mannixp/SphereManOpt; SH = Example_Problems/Periodic_Domain(Fourier)/Swift_Hohenberg/
 * FWD_Solve_SH23.py, KD = .../Kinematic_Dynamo/FWD_Solve_KDyn.py, SGD = Sphere_Grad_Descent.py).
 *
 * Conventions
 *  - all functions return 0 on success or a negative error code; smo_last_error() gives the message of the
 *    calling thread's last failure (the Python wrappers raise RuntimeError with it);
 *  - pointers named *_dev are device pointers on the handle's device, *_host are host pointers; no library
 *    types appear in the signatures, `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *  - vectors use the reference's layout: the dealiased grid in C order, float64.  SH23: M = 2*Npts values per
 *    instance (SH:110-128).  Kinematic dynamo: concat(Fx.ravel(), Fy.ravel(), Fz.ravel()), 3*M^3 values with
 *    M = 3*Npts/2 (KD:137); with more than one rank every rank holds the z-slab [x][y][z0:z0+nz] of each
 *    component (3*M*M*nz values);
 *  - handles are not thread safe; use one handle per thread.  No global mutable state besides the CUDA context.
 */
#ifndef SMO_B200_H
#define SMO_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct smo_sh23 smo_sh23_t;
typedef struct smo_kdyn smo_kdyn_t;

int smo_version(void);
const char* smo_last_error(void);
/* number of kernels of this library launched by the calling process so far (bench.py's gpu_launches) */
long long smo_launch_count(void);

/* flags */
#define SMO_ADJOINT_CONTINUOUS 1 /* Adjoint_type="Continuous" (SH:654-656, KD:904-910) */
#define SMO_COST_INTEGRATED 2    /* Cost_function="Integrated" (KD:655-669, 738-742, 861-864): J = dt * sum_n <B^n,B^n> */

/* ------------------------------------------------------------------------------------------------------------
 * Swift-Hohenberg SH23 (1-D periodic Fourier, SBDF1, dealias 2)
 * ---------------------------------------------------------------------------------------------------------- */
/* Replaces FWD_Solve_Build_Lin (SH:279-332): Npts Fourier modes on [0,L), parameter a (= -0.3 at SH:309). */
int smo_sh23_create(smo_sh23_t** h, int Npts, double L, double a);
int smo_sh23_destroy(smo_sh23_t* h);
/* bytes of snapshot storage per instance for n_iters steps (GEN_BUFFER, SH:238-272).  The store is opaque: every forward state
 * is kept ON THE GRID ((n_iters+1)*M doubles: what the adjoint's pointwise product consumes, written straight from the
 * registers of the forward solve's c2r transform) plus the coefficients of the final state (Npts/2 complex);
 * smo_sh23_snapshot_coef converts a stored state back to the reference's coefficient form A_fwd[:, n]. */
size_t smo_sh23_snapshot_bytes(const smo_sh23_t* h, int n_iters);
/* coefficients [batch][Npts/2] complex128 of stored state n (0..n_iters) of every instance (inspection / tests / side outputs) */
int smo_sh23_snapshot_coef(smo_sh23_t* h, const void* snaps_dev, int batch, int n_iters, int n, void* coef_dev, void* stream);
/* Replaces FWD_Solve_IVP_Lin (SH:409-545) for `batch` independent instances.
 * X_dev [batch][M] in; snaps_dev [batch][smo_sh23_snapshot_bytes] out; J_dev [batch] out with
 * J = dt*sum_{n=0..n_iters} mean(u_n^2)  (the reference returns -J, SH:545). */
int smo_sh23_forward(smo_sh23_t* h, const double* X_dev, int batch, double dt, int n_iters, void* snaps_dev,
                     double* J_dev, void* stream);
/* Replaces Compatib_Cond + ADJ_Solve_IVP_Lin (SH:552-596, 598-729).  snaps_dev as written by smo_sh23_forward;
 * grad_dev [batch][M] out (dJ/du0 of the returned -J).  flags: SMO_ADJOINT_CONTINUOUS. */
int smo_sh23_adjoint(smo_sh23_t* h, int batch, double dt, int n_iters, const void* snaps_dev, double* grad_dev,
                     int flags, void* stream);
/* Replaces FWD_Solve_IVP_PREP (SH:334-407): n_iters+1 SBDF1 steps, final state on the grid -> out_dev [batch][M]. */
int smo_sh23_prep(smo_sh23_t* h, const double* X_dev, int batch, double dt, int n_iters, double* out_dev,
                  void* stream);
/* transform helpers (initial conditions, tests): grid [batch][M] <-> retained coefficients [batch][Npts/2] complex128,
 * Dedalus conventions ([D2-1..3]): coefficients are amplitudes, scale change = zero-pad / truncate */
int smo_sh23_to_coef(smo_sh23_t* h, const double* X_dev, int batch, void* coef_dev, void* stream);
int smo_sh23_to_grid(smo_sh23_t* h, const void* coef_dev, int batch, double* out_dev, void* stream);
/* Host-buffer forms of the calls above (copies inside, synchronous on return): the drop-in for the reference's
 * numpy-in / numpy-out callables.  The snapshot store stays on the device: snaps_dev as above, or NULL to use a
 * store owned by the handle (grown on demand). */
int smo_sh23_forward_host(smo_sh23_t* h, const double* X_host, int batch, double dt, int n_iters, void* snaps_dev,
                          double* J_host, void* stream);
int smo_sh23_adjoint_host(smo_sh23_t* h, int batch, double dt, int n_iters, const void* snaps_dev, double* grad_host,
                          int flags, void* stream);
int smo_sh23_prep_host(smo_sh23_t* h, const double* X_host, int batch, double dt, int n_iters, double* out_host,
                       void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Kinematic dynamo (3-D periodic Fourier, CNAB1, dealias 3/2), slab-decomposed over nranks GPUs
 * ---------------------------------------------------------------------------------------------------------- */
/* Replaces FWD_Solve_Build_Lin (KD:362-450) and the solver construction in ADJ_Solve_IVP_Lin (KD:807-886).
 * rank/nranks: coefficient space is split along kx (Npts/2 planes), grid space along z; nranks must divide
 * Npts/2 and 3*Npts/2.  nccl_comm: an initialised ncclComm_t (as void*) when nranks > 1, else NULL. */
int smo_kdyn_create(smo_kdyn_t** h, int Npts, double L, int rank, int nranks, void* nccl_comm);
int smo_kdyn_destroy(smo_kdyn_t* h);
/* local sizes: elements of one grid component (M*M*nz doubles), of one coefficient component (complex128),
 * and bytes of the snapshot store for n_iters steps (GEN_BUFFER, KD:319-355).  The store is opaque: forward states are
 * kept as x-spectra on the z-slab (the form the adjoint x pass consumes: (n_iters+1) * 3 * (Npts/2)*M*nz complex) plus
 * the coefficients of the final state; smo_kdyn_snapshot_coef converts a stored state back to coefficients. */
size_t smo_kdyn_grid_elems(const smo_kdyn_t* h);
size_t smo_kdyn_coef_elems(const smo_kdyn_t* h);
size_t smo_kdyn_snapshot_bytes(const smo_kdyn_t* h, int n_iters);
size_t smo_kdyn_segment_bytes(const smo_kdyn_t* h, int every);
/* coefficients [3][coef_elems] of stored state n (0..n_iters) of a filled snapshot store (inspection / tests) */
int smo_kdyn_snapshot_coef(smo_kdyn_t* h, const void* snaps_dev, int n_iters, int n, void* coef_dev, void* stream);
/* Replaces FWD_Solve_IVP_Lin (KD:529-689), Cost_function="Final".  B0_dev, U_dev: [3][grid_elems] local slabs.
 * snaps_dev out.  J_host out: this rank's share of mean_grid(|B_N|^2) (sum over ranks = J; reference returns -J). */
int smo_kdyn_forward(smo_kdyn_t* h, const double* B0_dev, const double* U_dev, double Rm, double dt, int n_iters,
                     void* snaps_dev, double* J_host, int flags, void* stream);
/* Replaces Compatib_Cond + ADJ_Solve_IVP_Lin (KD:696-764, 766-1004).  Uses the velocity field cached by the
 * preceding smo_kdyn_forward on the same handle (the reference's f -> Grad_f state coupling).
 * gradB_dev, gradU_dev: [3][grid_elems] out. */
int smo_kdyn_adjoint(smo_kdyn_t* h, double Rm, double dt, int n_iters, const void* snaps_dev, double* gradB_dev,
                     double* gradU_dev, int flags, void* stream);
/* Checkpointed forms of the two calls above (two-level, revolve style; for grids whose n_iters+1 states exceed HBM, e.g.
 * 256^3 x 1000 steps = 401 GB).  The forward solve keeps the states 0, every, 2*every, ... and N in ckpt_dev
 * (smo_kdyn_checkpoint_bytes, coefficient form); the adjoint sweep recomputes one segment at a time into seg_dev
 * (smo_kdyn_segment_bytes(h, every) bytes, x-spectral form): at most n_iters extra forward steps.  Same results bit for bit. */
size_t smo_kdyn_checkpoint_bytes(const smo_kdyn_t* h, int n_iters, int every);
int smo_kdyn_forward_ckpt(smo_kdyn_t* h, const double* B0_dev, const double* U_dev, double Rm, double dt, int n_iters,
                          int every, void* ckpt_dev, double* J_host, int flags, void* stream);
int smo_kdyn_adjoint_ckpt(smo_kdyn_t* h, double Rm, double dt, int n_iters, int every, const void* ckpt_dev, void* seg_dev,
                          double* gradB_dev, double* gradU_dev, int flags, void* stream);
/* Replaces FWD_Solve_IVP_Prep (KD:452-527): n_iters+1 CNAB1 steps from B0 with velocity U; the final field on
 * the grid -> out_dev [3][grid_elems]. */
int smo_kdyn_prep(smo_kdyn_t* h, const double* B0_dev, const double* U_dev, double Rm, double dt, int n_iters,
                  double* out_dev, void* stream);
/* Host-buffer forms: full-size reference vectors (3*M^3 doubles) on the host; each rank reads/writes its z-slab
 * (the other entries of the gradient outputs are left untouched).  snaps_dev as above, or NULL to use a store owned
 * by the handle.  Synchronous on return. */
int smo_kdyn_forward_host(smo_kdyn_t* h, const double* B0_host, const double* U_host, double Rm, double dt,
                          int n_iters, void* snaps_dev, double* J_host, int flags, void* stream);
int smo_kdyn_adjoint_host(smo_kdyn_t* h, double Rm, double dt, int n_iters, const void* snaps_dev, double* gradB_host,
                          double* gradU_host, int flags, void* stream);
int smo_kdyn_prep_host(smo_kdyn_t* h, const double* B0_host, const double* U_host, double Rm, double dt, int n_iters,
                       double* out_host, void* stream);
/* this rank's z-slab [3][M][M][nz] on the device <-> the full reference vector (3*M^3 doubles) on the host: one strided 2-D copy
 * (Vec_to_Field's local slicing, KD:156-169).  Give exactly one of host_in (H2D) / host_out (D2H; only this rank's entries are
 * written).  Asynchronous on `stream`. */
int smo_kdyn_slab_copy(smo_kdyn_t* h, double* slab_dev, const double* host_in, double* host_out, void* stream);
/* transform helpers (tests, initial conditions): grid [3][grid_elems] <-> coefficients [3][coef_elems] */
int smo_kdyn_to_coef(smo_kdyn_t* h, const double* grid_dev, void* coef_dev, void* stream);
int smo_kdyn_to_grid(smo_kdyn_t* h, const void* coef_dev, double* grid_dev, void* stream);
/* per-kernel timing of the time loops: which = 0 off, 1 = z passes outside the fused step, 2 = y passes, 3 = fused forward x pass,
 * 5 = all-to-all / barrier launches, 6 = fused adjoint x pass, 7 = fused z step.
 * smo_kdyn_profile_read returns the accumulated CUDA-event time (ms) and launch count since the last set. */
int smo_kdyn_profile_set(smo_kdyn_t* h, int which);
int smo_kdyn_profile_read(smo_kdyn_t* h, double* total_ms, long long* launches);
/* Peer-memory transposes (replaces the all-to-all by stores over NVLink fused into the FFT passes).  Every rank
 * calls smo_kdyn_peer_export (writes smo_kdyn_peer_handle_bytes() bytes of CUDA IPC handles), the host program
 * all-gathers the blobs in rank order and every rank calls smo_kdyn_peer_attach with the concatenation.  Without
 * attachment a multi-rank handle uses grouped ncclSend/ncclRecv. */
int smo_kdyn_peer_handle_bytes(void);
int smo_kdyn_peer_export(smo_kdyn_t* h, void* handles_out);
int smo_kdyn_peer_attach(smo_kdyn_t* h, const void* all_handles);
/* tuning: number of z chunks of the y-pass -> fused x-pass -> y-pass sequence of a forward / adjoint step (keeps the
 * y-padded arrays L2 resident); -1 = choose from the problem size, 1 = off (default) */
int smo_kdyn_set_chunks(smo_kdyn_t* h, int chunks_fwd, int chunks_adj);
/* tuning / A-B switches.  SMO_OPT_FUSED_Z: 1 (default) = forward-z FFT + implicit update + inverse-z FFT of a time step
 * run as ONE kernel (csrc/zstep.cuh); 0 = the three separate kernels. */
#define SMO_OPT_FUSED_Z 1
/* SMO_OPT_KERNEL_SYNC: 1 (default) = with peer-memory transposes attached, the cross-GPU hand-shakes are fused into the
 * kernels (the producer's last CTA signals, the consumer's CTAs wait); 0 = one barrier launch per transpose. */
#define SMO_OPT_KERNEL_SYNC 2
/* SMO_OPT_PEER_PULL: 0 (default) = peer-memory transposes are fused into the producers' stores (push over NVLink);
 * 1 = fused into the consumers' loads (cp.async straight out of the peers' buffers, stores stay local). */
#define SMO_OPT_PEER_PULL 3
/* SMO_OPT_L2_HINTS: 1 (default) = the pencil data exchanged between the y passes and the fused z step of a time step is
 * stored evict_last / read evict_first (single rank), so that it can stay L2 resident between the launches. */
#define SMO_OPT_L2_HINTS 4
/* SMO_OPT_PUSH_WAVES: n >= 1 (default 1): kernels that push their results into the peers' memory run about n work items per
 * CTA one after the other, so that the remote stores of the first items drain over NVLink while the later ones compute. */
#define SMO_OPT_PUSH_WAVES 5
/* SMO_OPT_TWO_STREAMS: 1 = the z chunks of the y -> fused x -> y section (smo_kdyn_set_chunks) alternate between two CUDA
 * streams; 2 = the same as a software pipeline (a kernel of chunk c+1 starts after the same kernel of chunk c), so that one
 * chunk's NVLink transfer and hand-shake run beside the next chunk's x pass; 0 (default) = one stream. */
#define SMO_OPT_TWO_STREAMS 6
/* SMO_OPT_GRID_ACC: 1 (default) = the fused adjoint x pass adds its (curl G) x B_f products to a running sum ON THE GRID (tile-major,
 * read-modify-write straight from registers) and skips their r2c transform; one transform after the sweep replaces 3 of the 6
 * forward FFTs of every adjoint step (KD:874-877 is linear in the products).  0 = running sum on the x-spectra. */
#define SMO_OPT_GRID_ACC 7
/* SMO_OPT_BULK_U: 1 (default) = the fused x passes fetch the velocity tile of a column tile (18 KB at 128^3, stored in HBM in its
 * swizzled shared-memory order) with ONE TMA bulk copy (cp.async.bulk + mbarrier) instead of 16-byte cp.async copies. */
#define SMO_OPT_BULK_U 8
/* SMO_OPT_TMA_SIN: 1 (default) = the fused x passes fetch their spectral tiles (4 columns x Npts/2 rows per field, 64-byte pieces at the
 * row pitch of the x-spectral arrays) with TMA tensor copies (cp.async.bulk.tensor, 3-D tiled maps with hardware swizzle,
 * one copy per field and tile) instead of 16-byte cp.async copies; 0 = cp.async. */
#define SMO_OPT_TMA_SIN 9
/* SMO_OPT_PDL: 1 = inside the time loops of a single-rank handle every kernel is launched with programmatic stream
 * serialisation (programmatic dependent launch): the CTAs of launch n+1 become resident and run their prologue (shared-memory
 * set-up, mbarrier initialisation, twiddle tables) while the last CTAs of launch n drain, and block in griddepcontrol.wait until
 * launch n has completed and flushed.  Captured into the CUDA graphs as programmatic edges; results are bit-identical.
 * 0 = plain stream order; -1 (default) = automatic: on for Npts <= 64 (launch-latency-bound steps: -16 % at 16^3, -10 % at 24^3,
 * -7 % at 32^3), off above (measured 5 % slower at 128^3 and 256^3).  Always off with several ranks, two streams or per-kernel
 * profiling events. */
#define SMO_OPT_PDL 10
/* SMO_OPT_BULK_PUSH: peer-memory transposes of the time loops (push mode) through the TMA: the results are staged in shared memory
 * and shipped to their owner with bulk stores (cp.async.bulk.global.shared::cta) instead of 16-byte stores from every thread, so that
 * remote stores do not queue in the SM's load/store path in front of the local loads.  bit 0 = fused z step (one store per warp and
 * peer), bit 1 = forward y pass (one store per truncated row). */
#define SMO_OPT_BULK_PUSH 11
int smo_kdyn_set_option(smo_kdyn_t* h, int key, int value);
/* capture each time loop into a CUDA graph and replay it (launch-bound small grids); 0 = off (default) */
int smo_kdyn_use_graph(smo_kdyn_t* h, int on);

/* ------------------------------------------------------------------------------------------------------------
 * Communicator for the slab-decomposed dynamo (replaces the MPI communicator Dedalus uses for its layout
 * transposes, [D2-9]; FWD_Solve_KDyn.py:3 `from mpi4py import MPI`).  One process per GPU: rank 0 calls
 * smo_comm_get_unique_id, the bytes are broadcast by the host program (e.g. torch.distributed), every rank calls
 * smo_comm_create and passes the result as nccl_comm to smo_kdyn_create.
 * ---------------------------------------------------------------------------------------------------------- */
int smo_comm_unique_id_bytes(void);
int smo_comm_get_unique_id(void* id_out);
int smo_comm_create(void** comm, const void* id, int nranks, int rank);
int smo_comm_destroy(void* comm);

/* ------------------------------------------------------------------------------------------------------------
 * Inner product and sphere geometry on device vectors (rows A5/B5 and C1-C3)
 * ---------------------------------------------------------------------------------------------------------- */
/* bytes of device workspace needed by the reductions below for vectors of n elements */
size_t smo_vec_work_bytes(long long n);
/* Replaces Inner_Prod / Inner_Prod_3 (SH:158-172, KD:173-181): *out_host = scale * sum_j x_j*y_j (scale = 1/M or
 * 1/M^3 gives the reference's grid mean).  Deterministic; synchronises the stream. */
int smo_vec_dot(const double* x_dev, const double* y_dev, long long n, double scale, double* out_host,
                void* work_dev, void* stream);
/* the same without the D2H copy / synchronisation: the scaled sum is left in ((double*)work_dev)[0] */
int smo_vec_dot_dev(const double* x_dev, const double* y_dev, long long n, double scale, void* work_dev, void* stream);
/* batched form of Inner_Prod (SH:158-172) for many short vectors (ensembles of SH23 problems):
 * out_dev[r] = scale * sum_j x[r][j]*y[r][j], r < rows; one launch, no sync */
int smo_vec_dot_rows(const double* x_dev, const double* y_dev, int rows, long long len, double scale, double* out_dev, void* stream);
/* 64-bit position-sensitive checksum of the bit patterns of x (sum_i bits(x_i)*(2i+1) mod 2^64): the host layer's identity check
 * of the X a snapshot store was filled for (the f -> Grad_f coupling of SGD:740-796).  Synchronises the stream. */
int smo_vec_checksum(const double* x_dev, long long n, unsigned long long* out_host, void* work_dev, void* stream);
/* out = a*x + b*y (y may be NULL when b == 0): the numpy algebra of SGD:642, 659, 687-690, 772, 776 */
int smo_vec_axpby(double a, const double* x_dev, double b, const double* y_dev, double* out_dev, long long n,
                  void* stream);
/* Replaces tangent_vector / transport_vector (SGD:625-659): out = v - (<x,v>/<x,x>) x, no host round trip */
int smo_vec_project(const double* x_dev, const double* v_dev, double* out_dev, long long n, void* work_dev,
                    void* stream);
/* Replaces Update_vector (SGD:661-690): f = x + alpha*d ; out = f*sqrt(M0/(scale*sum f_j^2)) */
int smo_vec_retract(const double* x_dev, double alpha, const double* d_dev, double M0, double scale, double* out_dev,
                    long long n, void* work_dev, void* stream);

/* measurement aid (bench.py): one launch of dependent fp64 FMA chains on every SM (ctas_per_sm x 256 threads, 8 chains each);
 * *flops_out = flop of the launch.  Timed by the caller with CUDA events: the measured DFMA peak that the SH23 ensemble's
 * fp64 roofline is quoted against.  out_dev: >= 256*ctas_per_sm*#SMs doubles (never written in practice). */
int smo_microbench_dfma(double* out_dev, int iters, int ctas_per_sm, double* flops_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SMO_B200_H */
