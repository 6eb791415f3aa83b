/* =============================================================================
 * cfdb200.h -- C ABI of the B200-native fractional-step FE Navier-Stokes time loop.
 *
 * Drop-in boundary for the GPU seam of mmgocmen/cfd-with-cuda: the set of
 * `...GPU()` functions that the reference's host code calls under `#ifdef USECUDA`
 *   explicit (Blasco-Codina-Huerta):  fractionalStep/explicit/Cpp/blascoCodinaHuerta.cpp:365-377
 *   implicit (Guermond-Quartapelle):  fractionalStep/implicit/Cpp/guermondQuartapelle.cpp:350-361
 * The reference shares ~60 `extern` globals with its CUDAcodes.cu; here the same
 * arrays travel in one POD struct (`cfd_problem`, caller-owned, copied at init) and
 * all device state lives in an opaque `cfd_ctx`.  Names keep the reference's
 * (`step1GPUpart`, `step2GPU`, ...) with a `bch_` / `gq_` prefix for the two solvers.
 *
 * All functions return 0 on success, a negative code on failure
 * (`cfd_last_error()` gives the message).  Nothing here blocks on stdin and no
 * function allocates device memory after `cfd_initializeAndAllocateGPU`.
 * Plain pointers and sizes only: no C++ or torch types cross this boundary.
 * ============================================================================= */
#ifndef CFDB200_H
#define CFDB200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CFD_EXPLICIT 0 /* blascoCodinaHuerta */
#define CFD_IMPLICIT 1 /* guermondQuartapelle */

#define CFD_OK 0
#define CFD_ERR_ARG (-1)
#define CFD_ERR_CUDA (-2)
#define CFD_ERR_IO (-3)
#define CFD_ERR_NCCL (-4)
#define CFD_ERR_STATE (-5)

typedef struct cfd_ctx cfd_ctx;

/* The data model a replacement consumes.  Field names are the
 * reference's globals; every array is host memory, read-only, copied during init. */
typedef struct cfd_problem {
   int scheme;                /* CFD_EXPLICIT | CFD_IMPLICIT */
   int NE, NN, NNp;           /* elements, velocity nodes, pressure nodes (BCH.cpp:150-160) */
   int NENv, NENp, NGP;       /* 27|8, 8, 8 */
   double dt, density, viscosity;
   int maxIter;               /* explicit inner iterations per step (inp `maxIter`) */
   double tolerance;          /* explicit inner-iteration tolerance (inp `tolerance`) */
   double convergenceCriteria;/* steady-state tolerance (inp `converge`) */
   const double* coord;       /* [NN][3]   (BCH.cpp coord) */
   const int* LtoGnode;       /* [NE][NENv] (BCH.cpp LtoGnode; LtoGvel_1d = c*NN + LtoGnode) */
   /* one NN x NN pattern shared by K, M, A (BCH.cpp:212-219) */
   int sparseM_nnz;           /* nnz of ONE block (= reference sparseM_NNZ / 3) */
   const int* sparseMrowStarts;
   const int* sparseMcol;
   const double* sparseKvalue;
   const double* sparseMvalue;  /* implicit only: int N_i N_j / dt (GQ.cpp:3495) */
   /* one NN x NNp pattern shared by G1, G2, G3 (BCH.cpp:221-229) */
   int sparseG_nnz;           /* nnz of ONE block (= reference sparseG_NNZ / 3) */
   const int* sparseGrowStarts;
   const int* sparseGcol;
   const double* sparseG1value;
   const double* sparseG2value;
   const double* sparseG3value;
   /* pressure matrix Z, full (both triangles), 0-based, pin already applied (BCH.cu:281-411, GQ.cpp:3677-3772) */
   int sparseZ_nnz;
   const int* sparseZrowStarts;
   const int* sparseZcol;
   const double* sparseZvalue;
   const int* sparseMapM_1d;  /* implicit only: [NE][NENv][NENv] CSR positions (BCH.cpp:1887-1898) */
   const double* MdInv;       /* explicit only: [3NN], BC rows = 1 (BCH.cpp:3283-3292) */
   const double* MdOrigInv;   /* explicit only: [3NN] */
   int BCnVelNodes;
   const int* BCvelNodes;     /* [BCnVelNodes][2] = node, whichBC (BCH.cpp:199) */
   int nBC;
   const double* BCstr;       /* [nBC][3] */
   int zeroPressureNode;      /* 0-based; applied only when > 0 (BCH.cpp:4172) */
   int monPoint;              /* monitor node (BCH.cpp:1646-1668) */
   const double* U0;          /* [3NN] initial Un (explicit) / Uk_prev (implicit), BCs applied; NULL = zeros + BCs */
   const double* P0;          /* [NNp] initial pressure; NULL = zeros */
   /* ---- element-block partition (multi-GPU; all zero / NULL for a whole mesh) -------------------------------
    * The reference is single-GPU; this part has no counterpart there.  NN / NNp above are then the
    * LOCAL counts (owned + ghost), owned entries numbered first; CSR rows beyond the owned ones are ignored. */
   int world, rank;
   int NN_owned, NNp_owned;   /* 0 = all */
   int NE_apply;              /* leading elements the element-by-element kernels loop over (0 = NE) */
   int pinLocal;              /* partitioned runs: local pressure DOF to pin or -1 (zeroPressureNode's `> 0` rule already applied) */
   int monOwned;              /* partitioned runs: 1 if monPoint is owned by this rank */
   int nPeers;
   const int* peerRank;       /* [nPeers] */
   const int* sendVptr;       /* [nPeers+1] offsets into sendVidx */
   const int* sendVidx;       /* owned local velocity nodes each peer holds as ghosts */
   const int* recvVoff;       /* [nPeers] first ghost node filled by the peer */
   const int* recvVcnt;       /* [nPeers] */
   const int* sendPptr;       /* same for pressure DOFs */
   const int* sendPidx;
   const int* recvPoff;
   const int* recvPcnt;
   const int* remoteVoff;     /* [nPeers] start of MY values inside the peer's velocity ghost range (peer-memory pushes) */
   const int* remotePoff;     /* [nPeers] same for pressure DOFs */
} cfd_problem;

/* Solver settings the reference hard-codes (BCH.cu:433, GQ.cu:344,404) plus the
 * switches that select between its four mutually inconsistent builds. */
typedef struct cfd_options {
   double cg_tol;     /* default 1e-6: stop when ||r||_2 <= cg_tol * ||b||_2 */
   int cg_maxit;      /* default 1000 */
   int cg_warm_start; /* explicit only. 0 (default) = x0 = 0 (BCH.cpp:3997); 1 = previous Pdot (BCH.cu:418-451) */
   double bicg_tol;   /* default 1e-6 */
   int bicg_maxit;    /* default 1000 */
   int maxacc_rule;   /* 0 (default) = max|dU|/dt (BCH.cpp:3053); 1 = max|dU| (BCH.cu:107-117) */
   int device;        /* CUDA device ordinal; -1 = current device */
   int ebe_scatter;   /* element-by-element scatter: 0 (default) = FP64 atomicAdd, one launch; 1 = element colouring, one launch
                         per colour, plain stores, bitwise reproducible (the reference's scheme, BCH.cu:906-939) */
   int cg_no_precond; /* 0 (default) = Jacobi-preconditioned CG (BCH.cpp:4017, BCH.cu:437, GQ.cu:348); 1 = plain CG, what the
                         reference's implicit CPU build runs (its Jacobi line is commented out, GQ.cpp:4514) */
   int deterministic; /* 0 (default) = the persistent Krylov kernels hand the tail of each SpMV sweep out through a shared counter
                         (load balance; dot products then depend on which warp drew which slice: reproducible to rounding);
                         1 = static slice assignment only: bitwise reproducible from run to run (with ebe_scatter = 1) */
   int bicg_no_halfstep; /* 0 (default) = BiCGStab also tests |s| after the half step, like cusp::krylov::bicgstab (GQ.cu:404-414);
                            1 = one test per iteration, where a PARALUTION-style solver (GQ.cpp:4190-4229) stops */
   int bicg_no_verify;   /* 0 (default) = a momentum solve is accepted on its TRUE residual |b - A x| <= bicg_tol |b| and restarted from it
                            otherwise (breakdowns included); 1 = trust the recurrence residual, as the reference's libraries do */
   int cg_variant;    /* pressure CG recurrences: 0 / 1 (default) = classical two-reduction PCG, as the reference runs it (BCH.cu:418-451);
                         2 = Chronopoulos-Gear single-reduction form (same iterates in exact arithmetic, one grid-wide / cross-GPU
                         reduction per iteration instead of two, 16 B per row more vector traffic).  Both meet the parity bar; the
                         classical form measured faster on B200 at every size tried (profiles/r02_cg_variants.md) */
   int operator_path; /* 0 (default) = K, G, G^T applied from the assembled CSR arrays (Sell layout); 1 = matrix-free: the explicit scheme's
                         operator applications (step 1: -C(U)U - K U - G P in ONE element pass, step 2: G^T d, step 3: G Pdot, K Acc) are
                         evaluated element by element from the integrals of step0 (BCH.cpp:3127-3315), no matrix stream.  Same
                         results to rounding.  Single GPU; the pressure solve keeps its assembled Z */
   int step0_device;  /* 0 (default) = the matrix value arrays of cfd_problem (K, M, G1..3, Z, MdInv, MdOrigInv) come from the host, as the
                         reference's step0 / calculateZ leave them (BCH.cpp:3127-3479, GQ.cpp:3425-3772); 1 = those pointers may be NULL: the
                         element integrals, the lumped mass, Z (explicit: G^T Md^-1 G as a device sparse product; implicit: FE Laplacian)
                         and the pin are evaluated on the GPU from coord / LtoGnode and the three CSR patterns (columns ascending within
                         each row).  Equal to the host arrays to rounding.  Single GPU */
   void* cuda_stream; /* cudaStream_t to launch on (e.g. torch's current stream); NULL = own stream */
} cfd_options;

void cfd_default_options(cfd_options* opt);
const char* cfd_last_error(void);
const char* cfd_version(void);
/* sizeof(cfd_problem), sizeof(cfd_options), sizeof(cfd_step_report) as this library was compiled: lets a binding written in
 * another language (ctypes, cgo, JNI) check its mirror of the three structs at load time */
void cfd_struct_sizes(int out[3]);

/* replaces selectCUDAdevice()            BCH.cu:126-154 (no "last device" heuristic: explicit ordinal) */
int cfd_selectCUDAdevice(int device);
/* replaces initializeAndAllocateGPU()    BCH.cu:161-273 / GQ.cu:166-295: uploads the problem, converts
 * the CSR matrices to the device layout, allocates every work vector. */
int cfd_initializeAndAllocateGPU(const cfd_problem* prob, const cfd_options* opt, cfd_ctx** out);
void cfd_destroy(cfd_ctx* ctx);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink ------------------------------------------------------
 * rank 0 calls cfd_nccl_unique_id, the bytes are broadcast by the launcher (torch.distributed / MPI), then every
 * rank attaches its partitioned context.  Collectives used: grouped send/recv of interface nodal values before
 * each operator application, allreduce (sum / max) of Krylov dot products, convergence sums and maxAcc. */
int cfd_nccl_unique_id(char out[128]);
int cfd_comm_init(cfd_ctx* ctx, const char id[128]);
/* Optional peer-memory path (NVLink P2P through CUDA IPC): each rank exports its exchange window, the launcher
 * all-gathers the 96-byte descriptors, every rank maps its peers' windows.  With it the whole pressure solve runs
 * as ONE persistent kernel per GPU that pushes interface values of p straight into the neighbours' ghost entries
 * and all-reduces its dot products through peer mailboxes -- no host round trip and no NCCL call per iteration. */
int cfd_p2p_export(cfd_ctx* ctx, char desc[96]);
int cfd_p2p_attach(cfd_ctx* ctx, const char* all_desc /* [world][96] */);
/* After a peer-memory time-out (CFD_ERR_STATE from a step: a rank stopped participating for longer than CFD_P2P_TIMEOUT_MS,
 * default 4000): zero this rank's exchange window and epochs.  Collective in the sense that EVERY rank must call it, with a
 * host-side barrier before and after, before the contexts are used again. */
int cfd_p2p_reset(cfd_ctx* ctx);
/* Orderly teardown of the peer-memory path: unmaps the peers' windows (this rank's own window stays until cfd_destroy).
 * Call on EVERY rank, then synchronise the ranks on the host (barrier), then destroy the contexts: CUDA forbids freeing
 * exported memory while another process still has it mapped.  (cfd_destroy unmaps on its own if this was not called, which is
 * only safe when the peers are known to be gone.) */
int cfd_p2p_detach(cfd_ctx* ctx);

/* named device vectors: "Un","Unp1","Unp1_prev","UnpHalf","UnpHalf_prev","Acc","Acc_prev",
 * "KtimesAcc_prev","R1","R2","R3","Pn","Pnp1","Pnp1_prev","Pdot" (explicit);
 * "Uk","Uk_prev","Pk","Pk_prev","Pk_prevprev","Pdiff","R1","R2" (implicit; plus "Adinv", download only: the Jacobi diagonal of the
 * momentum matrix -- A itself is kept in the device layout, not in CSR order, and is not downloadable). */
int cfd_download(cfd_ctx* ctx, const char* name, double* host, long long count);
int cfd_upload(cfd_ctx* ctx, const char* name, const double* host, long long count);
/* Raw device address of a named vector (names as above; NULL if unknown).  For host programs that -- like the reference's
 * own timeLoop under USECUDA (BCH.cpp:2866-2869, 2929-2932, 3070-3071, 3099-3100; GQ.cpp:3366-3368) -- move data between
 * iterates with their own cudaMemcpy calls on the `*_d` globals: integration/cfdb200_shim_*.cpp points those globals here.
 * Buffers are never re-allocated.  The seam functions the reference calls leave the addresses alone (bch_rotateIterates and
 * bch_endTimeStep copy, like BCH.cpp:2929-2932 / 3070-3071); gq_endTimeStep and cfd_timeLoop rotate the iterates by swapping
 * buffers instead (implicit: Uk/Uk_prev, Pk/Pk_prev/Pk_prevprev; explicit: UnpHalf/UnpHalf_prev, Unp1/Unp1_prev, Acc/Acc_prev,
 * Pnp1/Pnp1_prev between inner iterations, Un/Unp1 and Pn/Pnp1 at the end of a step) -- a caller that uses those must fetch the
 * pointers again afterwards; the named download / upload calls always follow the current assignment.  Launch on the same stream (cfd_options.cuda_stream) or synchronise the device
 * between such copies and the seam calls, as the reference does (cudaThreadSynchronize). */
void* cfd_device_pointer(cfd_ctx* ctx, const char* name);
/* current velocity [3NN] and pressure [NNp] (Un/Pn or Uk_prev/Pk_prev after a completed step) */
int cfd_get_fields(cfd_ctx* ctx, double* U, double* P);

/* ---- explicit seam (call order = timeLoop, BCH.cpp:2859-3118) ------------------------------- */
int bch_beginTimeStep(cfd_ctx* ctx);               /* D2D copies + memset at BCH.cpp:2866-2869 */
int bch_calculateMatrixAGPU(cfd_ctx* ctx);         /* BCH.cu:906-939  R1 = -C(Un) UnpHalf_prev */
int bch_step1GPUpart(cfd_ctx* ctx, int iter);      /* BCH.cu:946-973 */
int bch_step2GPU(cfd_ctx* ctx, int iter);          /* BCH.cu:980-1022 (+ CUSP_CG_solver :418-451) */
int bch_step3GPU(cfd_ctx* ctx, int iter);          /* BCH.cu:1029-1060 */
int bch_checkConvergenceGPU(cfd_ctx* ctx, int* converged); /* BCH.cu:1084-1121 */
int bch_rotateIterates(cfd_ctx* ctx);              /* the four D2D copies at BCH.cpp:2929-2932 */
int bch_calculate_KtimesAcc_prevGPU(cfd_ctx* ctx); /* BCH.cu:1067-1077 */
int bch_checkConvergenceInTimeGPU(cfd_ctx* ctx, int* steady, double* maxAcc); /* BCH.cu:1151-1178 */
int bch_endTimeStep(cfd_ctx* ctx);                 /* Un<-Unp1, Pn<-Pnp1, BCH.cpp:3070-3071 */

/* ---- implicit seam (GQ.cpp:3308-3416) ------------------------------------------------------- */
int gq_calculateMatrixAGPU(cfd_ctx* ctx);          /* GQ.cu:726-754 (A += C(Uk_prev); A must hold M+K) */
int gq_step1GPU(cfd_ctx* ctx);                     /* GQ.cu:761-890 */
int gq_step2GPU(cfd_ctx* ctx);                     /* GQ.cu:897-944 */
int gq_checkConvergenceInTimeGPU(cfd_ctx* ctx, int* steady, double* maxAcc); /* GQ.cu:975-1002 */
int gq_endTimeStep(cfd_ctx* ctx);                  /* rotations at GQ.cpp:3366-3368 */

/* replaces printMonitorDataGPU()  BCH.cu:1128-1144 / GQ.cu:951-967: out[4] = u,v,w,p at monPoint
 * (a 4-double read-back instead of the reference's full-field D2H). */
int cfd_printMonitorDataGPU(cfd_ctx* ctx, double out[4]);

/* Krylov statistics of the most recent step2 / step1 call */
int cfd_last_cg_iterations(cfd_ctx* ctx);
int cfd_last_bicgstab_iterations(cfd_ctx* ctx, int its[3]);
/* status of the most recent solves: converged flags, relative residuals at exit (pressure: recurrence residual; momentum: true
 * residual ||b - A x|| / ||b||, -1 if bicg_no_verify), restarts of each momentum system.  Any pointer may be NULL. */
int cfd_last_krylov_status(cfd_ctx* ctx, int* cg_converged, double* cg_residual, int bicg_converged[3], double bicg_true_residual[3],
                           int bicg_restarts[3]);

/* ---- whole time steps (the fused fast path; same arithmetic as the seam sequence) ------------ */
typedef struct cfd_step_report {
   int timeN;            /* step counter after this step */
   int iter;             /* explicit: value the reference prints (maxIter+1 if not converged); implicit: 1 */
   int cg_iterations;    /* summed over the step's pressure solves */
   int cg_solves;
   int bicg_iterations[3];
   double maxAcc;
   int steady;           /* maxAcc <= convergenceCriteria */
   double monitor[4];    /* u,v,w,p at monPoint */
   /* Krylov status of the step (the reference's solvers report nothing: MKL failure -> print and return, BCH.cpp:4007) */
   int cg_converged;     /* every pressure solve of the step met ||r|| <= cg_tol ||b|| (recurrence residual) within cg_maxit */
   int bicg_converged[3];        /* implicit: momentum system d met the stopping rule on its TRUE residual (0 = stopped unconverged) */
   int bicg_restarts[3];         /* restarts from the true residual (0 in the normal case) */
   double cg_residual;           /* max over the step's pressure solves of ||r|| / ||b|| at exit */
   double bicg_true_residual[3]; /* ||b - A x|| / ||b|| at exit (-1 when verification is off) */
} cfd_step_report;
/* runs `nsteps` passes of the time-loop body; `reports` may be NULL or hold nsteps entries.
 * Stops early (returning the number of steps done in *done) when a step reports steady state
 * and stop_on_steady != 0. */
int cfd_timeLoop(cfd_ctx* ctx, int nsteps, int stop_on_steady, cfd_step_report* reports, int* done);
/* same, bracketed by CUDA events on the context's stream: *device_ms = elapsed device time */
int cfd_timeLoop_timed(cfd_ctx* ctx, int nsteps, cfd_step_report* reports, double* device_ms);
/* One time step driven from HOST buffers (the call a host program makes): copies U [3NN] and P [NNp]
 * to the device, runs one time-loop pass, copies the new U and P back into the same buffers.
 * *device_ms (optional) = CUDA-event time of copy-in + step + copy-out. */
int cfd_timeStep_host(cfd_ctx* ctx, double* U, double* P, cfd_step_report* report, double* device_ms);

/* ---- timing / roofline instrumentation --------------------------------------------------------
 * Per-kernel-class CUDA-event time on the context's stream (enabled by cfd_profile(ctx,1)): event
 * pairs are recorded around every launch without synchronising and resolved on cfd_profile_get.
 * classes: "ebe","step1","gt_rhs","pcg","update","step3","spmv_k","assemble","gq_rhs","bicgstab";
 * `bytes` = algorithmic bytes of the launches (DESIGN.md byte model). */
int cfd_profile(cfd_ctx* ctx, int enable);
int cfd_profile_get(cfd_ctx* ctx, const char* name, double* ms, long long* launches, double* bytes);
/* the same launches counted in the bytes the device layout STORES (zero padding of the sliced layout included, per-slice compressed
 * column indices counted once): the traffic the kernel cannot avoid, so achieved-GB/s figures formed with it cannot exceed the
 * memory bandwidth -- the algorithmic count above can (structured meshes store fewer index bytes than the CSR model charges) */
int cfd_profile_get_layout_bytes(cfd_ctx* ctx, const char* name, double* layout_bytes);
long long cfd_kernel_launches(cfd_ctx* ctx); /* kernels launched by this context so far */
/* layout statistics of a device matrix ("K", "G", "GT", "Z"): true non-zeros, stored value slots (padding included),
 * stored index integers (after per-slice offset compression), lanes per row */
int cfd_matrix_stats(cfd_ctx* ctx, const char* name, long long* nnz, long long* stored_values, long long* stored_indices, int* lanes_per_row);

/* One operator of the path applied both ways on a context's mesh (micro-benchmark "EBE vs CSR", parity test of the matrix-free
 * kernels): op = "K" (x [3NN] -> y [3NN]), "G" (x [NNp] -> y [3NN]), "GT" (x [3NN] -> y [NNp]), "Z" (implicit contexts: the FE
 * Laplacian, x [NNp] -> y [NNp]); path 0 = assembled (Sell SpMV), 1 = matrix-free element kernels.  x, y are host arrays;
 * *ms_per_call = best CUDA-event time of `repeat` launches; *bytes = algorithmic bytes of one launch in that path's own model. */
int cfd_operator_apply(cfd_ctx* ctx, const char* op, int path, const double* x, double* y, int repeat, double* ms_per_call, double* bytes);

/* ---- stand-alone kernels on caller data (micro-benchmarks and parity tests) ------------------- */
/* y = alpha*A*x + beta*y for `nvec` vectors sharing one CSR matrix; x is [nvec][ncols], y [nvec][nrows] host arrays */
int cfd_csr_spmv(int nrows, int ncols, const int* rowStarts, const int* col, const double* val, int nvec,
                 const double* x, double* y, double alpha, double beta, int repeat, double* ms_per_call);
/* Host-side self-check of the device matrix layout (no GPU needed): converts the CSR matrix to the sliced layout with T
 * lanes per row (T = 0: the library's own choice), decodes it again -- explicit and offset-compressed slices -- and
 * returns the number of (row, col, value) triples that differ (0 = exact); *compressed_slices / *slices report how many
 * slices store a shared column-offset list instead of explicit columns. */
long long cfd_sell_layout_check(int nrows, int ncols, const int* rowStarts, const int* col, const double* val, int T,
                                long long* slices, long long* compressed_slices, double* stored_over_nnz);
/* diagnostic: microseconds per grid-wide sync (mode 0) / deterministic grid reduction of 1 (mode 1) or 3 (mode 2) values */
int cfd_diag_sync(int blocks, int threads, int iters, int mode, double* us_per_op);
/* Jacobi-PCG on a CSR matrix: returns iterations in *its */
int cfd_csr_pcg(int n, const int* rowStarts, const int* col, const double* val, const double* b, double* x,
                double tol, int maxit, int* its, double* ms);
/* same with the recurrences chosen explicitly: variant 1 = classical, 2 = Chronopoulos-Gear (see cfd_options.cg_variant) */
int cfd_csr_pcg_variant(int n, const int* rowStarts, const int* col, const double* val, const double* b, double* x,
                        double tol, int maxit, int variant, int* its, double* ms);

/* =============================================================================
 * Solver-level API: the whole `solverGPU` program of the reference as a library
 * (readInputFile -> preprocessing -> step0 -> timeLoop -> createTecplot), host side
 * written from scratch in C++ (csrc/host_*.cpp).
 * ============================================================================= */
typedef struct cfd_solver cfd_solver;

/* Reads `<path>` (the reference's .inp grammar, BCH.cpp:528-725), generates mid-edge/face/centre
 * nodes with the reference's numbering, builds CSR patterns, assembles M,K,G,Z (step0). */
int cfd_solver_create_from_inp(const char* inp_path, int scheme, const cfd_options* opt, int init_gpu, cfd_solver** out);
/* Same, for an in-memory structured nx*ny*nz cavity (cube generator numbering)
 * without the text round trip: used for the 128^3 / 256^3 synthetic configs.  NENv = 8 or 27. */
int cfd_solver_create_cavity(int nx, int ny, int nz, double cluster, int NENv, int scheme, double dt,
                             double viscosity, int maxIter, double tolerance, const cfd_options* opt,
                             int init_gpu, cfd_solver** out);
/* Partitioned variants: this process holds rank `rank` of `world` element blocks (own elements + halo layers). */
int cfd_solver_create_from_inp_part(const char* inp_path, int scheme, const cfd_options* opt, int init_gpu, int world, int rank,
                                    cfd_solver** out);
int cfd_solver_create_cavity_part(int nx, int ny, int nz, double cluster, int NENv, int scheme, double dt, double viscosity,
                                  int maxIter, double tolerance, const cfd_options* opt, int init_gpu, int world, int rank,
                                  cfd_solver** out);
/* init_gpu = 0 above stops after step0 (host arrays only, no CUDA call: what the CPU-side tests
 * use); this uploads the problem afterwards. */
int cfd_solver_init_gpu(cfd_solver* s);
void cfd_solver_destroy(cfd_solver* s);
cfd_ctx* cfd_solver_ctx(cfd_solver* s);
const cfd_problem* cfd_solver_problem(cfd_solver* s);
/* integer facts by name: "NE","NCN","NN","NNp","nnzM","nnzG","nnzZ","BCnVelNodes","monPoint",
 * "zeroPressureNode","nActiveColors" */
long long cfd_solver_int(cfd_solver* s, const char* name);
/* host arrays by name (reference global names, see cfd_problem) -> pointer + count */
const void* cfd_solver_array(cfd_solver* s, const char* name, long long* count, int* is_int);
/* createTecplot() BCH.cpp:4249-4482 and readRestartFile() BCH.cpp:4214-4242 */
int cfd_solver_write_tecplot(cfd_solver* s, const char* dat_path, const char* title);
/* same writer on caller-supplied host fields U [3NN], P [NNp] (no GPU needed) */
int cfd_solver_write_tecplot_fields(cfd_solver* s, const char* dat_path, const char* title, const double* U, const double* P);
/* restart reader into host buffers U [3NN], P [NNp] (no GPU needed) */
int cfd_solver_read_restart_fields(cfd_solver* s, const char* dat_path, double* U, double* P);
int cfd_solver_read_restart(cfd_solver* s, const char* dat_path);
/* the reference's main(): reads ./ProblemName.txt, runs to t_final, prints the monitor table,
 * writes <name>.dat.  max_steps > 0 bounds the run.  Solver settings: cfd_default_options (= the reference's GPU builds),
 * overridden by the environment variables CFD_CG_TOL, CFD_CG_MAXIT, CFD_CG_WARM_START, CFD_CG_NO_PRECOND, CFD_BICG_TOL,
 * CFD_BICG_MAXIT, CFD_MAXACC_RULE, CFD_DEVICE, CFD_EBE_SCATTER, CFD_DETERMINISTIC, CFD_BICG_NO_HALFSTEP, CFD_BICG_NO_VERIFY,
 * CFD_CG_VARIANT (same meaning as the cfd_options fields).
 * csrc/solver_main.cpp wraps this into the executables bin/solverGPU_explicit and bin/solverGPU_implicit, which -- like the
 * reference's solverGPU (makeBCH:16-18) -- take no arguments and work on the current directory. */
int cfd_solver_main(const char* workdir, int scheme, int max_steps);

#ifdef __cplusplus
}
#endif
#endif /* CFDB200_H */
