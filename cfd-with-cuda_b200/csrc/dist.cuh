// Multi-GPU layer (product code): one process per GPU, NCCL over NVLink / NVSwitch.
//
// The reference is single-GPU (selectCUDAdevice picks exactly one device, BCH.cu:126-154), so
// nothing here mirrors reference code.  Partition: contiguous element blocks, owner-computes rows, ghost copies
// of the interface nodes (host_mesh.cpp: partition_mesh).  Exchange steps:
//   * halo swap  -- before an operator gathers a vector, every rank sends its owned interface values to the
//                   ranks holding them as ghosts (one grouped ncclSend/ncclRecv per neighbour and component;
//                   ghosts of one owner are contiguous, so receives land in place)
//   * allreduce  -- Krylov dot products, convergence sums (sum) and maxAcc (max)
// The Krylov solvers run as phase kernels driven from the host WITHOUT host synchronisation inside a batch of
// iterations: all scalars (alpha, beta, residual norms, iteration counter, `done`) live in device memory, kernels
// of an already converged solve return immediately, and the host reads the state once per batch.
// NCCL is loaded with dlopen so that the library has no link-time dependency on it (CPU-only test boxes) and
// shares the copy a host process (e.g. PyTorch) has already loaded.
#pragma once
// (included from device.cu inside namespace cfdb, after the row-kernel helpers)

typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId { char internal[128]; };
enum { kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2, kNcclMin = 3 };

struct Comm {
   void* lib = nullptr;
   ncclComm_t comm = nullptr;
   int world = 1, rank = 0;
   int (*GetUniqueId)(NcclUniqueId*) = nullptr;
   int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
   int (*CommDestroy)(ncclComm_t) = nullptr;
   int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
   int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
   int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
   int (*GroupStart)() = nullptr;
   int (*GroupEnd)() = nullptr;
   const char* (*GetErrorString)(int) = nullptr;
};

inline bool nccl_load(Comm& c) {
   if (c.lib) return true;
   c.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
   if (!c.lib) c.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
   if (!c.lib) { set_error("cannot load libnccl.so.2: %s", dlerror()); return false; }
#define LOADSYM(field, name)                                                        \
   *(void**)(&c.field) = dlsym(c.lib, name);                                        \
   if (!c.field) { set_error("libnccl: missing symbol %s", name); return false; }
   LOADSYM(GetUniqueId, "ncclGetUniqueId")
   LOADSYM(CommInitRank, "ncclCommInitRank")
   LOADSYM(CommDestroy, "ncclCommDestroy")
   LOADSYM(AllReduce, "ncclAllReduce")
   LOADSYM(Send, "ncclSend")
   LOADSYM(Recv, "ncclRecv")
   LOADSYM(GroupStart, "ncclGroupStart")
   LOADSYM(GroupEnd, "ncclGroupEnd")
   LOADSYM(GetErrorString, "ncclGetErrorString")
#undef LOADSYM
   return true;
}

#define CFD_NCCL(c, call)                                                                     \
   do {                                                                                       \
      int r_ = (call);                                                                        \
      if (r_ != 0) {                                                                          \
         cfdb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, (c)->GetErrorString(r_)); \
         return CFD_ERR_NCCL;                                                                 \
      }                                                                                       \
   } while (0)

// ---- pack the owned interface values of an ncomp-component vector into the staging buffer -----------------
// layout per peer j: [comp][k]  at  ncomp*ptr[j] + comp*cnt_j + k
__global__ void k_pack(int nsendTotal, int ncomp, long long ld, const int* __restrict__ idx, const int* __restrict__ segOfEntry,
                       const int* __restrict__ ptr, const double* __restrict__ vec, double* __restrict__ buf) {
   const int t = blockIdx.x * blockDim.x + threadIdx.x;
   if (t >= nsendTotal) return;
   const int j = segOfEntry[t];
   const int base = ptr[j], cnt = ptr[j + 1] - base, k = t - base;
   const int node = idx[t];
   for (int c = 0; c < ncomp; c++) buf[(long long)ncomp * base + (long long)c * cnt + k] = vec[c * ld + node];
}

// ---- distributed Jacobi-PCG phase kernels (same recurrences as k_pcg) ---------------------------------------
struct DPcgState {
   double rz, bb, rr;
   int it, done, maxit;
   double tol;
};
// acc layout: [0] pq | [1] rz' | [2] rr | [3..5] init sums (rz, bb, rr)

__global__ void k_dpcg_init(int n, int x0zero, const double* __restrict__ b, double* __restrict__ x, const double* __restrict__ q,
                            const double* __restrict__ dinv, double* __restrict__ r, double* __restrict__ p, double* __restrict__ acc) {
   __shared__ double sh[3 * 32];
   double v[3] = {0, 0, 0};
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const double bi = b[i];
      double ri = bi;
      if (x0zero) x[i] = 0.0; else ri = bi - q[i];
      const double zi = dinv[i] * ri;
      r[i] = ri; p[i] = zi;
      v[0] += ri * zi; v[1] += bi * bi; v[2] += ri * ri;
   }
   block_sum<3>(v, sh);
   if (threadIdx.x == 0) { atomicAdd(acc + 3, v[0]); atomicAdd(acc + 4, v[1]); atomicAdd(acc + 5, v[2]); }
}
__global__ void k_dpcg_begin(DPcgState* st, double* acc, double tol, int maxit) {
   st->rz = acc[3]; st->bb = acc[4]; st->rr = acc[5];
   st->it = 0; st->maxit = maxit; st->tol = tol;
   st->done = (sqrt(acc[5]) <= tol * sqrt(acc[4]) || maxit <= 0) ? 1 : 0;
   for (int i = 0; i < 6; i++) acc[i] = 0.0;
}
__global__ void __launch_bounds__(256) k_dpcg_spmv(Sell Z, const DPcgState* __restrict__ st, const double* __restrict__ p,
                                                   double* __restrict__ q, double* __restrict__ acc) {
   if (st->done) return;
   __shared__ double sh[32];
   RowMap m;
   double v[1] = {0.0};
   if (row_map(Z, m)) {
      const double* __restrict__ val = Z.vals[0];
      double a = 0.0;
      sell_pairs<4>(Z, m.slice, m.lane, [&](long long e, int2 c2) {
         const double2 v2 = ldnc_d2(val + e);
         a += v2.x * __ldg(p + c2.x) + v2.y * __ldg(p + c2.y);
      });
      a = lanes_sum(a, Z.T);
      if (m.owner) { q[m.row] = a; v[0] = p[m.row] * a; }
   }
   block_sum<1>(v, sh);
   if (threadIdx.x == 0) atomicAdd(acc + 0, v[0]);
}
__global__ void k_dpcg_update(int n, const DPcgState* __restrict__ st, const double* __restrict__ p, const double* __restrict__ q,
                              const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r, double* __restrict__ acc) {
   if (st->done) return;
   __shared__ double sh[2 * 32];
   const double alpha = st->rz / acc[0];
   double v[2] = {0, 0};
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      x[i] += alpha * p[i];
      const double ri = r[i] - alpha * q[i];
      r[i] = ri;
      v[0] += ri * (dinv[i] * ri);
      v[1] += ri * ri;
   }
   block_sum<2>(v, sh);
   if (threadIdx.x == 0) { atomicAdd(acc + 1, v[0]); atomicAdd(acc + 2, v[1]); }
}
__global__ void k_dpcg_pupdate(int n, const DPcgState* __restrict__ st, const double* __restrict__ r, const double* __restrict__ dinv,
                               double* __restrict__ p, const double* __restrict__ acc) {
   if (st->done) return;
   const double beta = acc[1] / st->rz;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = dinv[i] * r[i] + beta * p[i];
}
__global__ void k_dpcg_next(DPcgState* st, double* acc) {
   if (st->done) return;
   st->rz = acc[1]; st->rr = acc[2]; st->it += 1;
   if (sqrt(st->rr) <= st->tol * sqrt(st->bb) || st->it >= st->maxit) st->done = 1;
   acc[0] = acc[1] = acc[2] = 0.0;
}

// ---- distributed Jacobi-BiCGStab for the three momentum systems (same recurrences as k_bicgstab3) ------------
struct DBicgState {
   double bb[3], rr[3], rrs[3], alpha[3], omega[3], thresh[3];
   int its[3], act[3], half[3], act2[3];
   int done, maxit;
};
// acc layout: 4 sets of 6: [0..5] spmv1 (rs.AMp) | [6..11] |s|^2 | [12..17] spmv2 (AMs.s, AMs.AMs) | [18..23] (rs.r, r.r)

__global__ void k_dbicg_init(int n, long long ld, const double* __restrict__ b, const double* __restrict__ dinv, double* __restrict__ x,
                             double* __restrict__ r, double* __restrict__ rs, double* __restrict__ p, double* __restrict__ ph,
                             double* __restrict__ acc) {
   __shared__ double sh[3 * 32];
   double v[3] = {0, 0, 0};
   for (int c = 0; c < 3; c++)
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
         const long long k = c * ld + i;
         const double bi = b[k];
         x[k] = 0.0; r[k] = bi; rs[k] = bi; p[k] = bi;
         ph[k] = dinv[i] * bi;
         v[c] += bi * bi;
      }
   block_sum<3>(v, sh);
   if (threadIdx.x == 0)
      for (int c = 0; c < 3; c++) atomicAdd(acc + c, v[c]);
}
__global__ void k_dbicg_begin(DBicgState* st, double* acc, double tol, int maxit) {
   st->maxit = maxit;
   int any = 0;
   for (int c = 0; c < 3; c++) {
      st->bb[c] = st->rr[c] = st->rrs[c] = acc[c];
      st->thresh[c] = tol * sqrt(acc[c]);
      st->its[c] = 0; st->half[c] = 0; st->act2[c] = 0; st->alpha[c] = st->omega[c] = 0.0;
      st->act[c] = (sqrt(st->rr[c]) <= st->thresh[c] || maxit <= 0) ? 0 : 1;
      any |= st->act[c];
   }
   st->done = any ? 0 : 1;
   for (int i = 0; i < 24; i++) acc[i] = 0.0;
}
// out_c = A in_c for systems flagged in `flags` (act or act2); acc[o0+c] += l1_c . out_c ; acc[o1+c] += out_c . out_c
__global__ void __launch_bounds__(256) k_dbicg_spmv(Sell A, int vidx, long long ld, const DBicgState* __restrict__ st, int useAct2,
                                                    const double* __restrict__ in, double* __restrict__ out, const double* __restrict__ l1,
                                                    double* __restrict__ acc, int o0, int o1) {
   if (st->done) return;
   __shared__ double sh[6 * 32];
   const int* flags = useAct2 ? st->act2 : st->act;
   const bool f0 = flags[0], f1 = flags[1], f2 = flags[2];
   RowMap m;
   double v[6] = {0, 0, 0, 0, 0, 0};
   if (row_map(A, m)) {
      const double* __restrict__ val = sell_vals(A, vidx);
      double a0 = 0, a1 = 0, a2 = 0;
      sell_pairs<2>(A, m.slice, m.lane, [&](long long e, int2 c2) {
         const double2 v2 = ldnc_d2(val + e);
         if (f0) a0 += v2.x * __ldg(in + c2.x) + v2.y * __ldg(in + c2.y);
         if (f1) a1 += v2.x * __ldg(in + ld + c2.x) + v2.y * __ldg(in + ld + c2.y);
         if (f2) a2 += v2.x * __ldg(in + 2 * ld + c2.x) + v2.y * __ldg(in + 2 * ld + c2.y);
      });
      a0 = lanes_sum(a0, A.T); a1 = lanes_sum(a1, A.T); a2 = lanes_sum(a2, A.T);
      if (m.owner) {
         const double a[3] = {a0, a1, a2};
#pragma unroll
         for (int c = 0; c < 3; c++)
            if (flags[c]) {
               const long long i = c * ld + m.row;
               out[i] = a[c];
               v[c] = l1[i] * a[c];
               v[3 + c] = a[c] * a[c];
            }
      }
   }
   block_sum<6>(v, sh);
   if (threadIdx.x == 0)
      for (int c = 0; c < 3; c++) { atomicAdd(acc + o0 + c, v[c]); atomicAdd(acc + o1 + c, v[3 + c]); }
}
// s = r - alpha AMp (kept in r), sh = M s, |s|^2
__global__ void k_dbicg_s(int n, long long ld, const DBicgState* __restrict__ st, const double* __restrict__ AMp, const double* __restrict__ dinv,
                          double* __restrict__ r, double* __restrict__ shv, double* __restrict__ acc) {
   if (st->done) return;
   __shared__ double sh[3 * 32];
   double v[3] = {0, 0, 0};
   for (int c = 0; c < 3; c++) {
      if (!st->act[c]) continue;
      const double alpha = st->rrs[c] / acc[c];
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
         const long long k = c * ld + i;
         const double si = r[k] - alpha * AMp[k];
         r[k] = si;
         shv[k] = dinv[i] * si;
         v[c] += si * si;
      }
   }
   block_sum<3>(v, sh);
   if (threadIdx.x == 0)
      for (int c = 0; c < 3; c++) atomicAdd(acc + 6 + c, v[c]);
}
__global__ void k_dbicg_mid(DBicgState* st, const double* acc) {
   if (st->done) return;
   for (int c = 0; c < 3; c++) {
      st->alpha[c] = st->act[c] ? st->rrs[c] / acc[c] : 0.0;
      st->half[c] = (st->act[c] && sqrt(acc[6 + c]) <= st->thresh[c]) ? 1 : 0;
      st->act2[c] = (st->act[c] && !st->half[c]) ? 1 : 0;
   }
}
// x += alpha M p (+ omega M s) ; r = s - omega AMs ; rs.r , r.r
__global__ void k_dbicg_x(int n, long long ld, const DBicgState* __restrict__ st, const double* __restrict__ ph, const double* __restrict__ shv,
                          const double* __restrict__ AMs, const double* __restrict__ rs, double* __restrict__ x, double* __restrict__ r,
                          double* __restrict__ acc) {
   if (st->done) return;
   __shared__ double sh[6 * 32];
   double v[6] = {0, 0, 0, 0, 0, 0};
   for (int c = 0; c < 3; c++) {
      const double alpha = st->alpha[c];
      if (st->half[c]) {
         for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[c * ld + i] += alpha * ph[c * ld + i];
      } else if (st->act2[c]) {
         const double omega = acc[12 + c] / acc[15 + c];
         for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const long long k = c * ld + i;
            x[k] += alpha * ph[k] + omega * shv[k];
            const double ri = r[k] - omega * AMs[k];
            r[k] = ri;
            v[c] += rs[k] * ri;
            v[3 + c] += ri * ri;
         }
      }
   }
   block_sum<6>(v, sh);
   if (threadIdx.x == 0)
      for (int c = 0; c < 6; c++) atomicAdd(acc + 18 + c, v[c]);
}
// p = r + beta (p - omega AMp) ; ph = M p
__global__ void k_dbicg_p(int n, long long ld, const DBicgState* __restrict__ st, const double* __restrict__ r, const double* __restrict__ AMp,
                          const double* __restrict__ dinv, double* __restrict__ p, double* __restrict__ ph, const double* __restrict__ acc) {
   if (st->done) return;
   for (int c = 0; c < 3; c++) {
      if (!st->act2[c]) continue;
      const double omega = acc[12 + c] / acc[15 + c];
      const double beta = (acc[18 + c] / st->rrs[c]) * (st->alpha[c] / omega);
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
         const long long k = c * ld + i;
         const double pi = r[k] + beta * (p[k] - omega * AMp[k]);
         p[k] = pi;
         ph[k] = dinv[i] * pi;
      }
   }
}
__global__ void k_dbicg_next(DBicgState* st, double* acc) {
   if (st->done) return;
   int any = 0;
   for (int c = 0; c < 3; c++) {
      if (st->half[c]) { st->its[c] += 1; st->act[c] = 0; }
      else if (st->act2[c]) {
         st->rrs[c] = acc[18 + c]; st->rr[c] = acc[21 + c]; st->its[c] += 1;
         if (sqrt(st->rr[c]) <= st->thresh[c] || st->its[c] >= st->maxit) st->act[c] = 0;
      }
      st->half[c] = 0; st->act2[c] = 0;
      any |= st->act[c];
   }
   st->done = any ? 0 : 1;
   for (int i = 0; i < 24; i++) acc[i] = 0.0;
}

// max over owned entries of |a - b| for a 3-component vector with stride ld
__global__ void __launch_bounds__(256) k_max_abs_diff3(int n, long long ld, const double* __restrict__ a, const double* __restrict__ b,
                                                       double* __restrict__ out) {
   __shared__ double sh[32];
   double mx = 0.0;
   for (int c = 0; c < 3; c++)
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) mx = fmax(mx, fabs(a[c * ld + i] - b[c * ld + i]));
   mx = warp_max(mx);
   if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
   __syncthreads();
   if (threadIdx.x < 32) {
      mx = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
      mx = warp_max(mx);
      if (threadIdx.x == 0) atomic_max_nonneg(out, mx);
   }
}


// =================================================================================================
// Peer-memory path: the pressure solve as ONE persistent cooperative kernel per GPU.  Interface values of p are
// stored straight into the neighbours' ghost entries over NVLink, dot products are all-reduced through mailboxes in
// the peers' windows; ordering by system-scope release/acquire flags carrying monotonically increasing epochs.
// =================================================================================================
// window header: 4 KiB of flags, then 3 (slots) x 64 (ranks) mailboxes of 16 words (128 B: up to 8 doubles as flag-in-data word pairs)
enum { kWinFlagsBytes = 4096, kWinMailWords = 16, kWinMailBytes = 3 * 64 * 16 * 8, kWinHeader = 32768, kMaxRanks = 64 };

struct XComm {
   int world, rank, npeers;
   char* const* peerWin;               // [world] window bases (own included)
   const int* peerRank;                // [npeers]
   double* const* peerP;               // [npeers] start of my block inside each peer's ghost range of the gathered pressure vector
   const int* pushPtr;                 // [n+1]
   const int2* pushDst;                // (peer index, slot)
   const int* ifaceRows;               // [nIface] owned pressure DOFs that some peer holds as a ghost
   int nIface;
   unsigned long long* xstate;         // [0] allreduce epoch [1] halo epoch [2] error [3] reduction ticket [4] interface ticket
   unsigned long long timeoutNs;       // spin-waits give up after this long (a peer died): error word set, every block leaves the kernel
};
__device__ __forceinline__ unsigned long long* win_arflag(char* w, int slot, int src) { return (unsigned long long*)w + slot * kMaxRanks + src; }
__device__ __forceinline__ unsigned long long* win_haloflag(char* w, int src) { return (unsigned long long*)w + 3 * kMaxRanks + src; }
__device__ __forceinline__ double* win_mail(char* w, int slot, int src) { return (double*)(w + kWinFlagsBytes) + (slot * kMaxRanks + src) * kWinMailWords; }

__device__ __forceinline__ void st_rel_sys(unsigned long long* p, unsigned long long v) {
   asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acq_sys(const unsigned long long* p) {
   unsigned long long v;
   asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
   return v;
}
__device__ __forceinline__ void st_sys_f64(double* p, double v) { asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ double ld_sys_f64(const double* p) {
   double v;
   asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
   return v;
}
__device__ __forceinline__ unsigned long long gtimer() {
   unsigned long long t;
   asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
   return t;
}
// Spin until *f >= want; false after the time-out (a peer died) or when another block already raised the error word.
// (err[0] = error word; on the first time-out err[3..5] = what was waited for: 1 = halo flag | wanted epoch | value seen)
__device__ __forceinline__ bool wait_flag(const unsigned long long* f, unsigned long long want, unsigned long long* err, unsigned long long timeoutNs) {
   const unsigned long long t0 = gtimer();
   unsigned spins = 0;
   unsigned long long seen;
   while ((seen = ld_acq_sys(f)) < want) {
      if ((++spins & 255u) == 0) {
         if (*(volatile unsigned long long*)err) return false;
         if (gtimer() - t0 > timeoutNs) {
            if (atomicCAS(err, 0ull, 1ull) == 0ull) { err[3] = 1; err[4] = want; err[5] = seen; }
            return false;
         }
      }
   }
   return true;
}

__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
   asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
   unsigned long long v;
   asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
   return v;
}

// Sum N (<= 8) values over all blocks of all ranks; the result is bit-identical in every block of every rank
// (local partials in a fixed order, then ranks in rank order), so every block takes the same branches.
// One fused barrier instead of "local grid reduction, then exchange":
//   * every block publishes its partial and takes a ticket; the LAST block to arrive reduces the partials and
//     delivers this rank's sums to every rank's mailbox (its own included).  Tickets only ever count up (64 bit): the last
//     block of reduction k is the one that draws k*G - 1, so there is no reset that a fast block could overtake
//   * delivery uses flag-in-data words (low / high half of the double + the 32-bit epoch in one 8-byte store, the
//     scheme of NCCL's LL protocol): no separate flag, one NVLink traversal.  The delivering lanes fence at system scope
//     first, so that everything this rank wrote before the reduction (peer-memory pushes included) is visible to whoever
//     sees the words
//   * every block polls its own mailbox; seeing all ranks' words of this epoch implies that every block of every
//     rank has finished the phase before the reduction -> this IS the grid + cross-GPU barrier.  Every polling lane
//     fences after its own poll (acquire side).
// Returns false -- in every thread of the block -- after a time-out: the caller must leave the kernel at once (no further
// grid-wide or cross-GPU synchronisation; every block of every rank runs into the same time-out or sees the error word).
// PUBLISH = true: peer-memory stores made before the call (each storing thread fenced them at system scope) are handed over with
// the reduction: the delivering warp fences at system scope before it writes the words.  System-scope fences are expensive
// (measured: one in every polling block cost ~15 us per reduction), so nothing else in here uses one.
template <int N, bool PUBLISH = false>
__device__ __forceinline__ bool xgrid_sum(double (&v)[N], double* buf, double* sh, const XComm& xc, unsigned long long& epoch) {
   static_assert(2 * N <= kWinMailWords, "mailbox too small for this many values");
   __shared__ int s_last, s_fail;
   block_sum<N>(v, sh);
   epoch++;
   const unsigned long long tag = (epoch & 0xffffffffull) << 32;
   const int slot = (int)(epoch % 3);
   const unsigned long long G = gridDim.x;
   unsigned long long* ticket = xc.xstate + 3;
   if (threadIdx.x == 0) {
#pragma unroll
      for (int i = 0; i < N; i++) buf[(size_t)i * G + blockIdx.x] = v[i];
      __threadfence();
      s_last = (atomicAdd(ticket, 1ull) % G == G - 1) ? 1 : 0;
      s_fail = 0;
   }
   __syncthreads();
   if (s_last && threadIdx.x < 32) {
      __threadfence();   // acquire side of the tickets: the other blocks' partials (and whatever they published before them)
      if (PUBLISH) __threadfence_system();
      double tot[N];
#pragma unroll
      for (int i = 0; i < N; i++) {
         double s = 0.0;
         for (int b = threadIdx.x; b < (int)G; b += 32) s += __ldcg(buf + (size_t)i * G + b);
         tot[i] = warp_sum(s);
      }
      if (threadIdx.x < xc.world) {   // lane q delivers to rank q
         unsigned long long* mb = (unsigned long long*)win_mail(xc.peerWin[threadIdx.x], slot, xc.rank);
#pragma unroll
         for (int i = 0; i < N; i++) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(tot[i]);
            st_sys_u64(mb + 2 * i, tag | (bits & 0xffffffffull));
            st_sys_u64(mb + 2 * i + 1, tag | (bits >> 32));
         }
      }
   }
   if (threadIdx.x < 32) {
      double mine[N];
#pragma unroll
      for (int i = 0; i < N; i++) mine[i] = 0.0;
      if (threadIdx.x < xc.world) {
         const unsigned long long* mb = (const unsigned long long*)win_mail(xc.peerWin[xc.rank], slot, threadIdx.x);
         const unsigned long long t0 = gtimer();
         unsigned spins = 0;
         bool ok = true;
#pragma unroll
         for (int i = 0; i < N; i++) {
            unsigned long long lo = 0, hi = 0;
            while (ok) {
               lo = ld_sys_u64(mb + 2 * i);
               hi = ld_sys_u64(mb + 2 * i + 1);
               if ((lo >> 32) == (tag >> 32) && (hi >> 32) == (tag >> 32)) break;
               if ((++spins & 255u) == 0) {
                  if (*(volatile unsigned long long*)(xc.xstate + 2)) ok = false;
                  else if (gtimer() - t0 > xc.timeoutNs) {
                     if (atomicCAS(xc.xstate + 2, 0ull, 1ull) == 0ull) { xc.xstate[5] = 2 + 16 * (unsigned long long)(N * 64 + threadIdx.x); xc.xstate[6] = epoch; xc.xstate[7] = lo; }
                     ok = false;
                  }
               }
            }
            mine[i] = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
         }
         if (!ok) s_fail = 1;
         __threadfence();   // acquire side, in every lane that observed a word (drops this SM's stale L1 lines before the block reads on)
      }
#pragma unroll
      for (int i = 0; i < N; i++) {
         double s = 0.0;
         for (int q = 0; q < xc.world; q++) s += __shfl_sync(0xffffffffu, mine[i], q);   // rank order
         if (threadIdx.x == 0) sh[i] = s;
      }
   }
   __syncthreads();
#pragma unroll
   for (int i = 0; i < N; i++) v[i] = sh[i];
   const bool ok = s_fail == 0;
   __syncthreads();
   return ok;
}

// p <- values (owned rows), pushed to the neighbours' ghost entries; then signal and wait for their pushes.
// Returns false (block-uniform) after a time-out: leave the kernel.
template <typename F>
__device__ __forceinline__ bool p_update_and_push(cg::grid_group& grid, int n, double* p, const XComm& xc, unsigned long long& hepoch,
                                                  int tid, int nthreads, F value) {
   bool pushed = false;
   for (int i = tid; i < n; i += nthreads) {
      const double pn = value(i);
      p[i] = pn;
      for (int e = xc.pushPtr[i]; e < xc.pushPtr[i + 1]; e++) {
         const int2 d = xc.pushDst[e];
         st_sys_f64(xc.peerP[d.x] + d.y, pn);
         pushed = true;
      }
   }
   if (pushed) __threadfence_system();   // only the threads that stored into a neighbour's window pay for the system-scope fence
   grid.sync();
   hepoch++;
   if (blockIdx.x == 0 && threadIdx.x < xc.npeers) st_rel_sys(win_haloflag(xc.peerWin[xc.peerRank[threadIdx.x]], xc.rank), hepoch);
   bool ok = true;
   if (threadIdx.x < xc.npeers) ok = wait_flag(win_haloflag(xc.peerWin[xc.rank], xc.peerRank[threadIdx.x]), hepoch, xc.xstate + 2, xc.timeoutNs);
   return __syncthreads_and(ok ? 1 : 0) != 0;
}
