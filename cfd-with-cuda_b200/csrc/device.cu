// Device side of the fractional-step time loop: layout conversion, kernels, seam functions.
// Reference seam being replaced:
//   explicit  fractionalStep/explicit/Cpp/CUDAcodes.cu            (BCH.cu below)
//   implicit  fractionalStep/implicit/Cpp/CUDAcodes.cu            (GQ.cu below)
// and their call sites in blascoCodinaHuerta.cpp (BCH.cpp) / guermondQuartapelle.cpp (GQ.cpp).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <memory>

#include <dlfcn.h>

#include "kernels.cuh"

namespace cfdb {

static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
   char buf[1024];
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(buf, sizeof buf, fmt, ap);
   va_end(ap);
   g_err = buf;
}
const char* last_error() { return g_err.c_str(); }

// =================================================================================================
// Sell: host conversion + upload
// =================================================================================================
int choose_T(int nrows, long long nnz) {
   if (const char* e = getenv("CFD_Z_T")) {   // tuning override (experiments only)
      const int t = atoi(e);
      if (t == 1 || t == 2 || t == 4 || t == 8 || t == 16 || t == 32) return t;
   }
   if (nrows <= 0) return 1;
   const double avg = (double)nnz / nrows;
   int Tmax = 1;   // keep at least ~8 entries per lane
   while (Tmax < 8 && avg / (4 * Tmax) >= 4.0) Tmax *= 2;
   // mid-size and large matrices: >= 4 slices per resident warp of the persistent kernels (blocked + dynamic-tail
   // scheduling needs that); small L2-resident ones: about one slice per warp at 2 CTAs/SM -- more lanes would only
   // add CTAs to every grid-wide sync, which is what bounds those solves
   const long long big = 909000, small = 100000;
   int T = 1;
   if ((long long)nrows * Tmax >= big) { while ((long long)nrows * T < big) T *= 2; }
   else { while ((long long)nrows * T < small && T < Tmax) T *= 2; }
   return T;
}

// lanes per row for the velocity-row matrices: 1 once there are enough rows to fill the machine
int choose_Tv(int nrows, long long nnz) {
   if (const char* e = getenv("CFD_V_T")) {   // tuning override (experiments only)
      const int t = atoi(e);
      if (t == 1 || t == 2 || t == 4 || t == 8) return t;
   }
   if (nrows <= 0) return 1;
   const double avg = (double)nnz / nrows;
   int T = 1;
   while ((long long)nrows * T < 400000 && T < 4 && avg / (2 * T) >= 8.0) T *= 2;
   return T;
}

void sell_build_host(int nrows, int ncols, const int* rs, const int* col, int T, SellHost& out) {
   const int R = 32 / T;
   const int ns = (nrows + R - 1) / R;
   Sell& m = out.meta;
   m.nrows = nrows; m.ncols = ncols; m.T = T; m.R = R; m.nslices = ns; m.nnz = rs[nrows];
   out.sliceOff.assign((size_t)ns + 1, 0);
   out.colOff.assign((size_t)ns, 0);
   std::vector<int> width(ns, 0);
   std::vector<unsigned char> comp(ns, 0);
   std::vector<std::vector<int> > uni(ns);   // offset list of the compressed slices
   const bool allowComp = getenv("CFD_NO_COLCOMP") == nullptr;
#pragma omp parallel
   {
      std::vector<int> u;
#pragma omp for schedule(static)
      for (int s = 0; s < ns; s++) {
         int w = 0;
         for (int rl = 0; rl < R; rl++) {
            const int r = s * R + rl;
            if (r < nrows) w = std::max(w, (rs[r + 1] - rs[r] + T - 1) / T);
         }
         width[s] = (w + 1) & ~1;
         if (!allowComp || w == 0) continue;
         // union of the (col - row) offsets over the slice's rows
         u.clear();
         for (int rl = 0; rl < R; rl++) {
            const int r = s * R + rl;
            if (r >= nrows) break;
            for (int k = rs[r]; k < rs[r + 1]; k++) u.push_back(col[k] - r);
         }
         std::sort(u.begin(), u.end());
         u.erase(std::unique(u.begin(), u.end()), u.end());
         // offset u of the sorted union goes to lane u % T, slot u / T (the same dealing as the explicit layout)
         const int wc = (((int)u.size() + T - 1) / T + 1) & ~1;
         // worth it while the extra zero values (8 B each) cost less than the saved indices (4 B each)
         if (4 * wc <= 5 * width[s]) { comp[s] = 1; width[s] = wc; uni[s] = u; }
      }
   }
   long long cpos = 0;
   for (int s = 0; s < ns; s++) {
      out.sliceOff[s + 1] = out.sliceOff[s] + (long long)width[s] * 32;
      out.colOff[s] = cpos | (comp[s] ? 1 : 0);
      cpos += comp[s] ? (long long)width[s] * T : (long long)width[s] * 32;
      out.compressedSlices += comp[s];
   }
   m.padded = out.sliceOff[ns];
   m.anyComp = out.compressedSlices > 0 ? 1 : 0;
   m.nindex = cpos;
   out.cols.assign((size_t)std::max<long long>(cpos, 2), 0);
   out.perm.assign((size_t)m.nnz, 0);
#pragma omp parallel for schedule(static)
   for (int s = 0; s < ns; s++) {
      const long long off = out.sliceOff[s];
      const long long co = out.colOff[s] & ~1LL;
      const int w = width[s];
      if (comp[s]) {
         const std::vector<int>& u = uni[s];
         // offset list stored as [slot j][lane-in-row t]; missing tail = offset 0 (the row itself, value 0)
         for (int q = 0; q < w * T; q++) out.cols[(size_t)co + q] = q < (int)u.size() ? u[q] : 0;
         for (int rl = 0; rl < R; rl++) {
            const int r = s * R + rl;
            if (r >= nrows) break;
            for (int k = rs[r]; k < rs[r + 1]; k++) {
               const int q = (int)(std::lower_bound(u.begin(), u.end(), col[k] - r) - u.begin());
               const int j = q / T, t = rl * T + q % T;
               out.perm[(size_t)k] = off + (long long)(j >> 1) * 64 + 2 * t + (j & 1);
            }
         }
         continue;
      }
      for (int rl = 0; rl < R; rl++) {
         const int r = s * R + rl;
         const int len = (r < nrows) ? rs[r + 1] - rs[r] : 0;
         const int padcol = len > 0 ? col[rs[r]] : 0;
         for (int lane = 0; lane < T; lane++) {
            const int t = rl * T + lane;
            for (int j = 0; j < w; j++) {
               const long long rel = (long long)(j >> 1) * 64 + 2 * t + (j & 1);
               const int k = j * T + lane;
               if (k < len) {
                  out.cols[(size_t)(co + rel)] = col[rs[r] + k];
                  out.perm[(size_t)rs[r] + k] = off + rel;
               } else {
                  out.cols[(size_t)(co + rel)] = padcol;
               }
            }
         }
      }
   }
}

int sell_upload(const SellHost& h, int nval, const double* const* csrVals, Sell& d) {
   d = h.meta;
   d.nval = nval;
   CFD_CK(cudaMalloc(&d.sliceOff, sizeof(long long) * h.sliceOff.size()));
   CFD_CK(cudaMemcpy(d.sliceOff, h.sliceOff.data(), sizeof(long long) * h.sliceOff.size(), cudaMemcpyHostToDevice));
   CFD_CK(cudaMalloc(&d.colOff, sizeof(long long) * std::max<size_t>(h.colOff.size(), 1)));
   CFD_CK(cudaMemcpy(d.colOff, h.colOff.data(), sizeof(long long) * h.colOff.size(), cudaMemcpyHostToDevice));
   CFD_CK(cudaMalloc(&d.cols, sizeof(int) * std::max<size_t>(h.cols.size(), 2)));
   CFD_CK(cudaMemcpy(d.cols, h.cols.data(), sizeof(int) * h.cols.size(), cudaMemcpyHostToDevice));
   std::unique_ptr<double[]> tmp;   // staging in layout order: allocated once, zero-filled and scattered in parallel (multi-GB arrays)
   for (int v = 0; v < nval; v++) {
      CFD_CK(cudaMalloc(&d.vals[v], sizeof(double) * std::max<size_t>((size_t)d.padded, 2)));
      if (csrVals && csrVals[v]) {
         if (!tmp) tmp.reset(new double[std::max<size_t>((size_t)d.padded, 2)]);
         double* t = tmp.get();
         const double* src = csrVals[v];
#pragma omp parallel for schedule(static)
         for (long long k = 0; k < d.padded; k++) t[k] = 0.0;
#pragma omp parallel for schedule(static)
         for (long long k = 0; k < d.nnz; k++) t[(size_t)h.perm[(size_t)k]] = src[k];
         CFD_CK(cudaMemcpy(d.vals[v], t, sizeof(double) * (size_t)d.padded, cudaMemcpyHostToDevice));
      } else {
         CFD_CK(cudaMemset(d.vals[v], 0, sizeof(double) * (size_t)d.padded));
      }
   }
   return CFD_OK;
}

void sell_free(Sell& s) {
   cudaFree(s.sliceOff); cudaFree(s.colOff); cudaFree(s.cols);
   for (int v = 0; v < 3; v++) cudaFree(s.vals[v]);
   s = Sell();
}

// =================================================================================================
// Reference-element tables in constant memory
// =================================================================================================
// Both element families are resident so that Q2/Q1 and Q1/Q1 contexts can coexist in one process.
__constant__ double cSv27[8][27];       // [k][i]
__constant__ double cdSv27[3][27][8];   // [d][i][k]
__constant__ double cSv8[8][8];
__constant__ double cdSv8[3][8][8];
__constant__ double cdSp[3][8][8];      // [d][i][k] trilinear (geometry + pressure)
template <int NV> struct Tab;
template <> struct Tab<27> {
   static __device__ __forceinline__ double Sv(int k, int i) { return cSv27[k][i]; }
   static __device__ __forceinline__ double dSv(int d, int i, int k) { return cdSv27[d][i][k]; }
};
template <> struct Tab<8> {
   static __device__ __forceinline__ double Sv(int k, int i) { return cSv8[k][i]; }
   static __device__ __forceinline__ double dSv(int d, int i, int k) { return cdSv8[d][i][k]; }
};

// =================================================================================================
// Row kernels (one thread per velocity row, T = 1 matrices)
// =================================================================================================

// Thread -> (slice, lane, row, sub-lane) map shared by the row kernels: one warp per slice, T lanes per row.
struct RowMap {
   int slice, lane, row, sub;
   bool owner;   // this lane writes the row's result
};
__device__ __forceinline__ bool row_map(const Sell& A, RowMap& m) {
   const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
   m.slice = (int)(g >> 5);
   m.lane = (int)(g & 31);
   if (m.slice >= A.nslices) { m.owner = false; return false; }   // warp-uniform
   m.row = m.slice * A.R + m.lane / A.T;
   m.sub = m.lane % A.T;
   m.owner = (m.sub == 0) && (m.row < A.nrows);
   return true;
}
__device__ __forceinline__ double lanes_sum(double v, int T) {
   for (int o = T >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}
__device__ __forceinline__ const double* sell_vals(const Sell& A, int vidx) {
   return vidx == 0 ? A.vals[0] : (vidx == 1 ? A.vals[1] : A.vals[2]);   // no dynamic indexing of the by-value struct
}

// Visit the (value index, column pair) of every entry pair this lane owns in `slice`:  f(e, c2) with e the index
// of the pair inside the value arrays.  Three specialised loops: matrices without compressed slices (columns laid
// out exactly like the values, no extra loads), offset-compressed slices (one broadcast 8-byte load per pair, added
// to the row and clamped: padded zero entries near the ends of the matrix may point outside), explicit slices of a
// matrix that also has compressed ones.
template <int UNROLL, typename F>
__device__ __forceinline__ void sell_pairs(const Sell& A, int slice, int lane, F&& f) {
   const long long off = A.sliceOff[slice];
   const int np = (int)((A.sliceOff[slice + 1] - off) >> 6);
   const long long e0 = off + 2 * lane;
   if (!A.anyComp) {
      const int* __restrict__ cp = A.cols + e0;
#pragma unroll(UNROLL)
      for (int jp = 0; jp < np; jp++) f(e0 + ((long long)jp << 6), ldnc_i2(cp + ((long long)jp << 6)));
      return;
   }
   const long long co = A.colOff[slice];
   if (co & 1) {
      const int row = slice * A.R + lane / A.T, hi = A.ncols - 1;
      if (A.T == 1) {
         const int2* __restrict__ op = reinterpret_cast<const int2*>(A.cols + (co & ~1LL));
#pragma unroll(UNROLL)
         for (int jp = 0; jp < np; jp++) {
            const int2 o = __ldg(op + jp);
            f(e0 + ((long long)jp << 6), make_int2(min(max(row + o.x, 0), hi), min(max(row + o.y, 0), hi)));
         }
      } else {   // T lanes per row: slot j of lane-in-row t holds offset [j][t]
         const int T = A.T;
         const int* __restrict__ op = A.cols + (co & ~1LL) + lane % T;
#pragma unroll(UNROLL)
         for (int jp = 0; jp < np; jp++) {
            const int ox = __ldg(op + (2 * jp) * T), oy = __ldg(op + (2 * jp + 1) * T);
            f(e0 + ((long long)jp << 6), make_int2(min(max(row + ox, 0), hi), min(max(row + oy, 0), hi)));
         }
      }
   } else {
      const int* __restrict__ cp = A.cols + co + 2 * lane;
#pragma unroll(UNROLL)
      for (int jp = 0; jp < np; jp++) f(e0 + ((long long)jp << 6), ldnc_i2(cp + ((long long)jp << 6)));
   }
}

// y_c = alpha * A x_c + beta * y_c for NV vectors sharing one Sell value array (vectors are
// component-blocked with stride ldx / ldy).  Multi-RHS: the matrix stream is read once.
template <int NV>
__global__ void __launch_bounds__(256) k_spmv_multi(Sell A, int vidx, const double* __restrict__ x, long long ldx,
                                                    double* __restrict__ y, long long ldy, double alpha, double beta) {
   RowMap m;
   if (!row_map(A, m)) return;
   const double* __restrict__ val = sell_vals(A, vidx);
   double acc[NV];
#pragma unroll
   for (int c = 0; c < NV; c++) acc[c] = 0.0;
   sell_pairs<4>(A, m.slice, m.lane, [&](long long e, int2 c2) {
      const double2 v2 = ldnc_d2(val + e);
#pragma unroll
      for (int c = 0; c < NV; c++) acc[c] += v2.x * __ldg(x + c * ldx + c2.x) + v2.y * __ldg(x + c * ldx + c2.y);
   });
#pragma unroll
   for (int c = 0; c < NV; c++) acc[c] = lanes_sum(acc[c], A.T);
   if (m.owner) {
#pragma unroll
      for (int c = 0; c < NV; c++) {
         double r = alpha * acc[c];
         if (beta != 0.0) r += beta * y[c * ldy + m.row];
         y[c * ldy + m.row] = r;
      }
   }
}

// Accumulate a0..a2 += K * U_c (one value array, three component vectors) over this lane's entries.
__device__ __forceinline__ void lane_dot_k3(const Sell& K, const double* __restrict__ val, const RowMap& m, const double* __restrict__ U,
                                            long long NN, double& a0, double& a1, double& a2) {
   sell_pairs<4>(K, m.slice, m.lane, [&](long long e, int2 c2) {
      const double2 v2 = ldnc_d2(val + e);
      a0 += v2.x * __ldg(U + c2.x) + v2.y * __ldg(U + c2.y);
      a1 += v2.x * __ldg(U + NN + c2.x) + v2.y * __ldg(U + NN + c2.y);
      a2 += v2.x * __ldg(U + 2 * NN + c2.x) + v2.y * __ldg(U + 2 * NN + c2.y);
   });
}
// Accumulate a_c += sign * G_c * p (three value arrays, one vector); p(col) is supplied by a functor so
// that callers can form it on the fly (e.g. 2 Pk_prev - Pk_prevprev).
template <typename PF>
__device__ __forceinline__ void lane_dot_g3(const Sell& G, const RowMap& m, PF pf, double sign, double& a0, double& a1, double& a2) {
   sell_pairs<2>(G, m.slice, m.lane, [&](long long e, int2 c2) {
      const double2 g1 = ldnc_d2(G.vals[0] + e), g2 = ldnc_d2(G.vals[1] + e), g3 = ldnc_d2(G.vals[2] + e);
      const double px = sign * pf(c2.x), py = sign * pf(c2.y);
      a0 += g1.x * px + g1.y * py;
      a1 += g2.x * px + g2.y * py;
      a2 += g3.x * px + g3.y * py;
   });
}

// step 1 of the explicit scheme, fused (replaces step1GPUpart, BCH.cu:946-973: 6 csrmv + applyVelBC +
// calculate_UnpHalf):   R1_c = R1_c - K*Uprev_c - G_c*Pn ;  R1[bc] = 0 ;  UnpHalf = Un + dt*R1*MdInv
// K and G share the row partition (same T), so one thread owns the same row in both.
__global__ void __launch_bounds__(256) k_step1(Sell K, Sell G, int NN, const double* __restrict__ Uprev,
                                               const double* __restrict__ Pn, const double* __restrict__ Un,
                                               const double* __restrict__ MdInv, const int* __restrict__ bcIndex,
                                               double* __restrict__ R1, double* __restrict__ UnpHalf, double dt) {
   RowMap m;
   if (!row_map(K, m)) return;
   double a0 = 0, a1 = 0, a2 = 0;
   lane_dot_k3(K, K.vals[0], m, Uprev, NN, a0, a1, a2);
   lane_dot_g3(G, m, [Pn](int c) { return __ldg(Pn + c); }, 1.0, a0, a1, a2);
   a0 = lanes_sum(a0, K.T); a1 = lanes_sum(a1, K.T); a2 = lanes_sum(a2, K.T);
   if (m.owner) {
      const bool bc = bcIndex[m.row] >= 0;
      const double acc[3] = {a0, a1, a2};
#pragma unroll
      for (int c = 0; c < 3; c++) {
         const long long i = (long long)c * NN + m.row;
         const double r = bc ? 0.0 : (R1[i] - acc[c]);
         R1[i] = r;
         UnpHalf[i] = Un[i] + dt * r * MdInv[i];
      }
   }
}

// step 3 of the explicit scheme, fused (replaces step3GPU, BCH.cu:1029-1060: 3 csrmv + calculate_R3 +
// applyVelBC + multiplyVectors + calculate_Unp1) plus the velocity half of checkConvergenceGPU's norms:
//   R3_c = -dt*G_c*Pdot - [iter>1] dt*KAcc_c ; R3[bc]=0 ; Acc = R3*MdInv ; Unp1 = UnpHalf + dt*Acc
//   sums[0] += |Unp1|^2 ; sums[1] += |Unp1 - Unp1_prev|^2
__global__ void __launch_bounds__(256) k_step3(Sell G, int NN, const double* __restrict__ Pdot,
                                               const double* __restrict__ KAcc, int useKAcc,
                                               const double* __restrict__ MdInv, const int* __restrict__ bcIndex,
                                               const double* __restrict__ UnpHalf, const double* __restrict__ Unp1_prev,
                                               double* __restrict__ R3, double* __restrict__ Acc,
                                               double* __restrict__ Unp1, double dt, double* __restrict__ sums) {
   __shared__ double sh[64];
   RowMap m;
   double s[2] = {0.0, 0.0};
   if (row_map(G, m)) {
      double a0 = 0, a1 = 0, a2 = 0;
      lane_dot_g3(G, m, [Pdot](int c) { return __ldg(Pdot + c); }, 1.0, a0, a1, a2);
      a0 = lanes_sum(a0, G.T); a1 = lanes_sum(a1, G.T); a2 = lanes_sum(a2, G.T);
      if (m.owner) {
         const bool bc = bcIndex[m.row] >= 0;
         const double acc[3] = {a0, a1, a2};
#pragma unroll
         for (int c = 0; c < 3; c++) {
            const long long i = (long long)c * NN + m.row;
            double r = -dt * acc[c];
            if (useKAcc) r = r - dt * KAcc[i];
            if (bc) r = 0.0;
            const double a = r * MdInv[i];
            const double u = UnpHalf[i] + dt * a;
            R3[i] = r; Acc[i] = a; Unp1[i] = u;
            const double d = u - Unp1_prev[i];
            s[0] += u * u; s[1] += d * d;
         }
      }
   }
   block_sum<2>(s, sh);
   if (threadIdx.x == 0) { atomicAdd(sums + 0, s[0]); atomicAdd(sums + 1, s[1]); }
}

// Pnp1 = Pn + dt*Pdot (calculate_Pnp1, BCH.cu:623-634) fused with the pressure half of the
// convergence norms: sums[2] += |Pnp1|^2 ; sums[3] += |Pnp1 - Pnp1_prev|^2
__global__ void __launch_bounds__(256) k_pressure_update(int n, const double* __restrict__ Pn, const double* __restrict__ Pdot,
                                                         const double* __restrict__ Pprev, double* __restrict__ Pnp1, double dt,
                                                         double* __restrict__ sums) {
   __shared__ double sh[64];
   double s[2] = {0.0, 0.0};
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const double p = Pn[i] + dt * Pdot[i];
      Pnp1[i] = p;
      const double d = p - Pprev[i];
      s[0] += p * p; s[1] += d * d;
   }
   block_sum<2>(s, sh);
   if (threadIdx.x == 0) { atomicAdd(sums + 2, s[0]); atomicAdd(sums + 3, s[1]); }
}

// max |a - b| (checkConvergenceInTimeGPU, BCH.cu:1151-1178 / GQ.cu:975-1002)
__global__ void __launch_bounds__(256) k_max_abs_diff(long long n, const double* __restrict__ a, const double* __restrict__ b,
                                                      double* __restrict__ out) {
   __shared__ double sh[32];
   double m = 0.0;
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      m = fmax(m, fabs(a[i] - b[i]));
   m = warp_max(m);
   if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
   __syncthreads();
   if (threadIdx.x < 32) {
      m = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
      m = warp_max(m);
      if (threadIdx.x == 0) atomic_max_nonneg(out, m);
   }
}

// R2 = sum_c G_c^T d_c on the pressure rows, d formed on the fly (replaces calculate_step2dummyV1/V2 +
// 3 transposed csrmv + applyPresBC of step2GPU, BCH.cu:980-1022; and the implicit variant GQ.cu:897-944):
//   mode 0 (explicit): d = U/dt^2 - [useKAcc] MdOrigInv*KAcc        mode 1 (implicit): d = -U/dt
__global__ void __launch_bounds__(256) k_gt_rhs(Sell GT, int NN, int mode, const double* __restrict__ U,
                                                const double* __restrict__ KAcc, int useKAcc,
                                                const double* __restrict__ MdOrigInv, double scale, int pinNode,
                                                double* __restrict__ R2, double* __restrict__ zero4 = nullptr) {
   const int gthread = blockIdx.x * blockDim.x + threadIdx.x;
   const int slice = gthread >> 5, lane = gthread & 31;
   // the four convergence sums k_pressure_update / k_step3 accumulate later in this inner iteration start from zero (was a memset launch)
   if (zero4 != nullptr && gthread < 4) zero4[gthread] = 0.0;
   if (slice >= GT.nslices) return;
   double acc = 0.0;
   sell_pairs<1>(GT, slice, lane, [&](long long e, int2 c2) {
#pragma unroll
      for (int c = 0; c < 3; c++) {
         const double2 g = ldnc_d2(GT.vals[c] + e);
         const long long ix = (long long)c * NN + c2.x, iy = (long long)c * NN + c2.y;
         double dx = scale * __ldg(U + ix), dy = scale * __ldg(U + iy);
         if (mode == 0 && useKAcc) {
            dx -= __ldg(MdOrigInv + ix) * __ldg(KAcc + ix);
            dy -= __ldg(MdOrigInv + iy) * __ldg(KAcc + iy);
         }
         acc += g.x * dx + g.y * dy;
      }
   });
   for (int o = GT.T >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
   const int row = slice * GT.R + lane / GT.T;
   if ((lane % GT.T) == 0 && row < GT.nrows) R2[row] = (row == pinNode) ? 0.0 : acc;   // pinNode = -1: no pin
}

// =================================================================================================
// Element-by-element convective operator (thread per element, geometry recomputed from the 8 corners)
//   C(u)_ij = sum_k (u_k . grad N_j) N_i w_k detJ_k              (BCH.cpp:3520-3685, BCH.cu:695-899)
// Sum-factorised: with b_k = w_k adj(J_k)^T u_k (adj = detJ * J^-1, so no division is needed)
//   (C(u) x)_i = sum_k N_i(k) * ( b_k . grad_xi x(k) ),   grad_xi x(k) = sum_j dN_j/dxi(k) x_j
// The 27x27 element matrix is never formed in the apply kernel.
// =================================================================================================
template <int NV>
__device__ __forceinline__ void ebe_gauss_b(const int (&nd)[NV], int NN, const double* __restrict__ coord,
                                            const double* __restrict__ Uadv, double (&b)[8][3]) {
   double ug[8][3];
#pragma unroll
   for (int k = 0; k < 8; k++) { ug[k][0] = 0; ug[k][1] = 0; ug[k][2] = 0; }
#pragma unroll
   for (int c = 0; c < 3; c++) {
#pragma unroll
      for (int i = 0; i < NV; i++) {
         const double v = __ldg(Uadv + (long long)c * NN + nd[i]);
#pragma unroll
         for (int k = 0; k < 8; k++) ug[k][c] += Tab<NV>::Sv(k, i) * v;
      }
   }
   double xc[8][3];
#pragma unroll
   for (int m = 0; m < 8; m++) {
#pragma unroll
      for (int d = 0; d < 3; d++) xc[m][d] = __ldg(coord + 3 * (long long)nd[m] + d);
   }
#pragma unroll
   for (int k = 0; k < 8; k++) {
      double J[3][3];
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
         for (int d = 0; d < 3; d++) {
            double s = 0;
#pragma unroll
            for (int m = 0; m < 8; m++) s += cdSp[a][m][k] * xc[m][d];
            J[a][d] = s;
         }
      // adj(J) (cofactor transpose), same expressions as calcJacob's un-normalised inverse (BCH.cpp:2577-2595)
      double A[3][3];
      A[0][0] = J[1][1] * J[2][2] - J[2][1] * J[1][2];
      A[0][1] = -(J[0][1] * J[2][2] - J[0][2] * J[2][1]);
      A[0][2] = J[0][1] * J[1][2] - J[1][1] * J[0][2];
      A[1][0] = -(J[1][0] * J[2][2] - J[1][2] * J[2][0]);
      A[1][1] = J[2][2] * J[0][0] - J[2][0] * J[0][2];
      A[1][2] = -(J[1][2] * J[0][0] - J[1][0] * J[0][2]);
      A[2][0] = J[1][0] * J[2][1] - J[2][0] * J[1][1];
      A[2][1] = -(J[2][1] * J[0][0] - J[2][0] * J[0][1]);
      A[2][2] = J[1][1] * J[0][0] - J[1][0] * J[0][1];
      // b_k[m] = sum_i adj[i][m] u_k[i]   (GQ weight = 1, BCH.cpp:2189-2196)
#pragma unroll
      for (int m = 0; m < 3; m++) b[k][m] = A[0][m] * ug[k][0] + A[1][m] * ug[k][1] + A[2][m] * ug[k][2];
   }
}

// R1 = -C(Uadv) * Ux   (calculateMatrixAGPU, BCH.cu:906-939; R1 must be zeroed first).  Scatter by
// FP64 atomicAdd (RED.ADD.F64) instead of the reference's 8 coloured launches.
// ATOMIC = false: the launch covers one colour (elist = its elements, no two share a node): plain read-modify-write,
// bitwise reproducible -- the reference's scatter scheme with a valid colouring.
template <int NV, bool ATOMIC>
__global__ void __launch_bounds__(128) k_ebe_convect_apply(int NE, int NN, int ldE, const int* __restrict__ LtoGT,
                                                           const double* __restrict__ coord, const double* __restrict__ Uadv,
                                                           const double* __restrict__ Ux, double* __restrict__ R1,
                                                           const int* __restrict__ elist) {
   const int t = blockIdx.x * blockDim.x + threadIdx.x;
   if (t >= NE) return;
   const int e = elist ? elist[t] : t;
   int nd[NV];
#pragma unroll
   for (int i = 0; i < NV; i++) nd[i] = __ldg(LtoGT + (long long)i * ldE + e);
   double b[8][3];
   ebe_gauss_b<NV>(nd, NN, coord, Uadv, b);
#pragma unroll
   for (int c = 0; c < 3; c++) {
      double g[8][3];
#pragma unroll
      for (int k = 0; k < 8; k++) { g[k][0] = 0; g[k][1] = 0; g[k][2] = 0; }
#pragma unroll
      for (int j = 0; j < NV; j++) {
         const double v = __ldg(Ux + (long long)c * NN + nd[j]);
#pragma unroll
         for (int k = 0; k < 8; k++) {
            g[k][0] += Tab<NV>::dSv(0, j, k) * v;
            g[k][1] += Tab<NV>::dSv(1, j, k) * v;
            g[k][2] += Tab<NV>::dSv(2, j, k) * v;
         }
      }
      double t[8];
#pragma unroll
      for (int k = 0; k < 8; k++) t[k] = b[k][0] * g[k][0] + b[k][1] * g[k][1] + b[k][2] * g[k][2];
#pragma unroll
      for (int i = 0; i < NV; i++) {
         double r = 0;
#pragma unroll
         for (int k = 0; k < 8; k++) r += Tab<NV>::Sv(k, i) * t[k];
         if (ATOMIC) atomicAdd(R1 + (long long)c * NN + nd[i], -r);
         else R1[(long long)c * NN + nd[i]] -= r;
      }
   }
}

// A += C(Uadv) assembled into the Sell value array through the element map
// (calculateMatrixAGPU + calcAndAssembleMatrixA, GQ.cu:590-754).
template <int NV, bool ATOMIC>
__global__ void __launch_bounds__(128) k_ebe_convect_assemble(int NE, int NN, int ldE, const int* __restrict__ LtoGT,
                                                              const double* __restrict__ coord, const double* __restrict__ Uadv,
                                                              const int* __restrict__ mapT, double* __restrict__ Aval,
                                                              const int* __restrict__ elist) {
   const int t = blockIdx.x * blockDim.x + threadIdx.x;
   if (t >= NE) return;
   const int e = elist ? elist[t] : t;
   double b[8][3];
   {
      int nd[NV];
#pragma unroll
      for (int i = 0; i < NV; i++) nd[i] = __ldg(LtoGT + (long long)i * ldE + e);
      ebe_gauss_b<NV>(nd, NN, coord, Uadv, b);
   }
   for (int j = 0; j < NV; j++) {
      double cj[8];
#pragma unroll
      for (int k = 0; k < 8; k++) cj[k] = b[k][0] * Tab<NV>::dSv(0, j, k) + b[k][1] * Tab<NV>::dSv(1, j, k) + b[k][2] * Tab<NV>::dSv(2, j, k);
      for (int i = 0; i < NV; i++) {
         double a = 0;
#pragma unroll
         for (int k = 0; k < 8; k++) a += Tab<NV>::Sv(k, i) * cj[k];
         const int pos = __ldg(mapT + (long long)(i * NV + j) * ldE + e);
         if (pos >= 0) {   // -1: row owned by another rank
            if (ATOMIC) atomicAdd(Aval + pos, a);
            else Aval[pos] += a;
         }
      }
   }
}

// =================================================================================================
// Matrix-free (element-by-element) operator path: K, G, G^T and the FE Laplacian applied straight from the
// element integrals of step0 (BCH.cpp:3127-3315, GQ.cpp:3488-3500, 3677-3772) instead of streaming the
// assembled CSR arrays -- cfd_options.operator_path = 1, bench.py --path ebe.
// Thread per element, geometry recomputed from the 8 corners, sum-factorised: with C_k = cofactor matrix of
// the Jacobian at Gauss point k (C = det J * J^-1, element_geometry in host_solver.cpp) and reference gradients
// g = sum_j dN_j/dxi(k) x_j,
//    (K x)_i   = nu  sum_k dN_i/dxi(k) . ( C_k^T C_k g / det_k )          Ke = nu int grad N_i . grad N_j
//    (G_c p)_i = -1/rho sum_k dN_i/dxi(k) . ( C_k[c][:] p(k) )            Ge_c = -1/rho int Sp_j dN_i/dx_c   (det cancels)
//    (G^T d)_j = -1/rho sum_k Sp_j(k) sum_c C_k[c][:] . g_c                (g_c = reference gradient of d_c)
//    (Z p)_j   = -   sum_k dSp_j/dxi(k) . ( C_k^T C_k g / det_k )         FE Laplacian of the implicit scheme, GQ.cpp:3677-3772
// Algorithmic bytes per element: node ids + corner coordinates + gathered and scattered nodal values only -- no matrix stream.
// =================================================================================================
// Cofactor matrix C (C[a][b] = det * (J^-1)[a][b], the arithmetic of element_geometry) and det of the Jacobian at Gauss point k.
__device__ __forceinline__ double ebe_cofactor(const double (&xc)[8][3], int k, double (&C)[3][3]) {
   double J[3][3];
#pragma unroll
   for (int a = 0; a < 3; a++)
#pragma unroll
      for (int d = 0; d < 3; d++) {
         double s = 0;
#pragma unroll
         for (int m = 0; m < 8; m++) s += cdSp[a][m][k] * xc[m][d];
         J[a][d] = s;
      }
   C[0][0] = J[1][1] * J[2][2] - J[2][1] * J[1][2];
   C[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
   C[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
   C[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
   C[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
   C[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
   C[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
   C[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
   C[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
   return J[0][0] * C[0][0] + J[0][1] * C[1][0] + J[0][2] * C[2][0];
}

// out_c[i] += sign * ( [CONV] C(Uadv) X_c + [VISC] K X_c + [PRES] G_c P )_i   for the three components, one pass over the elements.
//   step 1 (CONV | VISC | PRES, sign -1):  R1 = -C(Un) U - K U - G Pn        (calculateMatrixA + step1, BCH.cpp:3520-3801)
//   step 3 (PRES):                          G Pdot                             (BCH.cpp:3912-3974)
//   K * Acc_prev (VISC):                                                        (BCH.cpp:2985-3033)
enum { kEbeConv = 1, kEbeVisc = 2, kEbePres = 4 };
template <int NV, int MODE>
__global__ void __launch_bounds__(128) k_ebe_momentum_apply(int NE, int NN, int ldE, const int* __restrict__ LtoGT, const double* __restrict__ coord,
                                                            const double* __restrict__ Uadv, const double* __restrict__ X, const double* __restrict__ Pv,
                                                            double nu, double invRho, double sign, double* __restrict__ out) {
   const int e = blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= NE) return;
   int nd[NV];
#pragma unroll
   for (int i = 0; i < NV; i++) nd[i] = __ldg(LtoGT + (long long)i * ldE + e);
   double xc[8][3];
#pragma unroll
   for (int m = 0; m < 8; m++)
#pragma unroll
      for (int d = 0; d < 3; d++) xc[m][d] = __ldg(coord + 3 * (long long)nd[m] + d);
   double pk[8];   // pressure at the Gauss points
   if (MODE & kEbePres) {
#pragma unroll
      for (int k = 0; k < 8; k++) pk[k] = 0.0;
#pragma unroll
      for (int j = 0; j < 8; j++) {
         const double v = __ldg(Pv + nd[j]);
#pragma unroll
         for (int k = 0; k < 8; k++) pk[k] += cSv8[k][j] * v;
      }
   }
   double ug[8][3];   // advecting velocity at the Gauss points
   if (MODE & kEbeConv) {
#pragma unroll
      for (int k = 0; k < 8; k++) { ug[k][0] = 0; ug[k][1] = 0; ug[k][2] = 0; }
#pragma unroll
      for (int c = 0; c < 3; c++)
#pragma unroll
         for (int i = 0; i < NV; i++) {
            const double v = __ldg(Uadv + (long long)c * NN + nd[i]);
#pragma unroll
            for (int k = 0; k < 8; k++) ug[k][c] += Tab<NV>::Sv(k, i) * v;
         }
   }
#pragma unroll 1
   for (int c = 0; c < 3; c++) {
      double g[8][3];   // reference gradient of X_c at the Gauss points, turned into the flux h below
      if (MODE & (kEbeConv | kEbeVisc)) {
#pragma unroll
         for (int k = 0; k < 8; k++) { g[k][0] = 0; g[k][1] = 0; g[k][2] = 0; }
#pragma unroll
         for (int j = 0; j < NV; j++) {
            const double v = __ldg(X + (long long)c * NN + nd[j]);
#pragma unroll
            for (int k = 0; k < 8; k++) {
               g[k][0] += Tab<NV>::dSv(0, j, k) * v;
               g[k][1] += Tab<NV>::dSv(1, j, k) * v;
               g[k][2] += Tab<NV>::dSv(2, j, k) * v;
            }
         }
      }
      double t[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
         double C[3][3];
         const double det = ebe_cofactor(xc, k, C);
         double h0 = 0, h1 = 0, h2 = 0;
         t[k] = 0.0;
         if (MODE & (kEbeConv | kEbeVisc)) {
            // w = C g (= det * physical gradient of X_c)
            const double w0 = C[0][0] * g[k][0] + C[0][1] * g[k][1] + C[0][2] * g[k][2];
            const double w1 = C[1][0] * g[k][0] + C[1][1] * g[k][1] + C[1][2] * g[k][2];
            const double w2 = C[2][0] * g[k][0] + C[2][1] * g[k][1] + C[2][2] * g[k][2];
            if (MODE & kEbeConv) t[k] = ug[k][0] * w0 + ug[k][1] * w1 + ug[k][2] * w2;
            if (MODE & kEbeVisc) {
               const double f = nu / det;
               h0 = f * (C[0][0] * w0 + C[1][0] * w1 + C[2][0] * w2);
               h1 = f * (C[0][1] * w0 + C[1][1] * w1 + C[2][1] * w2);
               h2 = f * (C[0][2] * w0 + C[1][2] * w1 + C[2][2] * w2);
            }
         }
         if (MODE & kEbePres) {
            const double q = -invRho * pk[k];
            h0 += q * C[c][0]; h1 += q * C[c][1]; h2 += q * C[c][2];
         }
         g[k][0] = h0; g[k][1] = h1; g[k][2] = h2;
      }
#pragma unroll
      for (int i = 0; i < NV; i++) {
         double r = 0;
#pragma unroll
         for (int k = 0; k < 8; k++) {
            if (MODE & kEbeConv) r += Tab<NV>::Sv(k, i) * t[k];
            if (MODE & (kEbeVisc | kEbePres))
               r += Tab<NV>::dSv(0, i, k) * g[k][0] + Tab<NV>::dSv(1, i, k) * g[k][1] + Tab<NV>::dSv(2, i, k) * g[k][2];
         }
         atomicAdd(out + (long long)c * NN + nd[i], sign * r);
      }
   }
}

// R2[j] += sum_c (G_c^T d_c)_j on the 8 pressure nodes of every element, d formed on the fly like k_gt_rhs:
//   mode 0 (explicit): d = scale * U - [useKAcc] MdOrigInv * KAcc        mode 1 (implicit): d = scale * U  (scale = -1/dt)
template <int NV>
__global__ void __launch_bounds__(128) k_ebe_gt_apply(int NE, int NN, int ldE, const int* __restrict__ LtoGT, const double* __restrict__ coord, int mode,
                                                      const double* __restrict__ U, const double* __restrict__ KAcc, int useKAcc,
                                                      const double* __restrict__ MdOrigInv, double scale, double invRho, double* __restrict__ R2) {
   const int e = blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= NE) return;
   int nd[NV];
#pragma unroll
   for (int i = 0; i < NV; i++) nd[i] = __ldg(LtoGT + (long long)i * ldE + e);
   double xc[8][3];
#pragma unroll
   for (int m = 0; m < 8; m++)
#pragma unroll
      for (int d = 0; d < 3; d++) xc[m][d] = __ldg(coord + 3 * (long long)nd[m] + d);
   double sk[8];
#pragma unroll
   for (int k = 0; k < 8; k++) sk[k] = 0.0;
#pragma unroll 1
   for (int c = 0; c < 3; c++) {
      double g[8][3];
#pragma unroll
      for (int k = 0; k < 8; k++) { g[k][0] = 0; g[k][1] = 0; g[k][2] = 0; }
#pragma unroll
      for (int j = 0; j < NV; j++) {
         const long long ix = (long long)c * NN + nd[j];
         double v = scale * __ldg(U + ix);
         if (mode == 0 && useKAcc) v -= __ldg(MdOrigInv + ix) * __ldg(KAcc + ix);
#pragma unroll
         for (int k = 0; k < 8; k++) {
            g[k][0] += Tab<NV>::dSv(0, j, k) * v;
            g[k][1] += Tab<NV>::dSv(1, j, k) * v;
            g[k][2] += Tab<NV>::dSv(2, j, k) * v;
         }
      }
#pragma unroll
      for (int k = 0; k < 8; k++) {
         double C[3][3];
         ebe_cofactor(xc, k, C);
         sk[k] += C[c][0] * g[k][0] + C[c][1] * g[k][1] + C[c][2] * g[k][2];
      }
   }
#pragma unroll
   for (int j = 0; j < 8; j++) {
      double r = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) r += cSv8[k][j] * sk[k];
      atomicAdd(R2 + nd[j], -invRho * r);
   }
}

// q[j] += (Z p)_j with Z the FE Laplacian of the implicit scheme (GQ.cpp:3677-3772: Ze = - int grad Sp_i . grad Sp_j), on the corner nodes.
template <int NV>
__global__ void __launch_bounds__(128) k_ebe_laplace_apply(int NE, int ldE, const int* __restrict__ LtoGT, const double* __restrict__ coord,
                                                           const double* __restrict__ pv, double* __restrict__ q) {
   const int e = blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= NE) return;
   int nd[8];
#pragma unroll
   for (int i = 0; i < 8; i++) nd[i] = __ldg(LtoGT + (long long)i * ldE + e);
   double xc[8][3], g[8][3];
#pragma unroll
   for (int m = 0; m < 8; m++)
#pragma unroll
      for (int d = 0; d < 3; d++) xc[m][d] = __ldg(coord + 3 * (long long)nd[m] + d);
#pragma unroll
   for (int k = 0; k < 8; k++) { g[k][0] = 0; g[k][1] = 0; g[k][2] = 0; }
#pragma unroll
   for (int j = 0; j < 8; j++) {
      const double v = __ldg(pv + nd[j]);
#pragma unroll
      for (int k = 0; k < 8; k++) {
         g[k][0] += cdSp[0][j][k] * v; g[k][1] += cdSp[1][j][k] * v; g[k][2] += cdSp[2][j][k] * v;
      }
   }
#pragma unroll
   for (int k = 0; k < 8; k++) {
      double C[3][3];
      const double det = ebe_cofactor(xc, k, C);
      const double w0 = C[0][0] * g[k][0] + C[0][1] * g[k][1] + C[0][2] * g[k][2];
      const double w1 = C[1][0] * g[k][0] + C[1][1] * g[k][1] + C[1][2] * g[k][2];
      const double w2 = C[2][0] * g[k][0] + C[2][1] * g[k][1] + C[2][2] * g[k][2];
      const double f = -1.0 / det;
      g[k][0] = f * (C[0][0] * w0 + C[1][0] * w1 + C[2][0] * w2);
      g[k][1] = f * (C[0][1] * w0 + C[1][1] * w1 + C[2][1] * w2);
      g[k][2] = f * (C[0][2] * w0 + C[1][2] * w1 + C[2][2] * w2);
   }
#pragma unroll
   for (int i = 0; i < 8; i++) {
      double r = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) r += cdSp[0][i][k] * g[k][0] + cdSp[1][i][k] * g[k][1] + cdSp[2][i][k] * g[k][2];
      atomicAdd(q + nd[i], r);
   }
}

// The pointwise halves of step 1 and step 3 after the element pass has accumulated the operator part (matrix-free path):
//   step 1:  R1[bc] = 0 ; UnpHalf = Un + dt * R1 * MdInv
// =================================================================================================
// step0 on the device (cfd_options.step0_device = 1): the element integrals of step0
// (BCH.cpp:3127-3315, GQ.cpp:3425-3572, 3677-3772) evaluated by the GPU and summed straight into CSR-ordered value arrays, the
// lumped mass, the explicit pressure matrix Z = sum_c G_c^T Md^-1 G_c (calculateZ, BCH.cpp:3322-3479 / BCH.cu:281-411) as a row-wise
// sparse product, the pin -- then one permutation pass per matrix into the Sell layout.  Same formulas and the same order of the
// Gauss-point sums as the host path (assemble_step0, host_solver.cpp); the element contributions of an entry arrive through FP64
// atomics, i.e. in any order: equal to the host arrays to rounding.  Set-up only, not on the time-loop path.
// =================================================================================================
// position of column c in CSR row `row` (ascending columns); -1 if absent
__device__ __forceinline__ int csr_find(const int* __restrict__ rs, const int* __restrict__ col, int row, int c) {
   int lo = __ldg(rs + row), hi = __ldg(rs + row + 1) - 1;
   if (hi < lo) return -1;
   while (lo < hi) {
      const int mid = lo + ((hi - lo) >> 1);   // (lo + hi overflows an int on the 128^3 Q2/Q1 mesh: 1.08e9 entries)
      if (__ldg(col + mid) < c) lo = mid + 1; else hi = mid;
   }
   return __ldg(col + lo) == c ? lo : -1;
}

// one thread per (element, local row i): rows i of Me, Ke, Ge_1..3 and (implicit, i < 8) of the FE-Laplacian element matrix
template <int NV>
__global__ void __launch_bounds__(128) k_step0_elements(int NE, int ldE, const int* __restrict__ LtoGT, const double* __restrict__ coord,
                                                        const int* __restrict__ Mrs, const int* __restrict__ Mcol,
                                                        const int* __restrict__ Grs, const int* __restrict__ Gcol,
                                                        const int* __restrict__ Zrs, const int* __restrict__ Zcol,
                                                        double mscale, double visc, double invDens,
                                                        double* __restrict__ M, double* __restrict__ K, double* __restrict__ G1,
                                                        double* __restrict__ G2, double* __restrict__ G3, double* __restrict__ Z) {
   const int e = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
   if (e >= NE) return;
   double xc[8][3];
#pragma unroll
   for (int m = 0; m < 8; m++) {
      const long long nd = __ldg(LtoGT + (long long)m * ldE + e);
#pragma unroll
      for (int d = 0; d < 3; d++) xc[m][d] = __ldg(coord + 3 * nd + d);
   }
   // inverse Jacobian and quadrature factor at the 8 Gauss points (calcJacob, BCH.cpp:2495-2711: unit weights, fac = det J)
   double I[8][9], fac[8], gi[8][3];
#pragma unroll
   for (int k = 0; k < 8; k++) {
      double C[3][3];
      const double det = ebe_cofactor(xc, k, C);
      const double id = 1.0 / det;
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
         for (int b = 0; b < 3; b++) I[k][3 * a + b] = C[a][b] * id;
      fac[k] = det;
#pragma unroll
      for (int a = 0; a < 3; a++)
         gi[k][a] = I[k][3 * a] * Tab<NV>::dSv(0, i, k) + I[k][3 * a + 1] * Tab<NV>::dSv(1, i, k) + I[k][3 * a + 2] * Tab<NV>::dSv(2, i, k);
   }
   const int row = __ldg(LtoGT + (long long)i * ldE + e);
   for (int j = 0; j < NV; j++) {
      double me = 0.0, ke = 0.0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
         double gj[3];
#pragma unroll
         for (int a = 0; a < 3; a++)
            gj[a] = I[k][3 * a] * Tab<NV>::dSv(0, j, k) + I[k][3 * a + 1] * Tab<NV>::dSv(1, j, k) + I[k][3 * a + 2] * Tab<NV>::dSv(2, j, k);
         me += Tab<NV>::Sv(k, i) * Tab<NV>::Sv(k, j) * fac[k];
         ke += (gi[k][0] * gj[0] + gi[k][1] * gj[1] + gi[k][2] * gj[2]) * fac[k];
      }
      const int pos = csr_find(Mrs, Mcol, row, __ldg(LtoGT + (long long)j * ldE + e));
      if (pos >= 0) { atomicAdd(M + pos, mscale * me); atomicAdd(K + pos, visc * ke); }
   }
   for (int j = 0; j < 8; j++) {
      double g1 = 0.0, g2 = 0.0, g3 = 0.0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
         const double w = cSv8[k][j] * fac[k];
         g1 += w * gi[k][0]; g2 += w * gi[k][1]; g3 += w * gi[k][2];
      }
      const int pos = csr_find(Grs, Gcol, row, __ldg(LtoGT + (long long)j * ldE + e));
      if (pos >= 0) { atomicAdd(G1 + pos, -invDens * g1); atomicAdd(G2 + pos, -invDens * g2); atomicAdd(G3 + pos, -invDens * g3); }
   }
   if (Z != nullptr && i < 8) {   // FE Laplacian on the pressure basis (implicit Z, GQ.cpp:3677-3772)
      double gp[8][3];
#pragma unroll
      for (int k = 0; k < 8; k++)
#pragma unroll
         for (int a = 0; a < 3; a++) gp[k][a] = I[k][3 * a] * cdSp[0][i][k] + I[k][3 * a + 1] * cdSp[1][i][k] + I[k][3 * a + 2] * cdSp[2][i][k];
      for (int j = 0; j < 8; j++) {
         double z = 0.0;
#pragma unroll
         for (int k = 0; k < 8; k++) {
            double gj[3];
#pragma unroll
            for (int a = 0; a < 3; a++) gj[a] = I[k][3 * a] * cdSp[0][j][k] + I[k][3 * a + 1] * cdSp[1][j][k] + I[k][3 * a + 2] * cdSp[2][j][k];
            z += (gp[k][0] * gj[0] + gp[k][1] * gj[1] + gp[k][2] * gj[2]) * fac[k];
         }
         const int pos = csr_find(Zrs, Zcol, row, __ldg(LtoGT + (long long)j * ldE + e));
         if (pos >= 0) atomicAdd(Z + pos, -z);
      }
   }
}

// lumped mass = row sums of M in CSR order (BCH.cpp:3254-3292); MdOrigInv = 1/md, MdInv = 1/md with Dirichlet rows set to 1
__global__ void __launch_bounds__(256) k_step0_lump(int NN, const int* __restrict__ rs, const double* __restrict__ M, const int* __restrict__ bcIndex,
                                                    double* __restrict__ md, double* __restrict__ MdOrigInv, double* __restrict__ MdInv) {
   const int r = blockIdx.x * blockDim.x + threadIdx.x;
   if (r >= NN) return;
   double s = 0.0;
   for (int k = rs[r]; k < rs[r + 1]; k++) s += M[k];
   md[r] = s;
   const double inv = 1.0 / s, invBC = bcIndex[r] >= 0 ? 1.0 : inv;
#pragma unroll
   for (int c = 0; c < 3; c++) { MdOrigInv[(size_t)c * NN + r] = inv; MdInv[(size_t)c * NN + r] = invBC; }
}

// Z = sum_c G_c^T diag(1/md) G_c, one warp per pressure row i through the transposed pattern of G (ts / trow / tpos): for every
// velocity node r coupled to i (ascending), the lanes take the entries of row r of G and add w . G[r][j] to Z[i][j]
// -- the accumulation order of the host product (assemble_step0).  Z must be zero on entry.
__global__ void __launch_bounds__(256) k_step0_spgemm_z(int NNp, const int* __restrict__ ts, const int* __restrict__ trow, const int* __restrict__ tpos,
                                                        const int* __restrict__ Grs, const int* __restrict__ Gcol, const double* __restrict__ G1,
                                                        const double* __restrict__ G2, const double* __restrict__ G3, const double* __restrict__ md,
                                                        const int* __restrict__ Zrs, const int* __restrict__ Zcol, double* Z) {
   const int i = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
   if (i >= NNp) return;
   for (int t = ts[i]; t < ts[i + 1]; t++) {
      const int r = trow[t], q = tpos[t];
      const double d = md[r];
      const double w1 = G1[q] / d, w2 = G2[q] / d, w3 = G3[q] / d;
      for (int k = Grs[r] + lane; k < Grs[r + 1]; k += 32) {
         const int pos = csr_find(Zrs, Zcol, i, Gcol[k]);
         if (pos >= 0) Z[pos] += w1 * G1[k] + w2 * G2[k] + w3 * G3[k];
      }
      __syncwarp();   // the next r may touch the same Z entries from other lanes
   }
}

// implicit scheme: element -> device-layout position map of the K / M / A pattern ([i*NV+j][e], what k_ebe_convect_assemble scatters
// through) and the positions of the diagonal, found by bisection instead of being marshalled from the host's sparseMapM_1d
template <int NV>
__global__ void __launch_bounds__(128) k_step0_map(int NE, int ldE, const int* __restrict__ LtoGT, const int* __restrict__ Mrs, const int* __restrict__ Mcol,
                                                   const long long* __restrict__ perm, int* __restrict__ mapT) {
   const int e = blockIdx.x * blockDim.x + threadIdx.x, q = blockIdx.y;
   if (e >= ldE) return;
   int out = 0;
   if (e < NE) {
      const int pos = csr_find(Mrs, Mcol, __ldg(LtoGT + (long long)(q / NV) * ldE + e), __ldg(LtoGT + (long long)(q % NV) * ldE + e));
      out = pos >= 0 ? (int)perm[pos] : -1;
   }
   mapT[(long long)q * ldE + e] = out;
}
__global__ void __launch_bounds__(256) k_step0_diagpos(int n, const int* __restrict__ rs, const int* __restrict__ col, const long long* __restrict__ perm,
                                                       int* __restrict__ diagPos) {
   const int r = blockIdx.x * blockDim.x + threadIdx.x;
   if (r >= n) return;
   const int pos = csr_find(rs, col, r, r);
   diagPos[r] = pos >= 0 ? (int)perm[pos] : 0;
}

// dst[perm[k]] = a[k] (+ b[k]): CSR order -> Sell layout;  dst[k] = src[idx[k]]: transposed values;  1/diag of a CSR matrix
__global__ void __launch_bounds__(256) k_step0_scatter(long long n, const long long* __restrict__ perm, const double* __restrict__ a,
                                                       const double* __restrict__ b, double* __restrict__ dst) {
   const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
   if (k < n) dst[perm[k]] = b ? a[k] + b[k] : a[k];
}
__global__ void __launch_bounds__(256) k_step0_gather(long long n, const int* __restrict__ idx, const double* __restrict__ src, double* __restrict__ dst) {
   const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
   if (k < n) dst[k] = src[idx[k]];
}
__global__ void __launch_bounds__(256) k_step0_diag(int n, const int* __restrict__ rs, const int* __restrict__ col, double* val, int pin, int noPrecond,
                                                    double* __restrict__ dinv) {
   const int r = blockIdx.x * blockDim.x + threadIdx.x;
   if (r >= n) return;
   const int pos = csr_find(rs, col, r, r);
   if (pos >= 0 && r == pin) val[pos] *= 1000.0;   // pressure pin: diagonal x 1000 (applyBC_Step2, BCH.cpp:4167-4180)
   dinv[r] = (noPrecond || pos < 0) ? 1.0 : 1.0 / val[pos];
}

__global__ void __launch_bounds__(256) k_step1_finish(int NN, const int* __restrict__ bcIndex, const double* __restrict__ Un, const double* __restrict__ MdInv,
                                                      double* __restrict__ R1, double* __restrict__ UnpHalf, double dt) {
   const int row = blockIdx.x * blockDim.x + threadIdx.x;
   if (row >= NN) return;
   const bool bc = bcIndex[row] >= 0;
#pragma unroll
   for (int c = 0; c < 3; c++) {
      const long long i = (long long)c * NN + row;
      const double r = bc ? 0.0 : R1[i];
      R1[i] = r;
      UnpHalf[i] = Un[i] + dt * r * MdInv[i];
   }
}
//   step 3:  R3 = -dt * (G Pdot) - [iter>1] dt * KAcc ; R3[bc] = 0 ; Acc = R3 * MdInv ; Unp1 = UnpHalf + dt * Acc ; convergence sums
__global__ void __launch_bounds__(256) k_step3_finish(int NN, const double* __restrict__ KAcc, int useKAcc, const double* __restrict__ MdInv,
                                                      const int* __restrict__ bcIndex, const double* __restrict__ UnpHalf, const double* __restrict__ Unp1_prev,
                                                      double* __restrict__ R3, double* __restrict__ Acc, double* __restrict__ Unp1, double dt,
                                                      double* __restrict__ sums) {
   __shared__ double sh[64];
   const int row = blockIdx.x * blockDim.x + threadIdx.x;
   double s[2] = {0.0, 0.0};
   if (row < NN) {
      const bool bc = bcIndex[row] >= 0;
#pragma unroll
      for (int c = 0; c < 3; c++) {
         const long long i = (long long)c * NN + row;
         double r = -dt * R3[i];   // R3 holds G_c Pdot from the element pass
         if (useKAcc) r = r - dt * KAcc[i];
         if (bc) r = 0.0;
         const double a = r * MdInv[i];
         const double u = UnpHalf[i] + dt * a;
         R3[i] = r; Acc[i] = a; Unp1[i] = u;
         const double d = u - Unp1_prev[i];
         s[0] += u * u; s[1] += d * d;
      }
   }
   block_sum<2>(s, sh);
   if (threadIdx.x == 0) { atomicAdd(sums + 0, s[0]); atomicAdd(sums + 1, s[1]); }
}

// Dirichlet rows of A -> identity rows (applyVelBC_LHS, GQ.cu:478-498), on the Sell layout.
__global__ void k_bc_rows_identity(Sell A, int vidx, int nbc, const int* __restrict__ bcNodes, const int* __restrict__ diagPos) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= nbc) return;
   const int row = bcNodes[i];
   const int slice = row / A.R, rl = row % A.R;
   const long long off = A.sliceOff[slice];
   const int w = (int)((A.sliceOff[slice + 1] - off) >> 5);
   double* val = vidx == 0 ? A.vals[0] : (vidx == 1 ? A.vals[1] : A.vals[2]);
   const long long dpos = diagPos[row];
   for (int lane = 0; lane < A.T; lane++) {
      const int t = rl * A.T + lane;
      for (int j = 0; j < w; j++) {
         const long long pos = off + (long long)(j >> 1) * 64 + 2 * t + (j & 1);
         val[pos] = (pos == dpos) ? 1.0 : 0.0;
      }
   }
}

__global__ void k_extract_dinv(int n, const double* __restrict__ val, const int* __restrict__ diagPos, double* __restrict__ dinv) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) dinv[i] = 1.0 / val[diagPos[i]];
}

// Right-hand sides of the three implicit momentum systems in one pass (replaces calculate_dummyPdiff +
// per-direction 2 csrmv + applyVelBC_RHS, GQ.cu:809-855):
//   R1_d = M*Uk_prev_d - G_d*(2 Pk_prev - Pk_prevprev) ;  R1_d[bc] = BCstr[whichBC][d]
__global__ void __launch_bounds__(256) k_gq_rhs(Sell MK, int mIdx, Sell G, int NN, const double* __restrict__ Uprev,
                                                const double* __restrict__ Pk_prev, const double* __restrict__ Pk_prevprev,
                                                const int* __restrict__ bcIndex, const double* __restrict__ BCstr,
                                                double* __restrict__ R1) {
   RowMap m;
   if (!row_map(MK, m)) return;
   double a0 = 0, a1 = 0, a2 = 0;
   lane_dot_k3(MK, sell_vals(MK, mIdx), m, Uprev, NN, a0, a1, a2);
   lane_dot_g3(G, m, [Pk_prev, Pk_prevprev](int c) { return 2 * __ldg(Pk_prev + c) - __ldg(Pk_prevprev + c); }, -1.0, a0, a1, a2);
   a0 = lanes_sum(a0, MK.T); a1 = lanes_sum(a1, MK.T); a2 = lanes_sum(a2, MK.T);
   if (m.owner) {
      const int w = bcIndex[m.row];
      R1[m.row] = (w >= 0) ? BCstr[3 * w + 0] : a0;
      R1[(long long)NN + m.row] = (w >= 0) ? BCstr[3 * w + 1] : a1;
      R1[2 * (long long)NN + m.row] = (w >= 0) ? BCstr[3 * w + 2] : a2;
   }
}

// y_c = G_c x (three value arrays on one pattern, rows = velocity nodes): the assembled counterpart of the matrix-free G apply
__global__ void __launch_bounds__(256) k_g_apply(Sell G, int NN, const double* __restrict__ x, double* __restrict__ y) {
   RowMap m;
   if (!row_map(G, m)) return;
   double a0 = 0, a1 = 0, a2 = 0;
   lane_dot_g3(G, m, [x](int c) { return __ldg(x + c); }, 1.0, a0, a1, a2);
   a0 = lanes_sum(a0, G.T); a1 = lanes_sum(a1, G.T); a2 = lanes_sum(a2, G.T);
   if (m.owner) { y[m.row] = a0; y[(long long)NN + m.row] = a1; y[2 * (long long)NN + m.row] = a2; }
}

__global__ void k_axpy_out(long long n, const double* __restrict__ a, const double* __restrict__ b, double s, double* __restrict__ out) {
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      out[i] = a[i] + s * b[i];
}
__global__ void k_add(long long n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      out[i] = a[i] + b[i];
}

#include "dist.cuh"

// =================================================================================================
// Persistent cooperative Krylov kernels
// =================================================================================================
struct PcgArgs {
   Sell Z;
   const int* part;   // per-block slice ranges of the static share (may be NULL)
   int dynStart;      // first slice of the dynamically scheduled tail
   int* dynCounter;   // NULL = everything static
   int n;
   const double* b;
   double* x;
   double *r, *p, *q;
   const double* dinv;
   double* scratch;   // 2 buffers x 3 x gridDim doubles
   double* info;      // [0]=iterations [1]=|r|^2 [2]=|b|^2
   double tol;
   int maxit;
   int x0zero;
};

// q = Z*p for the slices owned by this warp; returns the partial sum of p_i q_i over its rows.
// p is written by this same kernel in other phases: plain (coherent) loads, not __ldg.
// Slice -> warp assignment inside the persistent kernels.  Many slices per warp: contiguous chunks per
// block (block b owns [b*chunk, (b+1)*chunk)), so the vector entries a block gathers form one short
// window that stays in its SM's L1 for the whole phase.  Few slices per warp: plain grid-stride, which
// balances the load to within one slice.
struct SliceRange { int begin, end, step; };
__device__ __forceinline__ SliceRange slice_range(int nslices, const int* __restrict__ part, int gwarp, int nwarps) {
   SliceRange r;
   if (part && nslices >= 4 * nwarps) {
      // part[b] .. part[b+1]: chunk of block b, balanced by stored entries (rows of a Q2 mesh differ 4.6x in length)
      r.begin = part[blockIdx.x] + (threadIdx.x >> 5);
      r.end = part[blockIdx.x + 1];
      r.step = blockDim.x >> 5;
   } else {
      r.begin = gwarp; r.end = nslices; r.step = nwarps;
   }
   return r;
}

// One slice of q = Z p by one warp; returns this lane's contribution to p.q.
__device__ __forceinline__ double pcg_slice(const Sell& Z, const double* __restrict__ val, int s, const double* p, double* q, int lane) {
   const int T = Z.T, sub = lane % T;
   double acc = 0.0;
   sell_pairs<4>(Z, s, lane, [&](long long e, int2 c2) {
      const double2 v2 = ldnc_d2(val + e);
      acc += v2.x * p[c2.x] + v2.y * p[c2.y];
   });
   for (int o = T >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
   const int row = s * Z.R + lane / T;
   if (sub == 0 && row < Z.nrows) {
      q[row] = acc;
      return p[row] * acc;
   }
   return 0.0;
}

// q = Z p over the whole matrix.  Slices [0, dynStart) are dealt statically (slice_range); the tail
// [dynStart, nslices) is handed out two slices at a time through an atomic counter, so that SMs which finish
// their static share early (finish times differ by ~15 %) take the remaining work instead of idling at the
// grid reduction that follows.  `dynCounter` must be 0 on entry; the caller resets it after the next grid sync.
__device__ __forceinline__ double pcg_spmv_phase(const Sell& Z, const int* part, int dynStart, int* dynCounter, const double* p, double* q,
                                                 int gwarp, int nwarps, int lane) {
   double pq = 0.0;
   const double* __restrict__ val = Z.vals[0];
   const int nstatic = dynCounter ? dynStart : Z.nslices;
   const SliceRange sr = slice_range(nstatic, part, gwarp, nwarps);
   for (int s = sr.begin; s < sr.end; s += sr.step) pq += pcg_slice(Z, val, s, p, q, lane);
   if (dynCounter) {
      while (true) {
         int base = 0;
         if (lane == 0) base = atomicAdd(dynCounter, 2);
         base = __shfl_sync(0xffffffffu, base, 0) + dynStart;
         if (base >= Z.nslices) break;
         pq += pcg_slice(Z, val, base, p, q, lane);
         if (base + 1 < Z.nslices) pq += pcg_slice(Z, val, base + 1, p, q, lane);
      }
   }
   return pq;
}

// Jacobi-preconditioned CG, whole solve in one launch (replaces CUSP_CG_solver, BCH.cu:418-451 /
// GQ.cu:302-362: cusp::krylov::cg + diagonal preconditioner + default_monitor(b, maxit, tol)).
// 3 grid syncs per iteration; all reductions on the device; no host round trip.
template <int MINB>
__global__ void __launch_bounds__(512, MINB) k_pcg_t(PcgArgs a) {
   cg::grid_group grid = cg::this_grid();
   __shared__ double sh[3 * 32];
   const int tid = blockIdx.x * blockDim.x + threadIdx.x;
   const int nthreads = gridDim.x * blockDim.x;
   const int gwarp = tid >> 5, nwarps = nthreads >> 5, lane = threadIdx.x & 31;
   double* buf0 = a.scratch;
   double* buf1 = a.scratch + 3 * (size_t)gridDim.x;
   const int n = a.n;

   if (!a.x0zero) {
      // r = b - Z x  (warm start, BCH.cu:418-451 keeps the previous Pdot)
      pcg_spmv_phase(a.Z, a.part, a.dynStart, a.dynCounter, a.x, a.q, gwarp, nwarps, lane);
      grid.sync();
      if (tid == 0 && a.dynCounter) *a.dynCounter = 0;
   }
   double v3[3] = {0, 0, 0};
   for (int i = tid; i < n; i += nthreads) {
      const double bi = a.b[i];
      double ri = bi;
      if (a.x0zero) a.x[i] = 0.0; else ri = bi - a.q[i];
      const double zi = a.dinv[i] * ri;
      a.r[i] = ri;
      a.p[i] = zi;
      v3[0] += ri * zi; v3[1] += bi * bi; v3[2] += ri * ri;
   }
   grid_sum<3>(grid, v3, buf0, sh);
   double rz = v3[0];
   const double bb = v3[1];
   double rr = v3[2];
   const double thresh = a.tol * sqrt(bb);
   int it = 0;
   // phase clock of block 0 (diagnostics: info[8..12] = ns spent in spmv / reduce pq / update / reduce rz / p-update+sync)
#ifdef CFD_PCG_PHASE_TIMERS
   unsigned long long tph[5] = {0, 0, 0, 0, 0}, t0 = 0, t1;
#define PCG_TICK(k) do { if (tid == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); tph[k] += t1 - t0; t0 = t1; } } while (0)
   if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
#else
#define PCG_TICK(k)
#endif
   while (true) {
      if (sqrt(rr) <= thresh || it >= a.maxit) break;
      double v1[1];
      v1[0] = pcg_spmv_phase(a.Z, a.part, a.dynStart, a.dynCounter, a.p, a.q, gwarp, nwarps, lane);
      PCG_TICK(0);
      grid_sum<1>(grid, v1, buf1, sh);
      if (tid == 0 && a.dynCounter) *a.dynCounter = 0;   // every block is past the SpMV phase
      PCG_TICK(1);
      const double alpha = rz / v1[0];
      double v2[2] = {0, 0};
      for (int i = tid; i < n; i += nthreads) {
         a.x[i] += alpha * a.p[i];
         const double ri = a.r[i] - alpha * a.q[i];
         a.r[i] = ri;
         v2[0] += ri * (a.dinv[i] * ri);
         v2[1] += ri * ri;
      }
      PCG_TICK(2);
      grid_sum<2>(grid, v2, buf0, sh);
      PCG_TICK(3);
      const double beta = v2[0] / rz;
      rz = v2[0];
      rr = v2[1];
      for (int i = tid; i < n; i += nthreads) a.p[i] = a.dinv[i] * a.r[i] + beta * a.p[i];
      it++;
      grid.sync();
      PCG_TICK(4);
   }
#undef PCG_TICK
#ifdef CFD_PCG_PHASE_TIMERS
   if (tid == 0)
      for (int k = 0; k < 5; k++) a.info[8 + k] = (double)tph[k];
#endif
   if (tid == 0) { a.info[0] = (double)it; a.info[1] = rr; a.info[2] = bb; }
}

// two register budgets of the same kernel: 3 CTAs/SM (40 registers) or 2 CTAs/SM (64 registers); chosen at init
static void* pcg_kernel(int minb) { return minb == 2 ? (void*)k_pcg_t<2> : (void*)k_pcg_t<3>; }

// ---------------------------------------------------------------------------------------------------------------------------
// Jacobi-PCG with the MATRIX RESIDENT IN SHARED MEMORY for the whole solve: small pressure systems (the reference's shipped
// 30^3 input: Z = G^T Md^-1 G, 29 791 rows, 3.3 M non-zeros = 27 MB of values) fit into the 148 x 227 KB of shared memory of a
// B200.  One CTA of 1024 threads per SM owns a contiguous range of slices (balanced by bytes, dealt to its 32 warps) and
// copies their values in once; every CG iteration then reads the matrix from shared memory (LDS.128) instead of streaming it
// from L2 -- what is left per iteration is the p gather, three vector passes and the three grid-wide synchronisations, now over
// 148 CTAs.  Same recurrences and the same arithmetic per row as k_pcg_t; the dot-product partials are grouped by other CTAs, so the
// iterates agree to rounding, not bitwise.  Statically scheduled: bitwise reproducible from run to run.  Column indices stay
// where they are (compressed offset lists are a few hundred bytes per slice and live in L1).
// ---------------------------------------------------------------------------------------------------------------------------
struct PcgSmemArgs {
   PcgArgs a;            // a.part = slice range per CTA [gridDim + 1]
   double *s, *pp;       // single-reduction form only: s = Z p, and the search direction (a.p then holds u = M^-1 r, a.q holds w = Z u)
};
// SINGLE = true: the single-reduction (Chronopoulos-Gear) recurrences of k_pcg_cg -- TWO grid-wide synchronisations per iteration
// instead of three.  Here everything but the synchronisations is cheap (matrix in shared memory, vectors in L2, 64 registers at one
// CTA per SM anyway), so the form that loses on large streamed systems wins: measured on the shipped 30^3 input, us per CG
// iteration: streaming classical 21.8, resident classical 16.4, resident single-reduction -> profiles/r02_cg_variants.md.
template <bool SINGLE>
__global__ void __launch_bounds__(1024, 1) k_pcg_smem(PcgSmemArgs pa) {
   cg::grid_group grid = cg::this_grid();
   extern __shared__ double zs[];   // this CTA's slices' values, in layout order
   __shared__ double sh[4 * 32];
   const PcgArgs& a = pa.a;
   const int tid = blockIdx.x * blockDim.x + threadIdx.x;
   const int nthreads = gridDim.x * blockDim.x;
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   double* buf0 = a.scratch;
   double* buf1 = a.scratch + 4 * (size_t)gridDim.x;
   const int n = a.n;
   const int s0 = a.part[blockIdx.x], s1 = a.part[blockIdx.x + 1];
   const long long e0 = a.Z.sliceOff[s0], e1 = a.Z.sliceOff[s1];
   {
      const double2* __restrict__ src = reinterpret_cast<const double2*>(a.Z.vals[0] + e0);
      double2* dst = reinterpret_cast<double2*>(zs);
      for (long long k = threadIdx.x; k < ((e1 - e0) >> 1); k += blockDim.x) dst[k] = src[k];
   }
   __syncthreads();
   // q = Z p on this CTA's slices, matrix values from shared memory; returns this lane's share of p.q
   auto spmv = [&](const double* p, double* q) {
      double pq = 0.0;
      for (int s = s0 + warp; s < s1; s += (blockDim.x >> 5)) {
         const int T = a.Z.T, sub = lane % T;
         double acc = 0.0;
         sell_pairs<4>(a.Z, s, lane, [&](long long e, int2 c2) {
            const double2 v2 = *reinterpret_cast<const double2*>(zs + (e - e0));
            acc += v2.x * p[c2.x] + v2.y * p[c2.y];
         });
         for (int o = T >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
         const int row = s * a.Z.R + lane / T;
         if (sub == 0 && row < a.Z.nrows) { q[row] = acc; pq += p[row] * acc; }
      }
      return pq;
   };
   if (!a.x0zero) {
      spmv(a.x, a.q);
      grid.sync();
   }
   if (SINGLE) {
      double* u = a.p; double* w = a.q; double* p = pa.pp; double* sv = pa.s;
      double v4[4] = {0, 0, 0, 0};
      for (int i = tid; i < n; i += nthreads) {
         const double bi = a.b[i];
         double ri = bi;
         if (a.x0zero) a.x[i] = 0.0; else ri = bi - w[i];
         a.r[i] = ri;
         const double un = a.dinv[i] * ri;
         u[i] = un;
         v4[0] += ri * un; v4[2] += ri * ri; v4[3] += bi * bi;
      }
      grid.sync();
      v4[1] = spmv(u, w);
      grid_sum<4>(grid, v4, buf0, sh);
      double gamma = v4[0], rr = v4[2];
      const double bb = v4[3];
      const double thresh = a.tol * sqrt(bb);
      double alpha = gamma / v4[1], beta = 0.0;
      int it = 0;
#ifdef CFD_PCG_PHASE_TIMERS
      unsigned long long tph[4] = {0, 0, 0, 0}, t0 = 0, t1;
#define PCG_TICK(k) do { if (tid == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); tph[k] += t1 - t0; t0 = t1; } } while (0)
      if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
#else
#define PCG_TICK(k)
#endif
      while (true) {
         if (sqrt(rr) <= thresh || it >= a.maxit) break;
         const bool first = it == 0;
         double v3[3] = {0, 0, 0};
         for (int i = tid; i < n; i += nthreads) {
            const double ui = u[i], wi = w[i];
            const double pi = first ? ui : ui + beta * p[i];
            const double si = first ? wi : wi + beta * sv[i];
            p[i] = pi; sv[i] = si;
            a.x[i] += alpha * pi;
            const double ri = a.r[i] - alpha * si;
            a.r[i] = ri;
            const double un = a.dinv[i] * ri;
            u[i] = un;
            v3[0] += ri * un; v3[2] += ri * ri;
         }
         it++;
         PCG_TICK(0);
         grid.sync();
         PCG_TICK(1);
         v3[1] = spmv(u, w);
         PCG_TICK(2);
         grid_sum<3>(grid, v3, buf0, sh);
         PCG_TICK(3);
         beta = v3[0] / gamma;
         alpha = v3[0] / (v3[1] - beta * v3[0] / alpha);
         gamma = v3[0];
         rr = v3[2];
      }
#undef PCG_TICK
#ifdef CFD_PCG_PHASE_TIMERS
      if (tid == 0)
         for (int k = 0; k < 4; k++) a.info[8 + k] = (double)tph[k];
#endif
      if (tid == 0) { a.info[0] = (double)it; a.info[1] = rr; a.info[2] = bb; }
      return;
   }
   double v3[3] = {0, 0, 0};
   for (int i = tid; i < n; i += nthreads) {
      const double bi = a.b[i];
      double ri = bi;
      if (a.x0zero) a.x[i] = 0.0; else ri = bi - a.q[i];
      const double zi = a.dinv[i] * ri;
      a.r[i] = ri;
      a.p[i] = zi;
      v3[0] += ri * zi; v3[1] += bi * bi; v3[2] += ri * ri;
   }
   grid_sum<3>(grid, v3, buf0, sh);
   double rz = v3[0];
   const double bb = v3[1];
   double rr = v3[2];
   const double thresh = a.tol * sqrt(bb);
   int it = 0;
   while (true) {
      if (sqrt(rr) <= thresh || it >= a.maxit) break;
      double v1[1];
      v1[0] = spmv(a.p, a.q);
      grid_sum<1>(grid, v1, buf1, sh);
      const double alpha = rz / v1[0];
      double v2[2] = {0, 0};
      for (int i = tid; i < n; i += nthreads) {
         a.x[i] += alpha * a.p[i];
         const double ri = a.r[i] - alpha * a.q[i];
         a.r[i] = ri;
         v2[0] += ri * (a.dinv[i] * ri);
         v2[1] += ri * ri;
      }
      grid_sum<2>(grid, v2, buf0, sh);
      const double beta = v2[0] / rz;
      rz = v2[0];
      rr = v2[1];
      for (int i = tid; i < n; i += nthreads) a.p[i] = a.dinv[i] * a.r[i] + beta * a.p[i];
      it++;
      grid.sync();
   }
   if (tid == 0) { a.info[0] = (double)it; a.info[1] = rr; a.info[2] = bb; }
}

// ---------------------------------------------------------------------------------------------------------------------------
// k_pcg_res: the resident PCG of small pressure systems, second form.  Phase clocks of k_pcg_smem on the shipped 30^3 input
// (profiles/r02_pcg_resident_phases.txt): 5.9 us of the 17 us iteration went into the sweep and 8.3 us into the reduction that waits
// for the slowest sweep -- with the matrix in shared memory the sweep is bound by its p GATHERS: L1 is carved down to ~30 KB and is
// invalidated by every grid-wide fence, so 3.3 M eight-byte gathers per iteration travel as 32-byte sectors from L2.  Here each CTA
// owns a block of ROWS outright:
//   * matrix values of its rows in shared memory, in a private layout (groups of 32/T rows, T lanes per row, entries in pairs),
//   * the columns its rows touch form a short list (a few planes of the pressure lattice); per sweep the CTA copies exactly those
//     vector entries into shared memory with coalesced loads -- the gathers become LDS; column ids are 16-bit positions in that list,
//   * x, r, p, s, u, w, 1/diag of a row live in the REGISTERS of one thread for the whole solve: the vector pass costs no memory
//     traffic but the 8-byte store of u.
// Single-reduction (Chronopoulos-Gear) recurrences: one grid.sync (u visible) + one grid-wide reduction per iteration.  Static
// assignment, fixed summation orders: bitwise reproducible from run to run; iterates agree with k_pcg_t to rounding.
// ---------------------------------------------------------------------------------------------------------------------------
struct ResCta {
   long long valOff;   // first entry of this CTA in vals / idx
   long long grpOff;   // first of its ngroups + 1 group offsets in grp
   long long colOff;   // first of its ncols staged columns in cols
   int row0, nrows, ngroups, ncols;
};
struct PcgResArgs {
   const ResCta* cta;
   const double* vals;           // per CTA: groups of R = 32/T rows; entry (slot j, row rl, lane-in-row t) at grp[g] + (j/2)*64 + 2*(rl*T + t) + (j&1)
   const unsigned short* idx;    // same layout: position of the entry's column in the CTA's staged list
   const int* cols;              // per CTA: the global columns it stages, ascending
   const int* grp;               // per CTA: ngroups + 1 entry offsets (multiples of 64)
   int T;
   int n;
   const double* b;
   double* x;
   double* u;                    // the one vector other CTAs read: u = M^-1 r
   const double* dinv;
   double* scratch;
   double* info;
   double tol;
   int maxit, x0zero;
};
constexpr int kResThreads = 512;   // 16 warps, 128 registers per thread: room for the cached column ids
constexpr int kResWarps = kResThreads / 32;
// (A pipelined form -- Ghysels-Vanroose recurrences, the dot-product partials riding on the barrier that publishes the vector: ONE
// grid-wide barrier per iteration -- was built and measured: 361 against 366 steps/s on the shipped mesh, because the wait at the
// reduction is the wait for the slowest sweep, not the barrier itself; and at a tolerance of 1e-12 its iteration counts leave the +-5
// window of the tight reference fixtures (recurrence drift).  Removed.)
__global__ void __launch_bounds__(kResThreads, 1) k_pcg_res(PcgResArgs a) {
   cg::grid_group grid = cg::this_grid();
   extern __shared__ double zs[];   // values | staged vector entries | this CTA's rows of w = Z u
   __shared__ double sh[4 * 32];
   const ResCta c = a.cta[blockIdx.x];
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   const int* __restrict__ grp = a.grp + c.grpOff;
   const int nvals = grp[c.ngroups];
   double* win = zs + nvals;
   double* wrow = win + ((c.ncols + 1) & ~1);
   {
      const double2* __restrict__ src = reinterpret_cast<const double2*>(a.vals + c.valOff);
      double2* dst = reinterpret_cast<double2*>(zs);
      for (int k = threadIdx.x; k < (nvals >> 1); k += blockDim.x) dst[k] = src[k];
   }
   const unsigned short* __restrict__ idx = a.idx + c.valOff;
   const int* __restrict__ cols = a.cols + c.colOff;
   const int T = a.T, R = 32 / T;
   double* buf0 = a.scratch;
   // wrow = (Z src) on this CTA's rows
#ifdef CFD_PCG_PHASE_TIMERS
   unsigned long long tsw[3] = {0, 0, 0}, u0 = 0, u1;
#define SWEEP_TICK(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(u1)); tsw[k] += u1 - u0; u0 = u1; } } while (0)
#define SWEEP_T0() do { if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(u0)); } while (0)
#else
#define SWEEP_TICK(k)
#define SWEEP_T0()
#endif
   // Column ids of this warp's FIRST row group stay in registers for the whole solve (they never change, and every grid-wide fence
   // invalidates L1: re-reading them costs an L2 round trip per dependent load -- measured 4.3 of the 5.2 us of a sweep).  Further
   // groups of the warp (more than 16 groups per CTA) and pairs beyond kResCache are requested per sweep, a whole chunk at a time.
   constexpr int kResCache = 32;
   unsigned ireg[kResCache];
   int off0 = 0, np0 = 0;
   if (warp < c.ngroups) { off0 = grp[warp]; np0 = (grp[warp + 1] - off0) >> 6; }
#pragma unroll
   for (int q = 0; q < kResCache; q++) ireg[q] = q < np0 ? __ldg(reinterpret_cast<const unsigned*>(idx + off0) + lane + q * 32) : 0u;
   // the staged columns likewise: up to kResCols per thread in registers (one L2 round trip per sweep instead of two)
   constexpr int kResCols = 6;
   int creg[kResCols];
   const bool colsCached = c.ncols <= kResCols * kResThreads;
#pragma unroll
   for (int q = 0; q < kResCols; q++) { const int k = (int)threadIdx.x + q * kResThreads; creg[q] = (colsCached && k < c.ncols) ? __ldg(cols + k) : -1; }
   auto sweep = [&](const double* src) {
      SWEEP_T0();
      if (colsCached) {
         double v[kResCols];
#pragma unroll
         for (int q = 0; q < kResCols; q++) v[q] = creg[q] >= 0 ? __ldcg(src + creg[q]) : 0.0;
#pragma unroll
         for (int q = 0; q < kResCols; q++) if (creg[q] >= 0) win[threadIdx.x + q * kResThreads] = v[q];
      } else {
         for (int k = threadIdx.x; k < c.ncols; k += blockDim.x) win[k] = __ldcg(src + __ldg(cols + k));
      }
      __syncthreads();
      SWEEP_TICK(0);
      for (int g = warp; g < c.ngroups; g += kResWarps) {
         const bool first = g == warp;
         const int off = first ? off0 : grp[g];
         const int np = first ? np0 : (grp[g + 1] - off) >> 6;
         const double2* vp = reinterpret_cast<const double2*>(zs + off) + lane;
         const unsigned* __restrict__ ip = reinterpret_cast<const unsigned*>(idx + off) + lane;
         double acc = 0.0;
         for (int base = 0; base < np; base += kResCache) {
            unsigned iq[kResCache];
            if (first && base == 0) {
#pragma unroll
               for (int q = 0; q < kResCache; q++) iq[q] = ireg[q];
            } else {
#pragma unroll
               for (int q = 0; q < kResCache; q++) iq[q] = base + q < np ? __ldg(ip + (base + q) * 32) : 0u;
            }
#pragma unroll
            for (int q = 0; q < kResCache; q++)
               if (base + q < np) {
                  const double2 v2 = vp[(base + q) * 32];
                  acc += v2.x * win[iq[q] & 0xffffu] + v2.y * win[iq[q] >> 16];
               }
         }
         for (int o = T >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
         const int r = g * R + lane / T;
         if ((lane % T) == 0 && r < c.nrows) wrow[r] = acc;
      }
      SWEEP_TICK(1);
      __syncthreads();
      SWEEP_TICK(2);
   };
   const bool own = (int)threadIdx.x < c.nrows;
   const int row = c.row0 + (int)threadIdx.x;
   double xi = 0.0, ri = 0.0, pi = 0.0, si = 0.0, ui = 0.0, wi = 0.0, di = 0.0, bi = 0.0;
   if (own) { bi = a.b[row]; di = a.dinv[row]; }
   if (!a.x0zero) {   // warm start: r = b - Z x
      if (own) xi = a.x[row];
      sweep(a.x);
      if (own) ri = bi - wrow[threadIdx.x];
   } else {
      ri = bi;
   }
   ui = di * ri;
   if (own) a.u[row] = ui;
   double v4[4] = {ri * ui, 0.0, ri * ri, bi * bi};
   grid.sync();
   sweep(a.u);
   if (own) wi = wrow[threadIdx.x];
   v4[1] = ui * wi;
   grid_sum<4>(grid, v4, buf0, sh);
   double gamma = v4[0], rr = v4[2];
   const double bb = v4[3];
   const double thresh = a.tol * sqrt(bb);
   double alpha = gamma / v4[1], beta = 0.0;
   int it = 0;
#ifdef CFD_PCG_PHASE_TIMERS
   unsigned long long tph[4] = {0, 0, 0, 0}, t0 = 0, t1;
#define PCG_TICK(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); tph[k] += t1 - t0; t0 = t1; } } while (0)
   if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
#else
#define PCG_TICK(k)
#endif
   while (true) {
      if (sqrt(rr) <= thresh || it >= a.maxit) break;
      pi = ui + beta * pi;
      si = wi + beta * si;
      xi += alpha * pi;
      ri -= alpha * si;
      ui = di * ri;
      if (own) a.u[row] = ui;
      double v3[3] = {ri * ui, 0.0, ri * ri};
      it++;
      PCG_TICK(0);
      grid.sync();
      PCG_TICK(1);
      sweep(a.u);
      if (own) wi = wrow[threadIdx.x];
      v3[1] = ui * wi;
      PCG_TICK(2);
      grid_sum<3>(grid, v3, buf0, sh);
      PCG_TICK(3);
      beta = v3[0] / gamma;
      alpha = v3[0] / (v3[1] - beta * v3[0] / alpha);
      gamma = v3[0];
      rr = v3[2];
   }
#undef PCG_TICK
   if (own) a.x[row] = xi;
   if (blockIdx.x == 0 && threadIdx.x == 0) {
      a.info[0] = (double)it; a.info[1] = rr; a.info[2] = bb;
#ifdef CFD_PCG_PHASE_TIMERS
      for (int k = 0; k < 4; k++) a.info[8 + k] = (double)tph[k];
      for (int k = 0; k < 3; k++) a.info[24 + k] = (double)tsw[k];   // inside the sweep: staging | warp 0's groups | wait for the CTA's other warps
#endif
   }
}

struct PcgP2PArgs {
   PcgArgs a;   // a.p lives in this rank's exchange window
   XComm xc;
};

// Multi-GPU Jacobi-PCG, classical recurrences: k_pcg with the two exchange steps folded in (fused compute + collective over
// peer memory).  One persistent cooperative kernel per GPU; per iteration: q = Z p on the owned rows (ghost p entries were
// written by the neighbours), p.q all-reduced through mailboxes, fused x/r update, (r.z, r.r) all-reduced, p update whose
// interface values are stored directly into the neighbours' ghost entries over NVLink.  Identical recurrences and
// stopping rule on every rank (the reduced scalars are bit-identical everywhere), so all ranks leave together.
// (cfd_options.cg_variant = 1 on partitioned runs; the default there is the single-reduction kernel k_pcg_cg<true> below.)
__global__ void __launch_bounds__(512, 3) k_pcg_p2p(PcgP2PArgs pa) {
   cg::grid_group grid = cg::this_grid();
   __shared__ double sh[3 * 32];
   const PcgArgs& a = pa.a;
   const XComm& xc = pa.xc;
   const int tid = blockIdx.x * blockDim.x + threadIdx.x;
   const int nthreads = gridDim.x * blockDim.x;
   const int gwarp = tid >> 5, nwarps = nthreads >> 5, lane = threadIdx.x & 31;
   double* buf0 = a.scratch;
   double* buf1 = a.scratch + 3 * (size_t)gridDim.x;
   const int n = a.n;
   unsigned long long epoch = xc.xstate[0], hepoch = xc.xstate[1];

   if (!a.x0zero) {   // ghost entries of x were refreshed by the caller
      pcg_spmv_phase(a.Z, a.part, a.dynStart, a.dynCounter, a.x, a.q, gwarp, nwarps, lane);
      grid.sync();
      if (tid == 0 && a.dynCounter) *a.dynCounter = 0;
   }
   double v3[3] = {0, 0, 0};
   for (int i = tid; i < n; i += nthreads) {
      const double bi = a.b[i];
      double ri = bi;
      if (a.x0zero) a.x[i] = 0.0; else ri = bi - a.q[i];
      a.r[i] = ri;
      const double zi = a.dinv[i] * ri;
      v3[0] += ri * zi; v3[1] += bi * bi; v3[2] += ri * ri;
   }
   if (!xgrid_sum<3>(v3, buf0, sh, xc, epoch)) return;
   double rz = v3[0];
   const double bb = v3[1];
   double rr = v3[2];
   const double thresh = a.tol * sqrt(bb);
   int it = 0;
   if (!(sqrt(rr) <= thresh || it >= a.maxit))
      if (!p_update_and_push(grid, n, a.p, xc, hepoch, tid, nthreads, [&](int i) { return a.dinv[i] * a.r[i]; })) return;
   // phase clocks (diagnostics build: info[8..12] = ns in spmv / all-reduce 1 / update / all-reduce 2 / p update + push + flags), of
   // block 0 and -- info[13..17] -- of the last block
#ifdef CFD_PCG_PHASE_TIMERS
   unsigned long long tph[5] = {0, 0, 0, 0, 0}, t0 = 0, t1;
   const bool clk = threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);
#define PCG_TICK(k) do { if (clk) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); tph[k] += t1 - t0; t0 = t1; } } while (0)
   if (clk) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
#else
#define PCG_TICK(k)
#endif
   while (true) {
      if (sqrt(rr) <= thresh || it >= a.maxit) break;
      double v1[1];
      v1[0] = pcg_spmv_phase(a.Z, a.part, a.dynStart, a.dynCounter, a.p, a.q, gwarp, nwarps, lane);
      PCG_TICK(0);
      if (!xgrid_sum<1>(v1, buf1, sh, xc, epoch)) return;
      if (tid == 0 && a.dynCounter) *a.dynCounter = 0;
      PCG_TICK(1);
      const double alpha = rz / v1[0];
      double v2[2] = {0, 0};
      for (int i = tid; i < n; i += nthreads) {
         a.x[i] += alpha * a.p[i];
         const double ri = a.r[i] - alpha * a.q[i];
         a.r[i] = ri;
         v2[0] += ri * (a.dinv[i] * ri);
         v2[1] += ri * ri;
      }
      PCG_TICK(2);
      if (!xgrid_sum<2>(v2, buf0, sh, xc, epoch)) return;
      PCG_TICK(3);
      const double beta = v2[0] / rz;
      rz = v2[0];
      rr = v2[1];
      it++;
      if (sqrt(rr) <= thresh || it >= a.maxit) break;   // converged: the next search direction is never used
      if (!p_update_and_push(grid, n, a.p, xc, hepoch, tid, nthreads, [&](int i) { return a.dinv[i] * a.r[i] + beta * a.p[i]; })) return;
      PCG_TICK(4);
   }
#undef PCG_TICK
#ifdef CFD_PCG_PHASE_TIMERS
   if (clk)
      for (int k = 0; k < 5; k++) a.info[(blockIdx.x == 0 ? 8 : 13) + k] = (double)tph[k];
#endif
   grid.sync();
   if (tid == 0) {
      a.info[0] = (double)it; a.info[1] = rr; a.info[2] = bb;
      xc.xstate[0] = epoch; xc.xstate[1] = hepoch;
      xc.xstate[3] = 0; xc.xstate[4] = 0;   // tickets count modulo the grid size, which differs between the Krylov kernels
   }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Jacobi-PCG in the single-reduction form of Chronopoulos & Gear: the same iterates as the classical recurrences in exact
// arithmetic, but ONE grid-wide (DIST: cross-GPU) reduction per iteration instead of two, and two synchronisation points
// instead of three.  With u = M^-1 r, w = Z u:
//    p = u + beta p ;  s = w + beta s (= Z p) ;  x += alpha p ;  r -= alpha s ;  u = M^-1 r           (one vector pass)
//    w = Z u ;  gamma = r.u , delta = w.u , |r|^2                                                        (one sweep + ONE reduction)
//    beta = gamma / gamma_old ;  alpha = gamma / (delta - beta gamma / alpha_old)
// Stopping rule and iteration count as in k_pcg (cusp::default_monitor): |r| <= tol |b| tested before each iteration.  The
// price is 16 bytes per row and iteration of extra vector traffic and one more sweep at the end, so the library uses it
// where synchronisation dominates: small (L2-resident) systems and partitioned runs (choose_cg_variant).
//
// DIST = true: fused compute + collective over peer memory.  u lives in this rank's exchange window; the vector pass does
// the rows some neighbour holds as ghosts FIRST and stores their new u straight into the neighbours' windows over NVLink,
// the last block to finish them raises the neighbours' halo flags, and only then come the interior rows -- the transfer
// and the flag flight overlap the bulk of the pass.  The one reduction per iteration is the fused cross-GPU barrier
// xgrid_sum.  No NCCL call and no host round trip inside a solve.
struct PcgCgArgs {
   Sell Z;
   const int* part;
   int dynStart;
   int* dynCounter;
   int n;
   const double* b;
   double* x;
   double *r, *u, *w, *p, *s;
   const double* dinv;
   double* scratch;   // 4 x gridDim doubles
   double* info;      // [0]=iterations [1]=|r|^2 [2]=|b|^2
   double tol;
   int maxit;
   int x0zero;
   XComm xc;          // DIST only
};

// The sweep w = Z u with this lane's share of delta = w.u is exactly the classical kernel's q = Z p with p.q (pcg_spmv_phase):
// gamma = r.u and |r|^2 are accumulated where r and u are produced (the vector pass) and ride along to the one reduction.  (A first
// version read r and u again in the sweep's epilogue: two more dependent loads per slice of a warp that works through its slices
// one after the other -- the sweep ran at 0.55 of the classical kernel's rate.)

// The vector pass over rows i = f(t), t in [t0, t1): `first` = no previous search direction (beta = 0: p = u, s = w exactly).
template <bool DIST, typename RowOf>
__device__ __forceinline__ void pcgcg_rows(const PcgCgArgs& a, int t0, int t1, int stride, bool first, double alpha, double beta, RowOf rowOf,
                                           double& g, double& rr) {
   for (int t = t0; t < t1; t += stride) {
      const int i = rowOf(t);
      if (i < 0) continue;
      const double ui = a.u[i], wi = a.w[i];
      const double pi = first ? ui : ui + beta * a.p[i];
      const double si = first ? wi : wi + beta * a.s[i];
      a.p[i] = pi; a.s[i] = si;
      a.x[i] += alpha * pi;
      const double ri = a.r[i] - alpha * si;
      a.r[i] = ri;
      const double un = a.dinv[i] * ri;
      a.u[i] = un;
      g += ri * un; rr += ri * ri;
      if (DIST) {
         bool pushed = false;
         for (int e = a.xc.pushPtr[i]; e < a.xc.pushPtr[i + 1]; e++) {
            const int2 dst = a.xc.pushDst[e];
            st_sys_f64(a.xc.peerP[dst.x] + dst.y, un);
            pushed = true;
         }
         if (pushed) __threadfence_system();   // system-scope fences are expensive: only the storing threads issue one
      }
   }
}

// After this rank's interface values have been stored into the neighbours' windows: the last block to get here raises the halo flags.
__device__ __forceinline__ void pcgcg_signal(const XComm& xc, unsigned long long hepoch) {
   __shared__ int s_lastI;
   __threadfence();   // (the threads that stored into a neighbour's window have fenced those stores at system scope themselves)
   __syncthreads();
   if (threadIdx.x == 0) {
      const unsigned long long G = gridDim.x;
      s_lastI = (atomicAdd(xc.xstate + 4, 1ull) % G == G - 1) ? 1 : 0;
   }
   __syncthreads();
   if (s_lastI && threadIdx.x < xc.npeers) {
      __threadfence_system();
      st_rel_sys(win_haloflag(xc.peerWin[xc.peerRank[threadIdx.x]], xc.rank), hepoch);
   }
}
__device__ __forceinline__ bool pcgcg_wait(const XComm& xc, unsigned long long hepoch) {
   bool ok = true;
   if (threadIdx.x < xc.npeers) ok = wait_flag(win_haloflag(xc.peerWin[xc.rank], xc.peerRank[threadIdx.x]), hepoch, xc.xstate + 2, xc.timeoutNs);
   return __syncthreads_and(ok ? 1 : 0) != 0;
}

template <bool DIST, int MINB>
__global__ void __launch_bounds__(512, MINB) k_pcg_cg(PcgCgArgs a) {
   cg::grid_group grid = cg::this_grid();
   __shared__ double sh[4 * 32];
   const int tid = blockIdx.x * blockDim.x + threadIdx.x;
   const int nthreads = gridDim.x * blockDim.x;
   const int gwarp = tid >> 5, nwarps = nthreads >> 5, lane = threadIdx.x & 31;
   const int n = a.n;
   unsigned long long epoch = 0, hepoch = 0;
   if (DIST) { epoch = a.xc.xstate[0]; hepoch = a.xc.xstate[1]; }

   if (!a.x0zero) {
      // r = b - Z x (warm start, BCH.cu:418-451 keeps the previous Pdot; DIST: ghost entries of x were refreshed by the caller)
      pcg_spmv_phase(a.Z, a.part, a.dynStart, a.dynCounter, a.x, a.w, gwarp, nwarps, lane);
      grid.sync();
      if (tid == 0 && a.dynCounter) *a.dynCounter = 0;
   }
   double v4[4] = {0, 0, 0, 0};
   // u = M^-1 r: interface rows first (DIST), then everything else
   auto initRow = [&](int i) {
      const double bi = a.b[i];
      double ri = bi;
      if (a.x0zero) a.x[i] = 0.0; else ri = bi - a.w[i];
      a.r[i] = ri;
      const double un = a.dinv[i] * ri;
      a.u[i] = un;
      v4[0] += ri * un; v4[2] += ri * ri; v4[3] += bi * bi;
      if (DIST) {
         bool pushed = false;
         for (int e = a.xc.pushPtr[i]; e < a.xc.pushPtr[i + 1]; e++) {
            const int2 dst = a.xc.pushDst[e];
            st_sys_f64(a.xc.peerP[dst.x] + dst.y, un);
            pushed = true;
         }
         if (pushed) __threadfence_system();
      }
   };
   if (DIST) {
      for (int t = tid; t < a.xc.nIface; t += nthreads) initRow(a.xc.ifaceRows[t]);
      hepoch++;
      pcgcg_signal(a.xc, hepoch);
      for (int i = tid; i < n; i += nthreads)
         if (a.xc.pushPtr[i + 1] == a.xc.pushPtr[i]) initRow(i);
   } else {
      for (int i = tid; i < n; i += nthreads) initRow(i);
   }
   grid.sync();
   if (DIST && !pcgcg_wait(a.xc, hepoch)) return;
   v4[1] = pcg_spmv_phase(a.Z, a.part, a.dynStart, a.dynCounter, a.u, a.w, gwarp, nwarps, lane);
   if (DIST) { if (!xgrid_sum<4>(v4, a.scratch, sh, a.xc, epoch)) return; }
   else grid_sum<4>(grid, v4, a.scratch, sh);
   if (tid == 0 && a.dynCounter) *a.dynCounter = 0;
   double gamma = v4[0], rr = v4[2];
   const double bb = v4[3];
   const double thresh = a.tol * sqrt(bb);
   double alpha = gamma / v4[1], beta = 0.0;
   int it = 0;
   // phase clock of block 0 (diagnostics build: info[8..11] = ns in vector pass / grid sync (+ halo wait) / sweep / reduction)
#ifdef CFD_PCG_PHASE_TIMERS
   unsigned long long tph[4] = {0, 0, 0, 0}, t0 = 0, t1;
#define PCG_TICK(k) do { if (tid == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); tph[k] += t1 - t0; t0 = t1; } } while (0)
   if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
#else
#define PCG_TICK(k)
#endif
   while (true) {
      if (sqrt(rr) <= thresh || it >= a.maxit) break;
      const bool first = it == 0;
      double v3[3] = {0, 0, 0};
      if (DIST) {
         pcgcg_rows<true>(a, tid, a.xc.nIface, nthreads, first, alpha, beta, [&](int t) { return a.xc.ifaceRows[t]; }, v3[0], v3[2]);
         hepoch++;
         pcgcg_signal(a.xc, hepoch);
         pcgcg_rows<false>(a, tid, n, nthreads, first, alpha, beta, [&](int t) { return a.xc.pushPtr[t + 1] == a.xc.pushPtr[t] ? t : -1; }, v3[0], v3[2]);
      } else {
         pcgcg_rows<false>(a, tid, n, nthreads, first, alpha, beta, [](int t) { return t; }, v3[0], v3[2]);
      }
      it++;
      PCG_TICK(0);
      grid.sync();   // every block's u is in place (DIST: and the neighbours' interface values, once their flags are up)
      if (DIST && !pcgcg_wait(a.xc, hepoch)) return;
      PCG_TICK(1);
      v3[1] = pcg_spmv_phase(a.Z, a.part, a.dynStart, a.dynCounter, a.u, a.w, gwarp, nwarps, lane);
      PCG_TICK(2);
      if (DIST) { if (!xgrid_sum<3>(v3, a.scratch, sh, a.xc, epoch)) return; }
      else grid_sum<3>(grid, v3, a.scratch, sh);
      PCG_TICK(3);
      if (tid == 0 && a.dynCounter) *a.dynCounter = 0;   // every block is past the sweep; the next one comes after grid.sync()
      beta = v3[0] / gamma;
      alpha = v3[0] / (v3[1] - beta * v3[0] / alpha);
      gamma = v3[0];
      rr = v3[2];
   }
#undef PCG_TICK
#ifdef CFD_PCG_PHASE_TIMERS
   if (tid == 0)
      for (int k = 0; k < 4; k++) a.info[8 + k] = (double)tph[k];
#endif
   if (tid == 0) {
      a.info[0] = (double)it; a.info[1] = rr; a.info[2] = bb;
      if (DIST) { a.xc.xstate[0] = epoch; a.xc.xstate[1] = hepoch; a.xc.xstate[3] = 0; a.xc.xstate[4] = 0; }
   }
}

// Diagnostic: cost of the grid-wide primitives the persistent Krylov kernels are built from.
// mode 0: grid.sync()   mode 1: grid_sum<1>   mode 2: grid_sum<3>
__global__ void __launch_bounds__(1024) k_diag_sync(int iters, int mode, double* scratch, double* out) {
   cg::grid_group grid = cg::this_grid();
   __shared__ double sh[3 * 32];
   double acc = 0.0;
   for (int it = 0; it < iters; it++) {
      if (mode == 0) {
         grid.sync();
      } else if (mode == 1) {
         double v[1] = {1.0 + threadIdx.x};
         grid_sum<1>(grid, v, scratch + (it & 1) * 3 * (size_t)gridDim.x, sh);
         acc += v[0];
      } else {
         double v[3] = {1.0, 2.0, 3.0};
         grid_sum<3>(grid, v, scratch + (it & 1) * 3 * (size_t)gridDim.x, sh);
         acc += v[0] + v[1] + v[2];
      }
   }
   if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = acc;
}

#ifndef CFD_BICG_UNROLL1
#define CFD_BICG_UNROLL1 8
#endif
#ifndef CFD_BICG_UNROLL2
#define CFD_BICG_UNROLL2 6
#endif
// peer-memory exchange of the velocity-row vectors the momentum SpMV gathers (partitioned runs)
struct XVel {
   double* const* peerV;   // [npeers] start of MY block inside the peer's ghost range of its ph region (shv region: + 3 * peerLd)
   const int* peerLd;      // [npeers] the peer's component stride (its local NN)
   const int* pushPtr;     // [n+1] CSR of (peer index, slot) pushes per owned velocity node
   const int2* pushDst;
};
struct BicgArgs {
   Sell A;
   const int* part;   // per-block slice ranges of the static share
   int dynStart;      // first slice handed out through dynCounter (0 = everything dynamic)
   int* dynCounter;   // NULL = everything static
   int vidx;
   int n;             // rows (owned velocity nodes)
   long long ld;      // component stride of every vector below (local velocity nodes, owned + ghost)
   const double* b;   // [3][ld]
   double* x;         // [3][ld]
   double *r, *rs, *p, *ph, *AMp, *shv, *AMs, *xbest;   // work vectors [3][ld]; DIST: ph and shv live in the exchange window
   const double* dinv;
   double* scratch;   // 3 buffers x 6 x gridDim
   double* info;      // see kBicgInfo*
   double tol;
   int maxit;
   int halfstep;      // 1 = also test the residual after the half step (cusp::krylov::bicgstab); 0 = once per iteration (PARALUTION)
   int verify;        // 1 = accept a solve only on its TRUE residual |b - A x|; restart from it otherwise
   XComm xc;          // DIST only
   XVel xv;           // DIST only
};
// result slots inside cfd_ctx::kinfo (doubles)
enum { kBicgInfoIts = 3, kBicgInfoRR = 6, kBicgInfoTrue = 9, kBicgInfoBB = 12, kBicgInfoFin = 15, kBicgInfoRestarts = 18, kBicgInfoPasses = 21,
       kBicgInfoTimers = 32 };

// out_s = A * in_s for the NA still-active systems sys[0..NA) (multi-RHS: one pass over A), fused with
// the two dot products every BiCGStab SpMV is followed by: d0[s] += l1_s . out_s, d1[s] += out_s . out_s.
// RES = true: residual form, out_s = l1_s - A * in_s (l1 = right-hand side), d1[s] += |out_s|^2  (true-residual check / restart).
// Specialised on NA so that the common late phase (one slow system left) runs a lean single-vector loop
// with four matrix pairs in flight per lane.
template <int NA, bool RES>
__device__ __forceinline__ void bicg_spmv_na(const Sell& A, const int* part, int dynStart, int* dynCounter, const double* __restrict__ val, int n, long long ld,
                                             const double* in, double* out, const int (&sys)[3], const double* l1, double (&d0)[3], double (&d1)[3],
                                             int gwarp, int nwarps, int lane) {
   const int T = A.T, sub = lane % T;
   // Static share [0, nstatic) dealt as in slice_range; the rest is taken one slice at a time from a shared counter, so that
   // warps whose slices were cheap (rows of a Q2 mesh differ 4.6x in length, and in how well their gathers hit L1) keep
   // working instead of waiting at the grid reduction: measured on the 30^3 Taylor-Hood cavity, block 0 used to wait 22 us
   // after a 29 us SpMV share.  `dynCounter` must be 0 on entry; the kernel resets it after the reduction that follows.
   // (With the counter the dot-product partials depend on which warp drew which slice: results are reproducible to rounding,
   // not bitwise; cfd_options.deterministic = 1 turns the counter off.)
   const int nstatic = dynCounter ? dynStart : A.nslices;
   const SliceRange sr = slice_range(nstatic, part, gwarp, nwarps);
   size_t base[NA];
   double s0[NA], s1[NA];   // dot-product partials stay in registers (no dynamically indexed local arrays in the loop)
#pragma unroll
   for (int q = 0; q < NA; q++) { base[q] = (size_t)sys[q] * (size_t)ld; s0[q] = 0.0; s1[q] = 0.0; }
   int next = sr.begin;
   while (true) {
      int s;
      if (next < sr.end) { s = next; next += sr.step; }
      else if (dynCounter) {
         int g = 0;
         if (lane == 0) g = atomicAdd(dynCounter, 1);
         s = __shfl_sync(0xffffffffu, g, 0) + dynStart;
         if (s >= A.nslices) break;
      } else break;
      double acc[NA];
#pragma unroll
      for (int q = 0; q < NA; q++) acc[q] = 0.0;
      // a warp works through its slices one after the other, so the sweep is latency-bound per warp at 2 CTAs per SM: the fewer
      // right-hand sides are left, the more matrix pairs are kept in flight per lane (CFD_BICG_UNROLL1 / 2: tuning override)
      sell_pairs<(NA == 1 ? CFD_BICG_UNROLL1 : (NA == 2 ? CFD_BICG_UNROLL2 : 4))>(A, s, lane, [&](long long e, int2 c2) {
         const double2 v2 = ldnc_d2(val + e);
#pragma unroll
         for (int q = 0; q < NA; q++) acc[q] += v2.x * in[base[q] + c2.x] + v2.y * in[base[q] + c2.y];
      });
#pragma unroll
      for (int q = 0; q < NA; q++) acc[q] = lanes_sum(acc[q], T);
      const int row = s * A.R + lane / T;
      if (sub == 0 && row < n) {
#pragma unroll
         for (int q = 0; q < NA; q++) {
            const size_t i = base[q] + row;
            if (RES) {
               const double ri = l1[i] - acc[q];
               out[i] = ri;
               s1[q] += ri * ri;
            } else {
               out[i] = acc[q];
               s0[q] += l1[i] * acc[q];
               s1[q] += acc[q] * acc[q];
            }
         }
      }
   }
#pragma unroll
   for (int q = 0; q < NA; q++) {
      const int c = sys[q];   // uniform; resolved with selects, not local-memory indexing
      if (c == 0) { d0[0] += s0[q]; d1[0] += s1[q]; }
      else if (c == 1) { d0[1] += s0[q]; d1[1] += s1[q]; }
      else { d0[2] += s0[q]; d1[2] += s1[q]; }
   }
}
template <bool RES>
__device__ __forceinline__ void bicg_spmv_phase(const BicgArgs& a, const double* in, double* out, const bool (&act)[3], const double* l1,
                                                double (&d0)[3], double (&d1)[3], int gwarp, int nwarps, int lane) {
   const double* __restrict__ val = sell_vals(a.A, a.vidx);
   int sys[3] = {0, 0, 0}, na = 0;
   for (int c = 0; c < 3; c++)
      if (act[c]) sys[na++] = c;
   if (na == 1) bicg_spmv_na<1, RES>(a.A, a.part, a.dynStart, a.dynCounter, val, a.n, a.ld, in, out, sys, l1, d0, d1, gwarp, nwarps, lane);
   else if (na == 2) bicg_spmv_na<2, RES>(a.A, a.part, a.dynStart, a.dynCounter, val, a.n, a.ld, in, out, sys, l1, d0, d1, gwarp, nwarps, lane);
   else if (na == 3) bicg_spmv_na<3, RES>(a.A, a.part, a.dynStart, a.dynCounter, val, a.n, a.ld, in, out, sys, l1, d0, d1, gwarp, nwarps, lane);
}

// Jacobi-preconditioned BiCGStab for the three momentum systems in lock-step (replaces 3 calls of
// CUSP_GMRES_solver -> cusp::krylov::bicgstab, GQ.cu:369-433, x0 = 0; CPU build: paralution_BiCGStab_solver, GQ.cpp:4183-4231).
// The three systems share A, so each of the two SpMV phases per iteration streams A once for all still-active right-hand
// sides.  Per-system scalars are block-uniform and live in shared memory (written by thread 0 after each grid reduction),
// which keeps the register budget for the SpMV sweep.
//
// Acceptance is on the TRUE residual.  When the recurrence residual of a system meets the stopping rule (at the top of an
// iteration or -- CUSP's rule -- after the half step), when it breaks down (alpha, omega or rho not finite / zero) or when
// it runs out of iterations, r = b - A x is recomputed.  If |r| <= tol |b| the system is done and flagged converged.
// Otherwise it is RESTARTED from that residual (new shadow residual rs = r, p = r), at most kBicgMaxRestarts times and
// only while restarts still gain; a system that stops unconverged returns the iterate with the smallest true residual it
// has verified, flags converged = 0 and reports that residual (cfd_step_report.bicg_true_residual).  In the common case
// (the recurrence residual is honest) this costs one extra SpMV per system and changes neither iterates nor counts.
// Round 1 accepted whatever iterate existed at a breakdown: on the reference's under-integrated Q2 momentum matrix at
// 1e-12 that silently returned solves with a 1e-9 residual, i.e. 1e-5 in the fields.
enum { kBicgMaxRestarts = 12 };
struct BicgScalars {
   double bb[3], rr[3], rrs[3], thresh[3], alpha[3], omega[3], trueRR[3], bestRR[3], prevTrue[3];
   int its[3], act[3], half[3], act2[3], chk[3], brk[3], fin[3], restarts[3], hasBest[3];
   int todo[3];   // after a check: 0 nothing, 1 restart, 3 restart + remember x as the best iterate, 2 give back the best iterate, 4 give back x = 0
   int passes;
};

// DIST: store a freshly computed entry of ph (which = 0) or shv (which = 1) into the ghost copies the neighbours hold
__device__ __forceinline__ void bicg_push(const BicgArgs& a, int which, int c, int i, double v) {
   bool pushed = false;
   for (int e = a.xv.pushPtr[i]; e < a.xv.pushPtr[i + 1]; e++) {
      const int2 d = a.xv.pushDst[e];
      st_sys_f64(a.xv.peerV[d.x] + (size_t)(which * 3 + c) * (size_t)a.xv.peerLd[d.x] + d.y, v);
      pushed = true;
   }
   if (pushed) __threadfence_system();   // only the storing threads pay for the system-scope fence
}

// DIST = true (partitioned runs): the same kernel as a fused compute + collective over peer memory.  ph and shv -- the vectors
// the two SpMV phases gather -- live in this rank's exchange window; the passes that produce them store interface values
// straight into the neighbours' ghost entries over NVLink, every reduction is the fused cross-GPU barrier xgrid_sum (which
// also publishes the pushes made before it), and the one hand-over without a reduction (after the p update) is a flag
// exchange.  All ranks see bit-identical scalars, take the same branches and leave together.  No NCCL call and no host
// round trip inside the solve (round 1: 4 ncclAllReduce + 2 halo swaps + 7 launches per iteration, driven from the host).
template <bool DIST>
__global__ void __launch_bounds__(512, 2) k_bicgstab3(BicgArgs a) {
   cg::grid_group grid = cg::this_grid();
   __shared__ double sh[6 * 32];
   __shared__ BicgScalars S;
   const int tid = blockIdx.x * blockDim.x + threadIdx.x;
   const int nthreads = gridDim.x * blockDim.x;
   const int gwarp = tid >> 5, nwarps = nthreads >> 5, lane = threadIdx.x & 31;
   double* buf0 = a.scratch;
   double* buf1 = a.scratch + 6 * (size_t)gridDim.x;
   double* buf2 = a.scratch + 12 * (size_t)gridDim.x;   // the true-residual check has its own (it may run between any two others)
   const int n = a.n;
   const size_t ld = (size_t)a.ld;
   double *r = a.r, *rs = a.rs, *p = a.p, *ph = a.ph, *AMp = a.AMp, *shv = a.shv, *AMs = a.AMs, *xbest = a.xbest;
   unsigned long long epoch = 0, hepoch = 0;
   if (DIST) { epoch = a.xc.xstate[0]; hepoch = a.xc.xstate[1]; }
   // grid-wide (DIST: cross-GPU) sum; false = a peer timed out: leave at once
#define BICG_SUM(NV, v, buf) (DIST ? xgrid_sum<NV>(v, buf, sh, a.xc, epoch) : (grid_sum<NV>(grid, v, buf, sh), true))
   // ... that also hands over the peer-memory stores made in the pass before it
#define BICG_SUM_PUBLISH(NV, v, buf) (DIST ? xgrid_sum<NV, true>(v, buf, sh, a.xc, epoch) : (grid_sum<NV>(grid, v, buf, sh), true))
   // hand-over of pushed values without a reduction: flags + local grid sync
#define BICG_HANDOVER()                                   \
   do {                                                   \
      if (DIST) {                                         \
         hepoch++;                                        \
         pcgcg_signal(a.xc, hepoch);                      \
         grid.sync();                                     \
         if (!pcgcg_wait(a.xc, hepoch)) return;           \
      } else grid.sync();                                 \
   } while (0)

   double v6[6] = {0, 0, 0, 0, 0, 0}, v3[3];
   for (int c = 0; c < 3; c++) {
      double acc = 0.0;
      for (int i = tid; i < n; i += nthreads) {
         const size_t k = (size_t)c * ld + i;
         const double bi = a.b[k];
         a.x[k] = 0.0;
         r[k] = bi; rs[k] = bi; p[k] = bi;
         const double phi = a.dinv[i] * bi;
         ph[k] = phi;
         if (DIST) bicg_push(a, 0, c, i, phi);
         acc += bi * bi;
      }
      v6[c] = acc;
   }
   if (!BICG_SUM_PUBLISH(6, v6, buf0)) return;
   if (threadIdx.x == 0) {
      for (int c = 0; c < 3; c++) {
         S.bb[c] = S.rr[c] = S.rrs[c] = v6[c];
         S.thresh[c] = a.tol * sqrt(v6[c]);
         S.its[c] = 0; S.half[c] = 0; S.act2[c] = 0; S.alpha[c] = 0.0; S.omega[c] = 0.0;
         S.chk[c] = 0; S.brk[c] = 0; S.restarts[c] = 0; S.todo[c] = 0; S.hasBest[c] = 0;
         S.trueRR[c] = v6[c]; S.bestRR[c] = v6[c]; S.prevTrue[c] = v6[c];
         // b = 0: x = 0 is exact (first steps of a cavity: v = w = 0)
         S.act[c] = v6[c] > 0.0 ? 1 : 0;
         S.fin[c] = v6[c] > 0.0 ? 0 : 1;
      }
      S.passes = 0;
   }
   __syncthreads();
   // phase clock of block 0 (diagnostics build only: info[kBicgInfoTimers + 0..8] = ns in spmv1 / reduce / s-update / reduce / spmv2 /
   // reduce / x,r-update / reduce / p-update + grid sync, [+9] = lock-step iterations)
#ifdef CFD_BICG_PHASE_TIMERS
   unsigned long long tph[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, t0 = 0, t1, nit = 0;
#define BICG_TICK(k) do { if (tid == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); tph[k] += t1 - t0; t0 = t1; } } while (0)
   if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
#else
#define BICG_TICK(k)
#endif
   while (true) {
      // ---- stopping rule on the recurrence residual (cusp::default_monitor: tested before each iteration), breakdowns, iteration limit
      if (threadIdx.x == 0)
         for (int c = 0; c < 3; c++) {
            S.chk[c] = 0; S.todo[c] = 0;
            if (S.act[c] && (S.brk[c] || sqrt(S.rr[c]) <= S.thresh[c] || S.its[c] >= a.maxit)) {
               if (a.verify) S.chk[c] = 1;
               else { S.act[c] = 0; S.fin[c] = (!S.brk[c] && sqrt(S.rr[c]) <= S.thresh[c]) ? 1 : 2; }
            }
         }
      __syncthreads();
      if (S.chk[0] | S.chk[1] | S.chk[2]) {
         // ---- true residual of the candidates: AMs = b - A x, |AMs|^2
         if (DIST) {   // x of the candidates goes through shv (free at this point) so that the neighbours' values can be gathered
            for (int c = 0; c < 3; c++) {
               if (!S.chk[c]) continue;
               for (int i = tid; i < n; i += nthreads) {
                  const size_t k = (size_t)c * ld + i;
                  const double xi = a.x[k];
                  shv[k] = xi;
                  bicg_push(a, 1, c, i, xi);
               }
            }
            BICG_HANDOVER();
         }
         double d0[3] = {0, 0, 0}, d1[3] = {0, 0, 0};
         const bool chk[3] = {S.chk[0] != 0, S.chk[1] != 0, S.chk[2] != 0};
         bicg_spmv_phase<true>(a, DIST ? shv : a.x, AMs, chk, a.b, d0, d1, gwarp, nwarps, lane);
         v3[0] = d1[0]; v3[1] = d1[1]; v3[2] = d1[2];
         if (!BICG_SUM(3, v3, buf2)) return;
         if (tid == 0 && a.dynCounter) *a.dynCounter = 0;
         if (threadIdx.x == 0) {
            S.passes++;
            for (int c = 0; c < 3; c++) {
               if (!S.chk[c]) continue;
               const double tr = v3[c];
               S.trueRR[c] = tr;
               const bool improved = tr < S.bestRR[c];
               if (sqrt(tr) <= S.thresh[c]) { S.act[c] = 0; S.fin[c] = 1; S.bestRR[c] = tr; }
               else if (S.its[c] >= a.maxit || S.restarts[c] >= kBicgMaxRestarts || !isfinite(tr) ||
                        (S.restarts[c] > 0 && !(tr < 0.25 * S.prevTrue[c]))) {
                  // out of iterations, or restarts no longer gain (attainable accuracy reached): stop, unconverged, best iterate
                  S.act[c] = 0; S.fin[c] = 2;
                  if (!improved && S.hasBest[c]) { S.todo[c] = 2; S.trueRR[c] = S.bestRR[c]; }
                  else if (!improved) { S.todo[c] = 4; S.trueRR[c] = S.bb[c]; }   // nothing verified beats the initial guess: return x = 0
                  else S.bestRR[c] = tr;
               } else {
                  S.todo[c] = improved ? 3 : 1;   // restart (3: and remember x as the best iterate so far)
                  if (improved) { S.bestRR[c] = tr; S.hasBest[c] = 1; }
                  S.prevTrue[c] = tr;
                  S.restarts[c]++;
                  S.rr[c] = S.rrs[c] = tr;
                  S.brk[c] = 0;
               }
            }
         }
         __syncthreads();
         if (S.todo[0] | S.todo[1] | S.todo[2]) {
            for (int c = 0; c < 3; c++) {
               const int todo = S.todo[c];
               if (!todo) continue;
               for (int i = tid; i < n; i += nthreads) {
                  const size_t k = (size_t)c * ld + i;
                  if (todo == 2) { a.x[k] = xbest[k]; continue; }
                  if (todo == 4) { a.x[k] = 0.0; continue; }
                  const double ri = AMs[k];
                  r[k] = ri; rs[k] = ri; p[k] = ri;
                  const double phi = a.dinv[i] * ri;
                  ph[k] = phi;
                  if (DIST) bicg_push(a, 0, c, i, phi);
                  if (todo == 3) xbest[k] = a.x[k];
               }
            }
         }
         BICG_HANDOVER();   // ph of the restarted systems is gathered by every block; the slice counter is reset for everybody
      }
      if (!S.act[0] && !S.act[1] && !S.act[2]) break;
      // ---- AMp = A (M p), d0 = rs . AMp
      {
         double d0[3] = {0, 0, 0}, d1[3] = {0, 0, 0};
         const bool act[3] = {S.act[0] != 0, S.act[1] != 0, S.act[2] != 0};
         bicg_spmv_phase<false>(a, ph, AMp, act, rs, d0, d1, gwarp, nwarps, lane);
         v3[0] = d0[0]; v3[1] = d0[1]; v3[2] = d0[2];
      }
      BICG_TICK(0);
      if (!BICG_SUM(3, v3, buf1)) return;
      if (tid == 0 && a.dynCounter) *a.dynCounter = 0;   // every block is past the SpMV phase
      BICG_TICK(1);
      // Breakdowns (the reference's libraries have no guards and hand back NaN fields): a non-finite alpha leaves x untouched, a
      // non-finite or zero omega keeps the half step, a non-finite beta the full step; each time the system goes through the
      // true-residual check above and is restarted from there.  All blocks see bit-identical reduced values, so they decide alike.
      if (threadIdx.x == 0) {
         S.passes++;
         for (int c = 0; c < 3; c++) {
            S.half[c] = 0; S.act2[c] = 0;
            const double al = S.act[c] ? S.rrs[c] / v3[c] : 0.0;
            if (S.act[c] && !isfinite(al)) { S.brk[c] = 1; S.alpha[c] = 0.0; }
            else S.alpha[c] = al;
         }
      }
      __syncthreads();
      // ---- s = r - alpha AMp (stored in r), sh = M s, |s|^2
      for (int c = 0; c < 3; c++) {
         double acc = 0.0;
         if (S.act[c] && !S.brk[c]) {
            const double alpha = S.alpha[c];
            for (int i = tid; i < n; i += nthreads) {
               const size_t k = (size_t)c * ld + i;
               const double si = r[k] - alpha * AMp[k];
               r[k] = si;
               const double shi = a.dinv[i] * si;
               shv[k] = shi;
               if (DIST) bicg_push(a, 1, c, i, shi);
               acc += si * si;
            }
         }
         v3[c] = acc;
      }
      BICG_TICK(2);
      if (!BICG_SUM_PUBLISH(3, v3, buf0)) return;   // publishes the shv values stored into the neighbours' windows
      BICG_TICK(3);
      if (threadIdx.x == 0)
         for (int c = 0; c < 3; c++) {   // converged at the half step (cusp::krylov::bicgstab): x += alpha M p, no iteration counted
            const int go = S.act[c] && !S.brk[c];
            S.half[c] = (go && a.halfstep && sqrt(v3[c]) <= S.thresh[c]) ? 1 : 0;
            S.act2[c] = (go && !S.half[c]) ? 1 : 0;
            if (S.half[c]) S.rr[c] = v3[c];
         }
      __syncthreads();
      if (S.act2[0] | S.act2[1] | S.act2[2]) {
         // ---- AMs = A (M s), omega = (AMs . s) / (AMs . AMs)
         {
            double d0[3] = {0, 0, 0}, d1[3] = {0, 0, 0};
            const bool act2[3] = {S.act2[0] != 0, S.act2[1] != 0, S.act2[2] != 0};
            bicg_spmv_phase<false>(a, shv, AMs, act2, r, d0, d1, gwarp, nwarps, lane);
            v6[0] = d0[0]; v6[1] = d0[1]; v6[2] = d0[2]; v6[3] = d1[0]; v6[4] = d1[1]; v6[5] = d1[2];
         }
         BICG_TICK(4);
         if (!BICG_SUM(6, v6, buf1)) return;
         if (tid == 0 && a.dynCounter) *a.dynCounter = 0;
         BICG_TICK(5);
         if (threadIdx.x == 0) {
            S.passes++;
            for (int c = 0; c < 3; c++) {
               const double om = S.act2[c] ? v6[c] / v6[3 + c] : 0.0;
               if (S.act2[c] && (!isfinite(om) || om == 0.0)) { S.act2[c] = 0; S.half[c] = 1; S.brk[c] = 1; S.omega[c] = 0.0; }   // keep x + alpha M p
               else S.omega[c] = om;
            }
         }
         __syncthreads();
      }
      // ---- x += alpha M p + omega M s ; r = s - omega AMs ; rs.r , r.r
      for (int c = 0; c < 3; c++) {
         double a0 = 0.0, a1 = 0.0;
         const double alpha = S.alpha[c], omega = S.omega[c];
         if (S.half[c]) {
            for (int i = tid; i < n; i += nthreads) {
               const size_t k = (size_t)c * ld + i;
               a.x[k] += alpha * ph[k];
            }
         } else if (S.act2[c]) {
            for (int i = tid; i < n; i += nthreads) {
               const size_t k = (size_t)c * ld + i;
               a.x[k] += alpha * ph[k] + omega * shv[k];
               const double ri = r[k] - omega * AMs[k];
               r[k] = ri;
               a0 += rs[k] * ri;
               a1 += ri * ri;
            }
         }
         v6[c] = a0; v6[3 + c] = a1;
      }
      BICG_TICK(6);
      if (S.act2[0] | S.act2[1] | S.act2[2]) { if (!BICG_SUM(6, v6, buf0)) return; }
      BICG_TICK(7);
      // ---- p = r + beta (p - omega AMp) ; ph = M p
      for (int c = 0; c < 3; c++) {
         if (!S.act2[c]) continue;
         const double omega = S.omega[c];
         const double beta = (v6[c] / S.rrs[c]) * (S.alpha[c] / omega);
         if (!isfinite(beta) || v6[c] == 0.0) continue;   // rho = 0: breakdown, handled below
         for (int i = tid; i < n; i += nthreads) {
            const size_t k = (size_t)c * ld + i;
            const double pi = r[k] + beta * (p[k] - omega * AMp[k]);
            p[k] = pi;
            const double phi = a.dinv[i] * pi;
            ph[k] = phi;
            if (DIST) bicg_push(a, 0, c, i, phi);
         }
      }
      __syncthreads();   // every thread has read S.rrs before thread 0 replaces it
      if (threadIdx.x == 0)
         for (int c = 0; c < 3; c++) {
            if (S.act2[c]) {   // (half steps: no iteration counted, like CUSP; the stopping rule at the top of the loop takes them)
               const double beta = (v6[c] / S.rrs[c]) * (S.alpha[c] / S.omega[c]);
               if (!isfinite(beta) || v6[c] == 0.0) S.brk[c] = 1;
               S.rrs[c] = v6[c]; S.rr[c] = v6[3 + c]; S.its[c]++;
            }
         }
      BICG_HANDOVER();
      BICG_TICK(8);
#ifdef CFD_BICG_PHASE_TIMERS
      nit++;
#endif
   }
#undef BICG_TICK
#undef BICG_SUM
#undef BICG_SUM_PUBLISH
#undef BICG_HANDOVER
#ifdef CFD_BICG_PHASE_TIMERS
   if (tid == 0) {
      for (int k = 0; k < 9; k++) a.info[kBicgInfoTimers + k] = (double)tph[k];
      a.info[kBicgInfoTimers + 9] = (double)nit;
   }
#endif
   if (tid == 0) {
      for (int c = 0; c < 3; c++) {
         a.info[kBicgInfoIts + c] = (double)S.its[c];
         a.info[kBicgInfoRR + c] = S.rr[c];
         a.info[kBicgInfoTrue + c] = a.verify ? S.trueRR[c] : -1.0;
         a.info[kBicgInfoBB + c] = S.bb[c];
         a.info[kBicgInfoFin + c] = (double)S.fin[c];
         a.info[kBicgInfoRestarts + c] = (double)S.restarts[c];
      }
      a.info[kBicgInfoPasses] = (double)S.passes;
      if (DIST) { a.xc.xstate[0] = epoch; a.xc.xstate[1] = hepoch; a.xc.xstate[3] = 0; a.xc.xstate[4] = 0; }
   }
}

}   // namespace cfdb

// =================================================================================================
// Host side of the context
// =================================================================================================
using namespace cfdb;

namespace {

// Event pair around one launch (no synchronisation here; resolved by profile_collect).
struct ProfScope {
   cfd_ctx* c;
   int idx = -1;
   ProfScope(cfd_ctx* c_, const char* n, Bytes b) : c(c_) {
      if (!c->profile || c->recsUsed >= (1u << 16)) return;
      if (c->recsUsed == c->recs.size()) {
         cfd_ctx::ProfRec r;
         r.cls = n; r.bytes = 0; r.lbytes = 0;
         if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
         c->recs.push_back(r);
      }
      idx = (int)c->recsUsed++;
      c->recs[idx].cls = n;
      c->recs[idx].bytes = b.a;
      c->recs[idx].lbytes = b.l;
      cudaEventRecord(c->recs[idx].a, c->stream);
   }
   ~ProfScope() {
      if (idx >= 0) cudaEventRecord(c->recs[idx].b, c->stream);
   }
};

void profile_collect(cfd_ctx* c) {
   if (!c->recsUsed) return;
   cudaStreamSynchronize(c->stream);
   for (size_t i = 0; i < c->recsUsed; i++) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, c->recs[i].a, c->recs[i].b) != cudaSuccess) continue;
      ProfEntry& p = c->prof[c->recs[i].cls];
      p.ms += ms; p.launches += 1; p.bytes += c->recs[i].bytes; p.lbytes += c->recs[i].lbytes;
   }
   c->recsUsed = 0;
   c->lastPcgRec = c->lastBicgRec = -1;
}

inline int gridFor(long long n, int block) { return (int)((n + block - 1) / block); }

template <typename T>
int dalloc(cfd_ctx* c, T** p, size_t count) {
   CFD_CK(cudaMalloc((void**)p, sizeof(T) * std::max<size_t>(count, 1)));
   CFD_CK(cudaMemset(*p, 0, sizeof(T) * std::max<size_t>(count, 1)));
   c->allocs.push_back((void*)*p);
   return CFD_OK;
}
template <typename T>
int dupload(cfd_ctx* c, T** p, const T* host, size_t count) {
   int rc = dalloc(c, p, count);
   if (rc) return rc;
   if (count) CFD_CK(cudaMemcpy(*p, host, sizeof(T) * count, cudaMemcpyHostToDevice));
   return CFD_OK;
}

// set-up scratch on the device, released on every exit path (device step0)
// Uploads and zero fills are issued on the context's stream (created non-blocking: work on the legacy stream is NOT ordered with it),
// so they are ordered with the step0 kernels that consume them.
struct DevScratch {
   std::vector<void*> v;
   cudaStream_t st = nullptr;
   ~DevScratch() { for (void* q : v) cudaFree(q); }
   int alloc(void** q, size_t bytes) {
      CFD_CK(cudaMalloc(q, std::max<size_t>(bytes, 16)));
      v.push_back(*q);
      return CFD_OK;
   }
   template <class T> int up(T** q, const T* host, size_t n) {
      int rc = alloc((void**)q, sizeof(T) * n);
      if (rc) return rc;
      CFD_CK(cudaMemcpyAsync(*q, host, sizeof(T) * n, cudaMemcpyHostToDevice, st));   // (pageable source: returns once it is staged)
      return CFD_OK;
   }
   int zeros(double** q, size_t n) {
      int rc = alloc((void**)q, sizeof(double) * n);
      if (rc) return rc;
      CFD_CK(cudaMemsetAsync(*q, 0, sizeof(double) * std::max<size_t>(n, 2), st));
      return CFD_OK;
   }
};

int regvec(cfd_ctx* c, const char* name, double** p, long long n) {
   int rc = dalloc(c, p, (size_t)n);
   if (rc) return rc;
   c->vecs[name] = std::make_pair(p, n);
   return CFD_OK;
}

// algorithmic bytes of a CSR SpMV with nvec right-hand sides
// ... and the same launch in the bytes the Sell layout stores: padded values, (compressed) column indices, slice offsets
double matrix_lbytes(const Sell& A, int nval) { return 8.0 * nval * (double)A.padded + 4.0 * (double)A.nindex + 16.0 * A.nslices; }
Bytes spmv_bytes(const Sell& A, int nval, int nvec, bool accumulate) {
   const double vec = 8.0 * nvec * ((double)A.ncols + (double)A.nrows * (accumulate ? 2 : 1));
   return Bytes((8.0 * nval + 4.0) * (double)A.nnz + 4.0 * (A.nrows + 1.0) + vec, matrix_lbytes(A, nval) + vec);
}

// ---- multi-GPU primitives -----------------------------------------------------------------------------------
inline bool parted(const cfd_ctx* c) { return c->P.world > 1; }

int need_comm(cfd_ctx* c) {
   if (parted(c) && (!c->comm || !c->comm->comm)) { set_error("partitioned context without communicator: call cfd_comm_init first"); return CFD_ERR_STATE; }
   return CFD_OK;
}

// Ghost entries of `vec` <- owners' values.  velocity = true: 3-component node vector (stride NN); false: pressure vector.
int halo(cfd_ctx* c, double* vec, bool velocity) {
   if (!parted(c)) return CFD_OK;
   int rc = need_comm(c);
   if (rc) return rc;
   Comm* cm = c->comm;
   const int np = (int)c->peerRank.size();
   const std::vector<int>& ptr = velocity ? c->sendVptr : c->sendPptr;
   const std::vector<int>& roff = velocity ? c->recvVoff : c->recvPoff;
   const std::vector<int>& rcnt = velocity ? c->recvVcnt : c->recvPcnt;
   const int ncomp = velocity ? 3 : 1;
   const long long ld = velocity ? c->P.NN : c->P.NNp;
   const int ns = ptr[np];
   const int* plan = velocity ? c->sendVidx : c->sendPidx;
   ProfScope ps(c, "halo", 8.0 * ncomp * ns);
   if (ns > 0) {
      k_pack<<<gridFor(ns, 256), 256, 0, c->stream>>>(ns, ncomp, ld, plan, plan + ns, plan + 2 * ns, vec, c->sendBuf);
      c->launches++;
      CFD_CK(cudaGetLastError());
   }
   CFD_NCCL(cm, cm->GroupStart());
   for (int j = 0; j < np; j++) {
      const int cnt = ptr[j + 1] - ptr[j];
      for (int q = 0; q < ncomp; q++) {
         if (cnt > 0) CFD_NCCL(cm, cm->Send(c->sendBuf + (size_t)ncomp * ptr[j] + (size_t)q * cnt, cnt, kNcclFloat64, c->peerRank[j], cm->comm, c->stream));
         if (rcnt[j] > 0) CFD_NCCL(cm, cm->Recv(vec + q * ld + roff[j], rcnt[j], kNcclFloat64, c->peerRank[j], cm->comm, c->stream));
      }
   }
   CFD_NCCL(cm, cm->GroupEnd());
   return CFD_OK;
}

int allreduce(cfd_ctx* c, double* dev, int n, int op) {
   if (!parted(c)) return CFD_OK;
   int rc = need_comm(c);
   if (rc) return rc;
   CFD_NCCL(c->comm, c->comm->AllReduce(dev, dev, n, kNcclFloat64, op, c->comm->comm, c->stream));
   return CFD_OK;
}

// Every rank must take the same Krylov path (peer-memory kernels or host-driven NCCL): a rank whose window mapping failed, or that
// was started with CFD_NO_P2P, switches ALL ranks to the NCCL path -- otherwise the ranks would wait for each other in different
// protocols.  (min-reduction of the local decision over the communicator; without a communicator the local decision stands.)
int p2p_agree(cfd_ctx* c) {
   if (!c->comm || !c->comm->comm) return CFD_OK;
   double v = c->p2p ? 1.0 : 0.0;
   CFD_CK(cudaMemcpyAsync(c->dacc + 48, &v, sizeof v, cudaMemcpyHostToDevice, c->stream));
   CFD_NCCL(c->comm, c->comm->AllReduce(c->dacc + 48, c->dacc + 48, 1, kNcclFloat64, kNcclMin, c->comm->comm, c->stream));
   CFD_CK(cudaMemcpyAsync(&v, c->dacc + 48, sizeof v, cudaMemcpyDeviceToHost, c->stream));
   CFD_CK(cudaStreamSynchronize(c->stream));
   c->p2p = v > 0.5;
   return CFD_OK;
}

// Distributed Jacobi-PCG (phase kernels + NCCL; see dist.cuh).  State read back once per batch of iterations.
int run_pcg_dist(cfd_ctx* c, const double* b, double* x, bool x0zero) {
   const int n = c->NPo;
   double* acc = c->dacc;
   DPcgState* st = (DPcgState*)(c->dacc + 32);
   const int gridV = std::min(gridFor(n, 256), 8 * c->numSMs);
   const int gridS = gridFor((long long)c->Z.nslices * 32, 256);
   int rc;
   ProfScope ps(c, "pcg", 0.0);
   c->lastPcgRec = ps.idx;
   CFD_CK(cudaMemsetAsync(acc, 0, 8 * sizeof(double), c->stream));
   if (!x0zero) {
      if ((rc = halo(c, x, false))) return rc;
      k_spmv_multi<1><<<gridS, 256, 0, c->stream>>>(c->Z, 0, x, 0, c->cg_q, 0, 1.0, 0.0);
      c->launches++;
   }
   k_dpcg_init<<<gridV, 256, 0, c->stream>>>(n, x0zero ? 1 : 0, b, x, c->cg_q, c->Zdinv, c->cg_r, c->cg_p, acc);
   if ((rc = allreduce(c, acc + 3, 3, kNcclSum))) return rc;
   k_dpcg_begin<<<1, 1, 0, c->stream>>>(st, acc, c->opt.cg_tol, c->opt.cg_maxit);
   c->launches += 2;
   const int batch = 8;
   for (int it0 = 0; it0 <= c->opt.cg_maxit; it0 += batch) {
      for (int k = 0; k < batch; k++) {
         if ((rc = halo(c, c->cg_p, false))) return rc;
         k_dpcg_spmv<<<gridS, 256, 0, c->stream>>>(c->Z, st, c->cg_p, c->cg_q, acc);
         if ((rc = allreduce(c, acc, 1, kNcclSum))) return rc;
         k_dpcg_update<<<gridV, 256, 0, c->stream>>>(n, st, c->cg_p, c->cg_q, c->Zdinv, x, c->cg_r, acc);
         if ((rc = allreduce(c, acc + 1, 2, kNcclSum))) return rc;
         k_dpcg_pupdate<<<gridV, 256, 0, c->stream>>>(n, st, c->cg_r, c->Zdinv, c->cg_p, acc);
         k_dpcg_next<<<1, 1, 0, c->stream>>>(st, acc);
         c->launches += 4;
      }
      CFD_CK(cudaMemcpyAsync(c->h_pinned + 40, st, sizeof(DPcgState), cudaMemcpyDeviceToHost, c->stream));
      CFD_CK(cudaStreamSynchronize(c->stream));
      const DPcgState* hs = (const DPcgState*)(c->h_pinned + 40);
      if (hs->done) {
         // mirror the persistent kernel's result slots
         const double info[3] = {(double)hs->it, hs->rr, hs->bb};
         CFD_CK(cudaMemcpyAsync(c->kinfo, info, sizeof info, cudaMemcpyHostToDevice, c->stream));
         CFD_CK(cudaStreamSynchronize(c->stream));
         break;
      }
   }
   CFD_CK(cudaGetLastError());
   return CFD_OK;
}

// Distributed Jacobi-BiCGStab for the three momentum systems (phase kernels + NCCL).
int run_bicg_dist(cfd_ctx* c) {
   const int n = c->NNo;
   const long long ld = c->P.NN;
   const size_t N3 = 3 * (size_t)ld;
   double* acc = c->dacc;
   DBicgState* st = (DBicgState*)(c->dacc + 32);
   double *r = c->bw, *rs = c->bw + N3, *p = c->bw + 2 * N3, *ph = c->bw + 3 * N3, *AMp = c->bw + 4 * N3, *shv = c->bw + 5 * N3, *AMs = c->bw + 6 * N3;
   const int gridV = std::min(gridFor(n, 256), 8 * c->numSMs);
   const int gridS = gridFor((long long)c->K.nslices * 32, 256);
   int rc;
   ProfScope ps(c, "bicgstab", 0.0);
   c->lastBicgRec = ps.idx;
   CFD_CK(cudaMemsetAsync(acc, 0, 24 * sizeof(double), c->stream));
   k_dbicg_init<<<gridV, 256, 0, c->stream>>>(n, ld, c->R1i, c->Adinv, c->Uk, r, rs, p, ph, acc);
   if ((rc = allreduce(c, acc, 3, kNcclSum))) return rc;
   k_dbicg_begin<<<1, 1, 0, c->stream>>>(st, acc, c->opt.bicg_tol, c->opt.bicg_maxit);
   c->launches += 2;
   const int batch = 4;
   for (int it0 = 0; it0 <= c->opt.bicg_maxit; it0 += batch) {
      for (int k = 0; k < batch; k++) {
         if ((rc = halo(c, ph, true))) return rc;
         k_dbicg_spmv<<<gridS, 256, 0, c->stream>>>(c->K, 2, ld, st, 0, ph, AMp, rs, acc, 0, 3);
         if ((rc = allreduce(c, acc, 3, kNcclSum))) return rc;
         k_dbicg_s<<<gridV, 256, 0, c->stream>>>(n, ld, st, AMp, c->Adinv, r, shv, acc);
         if ((rc = allreduce(c, acc + 6, 3, kNcclSum))) return rc;
         k_dbicg_mid<<<1, 1, 0, c->stream>>>(st, acc);
         if ((rc = halo(c, shv, true))) return rc;
         k_dbicg_spmv<<<gridS, 256, 0, c->stream>>>(c->K, 2, ld, st, 1, shv, AMs, r, acc, 12, 15);
         if ((rc = allreduce(c, acc + 12, 6, kNcclSum))) return rc;
         k_dbicg_x<<<gridV, 256, 0, c->stream>>>(n, ld, st, ph, shv, AMs, rs, c->Uk, r, acc);
         if ((rc = allreduce(c, acc + 18, 6, kNcclSum))) return rc;
         k_dbicg_p<<<gridV, 256, 0, c->stream>>>(n, ld, st, r, AMp, c->Adinv, p, ph, acc);
         k_dbicg_next<<<1, 1, 0, c->stream>>>(st, acc);
         c->launches += 7;
      }
      CFD_CK(cudaMemcpyAsync(c->h_pinned + 40, st, sizeof(DBicgState), cudaMemcpyDeviceToHost, c->stream));
      CFD_CK(cudaStreamSynchronize(c->stream));
      const DBicgState* hs = (const DBicgState*)(c->h_pinned + 40);
      if (hs->done) {
         // mirror the persistent kernel's result slots (this host-driven fallback trusts the recurrence residual: true = -1)
         double info[19];
         for (int d = 0; d < 3; d++) {
            info[kBicgInfoIts - 3 + d] = (double)hs->its[d];
            info[kBicgInfoRR - 3 + d] = hs->rr[d];
            info[kBicgInfoTrue - 3 + d] = -1.0;
            info[kBicgInfoBB - 3 + d] = hs->bb[d];
            info[kBicgInfoFin - 3 + d] = (sqrt(hs->rr[d]) <= hs->thresh[d]) ? 1.0 : 2.0;
            info[kBicgInfoRestarts - 3 + d] = 0.0;
         }
         info[kBicgInfoPasses - 3] = 0.0;
         CFD_CK(cudaMemcpyAsync(c->kinfo + 3, info, sizeof info, cudaMemcpyHostToDevice, c->stream));
         CFD_CK(cudaStreamSynchronize(c->stream));
         break;
      }
   }
   CFD_CK(cudaGetLastError());
   return CFD_OK;
}

static void* pcgcg_kernel(bool dist, int minb) {
   if (dist) return minb == 2 ? (void*)k_pcg_cg<true, 2> : (void*)k_pcg_cg<true, 3>;
   return minb == 2 ? (void*)k_pcg_cg<false, 2> : (void*)k_pcg_cg<false, 3>;
}
void fill_xcomm(cfd_ctx* c, XComm& xc) {
   xc.world = c->P.world; xc.rank = c->P.rank; xc.npeers = (int)c->peerRank.size();
   xc.peerWin = c->dPeerWin; xc.peerRank = c->dPeerRank; xc.peerP = c->dPeerP; xc.pushPtr = c->pushPtrP; xc.pushDst = c->pushDstP;
   xc.ifaceRows = c->ifaceP; xc.nIface = c->nIfaceP;
   xc.xstate = c->xstate; xc.timeoutNs = c->p2pTimeoutNs;
}

// Peer-memory PCG: one cooperative launch per rank, all exchange inside the kernel (k_pcg_p2p).
int run_pcg_p2p(cfd_ctx* c, const double* b, double* x, bool x0zero) {
   int rc;
   if (!x0zero && (rc = halo(c, x, false))) return rc;
   PcgP2PArgs pa;
   PcgArgs& a = pa.a;
   a.Z = c->Z; a.part = c->pcgPart; a.dynStart = c->pcgDynStart; a.dynCounter = c->pcgDynCounter; a.n = c->NPo; a.b = b; a.x = x; a.r = c->cg_r; a.p = c->cg_p; a.q = c->cg_q; a.dinv = c->Zdinv;
   a.scratch = c->partials; a.info = c->kinfo; a.tol = c->opt.cg_tol; a.maxit = c->opt.cg_maxit; a.x0zero = x0zero ? 1 : 0;
   fill_xcomm(c, pa.xc);
   void* args[] = {&pa};
   ProfScope ps(c, "pcg", x0zero ? Bytes(0.0) : spmv_bytes(c->Z, 1, 1, false));
   c->lastPcgRec = ps.idx;
   CFD_CK(cudaLaunchCooperativeKernel((void*)k_pcg_p2p, dim3(c->pcg_blocks), dim3(c->pcg_threads), args, 0, c->stream));
   c->launches++;
   return CFD_OK;
}

// Single-reduction (Chronopoulos-Gear) PCG: one cooperative launch; on partitions the peer-memory form.
int run_pcg_cg(cfd_ctx* c, const double* b, double* x, bool x0zero) {
   const bool dist = parted(c);
   int rc;
   if (dist && !x0zero && (rc = halo(c, x, false))) return rc;
   PcgCgArgs a;
   a.Z = c->Z; a.part = c->pcgcgPart; a.dynStart = c->pcgDynStart; a.dynCounter = c->pcgDynCounter; a.n = dist ? c->NPo : c->P.NNp; a.b = b; a.x = x;
   a.r = c->cg_r; a.u = c->cg_p; a.w = c->cg_q; a.p = c->cg_pp; a.s = c->cg_s; a.dinv = c->Zdinv;
   a.scratch = c->partials; a.info = c->kinfo; a.tol = c->opt.cg_tol; a.maxit = c->opt.cg_maxit; a.x0zero = x0zero ? 1 : 0;
   memset(&a.xc, 0, sizeof a.xc);
   if (dist) fill_xcomm(c, a.xc);
   void* args[] = {&a};
   ProfScope ps(c, "pcg", spmv_bytes(c->Z, 1, 1, false) + (x0zero ? Bytes(0.0) : spmv_bytes(c->Z, 1, 1, false)));   // one sweep more than iterations
   c->lastPcgRec = ps.idx;
   CFD_CK(cudaLaunchCooperativeKernel(pcgcg_kernel(dist, c->pcg_minb), dim3(c->pcgcg_blocks), dim3(c->pcg_threads), args, 0, c->stream));
   c->launches++;
   return CFD_OK;
}

int run_pcg(cfd_ctx* c, const double* b, double* x, bool x0zero) {
   if (parted(c) && !c->p2p) return run_pcg_dist(c, b, x, x0zero);
   if (c->cg_variant == 2 && !c->pcgSmemBytes && !c->resBytes) return run_pcg_cg(c, b, x, x0zero);
   if (parted(c)) return run_pcg_p2p(c, b, x, x0zero);
   PcgArgs a;
   a.Z = c->Z; a.part = c->pcgPart; a.dynStart = c->pcgDynStart; a.dynCounter = c->pcgDynCounter; a.n = c->P.NNp; a.b = b; a.x = x; a.r = c->cg_r; a.p = c->cg_p; a.q = c->cg_q; a.dinv = c->Zdinv;
   a.scratch = c->partials; a.info = c->kinfo; a.tol = c->opt.cg_tol; a.maxit = c->opt.cg_maxit; a.x0zero = x0zero ? 1 : 0;
   if (c->resBytes > 0) {   // small system: row blocks resident per CTA (values in shared memory, vectors in registers)
      PcgResArgs ra;
      ra.cta = (const ResCta*)c->resCta; ra.vals = c->resVals; ra.idx = c->resIdx; ra.cols = c->resCols; ra.grp = c->resGrp; ra.T = c->resT;
      ra.n = c->P.NNp; ra.b = b; ra.x = x; ra.u = c->cg_p; ra.dinv = c->Zdinv; ra.scratch = c->partials; ra.info = c->kinfo;
      ra.tol = c->opt.cg_tol; ra.maxit = c->opt.cg_maxit; ra.x0zero = x0zero ? 1 : 0;
      void* rargs[] = {&ra};
      ProfScope ps(c, "pcg", spmv_bytes(c->Z, 1, 1, false) + (x0zero ? Bytes(0.0) : spmv_bytes(c->Z, 1, 1, false)));
      c->lastPcgRec = ps.idx;
      CFD_CK(cudaLaunchCooperativeKernel((void*)k_pcg_res, dim3(c->numSMs), dim3(kResThreads), rargs, c->resBytes, c->stream));
      c->launches++;
      return CFD_OK;
   }
   if (c->pcgSmemBytes > 0) {   // small system: matrix resident in shared memory, one CTA per SM
      PcgSmemArgs pa;
      pa.a = a;
      pa.a.part = c->pcgSmemPart; pa.a.dynCounter = nullptr; pa.a.dynStart = c->Z.nslices;
      pa.s = c->cg_s; pa.pp = c->cg_pp;
      void* sargs[] = {&pa};
      ProfScope ps(c, "pcg", x0zero ? Bytes(0.0) : spmv_bytes(c->Z, 1, 1, false));
      c->lastPcgRec = ps.idx;
      CFD_CK(cudaLaunchCooperativeKernel(c->pcgSmemSingle ? (void*)k_pcg_smem<true> : (void*)k_pcg_smem<false>, dim3(c->numSMs), dim3(1024), sargs,
                                         c->pcgSmemBytes, c->stream));
      c->launches++;
      return CFD_OK;
   }
   void* args[] = {&a};
   ProfScope ps(c, "pcg", x0zero ? Bytes(0.0) : spmv_bytes(c->Z, 1, 1, false));
   c->lastPcgRec = ps.idx;
   CFD_CK(cudaLaunchCooperativeKernel(pcg_kernel(c->pcg_minb), dim3(c->pcg_blocks), dim3(c->pcg_threads), args, 0, c->stream));
   c->launches++;
   return CFD_OK;
}

int fetch_kinfo(cfd_ctx* c) {
   double* hk = c->h_pinned + 128;
   CFD_CK(cudaMemcpyAsync(hk, c->kinfo, 32 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
   CFD_CK(cudaStreamSynchronize(c->stream));
   c->last_cg = (int)hk[0];
   c->last_cg_res = hk[2] > 0.0 ? sqrt(hk[1] / hk[2]) : 0.0;
   c->last_cg_conv = (hk[2] == 0.0 || sqrt(hk[1]) <= c->opt.cg_tol * sqrt(hk[2])) ? 1 : 0;
   for (int d = 0; d < 3; d++) {
      c->last_bicg[d] = (int)hk[kBicgInfoIts + d];
      const double bb = hk[kBicgInfoBB + d], tr = hk[kBicgInfoTrue + d];
      c->last_bicg_conv[d] = (int)hk[kBicgInfoFin + d] == 1 ? 1 : 0;
      c->last_bicg_true[d] = tr < 0.0 ? -1.0 : (bb > 0.0 ? sqrt(tr / bb) : 0.0);
      c->last_bicg_restarts[d] = (int)hk[kBicgInfoRestarts + d];
   }
   c->last_bicg_passes = (int)hk[kBicgInfoPasses];
#ifdef CFD_PCG_PHASE_TIMERS
   if (c->cg_variant == 1 && !parted(c) && !c->pcgSmemBytes && !c->resBytes && c->last_cg > 0 && getenv("CFD_PCG_PHASES"))
      fprintf(stderr, "pcg (classical) phases, us / iteration, block 0 (%d iterations): spmv %.2f | reduction 1 %.2f | update %.2f | reduction 2 %.2f | p update + sync %.2f\n",
              c->last_cg, hk[8] / c->last_cg * 1e-3, hk[9] / c->last_cg * 1e-3, hk[10] / c->last_cg * 1e-3, hk[11] / c->last_cg * 1e-3, hk[12] / c->last_cg * 1e-3);
   if (c->cg_variant == 1 && parted(c) && c->p2p && c->last_cg > 0 && getenv("CFD_PCG_PHASES"))
      for (int b = 0; b < 2; b++)
         fprintf(stderr, "[rank %d] pcg (classical, peer memory) phases, us / iteration, %s block (%d iterations): spmv %.2f | all-reduce 1 %.2f | update %.2f | all-reduce 2 %.2f | "
                         "p update + push + flags %.2f\n", c->P.rank, b ? "last" : "first", c->last_cg, hk[8 + 5 * b] / c->last_cg * 1e-3, hk[9 + 5 * b] / c->last_cg * 1e-3,
                 hk[10 + 5 * b] / c->last_cg * 1e-3, hk[11 + 5 * b] / c->last_cg * 1e-3, hk[12 + 5 * b] / c->last_cg * 1e-3);
   if (c->resBytes && c->last_cg > 0 && getenv("CFD_PCG_PHASES"))
      fprintf(stderr, "pcg (k_pcg_res) phases, us / iteration, block 0 (%d iterations): vector pass %.2f | grid sync %.2f | sweep %.2f | reduction %.2f\n",
              c->last_cg, hk[8] / c->last_cg * 1e-3, hk[9] / c->last_cg * 1e-3, hk[10] / c->last_cg * 1e-3, hk[11] / c->last_cg * 1e-3);
   if (c->resBytes && c->last_cg > 0 && getenv("CFD_PCG_PHASES"))
      fprintf(stderr, "   inside the sweep (block 0, %d sweeps): staging %.2f | warp 0's row groups %.2f | wait for the other warps %.2f\n",
              c->last_cg + 1, hk[24] / (c->last_cg + 1) * 1e-3, hk[25] / (c->last_cg + 1) * 1e-3, hk[26] / (c->last_cg + 1) * 1e-3);
   else if (c->pcgSmemBytes && c->pcgSmemSingle && c->last_cg > 0 && getenv("CFD_PCG_PHASES"))
      fprintf(stderr, "pcg (resident, single-reduction) phases, us / iteration, block 0 (%d iterations): vector pass %.2f | grid sync %.2f | sweep %.2f | reduction %.2f\n",
              c->last_cg, hk[8] / c->last_cg * 1e-3, hk[9] / c->last_cg * 1e-3, hk[10] / c->last_cg * 1e-3, hk[11] / c->last_cg * 1e-3);
   else if (c->cg_variant == 2 && c->last_cg > 0 && getenv("CFD_PCG_PHASES"))
      fprintf(stderr, "[rank %d] pcg (single-reduction) phases, us / iteration, block 0 (%d iterations): vector pass %.2f | grid sync + halo wait %.2f | sweep %.2f | reduction %.2f\n",
              c->P.rank, c->last_cg, hk[8] / c->last_cg * 1e-3, hk[9] / c->last_cg * 1e-3, hk[10] / c->last_cg * 1e-3, hk[11] / c->last_cg * 1e-3);
#endif
   if (c->p2p && c->xstate) {   // a peer that never arrived: the spin-waits of k_pcg_p2p give up after 4 s and raise this word
      unsigned long long err[6] = {0, 0, 0, 0, 0, 0};
      CFD_CK(cudaMemcpy(err, c->xstate + 2, sizeof err, cudaMemcpyDeviceToHost));
      if (err[0]) {
         set_error("peer-memory exchange timed out (a rank stopped participating) [rank %d waited for %s: code %llu, epoch %llu, last value seen 0x%llx]",
                   c->P.rank, (err[3] & 15) == 1 ? "a neighbour's halo flag" : "the all-reduce words", err[3] >> 4, err[4], err[5]);
         return CFD_ERR_STATE;
      }
   }
   // algorithmic bytes of the Krylov launches are known only now (iteration counts): DESIGN.md byte model
   if (c->lastPcgRec >= 0) {
      const double N = c->P.NNp;
      // classical: 104 B per row and iteration; single-reduction form: 12 vector accesses in the pass + w, r, u in the sweep
      const double vecB = c->cg_variant == 2 && !(parted(c) && !c->p2p) ? 120.0 : 104.0;
      c->recs[c->lastPcgRec].bytes += 48.0 * N + c->last_cg * (12.0 * (double)c->Z.nnz + 4.0 * (N + 1) + vecB * N);
      c->recs[c->lastPcgRec].lbytes += 48.0 * N + c->last_cg * (matrix_lbytes(c->Z, 1) + vecB * N);
      c->lastPcgRec = -1;
   }
   if (c->lastBicgRec >= 0) {
      const double N = c->P.NN;
      const int mx = std::max(c->last_bicg[0], std::max(c->last_bicg[1], c->last_bicg[2]));
      // A is streamed once per SpMV phase: twice per lock-step iteration, plus the true-residual checks (counted by the kernel)
      const double passes = c->last_bicg_passes > 0 ? (double)c->last_bicg_passes : 2.0 * mx;
      double vec = 0;
      for (int d = 0; d < 3; d++) vec += c->last_bicg[d] * 192.0 * N;    // 24 vector reads/writes per system-iteration
      c->recs[c->lastBicgRec].bytes += passes * (12.0 * (double)c->K.nnz + 4.0 * (N + 1)) + vec;
      c->recs[c->lastBicgRec].lbytes += passes * matrix_lbytes(c->K, 1) + vec;
      c->lastBicgRec = -1;
   }
   return CFD_OK;
}

template <int NV>
int launch_ebe_apply(cfd_ctx* c, const double* Uadv, const double* Ux, double* R1) {
   const int NE = c->NEa;   // elements touching an owned node
   if (c->opt.ebe_scatter == 1) {
      for (size_t k = 0; k + 1 < c->colorStart.size(); k++) {
         const int n = c->colorStart[k + 1] - c->colorStart[k];
         k_ebe_convect_apply<NV, false><<<gridFor(n, 128), 128, 0, c->stream>>>(n, c->P.NN, c->ldE, c->LtoGT, c->coord, Uadv, Ux, R1,
                                                                                c->elemOfColor + c->colorStart[k]);
         c->launches++;
      }
   } else {
      k_ebe_convect_apply<NV, true><<<gridFor(NE, 128), 128, 0, c->stream>>>(NE, c->P.NN, c->ldE, c->LtoGT, c->coord, Uadv, Ux, R1, nullptr);
      c->launches++;
   }
   CFD_CK(cudaGetLastError());
   return CFD_OK;
}
template <int NV>
int launch_ebe_assemble(cfd_ctx* c, const double* Uadv, double* Aval) {
   const int NE = c->NEa;
   if (c->opt.ebe_scatter == 1) {
      for (size_t k = 0; k + 1 < c->colorStart.size(); k++) {
         const int n = c->colorStart[k + 1] - c->colorStart[k];
         k_ebe_convect_assemble<NV, false><<<gridFor(n, 128), 128, 0, c->stream>>>(n, c->P.NN, c->ldE, c->LtoGT, c->coord, Uadv, c->mapT, Aval,
                                                                                   c->elemOfColor + c->colorStart[k]);
         c->launches++;
      }
   } else {
      k_ebe_convect_assemble<NV, true><<<gridFor(NE, 128), 128, 0, c->stream>>>(NE, c->P.NN, c->ldE, c->LtoGT, c->coord, Uadv, c->mapT, Aval, nullptr);
      c->launches++;
   }
   CFD_CK(cudaGetLastError());
   return CFD_OK;
}

Bytes ebe_bytes(const cfd_ctx* c, bool assemble) {
   // per element: node ids + 8 corner coordinates + gather of the advecting field (+ second field and
   // scatter RMW for the apply form; + map and RMW of NENv^2 matrix entries for the assembly form)
   const double nv = c->P.NENv;
   double per = 4.0 * nv + 8.0 * 8 * 3 + 8.0 * 3 * nv;
   per += assemble ? (4.0 + 16.0) * nv * nv : (8.0 * 3 * nv + 16.0 * 3 * nv);
   return per * c->P.NE;
}

// ---- matrix-free operator path (cfd_options.operator_path = 1) -------------------------------------------------------
inline bool matrix_free(const cfd_ctx* c) { return c->opt.operator_path == 1; }
// algorithmic bytes of one element pass: node ids + corner coordinates + `gath` gathered and `scat` scattered (read-modify-write) values
double mf_bytes(const cfd_ctx* c, double gath, double scat) { return (4.0 * c->P.NENv + 192.0 + 8.0 * gath + 16.0 * scat) * (double)c->P.NE; }

template <int MODE>
int launch_momentum_apply(cfd_ctx* c, const double* Uadv, const double* X, const double* Pv, double sign, double* out) {
   const int NE = c->P.NE, NN = c->P.NN;
   const double invRho = 1.0 / c->P.density;
   if (c->P.NENv == 27)
      k_ebe_momentum_apply<27, MODE><<<gridFor(NE, 128), 128, 0, c->stream>>>(NE, NN, c->ldE, c->LtoGT, c->coord, Uadv, X, Pv, c->P.viscosity, invRho, sign, out);
   else
      k_ebe_momentum_apply<8, MODE><<<gridFor(NE, 128), 128, 0, c->stream>>>(NE, NN, c->ldE, c->LtoGT, c->coord, Uadv, X, Pv, c->P.viscosity, invRho, sign, out);
   c->launches++;
   CFD_CK(cudaGetLastError());
   return CFD_OK;
}
int launch_gt_apply(cfd_ctx* c, int mode, const double* U, const double* KAcc, int useKAcc, const double* MdOrigInv, double scale, double* R2) {
   const int NE = c->P.NE, NN = c->P.NN;
   const double invRho = 1.0 / c->P.density;
   CFD_CK(cudaMemsetAsync(R2, 0, sizeof(double) * (size_t)c->P.NNp, c->stream));
   if (c->P.NENv == 27)
      k_ebe_gt_apply<27><<<gridFor(NE, 128), 128, 0, c->stream>>>(NE, NN, c->ldE, c->LtoGT, c->coord, mode, U, KAcc, useKAcc, MdOrigInv, scale, invRho, R2);
   else
      k_ebe_gt_apply<8><<<gridFor(NE, 128), 128, 0, c->stream>>>(NE, NN, c->ldE, c->LtoGT, c->coord, mode, U, KAcc, useKAcc, MdOrigInv, scale, invRho, R2);
   c->launches++;
   CFD_CK(cudaGetLastError());
   if (c->pin >= 0) CFD_CK(cudaMemsetAsync(R2 + c->pin, 0, sizeof(double), c->stream));   // applyPresBC: R2[zeroPressureNode] = 0
   return CFD_OK;
}

}   // namespace

extern "C" {

const char* cfd_last_error(void) { return cfdb::last_error(); }
const char* cfd_version(void) { return "cfdb200 0.2 (sm_100a)"; }
void cfd_struct_sizes(int out[3]) { out[0] = (int)sizeof(cfd_problem); out[1] = (int)sizeof(cfd_options); out[2] = (int)sizeof(cfd_step_report); }

void cfd_default_options(cfd_options* o) {
   memset(o, 0, sizeof *o);
   o->cg_tol = 1e-6; o->cg_maxit = 1000; o->cg_warm_start = 0;
   o->bicg_tol = 1e-6; o->bicg_maxit = 1000; o->maxacc_rule = 0; o->device = -1;
}

int cfd_selectCUDAdevice(int device) {
   int n = 0;
   CFD_CK(cudaGetDeviceCount(&n));
   if (n == 0) { set_error("no CUDA device: this library has no CPU fallback"); return CFD_ERR_CUDA; }
   if (device < 0 || device >= n) { set_error("device %d out of range (0..%d)", device, n - 1); return CFD_ERR_ARG; }
   CFD_CK(cudaSetDevice(device));
   return CFD_OK;
}

void cfd_destroy(cfd_ctx* c) {
   if (!c) return;
   cudaSetDevice(c->device);
   if (c->stream) cudaStreamSynchronize(c->stream);
   for (void* p : c->allocs) cudaFree(p);
   sell_free(c->K); sell_free(c->G); sell_free(c->GT); sell_free(c->Z);
   if (c->comm) { if (c->comm->comm) c->comm->CommDestroy(c->comm->comm); delete c->comm; }
   for (size_t q = 0; q < c->peerWin.size(); q++)
      if (c->peerWin[q] && c->peerWin[q] != c->win) cudaIpcCloseMemHandle(c->peerWin[q]);
   if (c->win) cudaFree(c->win);
   if (c->h_pinned) cudaFreeHost(c->h_pinned);
   if (c->ev0) cudaEventDestroy(c->ev0);
   if (c->ev1) cudaEventDestroy(c->ev1);
   for (auto& r : c->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
   if (c->stream && c->ownStream) cudaStreamDestroy(c->stream);
   delete c;
}

int cfd_initializeAndAllocateGPU(const cfd_problem* P, const cfd_options* opt, cfd_ctx** out) {
   if (!P || !out) { set_error("null argument"); return CFD_ERR_ARG; }
   if ((P->NENv != 27 && P->NENv != 8) || P->NENp != 8 || P->NGP != 8) {
      set_error("unsupported element: NENv=%d NENp=%d NGP=%d (supported: 27/8/8 and 8/8/8)", P->NENv, P->NENp, P->NGP);
      return CFD_ERR_ARG;
   }
   // step0_device = 1: the value arrays (K, M, G1..3, Z, MdInv, MdOrigInv) are computed on the device from the mesh and the patterns
   const bool dev0 = opt && opt->step0_device == 1;
   if (opt && opt->step0_device != 0 && opt->step0_device != 1) { set_error("step0_device must be 0 (value arrays from the host) or 1 (computed on the device)"); return CFD_ERR_ARG; }
   if (dev0 && P->world > 1) { set_error("step0_device = 1 is single-GPU (partitioned sub-meshes number their pressure DOFs separately)"); return CFD_ERR_ARG; }
   if (!P->coord || !P->LtoGnode || !P->sparseMrowStarts || !P->sparseMcol || !P->sparseGrowStarts || !P->sparseGcol || !P->sparseZrowStarts ||
       !P->sparseZcol || (P->BCnVelNodes > 0 && (!P->BCvelNodes || !P->BCstr))) {
      set_error("cfd_problem: a required array is NULL");
      return CFD_ERR_ARG;
   }
   if (!dev0 && (!P->sparseKvalue || !P->sparseG1value || !P->sparseG2value || !P->sparseG3value || !P->sparseZvalue)) {
      set_error("cfd_problem: a matrix value array is NULL (cfd_options.step0_device = 1 computes them on the device)");
      return CFD_ERR_ARG;
   }
   if (!dev0 && P->scheme == CFD_EXPLICIT && (!P->MdInv || !P->MdOrigInv)) { set_error("explicit scheme needs MdInv and MdOrigInv"); return CFD_ERR_ARG; }
   if (opt && opt->operator_path == 1 && P->world > 1) { set_error("operator_path = 1 (matrix-free operators) is single-GPU: interface rows would need a sum exchange"); return CFD_ERR_ARG; }
   if (opt && opt->operator_path != 0 && opt->operator_path != 1) { set_error("operator_path must be 0 (assembled) or 1 (matrix-free)"); return CFD_ERR_ARG; }
   if (P->scheme == CFD_IMPLICIT && !dev0 && (!P->sparseMvalue || !P->sparseMapM_1d)) { set_error("implicit scheme needs sparseMvalue and sparseMapM_1d"); return CFD_ERR_ARG; }
   cfd_options o;
   if (opt) o = *opt; else cfd_default_options(&o);
   int ndev = 0;
   CFD_CK(cudaGetDeviceCount(&ndev));
   if (ndev == 0) { set_error("no CUDA device: this library has no CPU fallback"); return CFD_ERR_CUDA; }
   if (o.device >= 0) CFD_CK(cudaSetDevice(o.device));
   cfd_ctx* c = new cfd_ctx();
   *out = nullptr;
   // every early return below (CFD_CK, TRY, argument errors) releases what the context holds so far
   struct InitGuard { cfd_ctx* c; ~InitGuard() { if (c) cfd_destroy(c); } } guard{c};
   c->P = *P;
   c->opt = o;
   CFD_CK(cudaGetDevice(&c->device));
   cudaDeviceProp prop;
   CFD_CK(cudaGetDeviceProperties(&prop, c->device));
   c->numSMs = prop.multiProcessorCount;
   if (!prop.cooperativeLaunch) { set_error("device lacks cooperative launch"); return CFD_ERR_CUDA; }
   if (o.cuda_stream) { c->stream = (cudaStream_t)o.cuda_stream; c->ownStream = false; }
   else CFD_CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
   CFD_CK(cudaEventCreate(&c->ev0));
   CFD_CK(cudaEventCreate(&c->ev1));
   CFD_CK(cudaMallocHost((void**)&c->h_pinned, 256 * sizeof(double)));

   const int NN = P->NN, NNp = P->NNp, NE = P->NE, nv = P->NENv;
   // partition: NN / NNp are local (owned + ghost) counts, rows are the owned ones (numbered first)
   const bool parted = P->world > 1;
   const int NNo = (parted && P->NN_owned > 0) ? P->NN_owned : NN;
   const int NPo = (parted && P->NNp_owned > 0) ? P->NNp_owned : NNp;
   c->NNo = NNo; c->NPo = NPo;
   c->NEa = (parted && P->NE_apply > 0) ? P->NE_apply : NE;
   c->pin = parted ? P->pinLocal : (P->zeroPressureNode > 0 ? P->zeroPressureNode : -1);   // `node > 0` rule, BCH.cpp:4172
   c->monOwned = parted ? (P->monOwned != 0) : true;
   int rc;
#define TRY(x) do { rc = (x); if (rc) return rc; } while (0)

   const bool verb = getenv("CFD_VERBOSE") != nullptr;
   auto now = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
   double tph = now();
#define IPHASE(msg) do { if (verb) { cudaDeviceSynchronize(); double t_ = now(); fprintf(stderr, "[cfdb200 init] %-28s %8.2f s\n", msg, t_ - tph); tph = t_; } } while (0)
   // reference-element tables -> constant memory
   {
      FeTables T;
      build_fe_tables(27, T);
      CFD_CK(cudaMemcpyToSymbol(cSv27, T.Sv, sizeof(T.Sv)));
      CFD_CK(cudaMemcpyToSymbol(cdSv27, T.dSv, sizeof(T.dSv)));
      CFD_CK(cudaMemcpyToSymbol(cdSp, T.dSp, sizeof(T.dSp)));
      double sv8[8][8], dsv8[3][8][8];
      for (int k = 0; k < 8; k++)
         for (int i = 0; i < 8; i++) {
            sv8[k][i] = T.Sp[k][i];
            for (int d = 0; d < 3; d++) dsv8[d][i][k] = T.dSp[d][i][k];
         }
      CFD_CK(cudaMemcpyToSymbol(cSv8, sv8, sizeof sv8));
      CFD_CK(cudaMemcpyToSymbol(cdSv8, dsv8, sizeof dsv8));
   }
   // ---- matrices -> Sell.  The velocity-row matrices (K/M/A and G) share one row partition (same T).
   const int Tv = choose_Tv(NNo, P->sparseMrowStarts[NNo]);
   // Q1/Q1 on a whole mesh: M, G and the FE-Laplacian Z arrive on ONE pattern (same arrays): its layout is built once and shared
   SellHost hK, hG, hGT, hZ;   // (hG / hZ stay empty when they share hK)
   std::vector<int> gtSrc, gtRs, gtRow;   // device step0: transposed pattern of G (row starts, velocity node and G position of every entry)
   const bool sameG = P->sparseGrowStarts == P->sparseMrowStarts && P->sparseGcol == P->sparseMcol && NNp == NN && NPo == NNo;
   {
      SellHost& h = hK;
      sell_build_host(NNo, NN, P->sparseMrowStarts, P->sparseMcol, Tv, h);
      c->hostOffK = h.sliceOff;
      if (P->scheme == CFD_EXPLICIT) {
         const double* v[1] = {dev0 ? nullptr : P->sparseKvalue};
         TRY(sell_upload(h, 1, v, c->K));
      } else {
         const long long nnzOwned = P->sparseMrowStarts[NNo];
         std::unique_ptr<double[]> mk;
         if (!dev0) {
            mk.reset(new double[std::max<size_t>((size_t)nnzOwned, 1)]);
#pragma omp parallel for schedule(static)
            for (long long k = 0; k < nnzOwned; k++) mk[k] = P->sparseMvalue[k] + P->sparseKvalue[k];   // calculate_MplusK, GQ.cu:537-547 (constant: done once)
         }
         const double* v[3] = {mk.get(), dev0 ? nullptr : P->sparseMvalue, nullptr};
         TRY(sell_upload(h, 3, v, c->K));
         // element -> Sell position map, transposed to [q][e]
         c->ldE = (NE + 31) & ~31;
         const size_t nmt = (size_t)nv * nv * c->ldE;
         if (dev0 && !P->sparseMapM_1d) {   // built on the device by the step0 block below
            TRY(dalloc(c, &c->mapT, nmt));
            TRY(dalloc(c, &c->diagPosA, (size_t)NNo));
         } else {
         std::unique_ptr<int[]> mt(new int[nmt]);   // (1 G entries at 256^3: filled in parallel, padding columns included)
#pragma omp parallel for schedule(static)
         for (int e = 0; e < c->ldE; e++)
            for (int q = 0; q < nv * nv; q++) {
               if (e >= NE) { mt[(size_t)q * c->ldE + e] = 0; continue; }
               const int row = P->LtoGnode[(size_t)e * nv + q / nv];
               mt[(size_t)q * c->ldE + e] = row < NNo ? (int)h.perm[(size_t)P->sparseMapM_1d[(size_t)e * nv * nv + q]] : -1;
            }
         TRY(dupload(c, &c->mapT, mt.get(), nmt));
         std::vector<int> dp(NNo, 0);
#pragma omp parallel for schedule(static)
         for (int r = 0; r < NNo; r++)
            for (int k = P->sparseMrowStarts[r]; k < P->sparseMrowStarts[r + 1]; k++)
               if (P->sparseMcol[k] == r) { dp[r] = (int)h.perm[k]; break; }
         TRY(dupload(c, &c->diagPosA, dp.data(), dp.size()));
         }
      }
   }
   IPHASE("Sell K/M/A + maps");
   {
      if (!sameG) sell_build_host(NNo, NNp, P->sparseGrowStarts, P->sparseGcol, Tv, hG);
      const double* v[3] = {dev0 ? nullptr : P->sparseG1value, dev0 ? nullptr : P->sparseG2value, dev0 ? nullptr : P->sparseG3value};
      TRY(sell_upload(sameG ? hK : hG, 3, v, c->G));
      if (!sameG && !dev0) hG = SellHost();
   }
   IPHASE("Sell G");
   {
      // explicit transpose of the G pattern: rows = pressure nodes (instead of the reference's
      // OPERATION_TRANSPOSE csrmv scatter, BCH.cu:1006-1009).  Parallel: count, prefix sum, scatter in any order, then every
      // transposed row is put into ascending column order (deterministic result).
      const int nnzG = P->sparseG_nnz;
      std::vector<int> ts, tcol, tsrc((size_t)nnzG);
      std::unique_ptr<double[]> t1, t2, t3;
      if (!dev0) { t1.reset(new double[(size_t)nnzG]); t2.reset(new double[(size_t)nnzG]); t3.reset(new double[(size_t)nnzG]); }   // (no serial zero fill)
      // Q1/Q1 on a whole mesh: G lives on the (structurally symmetric) K / M pattern, so G^T has the SAME pattern and -- with the
      // same lanes per row -- the same device layout; only the values move: entry (i, j) of G^T = entry (j, i) of G, found by
      // bisection in row j.  (Any miss -- an unsymmetric or unsorted pattern from a caller -- falls back to the general transpose.)
      bool gtShared = false;
      if (sameG && choose_T(NPo, P->sparseGrowStarts[NPo]) == Tv && getenv("CFD_NO_GT_SHARED") == nullptr) {
         const int* rs = P->sparseGrowStarts; const int* cl = P->sparseGcol;
         int miss = 0;
#pragma omp parallel for schedule(static) reduction(+ : miss)
         for (int r = 0; r < NN; r++)
            for (int k = rs[r]; k < rs[r + 1]; k++) {
               if (k > rs[r] && cl[k] <= cl[k - 1]) miss++;
               const int j = cl[k];
               const int* q = std::lower_bound(cl + rs[j], cl + rs[j + 1], r);
               if (q == cl + rs[j + 1] || *q != r) { miss++; tsrc[k] = 0; } else tsrc[k] = (int)(q - cl);
            }
         gtShared = miss == 0;
      }
      if (gtShared) {
         if (!dev0) {
#pragma omp parallel for schedule(static)
            for (int k = 0; k < nnzG; k++) { t1[k] = P->sparseG1value[tsrc[k]]; t2[k] = P->sparseG2value[tsrc[k]]; t3[k] = P->sparseG3value[tsrc[k]]; }
         }
         const double* v[3] = {t1.get(), t2.get(), t3.get()};
         TRY(sell_upload(hK, 3, v, c->GT));
         if (dev0) { gtSrc.swap(tsrc); gtRs.assign(P->sparseGrowStarts, P->sparseGrowStarts + NNp + 1); gtRow.assign(P->sparseGcol, P->sparseGcol + nnzG); }
      } else {
      ts.assign((size_t)NNp + 1, 0); tcol.resize((size_t)nnzG);
#pragma omp parallel for schedule(static)
      for (int k = 0; k < nnzG; k++) {
#pragma omp atomic
         ts[(size_t)P->sparseGcol[k] + 1]++;
      }
      for (int i = 0; i < NNp; i++) ts[i + 1] += ts[i];
      {
         std::vector<int> fill(ts.begin(), ts.end() - 1);
#pragma omp parallel for schedule(static)
         for (int r = 0; r < NN; r++)
            for (int k = P->sparseGrowStarts[r]; k < P->sparseGrowStarts[r + 1]; k++) {
               int p;
#pragma omp atomic capture
               p = fill[P->sparseGcol[k]]++;
               tcol[p] = r; tsrc[p] = k;
            }
      }
#pragma omp parallel
      {
         std::vector<std::pair<int, int> > row;
#pragma omp for schedule(static)
         for (int j = 0; j < NNp; j++) {
            row.clear();
            for (int p = ts[j]; p < ts[j + 1]; p++) row.push_back(std::make_pair(tcol[p], tsrc[p]));
            std::sort(row.begin(), row.end());
            for (int p = ts[j], q = 0; p < ts[j + 1]; p++, q++) {
               const int k = row[q].second;
               tcol[p] = row[q].first;
               tsrc[p] = k;
               if (!dev0) { t1[p] = P->sparseG1value[k]; t2[p] = P->sparseG2value[k]; t3[p] = P->sparseG3value[k]; }
            }
         }
      }
      SellHost& h = hGT;
      sell_build_host(NPo, NN, ts.data(), tcol.data(), choose_T(NPo, ts[NPo]), h);   // rows = owned pressure DOFs, all local velocity columns
      const double* v[3] = {t1.get(), t2.get(), t3.get()};
      TRY(sell_upload(h, 3, v, c->GT));
      if (dev0) { gtSrc.swap(tsrc); gtRs.swap(ts); gtRow.swap(tcol); } else hGT = SellHost();
      }
   }
   IPHASE("Sell G^T");
   {
      const int Tz = choose_T(NPo, P->sparseZrowStarts[NPo]);
      const bool sameZ = sameG && Tz == Tv && P->sparseZrowStarts == P->sparseMrowStarts && P->sparseZcol == P->sparseMcol;
      if (!sameZ) sell_build_host(NPo, NNp, P->sparseZrowStarts, P->sparseZcol, Tz, hZ);
      SellHost& h = sameZ ? hK : hZ;
      c->hostOffZ = h.sliceOff;
      const double* v[1] = {dev0 ? nullptr : P->sparseZvalue};
      TRY(sell_upload(h, 1, v, c->Z));
      if (dev0) {
         TRY(dalloc(c, &c->Zdinv, (size_t)NPo));   // filled by the device step0 below
      } else {
         std::vector<double> dinv(NPo, 1.0);
#pragma omp parallel for schedule(static)
         for (int r = 0; r < NPo; r++)
            for (int k = P->sparseZrowStarts[r]; !o.cg_no_precond && k < P->sparseZrowStarts[r + 1]; k++)
               if (P->sparseZcol[k] == r) { dinv[r] = 1.0 / P->sparseZvalue[k]; break; }
         TRY(dupload(c, &c->Zdinv, dinv.data(), dinv.size()));
         if (!sameZ) hZ = SellHost();
      }
   }
   IPHASE("Sell Z");
   // ---- mesh tables
   {
      if (!c->ldE) c->ldE = (NE + 31) & ~31;
      std::vector<int> lt((size_t)nv * c->ldE, 0);
#pragma omp parallel for schedule(static)
      for (int e = 0; e < NE; e++)
         for (int i = 0; i < nv; i++) lt[(size_t)i * c->ldE + e] = P->LtoGnode[(size_t)e * nv + i];
      TRY(dupload(c, &c->LtoGT, lt.data(), lt.size()));
      if (o.ebe_scatter == 1) {
         // A VALID greedy colouring over corner-sharing neighbours (every shared node of two hexahedra lies on a shared
         // face/edge/corner and so implies a shared corner).  Unlike setupMeshColoring (BCH.cpp:853-947) the scan restarts
         // from the first neighbour after a clash, so no same-coloured neighbours can remain.
         const int NEa = c->NEa;
         std::vector<int> start((size_t)NN + 1, 0), items((size_t)NEa * 8), color(NEa, -1);
         for (int e = 0; e < NEa; e++)
            for (int i = 0; i < 8; i++) start[P->LtoGnode[(size_t)e * nv + i] + 1]++;
         for (int i = 0; i < NN; i++) start[i + 1] += start[i];
         {
            std::vector<int> fill(start.begin(), start.end() - 1);
            for (int e = 0; e < NEa; e++)
               for (int i = 0; i < 8; i++) items[fill[P->LtoGnode[(size_t)e * nv + i]]++] = e;
         }
         int ncol = 0;
         for (int e = 0; e < NEa; e++) {
            unsigned long long used = 0;
            for (int i = 0; i < 8; i++) {
               const int node = P->LtoGnode[(size_t)e * nv + i];
               for (int t = start[node]; t < start[node + 1]; t++)
                  if (color[items[t]] >= 0) used |= 1ull << color[items[t]];
            }
            int cc = 0;
            while (cc < 63 && (used >> cc & 1ull)) cc++;
            color[e] = cc;
            ncol = std::max(ncol, cc + 1);
         }
         c->colorStart.assign((size_t)ncol + 1, 0);
         for (int e = 0; e < NEa; e++) c->colorStart[color[e] + 1]++;
         for (int k = 0; k < ncol; k++) c->colorStart[k + 1] += c->colorStart[k];
         std::vector<int> eoc(NEa), fill(c->colorStart.begin(), c->colorStart.end() - 1);
         for (int e = 0; e < NEa; e++) eoc[fill[color[e]]++] = e;
         TRY(dupload(c, &c->elemOfColor, eoc.data(), eoc.size()));
      }
      TRY(dupload(c, &c->coord, P->coord, (size_t)NN * 3));
      std::vector<int> bi(NN, -1), bn;
      for (int i = 0; i < P->BCnVelNodes; i++) {
         const int node = P->BCvelNodes[2 * i];
         bi[node] = P->BCvelNodes[2 * i + 1];
         if (node < NNo) bn.push_back(node);   // identity rows only for owned Dirichlet nodes
      }
      c->P.BCnVelNodes = (int)bn.size();
      TRY(dupload(c, &c->bcIndex, bi.data(), bi.size()));
      TRY(dupload(c, &c->bcNodes, bn.data(), bn.size()));
      TRY(dupload(c, &c->BCstr, P->BCstr, (size_t)3 * P->nBC));
   }
   // ---- vectors
   const long long N3 = 3LL * NN;
   if (P->scheme == CFD_EXPLICIT) {
      if (dev0) { TRY(dalloc(c, &c->MdInv, (size_t)N3)); TRY(dalloc(c, &c->MdOrigInv, (size_t)N3)); }
      else { TRY(dupload(c, &c->MdInv, P->MdInv, (size_t)N3)); TRY(dupload(c, &c->MdOrigInv, P->MdOrigInv, (size_t)N3)); }
      TRY(regvec(c, "Un", &c->Un, N3)); TRY(regvec(c, "Unp1", &c->Unp1, N3)); TRY(regvec(c, "Unp1_prev", &c->Unp1_prev, N3));
      TRY(regvec(c, "UnpHalf", &c->UnpHalf, N3)); TRY(regvec(c, "UnpHalf_prev", &c->UnpHalf_prev, N3));
      TRY(regvec(c, "Acc", &c->Acc, N3)); TRY(regvec(c, "Acc_prev", &c->Acc_prev, N3));
      TRY(regvec(c, "KtimesAcc_prev", &c->KtimesAcc_prev, N3)); TRY(regvec(c, "R1", &c->R1, N3)); TRY(regvec(c, "R3", &c->R3, N3));
      TRY(regvec(c, "Pn", &c->Pn, NNp)); TRY(regvec(c, "Pnp1", &c->Pnp1, NNp)); TRY(regvec(c, "Pnp1_prev", &c->Pnp1_prev, NNp));
      TRY(regvec(c, "Pdot", &c->Pdot, NNp)); TRY(regvec(c, "R2", &c->R2, NNp));
   } else {
      TRY(regvec(c, "Uk", &c->Uk, N3)); TRY(regvec(c, "Uk_prev", &c->Uk_prev, N3));
      TRY(regvec(c, "Pk", &c->Pk, NNp)); TRY(regvec(c, "Pk_prev", &c->Pk_prev, NNp)); TRY(regvec(c, "Pk_prevprev", &c->Pk_prevprev, NNp));
      TRY(regvec(c, "Pdiff", &c->Pdiff, NNp)); TRY(regvec(c, "R2", &c->R2, NNp)); TRY(regvec(c, "R1", &c->R1i, N3));
      TRY(dalloc(c, &c->Adinv, (size_t)NN));
      TRY(dalloc(c, &c->bw, (size_t)(8 * N3)));
   }
   TRY(dalloc(c, &c->cg_r, (size_t)NNp)); TRY(dalloc(c, &c->cg_p, (size_t)NNp)); TRY(dalloc(c, &c->cg_q, (size_t)NNp));
   // pressure CG recurrences (cfd_options.cg_variant): the single-reduction form where synchronisation dominates
   // (default: classical.  Measured on B200, profiles/r02_cg_variants.md: the single-reduction kernel saves a synchronisation per
   // iteration but its larger vector set falls out of L2 and its sweep runs at 32 instead of 48 warps per SM -- slower at every size tried)
   c->cg_variant = o.cg_variant == 2 ? 2 : 1;
   if (c->cg_variant == 2) { TRY(dalloc(c, &c->cg_s, (size_t)NNp)); TRY(dalloc(c, &c->cg_pp, (size_t)NNp)); }
   TRY(dalloc(c, &c->kinfo, 64)); TRY(dalloc(c, &c->scal, 16));
   // initial state (applyBC_initial, BCH.cpp:4062-4110 / GQ.cpp:4559-4607)
   {
      std::vector<double> u0((size_t)N3, 0.0);
      if (P->U0) std::copy(P->U0, P->U0 + N3, u0.begin());
      else
         for (int i = 0; i < P->BCnVelNodes; i++) {
            const int node = P->BCvelNodes[2 * i], w = P->BCvelNodes[2 * i + 1];
            for (int d = 0; d < 3; d++) u0[(size_t)d * NN + node] = P->BCstr[3 * w + d];
         }
      double* Ud = P->scheme == CFD_EXPLICIT ? c->Un : c->Uk_prev;
      double* Pd = P->scheme == CFD_EXPLICIT ? c->Pn : c->Pk_prev;
      CFD_CK(cudaMemcpy(Ud, u0.data(), sizeof(double) * N3, cudaMemcpyHostToDevice));
      if (P->P0) CFD_CK(cudaMemcpy(Pd, P->P0, sizeof(double) * NNp, cudaMemcpyHostToDevice));
   }
   IPHASE("tables + vectors");
   // ---- step0 on the device (cfd_options.step0_device = 1): the value arrays are computed here from the mesh and the patterns
   std::vector<double> zHost;   // host copy of Z's values for the resident-PCG builder below (small systems only)
   if (dev0) {
      DevScratch tmp;
      tmp.st = c->stream;
      CFD_CK(cudaDeviceSynchronize());   // the zero fills and uploads above ran on the legacy stream; the kernels below run on c->stream
      const size_t nnzM = (size_t)P->sparseMrowStarts[NN], nnzG = (size_t)P->sparseGrowStarts[NN], nnzZ = (size_t)P->sparseZrowStarts[NNp];
      const bool sameZp = P->sparseZrowStarts == P->sparseMrowStarts && P->sparseZcol == P->sparseMcol;
      {
         // the position search is a bisection: columns must ascend within every row (the product's own patterns do)
         auto sorted = [](int nrows, const int* rs, const int* col) {
            int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
            for (int r = 0; r < nrows; r++)
               for (int k = rs[r] + 1; k < rs[r + 1]; k++) bad += col[k] <= col[k - 1];
            return bad == 0;
         };
         if (!sorted(NN, P->sparseMrowStarts, P->sparseMcol) || (!sameG && !sorted(NN, P->sparseGrowStarts, P->sparseGcol)) ||
             (!sameZp && !sorted(NNp, P->sparseZrowStarts, P->sparseZcol))) {
            set_error("step0_device = 1 needs ascending column indices within every CSR row");
            return CFD_ERR_ARG;
         }
      }
      int *dMrs = nullptr, *dMcol = nullptr, *dGrs = nullptr, *dGcol = nullptr, *dZrs = nullptr, *dZcol = nullptr;
      TRY(tmp.up(&dMrs, P->sparseMrowStarts, (size_t)NN + 1)); TRY(tmp.up(&dMcol, P->sparseMcol, nnzM));
      if (sameG) { dGrs = dMrs; dGcol = dMcol; } else { TRY(tmp.up(&dGrs, P->sparseGrowStarts, (size_t)NN + 1)); TRY(tmp.up(&dGcol, P->sparseGcol, nnzG)); }
      if (sameZp) { dZrs = dMrs; dZcol = dMcol; } else { TRY(tmp.up(&dZrs, P->sparseZrowStarts, (size_t)NNp + 1)); TRY(tmp.up(&dZcol, P->sparseZcol, nnzZ)); }
      double *dM = nullptr, *dK = nullptr, *dG[3] = {nullptr, nullptr, nullptr}, *dZ = nullptr, *dmd = nullptr;
      TRY(tmp.zeros(&dM, nnzM)); TRY(tmp.zeros(&dK, nnzM));
      for (int d = 0; d < 3; d++) TRY(tmp.zeros(&dG[d], nnzG));
      TRY(tmp.zeros(&dZ, nnzZ));
      const bool imp = P->scheme == CFD_IMPLICIT;
      const double mscale = imp ? 1.0 / P->dt : 1.0;   // implicit M carries 1/dt (GQ.cpp:3495)
      {
         const dim3 grid((unsigned)gridFor(NE, 128), (unsigned)nv);
         if (nv == 27)
            k_step0_elements<27><<<grid, 128, 0, c->stream>>>(NE, c->ldE, c->LtoGT, c->coord, dMrs, dMcol, dGrs, dGcol, dZrs, dZcol, mscale, P->viscosity,
                                                              1.0 / P->density, dM, dK, dG[0], dG[1], dG[2], imp ? dZ : nullptr);
         else
            k_step0_elements<8><<<grid, 128, 0, c->stream>>>(NE, c->ldE, c->LtoGT, c->coord, dMrs, dMcol, dGrs, dGcol, dZrs, dZcol, mscale, P->viscosity,
                                                             1.0 / P->density, dM, dK, dG[0], dG[1], dG[2], imp ? dZ : nullptr);
         c->launches++;
         CFD_CK(cudaGetLastError());
      }
      int* dGtSrc = nullptr;
      TRY(tmp.up(&dGtSrc, gtSrc.data(), nnzG));
      if (!imp) {
         TRY(tmp.zeros(&dmd, (size_t)NN));
         k_step0_lump<<<gridFor(NN, 256), 256, 0, c->stream>>>(NN, dMrs, dM, c->bcIndex, dmd, c->MdOrigInv, c->MdInv);
         int *dTs = nullptr, *dTrow = nullptr;
         TRY(tmp.up(&dTs, gtRs.data(), (size_t)NNp + 1)); TRY(tmp.up(&dTrow, gtRow.data(), nnzG));
         k_step0_spgemm_z<<<gridFor((long long)NNp * 32, 256), 256, 0, c->stream>>>(NNp, dTs, dTrow, dGtSrc, dGrs, dGcol, dG[0], dG[1], dG[2], dmd, dZrs, dZcol, dZ);
         c->launches += 2;
         CFD_CK(cudaGetLastError());
      }
      k_step0_diag<<<gridFor(NPo, 256), 256, 0, c->stream>>>(NPo, dZrs, dZcol, dZ, c->pin, o.cg_no_precond ? 1 : 0, c->Zdinv);
      c->launches++;
      // CSR order -> Sell layout (padding entries stay zero)
      auto scatter = [&](const std::vector<long long>& perm, long long** dperm, size_t n, const double* a, const double* b, double* dst) -> int {
         if (!*dperm) { int rc2 = tmp.up(dperm, perm.data(), n); if (rc2) return rc2; }
         k_step0_scatter<<<gridFor((long long)n, 256), 256, 0, c->stream>>>((long long)n, *dperm, a, b, dst);
         c->launches++;
         CFD_CK(cudaGetLastError());
         return CFD_OK;
      };
      long long *dPermK = nullptr, *dPermG = nullptr, *dPermGT = nullptr, *dPermZ = nullptr;
      if (imp && !P->sparseMapM_1d) {
         TRY(tmp.up(&dPermK, hK.perm.data(), nnzM));
         const dim3 grid((unsigned)gridFor(c->ldE, 128), (unsigned)(nv * nv));
         if (nv == 27) k_step0_map<27><<<grid, 128, 0, c->stream>>>(NE, c->ldE, c->LtoGT, dMrs, dMcol, dPermK, c->mapT);
         else k_step0_map<8><<<grid, 128, 0, c->stream>>>(NE, c->ldE, c->LtoGT, dMrs, dMcol, dPermK, c->mapT);
         k_step0_diagpos<<<gridFor(NNo, 256), 256, 0, c->stream>>>(NNo, dMrs, dMcol, dPermK, c->diagPosA);
         c->launches += 2;
         CFD_CK(cudaGetLastError());
      }
      if (imp) {
         TRY(scatter(hK.perm, &dPermK, nnzM, dM, dK, c->K.vals[0]));    // M + K (calculate_MplusK, GQ.cu:537-547)
         TRY(scatter(hK.perm, &dPermK, nnzM, dM, nullptr, c->K.vals[1]));
      } else {
         TRY(scatter(hK.perm, &dPermK, nnzM, dK, nullptr, c->K.vals[0]));
      }
      if (sameG) dPermG = dPermK;
      for (int d = 0; d < 3; d++) TRY(scatter(hG.perm, &dPermG, nnzG, dG[d], nullptr, c->G.vals[d]));
      {
         double* dT = nullptr;
         TRY(tmp.alloc((void**)&dT, sizeof(double) * nnzG));
         for (int d = 0; d < 3; d++) {
            k_step0_gather<<<gridFor((long long)nnzG, 256), 256, 0, c->stream>>>((long long)nnzG, dGtSrc, dG[d], dT);
            c->launches++;
            if (hGT.perm.empty()) dPermGT = dPermK;   // G^T shares the K / M / A layout (Q1/Q1 on a whole mesh)
            TRY(scatter(hGT.perm, &dPermGT, nnzG, dT, nullptr, c->GT.vals[d]));
         }
      }
      if (hZ.perm.empty()) dPermZ = dPermK;   // Z shares the K / M / A layout (Q1/Q1 on a whole mesh)
      TRY(scatter(hZ.perm, &dPermZ, nnzZ, dZ, nullptr, c->Z.vals[0]));
      CFD_CK(cudaStreamSynchronize(c->stream));
      if ((long long)nnzZ * 8 <= 232448LL * c->numSMs) {
         zHost.resize(nnzZ);
         CFD_CK(cudaMemcpy(zHost.data(), dZ, sizeof(double) * nnzZ, cudaMemcpyDeviceToHost));
      }
      hG = SellHost(); hGT = SellHost(); hZ = SellHost();
      IPHASE("step0 on the device");
   }
   // ---- partition plan
   if (parted) {
      const int np = P->nPeers;
      c->peerRank.assign(P->peerRank, P->peerRank + np);
      c->sendVptr.assign(P->sendVptr, P->sendVptr + np + 1); c->sendPptr.assign(P->sendPptr, P->sendPptr + np + 1);
      c->recvVoff.assign(P->recvVoff, P->recvVoff + np); c->recvVcnt.assign(P->recvVcnt, P->recvVcnt + np);
      c->recvPoff.assign(P->recvPoff, P->recvPoff + np); c->recvPcnt.assign(P->recvPcnt, P->recvPcnt + np);
      const int nsV = c->sendVptr[np], nsP = c->sendPptr[np];
      // device copies: idx, segment-of-entry, ptr  (velocity plan then pressure plan)
      std::vector<int> segV(nsV), segP(nsP);
      for (int j = 0; j < np; j++) {
         for (int k = c->sendVptr[j]; k < c->sendVptr[j + 1]; k++) segV[k] = j;
         for (int k = c->sendPptr[j]; k < c->sendPptr[j + 1]; k++) segP[k] = j;
      }
      std::vector<int> packV, packP;
      packV.insert(packV.end(), P->sendVidx, P->sendVidx + nsV); packV.insert(packV.end(), segV.begin(), segV.end());
      packV.insert(packV.end(), c->sendVptr.begin(), c->sendVptr.end());
      packP.insert(packP.end(), P->sendPidx, P->sendPidx + nsP); packP.insert(packP.end(), segP.begin(), segP.end());
      packP.insert(packP.end(), c->sendPptr.begin(), c->sendPptr.end());
      TRY(dupload(c, &c->sendVidx, packV.data(), packV.size()));
      TRY(dupload(c, &c->sendPidx, packP.data(), packP.size()));
      TRY(dalloc(c, &c->sendBuf, (size_t)std::max(3 * nsV, nsP) + 1));
      TRY(dalloc(c, &c->dacc, 64));
      c->remoteVoff.assign(P->remoteVoff ? P->remoteVoff : P->recvVoff, (P->remoteVoff ? P->remoteVoff : P->recvVoff) + np);
      c->remotePoff.assign(P->remotePoff ? P->remotePoff : P->recvPoff, (P->remotePoff ? P->remotePoff : P->recvPoff) + np);
      // exchange window for the peer-memory path: flags | mailboxes | p  (p moves into the window)
      c->winOffP = kWinHeader;
      c->winOffV = kWinHeader + (long long)sizeof(double) * (((long long)NNp + 15) & ~15LL);
      c->winBytes = (size_t)c->winOffV + (P->scheme == CFD_IMPLICIT ? sizeof(double) * 6 * (size_t)NN : 0);
      CFD_CK(cudaMalloc((void**)&c->win, c->winBytes));
      CFD_CK(cudaMemset(c->win, 0, c->winBytes));
      c->cg_p = (double*)(c->win + c->winOffP);
      if (P->scheme == CFD_IMPLICIT) { c->bw_ph = (double*)(c->win + c->winOffV); c->bw_shv = c->bw_ph + 3 * (size_t)NN; }
      TRY(dalloc(c, &c->xstate, 8));
      if (const char* e = getenv("CFD_P2P_TIMEOUT_MS")) c->p2pTimeoutNs = (unsigned long long)std::max(1.0, atof(e)) * 1000000ull;
   }
   // ---- cooperative launch geometry: persistent kernels, a whole number of CTAs per SM
   {
      int occ = 0;
      if (c->cg_variant == 2) c->pcg_minb = 2;   // the single-reduction kernel needs 64 registers to stay free of spills
      if (const char* e = getenv("CFD_PCG_OCC")) c->pcg_minb = atoi(e) == 2 ? 2 : 3;
      const void* pcgK = c->cg_variant == 2 ? pcgcg_kernel(parted, c->pcg_minb) : (parted ? (const void*)k_pcg_p2p : (const void*)pcg_kernel(c->pcg_minb));
      CFD_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pcgK, c->pcg_threads, 0));
      if (occ < 1) { set_error("k_pcg does not fit on an SM"); return CFD_ERR_CUDA; }
      const long long wantWarps = std::max<long long>(c->Z.nslices, ((long long)NNp + 31) / 32);
      int blocks = c->numSMs * occ;
      const int needed = (int)std::max<long long>(1, (wantWarps * 32 + c->pcg_threads - 1) / c->pcg_threads);
      if (needed < blocks) blocks = std::max(c->numSMs, ((needed + c->numSMs - 1) / c->numSMs) * c->numSMs);
      c->pcg_blocks = std::min(blocks, c->numSMs * occ);
      if (const char* e = getenv("CFD_PCG_BLOCKS")) c->pcg_blocks = std::max(1, std::min(atoi(e), c->numSMs * occ));
      int occ2 = 0;
      CFD_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, parted ? k_bicgstab3<true> : k_bicgstab3<false>, c->bicg_threads, 0));
      if (occ2 < 1) { set_error("k_bicgstab3 does not fit on an SM"); return CFD_ERR_CUDA; }
      const int needed2 = (int)std::max<long long>(1, ((long long)c->K.nslices * 32 + c->bicg_threads - 1) / c->bicg_threads);
      int blocks2 = c->numSMs * occ2;
      if (needed2 < blocks2) blocks2 = std::max(c->numSMs, ((needed2 + c->numSMs - 1) / c->numSMs) * c->numSMs);
      c->bicg_blocks = std::min(blocks2, c->numSMs * occ2);
      TRY(dalloc(c, &c->partials, (size_t)3 * 6 * std::max(c->pcg_blocks, c->bicg_blocks)));
      auto partition = [](const std::vector<long long>& off, int nblk) {
         const int ns = (int)off.size() - 1;
         std::vector<int> part(nblk + 1, ns);
         const long long total = off[ns];
         int s = 0;
         for (int b = 0; b < nblk; b++) {
            const long long target = (long long)((double)total * b / nblk);
            while (s < ns && off[s] < target) s++;
            part[b] = s;
         }
         part[0] = 0;
         return part;
      };
      {
         // static share = the first ~88 % of the stored entries, balanced over the CTAs; the rest is the dynamic tail
         const int nsZ = (int)c->hostOffZ.size() - 1;
         int dynStart = nsZ;
         const bool useDyn = !o.deterministic && getenv("CFD_NO_DYN") == nullptr && nsZ >= 4 * c->pcg_blocks * (c->pcg_threads / 32);
         if (useDyn) {
            double frac = 0.88;
            if (const char* e = getenv("CFD_DYN_STATIC_FRAC")) frac = std::min(1.0, std::max(0.5, atof(e)));
            const long long target = (long long)(frac * (double)c->hostOffZ[nsZ]);
            dynStart = (int)(std::lower_bound(c->hostOffZ.begin(), c->hostOffZ.end(), target) - c->hostOffZ.begin());
            dynStart = std::min(std::max(dynStart, 0), nsZ);
         }
         std::vector<long long> offStatic(c->hostOffZ.begin(), c->hostOffZ.begin() + dynStart + 1);
         std::vector<int> pz = partition(offStatic, c->pcg_blocks);
         TRY(dupload(c, &c->pcgPart, pz.data(), pz.size()));
         c->pcgcgPart = c->pcgPart; c->pcgcg_blocks = c->pcg_blocks;
         c->pcgDynStart = dynStart;
         if (useDyn) TRY(dalloc(c, &c->pcgDynCounter, 4));
         // Shared-memory-resident matrix (k_pcg_smem): one CTA per SM takes a contiguous slice range whose values fit into its
         // shared memory (its 32 warps deal the slices among themselves).  Greedy fill in slice order with an even byte target; not
         // eligible if the ranges do not cover the matrix (too large) -- then the streaming kernel above runs.
         // k_pcg_res: row blocks resident per CTA (see the kernel).  Built from the CSR arrays of Z; eligible when every CTA's block
         // (values + staged vector entries + its rows of w) fits into its shared memory.  cfd_options.cg_variant = 1 (classical
         // recurrences asked for explicitly) keeps k_pcg_smem<classical> / k_pcg_t.
         if (!parted && o.cg_variant != 1 && getenv("CFD_NO_PCG_SMEM") == nullptr && getenv("CFD_NO_PCG_RES") == nullptr && NPo > 0) {
            int maxOptin = 0;
            CFD_CK(cudaDeviceGetAttribute(&maxOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
            const long long cap = (long long)maxOptin - 4 * 32 * 8 - 256;   // static shared memory of the kernel + slack
            const int* rs = P->sparseZrowStarts; const int* cl = P->sparseZcol; const double* zv = dev0 ? zHost.data() : P->sparseZvalue;
            const int nC = c->numSMs, n = NPo;
            const long long nnzZ = rs[n];
            if (nnzZ * 8 <= cap * nC && (!dev0 || (long long)zHost.size() == nnzZ)) {
               // lanes per row: fewest sequential entry pairs per warp (groups dealt round-robin to the 32 warps of a CTA)
               const double avgLen = (double)nnzZ / n, rowsPer = std::max(1.0, (double)n / nC);
               int bestT = 8; double bestCost = 1e300;
               for (int T = 1; T <= 32; T *= 2) {
                  const double groups = std::ceil(rowsPer / (32 / T)), rounds = std::ceil(groups / (double)kResWarps);
                  const double pairs = std::ceil(std::ceil(avgLen / T) / 2.0);
                  // a warp's first group keeps up to 32 column-id pairs in registers; anything beyond is re-requested from L2 every sweep
                  const double cost = rounds * (pairs + 0.25 * std::log2((double)T)) + (rounds > 1.0 || pairs > 32.0 ? 24.0 : 0.0);
                  if (cost < bestCost) { bestCost = cost; bestT = T; }
               }
               if (const char* e = getenv("CFD_PCG_RES_T")) { const int t = atoi(e); if (t == 1 || t == 2 || t == 4 || t == 8 || t == 16 || t == 32) bestT = t; }
               const int T = bestT, R = 32 / T;
               std::vector<ResCta> cta(nC);
               std::vector<int> grp, cols;
               std::vector<std::vector<int> > ctaCols(nC);
               long long totalEntries = 0, maxBytes = 0;
               bool ok = true;
               int r = 0;
               // Row blocks: contiguous, an even share of the STORED entries each -- shared-memory traffic is what a sweep costs.  (Capping
               // the rows of a block at one row group per warp, so that all column ids stay in registers, does not work out: blocks of
               // short boundary rows then take less than their share and the last block overflows its shared memory.)
               {
                  const int rowCap = kResThreads;
                  auto rowEntries = [&](int row) { return (long long)((((rs[row + 1] - rs[row] + T - 1) / T) + 1) & ~1) * T; };   // (estimate: the row's own padded width)
                  long long remaining = 0;
                  for (int row = 0; row < n; row++) remaining += rowEntries(row);
                  for (int b = 0; b < nC; b++) {
                     ResCta& k = cta[b];
                     k.row0 = r;
                     const long long target = remaining / (nC - b);
                     long long took = 0;
                     while (r < n && (b == nC - 1 || (r - k.row0 < rowCap && took + rowEntries(r) <= target + rowEntries(r) / 2))) { took += rowEntries(r); r++; }
                     remaining -= took;
                     k.nrows = r - k.row0;
                     k.ngroups = (k.nrows + R - 1) / R;
                     k.valOff = totalEntries; k.grpOff = (long long)grp.size(); k.colOff = (long long)cols.size();
                     std::vector<int>& u = ctaCols[b];
                     u.clear();
                     for (int q = rs[k.row0]; q < rs[r]; q++) u.push_back(cl[q]);
                     std::sort(u.begin(), u.end());
                     u.erase(std::unique(u.begin(), u.end()), u.end());
                     k.ncols = (int)u.size();
                     cols.insert(cols.end(), u.begin(), u.end());
                     int off = 0;
                     for (int g = 0; g < k.ngroups; g++) {
                        grp.push_back(off);
                        int w = 0;
                        for (int rl = 0; rl < R && k.row0 + g * R + rl < r; rl++) {
                           const int rr2 = k.row0 + g * R + rl;
                           w = std::max(w, (rs[rr2 + 1] - rs[rr2] + T - 1) / T);
                        }
                        off += ((w + 1) & ~1) * 32;
                     }
                     grp.push_back(off);
                     totalEntries += off;
                     const long long bytes = 8LL * off + 8LL * ((k.ncols + 1) & ~1) + 8LL * k.nrows;
                     maxBytes = std::max(maxBytes, bytes);
                     if (k.nrows > kResThreads || k.ncols > 65535) ok = false;
                  }
               }
               if (ok && maxBytes <= cap && r == n) {
                  std::vector<double> vals((size_t)std::max<long long>(totalEntries, 2), 0.0);
                  std::vector<unsigned short> idx((size_t)std::max<long long>(totalEntries, 2), 0);
#pragma omp parallel for schedule(dynamic, 1)
                  for (int b = 0; b < nC; b++) {
                     const ResCta& k = cta[b];
                     const std::vector<int>& u = ctaCols[b];
                     for (int lr = 0; lr < k.nrows; lr++) {
                        const int g = lr / R, rl = lr % R, row = k.row0 + lr;
                        const long long base = k.valOff + grp[(size_t)k.grpOff + g];
                        for (int q = rs[row], e = 0; q < rs[row + 1]; q++, e++) {
                           const int t = e % T, j = e / T;
                           const long long pos = base + (long long)(j >> 1) * 64 + 2 * (rl * T + t) + (j & 1);
                           vals[(size_t)pos] = zv[q];
                           idx[(size_t)pos] = (unsigned short)(std::lower_bound(u.begin(), u.end(), cl[q]) - u.begin());
                        }
                     }
                  }
                  const void* kres = (const void*)k_pcg_res;
                  CFD_CK(cudaFuncSetAttribute(kres, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)maxBytes));
                  int occR = 0;
                  CFD_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occR, kres, kResThreads, (size_t)maxBytes));
                  if (occR >= 1) {
                     ResCta* dcta = nullptr;
                     TRY(dupload(c, &dcta, cta.data(), cta.size()));
                     c->resCta = dcta;
                     TRY(dupload(c, &c->resVals, vals.data(), vals.size()));
                     TRY(dupload(c, &c->resIdx, idx.data(), idx.size()));
                     if (cols.empty()) cols.push_back(0);
                     TRY(dupload(c, &c->resCols, cols.data(), cols.size()));
                     TRY(dupload(c, &c->resGrp, grp.data(), grp.size()));
                     c->resT = T;
                     c->resBytes = (size_t)maxBytes;
                     c->cg_variant = 2;   // (byte model of the iteration)
                     if (verb)
                        fprintf(stderr, "[cfdb200 init] k_pcg_res: %d rows in %d row blocks, T = %d lanes per row, %.1f MB of values, at most %.1f KB of shared memory per CTA\n",
                                n, nC, T, 8e-6 * (double)totalEntries, maxBytes / 1024.0);
                  }
               }
            }
         }
         if (!parted && !c->resBytes && getenv("CFD_NO_PCG_SMEM") == nullptr) {
            int maxOptin = 0;
            CFD_CK(cudaDeviceGetAttribute(&maxOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
            const long long cap = (long long)maxOptin - 1024 - 3 * 32 * 8;   // static shared memory of the kernel + slack
            const long long total = c->hostOffZ[nsZ] * 8;
            if (total <= cap * c->numSMs) {
               std::vector<int> ps(c->numSMs + 1, nsZ);
               long long maxBytes = 0;
               int s = 0;
               bool ok = true;
               for (int b = 0; b < c->numSMs; b++) {
                  ps[b] = s;
                  const long long target = (long long)((double)c->hostOffZ[nsZ] * (b + 1) / c->numSMs);   // even share, in entries
                  const int sBegin = s;
                  while (s < nsZ && (c->hostOffZ[s + 1] - c->hostOffZ[sBegin]) * 8 <= cap &&
                         (c->hostOffZ[s + 1] <= target || s == sBegin)) s++;
                  maxBytes = std::max(maxBytes, (c->hostOffZ[s] - c->hostOffZ[sBegin]) * 8);
               }
               ps[c->numSMs] = s;
               ok = s == nsZ;
               if (ok) {
                  // recurrences inside the resident kernel: cfd_options.cg_variant 1 = classical, 2 = single reduction, 0 = the library's
                  // choice = single reduction (two synchronisations per iteration; here nothing else costs anything)
                  c->pcgSmemSingle = o.cg_variant != 1;
                  const void* kk = c->pcgSmemSingle ? (const void*)k_pcg_smem<true> : (const void*)k_pcg_smem<false>;
                  CFD_CK(cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)maxBytes));
                  int occS = 0;
                  CFD_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occS, kk, 1024, (size_t)maxBytes));
                  if (occS >= 1) {
                     TRY(dupload(c, &c->pcgSmemPart, ps.data(), ps.size()));
                     c->pcgSmemBytes = (size_t)maxBytes;
                     if (c->pcgSmemSingle && !c->cg_s) { TRY(dalloc(c, &c->cg_s, (size_t)NNp)); TRY(dalloc(c, &c->cg_pp, (size_t)NNp)); }
                     c->cg_variant = c->pcgSmemSingle ? 2 : 1;   // (byte model of the iteration)
                  }
               }
            }
         }
      }
      if (verb)
         fprintf(stderr, "[cfdb200 init] pressure CG: %s, Z %d rows, %d slices (T = %d), %.1f MB of values%s\n",
                 c->resBytes ? "k_pcg_res (row blocks resident per CTA: values in shared memory, staged vector window, vectors in registers)" : c->pcgSmemBytes ? (c->pcgSmemSingle ? "k_pcg_smem<single reduction> (matrix resident in shared memory)" : "k_pcg_smem<classical> (matrix resident in shared memory)") : (c->cg_variant == 2 ? "k_pcg_cg (single reduction)" : "k_pcg_t (streaming)"),
                 c->Z.nrows, c->Z.nslices, c->Z.T, 8e-6 * (double)c->Z.padded, c->pcgSmemBytes ? "" : "");
      if (P->scheme == CFD_IMPLICIT) {
         // large matrices: static share (contiguous chunk per CTA, its gathered window stays in L1) + dynamic tail as for Z;
         // fewer than 4 slices per warp: everything through the counter (slice order = node order, long corner rows first)
         const int nsK = (int)c->hostOffK.size() - 1;
         const int nwarps = c->bicg_blocks * (c->bicg_threads / 32);
         int dynStart = nsK;
         const bool useDyn = !o.deterministic && getenv("CFD_NO_DYN") == nullptr && nsK > nwarps;
         if (useDyn) {
            dynStart = 0;
            if (nsK >= 4 * nwarps) {
               double frac = 0.88;
               if (const char* e = getenv("CFD_DYN_STATIC_FRAC")) frac = std::min(1.0, std::max(0.0, atof(e)));
               const long long target = (long long)(frac * (double)c->hostOffK[nsK]);
               dynStart = (int)(std::lower_bound(c->hostOffK.begin(), c->hostOffK.end(), target) - c->hostOffK.begin());
               dynStart = std::min(std::max(dynStart, 0), nsK);
            }
         }
         std::vector<long long> offStatic(c->hostOffK.begin(), c->hostOffK.begin() + dynStart + 1);
         std::vector<int> pk = partition(offStatic, c->bicg_blocks);
         TRY(dupload(c, &c->bicgPart, pk.data(), pk.size()));
         c->bicgDynStart = dynStart;
         if (useDyn) TRY(dalloc(c, &c->bicgDynCounter, 4));
      }
   }
#undef TRY
   // scalar fields stay; host pointers are not ours to keep
   c->P.coord = nullptr; c->P.LtoGnode = nullptr; c->P.sparseMrowStarts = nullptr; c->P.sparseMcol = nullptr;
   c->P.sparseKvalue = nullptr; c->P.sparseMvalue = nullptr; c->P.sparseGrowStarts = nullptr; c->P.sparseGcol = nullptr;
   c->P.sparseG1value = c->P.sparseG2value = c->P.sparseG3value = nullptr;
   c->P.sparseZrowStarts = nullptr; c->P.sparseZcol = nullptr; c->P.sparseZvalue = nullptr; c->P.sparseMapM_1d = nullptr;
   c->P.MdInv = c->P.MdOrigInv = nullptr; c->P.BCvelNodes = nullptr; c->P.BCstr = nullptr; c->P.U0 = c->P.P0 = nullptr;
   CFD_CK(cudaDeviceSynchronize());
   guard.c = nullptr;
   *out = c;
   return CFD_OK;
}

int cfd_nccl_unique_id(char out[128]) {
   Comm tmp;
   if (!nccl_load(tmp)) return CFD_ERR_NCCL;
   NcclUniqueId id;
   CFD_NCCL(&tmp, tmp.GetUniqueId(&id));
   memcpy(out, id.internal, 128);
   return CFD_OK;
}

int cfd_comm_init(cfd_ctx* c, const char id[128]) {
   if (!c || !id) return CFD_ERR_ARG;
   if (c->P.world <= 1) return CFD_OK;
   if (!c->comm) c->comm = new Comm();
   if (!nccl_load(*c->comm)) return CFD_ERR_NCCL;
   c->comm->world = c->P.world; c->comm->rank = c->P.rank;
   NcclUniqueId uid;
   memcpy(uid.internal, id, 128);
   CFD_CK(cudaSetDevice(c->device));
   CFD_NCCL(c->comm, c->comm->CommInitRank(&c->comm->comm, c->P.world, uid, c->P.rank));
   // initial halo of the state vectors (the host supplied owned + ghost values, but be safe)
   return CFD_OK;
}

int cfd_p2p_export(cfd_ctx* c, char desc[96]) {
   if (!c || !desc) return CFD_ERR_ARG;
   if (!c->win) { set_error("no exchange window (context is not partitioned)"); return CFD_ERR_STATE; }
   memset(desc, 0, 96);
   cudaIpcMemHandle_t h;
   CFD_CK(cudaIpcGetMemHandle(&h, c->win));
   memcpy(desc, &h, sizeof h);
   memcpy(desc + 64, &c->winOffP, sizeof(long long));
   memcpy(desc + 72, &c->winOffV, sizeof(long long));
   memcpy(desc + 80, &c->P.NN, sizeof(int));
   return CFD_OK;
}

int cfd_p2p_attach(cfd_ctx* c, const char* all) {
   if (!c || !all) return CFD_ERR_ARG;
   if (!c->win) { set_error("no exchange window (context is not partitioned)"); return CFD_ERR_STATE; }
   const int world = c->P.world, rank = c->P.rank;
   if (world > kMaxRanks) { set_error("peer-memory path supports up to %d ranks", (int)kMaxRanks); return CFD_ERR_ARG; }
   CFD_CK(cudaSetDevice(c->device));
   c->peerWin.assign(world, nullptr);
   c->peerOffP.assign(world, 0);
   c->peerOffV.assign(world, 0);
   c->peerNN.assign(world, 0);
   for (int q = 0; q < world; q++) {
      memcpy(&c->peerOffP[q], all + 96 * q + 64, sizeof(long long));
      memcpy(&c->peerOffV[q], all + 96 * q + 72, sizeof(long long));
      memcpy(&c->peerNN[q], all + 96 * q + 80, sizeof(int));
      if (q == rank) { c->peerWin[q] = c->win; continue; }
      cudaIpcMemHandle_t h;
      memcpy(&h, all + 96 * q, sizeof h);
      void* ptr = nullptr;
      CFD_CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
      c->peerWin[q] = (char*)ptr;
   }
   const int np = (int)c->peerRank.size();
   std::vector<double*> peerP(np);
   for (int j = 0; j < np; j++)
      peerP[j] = (double*)(c->peerWin[c->peerRank[j]] + c->peerOffP[c->peerRank[j]]) + c->remotePoff[j];
   // push list per owned pressure DOF: (peer index, slot inside my block of the peer's ghosts)
   const int NPo = c->NPo;
   std::vector<int> sendP((size_t)c->sendPptr[np]);
   CFD_CK(cudaMemcpy(sendP.data(), c->sendPidx, sizeof(int) * sendP.size(), cudaMemcpyDeviceToHost));
   std::vector<int> ptr((size_t)NPo + 1, 0);
   for (int t = 0; t < (int)sendP.size(); t++) ptr[sendP[t] + 1]++;
   for (int i = 0; i < NPo; i++) ptr[i + 1] += ptr[i];
   std::vector<int2> dst((size_t)ptr[NPo]);
   {
      std::vector<int> fill(ptr.begin(), ptr.end() - 1);
      for (int j = 0; j < np; j++)
         for (int t = c->sendPptr[j]; t < c->sendPptr[j + 1]; t++) dst[fill[sendP[t]]++] = make_int2(j, t - c->sendPptr[j]);
   }
   int rc;
   if ((rc = dupload(c, &c->dPeerWin, c->peerWin.data(), c->peerWin.size()))) return rc;
   if ((rc = dupload(c, &c->dPeerP, peerP.data(), peerP.size()))) return rc;
   if ((rc = dupload(c, &c->dPeerRank, c->peerRank.data(), c->peerRank.size()))) return rc;
   if ((rc = dupload(c, &c->pushPtrP, ptr.data(), ptr.size()))) return rc;
   if ((rc = dupload(c, &c->pushDstP, dst.data(), dst.size()))) return rc;
   if (c->P.scheme == CFD_IMPLICIT) {
      // the same for the velocity-row vectors of the momentum solve (ph | shv inside the windows)
      const int NNo = c->NNo;
      std::vector<double*> peerV(np);
      std::vector<int> peerLd(np);
      for (int j = 0; j < np; j++) {
         peerV[j] = (double*)(c->peerWin[c->peerRank[j]] + c->peerOffV[c->peerRank[j]]) + c->remoteVoff[j];
         peerLd[j] = c->peerNN[c->peerRank[j]];
      }
      std::vector<int> sendV((size_t)c->sendVptr[np]);
      CFD_CK(cudaMemcpy(sendV.data(), c->sendVidx, sizeof(int) * sendV.size(), cudaMemcpyDeviceToHost));
      std::vector<int> pv((size_t)NNo + 1, 0);
      for (int t = 0; t < (int)sendV.size(); t++) pv[sendV[t] + 1]++;
      for (int i = 0; i < NNo; i++) pv[i + 1] += pv[i];
      std::vector<int2> dv((size_t)pv[NNo]);
      {
         std::vector<int> fill(pv.begin(), pv.end() - 1);
         for (int j = 0; j < np; j++)
            for (int t = c->sendVptr[j]; t < c->sendVptr[j + 1]; t++) dv[fill[sendV[t]]++] = make_int2(j, t - c->sendVptr[j]);
      }
      if ((rc = dupload(c, &c->dPeerV, peerV.data(), peerV.size()))) return rc;
      if ((rc = dupload(c, &c->dPeerLd, peerLd.data(), peerLd.size()))) return rc;
      if ((rc = dupload(c, &c->pushPtrV, pv.data(), pv.size()))) return rc;
      if ((rc = dupload(c, &c->pushDstV, dv.data(), dv.size()))) return rc;
   }
   std::vector<int> iface;
   for (int i = 0; i < NPo; i++)
      if (ptr[i + 1] > ptr[i]) iface.push_back(i);
   c->nIfaceP = (int)iface.size();
   if ((rc = dupload(c, &c->ifaceP, iface.data(), iface.size()))) return rc;
   c->p2p = getenv("CFD_NO_P2P") == nullptr;
   return p2p_agree(c);
}

int cfd_p2p_detach(cfd_ctx* c) {
   if (!c) return CFD_ERR_ARG;
   CFD_CK(cudaSetDevice(c->device));
   if (c->stream) CFD_CK(cudaStreamSynchronize(c->stream));
   for (size_t q = 0; q < c->peerWin.size(); q++)
      if (c->peerWin[q] && c->peerWin[q] != c->win) cudaIpcCloseMemHandle(c->peerWin[q]);
   c->peerWin.clear();
   c->p2p = false;
   return CFD_OK;
}

int cfd_p2p_reset(cfd_ctx* c) {
   if (!c) return CFD_ERR_ARG;
   if (!c->win) { set_error("no exchange window (context is not partitioned)"); return CFD_ERR_STATE; }
   CFD_CK(cudaSetDevice(c->device));
   CFD_CK(cudaStreamSynchronize(c->stream));
   CFD_CK(cudaMemset(c->win, 0, kWinHeader));
   CFD_CK(cudaMemset(c->xstate, 0, 8 * sizeof(unsigned long long)));
   CFD_CK(cudaDeviceSynchronize());
   return CFD_OK;
}

int cfd_download(cfd_ctx* c, const char* name, double* host, long long count) {
   if (!c) return CFD_ERR_ARG;
   if (std::string(name) == "sparseAvalue" || std::string(name) == "Adinv") {
      // diagnostic: A in CSR order is not kept on the device; only the Jacobi diagonal is exposed
      if (std::string(name) == "Adinv" && count == c->P.NN) {
         CFD_CK(cudaStreamSynchronize(c->stream));
         CFD_CK(cudaMemcpy(host, c->Adinv, sizeof(double) * count, cudaMemcpyDeviceToHost));
         return CFD_OK;
      }
      set_error("vector %s not downloadable", name);
      return CFD_ERR_ARG;
   }
   auto it = c->vecs.find(name);
   if (it == c->vecs.end() || count != it->second.second) { set_error("unknown vector %s or wrong length %lld", name, count); return CFD_ERR_ARG; }
   CFD_CK(cudaStreamSynchronize(c->stream));
   CFD_CK(cudaMemcpy(host, *it->second.first, sizeof(double) * count, cudaMemcpyDeviceToHost));
   return CFD_OK;
}
void* cfd_device_pointer(cfd_ctx* c, const char* name) {
   if (!c || !name) return nullptr;
   auto it = c->vecs.find(name);
   return it == c->vecs.end() ? nullptr : (void*)*it->second.first;
}

int cfd_upload(cfd_ctx* c, const char* name, const double* host, long long count) {
   if (!c) return CFD_ERR_ARG;
   auto it = c->vecs.find(name);
   if (it == c->vecs.end() || count != it->second.second) { set_error("unknown vector %s or wrong length %lld", name, count); return CFD_ERR_ARG; }
   CFD_CK(cudaStreamSynchronize(c->stream));
   CFD_CK(cudaMemcpy(*it->second.first, host, sizeof(double) * count, cudaMemcpyHostToDevice));
   return CFD_OK;
}
int cfd_get_fields(cfd_ctx* c, double* U, double* P) {
   if (!c) return CFD_ERR_ARG;
   const bool ex = c->P.scheme == CFD_EXPLICIT;
   CFD_CK(cudaStreamSynchronize(c->stream));
   if (U) CFD_CK(cudaMemcpy(U, ex ? c->Un : c->Uk_prev, sizeof(double) * 3 * (size_t)c->P.NN, cudaMemcpyDeviceToHost));
   if (P) CFD_CK(cudaMemcpy(P, ex ? c->Pn : c->Pk_prev, sizeof(double) * (size_t)c->P.NNp, cudaMemcpyDeviceToHost));
   return CFD_OK;
}

// ------------------------------------------------------------------------------------------------
// explicit seam
// ------------------------------------------------------------------------------------------------
// UnpHalf_prev <- Un, Pnp1_prev <- Pn, Acc_prev <- 0 (BCH.cpp:2866-2869) in one launch
__global__ void __launch_bounds__(256) k_begin_step(long long n3, long long np, const double* __restrict__ Un, const double* __restrict__ Pn,
                                                    double* __restrict__ UnpHalf_prev, double* __restrict__ Pnp1_prev, double* __restrict__ Acc_prev) {
   const long long stride = (long long)gridDim.x * blockDim.x;
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n3; i += stride) {
      UnpHalf_prev[i] = Un[i];
      Acc_prev[i] = 0.0;
      if (i < np) Pnp1_prev[i] = Pn[i];
   }
}
int bch_beginTimeStep(cfd_ctx* c) {
   const long long N3 = 3LL * c->P.NN;
   k_begin_step<<<std::min(gridFor(N3, 256), 8 * c->numSMs), 256, 0, c->stream>>>(N3, c->P.NNp, c->Un, c->Pn, c->UnpHalf_prev, c->Pnp1_prev, c->Acc_prev);
   c->launches++;
   CFD_CK(cudaGetLastError());
   return CFD_OK;
}

int bch_calculateMatrixAGPU(cfd_ctx* c) {
   const size_t N3 = 3 * (size_t)c->P.NN;
   if (matrix_free(c)) {
      // the whole operator part of step 1 in ONE element pass: R1 = -C(Un) U - K U - G Pn
      const double nv = c->P.NENv;
      ProfScope ps(c, "ebe", mf_bytes(c, 3 * nv + 3 * nv + 8, 3 * nv));
      CFD_CK(cudaMemsetAsync(c->R1, 0, N3 * sizeof(double), c->stream));
      return launch_momentum_apply<kEbeConv | kEbeVisc | kEbePres>(c, c->Un, c->UnpHalf_prev, c->Pn, -1.0, c->R1);
   }
   ProfScope ps(c, "ebe", ebe_bytes(c, false));
   CFD_CK(cudaMemsetAsync(c->R1, 0, N3 * sizeof(double), c->stream));
   if (c->P.NENv == 27) return launch_ebe_apply<27>(c, c->Un, c->UnpHalf_prev, c->R1);
   return launch_ebe_apply<8>(c, c->Un, c->UnpHalf_prev, c->R1);
}

int bch_step1GPUpart(cfd_ctx* c, int /*iter*/) {
   const int NN = c->P.NN;
   if (matrix_free(c)) {   // K U + G P went into R1 with the convective term (bch_calculateMatrixAGPU): only the pointwise half is left
      ProfScope ps(c, "step1", 8.0 * 3 * NN * 5 + 4.0 * NN);
      k_step1_finish<<<gridFor(NN, 256), 256, 0, c->stream>>>(NN, c->bcIndex, c->Un, c->MdInv, c->R1, c->UnpHalf, c->P.dt);
      c->launches++;
      CFD_CK(cudaGetLastError());
      return CFD_OK;
   }
   ProfScope ps(c, "step1", spmv_bytes(c->K, 1, 3, true) + spmv_bytes(c->G, 3, 1, false) + 8.0 * 3 * NN * 3);
   k_step1<<<gridFor((long long)c->K.nslices * 32, 256), 256, 0, c->stream>>>(c->K, c->G, NN, c->UnpHalf_prev, c->Pn, c->Un, c->MdInv,
                                                                                c->bcIndex, c->R1, c->UnpHalf, c->P.dt);
   c->launches++;
   CFD_CK(cudaGetLastError());
   return halo(c, c->UnpHalf, true);   // gathered by step 2 (G^T) and, as UnpHalf_prev, by the next inner iteration
}

int bch_step2GPU(cfd_ctx* c, int iter) {
   const int NN = c->P.NN, NNp = c->P.NNp;
   const double dt = c->P.dt;
   if (matrix_free(c)) {
      const double nv = c->P.NENv;
      ProfScope ps(c, "gt_rhs", mf_bytes(c, (iter != 1 ? 9 : 3) * nv, 8));
      int rcm = launch_gt_apply(c, 0, c->UnpHalf, c->KtimesAcc_prev, iter != 1, c->MdOrigInv, 1.0 / (dt * dt), c->R2);
      if (rcm) return rcm;
   } else {
      ProfScope ps(c, "gt_rhs", spmv_bytes(c->GT, 3, 3, false));
      k_gt_rhs<<<gridFor((long long)c->GT.nslices * 32, 256), 256, 0, c->stream>>>(c->GT, NN, 0, c->UnpHalf, c->KtimesAcc_prev, iter != 1,
                                                                                    c->MdOrigInv, 1.0 / (dt * dt), c->pin, c->R2, c->scal);
      c->launches++;
      CFD_CK(cudaGetLastError());
   }
   int rc = run_pcg(c, c->R2, c->Pdot, !c->opt.cg_warm_start);
   if (rc) return rc;
   {
      ProfScope ps(c, "update", 40.0 * NNp);
      if (matrix_free(c)) CFD_CK(cudaMemsetAsync(c->scal, 0, 4 * sizeof(double), c->stream));   // (assembled path: zeroed by k_gt_rhs)
      k_pressure_update<<<std::min(gridFor(c->NPo, 256), 4 * c->numSMs), 256, 0, c->stream>>>(c->NPo, c->Pn, c->Pdot, c->Pnp1_prev, c->Pnp1, dt, c->scal);
      c->launches++;
      CFD_CK(cudaGetLastError());
   }
   if ((rc = halo(c, c->Pdot, false))) return rc;    // gathered by step 3 (G Pdot)
   return halo(c, c->Pnp1, false);                     // becomes Pn, gathered by step 1 (G Pn)
}

int bch_step3GPU(cfd_ctx* c, int iter) {
   const int NN = c->P.NN;
   if (matrix_free(c)) {
      const double nv = c->P.NENv;
      {
         ProfScope ps(c, "step3", mf_bytes(c, 8, 3 * nv));
         CFD_CK(cudaMemsetAsync(c->R3, 0, 3 * (size_t)NN * sizeof(double), c->stream));
         int rcm = launch_momentum_apply<kEbePres>(c, nullptr, nullptr, c->Pdot, 1.0, c->R3);
         if (rcm) return rcm;
      }
      ProfScope ps(c, "step3", 8.0 * 3 * NN * 8 + 4.0 * NN);
      k_step3_finish<<<gridFor(NN, 256), 256, 0, c->stream>>>(NN, c->KtimesAcc_prev, iter != 1, c->MdInv, c->bcIndex, c->UnpHalf, c->Unp1_prev, c->R3, c->Acc,
                                                              c->Unp1, c->P.dt, c->scal);
      c->launches++;
      CFD_CK(cudaGetLastError());
      return CFD_OK;
   }
   ProfScope ps(c, "step3", spmv_bytes(c->G, 3, 1, false) + 8.0 * 3 * NN * 7);
   k_step3<<<gridFor((long long)c->G.nslices * 32, 256), 256, 0, c->stream>>>(c->G, NN, c->Pdot, c->KtimesAcc_prev, iter != 1, c->MdInv, c->bcIndex,
                                                                                c->UnpHalf, c->Unp1_prev, c->R3, c->Acc, c->Unp1, c->P.dt, c->scal);
   c->launches++;
   CFD_CK(cudaGetLastError());
   int rc = halo(c, c->Acc, true);        // gathered by K * Acc_prev
   if (rc) return rc;
   return halo(c, c->Unp1, true);          // becomes Un, gathered by the next step's operators
}

int bch_checkConvergenceGPU(cfd_ctx* c, int* converged) {
   int rca = allreduce(c, c->scal, 4, kNcclSum);
   if (rca) return rca;
   CFD_CK(cudaMemcpyAsync(c->h_pinned, c->scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
   int rc = fetch_kinfo(c);
   if (rc) return rc;
   const double* s = c->h_pinned;
   const double n1 = sqrt(s[1]) / sqrt(s[0]);   // NaN when the field is still zero -> not converged, like BCH.cpp:2946-2959
   const double n2 = sqrt(s[3]) / sqrt(s[2]);
   if (converged) *converged = (n1 < c->P.tolerance && n2 < c->P.tolerance) ? 1 : 0;
   return CFD_OK;
}

int bch_rotateIterates(cfd_ctx* c) {
   const size_t N3 = 3 * (size_t)c->P.NN * sizeof(double);
   CFD_CK(cudaMemcpyAsync(c->UnpHalf_prev, c->UnpHalf, N3, cudaMemcpyDeviceToDevice, c->stream));
   CFD_CK(cudaMemcpyAsync(c->Unp1_prev, c->Unp1, N3, cudaMemcpyDeviceToDevice, c->stream));
   CFD_CK(cudaMemcpyAsync(c->Acc_prev, c->Acc, N3, cudaMemcpyDeviceToDevice, c->stream));
   CFD_CK(cudaMemcpyAsync(c->Pnp1_prev, c->Pnp1, c->P.NNp * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
   return CFD_OK;
}

// What cfd_timeLoop runs instead of the copies: the iterates rotate by exchanging buffers (every "current" array is rewritten
// completely by the next inner iteration before anything reads it).  Not for callers that hold raw device pointers.
static void rotate_iterates_by_swap(cfd_ctx* c) {
   std::swap(c->UnpHalf_prev, c->UnpHalf);
   std::swap(c->Unp1_prev, c->Unp1);
   std::swap(c->Acc_prev, c->Acc);
   std::swap(c->Pnp1_prev, c->Pnp1);
}
static void end_time_step_by_swap(cfd_ctx* c) {
   std::swap(c->Un, c->Unp1);   // Un <- Unp1: the old Un buffer becomes the next Unp1, fully rewritten by step 3
   std::swap(c->Pn, c->Pnp1);
   c->timeN++;
}

int bch_calculate_KtimesAcc_prevGPU(cfd_ctx* c) {
   const int NN = c->P.NN;
   if (matrix_free(c)) {
      const double nv = c->P.NENv;
      ProfScope ps(c, "spmv_k", mf_bytes(c, 3 * nv, 3 * nv));
      CFD_CK(cudaMemsetAsync(c->KtimesAcc_prev, 0, 3 * (size_t)NN * sizeof(double), c->stream));
      return launch_momentum_apply<kEbeVisc>(c, nullptr, c->Acc_prev, nullptr, 1.0, c->KtimesAcc_prev);
   }
   ProfScope ps(c, "spmv_k", spmv_bytes(c->K, 1, 3, false));
   k_spmv_multi<3><<<gridFor((long long)c->K.nslices * 32, 256), 256, 0, c->stream>>>(c->K, 0, c->Acc_prev, NN, c->KtimesAcc_prev, NN, 1.0, 0.0);
   c->launches++;
   CFD_CK(cudaGetLastError());
   return halo(c, c->KtimesAcc_prev, true);   // gathered by step 2 (G^T)
}

static int max_acc(cfd_ctx* c, const double* a, const double* b, int* steady, double* maxAcc) {
   const long long N3 = 3LL * c->P.NN;
   {
      ProfScope ps(c, "update", 16.0 * N3);
      CFD_CK(cudaMemsetAsync(c->scal + 4, 0, sizeof(double), c->stream));
      k_max_abs_diff3<<<std::min(gridFor(c->NNo, 256), 8 * c->numSMs), 256, 0, c->stream>>>(c->NNo, c->P.NN, a, b, c->scal + 4);
      c->launches++;
      CFD_CK(cudaGetLastError());
   }
   {
      int rca = allreduce(c, c->scal + 4, 1, kNcclMax);
      if (rca) return rca;
   }
   CFD_CK(cudaMemcpyAsync(c->h_pinned + 8, c->scal + 4, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
   CFD_CK(cudaStreamSynchronize(c->stream));
   double m = c->h_pinned[8];
   if (c->opt.maxacc_rule == 0) m = m * (1.0 / c->P.dt);
   if (maxAcc) *maxAcc = m;
   if (steady) *steady = (m > c->P.convergenceCriteria) ? 0 : 1;
   return CFD_OK;
}

int bch_checkConvergenceInTimeGPU(cfd_ctx* c, int* steady, double* maxAcc) { return max_acc(c, c->Unp1, c->Un, steady, maxAcc); }

int bch_endTimeStep(cfd_ctx* c) {
   CFD_CK(cudaMemcpyAsync(c->Un, c->Unp1, 3 * (size_t)c->P.NN * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
   CFD_CK(cudaMemcpyAsync(c->Pn, c->Pnp1, c->P.NNp * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
   c->timeN++;
   return CFD_OK;
}

// ------------------------------------------------------------------------------------------------
// implicit seam
// ------------------------------------------------------------------------------------------------
int gq_calculateMatrixAGPU(cfd_ctx* c) {
   ProfScope ps(c, "assemble", ebe_bytes(c, true));
   if (c->P.NENv == 27) return launch_ebe_assemble<27>(c, c->Uk_prev, c->K.vals[2]);
   return launch_ebe_assemble<8>(c, c->Uk_prev, c->K.vals[2]);
}

int gq_step1GPU(cfd_ctx* c) {
   const int NN = c->P.NN;
   // A <- M + K (kept pre-summed), then A += C(Uk_prev), Dirichlet rows -> identity
   {
      ProfScope ps(c, "assemble", 16.0 * c->K.padded);
      CFD_CK(cudaMemcpyAsync(c->K.vals[2], c->K.vals[0], sizeof(double) * (size_t)c->K.padded, cudaMemcpyDeviceToDevice, c->stream));
   }
   int rc = gq_calculateMatrixAGPU(c);
   if (rc) return rc;
   {
      ProfScope ps(c, "assemble", 0.0);
      if (c->P.BCnVelNodes > 0) {
         k_bc_rows_identity<<<gridFor(c->P.BCnVelNodes, 128), 128, 0, c->stream>>>(c->K, 2, c->P.BCnVelNodes, c->bcNodes, c->diagPosA);
         c->launches++;
      }
      k_extract_dinv<<<gridFor(c->NNo, 256), 256, 0, c->stream>>>(c->NNo, c->K.vals[2], c->diagPosA, c->Adinv);
      c->launches++;
      CFD_CK(cudaGetLastError());
   }
   {
      ProfScope ps(c, "gq_rhs", spmv_bytes(c->K, 1, 3, false) + spmv_bytes(c->G, 3, 1, false));
      k_gq_rhs<<<gridFor((long long)c->K.nslices * 32, 256), 256, 0, c->stream>>>(c->K, 1, c->G, NN, c->Uk_prev, c->Pk_prev, c->Pk_prevprev,
                                                                                   c->bcIndex, c->BCstr, c->R1i);
      c->launches++;
      CFD_CK(cudaGetLastError());
   }
   if (parted(c) && !c->p2p) {
      int rcd = run_bicg_dist(c);
      if (rcd) return rcd;
      return halo(c, c->Uk, true);   // gathered by step 2 (G^T Uk) and, as Uk_prev, by the next step
   }
   {
      const bool dist = parted(c);
      const size_t N3 = 3 * (size_t)NN;
      BicgArgs a;
      a.A = c->K; a.part = c->bicgPart; a.dynStart = c->bicgDynStart; a.dynCounter = c->bicgDynCounter; a.vidx = 2; a.n = c->NNo; a.ld = NN;
      a.b = c->R1i; a.x = c->Uk; a.dinv = c->Adinv; a.scratch = c->partials;
      a.r = c->bw; a.rs = c->bw + N3; a.p = c->bw + 2 * N3; a.AMp = c->bw + 3 * N3; a.AMs = c->bw + 4 * N3; a.xbest = c->bw + 5 * N3;
      a.ph = dist ? c->bw_ph : c->bw + 6 * N3; a.shv = dist ? c->bw_shv : c->bw + 7 * N3;
      a.info = c->kinfo; a.tol = c->opt.bicg_tol; a.maxit = c->opt.bicg_maxit;
      a.halfstep = c->opt.bicg_no_halfstep ? 0 : 1; a.verify = c->opt.bicg_no_verify ? 0 : 1;
      memset(&a.xc, 0, sizeof a.xc); memset(&a.xv, 0, sizeof a.xv);
      if (dist) {
         fill_xcomm(c, a.xc);
         a.xv.peerV = c->dPeerV; a.xv.peerLd = c->dPeerLd; a.xv.pushPtr = c->pushPtrV; a.xv.pushDst = c->pushDstV;
      }
      void* args[] = {&a};
      ProfScope ps(c, "bicgstab", 0.0);
      c->lastBicgRec = ps.idx;
      CFD_CK(cudaLaunchCooperativeKernel(dist ? (void*)k_bicgstab3<true> : (void*)k_bicgstab3<false>, dim3(c->bicg_blocks), dim3(c->bicg_threads), args, 0, c->stream));
      c->launches++;
#ifdef CFD_BICG_PHASE_TIMERS
      {
         double h[10];
         CFD_CK(cudaStreamSynchronize(c->stream));
         CFD_CK(cudaMemcpy(h, c->kinfo + kBicgInfoTimers, sizeof h, cudaMemcpyDeviceToHost));
         if (h[9] > 0)
            fprintf(stderr, "bicgstab phases (us / lock-step iteration, block 0; %d iterations): spmv1 %.2f red %.2f | s-upd %.2f red %.2f | spmv2 %.2f red %.2f | "
                            "xr-upd %.2f red %.2f | p-upd+sync %.2f\n", (int)h[9], h[0] / h[9] * 1e-3, h[1] / h[9] * 1e-3, h[2] / h[9] * 1e-3, h[3] / h[9] * 1e-3,
                    h[4] / h[9] * 1e-3, h[5] / h[9] * 1e-3, h[6] / h[9] * 1e-3, h[7] / h[9] * 1e-3, h[8] / h[9] * 1e-3);
      }
#endif
   }
   return halo(c, c->Uk, true);   // partitioned runs: gathered by step 2 (G^T Uk) and, as Uk_prev, by the next step
}

int gq_step2GPU(cfd_ctx* c) {
   const int NN = c->P.NN, NNp = c->P.NNp;
   if (matrix_free(c)) {
      ProfScope ps(c, "gt_rhs", mf_bytes(c, 3.0 * c->P.NENv, 8));
      int rcm = launch_gt_apply(c, 1, c->Uk, nullptr, 0, nullptr, -1.0 / c->P.dt, c->R2);
      if (rcm) return rcm;
   } else {
      ProfScope ps(c, "gt_rhs", spmv_bytes(c->GT, 3, 3, false));
      k_gt_rhs<<<gridFor((long long)c->GT.nslices * 32, 256), 256, 0, c->stream>>>(c->GT, NN, 1, c->Uk, nullptr, 0, nullptr, -1.0 / c->P.dt,
                                                                                    c->pin, c->R2);
      c->launches++;
      CFD_CK(cudaGetLastError());
   }
   int rc = run_pcg(c, c->R2, c->Pdiff, true);   // cudaMemset(Pdiff_d, 0) each call, GQ.cu:307
   if (rc) return rc;
   {
      ProfScope ps(c, "update", 24.0 * NNp);
      k_add<<<std::min(gridFor(c->NPo, 256), 4 * c->numSMs), 256, 0, c->stream>>>(c->NPo, c->Pk_prev, c->Pdiff, c->Pk);   // calculate_Pk, GQ.cu:554-565
      c->launches++;
      CFD_CK(cudaGetLastError());
   }
   return halo(c, c->Pk, false);   // gathered by the next step's right-hand sides
}

int gq_checkConvergenceInTimeGPU(cfd_ctx* c, int* steady, double* maxAcc) { return max_acc(c, c->Uk, c->Uk_prev, steady, maxAcc); }

int gq_endTimeStep(cfd_ctx* c) {
   std::swap(c->Uk, c->Uk_prev);   // Uk_prev <- Uk; the old Uk_prev buffer is fully rewritten by the next BiCGStab solve
   double* t = c->Pk_prevprev;     // Pk_prevprev <- Pk_prev <- Pk; the freed buffer becomes the next Pk
   c->Pk_prevprev = c->Pk_prev; c->Pk_prev = c->Pk; c->Pk = t;
   c->timeN++;
   return CFD_OK;
}

int cfd_printMonitorDataGPU(cfd_ctx* c, double out[4]) {
   const bool ex = c->P.scheme == CFD_EXPLICIT;
   const double* U = ex ? c->Un : c->Uk_prev;
   const double* Pp = ex ? c->Pn : c->Pk_prev;
   const int m = c->P.monPoint, NN = c->P.NN;
   for (int d = 0; d < 3; d++) CFD_CK(cudaMemcpyAsync(c->h_pinned + 32 + d, U + (size_t)d * NN + m, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
   CFD_CK(cudaMemcpyAsync(c->h_pinned + 35, Pp + m, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
   CFD_CK(cudaStreamSynchronize(c->stream));
   for (int d = 0; d < 4; d++) out[d] = c->h_pinned[32 + d];
   if (parted(c)) {   // the owner of the monitor node reports, the others contribute zeros
      double v[4];
      for (int d = 0; d < 4; d++) v[d] = c->monOwned ? out[d] : 0.0;
      CFD_CK(cudaMemcpyAsync(c->dacc + 60, v, sizeof v, cudaMemcpyHostToDevice, c->stream));
      int rca = allreduce(c, c->dacc + 60, 4, kNcclSum);
      if (rca) return rca;
      CFD_CK(cudaMemcpyAsync(c->h_pinned + 32, c->dacc + 60, sizeof v, cudaMemcpyDeviceToHost, c->stream));
      CFD_CK(cudaStreamSynchronize(c->stream));
      for (int d = 0; d < 4; d++) out[d] = c->h_pinned[32 + d];
   }
   return CFD_OK;
}

int cfd_last_cg_iterations(cfd_ctx* c) {
   if (fetch_kinfo(c)) return -1;
   return c->last_cg;
}
int cfd_last_bicgstab_iterations(cfd_ctx* c, int its[3]) {
   int rc = fetch_kinfo(c);
   if (rc) return rc;
   for (int d = 0; d < 3; d++) its[d] = c->last_bicg[d];
   return CFD_OK;
}
int cfd_last_krylov_status(cfd_ctx* c, int* cg_converged, double* cg_residual, int bicg_converged[3], double bicg_true_residual[3],
                           int bicg_restarts[3]) {
   if (!c) return CFD_ERR_ARG;
   int rc = fetch_kinfo(c);
   if (rc) return rc;
   if (cg_converged) *cg_converged = c->last_cg_conv;
   if (cg_residual) *cg_residual = c->last_cg_res;
   for (int d = 0; d < 3; d++) {
      if (bicg_converged) bicg_converged[d] = c->last_bicg_conv[d];
      if (bicg_true_residual) bicg_true_residual[d] = c->last_bicg_true[d];
      if (bicg_restarts) bicg_restarts[d] = c->last_bicg_restarts[d];
   }
   return CFD_OK;
}

// ------------------------------------------------------------------------------------------------
// whole time steps
// ------------------------------------------------------------------------------------------------
int cfd_timeLoop(cfd_ctx* c, int nsteps, int stop_on_steady, cfd_step_report* reports, int* done) {
   if (!c) return CFD_ERR_ARG;
   int rc = 0;
   int s = 0;
   for (; s < nsteps; s++) {
      cfd_step_report rep;
      memset(&rep, 0, sizeof rep);
      rep.cg_converged = 1;
      for (int d = 0; d < 3; d++) { rep.bicg_converged[d] = 1; rep.bicg_true_residual[d] = -1.0; }
      if (c->P.scheme == CFD_EXPLICIT) {
         if ((rc = bch_beginTimeStep(c))) return rc;
         int iter;
         for (iter = 1; iter <= c->P.maxIter; iter++) {
            if ((rc = bch_calculateMatrixAGPU(c))) return rc;
            if ((rc = bch_step1GPUpart(c, iter))) return rc;
            if ((rc = bch_step2GPU(c, iter))) return rc;
            if ((rc = bch_step3GPU(c, iter))) return rc;
            int conv = 0;
            if ((rc = bch_checkConvergenceGPU(c, &conv))) return rc;
            rep.cg_iterations += c->last_cg;
            rep.cg_solves++;
            rep.cg_converged &= c->last_cg_conv;
            rep.cg_residual = std::max(rep.cg_residual, c->last_cg_res);
            if (conv) break;
            // BCH.cpp:2929-2932.  Buffer exchange instead of four copies -- except after the LAST allowed iteration, where the loop
            // ends right here and the newest iterate has to stay in Unp1 / Pnp1 (and in the *_prev arrays, as the reference leaves them)
            if (iter < c->P.maxIter) rotate_iterates_by_swap(c);
            else if ((rc = bch_rotateIterates(c))) return rc;
            if ((rc = bch_calculate_KtimesAcc_prevGPU(c))) return rc;
         }
         rep.iter = iter;
         if ((rc = bch_checkConvergenceInTimeGPU(c, &rep.steady, &rep.maxAcc))) return rc;
         end_time_step_by_swap(c);   // BCH.cpp:3070-3071
      } else {
         if ((rc = gq_step1GPU(c))) return rc;
         if ((rc = gq_step2GPU(c))) return rc;
         if ((rc = gq_checkConvergenceInTimeGPU(c, &rep.steady, &rep.maxAcc))) return rc;
         if ((rc = fetch_kinfo(c))) return rc;
         rep.iter = 1;
         rep.cg_iterations = c->last_cg;
         rep.cg_solves = 1;
         rep.cg_converged = c->last_cg_conv;
         rep.cg_residual = c->last_cg_res;
         for (int d = 0; d < 3; d++) {
            rep.bicg_iterations[d] = c->last_bicg[d];
            rep.bicg_converged[d] = c->last_bicg_conv[d];
            rep.bicg_true_residual[d] = c->last_bicg_true[d];
            rep.bicg_restarts[d] = c->last_bicg_restarts[d];
         }
         if ((rc = gq_endTimeStep(c))) return rc;
      }
      rep.timeN = c->timeN;
      if ((rc = cfd_printMonitorDataGPU(c, rep.monitor))) return rc;
      if (reports) reports[s] = rep;
      if (stop_on_steady && rep.steady) { s++; break; }
   }
   if (done) *done = s;
   return CFD_OK;
}

int cfd_timeLoop_timed(cfd_ctx* c, int nsteps, cfd_step_report* reports, double* device_ms) {
   if (!c) return CFD_ERR_ARG;
   CFD_CK(cudaEventRecord(c->ev0, c->stream));
   int done = 0;
   int rc = cfd_timeLoop(c, nsteps, 0, reports, &done);
   if (rc) return rc;
   CFD_CK(cudaEventRecord(c->ev1, c->stream));
   CFD_CK(cudaEventSynchronize(c->ev1));
   float ms = 0;
   CFD_CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
   if (device_ms) *device_ms = ms;
   return CFD_OK;
}

int cfd_timeStep_host(cfd_ctx* c, double* U, double* P, cfd_step_report* report, double* device_ms) {
   if (!c || !U || !P) return CFD_ERR_ARG;
   const bool ex = c->P.scheme == CFD_EXPLICIT;
   const size_t nU = 3 * (size_t)c->P.NN * sizeof(double), nP = (size_t)c->P.NNp * sizeof(double);
   CFD_CK(cudaEventRecord(c->ev0, c->stream));
   CFD_CK(cudaMemcpyAsync(ex ? c->Un : c->Uk_prev, U, nU, cudaMemcpyHostToDevice, c->stream));
   CFD_CK(cudaMemcpyAsync(ex ? c->Pn : c->Pk_prev, P, nP, cudaMemcpyHostToDevice, c->stream));
   int done = 0;
   int rc = cfd_timeLoop(c, 1, 0, report, &done);
   if (rc) return rc;
   CFD_CK(cudaMemcpyAsync(U, ex ? c->Un : c->Uk_prev, nU, cudaMemcpyDeviceToHost, c->stream));
   CFD_CK(cudaMemcpyAsync(P, ex ? c->Pn : c->Pk_prev, nP, cudaMemcpyDeviceToHost, c->stream));
   CFD_CK(cudaEventRecord(c->ev1, c->stream));
   CFD_CK(cudaEventSynchronize(c->ev1));
   float ms = 0;
   CFD_CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
   if (device_ms) *device_ms = ms;
   return CFD_OK;
}

int cfd_profile(cfd_ctx* c, int enable) {
   profile_collect(c);
   c->profile = enable != 0;
   if (enable) c->prof.clear();
   return CFD_OK;
}
int cfd_profile_get(cfd_ctx* c, const char* name, double* ms, long long* launches, double* bytes) {
   profile_collect(c);
   auto it = c->prof.find(name);
   if (it == c->prof.end()) { if (ms) *ms = 0; if (launches) *launches = 0; if (bytes) *bytes = 0; return CFD_ERR_ARG; }
   if (ms) *ms = it->second.ms;
   if (launches) *launches = it->second.launches;
   if (bytes) *bytes = it->second.bytes;
   return CFD_OK;
}
int cfd_profile_get_layout_bytes(cfd_ctx* c, const char* name, double* layout_bytes) {
   profile_collect(c);
   auto it = c->prof.find(name);
   if (layout_bytes) *layout_bytes = it == c->prof.end() ? 0.0 : it->second.lbytes;
   return it == c->prof.end() ? CFD_ERR_ARG : CFD_OK;
}
long long cfd_kernel_launches(cfd_ctx* c) { return c->launches; }

int cfd_matrix_stats(cfd_ctx* c, const char* name, long long* nnz, long long* stored_values, long long* stored_indices, int* lanes_per_row) {
   if (!c || !name) return CFD_ERR_ARG;
   const std::string n(name);
   const Sell* m = n == "K" ? &c->K : n == "G" ? &c->G : n == "GT" ? &c->GT : n == "Z" ? &c->Z : nullptr;
   if (!m) { set_error("unknown matrix %s", name); return CFD_ERR_ARG; }
   if (nnz) *nnz = m->nnz;
   if (stored_values) *stored_values = m->padded;
   if (stored_indices) *stored_indices = m->nindex;
   if (lanes_per_row) *lanes_per_row = m->T;
   return CFD_OK;
}

// ------------------------------------------------------------------------------------------------
// one operator, both ways (assembled / matrix-free)
// ------------------------------------------------------------------------------------------------
int cfd_operator_apply(cfd_ctx* c, const char* op, int path, const double* x, double* y, int repeat, double* ms_per_call, double* bytes) {
   if (!c || !op || !x || !y) return CFD_ERR_ARG;
   if (parted(c)) { set_error("cfd_operator_apply: single-GPU contexts only"); return CFD_ERR_ARG; }
   const std::string o(op);
   const int NN = c->P.NN, NNp = c->P.NNp, NE = c->P.NE;
   const double nv = c->P.NENv, invRho = 1.0 / c->P.density;
   const bool toP = o == "GT" || o == "Z", fromP = o == "G" || o == "Z";
   if (o != "K" && o != "G" && o != "GT" && o != "Z") { set_error("unknown operator %s (K, G, GT, Z)", op); return CFD_ERR_ARG; }
   if (o == "Z" && c->P.scheme != CFD_IMPLICIT) { set_error("operator Z (FE Laplacian) needs an implicit context; the explicit Z = G^T Md^-1 G is a product of G applies"); return CFD_ERR_ARG; }
   const size_t nx = fromP ? (size_t)NNp : 3 * (size_t)NN, ny = toP ? (size_t)NNp : 3 * (size_t)NN;
   double *dx = nullptr, *dy = nullptr;
   CFD_CK(cudaMalloc(&dx, sizeof(double) * nx));
   CFD_CK(cudaMalloc(&dy, sizeof(double) * ny));
   CFD_CK(cudaMemcpy(dx, x, sizeof(double) * nx, cudaMemcpyHostToDevice));
   // K of an implicit context: vals[0] holds M + K, vals[1] M; the explicit layout keeps K alone
   if (o == "K" && path == 0 && c->P.scheme == CFD_IMPLICIT) { cudaFree(dx); cudaFree(dy); set_error("assembled K alone is kept by explicit contexts only"); return CFD_ERR_ARG; }
   Bytes nb(0.0);
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   float best = 1e30f;
   for (int it = 0; it < std::max(1, repeat); it++) {
      cudaEventRecord(e0, c->stream);
      if (path == 0) {
         if (o == "K") { k_spmv_multi<3><<<gridFor((long long)c->K.nslices * 32, 256), 256, 0, c->stream>>>(c->K, 0, dx, NN, dy, NN, 1.0, 0.0); nb = spmv_bytes(c->K, 1, 3, false); }
         else if (o == "G") { k_g_apply<<<gridFor((long long)c->G.nslices * 32, 256), 256, 0, c->stream>>>(c->G, NN, dx, dy); nb = spmv_bytes(c->G, 3, 1, false) + Bytes(16.0 * NN); }
         else if (o == "GT") { k_gt_rhs<<<gridFor((long long)c->GT.nslices * 32, 256), 256, 0, c->stream>>>(c->GT, NN, 1, dx, nullptr, 0, nullptr, 1.0, -1, dy); nb = spmv_bytes(c->GT, 3, 3, false); }
         else { k_spmv_multi<1><<<gridFor((long long)c->Z.nslices * 32, 256), 256, 0, c->stream>>>(c->Z, 0, dx, NNp, dy, NNp, 1.0, 0.0); nb = spmv_bytes(c->Z, 1, 1, false); }
      } else {
         CFD_CK(cudaMemsetAsync(dy, 0, sizeof(double) * ny, c->stream));
         if (o == "K") {
            if (c->P.NENv == 27) k_ebe_momentum_apply<27, kEbeVisc><<<gridFor(NE, 128), 128, 0, c->stream>>>(NE, NN, c->ldE, c->LtoGT, c->coord, nullptr, dx, nullptr, c->P.viscosity, invRho, 1.0, dy);
            else k_ebe_momentum_apply<8, kEbeVisc><<<gridFor(NE, 128), 128, 0, c->stream>>>(NE, NN, c->ldE, c->LtoGT, c->coord, nullptr, dx, nullptr, c->P.viscosity, invRho, 1.0, dy);
            nb = Bytes(mf_bytes(c, 3 * nv, 3 * nv));
         } else if (o == "G") {
            if (c->P.NENv == 27) k_ebe_momentum_apply<27, kEbePres><<<gridFor(NE, 128), 128, 0, c->stream>>>(NE, NN, c->ldE, c->LtoGT, c->coord, nullptr, nullptr, dx, c->P.viscosity, invRho, 1.0, dy);
            else k_ebe_momentum_apply<8, kEbePres><<<gridFor(NE, 128), 128, 0, c->stream>>>(NE, NN, c->ldE, c->LtoGT, c->coord, nullptr, nullptr, dx, c->P.viscosity, invRho, 1.0, dy);
            nb = Bytes(mf_bytes(c, 8, 3 * nv));
         } else if (o == "GT") {
            if (c->P.NENv == 27) k_ebe_gt_apply<27><<<gridFor(NE, 128), 128, 0, c->stream>>>(NE, NN, c->ldE, c->LtoGT, c->coord, 1, dx, nullptr, 0, nullptr, 1.0, invRho, dy);
            else k_ebe_gt_apply<8><<<gridFor(NE, 128), 128, 0, c->stream>>>(NE, NN, c->ldE, c->LtoGT, c->coord, 1, dx, nullptr, 0, nullptr, 1.0, invRho, dy);
            nb = Bytes(mf_bytes(c, 3 * nv, 8));
         } else {
            if (c->P.NENv == 27) k_ebe_laplace_apply<27><<<gridFor(NE, 128), 128, 0, c->stream>>>(NE, c->ldE, c->LtoGT, c->coord, dx, dy);
            else k_ebe_laplace_apply<8><<<gridFor(NE, 128), 128, 0, c->stream>>>(NE, c->ldE, c->LtoGT, c->coord, dx, dy);
            nb = Bytes((4.0 * 8 + 192.0 + 8.0 * 8 + 16.0 * 8) * (double)NE);
         }
      }
      cudaEventRecord(e1, c->stream);
      CFD_CK(cudaEventSynchronize(e1));
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      best = std::min(best, ms);
   }
   CFD_CK(cudaGetLastError());
   CFD_CK(cudaMemcpy(y, dy, sizeof(double) * ny, cudaMemcpyDeviceToHost));
   if (o == "Z" && path == 1 && c->pin >= 0) {
      // the assembled Z carries the pin (diagonal x 1000, applyBC_Step2, GQ.cpp:4677-4693): add the same term to the matrix-free product
      // (999 * diagonal * x[pin]); the diagonal is what the Jacobi preconditioner stores
      double dinv = 1.0;
      CFD_CK(cudaMemcpy(&dinv, c->Zdinv + c->pin, sizeof(double), cudaMemcpyDeviceToHost));
      if (!c->opt.cg_no_precond) y[c->pin] += (999.0 / 1000.0) / dinv * x[c->pin];
   }
   if (ms_per_call) *ms_per_call = best;
   if (bytes) *bytes = path == 0 ? nb.l : nb.a;
   cudaEventDestroy(e0); cudaEventDestroy(e1);
   cudaFree(dx); cudaFree(dy);
   return CFD_OK;
}

// ------------------------------------------------------------------------------------------------
// stand-alone kernels on caller data
// ------------------------------------------------------------------------------------------------
int cfd_csr_spmv(int nrows, int ncols, const int* rowStarts, const int* col, const double* val, int nvec, const double* x, double* y,
                 double alpha, double beta, int repeat, double* ms_per_call) {
   if (nvec != 1 && nvec != 3 && nvec != 6) { set_error("nvec must be 1, 3 or 6"); return CFD_ERR_ARG; }
   int ndev = 0;
   CFD_CK(cudaGetDeviceCount(&ndev));
   if (ndev == 0) { set_error("no CUDA device: this library has no CPU fallback"); return CFD_ERR_CUDA; }
   SellHost h;
   sell_build_host(nrows, ncols, rowStarts, col, choose_T(nrows, rowStarts[nrows]), h);
   Sell A;
   const double* v[1] = {val};
   int rc = sell_upload(h, 1, v, A);
   if (rc) return rc;
   double *dx = nullptr, *dy = nullptr;
   CFD_CK(cudaMalloc(&dx, sizeof(double) * (size_t)nvec * ncols));
   CFD_CK(cudaMalloc(&dy, sizeof(double) * (size_t)nvec * nrows));
   CFD_CK(cudaMemcpy(dx, x, sizeof(double) * (size_t)nvec * ncols, cudaMemcpyHostToDevice));
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   float best = 1e30f;
   for (int it = 0; it < std::max(1, repeat); it++) {
      CFD_CK(cudaMemcpy(dy, y, sizeof(double) * (size_t)nvec * nrows, cudaMemcpyHostToDevice));
      cudaEventRecord(e0);
      const int grid = gridFor((long long)A.nslices * 32, 256);
      if (nvec == 1) k_spmv_multi<1><<<grid, 256>>>(A, 0, dx, ncols, dy, nrows, alpha, beta);
      else if (nvec == 3) k_spmv_multi<3><<<grid, 256>>>(A, 0, dx, ncols, dy, nrows, alpha, beta);
      else k_spmv_multi<6><<<grid, 256>>>(A, 0, dx, ncols, dy, nrows, alpha, beta);
      cudaEventRecord(e1);
      CFD_CK(cudaEventSynchronize(e1));
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      best = std::min(best, ms);
   }
   CFD_CK(cudaGetLastError());
   CFD_CK(cudaMemcpy(y, dy, sizeof(double) * (size_t)nvec * nrows, cudaMemcpyDeviceToHost));
   if (ms_per_call) *ms_per_call = best;
   cudaEventDestroy(e0); cudaEventDestroy(e1);
   cudaFree(dx); cudaFree(dy);
   sell_free(A);
   return CFD_OK;
}

long long cfd_sell_layout_check(int nrows, int ncols, const int* rs, const int* col, const double* val, int T, long long* slices,
                                long long* compressed_slices, double* stored_over_nnz) {
   if (T <= 0) T = choose_T(nrows, rs[nrows]);
   SellHost h;
   sell_build_host(nrows, ncols, rs, col, T, h);
   const Sell& m = h.meta;
   std::vector<double> v((size_t)m.padded, 0.0);
   for (long long k = 0; k < m.nnz; k++) v[(size_t)h.perm[(size_t)k]] = val[k];
   if (slices) *slices = m.nslices;
   if (compressed_slices) *compressed_slices = h.compressedSlices;
   if (stored_over_nnz) *stored_over_nnz = m.nnz ? (double)m.padded / (double)m.nnz : 1.0;
   long long bad = 0;
   std::vector<double> dense;   // one decoded row at a time: accumulate by column, compare with the CSR row
   std::vector<std::pair<int, double> > ent;
   for (int s = 0; s < m.nslices; s++) {
      const long long off = h.sliceOff[s];
      const int w = (int)((h.sliceOff[s + 1] - off) >> 5);
      const long long co = h.colOff[s] & ~1LL;
      const bool cmp = h.colOff[s] & 1;
      for (int rl = 0; rl < m.R; rl++) {
         const int r = s * m.R + rl;
         ent.clear();
         for (int lane = 0; lane < T; lane++) {
            const int t = rl * T + lane;
            for (int j = 0; j < w; j++) {
               const long long rel = (long long)(j >> 1) * 64 + 2 * t + (j & 1);
               int c;
               if (cmp) c = std::min(std::max(r + h.cols[(size_t)co + (size_t)j * T + lane], 0), ncols - 1);
               else c = h.cols[(size_t)(co + rel)];
               const double x = v[(size_t)(off + rel)];
               if (r >= nrows) { if (x != 0.0) bad++; continue; }
               if (c < 0 || c >= ncols) { bad++; continue; }
               if (x != 0.0) ent.push_back(std::make_pair(c, x));
            }
         }
         if (r >= nrows) continue;
         std::sort(ent.begin(), ent.end());
         size_t q = 0;
         for (int k = rs[r]; k < rs[r + 1]; k++) {
            if (val[k] == 0.0) continue;   // explicit zeros of the CSR matrix are indistinguishable from padding
            if (q < ent.size() && ent[q].first == col[k] && ent[q].second == val[k]) q++;
            else bad++;
         }
         if (q != ent.size()) bad += (long long)(ent.size() - q);
      }
   }
   return bad;
}

int cfd_diag_sync(int blocks, int threads, int iters, int mode, double* us_per_op) {
   double *scr = nullptr, *out = nullptr;
   CFD_CK(cudaMalloc(&scr, 8 * (size_t)6 * blocks));
   CFD_CK(cudaMalloc(&out, 8));
   void* args[] = {&iters, &mode, &scr, &out};
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   float best = 1e30f;
   for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      CFD_CK(cudaLaunchCooperativeKernel((void*)k_diag_sync, dim3(blocks), dim3(threads), args, 0, 0));
      cudaEventRecord(e1);
      CFD_CK(cudaEventSynchronize(e1));
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      best = std::min(best, ms);
   }
   if (us_per_op) *us_per_op = 1e3 * best / iters;
   cudaEventDestroy(e0); cudaEventDestroy(e1);
   cudaFree(scr); cudaFree(out);
   return CFD_OK;
}

int cfd_csr_pcg_variant(int n, const int* rowStarts, const int* col, const double* val, const double* b, double* x, double tol, int maxit,
                        int variant, int* its, double* ms) {
   int ndev = 0;
   CFD_CK(cudaGetDeviceCount(&ndev));
   if (ndev == 0) { set_error("no CUDA device: this library has no CPU fallback"); return CFD_ERR_CUDA; }
   if (variant != 1 && variant != 2) { set_error("variant must be 1 (classical) or 2 (Chronopoulos-Gear)"); return CFD_ERR_ARG; }
   SellHost h;
   sell_build_host(n, n, rowStarts, col, choose_T(n, rowStarts[n]), h);
   Sell Z;
   const double* v[1] = {val};
   int rc = sell_upload(h, 1, v, Z);
   if (rc) return rc;
   std::vector<double> dinv(n, 1.0);
   for (int r = 0; r < n; r++)
      for (int k = rowStarts[r]; k < rowStarts[r + 1]; k++)
         if (col[k] == r) { dinv[r] = 1.0 / val[k]; break; }
   double *db, *dx, *dr, *dp, *dq, *dd, *ds, *dpp, *scr, *info;
   CFD_CK(cudaMalloc(&db, 8 * (size_t)n)); CFD_CK(cudaMalloc(&dx, 8 * (size_t)n)); CFD_CK(cudaMalloc(&dr, 16 * (size_t)n));
   CFD_CK(cudaMalloc(&dp, 16 * (size_t)n)); CFD_CK(cudaMalloc(&dq, 8 * (size_t)n)); CFD_CK(cudaMalloc(&dd, 8 * (size_t)n));
   CFD_CK(cudaMalloc(&ds, 8 * (size_t)n)); CFD_CK(cudaMalloc(&dpp, 8 * (size_t)n));
   CFD_CK(cudaMalloc(&info, 8 * 16));
   CFD_CK(cudaMemcpy(db, b, 8 * (size_t)n, cudaMemcpyHostToDevice));
   CFD_CK(cudaMemcpy(dx, x, 8 * (size_t)n, cudaMemcpyHostToDevice));
   CFD_CK(cudaMemcpy(dd, dinv.data(), 8 * (size_t)n, cudaMemcpyHostToDevice));
   int dev = 0, sms = 0, occ = 0;
   cudaGetDevice(&dev);
   cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
   const void* kern = variant == 2 ? pcgcg_kernel(false, 3) : pcg_kernel(3);
   CFD_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 512, 0));
   const int blocks = sms * std::max(1, occ);
   CFD_CK(cudaMalloc(&scr, 8 * (size_t)8 * blocks));
   PcgArgs a;
   a.Z = Z; a.part = nullptr; a.dynStart = Z.nslices; a.dynCounter = nullptr; a.n = n; a.b = db; a.x = dx; a.r = dr; a.p = dp; a.q = dq; a.dinv = dd; a.scratch = scr; a.info = info;
   a.tol = tol; a.maxit = maxit; a.x0zero = 0;
   PcgCgArgs g;
   g.Z = Z; g.part = nullptr; g.dynStart = Z.nslices; g.dynCounter = nullptr; g.n = n; g.b = db; g.x = dx; g.r = dr; g.u = dp; g.w = dq; g.p = dpp; g.s = ds;
   g.dinv = dd; g.scratch = scr; g.info = info; g.tol = tol; g.maxit = maxit; g.x0zero = 0;
   memset(&g.xc, 0, sizeof g.xc);
   void* args[] = {variant == 2 ? (void*)&g : (void*)&a};
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   cudaEventRecord(e0);
   CFD_CK(cudaLaunchCooperativeKernel(kern, dim3(blocks), dim3(512), args, 0, 0));
   cudaEventRecord(e1);
   CFD_CK(cudaEventSynchronize(e1));
   float t = 0;
   cudaEventElapsedTime(&t, e0, e1);
   double hinfo[16];
   CFD_CK(cudaMemcpy(hinfo, info, sizeof hinfo, cudaMemcpyDeviceToHost));
   CFD_CK(cudaMemcpy(x, dx, 8 * (size_t)n, cudaMemcpyDeviceToHost));
   if (its) *its = (int)hinfo[0];
   if (ms) *ms = t;
   if (variant == 1 && getenv("CFD_PCG_PHASES") && hinfo[0] > 0)
      fprintf(stderr, "pcg phases (us/iter, block 0): spmv %.2f  reduce1 %.2f  update %.2f  reduce2 %.2f  pupdate+sync %.2f\n", hinfo[8] / hinfo[0] * 1e-3,
              hinfo[9] / hinfo[0] * 1e-3, hinfo[10] / hinfo[0] * 1e-3, hinfo[11] / hinfo[0] * 1e-3, hinfo[12] / hinfo[0] * 1e-3);
   if (variant == 2 && getenv("CFD_PCG_PHASES") && hinfo[0] > 0)
      fprintf(stderr, "pcg (single-reduction) phases (us/iter, block 0): vector pass %.2f  grid sync %.2f  sweep %.2f  reduction %.2f\n",
              hinfo[8] / hinfo[0] * 1e-3, hinfo[9] / hinfo[0] * 1e-3, hinfo[10] / hinfo[0] * 1e-3, hinfo[11] / hinfo[0] * 1e-3);
   cudaEventDestroy(e0); cudaEventDestroy(e1);
   cudaFree(db); cudaFree(dx); cudaFree(dr); cudaFree(dp); cudaFree(dq); cudaFree(dd); cudaFree(ds); cudaFree(dpp); cudaFree(scr); cudaFree(info);
   sell_free(Z);
   return CFD_OK;
}
int cfd_csr_pcg(int n, const int* rowStarts, const int* col, const double* val, const double* b, double* x, double tol, int maxit,
                int* its, double* ms) {
   return cfd_csr_pcg_variant(n, rowStarts, col, val, b, x, tol, maxit, 1, its, ms);
}

}   // extern "C"
