// TEST INFRASTRUCTURE -- not product code.
//
// Stand-in for the slice of PARALUTION 1.0.0 (fractionalStep/implicit/Cpp/makeGQ:26) that the reference's implicit
// CPU build uses for its three momentum solves: guermondQuartapelle.cpp:110-112,393-397,517,4183-4231 --
// LocalMatrix / LocalVector wrapping caller arrays, Jacobi preconditioner, BiCGStab with
// Init(abs_tol, rel_tol, div_tol, max_iter).  PARALUTION is not in this image; this header restates the published
// algorithm (van der Vorst's BiCGStab with the preconditioner applied to the search directions; residual
// control: stop when ||r|| <= abs_tol, ||r||/||r0|| <= rel_tol, ||r||/||r0|| >= div_tol or after max_iter iterations,
// tested once per full iteration), so that the UNMODIFIED reference source compiles and runs here.
// What stays a restatement (and therefore "unpinned"): the arithmetic inside the solver, in particular where in
// the iteration PARALUTION tests the residual.
//
// Environment switches (tests/golden/make_ref_golden.py):
//   REF_SHIM_BICG_TOL=t  relative tolerance to use instead of the one passed to Init
//   REF_SHIM_LOG=path    append "bicgstab <iterations> <converged>" per solve
#pragma once
#include <sys/time.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace paralution {

inline void init_paralution() {}
inline void stop_paralution() {}
inline void set_omp_threads_paralution(int) {}
inline void info_paralution() {}
inline double paralution_time() {   // microseconds
   timeval tv;
   gettimeofday(&tv, nullptr);
   return tv.tv_sec * 1e6 + tv.tv_usec;
}

template <class T>
class LocalVector {
public:
   // takes the caller's array (and nulls the caller's pointer); LeaveDataPtr hands it back
   void SetDataPtr(T** ptr, const std::string&, int size) { v_ = *ptr; *ptr = nullptr; n_ = size; }
   void LeaveDataPtr(T** ptr) { *ptr = v_; v_ = nullptr; n_ = 0; }
   void Zeros() { for (int i = 0; i < n_; i++) v_[i] = T(0); }
   T* data() const { return v_; }
   int size() const { return n_; }
private:
   T* v_ = nullptr;
   int n_ = 0;
};

template <class T>
class LocalMatrix {
public:
   void SetDataPtrCSR(int** row_offset, int** col, T** val, const std::string&, int nnz, int nrow, int ncol) {
      rs_ = *row_offset; col_ = *col; val_ = *val;
      *row_offset = nullptr; *col = nullptr; *val = nullptr;
      nnz_ = nnz; nrow_ = nrow; ncol_ = ncol;
   }
   void LeaveDataPtrCSR(int** row_offset, int** col, T** val) {
      *row_offset = rs_; *col = col_; *val = val_;
      rs_ = col_ = nullptr; val_ = nullptr;
   }
   void Apply(const T* x, T* y) const {
#pragma omp parallel for schedule(static)
      for (int i = 0; i < nrow_; i++) {
         T s = T(0);
         for (int k = rs_[i]; k < rs_[i + 1]; k++) s += val_[k] * x[col_[k]];
         y[i] = s;
      }
   }
   void InverseDiagonal(std::vector<T>& d) const {
      d.assign(nrow_, T(1));
      for (int i = 0; i < nrow_; i++)
         for (int k = rs_[i]; k < rs_[i + 1]; k++)
            if (col_[k] == i) d[i] = T(1) / val_[k];
   }
   int nrow() const { return nrow_; }
private:
   int *rs_ = nullptr, *col_ = nullptr;
   T* val_ = nullptr;
   int nnz_ = 0, nrow_ = 0, ncol_ = 0;
};

template <class M, class V, class T>
class Jacobi {
public:
   void Build(const M& A) { A.InverseDiagonal(dinv_); }
   void Solve(const T* r, T* z) const { for (size_t i = 0; i < dinv_.size(); i++) z[i] = dinv_[i] * r[i]; }
private:
   std::vector<T> dinv_;
};

template <class M, class V, class T>
class BiCGStab {
public:
   void Init(double abs_tol, double rel_tol, double div_tol, int max_iter) {
      abs_ = abs_tol; rel_ = rel_tol; div_ = div_tol; maxit_ = max_iter;
      if (const char* t = getenv("REF_SHIM_BICG_TOL")) rel_ = atof(t);
   }
   void SetOperator(M& A) { A_ = &A; }
   void SetPreconditioner(Jacobi<M, V, T>& p) { p_ = &p; }
   void Build() { if (p_) p_->Build(*A_); }
   void Clear() {}

   void Solve(const V& rhs, V* sol) {
      const int n = A_->nrow();
      const T* b = rhs.data();
      T* x = sol->data();
      std::vector<T> r0(n), r(n), p(n), z(n), v(n), w(n), t(n);
      A_->Apply(x, r0.data());
      for (int i = 0; i < n; i++) { r0[i] = b[i] - r0[i]; r[i] = r0[i]; p[i] = r0[i]; }
      const double res0 = norm(r);
      double res = res0;
      T rho = dot(r0, r), alpha = T(1), omega = T(1);
      int it = 0, ok = 0;
      if (res0 <= abs_) ok = 1;
      while (!ok && it < maxit_) {
         precond(p, z);                     // M z = p
         A_->Apply(z.data(), v.data());     // v = A z
         alpha = rho / dot(r0, v);
         for (int i = 0; i < n; i++) r[i] -= alpha * v[i];
         precond(r, w);                     // M w = r
         A_->Apply(w.data(), t.data());     // t = A w
         omega = dot(t, r) / dot(t, t);
         for (int i = 0; i < n; i++) { x[i] += alpha * z[i] + omega * w[i]; r[i] -= omega * t[i]; }
         it++;
         res = norm(r);
         if (res <= abs_ || res / res0 <= rel_) { ok = 1; break; }
         if (!(res / res0 < div_)) break;   // diverged (or NaN)
         const T rho_old = rho;
         rho = dot(r0, r);
         const T beta = (rho / rho_old) * (alpha / omega);
         for (int i = 0; i < n; i++) p[i] = r[i] + beta * (p[i] - omega * v[i]);
      }
      if (const char* path = getenv("REF_SHIM_LOG"))
         if (FILE* f = fopen(path, "a")) { fprintf(f, "bicgstab %d %d\n", it, ok); fclose(f); }
   }

private:
   void precond(const std::vector<T>& in, std::vector<T>& out) const {
      if (p_) p_->Solve(in.data(), out.data());
      else out = in;
   }
   static T dot(const std::vector<T>& a, const std::vector<T>& b) {
      T s = T(0);
      const long long n = (long long)a.size();
#pragma omp parallel for reduction(+ : s) schedule(static)
      for (long long i = 0; i < n; i++) s += a[i] * b[i];
      return s;
   }
   static double norm(const std::vector<T>& a) { return std::sqrt((double)dot(a, a)); }

   M* A_ = nullptr;
   Jacobi<M, V, T>* p_ = nullptr;
   double abs_ = 1e-15, rel_ = 1e-6, div_ = 1e8;
   int maxit_ = 1000000;
};

}   // namespace paralution
