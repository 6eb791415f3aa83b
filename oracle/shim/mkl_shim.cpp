// TEST INFRASTRUCTURE -- not product code.  See mkl_shim.h.
//
// Restatement, from the published interface description (Intel MKL Reference Manual, chapters "Sparse BLAS
// Level 2 and Level 3 Routines" and "Iterative Sparse Solvers based on Reverse Communication Interface"),
// of the MKL routines the reference's CPU build calls.  MKL's version is not pinned by the reference
// (`-mkl=parallel`, fractionalStep/explicit/Cpp/linuxCompile.txt:8).
//
// Environment switches (used by tests/golden/make_ref_golden.py to isolate the reference's OWN arithmetic from the
// stopping rule of the third-party Krylov solver):
//   REF_SHIM_CG_TOL=t       use ||r||_2 <= t*||r0||_2 whatever dpar[0] the caller set
//   REF_SHIM_CG_PRECOND=0|1 force the preconditioner request off / on whatever ipar[10] the caller set
//   REF_SHIM_LOG=path       append one line per finished solve: "cg <iterations> <converged>"
#include "mkl_shim.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace {

int base_of(char c) { return (c == 'F' || c == 'f') ? 1 : 0; }

double dot(const double* a, const double* b, int n) {
   double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
   for (int i = 0; i < n; i++) s += a[i] * b[i];
   return s;
}

void shim_log(const char* what, int its, int converged) {
   const char* p = getenv("REF_SHIM_LOG");
   if (!p) return;
   FILE* f = fopen(p, "a");
   if (!f) return;
   fprintf(f, "%s %d %d\n", what, its, converged);
   fclose(f);
}

[[noreturn]] void unsupported(const char* name) {
   fprintf(stderr, "mkl_shim: %s is not provided (the reference never calls it on the fractional-step path)\n", name);
   abort();
}

struct SymCache {
   const double* a = nullptr;
   const int *ia = nullptr, *ja = nullptr;
   int n = 0, nnz = 0;
   double chk = 0.0;
   char uplo = 0;
   std::vector<int> rs, col;
   std::vector<double> val;
};

// private slots of ipar / dpar (the interface reserves everything past ipar[10] / dpar[7] for the solver)
enum { PHASE = 100 };
enum { RHO = 100, RHO_OLD = 101 };

}   // namespace

extern "C" {

void mkl_set_dynamic(int) {}
void mkl_set_num_threads(int) {}
void MKL_Free_Buffers(void) {}
void mkl_free_buffers(void) {}

void mkl_dcsrmv(const char* transa, const MKL_INT* m, const MKL_INT* k, const double* alpha, const char* matdescra, const double* val,
                const MKL_INT* indx, const MKL_INT* pntrb, const MKL_INT* pntre, const double* x, const double* beta, double* y) {
   if (matdescra[0] != 'g' && matdescra[0] != 'G') unsupported("mkl_dcsrmv for a non-general matrix");
   const int base = base_of(matdescra[3]);
   const int rows = *m, cols = *k;
   const double a = *alpha, b = *beta;
   if (*transa == 'n' || *transa == 'N') {
#pragma omp parallel for schedule(static)
      for (int i = 0; i < rows; i++) {
         double s = 0.0;
         for (int p = pntrb[i] - base; p < pntre[i] - base; p++) s += val[p] * x[indx[p] - base];
         y[i] = (b == 0.0 ? 0.0 : b * y[i]) + a * s;
      }
   } else {
      for (int j = 0; j < cols; j++) y[j] = (b == 0.0 ? 0.0 : b * y[j]);
      for (int i = 0; i < rows; i++) {
         const double xi = a * x[i];
         for (int p = pntrb[i] - base; p < pntre[i] - base; p++) y[indx[p] - base] += val[p] * xi;
      }
   }
}

// The stored triangle is expanded once into a full row-wise copy (kept for as long as the caller keeps passing the same
// arrays; a sampled checksum guards against in-place edits) so that every call is one row-parallel sweep -- the role MKL's
// threaded kernels play in the reference's `-mkl=parallel` build.  The reference calls this once per CG iteration with a
// matrix that never changes (BCH.cpp:4036, GQ.cpp:4533).
void mkl_dcsrsymv(const char* uplo, const MKL_INT* m, const double* a, const MKL_INT* ia, const MKL_INT* ja, const double* x, double* y) {
   static SymCache C;
   const int n = *m;
   const int nnz = ia[n] - 1;
   double chk = 0.0;
   for (int p = 0; p < nnz; p += 61) chk += a[p];
   if (C.a != a || C.ia != ia || C.ja != ja || C.n != n || C.nnz != nnz || C.chk != chk || C.uplo != *uplo) {
      const bool upper = (*uplo == 'u' || *uplo == 'U');
      C.a = a; C.ia = ia; C.ja = ja; C.n = n; C.nnz = nnz; C.chk = chk; C.uplo = *uplo;
      C.rs.assign(n + 1, 0);
      for (int i = 0; i < n; i++)
         for (int p = ia[i] - 1; p < ia[i + 1] - 1; p++) {
            const int j = ja[p] - 1;
            if (j == i) C.rs[i + 1]++;
            else if (upper ? (j > i) : (j < i)) { C.rs[i + 1]++; C.rs[j + 1]++; }
         }
      for (int i = 0; i < n; i++) C.rs[i + 1] += C.rs[i];
      C.col.resize(C.rs[n]); C.val.resize(C.rs[n]);
      std::vector<int> fill(C.rs.begin(), C.rs.end() - 1);
      for (int i = 0; i < n; i++)
         for (int p = ia[i] - 1; p < ia[i + 1] - 1; p++) {
            const int j = ja[p] - 1;
            if (j == i || (upper ? (j > i) : (j < i))) { C.col[fill[i]] = j; C.val[fill[i]++] = a[p]; }
            if (j != i && (upper ? (j > i) : (j < i))) { C.col[fill[j]] = i; C.val[fill[j]++] = a[p]; }
         }
   }
#pragma omp parallel for schedule(static)
   for (int i = 0; i < n; i++) {
      double s = 0.0;
      for (int p = C.rs[i]; p < C.rs[i + 1]; p++) s += C.val[p] * x[C.col[p]];
      y[i] = s;
   }
}

void mkl_dcsrgemv(const char* transa, const MKL_INT* m, const double* a, const MKL_INT* ia, const MKL_INT* ja, const double* x, double* y) {
   if (!(*transa == 'n' || *transa == 'N')) unsupported("mkl_dcsrgemv(transposed)");
   for (int i = 0; i < *m; i++) {
      double s = 0.0;
      for (int p = ia[i] - 1; p < ia[i + 1] - 1; p++) s += a[p] * x[ja[p] - 1];
      y[i] = s;
   }
}

void mkl_dcsrsv(const char*, const MKL_INT* m, const double* alpha, const char* matdescra, const double* val, const MKL_INT* indx,
                const MKL_INT* pntrb, const MKL_INT* pntre, const double* x, double* y) {
   // The reference passes a 3-character descriptor {'d','l','n'} (BCH.cpp:3993-3996), so the indexing base in
   // matdescra[3] is whatever lies behind it on the stack; take the base from the row-pointer array instead.
   if (matdescra[0] != 'd' && matdescra[0] != 'D') unsupported("mkl_dcsrsv for a non-diagonal matrix");
   const int n = *m, base = pntrb[0];
   const bool unit = (matdescra[2] == 'u' || matdescra[2] == 'U');
   for (int i = 0; i < n; i++) {
      double d = 1.0;
      if (!unit)
         for (int p = pntrb[i] - base; p < pntre[i] - base; p++)
            if (indx[p] - base == i) { d = val[p]; break; }
      y[i] = *alpha * x[i] / d;
   }
}

// ---- RCI conjugate gradients ---------------------------------------------------------------------------------
// ipar: [0] n, [1] 6 (message unit), [2] 1, [3] current iteration, [4] max iterations (default min(150,n)), [5] 1, [6] 1,
//       [7] iteration-count stopping test (1), [8] residual stopping test (0), [9] user stopping test (1), [10] preconditioned (0)
// dpar: [0] relative tol (1e-6), [1] absolute tol (0), [2] (r0,r0), [3] dpar[0]*dpar[2]+dpar[1], [4] (r,r) current,
//       [5] (r,r) previous, [6] alpha, [7] beta
// tmp : [0,n) search direction p, [n,2n) A*p, [2n,3n) residual r, [3n,4n) C^-1 r
// requests: 1 = multiply tmp[0,n) by A into tmp[n,2n); 2 = user stopping test; 3 = apply C^-1 to tmp[2n,3n) into tmp[3n,4n);
// result 0 = converged, -1 = iteration limit reached (the reference prints "FAILURE ... ERROR code -1" and keeps x).
// The residual test compares squared EUCLIDEAN norms: dpar[4] <= dpar[3].
void dcg_init(const MKL_INT* n, const double*, const double*, MKL_INT* rci_request, MKL_INT* ipar, double* dpar, double*) {
   for (int i = 0; i < 128; i++) { ipar[i] = 0; dpar[i] = 0.0; }
   ipar[0] = *n; ipar[1] = 6; ipar[2] = 1; ipar[3] = 0; ipar[4] = (*n < 150 ? *n : 150); ipar[5] = 1; ipar[6] = 1;
   ipar[7] = 1; ipar[8] = 0; ipar[9] = 1; ipar[10] = 0;
   dpar[0] = 1e-6; dpar[1] = 0.0;
   *rci_request = 0;
}

void dcg_check(const MKL_INT* n, const double*, const double*, MKL_INT* rci_request, MKL_INT* ipar, double* dpar, double*) {
   *rci_request = (ipar[0] == *n && ipar[4] > 0 && dpar[0] >= 0.0 && dpar[1] >= 0.0) ? 0 : -1100;
   if (const char* t = getenv("REF_SHIM_CG_TOL")) { const double v = atof(t); dpar[0] = v * v; }
   if (const char* p = getenv("REF_SHIM_CG_PRECOND")) ipar[10] = atoi(p);
}

void dcg(const MKL_INT* pn, double* x, const double* b, MKL_INT* rci_request, MKL_INT* ipar, double* dpar, double* tmp) {
   const int n = *pn;
   double *p = tmp, *q = tmp + n, *r = tmp + 2 * (size_t)n, *z = tmp + 3 * (size_t)n;
   for (;;) {
      switch (ipar[PHASE]) {
      case 0:   // r0 = b - A*x0
         for (int i = 0; i < n; i++) p[i] = x[i];
         ipar[PHASE] = 1; *rci_request = 1; return;
      case 1:
         for (int i = 0; i < n; i++) r[i] = b[i] - q[i];
         dpar[2] = dot(r, r, n); dpar[4] = dpar[2]; dpar[5] = dpar[2];
         dpar[3] = dpar[0] * dpar[2] + dpar[1];
         ipar[PHASE] = 2;
         break;
      case 2:   // stopping tests, then the preconditioner request
         if (ipar[8] && dpar[4] <= dpar[3]) { ipar[PHASE] = 0; shim_log("cg", ipar[3], 1); *rci_request = 0; return; }
         if (ipar[7] && ipar[3] >= ipar[4]) { ipar[PHASE] = 0; shim_log("cg", ipar[3], 0); *rci_request = -1; return; }
         ipar[PHASE] = 3;
         if (ipar[9]) { *rci_request = 2; return; }
         break;
      case 3:
         ipar[PHASE] = 4;
         if (ipar[10]) { *rci_request = 3; return; }
         for (int i = 0; i < n; i++) z[i] = r[i];
         break;
      case 4: {   // p = z + beta*p, ask for A*p
         dpar[RHO_OLD] = dpar[RHO];
         dpar[RHO] = dot(r, z, n);
         if (ipar[3] == 0) { for (int i = 0; i < n; i++) p[i] = z[i]; }
         else {
            dpar[7] = dpar[RHO] / dpar[RHO_OLD];
            for (int i = 0; i < n; i++) p[i] = z[i] + dpar[7] * p[i];
         }
         ipar[PHASE] = 5; *rci_request = 1; return;
      }
      case 5:   // x += alpha*p, r -= alpha*A*p
         dpar[6] = dpar[RHO] / dot(p, q, n);
         for (int i = 0; i < n; i++) { x[i] += dpar[6] * p[i]; r[i] -= dpar[6] * q[i]; }
         ipar[3]++;
         dpar[5] = dpar[4]; dpar[4] = dot(r, r, n);
         ipar[PHASE] = 2;
         break;
      default: *rci_request = -10000; return;
      }
   }
}

void dcg_get(const MKL_INT*, const double*, const double*, const MKL_INT*, const MKL_INT* ipar, const double*, const double*, MKL_INT* itercount) {
   *itercount = ipar[3];
}

void dfgmres_init(const MKL_INT*, const double*, const double*, MKL_INT*, MKL_INT*, double*, double*) { unsupported("dfgmres_init"); }
void dfgmres_check(const MKL_INT*, const double*, const double*, MKL_INT*, MKL_INT*, double*, double*) { unsupported("dfgmres_check"); }
void dfgmres(const MKL_INT*, double*, double*, MKL_INT*, MKL_INT*, double*, double*) { unsupported("dfgmres"); }
void dfgmres_get(const MKL_INT*, double*, double*, MKL_INT*, const MKL_INT*, const double*, double*, MKL_INT*) { unsupported("dfgmres_get"); }

}   // extern "C"
