// TEST INFRASTRUCTURE -- not product code.  See cs.h.
#include "cs.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

extern "C" {

cs* cs_spalloc(csi m, csi n, csi nzmax, csi values, csi triplet) {
   cs* A = (cs*)calloc(1, sizeof(cs));
   if (nzmax < 1) nzmax = 1;
   A->m = m; A->n = n; A->nzmax = nzmax;
   A->nz = triplet ? 0 : -1;
   A->p = (csi*)malloc(sizeof(csi) * (size_t)(triplet ? nzmax : n + 1));
   A->i = (csi*)malloc(sizeof(csi) * (size_t)nzmax);
   A->x = values ? (double*)malloc(sizeof(double) * (size_t)nzmax) : nullptr;
   return A;
}

cs* cs_spfree(cs* A) {
   if (!A) return nullptr;
   free(A->p); free(A->i); free(A->x); free(A);
   return nullptr;
}

// counting sort of the triplets by column
cs* cs_compress(const cs* T) {
   if (!T || T->nz < 0) return nullptr;
   const csi n = T->n, nz = T->nz;
   cs* C = cs_spalloc(T->m, n, nz, T->x != nullptr, 0);
   std::vector<csi> next(n + 1, 0);
   for (csi k = 0; k < nz; k++) next[T->p[k] + 1]++;
   for (csi j = 0; j < n; j++) next[j + 1] += next[j];
   for (csi j = 0; j <= n; j++) C->p[j] = next[j];
   for (csi k = 0; k < nz; k++) {
      const csi q = next[T->p[k]]++;
      C->i[q] = T->i[k];
      if (C->x) C->x[q] = T->x[k];
   }
   return C;
}

// counting sort of the entries by row: the result has sorted row indices inside every column
cs* cs_transpose(const cs* A, csi values) {
   if (!A || A->nz >= 0) return nullptr;
   const csi m = A->m, n = A->n, nnz = A->p[n];
   cs* C = cs_spalloc(n, m, nnz, values && A->x, 0);
   std::vector<csi> next(m + 1, 0);
   for (csi k = 0; k < nnz; k++) next[A->i[k] + 1]++;
   for (csi r = 0; r < m; r++) next[r + 1] += next[r];
   for (csi r = 0; r <= m; r++) C->p[r] = next[r];
   for (csi j = 0; j < n; j++)
      for (csi k = A->p[j]; k < A->p[j + 1]; k++) {
         const csi q = next[A->i[k]]++;
         C->i[q] = j;
         if (C->x) C->x[q] = A->x[k];
      }
   C->nzmax = nnz > 0 ? nnz : 1;
   return C;
}

// Column j of C = sum over the entries B(k,j) of column k of A scaled by B(k,j); a dense accumulator and a mark array
// record which rows the column already holds (Gustavson).  Rows appear in order of first touch.
cs* cs_multiply(const cs* A, const cs* B) {
   if (!A || !B || A->nz >= 0 || B->nz >= 0 || A->n != B->m) return nullptr;
   const csi m = A->m, n = B->n;
   const bool values = A->x && B->x;
   std::vector<csi> mark(m, -1), Ci, Cp(n + 1, 0);
   std::vector<double> acc(values ? m : 0, 0.0), Cx;
   for (csi j = 0; j < n; j++) {
      const size_t start = Ci.size();
      for (csi pb = B->p[j]; pb < B->p[j + 1]; pb++) {
         const csi k = B->i[pb];
         const double bkj = values ? B->x[pb] : 0.0;
         for (csi pa = A->p[k]; pa < A->p[k + 1]; pa++) {
            const csi r = A->i[pa];
            if (mark[r] != j) {
               mark[r] = j;
               Ci.push_back(r);
               if (values) acc[r] = bkj * A->x[pa];
            } else if (values) acc[r] += bkj * A->x[pa];
         }
      }
      if (values)
         for (size_t q = start; q < Ci.size(); q++) Cx.push_back(acc[Ci[q]]);
      Cp[j + 1] = (csi)Ci.size();
   }
   const csi nnz = (csi)Ci.size();
   cs* C = cs_spalloc(m, n, nnz, values, 0);
   for (csi j = 0; j <= n; j++) C->p[j] = Cp[j];
   for (csi q = 0; q < nnz; q++) { C->i[q] = Ci[q]; if (values) C->x[q] = Cx[q]; }
   C->nzmax = nnz > 0 ? nnz : 1;
   return C;
}

csi cs_print(const cs* A, csi) {
   if (!A) { printf("(null)\n"); return 0; }
   printf("cs (stand-in) %d-by-%d, nzmax %d, nz %d\n", A->m, A->n, A->nzmax, A->nz);
   return 1;
}

}   // extern "C"
