// TEST INFRASTRUCTURE -- not product code.
//
// Stand-in for the five CSparse entry points the reference's explicit CPU build uses to form
// Z = sum_c G_c^T diag(MdOrigInv) G_c  (fractionalStep/explicit/Cpp/blascoCodinaHuerta.cpp:125,234-238,3349-3477).
// CSparse (T. Davis, "Direct Methods for Sparse Linear Systems", SIAM 2006; version unpinned by the reference,
// linked as ../../CSparse/Lib/libcsparse.a) is not in this image.  The data structure below is the book's
// (chapter 2: a matrix is either a triplet list, nz >= 0, or compressed-column, nz == -1); cs_shim.cpp restates the
// book's algorithms for triplet->CSC conversion, transposition and the Gustavson column-by-column product.
// The reference stores `int` arrays into cs::i / cs::p (BCH.cpp:3350-3351), so the index type is int here.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif

#define csi int

typedef struct cs_sparse {
   csi nzmax;   /* number of entries the arrays can hold */
   csi m, n;    /* rows, columns */
   csi* p;      /* column pointers (n+1) or, for a triplet list, column indices (nzmax) */
   csi* i;      /* row indices (nzmax) */
   double* x;   /* values (nzmax) or NULL */
   csi nz;      /* entries in a triplet list; -1 = compressed-column */
} cs;

cs* cs_spalloc(csi m, csi n, csi nzmax, csi values, csi triplet);
cs* cs_spfree(cs* A);
cs* cs_compress(const cs* T);           /* triplet list -> compressed-column (duplicates kept, order of arrival inside a column) */
cs* cs_transpose(const cs* A, csi values);
cs* cs_multiply(const cs* A, const cs* B);   /* C = A*B, columns of C unsorted, nzmax trimmed to the entry count */
csi cs_print(const cs* A, csi brief);

#ifdef __cplusplus
}
#endif
