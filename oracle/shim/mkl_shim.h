// TEST INFRASTRUCTURE -- not product code.
//
// Stand-in for the handful of Intel MKL entry points the reference's CPU build calls
// (fractionalStep/explicit/Cpp/blascoCodinaHuerta.cpp:100-104,403-404,3006-3022,3746-3758,3865-3868,3933-3935,
//  4006-4043; fractionalStep/implicit/Cpp/guermondQuartapelle.cpp:100-104,387-388,3981-4003,4124-4127,4503-4540).
// MKL is not in this image, so the UNMODIFIED reference sources are compiled against these headers
// (oracle/Makefile target `ref`) and linked with mkl_shim.cpp, which restates the documented behaviour of
// each routine ("Intel MKL Reference Manual": Sparse BLAS level 2, RCI ISS conjugate-gradient interface).
// Everything the reference itself computes -- reader, node generation, sparsity, element integrals,
// step0/1/2/3, BC handling, Tecplot writer -- is then the reference's own code running here.
// What stays a restatement (and therefore "unpinned"): the arithmetic INSIDE these routines, i.e. MKL's
// summation order and the exact inner product its CG stopping test uses when a preconditioner is on.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif

typedef int MKL_INT;

// mkl_service.h
void mkl_set_dynamic(int flag);
void mkl_set_num_threads(int n);
void MKL_Free_Buffers(void);
void mkl_free_buffers(void);

// mkl_spblas.h -- y := alpha*op(A)*x + beta*y, CSR given as (val, indx, pntrb, pntre); matdescra[3] = 'C' zero-based, 'F' one-based
void mkl_dcsrmv(const char* transa, const MKL_INT* m, const MKL_INT* k, const double* alpha, const char* matdescra,
                const double* val, const MKL_INT* indx, const MKL_INT* pntrb, const MKL_INT* pntre, const double* x,
                const double* beta, double* y);
// y := A*x, A symmetric, one triangle stored (uplo), 3-array CSR, ONE-based
void mkl_dcsrsymv(const char* uplo, const MKL_INT* m, const double* a, const MKL_INT* ia, const MKL_INT* ja, const double* x, double* y);
// y := A*x, general 3-array CSR, one-based ('N') -- only referenced from the reference's unused MKL_GMRES_solver
void mkl_dcsrgemv(const char* transa, const MKL_INT* m, const double* a, const MKL_INT* ia, const MKL_INT* ja, const double* x, double* y);
// y := alpha*inv(op(A))*x for triangular / diagonal A; the reference only uses matdescra = {'d','l','n'} (diagonal = Jacobi)
void mkl_dcsrsv(const char* transa, const MKL_INT* m, const double* alpha, const char* matdescra, const double* val,
                const MKL_INT* indx, const MKL_INT* pntrb, const MKL_INT* pntre, const double* x, double* y);

// mkl_rci.h -- reverse-communication CG
void dcg_init(const MKL_INT* n, const double* x, const double* b, MKL_INT* rci_request, MKL_INT* ipar, double* dpar, double* tmp);
void dcg_check(const MKL_INT* n, const double* x, const double* b, MKL_INT* rci_request, MKL_INT* ipar, double* dpar, double* tmp);
void dcg(const MKL_INT* n, double* x, const double* b, MKL_INT* rci_request, MKL_INT* ipar, double* dpar, double* tmp);
void dcg_get(const MKL_INT* n, const double* x, const double* b, const MKL_INT* rci_request, const MKL_INT* ipar, const double* dpar,
             const double* tmp, MKL_INT* itercount);
// FGMRES: only referenced from the reference's unused MKL_GMRES_solver (GQ.cpp:4320-4470); aborts if ever called
void dfgmres_init(const MKL_INT* n, const double* x, const double* b, MKL_INT* rci_request, MKL_INT* ipar, double* dpar, double* tmp);
void dfgmres_check(const MKL_INT* n, const double* x, const double* b, MKL_INT* rci_request, MKL_INT* ipar, double* dpar, double* tmp);
void dfgmres(const MKL_INT* n, double* x, double* b, MKL_INT* rci_request, MKL_INT* ipar, double* dpar, double* tmp);
void dfgmres_get(const MKL_INT* n, double* x, double* b, MKL_INT* rci_request, const MKL_INT* ipar, const double* dpar, double* tmp,
                 MKL_INT* itercount);
#ifdef __cplusplus
}
#endif
