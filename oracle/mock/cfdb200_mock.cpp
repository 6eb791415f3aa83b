// TEST INFRASTRUCTURE -- not product code, never shipped.
//
// A CPU stand-in for libcfdb200.so with the same C ABI (include/cfdb200.h), used ONLY to exercise the reference-side
// binding (integration/cfdb200_shim_*.cpp) and the reference's own -DUSECUDA host programs on a machine without a GPU:
// every seam function is served by the CPU oracle (oracle/liboracle.so), "device pointers" are the oracle's host arrays, and
// mock/fake_cudart.cpp turns the host program's cudaMemcpy / cudaMemset into memcpy / memset.
//   make -C oracle refhost-mock  ->  oracle/_ref/mock/{libcfdb200_cpumock.so, libcudart_fake.so, bch_mock, gq_mock}
// What this checks that nothing else can check here: (1) the arrays the shim marshals into cfd_problem -- i.e. the
// REFERENCE's own preprocessing and step0 results, plus the shim's Z -- equal the oracle's, array by array; (2) the order
// in which the reference's timeLoop() drives the seam, its own rotations on the pointers cfd_device_pointer hands out, and
// the read-back that its createTecplot() relies on, produce the .dat the reference's CPU build writes.
// This is NOT a fallback of the product: the real library has no CPU path (tests/test_abi.py) and cannot load this one -- it has
// another file name, lives under oracle/ (test infrastructure) and is linked only by the two test programs above.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>

#include "../../include/cfdb200.h"

extern "C" {
void* orc_create(const char* inpPath, int scheme, int nthreads);
const char* orc_error(void* h);
long long orc_int(void* h, const char* name);
double orc_double(void* h, const char* name);
int orc_set(void* h, const char* name, double v);
const void* orc_array(void* h, const char* name, long long* count, int* isInt);
int orc_phase(void* h, const char* name, int iter);
}

struct cfd_ctx {
   void* o = nullptr;
   int scheme = 0, NN = 0, NNp = 0;
   double dt = 0, tolerance = 0, convergenceCriteria = 0;
   cfd_options opt;
};

namespace {
char g_err[512] = "";
double* dv(cfd_ctx* c, const char* name, long long* n = nullptr) {
   long long cnt = 0; int isInt = 0;
   const void* p = orc_array(c->o, name, &cnt, &isInt);
   if (n) *n = cnt;
   return (cnt < 0 || isInt) ? nullptr : (double*)p;
}
const int* iv(cfd_ctx* c, const char* name, long long* n) {
   int isInt = 0;
   const void* p = orc_array(c->o, name, n, &isInt);
   return (*n < 0 || !isInt) ? nullptr : (const int*)p;
}
int mismatches = 0;
void cmp_d(cfd_ctx* c, const char* what, const double* got, const char* oname, long long expectCount, double rtol) {
   long long n = 0;
   const double* ref = dv(c, oname, &n);
   if (!got || !ref || n < expectCount) { fprintf(stderr, "mock: %s: missing (%lld < %lld)\n", what, n, expectCount); mismatches++; return; }
   double scale = 0, worst = 0;
   for (long long i = 0; i < expectCount; i++) scale = std::fmax(scale, std::fabs(ref[i]));
   for (long long i = 0; i < expectCount; i++) worst = std::fmax(worst, std::fabs(got[i] - ref[i]));
   const bool ok = worst <= rtol * (scale > 0 ? scale : 1.0);
   fprintf(stderr, "mock: cfd_problem.%-18s %9lld values, max |diff| / max |ref| = %.2e %s\n", what, expectCount, scale > 0 ? worst / scale : worst, ok ? "ok" : "MISMATCH");
   if (!ok) mismatches++;
}
void cmp_i(cfd_ctx* c, const char* what, const int* got, const char* oname, long long expectCount) {
   long long n = 0;
   const int* ref = iv(c, oname, &n);
   bool ok = got && ref && n >= expectCount;
   for (long long i = 0; ok && i < expectCount; i++) ok = got[i] == ref[i];
   fprintf(stderr, "mock: cfd_problem.%-18s %9lld integers %s\n", what, expectCount, ok ? "ok" : "MISMATCH");
   if (!ok) mismatches++;
}
void cmp_s(const char* what, double got, double ref) {
   const bool ok = got == ref;
   if (!ok) { fprintf(stderr, "mock: cfd_problem.%s = %.17g, oracle %.17g MISMATCH\n", what, got, ref); mismatches++; }
}
}   // namespace

extern "C" {

const char* cfd_last_error(void) { return g_err; }
const char* cfd_version(void) { return "cfdb200-mock (CPU oracle behind the C ABI; test infrastructure)"; }
void cfd_default_options(cfd_options* o) {
   memset(o, 0, sizeof *o);
   o->cg_tol = 1e-6; o->cg_maxit = 1000; o->bicg_tol = 1e-6; o->bicg_maxit = 1000; o->device = -1;
}
int cfd_selectCUDAdevice(int) { return CFD_OK; }

int cfd_initializeAndAllocateGPU(const cfd_problem* P, const cfd_options* opt, cfd_ctx** out) {
   std::ifstream pn("ProblemName.txt");
   std::string name;
   if (!(pn >> name)) { snprintf(g_err, sizeof g_err, "mock: no ./ProblemName.txt"); return CFD_ERR_IO; }
   cfd_ctx* c = new cfd_ctx();
   c->o = orc_create((name + ".inp").c_str(), P->scheme, 0);
   if (orc_error(c->o)[0]) { snprintf(g_err, sizeof g_err, "mock: oracle: %s", orc_error(c->o)); return CFD_ERR_IO; }
   c->scheme = P->scheme; c->NN = P->NN; c->NNp = P->NNp; c->dt = P->dt; c->tolerance = P->tolerance;
   c->convergenceCriteria = P->convergenceCriteria;
   if (opt) c->opt = *opt; else cfd_default_options(&c->opt);
   orc_set(c->o, "cg_tol", c->opt.cg_tol); orc_set(c->o, "cg_maxit", c->opt.cg_maxit); orc_set(c->o, "cg_warm", c->opt.cg_warm_start);
   orc_set(c->o, "cg_precond", c->opt.cg_no_precond ? 0 : 1); orc_set(c->o, "bicg_tol", c->opt.bicg_tol);
   orc_set(c->o, "bicg_maxit", c->opt.bicg_maxit); orc_set(c->o, "maxacc_rule", c->opt.maxacc_rule);
   // ---- what the caller marshalled against what the oracle derived from the same input file
   const long long NE = P->NE, NN = P->NN, NNp = P->NNp, nv = P->NENv;
   cmp_s("NE", (double)P->NE, (double)orc_int(c->o, "NE")); cmp_s("NN", (double)P->NN, (double)orc_int(c->o, "NN"));
   cmp_s("NNp", (double)P->NNp, (double)orc_int(c->o, "NNp")); cmp_s("NENv", (double)P->NENv, (double)orc_int(c->o, "NENv"));
   cmp_s("dt", P->dt, orc_double(c->o, "dt")); cmp_s("density", P->density, orc_double(c->o, "density"));
   cmp_s("viscosity", P->viscosity, orc_double(c->o, "viscosity")); cmp_s("convergenceCriteria", P->convergenceCriteria, orc_double(c->o, "convergenceCriteria"));
   cmp_s("zeroPressureNode", P->zeroPressureNode, (double)orc_int(c->o, "zeroPressureNode")); cmp_s("monPoint", P->monPoint, (double)orc_int(c->o, "monPoint"));
   cmp_s("BCnVelNodes", P->BCnVelNodes, (double)orc_int(c->o, "BCnVelNodes")); cmp_s("sparseM_nnz", P->sparseM_nnz, (double)orc_int(c->o, "nnzM"));
   cmp_s("sparseG_nnz", P->sparseG_nnz, (double)orc_int(c->o, "nnzG")); cmp_s("sparseZ_nnz", P->sparseZ_nnz, (double)orc_int(c->o, "nnzZ"));
   if (P->scheme == CFD_EXPLICIT) { cmp_s("maxIter", P->maxIter, (double)orc_int(c->o, "maxIter")); cmp_s("tolerance", P->tolerance, orc_double(c->o, "tolerance")); }
   if (mismatches) { snprintf(g_err, sizeof g_err, "mock: scalar fields of cfd_problem differ from the oracle's"); return CFD_ERR_ARG; }
   cmp_d(c, "coord", P->coord, "coord", 3 * NN, 0.0);
   cmp_i(c, "LtoGnode", P->LtoGnode, "LtoGnode", NE * nv);
   cmp_i(c, "sparseMrowStarts", P->sparseMrowStarts, "sparseMrowStarts", NN + 1);
   cmp_i(c, "sparseMcol", P->sparseMcol, "sparseMcol", P->sparseM_nnz);
   cmp_d(c, "sparseKvalue", P->sparseKvalue, "sparseKvalue", P->sparseM_nnz, 1e-13);
   cmp_i(c, "sparseGrowStarts", P->sparseGrowStarts, "sparseGrowStarts", NN + 1);
   cmp_i(c, "sparseGcol", P->sparseGcol, "sparseGcol", P->sparseG_nnz);
   cmp_d(c, "sparseG1value", P->sparseG1value, "sparseG1value", P->sparseG_nnz, 1e-13);
   cmp_d(c, "sparseG2value", P->sparseG2value, "sparseG2value", P->sparseG_nnz, 1e-13);
   cmp_d(c, "sparseG3value", P->sparseG3value, "sparseG3value", P->sparseG_nnz, 1e-13);
   cmp_i(c, "sparseZrowStarts", P->sparseZrowStarts, "sparseZrowStarts", NNp + 1);
   cmp_i(c, "sparseZcol", P->sparseZcol, "sparseZcol", P->sparseZ_nnz);
   cmp_d(c, "sparseZvalue", P->sparseZvalue, "sparseZvalue", P->sparseZ_nnz, 1e-11);
   cmp_i(c, "BCvelNodes", P->BCvelNodes, "BCvelNodes", 2LL * P->BCnVelNodes);
   cmp_d(c, "BCstr", P->BCstr, "BCstr", 3LL * P->nBC, 0.0);
   if (P->scheme == CFD_EXPLICIT) {
      cmp_d(c, "MdInv", P->MdInv, "MdInv", 3 * NN, 1e-13);
      cmp_d(c, "MdOrigInv", P->MdOrigInv, "MdOrigInv", 3 * NN, 1e-13);
      cmp_d(c, "U0", P->U0, "Un", 3 * NN, 0.0);
      cmp_d(c, "P0", P->P0, "Pn", NNp, 0.0);
   } else {
      cmp_d(c, "sparseMvalue", P->sparseMvalue, "sparseMvalue", P->sparseM_nnz, 1e-13);
      cmp_i(c, "sparseMapM_1d", P->sparseMapM_1d, "sparseMapM_1d", NE * nv * nv);
      cmp_d(c, "U0", P->U0, "Uk_prev", 3 * NN, 0.0);
      cmp_d(c, "P0", P->P0, "Pk_prev", NNp, 0.0);
   }
   if (mismatches) { snprintf(g_err, sizeof g_err, "mock: %d array(s) of cfd_problem differ from the oracle's", mismatches); return CFD_ERR_ARG; }
   *out = c;
   return CFD_OK;
}
void cfd_destroy(cfd_ctx*) {}

void* cfd_device_pointer(cfd_ctx* c, const char* name) { return dv(c, name); }

int cfd_get_fields(cfd_ctx* c, double* U, double* P) {
   const bool ex = c->scheme == CFD_EXPLICIT;
   if (U) memcpy(U, dv(c, ex ? "Un" : "Uk_prev"), sizeof(double) * 3 * (size_t)c->NN);
   if (P) memcpy(P, dv(c, ex ? "Pn" : "Pk_prev"), sizeof(double) * (size_t)c->NNp);
   return CFD_OK;
}

int bch_calculateMatrixAGPU(cfd_ctx*) { return CFD_OK; }   // the oracle's step1 phase forms -C(Un) UnpHalf_prev itself (like BCH.cpp:3692-3707)
int bch_step1GPUpart(cfd_ctx* c, int iter) { return orc_phase(c->o, "bch_step1", iter); }
int bch_step2GPU(cfd_ctx* c, int iter) { return orc_phase(c->o, "bch_step2", iter); }
int bch_step3GPU(cfd_ctx* c, int iter) { return orc_phase(c->o, "bch_step3", iter); }
int bch_calculate_KtimesAcc_prevGPU(cfd_ctx* c) { return orc_phase(c->o, "bch_KtimesAcc", 0); }
int bch_checkConvergenceGPU(cfd_ctx* c, int* converged) {   // BCH.cpp:2936-2973
   const double *u = dv(c, "Unp1"), *up = dv(c, "Unp1_prev"), *p = dv(c, "Pnp1"), *pp = dv(c, "Pnp1_prev");
   double s1 = 0, s2 = 0, s3 = 0, s4 = 0;
   for (long long i = 0; i < 3LL * c->NN; i++) { s1 += u[i] * u[i]; s2 += (u[i] - up[i]) * (u[i] - up[i]); }
   for (long long i = 0; i < c->NNp; i++) { s3 += p[i] * p[i]; s4 += (p[i] - pp[i]) * (p[i] - pp[i]); }
   *converged = (std::sqrt(s2) / std::sqrt(s1) < c->tolerance && std::sqrt(s4) / std::sqrt(s3) < c->tolerance) ? 1 : 0;
   return CFD_OK;
}
static int max_acc(cfd_ctx* c, const double* a, const double* b, int* steady, double* maxAcc) {
   double m = 0;
   for (long long i = 0; i < 3LL * c->NN; i++) m = std::fmax(m, std::fabs(a[i] - b[i]));
   if (c->opt.maxacc_rule == 0) m /= c->dt;
   if (maxAcc) *maxAcc = m;
   if (steady) *steady = m <= c->convergenceCriteria ? 1 : 0;
   return CFD_OK;
}
int bch_checkConvergenceInTimeGPU(cfd_ctx* c, int* steady, double* maxAcc) { return max_acc(c, dv(c, "Unp1"), dv(c, "Un"), steady, maxAcc); }
int gq_step1GPU(cfd_ctx* c) { return orc_phase(c->o, "gq_step1", 0); }
int gq_step2GPU(cfd_ctx* c) { return orc_phase(c->o, "gq_step2", 0); }
int gq_checkConvergenceInTimeGPU(cfd_ctx* c, int* steady, double* maxAcc) { return max_acc(c, dv(c, "Uk"), dv(c, "Uk_prev"), steady, maxAcc); }

}   // extern "C"
