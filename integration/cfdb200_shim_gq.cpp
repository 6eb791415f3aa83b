// cfdb200_shim_gq.cpp -- drop-in replacement for fractionalStep/implicit/Cpp/CUDAcodes.cu.
//
// The reference's implicit host program (guermondQuartapelle.cpp, compiled UNMODIFIED with -DUSECUDA) calls six functions of
// CUDAcodes.cu (declared at GQ.cpp:350-361; calculateMatrixAGPU is called from inside the .cu's step1GPU) and shares its
// globals with it.  This file defines those functions on top of libcfdb200.so.  The host program keeps its own reader,
// preprocessing, step0() -- including the FE-Laplacian Z it builds itself for the GPU build (setupSparseZGPU / calculateZGPU,
// GQ.cpp:2418-2647, 3677-3772) and applyBC_Step2(1) --, timeLoop() with its cudaMemcpy rotations (GQ.cpp:3366-3368) and
// createTecplot().
//
//   makeGQ:  nvcc ... -DUSECUDA guermondQuartapelle.cpp CUDAcodes.cu -lcusparse -lcublas <CUSP headers> -lparalution
//   becomes       ... -DUSECUDA guermondQuartapelle.cpp cfdb200_shim_gq.cpp -lcfdb200 -lcudart
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cfdb200.h"

// ---- the reference's globals (defined in guermondQuartapelle.cpp) ------------------------------------------------------
extern int NE, NN, NNp, NENv, NENp, NGP, sparseM_NNZ, sparseG_NNZ, sparseZ_NNZ, BCnVelNodes, nBC, zeroPressureNode, monPoint, timeN;
extern int **LtoGnode, **BCvelNodes, *sparseMrowStarts, *sparseMcol, *sparseGrowStarts, *sparseGcol, *sparseZrowStarts, *sparseZcol;
extern int* sparseMapM_1d;
extern double **coord, **BCstr, dt, density, viscosity, convergenceCriteria, timeT, maxAcc, wallClockTimeCurrentTimeStep;
extern double *sparseKvalue, *sparseMvalue, *sparseG1value, *sparseG2value, *sparseG3value, *sparseZvalue, *Uk_prev, *Pk_prev;
// device pointers the host file defines under USECUDA (GQ.cpp:283-290) and copies between in timeLoop()
extern double *Uk_d, *Uk_prev_d, *Pk_d, *Pk_prev_d, *Pk_prevprev_d, *Pdiff_d, *R1_d, *R2_d;

namespace {
cfd_ctx* ctx = nullptr;
void die(int rc) {
   if (rc == CFD_OK) return;
   fprintf(stderr, "cfdb200: error %d: %s\n", rc, cfd_last_error());
   exit(1);
}
double* dev(const char* name) {
   double* p = (double*)cfd_device_pointer(ctx, name);
   if (!p) { fprintf(stderr, "cfdb200: no device vector %s\n", name); exit(1); }
   return p;
}
}   // namespace

void selectCUDAdevice() { die(cfd_selectCUDAdevice(0)); }   // GQ.cu:131-159

void initializeAndAllocateGPU() {   // GQ.cu:166-295
   std::vector<double> coord1d((size_t)NN * 3), bcstr1d((size_t)nBC * 3);
   std::vector<int> ltog1d((size_t)NE * NENv), bc1d((size_t)BCnVelNodes * 2);
   for (int i = 0; i < NN; i++) for (int d = 0; d < 3; d++) coord1d[(size_t)i * 3 + d] = coord[i][d];
   for (int e = 0; e < NE; e++) for (int i = 0; i < NENv; i++) ltog1d[(size_t)e * NENv + i] = LtoGnode[e][i];
   for (int i = 0; i < BCnVelNodes; i++) { bc1d[2 * (size_t)i] = BCvelNodes[i][0]; bc1d[2 * (size_t)i + 1] = BCvelNodes[i][1]; }
   for (int i = 0; i < nBC; i++) for (int d = 0; d < 3; d++) bcstr1d[(size_t)i * 3 + d] = BCstr[i][d];
   cfd_problem p = {};
   p.scheme = CFD_IMPLICIT; p.NE = NE; p.NN = NN; p.NNp = NNp; p.NENv = NENv; p.NENp = NENp; p.NGP = NGP;
   p.dt = dt; p.density = density; p.viscosity = viscosity; p.maxIter = 1; p.convergenceCriteria = convergenceCriteria;
   p.coord = coord1d.data(); p.LtoGnode = ltog1d.data();
   p.sparseM_nnz = sparseM_NNZ / 3; p.sparseMrowStarts = sparseMrowStarts; p.sparseMcol = sparseMcol;
   p.sparseKvalue = sparseKvalue; p.sparseMvalue = sparseMvalue;   // M already carries 1/dt (GQ.cpp:3495)
   p.sparseG_nnz = sparseG_NNZ / 3; p.sparseGrowStarts = sparseGrowStarts; p.sparseGcol = sparseGcol;
   p.sparseG1value = sparseG1value; p.sparseG2value = sparseG2value; p.sparseG3value = sparseG3value;
   // full 0-based FE-Laplacian Z of the GPU build, pin applied by the host's applyBC_Step2(1) (GQ.cpp:4677-4693)
   p.sparseZ_nnz = sparseZ_NNZ; p.sparseZrowStarts = sparseZrowStarts; p.sparseZcol = sparseZcol; p.sparseZvalue = sparseZvalue;
   p.sparseMapM_1d = sparseMapM_1d;
   p.BCnVelNodes = BCnVelNodes; p.BCvelNodes = bc1d.data(); p.nBC = nBC; p.BCstr = bcstr1d.data();
   p.zeroPressureNode = zeroPressureNode; p.monPoint = monPoint;
   p.U0 = Uk_prev; p.P0 = Pk_prev;   // GQ.cu:287-288
   cfd_options o;
   cfd_default_options(&o);          // = CUDAcodes.cu: Jacobi-CG and Jacobi-BiCGStab, 1e-6, 1000 iterations (GQ.cu:344-353, 404-414)
   o.maxacc_rule = 0;                // GQ.cu:975-1002 divides by dt
   o.cuda_stream = (void*)cudaStreamLegacy;   // same stream as the host file's own cudaMemcpy calls
   if (const char* v = getenv("CFD_CG_TOL")) o.cg_tol = atof(v);       // the original hard-codes 1e-6 for both solvers (GQ.cu:344, 404)
   if (const char* v = getenv("CFD_BICG_TOL")) o.bicg_tol = atof(v);
   if (const char* v = getenv("CFD_BICG_MAXIT")) o.bicg_maxit = atoi(v);
   if (const char* v = getenv("CFD_EBE_SCATTER")) o.ebe_scatter = atoi(v);
   if (const char* v = getenv("CFD_DETERMINISTIC")) o.deterministic = atoi(v);
   if (const char* v = getenv("CFD_BICG_NO_HALFSTEP")) o.bicg_no_halfstep = atoi(v);
   die(cfd_initializeAndAllocateGPU(&p, &o, &ctx));
   // the five buffers timeLoop() copies between (GQ.cpp:3366-3368) and reads back (GQ.cpp:3397-3398, 3407-3408) ...
   Uk_d = dev("Uk"); Uk_prev_d = dev("Uk_prev"); Pk_d = dev("Pk"); Pk_prev_d = dev("Pk_prev"); Pk_prevprev_d = dev("Pk_prevprev");
   // ... and, for completeness, the ones only CUDAcodes.cu itself used
   Pdiff_d = (double*)cfd_device_pointer(ctx, "Pdiff"); R1_d = (double*)cfd_device_pointer(ctx, "R1"); R2_d = (double*)cfd_device_pointer(ctx, "R2");
}

void step1GPU() { die(gq_step1GPU(ctx)); }   // GQ.cu:761-890; like the original it calls calculateMatrixAGPU (:726-754) itself
void step2GPU() { die(gq_step2GPU(ctx)); }                                     // GQ.cu:897-944
bool checkConvergenceInTimeGPU() { int s = 0; die(gq_checkConvergenceInTimeGPU(ctx, &s, &maxAcc)); return s != 0; }   // GQ.cu:975-1002
void printMonitorDataGPU() {                                                                                         // GQ.cu:951-967
   // Like the original: bring Uk and Pk back to the host arrays every step (createTecplot() writes those).  timeLoop() has
   // just copied Uk_d into Uk_prev_d and Pk_d into Pk_prev_d (GQ.cpp:3366-3368), which is what cfd_get_fields reads.
   extern double *Uk, *Pk;
   die(cfd_get_fields(ctx, Uk, Pk));
   printf("%6d  %10.5f  %12.5f  %12.5f  %12.5f  %12.5f  %12.5f  %12.5f\n", timeN, timeT, Uk[monPoint], Uk[monPoint + NN], Uk[monPoint + 2 * NN],
          Pk[monPoint], wallClockTimeCurrentTimeStep, maxAcc);
}
