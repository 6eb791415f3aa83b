// cfdb200_shim_bch.cpp -- drop-in replacement for fractionalStep/explicit/Cpp/CUDAcodes.cu.
//
// The reference's host program (blascoCodinaHuerta.cpp, compiled UNMODIFIED with -DUSECUDA) calls eleven functions of
// CUDAcodes.cu (declared at BCH.cpp:365-377) and shares its globals with it (BCH.cpp:145-322; CUDAcodes.cu:47-97 declares
// them `extern`).  This file defines those eleven functions on top of libcfdb200.so; nothing else of the reference changes:
// its main(), reader, preprocessing, step0(), timeLoop() (including its own cudaMemcpy rotations on the `*_d` globals,
// BCH.cpp:2866-2869, 2929-2932, 3070-3071, 3099-3100) and createTecplot() run as they are.
//
//   makeBCH:26   nvcc ... -DUSECUDA blascoCodinaHuerta.cpp CUDAcodes.cu  -lcusparse -lcublas <CUSP headers>
//   becomes           ... -DUSECUDA blascoCodinaHuerta.cpp cfdb200_shim_bch.cpp -lcfdb200 -lcudart
//
// (oracle/Makefile target `refhost` does exactly that here, with the MKL / CSparse stand-ins of oracle/shim for the CPU
//  functions the host file still contains.)
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cfdb200.h"
#include "cfdb200_explicit_z.hpp"

// ---- the reference's globals (defined in blascoCodinaHuerta.cpp) ------------------------------------------------------
extern int NE, NN, NNp, NENv, NENp, NGP, sparseM_NNZ, sparseG_NNZ, BCnVelNodes, nBC, zeroPressureNode, monPoint, maxIter, timeN;
extern int **LtoGnode, **BCvelNodes, *sparseMrowStarts, *sparseMcol, *sparseGrowStarts, *sparseGcol;
extern double **coord, **BCstr, dt, density, viscosity, tolerance, convergenceCriteria, timeT, maxAcc, wallClockTimeCurrentTimeStep;
extern double *sparseKvalue, *sparseG1value, *sparseG2value, *sparseG3value, *MdInv, *MdOrigInv, *Un, *Pn;
// device pointers the host file defines under USECUDA (BCH.cpp:298-306) and copies between in timeLoop()
extern double *Un_d, *Unp1_d, *Unp1_prev_d, *UnpHalf_d, *UnpHalf_prev_d, *Pn_d, *Pnp1_d, *Pnp1_prev_d, *Pdot_d;
extern double *R1_d, *R2_d, *R3_d, *Acc_d, *Acc_prev_d, *KtimesAcc_prev_d;

namespace {
cfd_ctx* ctx = nullptr;
std::vector<int> zrs, zcol;      // Z = sum_c G_c^T diag(MdOrigInv_c) G_c, full CSR, 0-based
std::vector<double> zval;

void die(int rc) {   // the reference prints "ErrorNN" and blocks on `cin >> dummyUserInput` (BCH.cu:184 pattern); a library user wants an exit code
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

void selectCUDAdevice() { die(cfd_selectCUDAdevice(0)); }   // BCH.cu:126-154 picks the LAST device of the box; here: device 0

// Called by step0() in place of the CPU/CSparse calculateZ() (BCH.cpp:3301-3306).  CUDAcodes.cu forms Z with CUSP on the
// host and converts it for the device (BCH.cu:281-411); cfdb200_explicit_z.hpp is a plain CSR product doing the same.
void calculateZ_CUSP() {
   cfdb200_form_explicit_Z(NN, NNp, sparseGrowStarts, sparseGcol, sparseG1value, sparseG2value, sparseG3value, MdOrigInv, zeroPressureNode,
                           zrs, zcol, zval);
}

void initializeAndAllocateGPU() {   // BCH.cu:161-273
   std::vector<double> coord1d((size_t)NN * 3), bcstr1d((size_t)nBC * 3);
   std::vector<int> ltog1d((size_t)NE * NENv), bc1d((size_t)BCnVelNodes * 2);
   for (int i = 0; i < NN; i++) for (int d = 0; d < 3; d++) coord1d[(size_t)i * 3 + d] = coord[i][d];
   for (int e = 0; e < NE; e++) for (int i = 0; i < NENv; i++) ltog1d[(size_t)e * NENv + i] = LtoGnode[e][i];
   for (int i = 0; i < BCnVelNodes; i++) { bc1d[2 * (size_t)i] = BCvelNodes[i][0]; bc1d[2 * (size_t)i + 1] = BCvelNodes[i][1]; }
   for (int i = 0; i < nBC; i++) for (int d = 0; d < 3; d++) bcstr1d[(size_t)i * 3 + d] = BCstr[i][d];
   cfd_problem p = {};
   p.scheme = CFD_EXPLICIT; p.NE = NE; p.NN = NN; p.NNp = NNp; p.NENv = NENv; p.NENp = NENp; p.NGP = NGP;
   p.dt = dt; p.density = density; p.viscosity = viscosity; p.maxIter = maxIter; p.tolerance = tolerance;
   p.convergenceCriteria = convergenceCriteria;
   p.coord = coord1d.data(); p.LtoGnode = ltog1d.data();
   p.sparseM_nnz = sparseM_NNZ / 3; p.sparseMrowStarts = sparseMrowStarts; p.sparseMcol = sparseMcol; p.sparseKvalue = sparseKvalue;
   p.sparseG_nnz = sparseG_NNZ / 3; p.sparseGrowStarts = sparseGrowStarts; p.sparseGcol = sparseGcol;
   p.sparseG1value = sparseG1value; p.sparseG2value = sparseG2value; p.sparseG3value = sparseG3value;
   p.sparseZ_nnz = (int)zcol.size(); p.sparseZrowStarts = zrs.data(); p.sparseZcol = zcol.data(); p.sparseZvalue = zval.data();
   p.MdInv = MdInv; p.MdOrigInv = MdOrigInv;
   p.BCnVelNodes = BCnVelNodes; p.BCvelNodes = bc1d.data(); p.nBC = nBC; p.BCstr = bcstr1d.data();
   p.zeroPressureNode = zeroPressureNode; p.monPoint = monPoint;
   p.U0 = Un; p.P0 = Pn;   // initial or restart state, BCs applied by the host (BCH.cpp:2759-2800)
   cfd_options o;
   cfd_default_options(&o);
   // what CUDAcodes.cu hard-codes: warm-started CG (BCH.cu:418-451) and max|dU| without the 1/dt (BCH.cu:107-117).
   // CFDB200_CPU_SEMANTICS=1 selects the CPU build's instead (x0 = 0, max|dU|/dt), e.g. to compare with solverCPU output.
   const bool cpuSemantics = getenv("CFDB200_CPU_SEMANTICS") && atoi(getenv("CFDB200_CPU_SEMANTICS")) != 0;
   o.cg_warm_start = cpuSemantics ? 0 : 1;
   o.maxacc_rule = cpuSemantics ? 0 : 1;
   o.cuda_stream = (void*)cudaStreamLegacy;   // the host file's own cudaMemcpy / cudaMemset calls run on the legacy stream: keep one order
   if (const char* v = getenv("CFD_CG_TOL")) o.cg_tol = atof(v);   // the original hard-codes 1e-6 (BCH.cu:433)
   die(cfd_initializeAndAllocateGPU(&p, &o, &ctx));
   zrs.clear(); zcol.clear(); zval.clear(); zrs.shrink_to_fit(); zcol.shrink_to_fit(); zval.shrink_to_fit();
   // the `*_d` globals timeLoop() copies between (BCH.cpp:2866-2869, 2929-2932, 3070-3071) and reads back (:3099-3100) ...
   Un_d = dev("Un"); Unp1_d = dev("Unp1"); Unp1_prev_d = dev("Unp1_prev"); UnpHalf_d = dev("UnpHalf"); UnpHalf_prev_d = dev("UnpHalf_prev");
   Pn_d = dev("Pn"); Pnp1_d = dev("Pnp1"); Pnp1_prev_d = dev("Pnp1_prev"); Acc_d = dev("Acc"); Acc_prev_d = dev("Acc_prev");
   // ... and, for completeness, the ones only CUDAcodes.cu itself used
   Pdot_d = (double*)cfd_device_pointer(ctx, "Pdot"); R1_d = (double*)cfd_device_pointer(ctx, "R1"); R2_d = (double*)cfd_device_pointer(ctx, "R2");
   R3_d = (double*)cfd_device_pointer(ctx, "R3"); KtimesAcc_prev_d = (double*)cfd_device_pointer(ctx, "KtimesAcc_prev");
}

void calculateMatrixAGPU()         { die(bch_calculateMatrixAGPU(ctx)); }            // BCH.cu:906-939
void step1GPUpart(int iter)        { die(bch_step1GPUpart(ctx, iter)); }             // BCH.cu:946-973
void step2GPU(int iter)            { die(bch_step2GPU(ctx, iter)); }                 // BCH.cu:980-1022
void step3GPU(int iter)            { die(bch_step3GPU(ctx, iter)); }                 // BCH.cu:1029-1060
void calculate_KtimesAcc_prevGPU() { die(bch_calculate_KtimesAcc_prevGPU(ctx)); }    // BCH.cu:1067-1077
bool checkConvergenceGPU()         { int c = 0; die(bch_checkConvergenceGPU(ctx, &c)); return c != 0; }                        // BCH.cu:1084-1121
bool checkConvergenceInTimeGPU()   { int s = 0; die(bch_checkConvergenceInTimeGPU(ctx, &s, &maxAcc)); return s != 0; }         // BCH.cu:1151-1178
void printMonitorDataGPU(int iter) {                                                                                           // BCH.cu:1128-1144
   // Like the original: bring the whole fields back every step.  The host program relies on it -- its createTecplot() at
   // the end of main() (BCH.cpp:517) writes the HOST arrays, and timeLoop() itself copies them back only every 1000 steps,
   // at steady state or when timeT hits t_final to 1e-10 (BCH.cpp:3097-3112).
   die(cfd_get_fields(ctx, Un, Pn));
   printf("%6d  %6d  %10.5f  %12.5f  %12.5f  %12.5f  %12.5f  %12.5f  %12.5f\n", timeN, iter, timeT, Un[monPoint], Un[monPoint + NN],
          Un[monPoint + 2 * NN], Pn[monPoint], wallClockTimeCurrentTimeStep, maxAcc);
}
