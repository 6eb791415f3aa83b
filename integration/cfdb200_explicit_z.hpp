// cfdb200_explicit_z.hpp -- the pressure matrix of the explicit scheme, Z = sum_c G_c^T diag(MdOrigInv_c) G_c, as a CSR product on
// the host.  The reference forms it in calculateZ_CUSP() (CUDAcodes.cu:281-411, CUSP on the host) for its GPU build and in
// calculateZ() (blascoCodinaHuerta.cpp:3322-3479, CSparse) for its CPU build; the replacement shim needs the same matrix to
// hand to cfd_initializeAndAllocateGPU.  Row i of Z gathers, over the velocity rows r that hold pressure column i,
// G_c(r,i) * MdOrigInv_c(r) * G_c(r,:).  Output: full matrix, 0-based, ascending columns, pin applied.
#pragma once
#include <algorithm>
#include <vector>

inline void cfdb200_form_explicit_Z(int NN, int NNp, const int* growStarts, const int* gcol, const double* G1, const double* G2, const double* G3,
                                    const double* MdOrigInv /* [3NN] */, int zeroPressureNode, std::vector<int>& zrs, std::vector<int>& zcol,
                                    std::vector<double>& zval) {
   const int nnzG = growStarts[NN];
   // transpose pattern of G: for every pressure node the (velocity row, position) pairs that reference it
   std::vector<int> ts(NNp + 1, 0), trow(nnzG), tpos(nnzG);
   for (int k = 0; k < nnzG; k++) ts[gcol[k] + 1]++;
   for (int i = 0; i < NNp; i++) ts[i + 1] += ts[i];
   {
      std::vector<int> fill(ts.begin(), ts.end() - 1);
      for (int r = 0; r < NN; r++)
         for (int k = growStarts[r]; k < growStarts[r + 1]; k++) {
            const int q = fill[gcol[k]]++;
            trow[q] = r; tpos[q] = k;
         }
   }
   const double* G[3] = {G1, G2, G3};
   zrs.assign(NNp + 1, 0);
   std::vector<std::vector<int> > cols(NNp);
   std::vector<std::vector<double> > vals(NNp);
#pragma omp parallel
   {
      std::vector<double> acc(NNp, 0.0);
      std::vector<int> mark(NNp, -1), list;
#pragma omp for schedule(dynamic, 64)
      for (int i = 0; i < NNp; i++) {
         list.clear();
         for (int q = ts[i]; q < ts[i + 1]; q++) {
            const int r = trow[q];
            for (int c = 0; c < 3; c++) {
               const double w = G[c][tpos[q]] * MdOrigInv[(size_t)c * NN + r];
               for (int k = growStarts[r]; k < growStarts[r + 1]; k++) {
                  const int j = gcol[k];
                  if (mark[j] != i) { mark[j] = i; acc[j] = 0.0; list.push_back(j); }
                  acc[j] += w * G[c][k];
               }
            }
         }
         std::sort(list.begin(), list.end());
         cols[i] = list;
         vals[i].resize(list.size());
         for (size_t t = 0; t < list.size(); t++) vals[i][t] = acc[list[t]];
      }
   }
   for (int i = 0; i < NNp; i++) zrs[i + 1] = zrs[i] + (int)cols[i].size();
   zcol.resize(zrs[NNp]); zval.resize(zrs[NNp]);
   for (int i = 0; i < NNp; i++) {
      std::copy(cols[i].begin(), cols[i].end(), zcol.begin() + zrs[i]);
      std::copy(vals[i].begin(), vals[i].end(), zval.begin() + zrs[i]);
   }
   // applyBC_Step2(1): the pinned node's diagonal times 1000, only for a node id > 0 (BCH.cu:357, BCH.cpp:4172)
   if (zeroPressureNode > 0)
      for (int k = zrs[zeroPressureNode]; k < zrs[zeroPressureNode + 1]; k++)
         if (zcol[k] == zeroPressureNode) zval[k] *= 1000.0;
}
