// step0 assembly, Tecplot/restart I/O and the solver-level C API (product host code).
//
// Replaces step0 / calculateZ / calculateZ_CUSP / calculateZGPU / createTecplot / readRestartFile /
// main of blascoCodinaHuerta.cpp (BCH.cpp) and guermondQuartapelle.cpp (GQ.cpp).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include <omp.h>
#include <sched.h>

#include <thread>

#include "host.hpp"

namespace cfdb {

namespace {

const int kEdgeCorner[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {0, 4}, {1, 5}, {2, 6}, {3, 7}, {4, 5}, {5, 6}, {6, 7}, {7, 4}};
const int kFaceCorner[6][4] = {{0, 1, 2, 3}, {0, 1, 4, 5}, {1, 2, 5, 6}, {2, 3, 6, 7}, {0, 3, 4, 7}, {4, 5, 6, 7}};

// Physical gradients and quadrature factors of one element at its 8 Gauss points
// (calcJacob, BCH.cpp:2495-2711; sub-parametric: Jacobian from the 8 corners through dSp).
struct ElemGeom {
   double fac[8];          // detJ * weight
   double gv[8][27][3];    // grad N_i (velocity basis)
   double gp[8][8][3];     // grad Sp_i (pressure basis)
};

void element_geometry(const Mesh& m, const FeTables& T, int e, ElemGeom& g) {
   const int nv = m.NENv;
   const int* L = &m.LtoGnode[(size_t)e * nv];
   double xc[8][3];
   for (int i = 0; i < 8; i++)
      for (int d = 0; d < 3; d++) xc[i][d] = m.coord[3 * (size_t)L[i] + d];
   for (int k = 0; k < 8; k++) {
      double J[3][3], I[3][3];
      for (int a = 0; a < 3; a++)
         for (int d = 0; d < 3; d++) {
            double s = 0;
            for (int q = 0; q < 8; q++) s += T.dSp[a][q][k] * xc[q][d];
            J[a][d] = s;
         }
      const double c00 = J[1][1] * J[2][2] - J[2][1] * J[1][2];
      const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
      const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
      const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
      const double id = 1.0 / det;
      I[0][0] = c00 * id;
      I[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
      I[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
      I[1][0] = c01 * id;
      I[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
      I[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
      I[2][0] = c02 * id;
      I[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
      I[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
      g.fac[k] = det;   // unit Gauss weights
      for (int i = 0; i < nv; i++)
         for (int a = 0; a < 3; a++)
            g.gv[k][i][a] = I[a][0] * T.dSv[0][i][k] + I[a][1] * T.dSv[1][i][k] + I[a][2] * T.dSv[2][i][k];
      for (int i = 0; i < 8; i++)
         for (int a = 0; a < 3; a++)
            g.gp[k][i][a] = I[a][0] * T.dSp[0][i][k] + I[a][1] * T.dSp[1][i][k] + I[a][2] * T.dSp[2][i][k];
   }
}

}   // namespace

// step0 (BCH.cpp:3127-3315, GQ.cpp:3425-3572): M, K, G1..3; lumped mass and Z = sum_c G_c^T Md^-1 G_c
// (explicit, BCH.cpp:3322-3479 / BCH.cu:281-411) or the FE Laplacian Z (implicit, GQ.cpp:3677-3772).
// Element loop in parallel, summation order of the reference's serial loop (owner-gather, see below).
static double wall_s() {
   timeval tv;
   gettimeofday(&tv, nullptr);
   return tv.tv_sec + 1e-6 * tv.tv_usec;
}
static bool verbose() { return getenv("CFD_VERBOSE") != nullptr; }
#define PHASE(msg) do { if (verbose()) { double t_ = wall_s(); fprintf(stderr, "[cfdb200 host] %-28s %8.2f s\n", msg, t_ - tph); tph = t_; } } while (0)

void assemble_step0(const Mesh& m, int scheme, Assembled& A, bool patternsOnly) {
   const int NE = m.NE, nv = m.NENv, NN = m.NN, NNp = m.NNp;
   double tph = wall_s();
   FeTables T;
   build_fe_tables(nv, T);
   const std::vector<int>* pidx = m.presIndex.empty() ? nullptr : &m.presIndex;
   // patternsOnly (cfd_options.step0_device = 1): the device evaluates the element integrals and finds the entries by bisection (the
   // element -> entry map of the implicit scheme's per-step assembly included), so only the patterns are built here
   build_pattern(m, nv, nv, NN, NN, A.Mp, patternsOnly ? nullptr : &A.mapM);
   const bool samePattern = (nv == 8 && !pidx);   // Q1/Q1 on a whole mesh: M, G and the FE-Laplacian Z share one pattern
   A.sharedG = samePattern;
   if (!samePattern) build_pattern(m, nv, 8, NN, NNp, A.GpOwn, patternsOnly ? nullptr : &A.mapGOwn, nullptr, pidx);
   PHASE("patterns M, G");
   const size_t nnzM = (size_t)A.Mp.nnz(), nnzG = (size_t)A.Gp().nnz();
   if (!patternsOnly) {
      par_zero(A.Mval, nnzM); par_zero(A.Kval, nnzM);
      par_zero(A.G1, nnzG); par_zero(A.G2, nnzG); par_zero(A.G3, nnzG);
   }
   const double invDens = 1.0 / m.density;
   const double mscale = scheme == CFD_IMPLICIT ? 1.0 / m.dt : 1.0;   // implicit M carries 1/dt (GQ.cpp:3495)
   if (scheme == CFD_IMPLICIT) {
      A.sharedZ = samePattern;
      if (!samePattern) build_pattern(m, 8, 8, NNp, NNp, A.ZpOwn, patternsOnly ? nullptr : &A.mapZOwn, pidx, pidx);
      if (!patternsOnly) par_zero(A.Zval, (size_t)A.Zp().nnz());
      PHASE("pattern Z");
   }
   if (patternsOnly) {
      if (scheme == CFD_EXPLICIT) {
         // pattern of Z = sum_c G_c^T Md^-1 G_c: row i = union of the rows of G over the velocity nodes coupled to pressure node i
         const CsrPattern& G = A.Gp();
         std::vector<int> ts((size_t)NNp + 1, 0), trow(nnzG);
         for (size_t k = 0; k < nnzG; k++) ts[G.col[k] + 1]++;
         for (int i = 0; i < NNp; i++) ts[i + 1] += ts[i];
         {
            std::vector<int> fill(ts.begin(), ts.end() - 1);
            for (int r = 0; r < NN; r++)
               for (int k = G.rowStarts[r]; k < G.rowStarts[r + 1]; k++) trow[fill[G.col[k]]++] = r;
         }
         CsrPattern& Z = A.Zp();
         Z.nrows = Z.ncols = NNp;
         Z.rowStarts.assign((size_t)NNp + 1, 0);
         std::vector<std::vector<int> > zc(NNp);
#pragma omp parallel
         {
            std::vector<int> mark((size_t)NNp, -1), list;
#pragma omp for schedule(dynamic, 256)
            for (int i = 0; i < NNp; i++) {
               list.clear();
               for (int t = ts[i]; t < ts[i + 1]; t++) {
                  const int r = trow[t];
                  for (int k = G.rowStarts[r]; k < G.rowStarts[r + 1]; k++) {
                     const int j = G.col[k];
                     if (mark[j] != i) { mark[j] = i; list.push_back(j); }
                  }
               }
               std::sort(list.begin(), list.end());
               zc[i] = list;
            }
         }
         for (int i = 0; i < NNp; i++) Z.rowStarts[i + 1] = Z.rowStarts[i] + (int)zc[i].size();
         Z.col.resize((size_t)Z.rowStarts[NNp]);
#pragma omp parallel for schedule(static)
         for (int i = 0; i < NNp; i++) std::copy(zc[i].begin(), zc[i].end(), Z.col.begin() + Z.rowStarts[i]);
         PHASE("pattern Z (symbolic product)");
      }
      return;
   }
   // Deterministic assembly: every CSR entry receives its element contributions in ASCENDING ELEMENT ORDER -- the order of the
   // reference's serial element loop (BCH.cpp:3127-3315) -- whatever the number of threads, so K, M, G and Z are bitwise the same
   // arrays on every host (the under-integrated implicit Q2 matrices amplify last-bit differences).
   // Elements are processed in blocks: (1) element matrices of the block in parallel into a buffer, (2) owner-gather: in parallel
   // over the nodes the block touches, each node adds the rows it owns from its elements of the block, in element order.
   std::vector<int> adjStart((size_t)NN + 1, 0);
   std::vector<unsigned> adj((size_t)NE * nv);   // (element << 5) | local node, grouped by node, ascending element
   for (size_t q = 0; q < (size_t)NE * nv; q++) adjStart[m.LtoGnode[q] + 1]++;
   for (int n = 0; n < NN; n++) adjStart[n + 1] += adjStart[n];
   {
      std::vector<int> fill(adjStart.begin(), adjStart.end() - 1);
      for (int e = 0; e < NE; e++)
         for (int i = 0; i < nv; i++) adj[fill[m.LtoGnode[(size_t)e * nv + i]]++] = ((unsigned)e << 5) | (unsigned)i;
   }
   const bool implicitZ = scheme == CFD_IMPLICIT;
   const size_t oM = 0, oK = (size_t)nv * nv, oG = 2 * (size_t)nv * nv, oZ = oG + 3 * (size_t)nv * 8, stride = oZ + (implicitZ ? 64 : 0);
   const int B = (int)std::min<long long>(NE, std::max<long long>(4096, (long long)(256u << 20) / (long long)(8 * stride)));
   std::vector<double> buf((size_t)B * stride);
   std::vector<int> cursor(adjStart.begin(), adjStart.end() - 1), mark((size_t)NN, -1), touched;
   for (int b0 = 0, blk = 0; b0 < NE; b0 += B, blk++) {
      const int b1 = std::min(NE, b0 + B);
#pragma omp parallel
      {
         ElemGeom* g = new ElemGeom;
#pragma omp for schedule(static)
         for (int e = b0; e < b1; e++) {
            element_geometry(m, T, e, *g);
            double* eb = &buf[(size_t)(e - b0) * stride];
            for (int i = 0; i < nv; i++)
               for (int j = 0; j < nv; j++) {
                  double me = 0, ke = 0;
                  for (int k = 0; k < 8; k++) {
                     me += T.Sv[k][i] * T.Sv[k][j] * g->fac[k];
                     ke += (g->gv[k][i][0] * g->gv[k][j][0] + g->gv[k][i][1] * g->gv[k][j][1] + g->gv[k][i][2] * g->gv[k][j][2]) * g->fac[k];
                  }
                  eb[oM + i * nv + j] = mscale * me;
                  eb[oK + i * nv + j] = m.viscosity * ke;
               }
            for (int i = 0; i < nv; i++)
               for (int j = 0; j < 8; j++) {
                  double g1 = 0, g2 = 0, g3 = 0;
                  for (int k = 0; k < 8; k++) {
                     const double w = T.Sp[k][j] * g->fac[k];
                     g1 += w * g->gv[k][i][0]; g2 += w * g->gv[k][i][1]; g3 += w * g->gv[k][i][2];
                  }
                  eb[oG + (size_t)(i * 8 + j) * 3 + 0] = -invDens * g1;
                  eb[oG + (size_t)(i * 8 + j) * 3 + 1] = -invDens * g2;
                  eb[oG + (size_t)(i * 8 + j) * 3 + 2] = -invDens * g3;
               }
            if (implicitZ)
               for (int i = 0; i < 8; i++)
                  for (int j = 0; j < 8; j++) {
                     double z = 0;
                     for (int k = 0; k < 8; k++)
                        z += (g->gp[k][i][0] * g->gp[k][j][0] + g->gp[k][i][1] * g->gp[k][j][1] + g->gp[k][i][2] * g->gp[k][j][2]) * g->fac[k];
                     eb[oZ + i * 8 + j] = -z;
                  }
         }
         delete g;
      }
      touched.clear();
      for (int e = b0; e < b1; e++)
         for (int i = 0; i < nv; i++) {
            const int n = m.LtoGnode[(size_t)e * nv + i];
            if (mark[n] != blk) { mark[n] = blk; touched.push_back(n); }
         }
      const long long nt = (long long)touched.size();
#pragma omp parallel for schedule(dynamic, 256)
      for (long long q = 0; q < nt; q++) {
         const int n = touched[(size_t)q];
         int c = cursor[n];
         for (; c < adjStart[n + 1]; c++) {
            const int e = (int)(adj[c] >> 5), i = (int)(adj[c] & 31u);
            if (e >= b1) break;
            const double* eb = &buf[(size_t)(e - b0) * stride];
            const int* mm = &A.mapM[(size_t)e * nv * nv + (size_t)i * nv];
            for (int j = 0; j < nv; j++) { A.Mval[mm[j]] += eb[oM + i * nv + j]; A.Kval[mm[j]] += eb[oK + i * nv + j]; }
            const int* mg = &A.mapG()[(size_t)e * nv * 8 + (size_t)i * 8];
            for (int j = 0; j < 8; j++) {
               A.G1[mg[j]] += eb[oG + (size_t)(i * 8 + j) * 3 + 0];
               A.G2[mg[j]] += eb[oG + (size_t)(i * 8 + j) * 3 + 1];
               A.G3[mg[j]] += eb[oG + (size_t)(i * 8 + j) * 3 + 2];
            }
            if (implicitZ && i < 8) {
               const int* mz = &A.mapZ()[(size_t)e * 64 + (size_t)i * 8];
               for (int j = 0; j < 8; j++) A.Zval[mz[j]] += eb[oZ + i * 8 + j];
            }
         }
         cursor[n] = c;
      }
   }
   PHASE("element assembly");
   if (scheme == CFD_EXPLICIT) {
      // lumped mass, BC rows -> 1 (BCH.cpp:3254-3292)
      std::vector<double> md((size_t)NN, 0.0);
#pragma omp parallel for schedule(static)
      for (int r = 0; r < NN; r++) {
         double s = 0;
         for (int k = A.Mp.rowStarts[r]; k < A.Mp.rowStarts[r + 1]; k++) s += A.Mval[k];
         md[r] = s;
      }
      A.MdOrigInv.resize((size_t)3 * NN); A.MdInv.resize((size_t)3 * NN);
      for (int c = 0; c < 3; c++)
         for (int r = 0; r < NN; r++) { A.MdOrigInv[(size_t)c * NN + r] = 1.0 / md[r]; A.MdInv[(size_t)c * NN + r] = 1.0 / md[r]; }
      for (int i = 0; i < m.BCnVelNodes; i++)
         for (int c = 0; c < 3; c++) A.MdInv[(size_t)c * NN + m.BCvelNodes[2 * i]] = 1.0;
      // Z = sum_c G_c^T diag(MdOrigInv) G_c : row-wise sparse product through the transposed pattern
      std::vector<int> ts((size_t)NNp + 1, 0), trow(nnzG), tpos(nnzG);
      for (size_t k = 0; k < nnzG; k++) ts[A.Gp().col[k] + 1]++;
      for (int i = 0; i < NNp; i++) ts[i + 1] += ts[i];
      {
         std::vector<int> fill(ts.begin(), ts.end() - 1);
         for (int r = 0; r < NN; r++)
            for (int k = A.Gp().rowStarts[r]; k < A.Gp().rowStarts[r + 1]; k++) {
               const int p = fill[A.Gp().col[k]]++;
               trow[p] = r; tpos[p] = k;
            }
      }
      A.Zp().nrows = A.Zp().ncols = NNp;
      A.Zp().rowStarts.assign((size_t)NNp + 1, 0);
      std::vector<std::vector<int> > zc(NNp);
      std::vector<std::vector<double> > zv(NNp);
#pragma omp parallel
      {
         std::vector<double> acc((size_t)NNp, 0.0);
         std::vector<int> mark((size_t)NNp, -1), list;
#pragma omp for schedule(dynamic, 256)
         for (int i = 0; i < NNp; i++) {
            list.clear();
            for (int t = ts[i]; t < ts[i + 1]; t++) {
               const int r = trow[t];
               const double d = md[r];
               const double w1 = A.G1[tpos[t]] / d, w2 = A.G2[tpos[t]] / d, w3 = A.G3[tpos[t]] / d;
               for (int k = A.Gp().rowStarts[r]; k < A.Gp().rowStarts[r + 1]; k++) {
                  const int j = A.Gp().col[k];
                  if (mark[j] != i) { mark[j] = i; acc[j] = 0.0; list.push_back(j); }
                  acc[j] += w1 * A.G1[k] + w2 * A.G2[k] + w3 * A.G3[k];
               }
            }
            std::sort(list.begin(), list.end());
            zc[i] = list;
            zv[i].resize(list.size());
            for (size_t q = 0; q < list.size(); q++) zv[i][q] = acc[list[q]];
         }
      }
      for (int i = 0; i < NNp; i++) A.Zp().rowStarts[i + 1] = A.Zp().rowStarts[i] + (int)zc[i].size();
      A.Zp().col.resize((size_t)A.Zp().rowStarts[NNp]);
      A.Zval.resize((size_t)A.Zp().rowStarts[NNp]);
#pragma omp parallel for schedule(static)
      for (int i = 0; i < NNp; i++) {
         std::copy(zc[i].begin(), zc[i].end(), A.Zp().col.begin() + A.Zp().rowStarts[i]);
         std::copy(zv[i].begin(), zv[i].end(), A.Zval.begin() + A.Zp().rowStarts[i]);
      }
   }
   PHASE("Z");
   // pressure pin: diagonal * 1000, only for node > 0 (applyBC_Step2(1), BCH.cpp:4167-4180)
   // (a partitioned sub-mesh carries the rule already applied: local DOF or -1)
   const int pin = m.presIndex.empty() ? (m.zeroPressureNode > 0 ? m.zeroPressureNode : -1) : m.zeroPressureNode;
   if (pin >= 0)
      for (int k = A.Zp().rowStarts[pin]; k < A.Zp().rowStarts[pin + 1]; k++)
         if (A.Zp().col[k] == pin) { A.Zval[k] *= 1000.0; break; }
}

// createTecplot (BCH.cpp:4249-4482): same bytes (TITLE / VARIABLES / ZONE lines, `scientific`
// precision 11 rows, 8 sub-bricks per Q2 element), written through one large buffer.
// Rows are formatted in parallel (blocks of 32 Ki rows, one buffer per block, written in order): the serial fprintf loop
// needed about a minute for the 3.2 GB file of a 256^3 mesh.  Same bytes as before.
namespace {
template <class F>
bool write_rows_parallel(FILE* f, long long nrows, size_t bytesPerRow, F formatRow) {
   const long long CH = 1 << 15;
   const long long nchunks = (nrows + CH - 1) / CH;
   const int batch = std::max(1, 2 * omp_get_max_threads());
   std::vector<std::vector<char> > buf(batch);
   std::vector<size_t> used(batch, 0);
   for (long long c0 = 0; c0 < nchunks; c0 += batch) {
      const int nb = (int)std::min<long long>(batch, nchunks - c0);
#pragma omp parallel for schedule(static, 1)
      for (int b = 0; b < nb; b++) {
         const long long r0 = (c0 + b) * CH, r1 = std::min(nrows, r0 + CH);
         std::vector<char>& B = buf[b];
         B.resize((size_t)(r1 - r0) * bytesPerRow);
         size_t pos = 0;
         for (long long r = r0; r < r1; r++) pos += formatRow(r, B.data() + pos, bytesPerRow);
         used[b] = pos;
      }
      for (int b = 0; b < nb; b++)
         if (fwrite(buf[b].data(), 1, used[b], f) != used[b]) return false;
   }
   return true;
}
}   // namespace

bool write_tecplot(const Mesh& m, const double* U, const double* P, const std::string& path, const std::string& title) {
   FILE* f = fopen(path.c_str(), "w");
   if (!f) { set_error("cannot write %s", path.c_str()); return false; }
   const int NN = m.NN, nv = m.NENv, NE = m.NE;
   std::vector<char> iobuf(1 << 22);
   setvbuf(f, iobuf.data(), _IOFBF, iobuf.size());
   fprintf(f, "TITLE = %s\n", title.c_str());
   fprintf(f, "VARIABLES = x,  y,  z,  u, v, w, p\n");
   fprintf(f, "ZONE N=%d, E=%d, F=FEPOINT, ET=BRICK\n", NN, nv == 27 ? 8 * NE : NE);
   std::vector<double> pn((size_t)NN, 0.0);
   std::copy(P, P + m.NNp, pn.begin());
   if (nv == 27)
      for (int e = 0; e < NE; e++) {   // element order matters: later elements overwrite shared nodes (BCH.cpp:4299-4436)
         const int* L = &m.LtoGnode[(size_t)e * nv];
         for (int ed = 0; ed < 12; ed++) pn[L[8 + ed]] = 0.5 * (pn[L[kEdgeCorner[ed][0]]] + pn[L[kEdgeCorner[ed][1]]]);
         for (int fc = 0; fc < 6; fc++)
            pn[L[20 + fc]] = 0.25 * (pn[L[kFaceCorner[fc][0]]] + pn[L[kFaceCorner[fc][1]]] + pn[L[kFaceCorner[fc][2]]] + pn[L[kFaceCorner[fc][3]]]);
         pn[L[26]] = 0.125 * (pn[L[0]] + pn[L[1]] + pn[L[2]] + pn[L[3]] + pn[L[4]] + pn[L[5]] + pn[L[6]] + pn[L[7]]);
      }
   // 7 numbers of at most 25 characters ("-1.23456789012e-308" is 19; inf/nan shorter) + separators
   bool ok = write_rows_parallel(f, NN, 7 * 26 + 2, [&](long long i, char* out, size_t cap) {
      return (size_t)snprintf(out, cap, "%.11e %.11e %.11e %.11e %.11e %.11e %.11e\n", m.coord[3 * (size_t)i], m.coord[3 * (size_t)i + 1],
                              m.coord[3 * (size_t)i + 2], U[i], U[(size_t)NN + i], U[2 * (size_t)NN + i], pn[i]);
   });
   static const int sub[8][8] = {{0, 8, 20, 11, 12, 21, 26, 24},  {8, 1, 9, 20, 21, 13, 22, 26},  {11, 20, 10, 3, 24, 26, 23, 15},
                                 {20, 9, 2, 10, 26, 22, 14, 23},  {12, 21, 26, 24, 4, 16, 25, 19}, {21, 13, 22, 26, 16, 5, 17, 25},
                                 {24, 26, 23, 15, 19, 25, 18, 7}, {26, 22, 14, 23, 25, 17, 6, 18}};
   const int nsub = nv == 27 ? 8 : 1;   // Q1/Q1: the reference writer reads local nodes 8..26 and cannot run; one brick per element
   ok = ok && write_rows_parallel(f, (long long)NE * nsub, 8 * 12 + 2, [&](long long r, char* out, size_t cap) {
      const int* L = &m.LtoGnode[(size_t)(r / nsub) * nv];
      if (nsub == 8) {
         const int* q = sub[r % 8];
         return (size_t)snprintf(out, cap, "%d %d %d %d %d %d %d %d\n", L[q[0]] + 1, L[q[1]] + 1, L[q[2]] + 1, L[q[3]] + 1, L[q[4]] + 1, L[q[5]] + 1,
                                 L[q[6]] + 1, L[q[7]] + 1);
      }
      return (size_t)snprintf(out, cap, "%d %d %d %d %d %d %d %d\n", L[0] + 1, L[1] + 1, L[2] + 1, L[3] + 1, L[4] + 1, L[5] + 1, L[6] + 1, L[7] + 1);
   });
   ok = ok && !ferror(f);
   fclose(f);
   if (!ok) set_error("write error on %s", path.c_str());
   return ok;
}

// readRestartFile (BCH.cpp:4214-4242): 3 header lines, then NN rows x y z u v w p (p kept for the
// first NCN rows only).  The file is mapped and its rows are parsed in parallel: byte range split per thread, each
// part aligned to the next line start, rows counted per part, prefix sum = first row of each part.
bool read_restart(const Mesh& m, const std::string& path, double* U, double* P) {
   const int fd = open(path.c_str(), O_RDONLY);
   if (fd < 0) { set_error("cannot open %s", path.c_str()); return false; }
   struct stat st;
   if (fstat(fd, &st) != 0 || st.st_size == 0) { close(fd); set_error("%s: empty or unreadable", path.c_str()); return false; }
   const size_t size = (size_t)st.st_size;
   const char* base = (const char*)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
   close(fd);
   if (base == MAP_FAILED) { set_error("cannot map %s", path.c_str()); return false; }
   const char* endp = base + size;
   const char* p = base;
   for (int i = 0; i < 3; i++) {
      const char* nl = (const char*)memchr(p, '\n', (size_t)(endp - p));
      if (!nl) { munmap((void*)base, size); set_error("%s: truncated header", path.c_str()); return false; }
      p = nl + 1;
   }
   const int nparts = std::max(1, omp_get_max_threads());
   std::vector<const char*> cut(nparts + 1, endp);
   cut[0] = p;
   for (int t = 1; t < nparts; t++) {
      const char* q = p + (size_t)(endp - p) / nparts * t;
      const char* nl = q < endp ? (const char*)memchr(q, '\n', (size_t)(endp - q)) : nullptr;
      cut[t] = nl ? nl + 1 : endp;
   }
   for (int t = 1; t <= nparts; t++) cut[t] = std::max(cut[t], cut[t - 1]);
   std::vector<long long> first(nparts + 1, 0);
#pragma omp parallel for schedule(static, 1)
   for (int t = 0; t < nparts; t++) {
      long long n = 0;
      for (const char* q = cut[t]; q < cut[t + 1];) {
         const char* nl = (const char*)memchr(q, '\n', (size_t)(cut[t + 1] - q));
         n++;
         if (!nl) break;
         q = nl + 1;
      }
      first[t + 1] = n;
   }
   for (int t = 0; t < nparts; t++) first[t + 1] += first[t];
   long long bad = -1;
   const long long NN = m.NN;
#pragma omp parallel for schedule(static, 1)
   for (int t = 0; t < nparts; t++) {
      long long row = first[t];
      for (const char* q = cut[t]; q < cut[t + 1] && row < NN; row++) {
         const char* nl = (const char*)memchr(q, '\n', (size_t)(cut[t + 1] - q));
         const char* le = nl ? nl : cut[t + 1];
         char line[512];
         const size_t len = std::min<size_t>((size_t)(le - q), sizeof line - 1);
         memcpy(line, q, len);
         line[len] = 0;
         double v[7];
         char* c = line;
         int got = 0;
         for (; got < 7; got++) {
            char* e2;
            v[got] = strtod(c, &e2);
            if (e2 == c) break;
            c = e2;
         }
         if (got != 7) {
#pragma omp critical
            if (bad < 0 || row < bad) bad = row;
         } else {
            U[row] = v[3]; U[(size_t)NN + row] = v[4]; U[2 * (size_t)NN + row] = v[5];
            if (row < m.NCN) P[row] = v[6];
         }
         q = nl ? nl + 1 : cut[t + 1];
      }
   }
   munmap((void*)base, size);
   if (bad < 0 && first[nparts] < NN) bad = first[nparts];
   if (bad >= 0) { set_error("%s: malformed row %lld", path.c_str(), bad); return false; }
   return true;
}

}   // namespace cfdb

// =================================================================================================
// Solver-level C API
// =================================================================================================
using namespace cfdb;

struct cfd_solver {
   Mesh mesh;          // the mesh this process computes on (the whole mesh, or its partition's sub-mesh)
   Partition part;
   bool partitioned = false;
   int globalNN = 0, globalNNp = 0, globalNE = 0;
   std::vector<int> peerRank, sendVptr, sendVidx, recvVoff, recvVcnt, sendPptr, sendPidx, recvPoff, recvPcnt, remoteVoff, remotePoff;
   Assembled A;
   cfd_problem prob;
   cfd_options opt;
   cfd_ctx* ctx = nullptr;
   int scheme = 0;
   std::vector<double> U0, P0;
   double timeT = 0;
};

namespace {

double now_s() {
   timeval tv;
   gettimeofday(&tv, nullptr);
   return tv.tv_sec + 1e-6 * tv.tv_usec;
}

int finish_create(cfd_solver* s, int scheme, const cfd_options* opt, const std::string& restartPath, int init_gpu, int world = 1,
                  int rank = 0) {
   s->scheme = scheme;
   if (scheme != CFD_EXPLICIT && scheme != CFD_IMPLICIT) { set_error("scheme must be CFD_EXPLICIT (0) or CFD_IMPLICIT (1), got %d", scheme); return CFD_ERR_ARG; }
   // the reference divides by dt and the density without looking (BCH.cpp:2822, 3240); a zero would only show up as NaN fields
   if (!(s->mesh.dt > 0.0) || !(s->mesh.density > 0.0) || !(s->mesh.viscosity >= 0.0)) {
      set_error("dt and density must be positive, viscosity non-negative (dt %g, density %g, viscosity %g)", s->mesh.dt, s->mesh.density, s->mesh.viscosity);
      return CFD_ERR_ARG;
   }
   if (world > s->mesh.NE) { set_error("%d ranks for %d elements: an element block per rank needs world <= NE", world, s->mesh.NE); return CFD_ERR_ARG; }
   if (opt) s->opt = *opt; else cfd_default_options(&s->opt);
   // the thread count chosen below is for this set-up only: the caller's OpenMP setting is put back on every way out
   struct ThreadsBack { int n; ~ThreadsBack() { omp_set_num_threads(n); } } threadsBack{omp_get_max_threads()};
   {
      // one process per GPU: launchers such as torchrun pin OMP_NUM_THREADS=1; give each rank its share of the cores
      int want = 0;
      if (const char* e = getenv("CFD_HOST_THREADS")) want = atoi(e);
      if (want <= 0 && world > 1) {
         cpu_set_t set;
         int ncpu = (int)std::thread::hardware_concurrency();
         if (sched_getaffinity(0, sizeof set, &set) == 0) ncpu = CPU_COUNT(&set);   // what `nproc` reports (cpuset-aware)
         want = std::max(1, ncpu / world);
      }
      if (want > 0) omp_set_num_threads(want);
   }
   generate_midnodes(s->mesh);
   find_bc_nodes(s->mesh);
   find_monitor(s->mesh);
   s->globalNN = s->mesh.NN; s->globalNNp = s->mesh.NNp; s->globalNE = s->mesh.NE;
   if (world > 1) {
      if (world > 64 || rank < 0 || rank >= world) { set_error("bad partition %d/%d", rank, world); return CFD_ERR_ARG; }
      // restart (readRestartFile, BCH.cpp:4214-4242) on a partition: every rank reads the one restart file of the WHOLE mesh before
      // the mesh is cut, and keeps the rows of its own nodes (owned and ghost) below
      std::vector<double> Ug, Pg;
      const bool restart = s->mesh.isRestart == 1 && !restartPath.empty();
      if (restart) {
         Ug.assign((size_t)3 * s->mesh.NN, 0.0); Pg.assign((size_t)s->mesh.NNp, 0.0);
         if (!read_restart(s->mesh, restartPath, Ug.data(), Pg.data())) return CFD_ERR_IO;
      }
      const size_t NNg = (size_t)s->mesh.NN;
      Mesh local;
      // explicit: rows of Z = G^T Md^-1 G reach two element layers; implicit: one
      partition_mesh(s->mesh, world, rank, scheme == CFD_EXPLICIT ? 2 : 1, local, s->part);
      s->mesh = std::move(local);
      s->partitioned = true;
      if (restart) {
         const size_t NNl = s->part.nodeGlobal.size(), NPl = s->part.presGlobal.size();
         s->U0.assign(3 * NNl, 0.0); s->P0.assign(NPl, 0.0);
         for (size_t i = 0; i < NNl; i++)
            for (int d = 0; d < 3; d++) s->U0[(size_t)d * NNl + i] = Ug[(size_t)d * NNg + (size_t)s->part.nodeGlobal[i]];
         for (size_t i = 0; i < NPl; i++) s->P0[i] = Pg[(size_t)s->part.presGlobal[i]];
         s->mesh.isRestart = 0;   // (already applied: the block further down reads the file of a whole mesh)
      }
   }
   Mesh& m = s->mesh;
   const bool dev0 = s->opt.step0_device == 1;
   if (dev0 && world > 1) { set_error("step0_device = 1 is single-GPU"); return CFD_ERR_ARG; }
   assemble_step0(m, scheme, s->A, dev0);
   cfd_problem& P = s->prob;
   memset(&P, 0, sizeof P);
   P.scheme = scheme; P.NE = m.NE; P.NN = m.NN; P.NNp = m.NNp; P.NENv = m.NENv; P.NENp = m.NENp; P.NGP = m.NGP;
   P.dt = m.dt; P.density = m.density; P.viscosity = m.viscosity; P.maxIter = m.maxIter; P.tolerance = m.tolerance;
   P.convergenceCriteria = m.converge;
   P.coord = m.coord.data(); P.LtoGnode = m.LtoGnode.data();
   P.sparseM_nnz = (int)s->A.Mp.nnz(); P.sparseMrowStarts = s->A.Mp.rowStarts.data(); P.sparseMcol = s->A.Mp.col.data();
   P.sparseKvalue = dev0 ? nullptr : s->A.Kval.data(); P.sparseMvalue = dev0 ? nullptr : s->A.Mval.data();
   P.sparseG_nnz = (int)s->A.Gp().nnz(); P.sparseGrowStarts = s->A.Gp().rowStarts.data(); P.sparseGcol = s->A.Gp().col.data();
   P.sparseG1value = dev0 ? nullptr : s->A.G1.data(); P.sparseG2value = dev0 ? nullptr : s->A.G2.data(); P.sparseG3value = dev0 ? nullptr : s->A.G3.data();
   P.sparseZ_nnz = (int)s->A.Zp().nnz(); P.sparseZrowStarts = s->A.Zp().rowStarts.data(); P.sparseZcol = s->A.Zp().col.data();
   P.sparseZvalue = dev0 ? nullptr : s->A.Zval.data();
   P.sparseMapM_1d = s->A.mapM.empty() ? nullptr : s->A.mapM.data();
   P.MdInv = s->A.MdInv.empty() ? nullptr : s->A.MdInv.data();
   P.MdOrigInv = s->A.MdOrigInv.empty() ? nullptr : s->A.MdOrigInv.data();
   P.BCnVelNodes = m.BCnVelNodes; P.BCvelNodes = m.BCvelNodes.data(); P.nBC = m.nBC; P.BCstr = m.BCstr.data();
   P.zeroPressureNode = m.zeroPressureNode; P.monPoint = m.monPoint;
   if (s->partitioned) {
      const Partition& pt = s->part;
      P.world = pt.world; P.rank = pt.rank; P.NN_owned = pt.NNo; P.NNp_owned = pt.NPo; P.NE_apply = pt.NEa;
      P.pinLocal = pt.pinLocal; P.monOwned = pt.monLocal >= 0 ? 1 : 0;
      s->sendVptr.assign(1, 0); s->sendPptr.assign(1, 0);
      for (const PeerPlan& pp : pt.peers) {
         s->peerRank.push_back(pp.rank);
         s->sendVidx.insert(s->sendVidx.end(), pp.sendV.begin(), pp.sendV.end());
         s->sendVptr.push_back((int)s->sendVidx.size());
         s->sendPidx.insert(s->sendPidx.end(), pp.sendP.begin(), pp.sendP.end());
         s->sendPptr.push_back((int)s->sendPidx.size());
         s->recvVoff.push_back(pp.recvVoff); s->recvVcnt.push_back(pp.recvVcnt);
         s->recvPoff.push_back(pp.recvPoff); s->recvPcnt.push_back(pp.recvPcnt);
         s->remoteVoff.push_back(pp.remoteVoff); s->remotePoff.push_back(pp.remotePoff);
      }
      P.nPeers = (int)pt.peers.size();
      P.peerRank = s->peerRank.data(); P.sendVptr = s->sendVptr.data(); P.sendVidx = s->sendVidx.data();
      P.recvVoff = s->recvVoff.data(); P.recvVcnt = s->recvVcnt.data();
      P.sendPptr = s->sendPptr.data(); P.sendPidx = s->sendPidx.data();
      P.recvPoff = s->recvPoff.data(); P.recvPcnt = s->recvPcnt.data();
      P.remoteVoff = s->remoteVoff.data(); P.remotePoff = s->remotePoff.data();
   }
   if (m.isRestart == 1 && !restartPath.empty()) {
      s->U0.assign((size_t)3 * m.NN, 0.0); s->P0.assign((size_t)m.NNp, 0.0);
      if (!read_restart(m, restartPath, s->U0.data(), s->P0.data())) return CFD_ERR_IO;
   }
   if (!s->U0.empty()) { P.U0 = s->U0.data(); P.P0 = s->P0.data(); }
   s->timeT = m.t_ini;
   if (!init_gpu) return CFD_OK;
   return cfd_initializeAndAllocateGPU(&P, &s->opt, &s->ctx);
}

}   // namespace

extern "C" {

int cfd_solver_create_from_inp(const char* inp_path, int scheme, const cfd_options* opt, int init_gpu, cfd_solver** out) {
   if (!inp_path || !out) { set_error("null argument"); return CFD_ERR_ARG; }
   *out = nullptr;
   cfd_solver* s = new cfd_solver();
   if (!read_inp(inp_path, s->mesh)) { delete s; return CFD_ERR_IO; }
   std::string base(inp_path);
   const size_t dot = base.rfind(".inp");
   if (dot != std::string::npos) base = base.substr(0, dot);
   int rc = finish_create(s, scheme, opt, base + "_restart.dat", init_gpu);
   if (rc) { delete s; return rc; }
   *out = s;
   return CFD_OK;
}

int cfd_solver_create_cavity(int nx, int ny, int nz, double cluster, int NENv, int scheme, double dt, double viscosity, int maxIter,
                             double tolerance, const cfd_options* opt, int init_gpu, cfd_solver** out) {
   if (!out || nx < 1 || ny < 1 || nz < 1 || (NENv != 8 && NENv != 27)) { set_error("bad cavity arguments"); return CFD_ERR_ARG; }
   if (cluster != 0.0 && ((nx | ny | nz) & 1)) { set_error("clustering needs even element counts"); return CFD_ERR_ARG; }
   *out = nullptr;
   cfd_solver* s = new cfd_solver();
   make_cavity(nx, ny, nz, cluster, NENv, s->mesh);
   s->mesh.dt = dt; s->mesh.viscosity = viscosity; s->mesh.maxIter = maxIter; s->mesh.tolerance = tolerance;
   s->mesh.converge = 1e-6; s->mesh.t_final = 1e30; s->mesh.density = 1.0;
   int rc = finish_create(s, scheme, opt, "", init_gpu);
   if (rc) { delete s; return rc; }
   *out = s;
   return CFD_OK;
}

int cfd_solver_create_from_inp_part(const char* inp_path, int scheme, const cfd_options* opt, int init_gpu, int world, int rank,
                                    cfd_solver** out) {
   if (!inp_path || !out) { set_error("null argument"); return CFD_ERR_ARG; }
   *out = nullptr;
   cfd_solver* s = new cfd_solver();
   if (!read_inp(inp_path, s->mesh)) { delete s; return CFD_ERR_IO; }
   std::string base(inp_path);
   const size_t dot = base.rfind(".inp");
   if (dot != std::string::npos) base = base.substr(0, dot);
   int rc = finish_create(s, scheme, opt, base + "_restart.dat", init_gpu, world, rank);
   if (rc) { delete s; return rc; }
   *out = s;
   return CFD_OK;
}

int cfd_solver_create_cavity_part(int nx, int ny, int nz, double cluster, int NENv, int scheme, double dt, double viscosity, int maxIter,
                                  double tolerance, const cfd_options* opt, int init_gpu, int world, int rank, cfd_solver** out) {
   if (!out || nx < 1 || ny < 1 || nz < 1 || (NENv != 8 && NENv != 27)) { set_error("bad cavity arguments"); return CFD_ERR_ARG; }
   *out = nullptr;
   cfd_solver* s = new cfd_solver();
   make_cavity(nx, ny, nz, cluster, NENv, s->mesh);
   s->mesh.dt = dt; s->mesh.viscosity = viscosity; s->mesh.maxIter = maxIter; s->mesh.tolerance = tolerance;
   s->mesh.converge = 1e-6; s->mesh.t_final = 1e30; s->mesh.density = 1.0;
   int rc = finish_create(s, scheme, opt, "", init_gpu, world, rank);
   if (rc) { delete s; return rc; }
   *out = s;
   return CFD_OK;
}

int cfd_solver_init_gpu(cfd_solver* s) {
   if (!s) return CFD_ERR_ARG;
   if (s->ctx) return CFD_OK;
   return cfd_initializeAndAllocateGPU(&s->prob, &s->opt, &s->ctx);
}

void cfd_solver_destroy(cfd_solver* s) {
   if (!s) return;
   cfd_destroy(s->ctx);
   delete s;
}
cfd_ctx* cfd_solver_ctx(cfd_solver* s) { return s ? s->ctx : nullptr; }
const cfd_problem* cfd_solver_problem(cfd_solver* s) { return s ? &s->prob : nullptr; }

long long cfd_solver_int(cfd_solver* s, const char* name) {
   const std::string n(name);
   const Mesh& m = s->mesh;
   if (n == "NE") return m.NE; if (n == "NCN") return m.NCN; if (n == "NN") return m.NN; if (n == "NNp") return m.NNp;
   if (n == "globalNN") return s->globalNN; if (n == "globalNNp") return s->globalNNp; if (n == "globalNE") return s->globalNE;
   if (n == "NN_owned") return s->partitioned ? s->part.NNo : m.NN; if (n == "NNp_owned") return s->partitioned ? s->part.NPo : m.NNp;
   if (n == "NE_apply") return s->partitioned ? s->part.NEa : m.NE; if (n == "nPeers") return (long long)s->part.peers.size();
   if (n == "pinLocal") return s->partitioned ? s->part.pinLocal : (m.zeroPressureNode > 0 ? m.zeroPressureNode : -1);
   if (n == "monLocal") return s->partitioned ? s->part.monLocal : m.monPoint;
   if (n == "NENv") return m.NENv; if (n == "nnzM") return s->A.Mp.nnz(); if (n == "nnzG") return s->A.Gp().nnz();
   if (n == "nnzZ") return s->A.Zp().nnz(); if (n == "BCnVelNodes") return m.BCnVelNodes; if (n == "monPoint") return m.monPoint;
   if (n == "zeroPressureNode") return m.zeroPressureNode; if (n == "maxIter") return m.maxIter; if (n == "isRestart") return m.isRestart;
   if (n == "nActiveColors") { if (m.meshColors.empty()) color_mesh(s->mesh); return s->mesh.nActiveColors; }
   return -999999;
}

const void* cfd_solver_array(cfd_solver* s, const char* name, long long* count, int* is_int) {
   const std::string n(name);
   Mesh& m = s->mesh;
   Assembled& A = s->A;
#define DV(nm, v) if (n == nm) { *count = (long long)(v).size(); *is_int = 0; return (v).data(); }
#define IV(nm, v) if (n == nm) { *count = (long long)(v).size(); *is_int = 1; return (v).data(); }
   IV("part_nodeGlobal", s->part.nodeGlobal) IV("part_presGlobal", s->part.presGlobal) IV("part_elemGlobal", s->part.elemGlobal)
   IV("peerRank", s->peerRank) IV("sendVptr", s->sendVptr) IV("sendVidx", s->sendVidx) IV("recvVoff", s->recvVoff) IV("recvVcnt", s->recvVcnt)
   IV("remoteVoff", s->remoteVoff) IV("remotePoff", s->remotePoff) IV("sendPptr", s->sendPptr) IV("sendPidx", s->sendPidx) IV("recvPoff", s->recvPoff) IV("recvPcnt", s->recvPcnt)
   DV("U0", s->U0) DV("P0", s->P0)
   DV("coord", m.coord) IV("LtoGnode", m.LtoGnode) IV("BCvelNodes", m.BCvelNodes) DV("BCstr", m.BCstr)
   IV("sparseMrowStarts", A.Mp.rowStarts) IV("sparseMcol", A.Mp.col) IV("sparseGrowStarts", A.Gp().rowStarts) IV("sparseGcol", A.Gp().col)
   IV("sparseZrowStarts", A.Zp().rowStarts) IV("sparseZcol", A.Zp().col) IV("sparseMapM_1d", A.mapM) IV("sparseMapG_1d", A.mapG())
   DV("sparseMvalue", A.Mval) DV("sparseKvalue", A.Kval) DV("sparseG1value", A.G1) DV("sparseG2value", A.G2) DV("sparseG3value", A.G3)
   DV("sparseZvalue", A.Zval) DV("MdInv", A.MdInv) DV("MdOrigInv", A.MdOrigInv)
   if (n == "meshColors" || n == "NmeshColors" || n == "elementsOfColor") {
      if (m.meshColors.empty()) color_mesh(m);
      IV("meshColors", m.meshColors) IV("NmeshColors", m.NmeshColors) IV("elementsOfColor", m.elementsOfColor)
   }
#undef DV
#undef IV
   *count = -1; *is_int = 0;
   return nullptr;
}

int cfd_solver_write_tecplot(cfd_solver* s, const char* dat_path, const char* title) {
   std::vector<double> U((size_t)3 * s->mesh.NN), P((size_t)s->mesh.NNp);
   int rc = cfd_get_fields(s->ctx, U.data(), P.data());
   if (rc) return rc;
   return write_tecplot(s->mesh, U.data(), P.data(), dat_path, title) ? CFD_OK : CFD_ERR_IO;
}

int cfd_solver_write_tecplot_fields(cfd_solver* s, const char* dat_path, const char* title, const double* U, const double* P) {
   if (!s || !U || !P) return CFD_ERR_ARG;
   return write_tecplot(s->mesh, U, P, dat_path, title) ? CFD_OK : CFD_ERR_IO;
}

int cfd_solver_read_restart_fields(cfd_solver* s, const char* dat_path, double* U, double* P) {
   if (!s || !U || !P) return CFD_ERR_ARG;
   return read_restart(s->mesh, dat_path, U, P) ? CFD_OK : CFD_ERR_IO;
}

int cfd_solver_read_restart(cfd_solver* s, const char* dat_path) {
   std::vector<double> U((size_t)3 * s->mesh.NN, 0.0), P((size_t)s->mesh.NNp, 0.0);
   if (!read_restart(s->mesh, dat_path, U.data(), P.data())) return CFD_ERR_IO;
   const bool ex = s->scheme == CFD_EXPLICIT;
   int rc = cfd_upload(s->ctx, ex ? "Un" : "Uk_prev", U.data(), (long long)U.size());
   if (rc) return rc;
   return cfd_upload(s->ctx, ex ? "Pn" : "Pk_prev", P.data(), (long long)P.size());
}

// The reference's main() + timeLoop() (BCH.cpp:383-520, 2815-3120 / GQ.cpp:367-500, 3266-3418):
// ./ProblemName.txt -> <name>.inp -> monitor table on stdout -> <name>.dat.
int cfd_solver_main(const char* workdir, int scheme, int max_steps) {
   const std::string dir = (workdir && *workdir) ? std::string(workdir) + "/" : std::string();
   std::ifstream pn((dir + "ProblemName.txt").c_str());
   std::string name;
   if (!(pn >> name)) { set_error("cannot read %sProblemName.txt", dir.c_str()); return CFD_ERR_IO; }
   printf("\n\n*********************************************************");
   printf("\n*    3D Unsteady Incompressible Navier-Stokes Solver    *");
   printf("\n*  %s  *", scheme == CFD_EXPLICIT ? "     Formulation of Blasco, Codina & Huerta        " : "     Formulation of Guermond & Quartapelle         ");
   printf("\n*        B200-native fractional-step time loop          *");
   printf("\n*********************************************************\n\n");
   double t0 = now_s();
   cfd_solver* s = nullptr;
   // the reference hard-codes its solver settings, differently in each of its four builds: the program
   // takes the GPU builds' (the defaults) and lets the environment select the others
   cfd_options opt;
   cfd_default_options(&opt);
   if (const char* v = getenv("CFD_CG_TOL")) opt.cg_tol = atof(v);
   if (const char* v = getenv("CFD_CG_MAXIT")) opt.cg_maxit = atoi(v);
   if (const char* v = getenv("CFD_CG_WARM_START")) opt.cg_warm_start = atoi(v);
   if (const char* v = getenv("CFD_CG_NO_PRECOND")) opt.cg_no_precond = atoi(v);
   if (const char* v = getenv("CFD_BICG_TOL")) opt.bicg_tol = atof(v);
   if (const char* v = getenv("CFD_BICG_MAXIT")) opt.bicg_maxit = atoi(v);
   if (const char* v = getenv("CFD_MAXACC_RULE")) opt.maxacc_rule = atoi(v);
   if (const char* v = getenv("CFD_DEVICE")) opt.device = atoi(v);
   if (const char* v = getenv("CFD_EBE_SCATTER")) opt.ebe_scatter = atoi(v);
   if (const char* v = getenv("CFD_DETERMINISTIC")) opt.deterministic = atoi(v);
   if (const char* v = getenv("CFD_BICG_NO_HALFSTEP")) opt.bicg_no_halfstep = atoi(v);
   if (const char* v = getenv("CFD_BICG_NO_VERIFY")) opt.bicg_no_verify = atoi(v);
   if (const char* v = getenv("CFD_CG_VARIANT")) opt.cg_variant = atoi(v);
   int rc = cfd_solver_create_from_inp((dir + name + ".inp").c_str(), scheme, &opt, 1, &s);
   if (rc) return rc;
   const Mesh& m = s->mesh;
   printf("readInputFile() .. step0() took  %8.3f seconds.\n", now_s() - t0);
   printf("\n NN = %d\n NNp = %d\n sparseM_NNZ = %lld\n sparseG_NNZ = %lld\n", m.NN, m.NNp, 3 * s->A.Mp.nnz(), 3 * s->A.Gp().nnz());
   printf("\n\nMonitoring node is %d, with coordinates [%f, %f, %f]\n\n\n", m.monPoint, m.coord[3 * (size_t)m.monPoint],
          m.coord[3 * (size_t)m.monPoint + 1], m.coord[3 * (size_t)m.monPoint + 2]);
   const std::string dat = dir + name + ".dat";
   if ((rc = cfd_solver_write_tecplot(s, dat.c_str(), name.c_str()))) { cfd_solver_destroy(s); return rc; }   // initial state (BCH.cpp:2801)
   if (scheme == CFD_EXPLICIT) {
      printf("Time step  Iter     Time       u_monitor     v_monitor     w_monitor     p_monitor     TimeSpend      maxAcc \n");
      printf("-------------------------------------------------------------------------------------------------------------\n");
   } else {
      printf("Time step  Time       u_monitor     v_monitor     w_monitor     p_monitor     TimeSpend      maxAcc \n");
      printf("----------------------------------------------------------------------------------------------------\n");
   }
   double timeT = m.t_ini;
   int timeN = 0;
   const double tStart = now_s();
   while (timeT < m.t_final) {
      if (max_steps > 0 && timeN >= max_steps) break;
      const double ts = now_s();
      timeN++;
      timeT += m.dt;
      cfd_step_report rep;
      int done = 0;
      if ((rc = cfd_timeLoop(s->ctx, 1, 0, &rep, &done))) { cfd_solver_destroy(s); return rc; }
      const double spent = now_s() - ts;
      if (scheme == CFD_EXPLICIT)
         printf("%6d  %6d  %10.5f  %12.5f  %12.5f  %12.5f  %12.5f  %12.5f  %12.5f\n", timeN, rep.iter, timeT, rep.monitor[0], rep.monitor[1],
                rep.monitor[2], rep.monitor[3], spent, rep.maxAcc);
      else
         printf("%6d  %10.5f  %12.5f  %12.5f  %12.5f  %12.5f  %12.5f  %12.5f\n", timeN, timeT, rep.monitor[0], rep.monitor[1], rep.monitor[2],
                rep.monitor[3], spent, rep.maxAcc);
      if (rep.steady) {
         cfd_solver_write_tecplot(s, dat.c_str(), name.c_str());
         break;
      }
      if (timeN % 1000 == 0 || std::fabs(timeT - m.t_final) < 1e-10) cfd_solver_write_tecplot(s, dat.c_str(), name.c_str());
   }
   printf("\nTotal run took %10.3f seconds.\n", now_s() - tStart);
   printf("\nThe program is terminated successfully.\n\n\n");
   rc = cfd_solver_write_tecplot(s, dat.c_str(), name.c_str());   // BCH.cpp:517
   cfd_solver_destroy(s);
   return rc;
}

}   // extern "C"
