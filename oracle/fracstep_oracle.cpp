// =============================================================================
// fracstep_oracle.cpp -- TEST INFRASTRUCTURE ONLY (CPU oracle). NOT the product.
//
// A dependency-free CPU restatement of the reference's fractional-step FE
// Navier-Stokes solvers (CPU branches, `#ifndef USECUDA`):
//   BCH = fractionalStep/explicit/Cpp/blascoCodinaHuerta.cpp
//   GQ  = fractionalStep/implicit/Cpp/guermondQuartapelle.cpp
// Every function cites the reference file:line it follows.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library; the product (cfd-with-cuda_b200/) never does.
//
// PINNED AGAINST THE REFERENCE ITSELF (tests/test_reference_pin.py): the reference's two programs are compiled
// unmodified from /root/reference against stand-ins for Intel MKL, CSparse and PARALUTION (oracle/shim/, `make ref`
// -> oracle/_ref/bch_cpu, gq_cpu) and run on generated inputs and on the shipped lidDrivenCavity_NE27000.inp; this
// restatement reproduces the .dat files they write to 1e-12..1e-15 (explicit, identical CG iteration counts) and
// 1e-8..1e-11 (implicit with both sides iterated to 1e-12).  Fixtures: tests/golden/ref_*.npz, generator
// tests/golden/make_ref_golden.py.
// STILL UNPINNED -- the arithmetic INSIDE the third-party libraries, which neither the reference nor this image holds
// (MKL / CSparse / PARALUTION 1.0.0 / CUSP absent).  It is restated from their published algorithms, here and in oracle/shim:
//   * mkl_dcsrmv            -> plain CSR SpMV (row-sequential sums)
//   * CSparse cs_multiply   -> row-wise (Gustavson) sparse product, same pattern
//   * MKL RCI dcg / cusp::krylov::cg + diagonal precond -> textbook Jacobi-PCG,
//       stop when ||r||_2 <= tol*||b||_2 (checked before every iteration), <=1000
//   * Paralution BiCGStab / cusp::krylov::bicgstab + Jacobi -> textbook
//       right-preconditioned BiCGStab, x0 = 0, same stopping rule
//   (which inner product MKL's preconditioned stopping test uses, and where in its iteration PARALUTION tests the
//    residual, cannot be verified here; on the reference's under-integrated implicit Q2 momentum matrix the latter moves
//    the fields by ~1e-5 at the reference's 1e-6 tolerance).
// Also pinned (tests/test_oracle.py): closed-form node / nnz counts of the shipped lidDrivenCavity_NE27000.inp,
// colouring, BC-node counts, analytic FE invariants (partition of unity, K row sums, lumped-mass volume, G*1 = 0, ...).
// =============================================================================
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <map>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

typedef std::vector<double> dvec;
typedef std::vector<int> ivec;

// Corner pairs of the 12 edges and corner quadruples of the 6 faces
// (BCH.cpp:1005-1054 and 1143-1180).
const int kEdge[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {0, 4}, {1, 5},
                          {2, 6}, {3, 7}, {4, 5}, {5, 6}, {6, 7}, {7, 4}};
const int kFace[6][4] = {{0, 1, 2, 3}, {0, 1, 4, 5}, {1, 2, 5, 6},
                         {2, 3, 6, 7}, {0, 3, 4, 7}, {4, 5, 6, 7}};
// Non-corner nodes lying on each face (BCH.cpp:1485-1528).
const int kFaceMid[6][5] = {{8, 9, 10, 11, 20},  {8, 12, 13, 16, 21},
                            {9, 13, 14, 17, 22}, {10, 14, 15, 18, 23},
                            {11, 12, 15, 19, 24}, {16, 17, 18, 19, 25}};

struct Csr {
   int nrows = 0, ncols = 0;
   ivec rowStarts, col, row;
   int nnz() const { return (int)col.size(); }
};

struct Options {
   double cg_tol = 1e-6;
   int cg_maxit = 1000;
   int cg_warm = 0;       // 1 = explicit GPU/CUSP behaviour (BCH.cu:418-451): x0 = previous Pdot
   int cg_precond = 1;    // 0 = implicit CPU build (GQ.cpp:4514 Jacobi line commented out)
   double bicg_tol = 1e-6;
   int bicg_maxit = 1000;
   int bicg_halfstep = 1;   // 1 = also test |s| after the half step (CUSP, GQ.cu:404-414); 0 = once per iteration (PARALUTION, GQ.cpp:4190-4229)
   int maxacc_rule = 0;   // 0: max|dU|/dt (BCH.cpp:3053); 1: max|dU| (BCH.cu:107-117)
};

struct Orc {
   int scheme = 0;   // 0 explicit (BCH), 1 implicit (GQ)
   Options opt;
   // ---- input file (BCH.cpp:528-725)
   int eType = 1, NE = 0, NCN = 0, NENv = 0, NENp = 0, NGP = 0;
   double alpha = 0, dt = 0, t_ini = 0, t_final = 0;
   int maxIter = 0;
   double tolerance = 0, convergenceCriteria = 0;
   int isRestart = 0;
   double density = 1, viscosity = 0, fx = 0, fy = 0;
   int NEC = 8, NEF = 6, NEE = 12;
   dvec coord;      // [node][3]
   ivec LtoGnode;   // [e][NENv]
   int nBC = 0;
   dvec BCtype, BCstr;   // BCstr [nBC][3]
   int BCnVelFaces = 0, BCnOutFaces = 0;
   ivec BCvelFaces, BCoutFaces;   // [i][3], 0-based
   int zeroPressureNode = -1;
   double monPointCoord[3] = {0, 0, 0};
   std::string name, dir;
   // ---- preprocessing
   int NN = 0, NNp = 0;
   std::vector<ivec> elemsOfPresNodes, elemNeighbors;
   ivec meshColors, NmeshColors, elementsOfColor;
   int nActiveColors = 0;
   int colorConflicts = 0;   // pairs of corner-sharing elements that received the same colour
   ivec LtoGvel_1d;   // [e][3*NENv]
   ivec BCvelNodes;   // [i][2] = node, whichBC
   int BCnVelNodes = 0, monPoint = 0;
   Csr Mp, Gp, Zp;    // patterns: NNxNN, NNxNNp, NNpxNNp
   ivec sparseMapM, sparseMapG, sparseMapZ;   // [e][i][j]
   // ---- reference element + geometry
   dvec GQpoint, GQweight;   // [k][3], [k]
   dvec Sv, Sp;              // [k][i]
   dvec dSv, dSp;            // [d][i][k]
   dvec detJacob;            // [e][k]
   dvec gDSv, gDSp;          // [e][k][i][d]
   dvec GQfactor_1d, Sv_1d, gDSv_1d;
   // ---- matrices
   dvec Mval, Kval, Aval, G1, G2, G3, Zval;
   dvec Md, MdInv, MdOrig, MdOrigInv;
   // ---- explicit state (BCH.cpp:2718-2808)
   dvec Un, Unp1, Unp1_prev, UnpHalf, UnpHalf_prev, Acc, Acc_prev, KtimesAcc_prev, R1, R3;
   dvec Pn, Pnp1, Pnp1_prev, Pdot, R2;
   // ---- implicit state (GQ.cpp:3206-3259)
   dvec Uk, Uk_prev, Pk, Pk_prev, Pk_prevprev, Pdiff;
   int direction = 0;
   // ---- bookkeeping
   int timeN = 0;
   double timeT = 0, maxAcc = 0;
   int lastIter = 0;
   ivec cgIters;     // one entry per CG solve
   ivec bicgIters;   // one entry per BiCGStab solve (implicit: 3 per step)
   ivec innerIters;  // explicit: iter value printed per time step
   double cgSeconds = 0;
   std::string err;
};

double wall() {
#ifdef _OPENMP
   return omp_get_wtime();
#else
   return (double)clock() / CLOCKS_PER_SEC;
#endif
}

// ---------------------------------------------------------------------------
// readInputFile  (BCH.cpp:528-725, GQ.cpp:537-740)
// ---------------------------------------------------------------------------
bool readInputFile(Orc& o, const std::string& path) {
   std::ifstream f(path.c_str());
   if (!f) { o.err = "cannot open " + path; return false; }
   auto skipLine = [&]() { f.ignore(1 << 20, '\n'); };
   auto afterColon = [&]() -> std::istream& { f.ignore(256, ':'); return f; };
   skipLine(); skipLine();
   afterColon() >> o.eType; skipLine();
   afterColon() >> o.NE; skipLine();
   afterColon() >> o.NCN; skipLine();
   afterColon() >> o.NENv; skipLine();
   afterColon() >> o.NENp; skipLine();
   afterColon() >> o.NGP; skipLine();
   afterColon() >> o.alpha; skipLine();
   afterColon() >> o.dt; skipLine();
   afterColon() >> o.t_ini; skipLine();
   afterColon() >> o.t_final; skipLine();
   afterColon() >> o.maxIter; skipLine();
   afterColon() >> o.tolerance; skipLine();
   afterColon() >> o.convergenceCriteria; skipLine();
   afterColon() >> o.isRestart; skipLine();
   afterColon() >> o.density; skipLine();
   afterColon() >> o.viscosity; skipLine();
   afterColon() >> o.fx; skipLine();
   afterColon() >> o.fy; skipLine();
   if (!f || o.eType != 1 || o.NENp != 8 || (o.NENv != 27 && o.NENv != 8) || o.NGP != 8) {
      o.err = "unsupported or malformed header in " + path;
      return false;
   }
   o.coord.assign((size_t)o.NE * o.NENv * 3, 0.0);   // over-allocated like BCH.cpp:568
   skipLine(); skipLine();
   int id;
   for (int i = 0; i < o.NCN; i++) {
      f >> id >> o.coord[3 * i] >> o.coord[3 * i + 1] >> o.coord[3 * i + 2];
      skipLine();
   }
   o.LtoGnode.assign((size_t)o.NE * o.NENv, -1);
   skipLine(); skipLine();
   for (int e = 0; e < o.NE; e++) {
      f >> id;
      for (int i = 0; i < o.NEC; i++) {
         int n; f >> n;
         o.LtoGnode[(size_t)e * o.NENv + i] = n - 1;
      }
      skipLine();
   }
   skipLine(); skipLine();
   afterColon() >> o.nBC; skipLine();
   o.BCtype.assign(o.nBC, 0.0);
   o.BCstr.assign(3 * o.nBC, 0.0);
   std::string dummy;
   for (int i = 0; i < o.nBC; i++) {
      f.ignore(256, ':');
      f >> o.BCtype[i];
      f >> o.BCstr[3 * i]; f >> dummy;
      f >> o.BCstr[3 * i + 1]; f >> dummy;
      f >> o.BCstr[3 * i + 2];
      skipLine();
   }
   skipLine();
   afterColon() >> o.BCnVelFaces; skipLine();
   afterColon() >> o.BCnOutFaces; skipLine();
   skipLine(); skipLine();
   o.BCvelFaces.resize(3 * (size_t)o.BCnVelFaces);
   for (int i = 0; i < o.BCnVelFaces; i++) {
      int a, b, c; f >> a >> b >> c;
      o.BCvelFaces[3 * i] = a - 1; o.BCvelFaces[3 * i + 1] = b - 1; o.BCvelFaces[3 * i + 2] = c - 1;
      skipLine();
   }
   skipLine(); skipLine();
   o.BCoutFaces.resize(3 * (size_t)o.BCnOutFaces);
   for (int i = 0; i < o.BCnOutFaces; i++) {
      int a, b, c; f >> a >> b >> c;
      o.BCoutFaces[3 * i] = a - 1; o.BCoutFaces[3 * i + 1] = b - 1; o.BCoutFaces[3 * i + 2] = c - 1;
      skipLine();
   }
   skipLine(); skipLine();
   f >> o.zeroPressureNode; o.zeroPressureNode -= 1; skipLine();
   skipLine(); skipLine();
   f >> o.monPointCoord[0] >> o.monPointCoord[1] >> o.monPointCoord[2];
   if (!f) { o.err = "malformed body in " + path; return false; }
   o.NNp = (o.NENp == 1) ? o.NE : o.NCN;
   return true;
}

// findElemsOfPresNodes (BCH.cpp:732-783) + findElemNeighbors (BCH.cpp:790-846)
void findNeighbors(Orc& o) {
   o.elemsOfPresNodes.assign(o.NNp, ivec());
   for (int e = 0; e < o.NE; e++)
      for (int i = 0; i < o.NENp; i++) o.elemsOfPresNodes[o.LtoGnode[(size_t)e * o.NENv + i]].push_back(e);
   o.elemNeighbors.assign(o.NE, ivec());
   for (int e = 0; e < o.NE; e++) {
      ivec& nb = o.elemNeighbors[e];
      for (int i = 0; i < o.NEC; i++) {
         int node = o.LtoGnode[(size_t)e * o.NENv + i];
         for (int elem : o.elemsOfPresNodes[node]) {
            if (elem == e) continue;
            if (std::find(nb.begin(), nb.end(), elem) == nb.end()) nb.push_back(elem);
         }
      }
   }
}

// setupMeshColoring (BCH.cpp:853-947).  The restart quirk at :880-886 (`i = 0`
// followed by the loop's own i++) is kept: after a clash the scan resumes at
// neighbour 1, so neighbour 0 is not rechecked.
void setupMeshColoring(Orc& o) {
   const int LARGE = 30;
   o.meshColors.assign(o.NE, -1);
   o.NmeshColors.assign(LARGE, 0);
   o.elementsOfColor.assign(o.NE, 0);
   for (int e = 0; e < o.NE; e++) {
      int cand = 0;
      const ivec& nb = o.elemNeighbors[e];
      for (int i = 0; i < (int)nb.size(); i++) {
         if (cand == o.meshColors[nb[i]]) { cand++; i = 0; }
      }
      if (cand >= LARGE) cand = LARGE - 1;
      o.meshColors[e] = cand;
      o.NmeshColors[cand]++;
   }
   int count = 0;
   o.nActiveColors = 0;
   for (int c = 0; c < LARGE; c++) {
      if (o.NmeshColors[c] > 0) {
         o.nActiveColors++;
         for (int e = 0; e < o.NE; e++)
            if (o.meshColors[e] == c) o.elementsOfColor[count++] = e;
      }
   }
   // The restart quirk above leaves same-coloured neighbours on some meshes (e.g. odd element counts
   // per edge: 8 conflicting pairs on a 5^3 cube) -- a data race in the reference's coloured OpenMP /
   // CUDA loops.  The oracle must be deterministic: when the colouring is
   // invalid the element loops below run serially, i.e. they compute what the reference computes
   // when the race does not strike.
   o.colorConflicts = 0;
   for (int e = 0; e < o.NE; e++)
      for (int nb : o.elemNeighbors[e])
         if (nb > e && o.meshColors[nb] == o.meshColors[e]) o.colorConflicts++;
}

// setupNonCornerNodes (BCH.cpp:954-1320): mid-edge, then mid-face, then centre
// nodes; an existing node is reused only when a NEIGHBOUR element already owns
// one within 1e-10 of the same coordinates.
void setupNonCornerNodes(Orc& o) {
   if (o.NENv == o.NENp) { o.NN = o.NCN; o.coord.resize((size_t)o.NN * 3); return; }
   const double SMALL = 1e-10;
   int nodeCount = o.NCN;
   const int nv = o.NENv;
   auto X = [&](int n, int d) -> double& { return o.coord[(size_t)3 * n + d]; };
   auto pass = [&](int first, int count, bool isEdge) {
      ivec cand;
      for (int e = 0; e < o.NE; e++) {
         cand.clear();
         for (int ne : o.elemNeighbors[e])
            for (int i = first; i < first + count; i++) {
               int g = o.LtoGnode[(size_t)ne * nv + i];
               if (g != -1) cand.push_back(g);
            }
         for (int q = 0; q < count; q++) {
            double mid[3];
            for (int d = 0; d < 3; d++) {
               if (isEdge) {
                  int n1 = o.LtoGnode[(size_t)e * nv + kEdge[q][0]], n2 = o.LtoGnode[(size_t)e * nv + kEdge[q][1]];
                  mid[d] = 0.5 * (X(n1, d) + X(n2, d));
               } else {
                  const int* c = kFace[q];
                  mid[d] = 0.25 * (X(o.LtoGnode[(size_t)e * nv + c[0]], d) + X(o.LtoGnode[(size_t)e * nv + c[1]], d) +
                                   X(o.LtoGnode[(size_t)e * nv + c[2]], d) + X(o.LtoGnode[(size_t)e * nv + c[3]], d));
               }
            }
            int found = -1;
            for (int g : cand) {
               if (std::fabs(mid[0] - X(g, 0)) < SMALL && std::fabs(mid[1] - X(g, 1)) < SMALL &&
                   std::fabs(mid[2] - X(g, 2)) < SMALL) { found = g; break; }
            }
            if (found < 0) {
               found = nodeCount++;
               for (int d = 0; d < 3; d++) X(found, d) = mid[d];
            }
            o.LtoGnode[(size_t)e * nv + first + q] = found;
         }
      }
   };
   pass(o.NEC, o.NEE, true);
   pass(o.NEC + o.NEE, o.NEF, false);
   for (int e = 0; e < o.NE; e++) {
      for (int d = 0; d < 3; d++) {
         double s = 0;
         // same left-to-right sum as BCH.cpp:1246-1248
         s = X(o.LtoGnode[(size_t)e * nv + 0], d) + X(o.LtoGnode[(size_t)e * nv + 1], d) + X(o.LtoGnode[(size_t)e * nv + 2], d) +
             X(o.LtoGnode[(size_t)e * nv + 3], d) + X(o.LtoGnode[(size_t)e * nv + 4], d) + X(o.LtoGnode[(size_t)e * nv + 5], d) +
             X(o.LtoGnode[(size_t)e * nv + 6], d) + X(o.LtoGnode[(size_t)e * nv + 7], d);
         X(nodeCount, d) = 0.125 * s;
      }
      o.LtoGnode[(size_t)e * nv + o.NEC + o.NEE + o.NEF] = nodeCount++;
   }
   o.NN = nodeCount;
   o.coord.resize((size_t)o.NN * 3);
}

// setupLtoGdof (BCH.cpp:1327-1403): component-blocked velocity DOFs.
void setupLtoGdof(Orc& o) {
   const int nv = o.NENv;
   o.LtoGvel_1d.resize((size_t)o.NE * 3 * nv);
   for (int e = 0; e < o.NE; e++)
      for (int c = 0; c < 3; c++)
         for (int i = 0; i < nv; i++) o.LtoGvel_1d[(size_t)e * 3 * nv + c * nv + i] = c * o.NN + o.LtoGnode[(size_t)e * nv + i];
}

// determineVelBCnodes (BCH.cpp:1410-1580): later faces overwrite earlier ones.
void determineVelBCnodes(Orc& o) {
   dvec info(o.NN, -1.0);
   const int nv = o.NENv;
   for (int i = 0; i < o.BCnVelFaces; i++) {
      int e = o.BCvelFaces[3 * i], f = o.BCvelFaces[3 * i + 1], w = o.BCvelFaces[3 * i + 2];
      for (int q = 0; q < 4; q++) info[o.LtoGnode[(size_t)e * nv + kFace[f][q]]] = w;
      if (o.NENp != o.NENv)
         for (int q = 0; q < 5; q++) info[o.LtoGnode[(size_t)e * nv + kFaceMid[f][q]]] = w;
   }
   o.BCvelNodes.clear();
   for (int i = 0; i < o.NN; i++)
      if (info[i] != -1) { o.BCvelNodes.push_back(i); o.BCvelNodes.push_back((int)info[i]); }
   o.BCnVelNodes = (int)o.BCvelNodes.size() / 2;
}

// findMonitorPoint (BCH.cpp:1646-1668): nearest CORNER node, first wins on ties.
void findMonitorPoint(Orc& o) {
   double distance = 1e6;
   for (int i = 0; i < o.NCN; i++) {
      double dx = o.coord[3 * i] - o.monPointCoord[0], dy = o.coord[3 * i + 1] - o.monPointCoord[1],
             dz = o.coord[3 * i + 2] - o.monPointCoord[2];
      double d = std::sqrt(dx * dx + dy * dy + dz * dz);
      if (d < distance) { distance = d; o.monPoint = i; }
   }
}

// setupSparseM / setupSparseG (BCH.cpp:1675-2159), setupSparseZGPU (GQ.cpp:2418-2647).
// Same result as the reference (sorted columns per row, [e][i][j] position maps),
// built by a per-row gather instead of the O(NN*216) NZcolsInARow temporary.
void buildPattern(Orc& o, int nrows, int ncols, int nRowLoc, int nColLoc, Csr& P, ivec& map) {
   const int nv = o.NENv;
   std::vector<ivec> elemsOfRow(nrows);
   for (int e = 0; e < o.NE; e++)
      for (int i = 0; i < nRowLoc; i++) elemsOfRow[o.LtoGnode[(size_t)e * nv + i]].push_back(e);
   P.nrows = nrows; P.ncols = ncols;
   P.rowStarts.assign(nrows + 1, 0);
   std::vector<ivec> cols(nrows);
   for (int r = 0; r < nrows; r++) {
      ivec& c = cols[r];
      for (int e : elemsOfRow[r])
         for (int j = 0; j < nColLoc; j++) c.push_back(o.LtoGnode[(size_t)e * nv + j]);
      std::sort(c.begin(), c.end());
      c.erase(std::unique(c.begin(), c.end()), c.end());
      P.rowStarts[r + 1] = P.rowStarts[r] + (int)c.size();
   }
   P.col.resize(P.rowStarts[nrows]);
   P.row.resize(P.rowStarts[nrows]);
   for (int r = 0; r < nrows; r++) {
      std::copy(cols[r].begin(), cols[r].end(), P.col.begin() + P.rowStarts[r]);
      std::fill(P.row.begin() + P.rowStarts[r], P.row.begin() + P.rowStarts[r + 1], r);
   }
   map.resize((size_t)o.NE * nRowLoc * nColLoc);
   for (int e = 0; e < o.NE; e++)
      for (int i = 0; i < nRowLoc; i++) {
         int r = o.LtoGnode[(size_t)e * nv + i];
         const int* b = &P.col[P.rowStarts[r]];
         const int* en = &P.col[0] + P.rowStarts[r + 1];
         for (int j = 0; j < nColLoc; j++) {
            int c = o.LtoGnode[(size_t)e * nv + j];
            map[((size_t)e * nRowLoc + i) * nColLoc + j] = (int)(std::lower_bound(b, en, c) - &P.col[0]);
         }
      }
}

// setupGQ (BCH.cpp:2166-2208)
void setupGQ(Orc& o) {
   const double g = std::sqrt(1. / 3);
   o.GQpoint.resize(8 * 3);
   o.GQweight.assign(8, 1.0);
   for (int k = 0; k < 8; k++) {
      o.GQpoint[3 * k + 0] = (k & 1) ? g : -g;
      o.GQpoint[3 * k + 1] = (k & 2) ? g : -g;
      o.GQpoint[3 * k + 2] = (k & 4) ? g : -g;
   }
}

// calcShape (BCH.cpp:2215-2488).  The 27 triquadratic functions are written as
// coef * fx(ksi) * fy(eta) * fz(zeta) with f in {(x^2-x), (1-x^2), (x^2+x)},
// the same factor order as BCH.cpp:2312-2343; node positions from Appendix D.
void calcShape(Orc& o) {
   const int G = o.NGP;
   o.Sp.assign(G * 8, 0); o.dSp.assign(3 * 8 * G, 0);
   static const int cs[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                                {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};
   for (int k = 0; k < G; k++) {
      double x = o.GQpoint[3 * k], y = o.GQpoint[3 * k + 1], z = o.GQpoint[3 * k + 2];
      for (int i = 0; i < 8; i++) {
         double a = cs[i][0], b = cs[i][1], c = cs[i][2];
         o.Sp[k * 8 + i] = 0.125 * (1 + a * x) * (1 + b * y) * (1 + c * z);
         o.dSp[(0 * 8 + i) * G + k] = a * 0.125 * (1 + b * y) * (1 + c * z);
         o.dSp[(1 * 8 + i) * G + k] = b * 0.125 * (1 + a * x) * (1 + c * z);
         o.dSp[(2 * 8 + i) * G + k] = c * 0.125 * (1 + a * x) * (1 + b * y);
      }
   }
   const int nv = o.NENv;
   if (nv == 8) {   // BCH.cpp:2303-2305 (commented `Sv = Sp; dSv = dSp;`) -- the evident intent
      o.Sv = o.Sp; o.dSv = o.dSp;
      return;
   }
   static const int pos[27][3] = {
       {-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1},
       {0, -1, -1},  {1, 0, -1},  {0, 1, -1}, {-1, 0, -1}, {-1, -1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 1, 0},
       {0, -1, 1},   {1, 0, 1},   {0, 1, 1},  {-1, 0, 1},  {0, 0, -1},  {0, -1, 0}, {1, 0, 0}, {0, 1, 0},
       {-1, 0, 0},   {0, 0, 1},   {0, 0, 0}};
   auto f = [](int p, double t) { return p < 0 ? (t * t - t) : (p > 0 ? (t * t + t) : (1 - t * t)); };
   auto df = [](int p, double t) { return p < 0 ? (2 * t - 1) : (p > 0 ? (2 * t + 1) : (-2 * t)); };
   o.Sv.assign(G * 27, 0); o.dSv.assign(3 * 27 * G, 0);
   for (int k = 0; k < G; k++) {
      double x = o.GQpoint[3 * k], y = o.GQpoint[3 * k + 1], z = o.GQpoint[3 * k + 2];
      for (int i = 0; i < 27; i++) {
         int nz = (pos[i][0] == 0) + (pos[i][1] == 0) + (pos[i][2] == 0);
         double coef = nz == 0 ? 0.125 : nz == 1 ? 0.25 : nz == 2 ? 0.5 : 1.0;
         o.Sv[k * 27 + i] = coef * f(pos[i][0], x) * f(pos[i][1], y) * f(pos[i][2], z);
         o.dSv[(0 * 27 + i) * G + k] = coef * df(pos[i][0], x) * f(pos[i][1], y) * f(pos[i][2], z);
         o.dSv[(1 * 27 + i) * G + k] = coef * f(pos[i][0], x) * df(pos[i][1], y) * f(pos[i][2], z);
         o.dSv[(2 * 27 + i) * G + k] = coef * f(pos[i][0], x) * f(pos[i][1], y) * df(pos[i][2], z);
      }
   }
}

// calcJacob (BCH.cpp:2495-2711): sub-parametric Jacobian from the 8 corners.
void calcJacob(Orc& o) {
   const int G = o.NGP, nv = o.NENv, np = o.NENp;
   o.detJacob.resize((size_t)o.NE * G);
   o.gDSp.resize((size_t)o.NE * G * np * 3);
   o.gDSv.resize((size_t)o.NE * G * nv * 3);
#pragma omp parallel for schedule(static)
   for (int e = 0; e < o.NE; e++) {
      double ec[8][3], J[3][3], inv[3][3];
      for (int i = 0; i < 8; i++)
         for (int d = 0; d < 3; d++) ec[i][d] = o.coord[(size_t)3 * o.LtoGnode[(size_t)e * nv + i] + d];
      for (int k = 0; k < G; k++) {
         for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
               double s = 0;
               for (int m = 0; m < np; m++) s = s + o.dSp[(i * np + m) * G + k] * ec[m][j];
               J[i][j] = s;
            }
         inv[0][0] = J[1][1] * J[2][2] - J[2][1] * J[1][2];
         inv[0][1] = -(J[0][1] * J[2][2] - J[0][2] * J[2][1]);
         inv[0][2] = J[0][1] * J[1][2] - J[1][1] * J[0][2];
         inv[1][0] = -(J[1][0] * J[2][2] - J[1][2] * J[2][0]);
         inv[1][1] = J[2][2] * J[0][0] - J[2][0] * J[0][2];
         inv[1][2] = -(J[1][2] * J[0][0] - J[1][0] * J[0][2]);
         inv[2][0] = J[1][0] * J[2][1] - J[2][0] * J[1][1];
         inv[2][1] = -(J[2][1] * J[0][0] - J[2][0] * J[0][1]);
         inv[2][2] = J[1][1] * J[0][0] - J[1][0] * J[0][1];
         double det = J[0][0] * (J[1][1] * J[2][2] - J[2][1] * J[1][2]) + J[0][1] * (J[1][2] * J[2][0] - J[1][0] * J[2][2]) +
                      J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
         o.detJacob[(size_t)e * G + k] = det;
         for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) inv[i][j] = inv[i][j] / det;
         for (int i = 0; i < 3; i++)
            for (int j = 0; j < np; j++) {
               double s = 0;
               for (int m = 0; m < 3; m++) s = s + inv[i][m] * o.dSp[(m * np + j) * G + k];
               o.gDSp[(((size_t)e * G + k) * np + j) * 3 + i] = s;
            }
         for (int i = 0; i < 3; i++)
            for (int j = 0; j < nv; j++) {
               double s = 0;
               for (int m = 0; m < 3; m++) s = s + inv[i][m] * o.dSv[(m * nv + j) * G + k];
               o.gDSv[(((size_t)e * G + k) * nv + j) * 3 + i] = s;
            }
      }
   }
   o.GQfactor_1d.resize((size_t)o.NE * G);
   for (size_t i = 0; i < o.GQfactor_1d.size(); i++) o.GQfactor_1d[i] = o.detJacob[i] * o.GQweight[i % G];
   o.gDSv_1d = o.gDSv;   // already [e][k][j][d] (BCH.cpp:2658-2676)
   o.Sv_1d = o.Sv;       // [k][i]       (BCH.cpp:2460-2469)
}

// y = alpha*A*x + beta*y for a CSR pattern + value array (stands in for mkl_dcsrmv 'n').
void spmv(const Csr& P, const double* val, const double* x, double* y, double alpha, double beta) {
#pragma omp parallel for schedule(static)
   for (int r = 0; r < P.nrows; r++) {
      double s = 0;
      for (int k = P.rowStarts[r]; k < P.rowStarts[r + 1]; k++) s += val[k] * x[P.col[k]];
      y[r] = (beta == 0.0 ? 0.0 : beta * y[r]) + alpha * s;
   }
}
// y = alpha*A^T*x + beta*y (stands in for mkl_dcsrmv 't'): serial scatter.
void spmvT(const Csr& P, const double* val, const double* x, double* y, double alpha, double beta) {
   for (int c = 0; c < P.ncols; c++) y[c] = (beta == 0.0 ? 0.0 : beta * y[c]);
   for (int r = 0; r < P.nrows; r++) {
      double xr = alpha * x[r];
      for (int k = P.rowStarts[r]; k < P.rowStarts[r + 1]; k++) y[P.col[k]] += val[k] * xr;
   }
}

// ---------------------------------------------------------------------------
// Boundary conditions
// ---------------------------------------------------------------------------
// explicit: applyBC_initial (BCH.cpp:4062-4110); implicit: GQ.cpp:4559-4607 (Uk_prev)
void applyBC_initial(Orc& o, dvec& U) {
   for (int i = 0; i < o.BCnVelNodes; i++) {
      int node = o.BCvelNodes[2 * i], w = o.BCvelNodes[2 * i + 1];
      for (int c = 0; c < 3; c++) U[node + c * o.NN] = o.BCstr[3 * w + c];
   }
}
// explicit applyBC_Step1(2) / applyBC_Step3 (BCH.cpp:4117-4160, 4194-4207): zero increments
void zeroVelBC(Orc& o, dvec& R) {
   for (int i = 0; i < o.BCnVelNodes; i++) {
      int node = o.BCvelNodes[2 * i];
      R[node] = 0; R[node + o.NN] = 0; R[node + 2 * o.NN] = 0;
   }
}
// applyBC_Step2(1) (BCH.cpp:4167-4180, GQ.cpp:4650-4680): diag*1000 at the pin iff node > 0
void pinPressureLHS(Orc& o) {
   int node = o.zeroPressureNode;
   if (node > 0)
      for (int j = o.Zp.rowStarts[node]; j < o.Zp.rowStarts[node + 1]; j++)
         if (o.Zp.col[j] == node) { o.Zval[j] = o.Zval[j] * 1000; break; }
}

// ---------------------------------------------------------------------------
// step0: element matrices M, K, G1..3 (BCH.cpp:3127-3315; GQ.cpp:3425-3572 where
// M carries 1/dt).
// ---------------------------------------------------------------------------
void assembleMKG(Orc& o) {
   const int nv = o.NENv, np = o.NENp, G = o.NGP;
   const int nnzM = o.Mp.nnz(), nnzG = o.Gp.nnz();
   o.Mval.assign(nnzM, 0); o.Kval.assign(nnzM, 0);
   o.G1.assign(nnzG, 0); o.G2.assign(nnzG, 0); o.G3.assign(nnzG, 0);
   const double invDens = 1.0 / o.density, invDt = 1.0 / o.dt;
   dvec Me(nv * nv), Ke(nv * nv), Ge1(nv * np), Ge2(nv * np), Ge3(nv * np);
   for (int e = 0; e < o.NE; e++) {
      std::fill(Me.begin(), Me.end(), 0); std::fill(Ke.begin(), Ke.end(), 0);
      std::fill(Ge1.begin(), Ge1.end(), 0); std::fill(Ge2.begin(), Ge2.end(), 0); std::fill(Ge3.begin(), Ge3.end(), 0);
      for (int k = 0; k < G; k++) {
         double fac = o.detJacob[(size_t)e * G + k] * o.GQweight[k];
         const double* gd = &o.gDSv[((size_t)e * G + k) * nv * 3];
         const double* sv = &o.Sv[k * nv];
         for (int i = 0; i < nv; i++)
            for (int j = 0; j < nv; j++) {
               if (o.scheme == 0) Me[i * nv + j] = Me[i * nv + j] + sv[i] * sv[j] * fac;
               else Me[i * nv + j] = Me[i * nv + j] + invDt * sv[i] * sv[j] * fac;
               Ke[i * nv + j] = Ke[i * nv + j] + o.viscosity * (gd[i * 3] * gd[j * 3] + gd[i * 3 + 1] * gd[j * 3 + 1] + gd[i * 3 + 2] * gd[j * 3 + 2]) * fac;
            }
         for (int i = 0; i < nv; i++)
            for (int j = 0; j < np; j++) {
               double sp = o.Sp[k * np + j];
               Ge1[i * np + j] = Ge1[i * np + j] - invDens * sp * gd[i * 3] * fac;
               Ge2[i * np + j] = Ge2[i * np + j] - invDens * sp * gd[i * 3 + 1] * fac;
               Ge3[i * np + j] = Ge3[i * np + j] - invDens * sp * gd[i * 3 + 2] * fac;
            }
      }
      const int* mm = &o.sparseMapM[(size_t)e * nv * nv];
      for (int q = 0; q < nv * nv; q++) { o.Mval[mm[q]] += Me[q]; o.Kval[mm[q]] += Ke[q]; }
      const int* mg = &o.sparseMapG[(size_t)e * nv * np];
      for (int q = 0; q < nv * np; q++) { o.G1[mg[q]] += Ge1[q]; o.G2[mg[q]] += Ge2[q]; o.G3[mg[q]] += Ge3[q]; }
   }
}

// explicit Z = sum_c G_c^T diag(MdOrigInv) G_c  (calculateZ, BCH.cpp:3322-3479, the
// CSparse transpose/multiply/add sequence restated as a row-wise sparse product).
void calculateZ_explicit(Orc& o) {
   const Csr& Gp = o.Gp;
   const int NN = o.NN, NNp = o.NNp, nnzG = Gp.nnz();
   // transpose pattern (rows = pressure nodes) with positions into the G value arrays
   ivec tStart(NNp + 1, 0), tRow(nnzG), tPos(nnzG);
   for (int k = 0; k < nnzG; k++) tStart[Gp.col[k] + 1]++;
   for (int i = 0; i < NNp; i++) tStart[i + 1] += tStart[i];
   {
      ivec fill(tStart.begin(), tStart.end() - 1);
      for (int r = 0; r < NN; r++)
         for (int k = Gp.rowStarts[r]; k < Gp.rowStarts[r + 1]; k++) {
            int p = fill[Gp.col[k]]++;
            tRow[p] = r; tPos[p] = k;
         }
   }
   Csr& Z = o.Zp;
   Z.nrows = Z.ncols = NNp;
   Z.rowStarts.assign(NNp + 1, 0);
   std::vector<ivec> zc(NNp);
   std::vector<dvec> zv(NNp);
#pragma omp parallel
   {
      dvec acc(NNp, 0.0);
      ivec mark(NNp, -1), list;
#pragma omp for schedule(dynamic, 64)
      for (int i = 0; i < NNp; i++) {
         list.clear();
         for (int comp = 0; comp < 3; comp++) {
            const dvec& Gv = comp == 0 ? o.G1 : comp == 1 ? o.G2 : o.G3;
            for (int t = tStart[i]; t < tStart[i + 1]; t++) {
               int r = tRow[t];
               double w = Gv[tPos[t]];
               double d = o.MdOrigInv[r];
               for (int k = Gp.rowStarts[r]; k < Gp.rowStarts[r + 1]; k++) {
                  int j = Gp.col[k];
                  if (mark[j] != i) { mark[j] = i; acc[j] = 0; list.push_back(j); }
                  acc[j] += w * (Gv[k] * d);
               }
            }
         }
         std::sort(list.begin(), list.end());
         zc[i] = list;
         zv[i].resize(list.size());
         for (size_t q = 0; q < list.size(); q++) zv[i][q] = acc[list[q]];
      }
   }
   for (int i = 0; i < NNp; i++) Z.rowStarts[i + 1] = Z.rowStarts[i] + (int)zc[i].size();
   Z.col.resize(Z.rowStarts[NNp]); Z.row.resize(Z.rowStarts[NNp]);
   o.Zval.resize(Z.rowStarts[NNp]);
   for (int i = 0; i < NNp; i++) {
      std::copy(zc[i].begin(), zc[i].end(), Z.col.begin() + Z.rowStarts[i]);
      std::copy(zv[i].begin(), zv[i].end(), o.Zval.begin() + Z.rowStarts[i]);
      std::fill(Z.row.begin() + Z.rowStarts[i], Z.row.begin() + Z.rowStarts[i + 1], i);
   }
}

// implicit Z = -int grad(Sp_i).grad(Sp_j)  (calculateZGPU, GQ.cpp:3677-3772; full matrix)
void calculateZ_implicit(Orc& o) {
   const int np = o.NENp, G = o.NGP;
   o.Zval.assign(o.Zp.nnz(), 0);
   dvec Ze(np * np);
   for (int e = 0; e < o.NE; e++) {
      std::fill(Ze.begin(), Ze.end(), 0);
      for (int k = 0; k < G; k++) {
         double fac = o.detJacob[(size_t)e * G + k] * o.GQweight[k];
         const double* gd = &o.gDSp[((size_t)e * G + k) * np * 3];
         for (int i = 0; i < np; i++)
            for (int j = 0; j < np; j++)
               Ze[i * np + j] = Ze[i * np + j] - (gd[i * 3] * gd[j * 3] + gd[i * 3 + 1] * gd[j * 3 + 1] + gd[i * 3 + 2] * gd[j * 3 + 2]) * fac;
      }
      const int* mz = &o.sparseMapZ[(size_t)e * np * np];
      for (int q = 0; q < np * np; q++) o.Zval[mz[q]] += Ze[q];
   }
}

void step0(Orc& o) {
   assembleMKG(o);
   const int NN = o.NN;
   if (o.scheme == 0) {
      // lumped mass (BCH.cpp:3254-3292)
      o.Md.assign(3 * NN, 0);
      for (int k = 0; k < o.Mp.nnz(); k++) o.Md[o.Mp.row[k]] += o.Mval[k];
      for (int i = 0; i < NN; i++) { o.Md[i + NN] = o.Md[i]; o.Md[i + 2 * NN] = o.Md[i]; }
      o.MdOrig = o.Md;
      o.MdOrigInv.resize(3 * NN);
      for (int i = 0; i < 3 * NN; i++) o.MdOrigInv[i] = 1.0 / o.MdOrig[i];
      for (int i = 0; i < o.BCnVelNodes; i++) {   // applyBC_Step1(1), BCH.cpp:4121-4128
         int node = o.BCvelNodes[2 * i];
         o.Md[node] = 1.0; o.Md[node + NN] = 1.0; o.Md[node + 2 * NN] = 1.0;
      }
      o.MdInv.resize(3 * NN);
      for (int i = 0; i < 3 * NN; i++) o.MdInv[i] = 1.0 / o.Md[i];
      calculateZ_explicit(o);
      pinPressureLHS(o);
   } else {
      calculateZ_implicit(o);
      pinPressureLHS(o);
   }
}

// ---------------------------------------------------------------------------
// Krylov solvers
// ---------------------------------------------------------------------------
double dot(const double* a, const double* b, int n) {
   double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
   for (int i = 0; i < n; i++) s += a[i] * b[i];
   return s;
}

// Jacobi-PCG on Z (MKL_CG_solver BCH.cpp:3981-4055 / GQ.cpp:4478-4552;
// CUSP_CG_solver BCH.cu:418-451 / GQ.cu:302-362).  Monitor rule of
// cusp::default_monitor(b, maxit, tol): finished when ||r|| <= tol*||b|| or the
// iteration limit is hit, tested before each iteration.
int pcg(Orc& o, const double* b, double* x) {
   const int n = o.NNp;
   const Csr& Z = o.Zp;
   dvec r(n), z(n), p(n), q(n), dinv(n, 1.0);
   if (o.opt.cg_precond)
      for (int i = 0; i < n; i++)
         for (int k = Z.rowStarts[i]; k < Z.rowStarts[i + 1]; k++)
            if (Z.col[k] == i) dinv[i] = 1.0 / o.Zval[k];
   spmv(Z, o.Zval.data(), x, q.data(), 1.0, 0.0);
   for (int i = 0; i < n; i++) { r[i] = b[i] - q[i]; z[i] = dinv[i] * r[i]; p[i] = z[i]; }
   double rz = dot(r.data(), z.data(), n);
   const double bnorm = std::sqrt(dot(b, b, n));
   int it = 0;
   while (true) {
      double rnorm = std::sqrt(dot(r.data(), r.data(), n));
      if (rnorm <= o.opt.cg_tol * bnorm || it >= o.opt.cg_maxit) break;
      spmv(Z, o.Zval.data(), p.data(), q.data(), 1.0, 0.0);
      double alpha = rz / dot(q.data(), p.data(), n);
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; i++) { x[i] += alpha * p[i]; r[i] -= alpha * q[i]; z[i] = dinv[i] * r[i]; }
      double rzNew = dot(r.data(), z.data(), n);
      double beta = rzNew / rz;
      rz = rzNew;
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; i++) p[i] = z[i] + beta * p[i];
      it++;
   }
   return it;
}

// Jacobi-BiCGStab (paralution_BiCGStab_solver GQ.cpp:4183-4231; cusp::krylov::bicgstab
// via CUSP_GMRES_solver GQ.cu:369-433).  x0 = 0, right preconditioning M = diag(A)^-1.
int bicgstab(Orc& o, const double* A, const double* b, double* x) {
   const int n = o.NN;
   const Csr& P = o.Mp;
   dvec r(n), rs(n), p(n), Mp_(n), AMp(n), s(n), Ms(n), AMs(n), dinv(n, 1.0);
   for (int i = 0; i < n; i++)
      for (int k = P.rowStarts[i]; k < P.rowStarts[i + 1]; k++)
         if (P.col[k] == i) dinv[i] = 1.0 / A[k];
#pragma omp parallel for schedule(static)
   for (int i = 0; i < n; i++) { x[i] = 0; r[i] = b[i]; p[i] = r[i]; rs[i] = r[i]; }
   double rrsOld = dot(rs.data(), r.data(), n);
   const double bnorm = std::sqrt(dot(b, b, n));
   int it = 0;
   while (true) {
      double rnorm = std::sqrt(dot(r.data(), r.data(), n));
      if (rnorm <= o.opt.bicg_tol * bnorm || it >= o.opt.bicg_maxit) break;
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; i++) Mp_[i] = dinv[i] * p[i];
      spmv(P, A, Mp_.data(), AMp.data(), 1.0, 0.0);
      double alpha = rrsOld / dot(rs.data(), AMp.data(), n);
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; i++) s[i] = r[i] - alpha * AMp[i];
      double snorm = std::sqrt(dot(s.data(), s.data(), n));
      if (o.opt.bicg_halfstep && snorm <= o.opt.bicg_tol * bnorm) {
         // cusp::krylov::bicgstab leaves here BEFORE `++monitor`: the half step is not counted as an iteration
         for (int i = 0; i < n; i++) x[i] += alpha * Mp_[i];
         break;
      }
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; i++) Ms[i] = dinv[i] * s[i];
      spmv(P, A, Ms.data(), AMs.data(), 1.0, 0.0);
      double omega = dot(AMs.data(), s.data(), n) / dot(AMs.data(), AMs.data(), n);
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; i++) {
         x[i] += alpha * Mp_[i] + omega * Ms[i];
         r[i] = s[i] - omega * AMs[i];
      }
      double rrsNew = dot(rs.data(), r.data(), n);
      double beta = (rrsNew / rrsOld) * (alpha / omega);
      rrsOld = rrsNew;
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; i++) p[i] = r[i] + beta * (p[i] - omega * AMp[i]);
      it++;
   }
   return it;
}

// ---------------------------------------------------------------------------
// explicit scheme (BCH)
// ---------------------------------------------------------------------------
// calculateMatrixA (BCH.cpp:3520-3685): R1 = -C(Un) * UnpHalf_prev, element by
// element, forming the 27x27 Ae exactly as the reference does; coloured OpenMP loop.
void bch_calculateMatrixA(Orc& o) {
   const int nv = o.NENv, G = o.NGP, NN = o.NN;
   std::fill(o.R1.begin(), o.R1.end(), 0.0);
   int offset = 0;
   for (int color = 0; color < o.nActiveColors; color++) {
      // colours are stored compacted in elementsOfColor; NmeshColors is indexed by colour id
      int cid = -1, seen = -1;
      for (int c = 0; c < (int)o.NmeshColors.size(); c++)
         if (o.NmeshColors[c] > 0 && ++seen == color) { cid = c; break; }
      const int cnt = o.NmeshColors[cid];
#pragma omp parallel if (o.colorConflicts == 0)
      {
         dvec Ae(nv * nv), u0n(nv), v0n(nv), w0n(nv), up(nv), vp(nv), wp(nv);
#pragma omp for schedule(static)
         for (int q = 0; q < cnt; q++) {
            int e = o.elementsOfColor[offset + q];
            std::fill(Ae.begin(), Ae.end(), 0.0);
            const int* lg = &o.LtoGvel_1d[(size_t)e * 3 * nv];
            for (int i = 0; i < nv; i++) {
               u0n[i] = o.Un[lg[i]]; up[i] = o.UnpHalf_prev[lg[i]];
               v0n[i] = o.Un[lg[i + nv]]; vp[i] = o.UnpHalf_prev[lg[i + nv]];
               w0n[i] = o.Un[lg[i + 2 * nv]]; wp[i] = o.UnpHalf_prev[lg[i + 2 * nv]];
            }
            for (int k = 0; k < G; k++) {
               double fac = o.detJacob[(size_t)e * G + k] * o.GQweight[k];
               const double* sv = &o.Sv[k * nv];
               const double* gd = &o.gDSv[((size_t)e * G + k) * nv * 3];
               double u0 = 0, v0 = 0, w0 = 0;
               for (int i = 0; i < nv; i++) { u0 = u0 + sv[i] * u0n[i]; v0 = v0 + sv[i] * v0n[i]; w0 = w0 + sv[i] * w0n[i]; }
               for (int i = 0; i < nv; i++)
                  for (int j = 0; j < nv; j++)
                     Ae[i * nv + j] = Ae[i * nv + j] + (u0 * gd[j * 3] + v0 * gd[j * 3 + 1] + w0 * gd[j * 3 + 2]) * sv[i] * fac;
            }
            for (int i = 0; i < nv; i++) {
               double ru = 0, rv = 0, rw = 0;
               for (int j = 0; j < nv; j++) { ru += Ae[i * nv + j] * up[j]; rv += Ae[i * nv + j] * vp[j]; rw += Ae[i * nv + j] * wp[j]; }
               int g = lg[i];
               o.R1[g] -= ru; o.R1[g + NN] -= rv; o.R1[g + 2 * NN] -= rw;
            }
         }
      }
      offset += cnt;
   }
}

// step1 (BCH.cpp:3692-3801)
void bch_step1(Orc& o) {
   const int NN = o.NN;
   bch_calculateMatrixA(o);
   for (int c = 0; c < 3; c++) spmv(o.Mp, o.Kval.data(), &o.UnpHalf_prev[c * NN], &o.R1[c * NN], -1.0, 1.0);
   spmv(o.Gp, o.G1.data(), o.Pn.data(), &o.R1[0], -1.0, 1.0);
   spmv(o.Gp, o.G2.data(), o.Pn.data(), &o.R1[NN], -1.0, 1.0);
   spmv(o.Gp, o.G3.data(), o.Pn.data(), &o.R1[2 * NN], -1.0, 1.0);
   zeroVelBC(o, o.R1);
   for (int i = 0; i < 3 * NN; i++) o.UnpHalf[i] = o.Un[i] + o.dt * o.R1[i] * o.MdInv[i];
}

// step2 (BCH.cpp:3808-3905)
void bch_step2(Orc& o, int iter) {
   const int NN = o.NN, NNp = o.NNp;
   dvec dummy(3 * NN);
   const double oneOverdt2 = 1.0 / (o.dt * o.dt);
   for (int i = 0; i < 3 * NN; i++) dummy[i] = oneOverdt2 * o.UnpHalf[i];
   if (iter != 1)
      for (int i = 0; i < 3 * NN; i++) dummy[i] = dummy[i] - o.MdOrigInv[i] * o.KtimesAcc_prev[i];
   spmvT(o.Gp, o.G1.data(), &dummy[0], o.R2.data(), 1.0, 0.0);
   spmvT(o.Gp, o.G2.data(), &dummy[NN], o.R2.data(), 1.0, 1.0);
   spmvT(o.Gp, o.G3.data(), &dummy[2 * NN], o.R2.data(), 1.0, 1.0);
   if (o.zeroPressureNode > 0) o.R2[o.zeroPressureNode] = 0.0;   // applyBC_Step2(2)
   if (!o.opt.cg_warm) std::fill(o.Pdot.begin(), o.Pdot.end(), 0.0);   // BCH.cpp:3997-3999
   double t0 = wall();
   o.cgIters.push_back(pcg(o, o.R2.data(), o.Pdot.data()));
   o.cgSeconds += wall() - t0;
   for (int i = 0; i < NNp; i++) o.Pnp1[i] = o.Pn[i] + o.dt * o.Pdot[i];
}

// step3 (BCH.cpp:3912-3974)
void bch_step3(Orc& o, int iter) {
   const int NN = o.NN;
   spmv(o.Gp, o.G1.data(), o.Pdot.data(), &o.R3[0], -o.dt, 0.0);
   spmv(o.Gp, o.G2.data(), o.Pdot.data(), &o.R3[NN], -o.dt, 0.0);
   spmv(o.Gp, o.G3.data(), o.Pdot.data(), &o.R3[2 * NN], -o.dt, 0.0);
   if (iter != 1)
      for (int i = 0; i < 3 * NN; i++) o.R3[i] = o.R3[i] - o.dt * o.KtimesAcc_prev[i];
   zeroVelBC(o, o.R3);
   for (int i = 0; i < 3 * NN; i++) o.Acc[i] = o.R3[i] * o.MdInv[i];
   for (int i = 0; i < 3 * NN; i++) o.Unp1[i] = o.UnpHalf[i] + o.dt * o.Acc[i];
}

// one pass of the `while (timeT < t_final)` body (BCH.cpp:2859-3118)
void bch_timeStep(Orc& o) {
   const int NN = o.NN, NNp = o.NNp;
   o.timeN++; o.timeT += o.dt;
   for (int i = 0; i < 3 * NN; i++) { o.UnpHalf_prev[i] = o.Un[i]; o.Acc_prev[i] = 0.0; }
   for (int i = 0; i < NNp; i++) o.Pnp1_prev[i] = o.Pn[i];
   int iter;
   for (iter = 1; iter <= o.maxIter; iter++) {
      bch_step1(o);
      bch_step2(o, iter);
      bch_step3(o, iter);
      double s1 = 0, s2 = 0;
      for (int i = 0; i < 3 * NN; i++) { s1 += o.Unp1[i] * o.Unp1[i]; double d = o.Unp1[i] - o.Unp1_prev[i]; s2 += d * d; }
      double n1 = std::sqrt(s2) / std::sqrt(s1);
      s1 = 0; s2 = 0;
      for (int i = 0; i < NNp; i++) { s1 += o.Pnp1[i] * o.Pnp1[i]; double d = o.Pnp1[i] - o.Pnp1_prev[i]; s2 += d * d; }
      double n2 = std::sqrt(s2) / std::sqrt(s1);
      if (n1 < o.tolerance && n2 < o.tolerance) break;
      o.UnpHalf_prev = o.UnpHalf; o.Unp1_prev = o.Unp1; o.Acc_prev = o.Acc; o.Pnp1_prev = o.Pnp1;
      for (int c = 0; c < 3; c++) spmv(o.Mp, o.Kval.data(), &o.Acc_prev[c * NN], &o.KtimesAcc_prev[c * NN], 1.0, 0.0);
   }
   o.lastIter = iter;
   o.innerIters.push_back(iter);
   const double oneOverdt = 1.0 / o.dt;
   o.maxAcc = 0;
   for (int i = 0; i < 3 * NN; i++) {
      double d = (o.Unp1[i] - o.Un[i]);
      if (o.opt.maxacc_rule == 0) d *= oneOverdt;
      if (std::fabs(d) > o.maxAcc) o.maxAcc = std::fabs(d);
   }
   o.Un = o.Unp1; o.Pn = o.Pnp1;
}

// ---------------------------------------------------------------------------
// implicit scheme (GQ)
// ---------------------------------------------------------------------------
// calculateMatrixA (GQ.cpp:3779-3899): A += C(Uk_prev) through sparseMapM (the
// `0.0 * (Du0+Dv0+Dw0)` term at :3857 contributes nothing and is dropped).
void gq_calculateMatrixA(Orc& o) {
   const int nv = o.NENv, G = o.NGP;
   int offset = 0;
   for (int color = 0; color < o.nActiveColors; color++) {
      int cid = -1, seen = -1;
      for (int c = 0; c < (int)o.NmeshColors.size(); c++)
         if (o.NmeshColors[c] > 0 && ++seen == color) { cid = c; break; }
      const int cnt = o.NmeshColors[cid];
#pragma omp parallel if (o.colorConflicts == 0)
      {
         dvec Ae(nv * nv), u0n(nv), v0n(nv), w0n(nv);
#pragma omp for schedule(static)
         for (int q = 0; q < cnt; q++) {
            int e = o.elementsOfColor[offset + q];
            std::fill(Ae.begin(), Ae.end(), 0.0);
            const int* lg = &o.LtoGvel_1d[(size_t)e * 3 * nv];
            for (int i = 0; i < nv; i++) { u0n[i] = o.Uk_prev[lg[i]]; v0n[i] = o.Uk_prev[lg[i + nv]]; w0n[i] = o.Uk_prev[lg[i + 2 * nv]]; }
            for (int k = 0; k < G; k++) {
               double fac = o.detJacob[(size_t)e * G + k] * o.GQweight[k];
               const double* sv = &o.Sv[k * nv];
               const double* gd = &o.gDSv[((size_t)e * G + k) * nv * 3];
               double u0 = 0, v0 = 0, w0 = 0;
               for (int i = 0; i < nv; i++) { u0 = u0 + sv[i] * u0n[i]; v0 = v0 + sv[i] * v0n[i]; w0 = w0 + sv[i] * w0n[i]; }
               for (int i = 0; i < nv; i++)
                  for (int j = 0; j < nv; j++)
                     Ae[i * nv + j] = Ae[i * nv + j] + (u0 * gd[j * 3] + v0 * gd[j * 3 + 1] + w0 * gd[j * 3 + 2]) * sv[i] * fac;
            }
            const int* mm = &o.sparseMapM[(size_t)e * nv * nv];
            for (int qq = 0; qq < nv * nv; qq++) o.Aval[mm[qq]] += Ae[qq];
         }
      }
      offset += cnt;
   }
}

// step1 (GQ.cpp:3906-4083) with applyBC_Step1 (GQ.cpp:4614-4643)
void gq_step1(Orc& o) {
   const int NN = o.NN, NNp = o.NNp, nnzM = o.Mp.nnz();
   o.Aval.resize(nnzM);
   for (int i = 0; i < nnzM; i++) o.Aval[i] = o.Mval[i] + o.Kval[i];
   gq_calculateMatrixA(o);
   for (int i = 0; i < o.BCnVelNodes; i++) {
      int node = o.BCvelNodes[2 * i];
      for (int j = o.Mp.rowStarts[node]; j < o.Mp.rowStarts[node + 1]; j++) o.Aval[j] = (o.Mp.col[j] == node) ? 1.0 : 0.0;
   }
   dvec ptil(NNp), R1(NN), x(NN);
   for (int i = 0; i < NNp; i++) ptil[i] = 2 * o.Pk_prev[i] - o.Pk_prevprev[i];
   for (int d = 0; d < 3; d++) {
      o.direction = d;
      spmv(o.Mp, o.Mval.data(), &o.Uk_prev[d * NN], R1.data(), 1.0, 0.0);
      const dvec& Gv = d == 0 ? o.G1 : d == 1 ? o.G2 : o.G3;
      spmv(o.Gp, Gv.data(), ptil.data(), R1.data(), -1.0, 1.0);
      for (int i = 0; i < o.BCnVelNodes; i++) R1[o.BCvelNodes[2 * i]] = o.BCstr[3 * o.BCvelNodes[2 * i + 1] + d];
      o.bicgIters.push_back(bicgstab(o, o.Aval.data(), R1.data(), x.data()));
      for (int i = 0; i < NN; i++) o.Uk[i + d * NN] = x[i];
   }
}

// step2 (GQ.cpp:4090-4176)
void gq_step2(Orc& o) {
   const int NN = o.NN, NNp = o.NNp;
   const double a = -1.0 / o.dt;
   spmvT(o.Gp, o.G1.data(), &o.Uk[0], o.R2.data(), a, 0.0);
   spmvT(o.Gp, o.G2.data(), &o.Uk[NN], o.R2.data(), a, 1.0);
   spmvT(o.Gp, o.G3.data(), &o.Uk[2 * NN], o.R2.data(), a, 1.0);
   if (o.zeroPressureNode > 0) o.R2[o.zeroPressureNode] = 0.0;
   std::fill(o.Pdiff.begin(), o.Pdiff.end(), 0.0);   // GQ.cpp:4494-4496, GQ.cu:307
   double t0 = wall();
   o.cgIters.push_back(pcg(o, o.R2.data(), o.Pdiff.data()));
   o.cgSeconds += wall() - t0;
   for (int i = 0; i < NNp; i++) o.Pk[i] = o.Pk_prev[i] + o.Pdiff[i];
}

// one pass of the time loop body (GQ.cpp:3308-3416)
void gq_timeStep(Orc& o) {
   const int NN = o.NN;
   o.timeN++; o.timeT += o.dt;
   gq_step1(o);
   gq_step2(o);
   const double oneOverdt = 1.0 / o.dt;
   o.maxAcc = 0;
   for (int i = 0; i < 3 * NN; i++) {
      double d = (o.Uk[i] - o.Uk_prev[i]);
      if (o.opt.maxacc_rule == 0) d *= oneOverdt;
      if (std::fabs(d) > o.maxAcc) o.maxAcc = std::fabs(d);   // GQ.cpp:3349-3352 loses the abs; intent kept
   }
   o.Uk_prev = o.Uk; o.Pk_prevprev = o.Pk_prev; o.Pk_prev = o.Pk;
}

void allocateState(Orc& o) {
   const int NN = o.NN, NNp = o.NNp;
   if (o.scheme == 0) {
      for (dvec* v : {&o.Un, &o.Unp1, &o.Unp1_prev, &o.UnpHalf, &o.UnpHalf_prev, &o.Acc, &o.Acc_prev, &o.KtimesAcc_prev, &o.R1, &o.R3})
         v->assign(3 * NN, 0.0);
      for (dvec* v : {&o.Pn, &o.Pnp1, &o.Pnp1_prev, &o.Pdot, &o.R2}) v->assign(NNp, 0.0);
      applyBC_initial(o, o.Un);
   } else {
      o.Uk.assign(3 * NN, 0.0); o.Uk_prev.assign(3 * NN, 0.0);
      for (dvec* v : {&o.Pk, &o.Pk_prev, &o.Pk_prevprev, &o.Pdiff, &o.R2}) v->assign(NNp, 0.0);
      applyBC_initial(o, o.Uk_prev);
   }
   o.timeN = 0; o.timeT = o.t_ini;
}

// readRestartFile (BCH.cpp:4214-4242): velocities of all NN rows, pressure of the first NCN.
bool readRestart(Orc& o, const std::string& path) {
   std::ifstream f(path.c_str());
   if (!f) { o.err = "cannot open " + path; return false; }
   f.ignore(1 << 20, '\n'); f.ignore(1 << 20, '\n'); f.ignore(1 << 20, '\n');
   dvec& U = o.scheme == 0 ? o.Un : o.Uk_prev;
   dvec& P = o.scheme == 0 ? o.Pn : o.Pk_prev;
   double x, y, z, p;
   for (int i = 0; i < o.NN; i++) {
      f >> x >> y >> z >> U[i] >> U[o.NN + i] >> U[2 * o.NN + i] >> p;
      if (i < o.NCN) P[i] = p;
      f.ignore(1 << 20, '\n');
   }
   return (bool)f;
}

// createTecplot (BCH.cpp:4249-4482)
bool createTecplot(Orc& o, const std::string& path, const std::string& title) {
   std::ofstream f(path.c_str());
   if (!f) { o.err = "cannot write " + path; return false; }
   const int NN = o.NN, nv = o.NENv;
   const dvec& U = o.scheme == 0 ? o.Un : o.Uk;
   const dvec& P = o.scheme == 0 ? o.Pn : o.Pk;
   f << "TITLE = " << title << std::endl;
   f << "VARIABLES = x,  y,  z,  u, v, w, p" << std::endl;
   f << "ZONE N=" << NN << ", E=" << (nv == 27 ? 8 * o.NE : o.NE) << ", F=FEPOINT, ET=BRICK" << std::endl;
   dvec pNode(NN, 0.0);
   for (int i = 0; i < o.NNp; i++) pNode[i] = P[i];
   if (nv == 27)
      for (int e = 0; e < o.NE; e++) {
         const int* L = &o.LtoGnode[(size_t)e * nv];
         for (int ed = 0; ed < 12; ed++) pNode[L[8 + ed]] = 0.5 * (pNode[L[kEdge[ed][0]]] + pNode[L[kEdge[ed][1]]]);
         for (int fc = 0; fc < 6; fc++)
            pNode[L[20 + fc]] = 0.25 * (pNode[L[kFace[fc][0]]] + pNode[L[kFace[fc][1]]] + pNode[L[kFace[fc][2]]] + pNode[L[kFace[fc][3]]]);
         pNode[L[26]] = 0.125 * (pNode[L[0]] + pNode[L[1]] + pNode[L[2]] + pNode[L[3]] + pNode[L[4]] + pNode[L[5]] + pNode[L[6]] + pNode[L[7]]);
      }
   f.precision(11);
   f << std::scientific;
   for (int i = 0; i < NN; i++)
      f << o.coord[3 * i] << " " << o.coord[3 * i + 1] << " " << o.coord[3 * i + 2] << " " << U[i] << " " << U[i + NN] << " "
        << U[i + 2 * NN] << " " << pNode[i] << std::endl;
   static const int sub[8][8] = {{0, 8, 20, 11, 12, 21, 26, 24},  {8, 1, 9, 20, 21, 13, 22, 26},  {11, 20, 10, 3, 24, 26, 23, 15},
                                 {20, 9, 2, 10, 26, 22, 14, 23},  {12, 21, 26, 24, 4, 16, 25, 19}, {21, 13, 22, 26, 16, 5, 17, 25},
                                 {24, 26, 23, 15, 19, 25, 18, 7}, {26, 22, 14, 23, 25, 17, 6, 18}};
   for (int e = 0; e < o.NE; e++) {
      const int* L = &o.LtoGnode[(size_t)e * nv];
      if (nv == 27) {
         for (int s = 0; s < 8; s++) {
            for (int q = 0; q < 8; q++) f << L[sub[s][q]] + 1 << (q < 7 ? " " : "");
            f << std::endl;
         }
      } else {   // Q1: the reference writer indexes local nodes 8..26 and cannot run; one brick per element
         for (int q = 0; q < 8; q++) f << L[q] + 1 << (q < 7 ? " " : "");
         f << std::endl;
      }
   }
   return true;
}

}   // namespace

// =============================================================================
// C interface for ctypes (tests / bench cpu_baseline only)
// =============================================================================
extern "C" {

void* orc_create(const char* inpPath, int scheme, int nthreads) {
   Orc* o = new Orc();
   o->scheme = scheme;
#ifdef _OPENMP
   // the thread count is process-wide: nthreads = 0 means "the default again" (OMP_NUM_THREADS or all cores), so that a
   // single-threaded (deterministic) instance does not slow down the ones created after it
   static const int kDefaultThreads = omp_get_max_threads();
   omp_set_num_threads(nthreads > 0 ? nthreads : kDefaultThreads);
#endif
   if (!readInputFile(*o, inpPath)) return o;
   findNeighbors(*o);
   setupMeshColoring(*o);
   setupNonCornerNodes(*o);
   setupLtoGdof(*o);
   determineVelBCnodes(*o);
   findMonitorPoint(*o);
   buildPattern(*o, o->NN, o->NN, o->NENv, o->NENv, o->Mp, o->sparseMapM);
   buildPattern(*o, o->NN, o->NNp, o->NENv, o->NENp, o->Gp, o->sparseMapG);
   if (scheme == 1) buildPattern(*o, o->NNp, o->NNp, o->NENp, o->NENp, o->Zp, o->sparseMapZ);
   setupGQ(*o);
   calcShape(*o);
   calcJacob(*o);
   allocateState(*o);
   step0(*o);
   return o;
}
void orc_destroy(void* h) { delete (Orc*)h; }
const char* orc_error(void* h) { return ((Orc*)h)->err.c_str(); }
int orc_threads() {
#ifdef _OPENMP
   return omp_get_max_threads();
#else
   return 1;
#endif
}

long long orc_int(void* h, const char* name) {
   Orc& o = *(Orc*)h;
   std::string n(name);
   if (n == "NE") return o.NE; if (n == "NCN") return o.NCN; if (n == "NN") return o.NN; if (n == "NNp") return o.NNp;
   if (n == "NENv") return o.NENv; if (n == "NENp") return o.NENp; if (n == "NGP") return o.NGP;
   if (n == "nnzM") return o.Mp.nnz(); if (n == "nnzG") return o.Gp.nnz(); if (n == "nnzZ") return o.Zp.nnz();
   if (n == "BCnVelNodes") return o.BCnVelNodes; if (n == "BCnVelFaces") return o.BCnVelFaces; if (n == "nBC") return o.nBC;
   if (n == "zeroPressureNode") return o.zeroPressureNode; if (n == "monPoint") return o.monPoint;
   if (n == "nActiveColors") return o.nActiveColors; if (n == "colorConflicts") return o.colorConflicts; if (n == "maxIter") return o.maxIter; if (n == "timeN") return o.timeN;
   if (n == "lastIter") return o.lastIter; if (n == "isRestart") return o.isRestart;
   if (n == "n_cgIters") return (long long)o.cgIters.size(); if (n == "n_bicgIters") return (long long)o.bicgIters.size();
   if (n == "n_innerIters") return (long long)o.innerIters.size();
   return -999999;
}
double orc_double(void* h, const char* name) {
   Orc& o = *(Orc*)h;
   std::string n(name);
   if (n == "dt") return o.dt; if (n == "t_ini") return o.t_ini; if (n == "t_final") return o.t_final;
   if (n == "tolerance") return o.tolerance; if (n == "convergenceCriteria") return o.convergenceCriteria;
   if (n == "density") return o.density; if (n == "viscosity") return o.viscosity; if (n == "timeT") return o.timeT;
   if (n == "maxAcc") return o.maxAcc; if (n == "cgSeconds") return o.cgSeconds; if (n == "alpha") return o.alpha;
   return NAN;
}
int orc_set(void* h, const char* name, double v) {
   Orc& o = *(Orc*)h;
   std::string n(name);
   if (n == "cg_tol") o.opt.cg_tol = v; else if (n == "cg_maxit") o.opt.cg_maxit = (int)v;
   else if (n == "cg_warm") o.opt.cg_warm = (int)v; else if (n == "cg_precond") o.opt.cg_precond = (int)v;
   else if (n == "bicg_tol") o.opt.bicg_tol = v; else if (n == "bicg_maxit") o.opt.bicg_maxit = (int)v;
   else if (n == "bicg_halfstep") o.opt.bicg_halfstep = (int)v;
   else if (n == "maxacc_rule") o.opt.maxacc_rule = (int)v; else if (n == "maxIter") o.maxIter = (int)v;
   else if (n == "tolerance") o.tolerance = v;
   else return -1;
   return 0;
}

// Pointer + element count of a named array (valid until destroy / next run).
const void* orc_array(void* h, const char* name, long long* count, int* isInt) {
   Orc& o = *(Orc*)h;
   std::string n(name);
#define DV(nm, v) if (n == nm) { *count = (long long)(v).size(); *isInt = 0; return (v).data(); }
#define IV(nm, v) if (n == nm) { *count = (long long)(v).size(); *isInt = 1; return (v).data(); }
   DV("coord", o.coord) IV("LtoGnode", o.LtoGnode) IV("LtoGvel_1d", o.LtoGvel_1d) IV("BCvelNodes", o.BCvelNodes)
   DV("BCstr", o.BCstr) IV("meshColors", o.meshColors) IV("NmeshColors", o.NmeshColors) IV("elementsOfColor", o.elementsOfColor)
   IV("sparseMrowStarts", o.Mp.rowStarts) IV("sparseMcol", o.Mp.col) IV("sparseGrowStarts", o.Gp.rowStarts) IV("sparseGcol", o.Gp.col)
   IV("sparseZrowStarts", o.Zp.rowStarts) IV("sparseZcol", o.Zp.col) IV("sparseMapM_1d", o.sparseMapM) IV("sparseMapG_1d", o.sparseMapG)
   IV("sparseMapZ_1d", o.sparseMapZ)
   DV("Sv_1d", o.Sv_1d) DV("Sp", o.Sp) DV("dSv", o.dSv) DV("dSp", o.dSp) DV("detJacob", o.detJacob) DV("gDSv_1d", o.gDSv_1d)
   DV("gDSp", o.gDSp) DV("GQfactor_1d", o.GQfactor_1d)
   DV("sparseMvalue", o.Mval) DV("sparseKvalue", o.Kval) DV("sparseAvalue", o.Aval) DV("sparseG1value", o.G1) DV("sparseG2value", o.G2)
   DV("sparseG3value", o.G3) DV("sparseZvalue", o.Zval) DV("MdInv", o.MdInv) DV("MdOrigInv", o.MdOrigInv) DV("MdOrig", o.MdOrig)
   DV("Un", o.Un) DV("Unp1", o.Unp1) DV("Unp1_prev", o.Unp1_prev) DV("UnpHalf", o.UnpHalf) DV("UnpHalf_prev", o.UnpHalf_prev)
   DV("Acc", o.Acc) DV("Acc_prev", o.Acc_prev) DV("KtimesAcc_prev", o.KtimesAcc_prev) DV("R1", o.R1) DV("R2", o.R2) DV("R3", o.R3)
   DV("Pn", o.Pn) DV("Pnp1", o.Pnp1) DV("Pnp1_prev", o.Pnp1_prev) DV("Pdot", o.Pdot)
   DV("Uk", o.Uk) DV("Uk_prev", o.Uk_prev) DV("Pk", o.Pk) DV("Pk_prev", o.Pk_prev) DV("Pk_prevprev", o.Pk_prevprev) DV("Pdiff", o.Pdiff)
   IV("cgIters", o.cgIters) IV("bicgIters", o.bicgIters) IV("innerIters", o.innerIters)
#undef DV
#undef IV
   *count = -1; *isInt = 0;
   return nullptr;
}

// Run n time steps (fixed count; the reference's `while (timeT < t_final)` is FP-governed).
int orc_run(void* h, int nsteps) {
   Orc& o = *(Orc*)h;
   for (int s = 0; s < nsteps; s++) {
      if (o.scheme == 0) bch_timeStep(o); else gq_timeStep(o);
   }
   return 0;
}
// Single phases, for phase-level parity tests of the seam functions.
int orc_phase(void* h, const char* name, int iter) {
   Orc& o = *(Orc*)h;
   std::string n(name);
   if (n == "bch_calculateMatrixA") bch_calculateMatrixA(o);
   else if (n == "bch_step1") bch_step1(o);
   else if (n == "bch_step2") bch_step2(o, iter);
   else if (n == "bch_step3") bch_step3(o, iter);
   else if (n == "bch_KtimesAcc")   // BCH.cpp:2985-3033
      for (int c = 0; c < 3; c++) spmv(o.Mp, o.Kval.data(), &o.Acc_prev[c * o.NN], &o.KtimesAcc_prev[c * o.NN], 1.0, 0.0);
   else if (n == "gq_step1") gq_step1(o);
   else if (n == "gq_step2") gq_step2(o);
   else return -1;
   return 0;
}
int orc_write_tecplot(void* h, const char* path, const char* title) { return createTecplot(*(Orc*)h, path, title) ? 0 : -1; }
int orc_read_restart(void* h, const char* path) { return readRestart(*(Orc*)h, path) ? 0 : -1; }

}   // extern "C"
