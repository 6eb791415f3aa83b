// Reference-element tables for the Q2/Q1 (27/8-node) and Q1/Q1 hexahedra.
//
// Same functions as the reference's setupGQ()/calcShape()
// (fractionalStep/explicit/Cpp/blascoCodinaHuerta.cpp:2166-2488) but generated as tensor
// products of 1-D Lagrange bases from a node-position table instead of 27 hand-expanded lines.
#include <cmath>

#include "common.hpp"

namespace cfdb {

namespace {
// (xi, eta, zeta) in {-1,0,+1} of local nodes 0..26: 8 corners, 12 mid-edges, 6 mid-faces, centre
// (node order implied by blascoCodinaHuerta.cpp:2312-2343).
const signed char kNodePos[27][3] = {
    {-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1},
    {0, -1, -1},  {1, 0, -1},  {0, 1, -1}, {-1, 0, -1}, {-1, -1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 1, 0},
    {0, -1, 1},   {1, 0, 1},   {0, 1, 1},  {-1, 0, 1},  {0, 0, -1},  {0, -1, 0}, {1, 0, 0}, {0, 1, 0},
    {-1, 0, 0},   {0, 0, 1},   {0, 0, 0}};

inline double quad(int p, double t) { return p < 0 ? 0.5 * t * (t - 1.0) : (p > 0 ? 0.5 * t * (t + 1.0) : 1.0 - t * t); }
inline double dquad(int p, double t) { return p < 0 ? t - 0.5 : (p > 0 ? t + 0.5 : -2.0 * t); }
inline double lin(int p, double t) { return 0.5 * (1.0 + p * t); }
inline double dlin(int p, double) { return 0.5 * p; }
}   // namespace

void build_fe_tables(int NENv, FeTables& T) {
   T.NENv = NENv;
   const double g = 1.0 / std::sqrt(3.0);
   for (int k = 0; k < 8; k++) {
      T.gq[k][0] = (k & 1) ? g : -g;
      T.gq[k][1] = (k & 2) ? g : -g;
      T.gq[k][2] = (k & 4) ? g : -g;
   }
   for (int k = 0; k < 8; k++) {
      const double x = T.gq[k][0], y = T.gq[k][1], z = T.gq[k][2];
      for (int i = 0; i < 8; i++) {
         const int a = kNodePos[i][0], b = kNodePos[i][1], c = kNodePos[i][2];
         T.Sp[k][i] = lin(a, x) * lin(b, y) * lin(c, z);
         T.dSp[0][i][k] = dlin(a, x) * lin(b, y) * lin(c, z);
         T.dSp[1][i][k] = lin(a, x) * dlin(b, y) * lin(c, z);
         T.dSp[2][i][k] = lin(a, x) * lin(b, y) * dlin(c, z);
      }
      for (int i = 0; i < 27; i++) {
         const int a = kNodePos[i][0], b = kNodePos[i][1], c = kNodePos[i][2];
         if (NENv == 27) {
            T.Sv[k][i] = quad(a, x) * quad(b, y) * quad(c, z);
            T.dSv[0][i][k] = dquad(a, x) * quad(b, y) * quad(c, z);
            T.dSv[1][i][k] = quad(a, x) * dquad(b, y) * quad(c, z);
            T.dSv[2][i][k] = quad(a, x) * quad(b, y) * dquad(c, z);
         } else if (i < 8) {   // Q1/Q1: Sv = Sp, dSv = dSp (the commented intent at blascoCodinaHuerta.cpp:2303-2305)
            T.Sv[k][i] = T.Sp[k][i];
            for (int d = 0; d < 3; d++) T.dSv[d][i][k] = T.dSp[d][i][k];
         } else {
            T.Sv[k][i] = 0.0;
            for (int d = 0; d < 3; d++) T.dSv[d][i][k] = 0.0;
         }
      }
   }
}

}   // namespace cfdb
