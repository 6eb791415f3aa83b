// Shared declarations of the B200-native fractional-step library (product code).
#pragma once
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/cfdb200.h"

namespace cfdb {

void set_error(const char* fmt, ...);

// Reference-element tables (blascoCodinaHuerta.cpp:2166-2488), built by fe_tables.cpp.
struct FeTables {
   int NENv = 0;
   double gq[8][3];          // Gauss points, xi fastest
   double Sp[8][8];          // [k][i]   trilinear
   double dSp[3][8][8];      // [d][i][k]
   double Sv[8][27];         // [k][i]   triquadratic (or = Sp when NENv == 8)
   double dSv[3][27][8];     // [d][i][k]
};
void build_fe_tables(int NENv, FeTables& t);

}   // namespace cfdb
