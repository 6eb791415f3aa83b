// solverGPU -- the reference's executable as a drop-in program (process-level contract).
// Like the reference's `solverGPU` (fractionalStep/explicit/Cpp/makeBCH:16-18, blascoCodinaHuerta.cpp:383-520; implicit:
// makeGQ, guermondQuartapelle.cpp:367-520) it takes no arguments: it reads ./ProblemName.txt, then <name>.inp (and
// <name>_restart.dat when isRestart is 1), prints the monitor table and writes <name>.dat.  The reference builds one
// program per scheme; so does this file (-DCFD_SCHEME=0 explicit / 1 implicit -> bin/solverGPU_explicit, bin/solverGPU_implicit).
// All work happens in libcfdb200.so (cfd_solver_main); solver settings the reference hard-codes differently in its four
// builds can be chosen through the CFD_* environment variables listed in include/cfdb200.h.
#include <cstdio>
#include <cstdlib>

#include "../../include/cfdb200.h"

#ifndef CFD_SCHEME
#define CFD_SCHEME CFD_EXPLICIT
#endif

int main() {
   const char* ms = getenv("CFD_MAX_STEPS");
   const int rc = cfd_solver_main("", CFD_SCHEME, ms ? atoi(ms) : 0);
   if (rc != CFD_OK) {
      fprintf(stderr, "solverGPU: error %d: %s\n", rc, cfd_last_error());
      return 1;
   }
   return 0;
}
