// TEST INFRASTRUCTURE -- not product code.
// C entry point around the PARALUTION stand-in, used by tests/test_shim.py only: drives the classes with exactly the call
// sequence of the reference's paralution_BiCGStab_solver() (guermondQuartapelle.cpp:4183-4231).
#include "paralution.hpp"

using namespace paralution;

extern "C" int shim_paralution_bicgstab(int n, int nnz, int* rowStarts, int* col, double* val, double* rhs, double* x, double abs_tol,
                                        double rel_tol, double div_tol, int max_iter) {
   int *rs = rowStarts, *c = col;
   double *v = val, *b = rhs, *sol = x;
   LocalMatrix<double> A_p;
   LocalVector<double> R1_p, Uk_dir_p;
   A_p.SetDataPtrCSR(&rs, &c, &v, "BiCGStab LHS", nnz, n, n);
   R1_p.SetDataPtr(&b, "BiCGStab RHS", n);
   Uk_dir_p.SetDataPtr(&sol, "BiCGStab result", n);
   if (rs || c || v || b || sol) return -1;   // SetDataPtr takes the arrays away from the caller
   Uk_dir_p.Zeros();
   BiCGStab<LocalMatrix<double>, LocalVector<double>, double> ls;
   ls.Init(abs_tol, rel_tol, div_tol, max_iter);
   Jacobi<LocalMatrix<double>, LocalVector<double>, double> p;
   ls.SetOperator(A_p);
   ls.SetPreconditioner(p);
   ls.Build();
   ls.Solve(R1_p, &Uk_dir_p);
   A_p.LeaveDataPtrCSR(&rs, &c, &v);
   R1_p.LeaveDataPtr(&b);
   Uk_dir_p.LeaveDataPtr(&sol);
   ls.Clear();
   return (rs == rowStarts && c == col && v == val && b == rhs && sol == x) ? 0 : -2;   // ... and LeaveDataPtr hands them back
}

// ---- the explicit-Z product of integration/cfdb200_explicit_z.hpp, for tests/test_shim.py ---------------------------------
#include "../../integration/cfdb200_explicit_z.hpp"
extern "C" int shim_form_explicit_Z(int NN, int NNp, const int* growStarts, const int* gcol, const double* G1, const double* G2, const double* G3,
                                    const double* MdOrigInv, int pin, int* zrs, int* zcol, double* zval, int cap) {
   std::vector<int> rs, col;
   std::vector<double> val;
   cfdb200_form_explicit_Z(NN, NNp, growStarts, gcol, G1, G2, G3, MdOrigInv, pin, rs, col, val);
   if ((int)col.size() > cap) return -(int)col.size();
   std::copy(rs.begin(), rs.end(), zrs);
   std::copy(col.begin(), col.end(), zcol);
   std::copy(val.begin(), val.end(), zval);
   return (int)col.size();
}
