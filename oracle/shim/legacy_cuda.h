// TEST INFRASTRUCTURE -- force-included (-include) when the reference's host files are compiled with -DUSECUDA against
// today's CUDA headers: the legacy cuSPARSE type they declare a global of (BCH.cpp:318, GQ.cpp likewise) was removed in
// CUDA 11.  Nothing uses the objects once CUDAcodes.cu is replaced by integration/cfdb200_shim_*.cpp.
#pragma once
typedef void* cusparseSolveAnalysisInfo_t;
