// TEST INFRASTRUCTURE -- the five CUDA runtime calls the reference's -DUSECUDA host files make themselves
// (blascoCodinaHuerta.cpp:2866-2870, 2929-2934, 3070-3072, 3099-3111, 3310; guermondQuartapelle.cpp:3366-3370, 3396-3410),
// on HOST memory, for the CPU mock of the seam (mock/cfdb200_mock.cpp).  Linked instead of libcudart by `refhost-mock` only.
#include <cstddef>
#include <cstring>
extern "C" {
int cudaMemcpy(void* dst, const void* src, size_t n, int /*kind*/) { memmove(dst, src, n); return 0; }
int cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return 0; }
int cudaThreadSynchronize(void) { return 0; }
int cudaDeviceSynchronize(void) { return 0; }
int cudaMemGetInfo(size_t* freeB, size_t* totalB) { *freeB = *totalB = 0; return 0; }
const char* cudaGetErrorString(int) { return "fake cudart (host memory)"; }
}
