// Secondary GPU baseline (NOT the product, NOT the oracle): the reference's GPU pressure solve re-created with the
// libraries' current APIs.  The reference calls cusp::krylov::cg with a diagonal preconditioner (CUDAcodes.cu:418-451);
// its older sibling in oldFiles/ is a host loop of cusparseDcsrmv + cublasDdot + cublasDaxpy with the dot products read
// back to the host every iteration (oldFiles/poissonSolver/CUSPARSEsolver.cu:80-128).  CUSP and the legacy csrmv API no
// longer exist in CUDA 12.9, so this file is that call sequence on cusparseSpMV + cuBLAS: one library call per vector
// operation, blocking scalar read-backs, Jacobi preconditioner, monitor rule ||r|| <= tol ||b||.  It exists only so that
// bench / DESIGN can quote "library-call implementation on the same B200" beside the fused persistent kernel.
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusparse.h>

#include <cmath>
#include <cstdio>

__global__ void k_jacobi(int n, const double* dinv, const double* r, double* z) {
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) z[i] = dinv[i] * r[i];
}

extern "C" int lib_pcg(int n, int nnz, const int* rs, const int* col, const double* val, const double* dinv, const double* b, double* x,
                       double tol, int maxit, int* its, double* ms) {
   int *d_rs, *d_col;
   double *d_val, *d_dinv, *d_b, *d_x, *d_r, *d_z, *d_p, *d_q;
   cudaMalloc(&d_rs, sizeof(int) * (n + 1)); cudaMalloc(&d_col, sizeof(int) * nnz); cudaMalloc(&d_val, sizeof(double) * nnz);
   cudaMalloc(&d_dinv, 8 * n); cudaMalloc(&d_b, 8 * n); cudaMalloc(&d_x, 8 * n); cudaMalloc(&d_r, 8 * n); cudaMalloc(&d_z, 8 * n);
   cudaMalloc(&d_p, 8 * n); cudaMalloc(&d_q, 8 * n);
   cudaMemcpy(d_rs, rs, sizeof(int) * (n + 1), cudaMemcpyHostToDevice); cudaMemcpy(d_col, col, sizeof(int) * nnz, cudaMemcpyHostToDevice);
   cudaMemcpy(d_val, val, sizeof(double) * nnz, cudaMemcpyHostToDevice); cudaMemcpy(d_dinv, dinv, 8 * n, cudaMemcpyHostToDevice);
   cudaMemcpy(d_b, b, 8 * n, cudaMemcpyHostToDevice);
   cudaMemset(d_x, 0, 8 * n);
   cusparseHandle_t sp; cublasHandle_t bl;
   cusparseCreate(&sp); cublasCreate(&bl);
   cusparseSpMatDescr_t A; cusparseDnVecDescr_t vp, vq;
   cusparseCreateCsr(&A, n, n, nnz, d_rs, d_col, d_val, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F);
   cusparseCreateDnVec(&vp, n, d_p, CUDA_R_64F); cusparseCreateDnVec(&vq, n, d_q, CUDA_R_64F);
   const double one = 1.0, zero = 0.0;
   size_t bufSize = 0; void* buf = nullptr;
   cusparseSpMV_bufferSize(sp, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, A, vp, &zero, vq, CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, &bufSize);
   cudaMalloc(&buf, bufSize ? bufSize : 8);
   cusparseSpMV_preprocess(sp, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, A, vp, &zero, vq, CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, buf);
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   cudaDeviceSynchronize();
   cudaEventRecord(e0);
   // x0 = 0: r = b, z = M r, p = z
   cudaMemcpy(d_r, d_b, 8 * n, cudaMemcpyDeviceToDevice);
   const int grid = (n + 255) / 256;
   k_jacobi<<<grid, 256>>>(n, d_dinv, d_r, d_z);
   cudaMemcpy(d_p, d_z, 8 * n, cudaMemcpyDeviceToDevice);
   double rz, bnorm, rnorm;
   cublasDdot(bl, n, d_r, 1, d_z, 1, &rz);
   cublasDnrm2(bl, n, d_b, 1, &bnorm);
   int it = 0;
   while (true) {
      cublasDnrm2(bl, n, d_r, 1, &rnorm);                       // monitor.finished(r): host read-back
      if (rnorm <= tol * bnorm || it >= maxit) break;
      cusparseSpMV(sp, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, A, vp, &zero, vq, CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, buf);
      double pq;
      cublasDdot(bl, n, d_p, 1, d_q, 1, &pq);                   // host read-back
      const double alpha = rz / pq, malpha = -alpha;
      cublasDaxpy(bl, n, &alpha, d_p, 1, d_x, 1);
      cublasDaxpy(bl, n, &malpha, d_q, 1, d_r, 1);
      k_jacobi<<<grid, 256>>>(n, d_dinv, d_r, d_z);
      double rzNew;
      cublasDdot(bl, n, d_r, 1, d_z, 1, &rzNew);                // host read-back
      const double beta = rzNew / rz;
      rz = rzNew;
      cublasDscal(bl, n, &beta, d_p, 1);
      cublasDaxpy(bl, n, &one, d_z, 1, d_p, 1);
      it++;
   }
   cudaEventRecord(e1);
   cudaEventSynchronize(e1);
   float t = 0;
   cudaEventElapsedTime(&t, e0, e1);
   cudaMemcpy(x, d_x, 8 * n, cudaMemcpyDeviceToHost);
   if (its) *its = it;
   if (ms) *ms = t;
   cusparseDestroySpMat(A); cusparseDestroyDnVec(vp); cusparseDestroyDnVec(vq);
   cusparseDestroy(sp); cublasDestroy(bl);
   cudaFree(buf); cudaFree(d_rs); cudaFree(d_col); cudaFree(d_val); cudaFree(d_dinv); cudaFree(d_b); cudaFree(d_x); cudaFree(d_r);
   cudaFree(d_z); cudaFree(d_p); cudaFree(d_q);
   return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
