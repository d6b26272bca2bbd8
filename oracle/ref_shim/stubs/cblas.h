/* TEST INFRASTRUCTURE — minimal <cblas.h> for the reference's TMatrix.cpp
 * (libvvhd/src/TMatrix.cpp:11,176). The four BLAS/LAPACK entry points are
 * provided by ref_stubs.cpp, which forwards them to scipy's bundled OpenBLAS
 * (dlopen at first use), because the image has no system LAPACK. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
void cblas_dgemv(CBLAS_ORDER order, CBLAS_TRANSPOSE trans, int M, int N, double alpha,
                 const double* A, int lda, const double* X, int incX, double beta,
                 double* Y, int incY);
#ifdef __cplusplus
}
#endif
