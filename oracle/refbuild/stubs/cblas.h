/* stand-in for <cblas.h> (no BLAS development package in this image): the prototypes src/Math/Blas.hh uses.
 * Implementations: oracle/refbuild/miniblas.cc (plain loops).  TEST INFRASTRUCTURE ONLY. */
#ifndef MINIBLAS_CBLAS_H
#define MINIBLAS_CBLAS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
typedef CBLAS_ORDER CBLAS_LAYOUT;
void   cblas_sswap(int N, float* X, int incX, float* Y, int incY);
float cblas_snrm2(int N, const float* X, int incX);
float cblas_sasum(int N, const float* X, int incX);
size_t cblas_isamax(int N, const float* X, int incX);
void   cblas_sscal(int N, float alpha, float* X, int incX);
void   cblas_saxpy(int N, float alpha, const float* X, int incX, float* Y, int incY);
float cblas_sdot(int N, const float* X, int incX, const float* Y, int incY);
void   cblas_scopy(int N, const float* X, int incX, float* Y, int incY);
void   cblas_sger(CBLAS_ORDER order, int M, int N, float alpha, const float* X, int incX, const float* Y, int incY, float* A, int lda);
void   cblas_sgemv(CBLAS_ORDER order, CBLAS_TRANSPOSE TransA, int M, int N, float alpha, const float* A, int lda, const float* X, int incX, float beta, float* Y, int incY);
void   cblas_sgemm(CBLAS_ORDER order, CBLAS_TRANSPOSE TransA, CBLAS_TRANSPOSE TransB, int M, int N, int K, float alpha, const float* A, int lda, const float* B, int ldb, float beta, float* C, int ldc);
void   cblas_dswap(int N, double* X, int incX, double* Y, int incY);
double cblas_dnrm2(int N, const double* X, int incX);
double cblas_dasum(int N, const double* X, int incX);
size_t cblas_idamax(int N, const double* X, int incX);
void   cblas_dscal(int N, double alpha, double* X, int incX);
void   cblas_daxpy(int N, double alpha, const double* X, int incX, double* Y, int incY);
double cblas_ddot(int N, const double* X, int incX, const double* Y, int incY);
void   cblas_dcopy(int N, const double* X, int incX, double* Y, int incY);
void   cblas_dger(CBLAS_ORDER order, int M, int N, double alpha, const double* X, int incX, const double* Y, int incY, double* A, int lda);
void   cblas_dgemv(CBLAS_ORDER order, CBLAS_TRANSPOSE TransA, int M, int N, double alpha, const double* A, int lda, const double* X, int incX, double beta, double* Y, int incY);
void   cblas_dgemm(CBLAS_ORDER order, CBLAS_TRANSPOSE TransA, CBLAS_TRANSPOSE TransB, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double beta, double* C, int ldc);
#ifdef __cplusplus
}
#endif
#endif
