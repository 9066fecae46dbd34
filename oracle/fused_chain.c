/*
 * ORACLE (test / baseline infrastructure, not product code) — plain C restatement of the loops
 * the reference generates for the BASELINE programs and runs under Numba with
 * parallel=True, fastmath=True (ramba/ramba.py:8247-8265, 3758-3780; ramba/common.py:28).
 *
 * chain_f64: the fused op of sample/test-ramba.py:12-17 as the reference prints it with
 * RAMBA_SHOW_CODE=1:
 *     B[index] = numpy.sin(A[index])
 *     t4 = B[index] * B[index]
 *     C[index] = numpy.cos(A[index])
 *     t6 = C[index] ** 2                 # int exponent -> powi -> C*C (Numba int_power)
 *     D[index] = t4 + t6
 * and, with make_A != 0, the one-shot README form that first writes
 *     A[index] = (index[0] + global_start[0]) * 0.001   (ramba/ramba.py:8955-8960, 6121-6126)
 * numba.pndindex -> OpenMP static schedule; libm sin/cos like Numba's lowering of numpy.sin/cos.
 *
 * sum_affine_f32: `(X*2.0 + 1.0).sum()` stage 1 (ramba/ramba.py:5798-5807): float32 element times
 * float64 scalar -> float64, float64 accumulator.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp; no -ffast-math so results are reproducible).
 */
#include <math.h>
#include <stdint.h>
#include <omp.h>

void chain_f64(double* A, double* B, double* C, double* D, int64_t n, int64_t global_start, int make_A) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double a;
    if (make_A) {
      a = (double)(i + global_start) * 0.001;
      A[i] = a;
    } else {
      a = A[i];
    }
    double b = sin(a);
    B[i] = b;
    double t4 = b * b;
    double c = cos(a);
    C[i] = c;
    double t6 = c * c;
    D[i] = t4 + t6;
  }
}

double sum_affine_f32(const float* X, int64_t n, double mul, double add) {
  double acc = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : acc)
  for (int64_t i = 0; i < n; ++i) acc += (double)X[i] * mul + add;
  return acc;
}

/* laplace7_f32: BASELINE config 4, the optimised kernel the reference generates for
 *   V[1:-1,1:-1,1:-1] = U[:-2,..] + U[2:,..] + U[..,:-2,..] + U[..,2:,..] + U[..,:-2] + U[..,2:] - 6.0*U[1:-1,..]
 * (ramba/ramba.py:8146-8188): float32 neighbour sum, float64 `6.0*u` and subtraction, rounded on store. */
void laplace7_f32(const float* U, float* V, int64_t n) {
#pragma omp parallel for schedule(static) collapse(2)
  for (int64_t i = 1; i < n - 1; ++i)
    for (int64_t j = 1; j < n - 1; ++j) {
      const float* u = U + (i * n + j) * n;
      float* v = V + (i * n + j) * n;
      for (int64_t k = 1; k < n - 1; ++k) {
        float s = u[k - n * n] + u[k + n * n];
        s = s + u[k - n];
        s = s + u[k + n];
        s = s + u[k - 1];
        s = s + u[k + 1];
        v[k] = (float)((double)s - 6.0 * (double)u[k]);
      }
    }
}

/* bcast_add_axis0_sum_f32: BASELINE config 5 stage 1, `red[j] = red[j] + (M[i,j] + v[j])` walked row by row through the
 * float32 partial array (ramba/ramba.py:8231-8244); columns are split over the threads. */
void bcast_add_axis0_sum_f32(const float* M, const float* v, float* red, int64_t rows, int64_t cols) {
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < cols; ++j) red[j] = 0.0f;
#pragma omp parallel
  {
    const int nt = omp_get_num_threads(), t = omp_get_thread_num();
    const int64_t c0 = cols * t / nt, c1 = cols * (t + 1) / nt;
    for (int64_t i = 0; i < rows; ++i) {
      const float* m = M + i * cols;
      for (int64_t j = c0; j < c1; ++j) red[j] = red[j] + (m[j] + v[j]);
    }
  }
}

int oracle_num_threads(void) { return omp_get_max_threads(); }
void oracle_set_num_threads(int n) { omp_set_num_threads(n); }
