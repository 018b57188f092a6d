// exact_div_test.cu -- diagnostic kernels behind chl_debug_exact_div / chl_debug_exact_sqrt (include/cholla_b200.h): evaluates the
// shared-reciprocal division of exact_div.cuh and the compiler's own a / b on the same operands so that
// tests/test_gpu_exact_div.py can compare them bit for bit.  Built -fmad=false like the strict kernels.
#define CHL_STRICT 1
#define CHL_MODE f64_divtest
#include "hydro_physics.cuh"  // exact_div.cuh + sqrt_r
#include "chl_aux.h"

namespace chl_f64_divtest
{
__global__ void div_kernel(const double *a, const double *b, double *q_shared, double *q_plain, long n)
{
  long const i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Den const d = den(b[i]);
  q_shared[i] = qdiv(a[i], d);
  q_plain[i]  = a[i] / b[i];
}
// the branch-free policy (kDivFlag): raw quotient + whether the sticky flag came up; the kernels redo a flagged block
// with IEEE divisions, so the contract is "flag clear => the bits of a / b"
__global__ void div_flag_kernel(const double *a, const double *b, double *q, int *flag, int zero_ok, long n)
{
  long const i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  KCF k;
  k.bad       = false;
  Den const d = den(b[i], k);
  q[i]        = zero_ok ? qdiv<true>(a[i], d, k) : qdiv<false>(a[i], d, k);
  int f       = k.bad ? 1 : 0;
  k.bad       = false;
  double const r = sqrt_r(a[i], k);
  if (k.bad) f |= 2;
  else if (r != sqrt(a[i]) && !(r != r)) f |= 4;  // a clear flag with the wrong root: reported as bit 2
  flag[i] = f;
}
// strict mode's sqrt_r (the compiler's fast-path sequence without its marshalling) against sqrt()
__global__ void sqrt_kernel(const double *x, double *r_own, double *r_plain, long n)
{
  long const i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  r_own[i]   = sqrt_r(x[i]);
  r_plain[i] = sqrt(x[i]);
}
}  // namespace chl_f64_divtest

namespace chl
{
cudaError_t launch_exact_div_test(const double *a, const double *b, double *q_shared, double *q_plain, long n,
                                  cudaStream_t s)
{
  if (n <= 0) return cudaSuccess;
  chl_f64_divtest::div_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a, b, q_shared, q_plain, n);
  return cudaGetLastError();
}
cudaError_t launch_exact_div_flag_test(const double *a, const double *b, double *q, int *flag, int zero_ok, long n,
                                       cudaStream_t s)
{
  if (n <= 0) return cudaSuccess;
  chl_f64_divtest::div_flag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a, b, q, flag, zero_ok, n);
  return cudaGetLastError();
}
cudaError_t launch_exact_sqrt_test(const double *x, double *r_own, double *r_plain, long n, cudaStream_t s)
{
  if (n <= 0) return cudaSuccess;
  chl_f64_divtest::sqrt_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, r_own, r_plain, n);
  return cudaGetLastError();
}
}  // namespace chl
