// exact_div_test.cu -- diagnostic kernel behind chl_debug_exact_div (include/cholla_b200.h): evaluates the
// shared-reciprocal division of exact_div.cuh and the compiler's own a / b on the same operands so that
// tests/test_gpu_exact_div.py can compare them bit for bit.  Built -fmad=false like the strict kernels.
#define CHL_STRICT 1
#define CHL_MODE f64_divtest
#define CHL_NS chl_f64_divtest
#include "exact_div.cuh"
#include "chl_internal.h"

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
}  // namespace chl
