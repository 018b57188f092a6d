// cooling_kernels.cu -- the optically-thin cooling source step that follows the integrator in a -DCOOLING_GPU build
// (Grid3D::Update_Hydro_Grid, src/grid/grid3D.cpp:512-522 -> Cooling_Update / cooling_kernel,
// src/cooling/cooling_cuda.cu:20-196) with the analytic collisional-ionisation-equilibrium fit CIE_cool (:293-317), which is
// what builds/make.type.cooling selects.  (The -DCLOUDY_COOL variant interpolates tables that are read from files outside
// the source tree; it is not built.)  One elementwise pass over the real cells: temperature from the pressure -- from the
// advected internal energy in a dual-energy context --, sub-cycled so that no sub-step changes T by more than 1 %, then the
// energy (and GasEnergy) lowered by the total change.
//
// Built -fmad=false for every arithmetic mode: the expressions keep the reference's order and log10 / pow are the same
// device library the reference's build links, so its bits come out.
#include <math.h>

#include "chl_aux.h"

namespace chl
{
namespace
{
// src/global/global.h:26-47
constexpr double kMP = 1.672622e-24, kKB = 1.380658e-16;
constexpr double kTimeUnit = 3.15569e10, kLengthUnit = 3.08567758e21, kMassUnit = 1.98847e33;
constexpr double kDensityUnit  = (kMassUnit / (kLengthUnit * kLengthUnit * kLengthUnit));
constexpr double kVelocityUnit = (kLengthUnit / kTimeUnit);
constexpr double kEnergyUnit   = (kDensityUnit * kVelocityUnit * kVelocityUnit);
constexpr double kPressureUnit = (kDensityUnit * kVelocityUnit * kVelocityUnit);
constexpr double kSpEnergyUnit = (kVelocityUnit * kVelocityUnit);
constexpr double kTiny = 1.0e-20;

template <typename T>
__device__ T cie_cool(T n, T Tk)
{  // analytic fit to a solar-metallicity CIE cooling curve, cooling_cuda.cu:293-317
  T lambda = 0.0;
  if (log10(Tk) < (T)4.0) {
    lambda = 0.0;
  } else if (log10(Tk) >= (T)4.0 && log10(Tk) < (T)5.9) {
    lambda = pow((T)10.0, ((T)-1.3 * (log10(Tk) - (T)5.25) * (log10(Tk) - (T)5.25) - (T)21.25));
  } else if (log10(Tk) >= (T)5.9 && log10(Tk) < (T)7.4) {
    lambda = pow((T)10.0, ((T)0.7 * (log10(Tk) - (T)7.1) * (log10(Tk) - (T)7.1) - (T)22.8));
  } else {
    lambda = pow((T)10.0, ((T)0.45 * log10(Tk) - (T)26.065));
  }
  return n * n * lambda;
}

template <typename T>
__global__ void __launch_bounds__(256) cooling_kernel(T *__restrict__ U, int nx, int ny, int nz, int g, int n_fields,
                                                      int dual_energy, T dt_in, T gamma)
{
  long const N = (long)nx * ny * nz;
  int const rx = nx - 2 * g, ry = ny > 1 ? ny - 2 * g : 1, rz = nz > 1 ? nz - 2 * g : 1;
  long const nreal = (long)rx * ry * rz;
  T const mu = (T)0.6;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < nreal; t += (long)gridDim.x * blockDim.x) {
    int const x = (int)(t % rx) + g;
    long const r = t / rx;
    int const y = ny > 1 ? (int)(r % ry) + g : 0, z = nz > 1 ? (int)(r / ry) + g : 0;
    long const id = x + (long)nx * (y + (long)ny * z);
    T dt = dt_in;
    T const d = U[id];
    T E = U[4 * N + id];
    if (E < (T)0.0 || E != E) continue;  // "don't apply cooling if this thread crashed"
    T const vx = U[N + id] / d, vy = U[2 * N + id] / d, vz = U[3 * N + id] / d;
    T p = (E - (T)0.5 * d * (vx * vx + vy * vy + vz * vz)) * (gamma - (T)1.0);
    p   = fmax(p, (T)kTiny);
    T ge = 0;
    if (dual_energy) {
      ge = U[(long)(n_fields - 1) * N + id] / d;
      ge = fmax(ge, (T)kTiny);
    }
    T const n = d * (T)kDensityUnit / (mu * (T)kMP);  // number density, cgs
    T T_init = p * (T)kPressureUnit / (n * (T)kKB);
    if (dual_energy) T_init = d * ge * (gamma - (T)1.0) * (T)kPressureUnit / (n * (T)kKB);
    T Tk   = T_init;
    T cool = cie_cool(n, Tk);
    T del_T = cool * dt * (T)kTimeUnit * (gamma - (T)1.0) / (n * (T)kKB);
    while (del_T / Tk > (T)0.01) {  // limit the change in temperature to 1 % per sub-step
      T const dt_sub = (T)0.01 * Tk * n * (T)kKB / (cool * (T)kTimeUnit * (gamma - (T)1.0));
      Tk -= cool * dt_sub * (T)kTimeUnit * (gamma - (T)1.0) / (n * (T)kKB);
      dt -= dt_sub;
      cool  = cie_cool(n, Tk);
      del_T = cool * dt * (T)kTimeUnit * (gamma - (T)1.0) / (n * (T)kKB);
    }
    Tk -= del_T;
    del_T = T_init - Tk;  // total change in T
    E -= n * (T)kKB * del_T / ((gamma - (T)1.0) * (T)kEnergyUnit);
    U[4 * N + id] = E;
    if (dual_energy) {
      ge -= (T)kKB * del_T / (mu * (T)kMP * (gamma - (T)1.0) * (T)kSpEnergyUnit);
      U[(long)(n_fields - 1) * N + id] = d * ge;
    }
  }
}
}  // namespace

cudaError_t launch_cooling_update(void *U, int nx, int ny, int nz, int n_ghost, int n_fields, int dual_energy, double dt,
                                  double gamma, int real_bytes, cudaStream_t s)
{
  long const nreal = (long)(nx - 2 * n_ghost) * (ny > 1 ? ny - 2 * n_ghost : 1) * (nz > 1 ? nz - 2 * n_ghost : 1);
  long b = (nreal + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  if (real_bytes == 8)
    cooling_kernel<double><<<(int)b, 256, 0, s>>>((double *)U, nx, ny, nz, n_ghost, n_fields, dual_energy, dt, gamma);
  else
    cooling_kernel<float><<<(int)b, 256, 0, s>>>((float *)U, nx, ny, nz, n_ghost, n_fields, dual_energy, (float)dt,
                                                 (float)gamma);
  return cudaGetLastError();
}
}  // namespace chl
