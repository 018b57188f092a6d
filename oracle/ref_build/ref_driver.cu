// ref_driver.cu -- C entry points over the UNMODIFIED reference kernels (cholla @ 8478c9e), compiled from
// /root/reference/src where they lie (see Makefile in this directory).  TEST/BENCH INFRASTRUCTURE ONLY: it is
// the "real reference" arm the CPU oracle and the B200-native kernels are validated and timed against.
// Nothing here is reference code: these are thin calls into the reference's own public host functions
//   VL_Algorithm_{1,2,3}D_CUDA (src/integrators/VL_*_cuda.h), Calc_dt_GPU, Apply_Temperature_Floor
//   (src/hydro/hydro_cuda.h), SetGhostCells (src/grid/cuda_boundaries.h)
// plus the symbols those objects need from translation units we do not build: `gama`
// (src/global/global.cpp:19), `chexit` (src/utils/error_handling.cpp:25) and `chprintf` (src/io/io.cpp).
#include <cstdio>
#include <cstdlib>

#include "global/global.h"
#include "global/global_cuda.h"
#include "grid/cuda_boundaries.h"
#include "hydro/hydro_cuda.h"
#include "integrators/VL_1D_cuda.h"
#include "integrators/VL_2D_cuda.h"
#include "integrators/VL_3D_cuda.h"
#include "integrators/simple_1D_cuda.h"
#include "integrators/simple_2D_cuda.h"
#include "integrators/simple_3D_cuda.h"
#ifdef COOLING_GPU
  #include "cooling/cooling_cuda.h"
#endif

Real gama;
[[noreturn]] void chexit(int code) { exit(code); }
// chprintf (src/io/io.h:29, defined in io.cpp: printf on the root rank): simple_*D_cuda.cu reports its allocation with it
#include <cstdarg>
int chprintf(const char *__restrict sdata, ...)
{
  va_list ap;
  va_start(ap, sdata);
  int const n = getenv("REF_DRIVER_VERBOSE") ? vprintf(sdata, ap) : 0;
  va_end(ap);
  return n;
}
#include "grid/grid_enum.h"
// 5; 8 in an -DMHD build; + NSCALARS with -DSCALAR; + 1 (GasEnergy, last) with -DDE
static const int kFields = grid_enum::num_fields;

static void boundary_extents(int dir, int nx, int ny, int nz, int g, int *lo, int *hi)
{  // Grid3D::Set_Boundary_Extents lives in a host TU (boundary_conditions.cpp:419-) that needs the whole Grid3D;
   // restated here: a face spans real cells of later axes, everything of earlier axes
  int n[3] = {nx, ny, nz};
  lo[0] = lo[1] = lo[2] = 0;
  hi[0] = nx; hi[1] = ny; hi[2] = nz;
  if (ny > 1) { lo[1] = g; hi[1] = ny - g; }
  if (nz > 1) { lo[2] = g; hi[2] = nz - g; }
  int ax = dir / 2;
  for (int d = 0; d < ax; d++) { lo[d] = 0; hi[d] = n[d]; }
  if (dir % 2 == 0) { lo[ax] = 0; hi[ax] = g; } else { lo[ax] = n[ax] - g; hi[ax] = n[ax]; }
}

// Parameters::custom_grav and the domain's lower corner, as Grid3D hands them to the integrators (grid3D.cpp:431-465);
// only a -DSTATIC_GRAV build reads them
static int g_custom_grav = 0;
static double g_bound[3] = {0.0, 0.0, 0.0};
// the device potential a -DGRAVITY -DGRAVITY_GPU build's Poisson solver would have left for the integrator
static double *g_potential = nullptr;

extern "C" {

const char *ref_build_flags(void) { return REF_BUILD_FLAGS; }
void ref_set_potential(double *d_potential) { g_potential = d_potential; }
void ref_set_gravity(int custom_grav, double xbound, double ybound, double zbound)
{
  g_custom_grav = custom_grav;
  g_bound[0] = xbound; g_bound[1] = ybound; g_bound[2] = zbound;
}

void ref_set_gamma(double g) { gama = g; }

// one integrator step (VL, or SIMPLE in a -DSIMPLE build), dimensionality from ny/nz as in
// Grid3D::Execute_Hydro_Integrator (grid3D.cpp:413-470)
void ref_vl_step(double *d_conserved, int nx, int ny, int nz, int n_ghost, double dx, double dy, double dz, double dt,
                 int n_fields, double density_floor)
{
#ifdef SIMPLE
  if (ny == 1 && nz == 1) Simple_Algorithm_1D_CUDA(d_conserved, nx, 0, n_ghost, dx, 0.0, dt, n_fields, 0);
  else if (nz == 1) Simple_Algorithm_2D_CUDA(d_conserved, nx, ny, 0, 0, n_ghost, dx, dy, 0.0, 0.0, dt, n_fields, 0);
  else
    Simple_Algorithm_3D_CUDA(d_conserved, nullptr, nx, ny, nz, 0, 0, 0, n_ghost, dx, dy, dz, 0.0, 0.0, 0.0, dt, n_fields,
                             0, density_floor, nullptr);
#else
  if (ny == 1 && nz == 1) VL_Algorithm_1D_CUDA(d_conserved, nx, 0, n_ghost, dx, g_bound[0], dt, n_fields, g_custom_grav);
  else if (nz == 1)
    VL_Algorithm_2D_CUDA(d_conserved, nx, ny, 0, 0, n_ghost, dx, dy, g_bound[0], g_bound[1], dt, n_fields, g_custom_grav);
  else {
  #ifdef REF_NO_VL3D  // the reference's VL_3D_cuda.cu does not compile in this configuration (see the Makefile)
    fprintf(stderr, "ref_driver: this reference build has no 3-D Van Leer integrator\n");
    chexit(-1);
  #else
    VL_Algorithm_3D_CUDA(d_conserved, g_potential, nx, ny, nz, 0, 0, 0, n_ghost, dx, dy, dz, g_bound[0], g_bound[1],
                         g_bound[2], dt, n_fields, g_custom_grav, density_floor, nullptr);
  #endif
  }
#endif
}

double ref_calc_dti(double *d_conserved, int nx, int ny, int nz, int n_ghost, int n_fields, double dx, double dy,
                    double dz)
{
  return Calc_dt_GPU(d_conserved, nx, ny, nz, n_ghost, n_fields, dx, dy, dz, gama);
}

void ref_temperature_floor(double *d_conserved, int nx, int ny, int nz, int n_ghost, int n_fields, double U_floor)
{
  Apply_Temperature_Floor(d_conserved, nx, ny, nz, n_ghost, n_fields, U_floor);
}

void ref_set_boundaries(double *d_conserved, int nx, int ny, int nz, int n_fields, int n_ghost, int *flags)
{
  int ndim = 1 + (ny > 1) + (nz > 1);
  for (int dir = 0; dir < 2 * ndim; dir++) {
    if (flags[dir] == 4 || flags[dir] == 5) continue;
    int lo[3], hi[3];
    boundary_extents(dir, nx, ny, nz, n_ghost, lo, hi);
    SetGhostCells(d_conserved, nx, ny, nz, n_fields, nx * ny * nz, n_ghost, flags, hi[0] - lo[0], hi[1] - lo[1],
                  hi[2] - lo[2], lo[0], lo[1], lo[2], dir);
  }
}

// The body of main()'s loop (main.cpp:247-379) n_steps times; returns the last dti.  When `timed_ms` is not
// null the loop is bracketed by CUDA events on the default stream (the stream every reference launch uses).
double ref_advance(double *d_conserved, int nx, int ny, int nz, int n_ghost, double dx, double dy, double dz,
                   int *flags, int n_steps, double cfl, double density_floor, double U_floor, double *dts,
                   float *timed_ms)
{
  cudaEvent_t e0, e1;
  if (timed_ms) {
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
  }
  double dti = Calc_dt_GPU(d_conserved, nx, ny, nz, n_ghost, kFields, dx, dy, dz, gama);
  if (timed_ms) cudaEventRecord(e0, 0);
  for (int s = 0; s < n_steps; s++) {
    double dt = cfl / dti;
    ref_vl_step(d_conserved, nx, ny, nz, n_ghost, dx, dy, dz, dt, kFields, density_floor);
    Apply_Temperature_Floor(d_conserved, nx, ny, nz, n_ghost, kFields, U_floor);
#ifdef COOLING_GPU  // grid3D.cpp:512-522
    Cooling_Update(d_conserved, nx, ny, nz, n_ghost, kFields, dt, gama);
#endif
    dti = Calc_dt_GPU(d_conserved, nx, ny, nz, n_ghost, kFields, dx, dy, dz, gama);
    ref_set_boundaries(d_conserved, nx, ny, nz, kFields, n_ghost, flags);
    if (dts) dts[s] = dt;
  }
  if (timed_ms) {
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(timed_ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  cudaDeviceSynchronize();
  return dti;
}

// The reference's integrator owns its scratch through file-scope globals and Free_Memory_VL_* also frees the
// caller's array (VL_3D_cuda.cu:362-377), so this driver never calls it: one process = one grid shape.
int ref_last_cuda_error(void) { return (int)cudaGetLastError(); }
int ref_n_fields(void) { return kFields; }
}
