// cholla_shims.cpp -- the reference's own host entry points, re-implemented over the C ABI of libcholla_b200.so.
//
// Compile this ONE file into cholla (it replaces src/integrators/VL_{1,2,3}D_cuda.cu, src/hydro/hydro_cuda.cu's
// Calc_dt_GPU / Apply_Temperature_Floor and src/grid/cuda_boundaries.cu's SetGhostCells / Pack/UnpackBuffers3D in the
// link) with the build's usual DFLAGS, and link libcholla_b200.so: Grid3D::Execute_Hydro_Integrator
// (grid3D.cpp:431-465), Grid3D::Calc_Inverse_Timestep (:115-120), Grid3D::Update_Hydro_Grid (:493-503),
// Grid3D::Set_Boundaries (boundary_conditions.cpp:413) and the Load/Unload_Hydro_DeviceBuffer_* routines
// (mpi_boundaries.cpp:113-298) then call the B200 kernels without a line of the reference's host code changing.
//
// Signatures are the reference's, C++ linkage (src/integrators/VL_*D_cuda.h, src/hydro/hydro_cuda.h:69,87,
// src/grid/cuda_boundaries.h).  The compile-time feature macros of builds/make.type.* select the run-time enums:
//   -DPCM / -DPLMC / -DPLMP / -DPPMP / -DPPMC -> chl_config.recon      -DHLLC / -DROE / -DHLL / -DHLLD -> chl_config.solver
//   -DPRECISION=1|2         -> Real and chl_config.precision        -DMHD -> chl_config.mhd
// `gama` is the reference's global (src/global/global.cpp:18).  CHOLLA_B200_STRICT=0 in the environment selects the
// product arithmetic; the default is the bit-reproducing strict mode, so a drop-in build produces the snapshots the
// reference's own kernels would.
//
// Like the reference this layer keeps ONE lazily created context in file-scope state ("memory_allocated",
// src/global/global_cuda.cu:7-16); unlike the reference Free_Memory_VL_* does not free the caller's d_conserved
// (VL_3D_cuda.cu:365 does, and Grid3D::FreeMemory relies on it: see INTEGRATION.md).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/cholla_b200.h"

#ifndef PRECISION
#define PRECISION 2
#endif
#if PRECISION == 1
typedef float Real;
#else
typedef double Real;
#endif

extern Real gama;                  // src/global/global.h:147
[[noreturn]] void chexit(int code);  // src/utils/error_handling.h

namespace
{
#if defined(PPMC)
constexpr int kRecon = CHL_RECON_PPMC;
#elif defined(PLMC)
constexpr int kRecon = CHL_RECON_PLMC;
#elif defined(PLMP)
constexpr int kRecon = CHL_RECON_PLMP;
#elif defined(PPMP)
constexpr int kRecon = CHL_RECON_PPMP;
#elif defined(PCM)
constexpr int kRecon = CHL_RECON_PCM;
#else
#error "cholla_shims.cpp: build with -DPCM, -DPLMC, -DPLMP, -DPPMP or -DPPMC"
#endif
#if defined(HLLD)
constexpr int kSolver = CHL_SOLVER_HLLD;
#elif defined(ROE)
constexpr int kSolver = CHL_SOLVER_ROE;
#elif defined(HLLC)
constexpr int kSolver = CHL_SOLVER_HLLC;
#elif defined(HLL)
constexpr int kSolver = CHL_SOLVER_HLL;
#elif defined(EXACT)
constexpr int kSolver = CHL_SOLVER_EXACT;
#else
#error "cholla_shims.cpp: build with -DHLLC, -DROE, -DHLL, -DEXACT or -DHLLD"
#endif
#ifdef MHD
constexpr int kMhd = 1;
#else
constexpr int kMhd = 0;
#endif

chl_ctx *g_ctx = nullptr;
chl_config g_cfg;

void die(const char *where, int rc)
{  // the reference's error path: print, then chexit (src/utils/gpu.hpp:139-156)
  fprintf(stderr, "cholla_b200: %s failed (%d): %s\n", where, rc, chl_last_error());
  chexit(rc);
}

chl_ctx *context(int nx, int ny, int nz, int n_ghost, int n_fields, Real dx, Real dy, Real dz, Real density_floor)
{
  chl_config c;
  memset(&c, 0, sizeof(c));
  c.nx = nx; c.ny = ny; c.nz = nz; c.n_ghost = n_ghost; c.n_fields = n_fields;
  c.recon = kRecon; c.solver = kSolver; c.precision = PRECISION; c.mhd = kMhd;
  const char *s = getenv("CHOLLA_B200_STRICT");
  c.strict = (PRECISION == 2) && !(s && s[0] == '0');
  c.gamma = gama; c.dx = dx; c.dy = dy; c.dz = dz; c.density_floor = density_floor;
  c.temperature_floor_U = 0.0;
#ifdef DE  // -DDE: GasEnergy is the last field; -DSCALAR: whatever lies between Energy and it (grid_enum.h:33-78)
  c.dual_energy = 1;
#endif
  if (!kMhd) c.n_scalars = n_fields - 5 - c.dual_energy;
  if (g_ctx && memcmp(&c, &g_cfg, sizeof(c)) == 0) return g_ctx;
  if (g_ctx) chl_destroy(g_ctx);
  g_ctx = nullptr;
  int const rc = chl_create(&c, &g_ctx);
  if (rc != CHL_OK) die("chl_create", rc);
  g_cfg = c;
  return g_ctx;
}
// -DSTATIC_GRAV: the integrators receive Parameters::custom_grav, the rank's offsets and the domain corner on every call
// (grid3D.cpp:431-465); the library evaluates the field once per distinct set
void gravity(chl_ctx *c, int custom_grav, int x_off, int y_off, int z_off, Real xbound, Real ybound, Real zbound)
{
#ifdef STATIC_GRAV
  static chl_ctx *last_ctx = nullptr;
  static double last[7];
  double const now[7] = {(double)custom_grav, (double)x_off, (double)y_off, (double)z_off, (double)xbound, (double)ybound,
                         (double)zbound};
  if (c == last_ctx && memcmp(now, last, sizeof(now)) == 0) return;
  int const rc = chl_set_static_gravity(c, custom_grav, x_off, y_off, z_off, xbound, ybound, zbound, nullptr);
  if (rc != CHL_OK) die("chl_set_static_gravity", rc);
  last_ctx = c;
  memcpy(last, now, sizeof(now));
#else
  (void)c; (void)custom_grav; (void)x_off; (void)y_off; (void)z_off; (void)xbound; (void)ybound; (void)zbound;
#endif
}
}  // namespace

// src/cooling/cooling_cuda.h: Cooling_Update (a -DCOOLING_GPU build calls it after the integrator, grid3D.cpp:512-522)
void Cooling_Update(Real *dev_conserved, int nx, int ny, int nz, int n_ghost, int n_fields, Real dt, Real gamma)
{
  (void)gamma;  // the context carries the global `gama` the reference passes here
  if (!g_ctx || g_cfg.nx != nx || g_cfg.ny != ny || g_cfg.nz != nz || g_cfg.n_ghost != n_ghost || g_cfg.n_fields != n_fields)
    die("Cooling_Update before the first integrator call on this grid", CHL_ERR_INVALID);
  int const rc = chl_cooling_update(g_ctx, dev_conserved, dt, nullptr);
  if (rc != CHL_OK) die("chl_cooling_update", rc);
}

// src/integrators/VL_3D_cuda.h:9-11
void VL_Algorithm_3D_CUDA(Real *d_conserved, Real *d_grav_potential, int nx, int ny, int nz, int x_off, int y_off,
                          int z_off, int n_ghost, Real dx, Real dy, Real dz, Real xbound, Real ybound, Real zbound,
                          Real dt, int n_fields, int custom_grav, Real density_floor, Real *host_grav_potential)
{
  (void)host_grav_potential;
  chl_ctx *const c = context(nx, ny, nz, n_ghost, n_fields, dx, dy, dz, density_floor);
  gravity(c, custom_grav, x_off, y_off, z_off, xbound, ybound, zbound);
#ifdef GRAVITY  // the potential-gradient source terms of the update kernel; the Poisson solve stays the reference's
  #ifndef GRAVITY_GPU
    #error "cholla_shims.cpp: -DGRAVITY needs -DGRAVITY_GPU (the potential must already be in device memory)"
  #endif
  #ifdef GRAVITY_5_POINTS_GRADIENT
  int const rcp = chl_set_gravity_potential(c, d_grav_potential, 1);
  #else
  int const rcp = chl_set_gravity_potential(c, d_grav_potential, 0);
  #endif
  if (rcp != CHL_OK) die("chl_set_gravity_potential", rcp);
#else
  (void)d_grav_potential;
#endif
  int const rc = chl_vl_step(c, d_conserved, dt, nullptr);
  if (rc != CHL_OK) die("chl_vl_step", rc);
}

// src/integrators/VL_2D_cuda.h:9-11
void VL_Algorithm_2D_CUDA(Real *d_conserved, int nx, int ny, int x_off, int y_off, int n_ghost, Real dx, Real dy,
                          Real xbound, Real ybound, Real dt, int n_fields, int custom_grav)
{
  chl_ctx *const c = context(nx, ny, 1, n_ghost, n_fields, dx, dy, dy, 0.0);
  gravity(c, custom_grav, x_off, y_off, 0, xbound, ybound, 0.0);
  int const rc = chl_vl_step(c, d_conserved, dt, nullptr);
  if (rc != CHL_OK) die("chl_vl_step", rc);
}

// src/integrators/VL_1D_cuda.h
void VL_Algorithm_1D_CUDA(Real *d_conserved, int nx, int x_off, int n_ghost, Real dx, Real xbound, Real dt,
                          int n_fields, int custom_grav)
{
  chl_ctx *const c = context(nx, 1, 1, n_ghost, n_fields, dx, dx, dx, 0.0);
  gravity(c, custom_grav, x_off, 0, 0, xbound, 0.0, 0.0);
  int const rc = chl_vl_step(c, d_conserved, dt, nullptr);
  if (rc != CHL_OK) die("chl_vl_step", rc);
}

// src/integrators/simple_3D_cuda.h, simple_2D_cuda.h, simple_1D_cuda.h: a -DSIMPLE build (builds/make.type.hydro:8-12) calls
// these from Grid3D::Execute_Hydro_Integrator instead of the VL entry points; both sets are defined so one shim object
// serves either build
void Simple_Algorithm_3D_CUDA(Real *d_conserved, Real *d_grav_potential, int nx, int ny, int nz, int x_off, int y_off,
                              int z_off, int n_ghost, Real dx, Real dy, Real dz, Real xbound, Real ybound, Real zbound,
                              Real dt, int n_fields, int custom_grav, Real density_floor, Real *host_grav_potential)
{
  (void)d_grav_potential; (void)x_off; (void)y_off; (void)z_off; (void)xbound; (void)ybound; (void)zbound;
  (void)custom_grav; (void)host_grav_potential;
  int const rc = chl_simple_step(context(nx, ny, nz, n_ghost, n_fields, dx, dy, dz, density_floor), d_conserved, dt, nullptr);
  if (rc != CHL_OK) die("chl_simple_step", rc);
}

void Simple_Algorithm_2D_CUDA(Real *d_conserved, int nx, int ny, int x_off, int y_off, int n_ghost, Real dx, Real dy,
                              Real xbound, Real ybound, Real dt, int n_fields, int custom_grav)
{
  (void)x_off; (void)y_off; (void)xbound; (void)ybound; (void)custom_grav;
  int const rc = chl_simple_step(context(nx, ny, 1, n_ghost, n_fields, dx, dy, dy, 0.0), d_conserved, dt, nullptr);
  if (rc != CHL_OK) die("chl_simple_step", rc);
}

void Simple_Algorithm_1D_CUDA(Real *d_conserved, int nx, int x_off, int n_ghost, Real dx, Real xbound, Real dt,
                              int n_fields, int custom_grav)
{
  (void)x_off; (void)xbound; (void)custom_grav;
  int const rc = chl_simple_step(context(nx, 1, 1, n_ghost, n_fields, dx, dx, dx, 0.0), d_conserved, dt, nullptr);
  if (rc != CHL_OK) die("chl_simple_step", rc);
}

static void free_memory()
{
  if (g_ctx) chl_destroy(g_ctx);
  g_ctx = nullptr;
}
void Free_Memory_VL_1D() { free_memory(); }
void Free_Memory_VL_2D() { free_memory(); }
void Free_Memory_VL_3D() { free_memory(); }
void Free_Memory_Simple_1D() { free_memory(); }
void Free_Memory_Simple_2D() { free_memory(); }
void Free_Memory_Simple_3D() { free_memory(); }

// src/hydro/hydro_cuda.h:69
Real Calc_dt_GPU(Real *dev_conserved, int nx, int ny, int nz, int n_ghost, int n_fields, Real dx, Real dy, Real dz,
                 Real gamma)
{
  (void)gamma;  // the reference passes its global `gama`
  chl_ctx *c = g_ctx;
  if (!c || g_cfg.nx != nx || g_cfg.ny != ny || g_cfg.nz != nz || g_cfg.n_ghost != n_ghost)
    c = context(nx, ny, nz, n_ghost, n_fields, dx, dy, dz, g_ctx ? g_cfg.density_floor : 0.0);
  double dti = 0.0;
  int const rc = chl_calc_dti(c, dev_conserved, &dti, nullptr);
  if (rc != CHL_OK) die("chl_calc_dti", rc);
  return (Real)dti;
}

// src/hydro/hydro_cuda.h:87
void Apply_Temperature_Floor(Real *dev_conserved, int nx, int ny, int nz, int n_ghost, int n_fields, Real U_floor)
{
  chl_ctx *c = g_ctx;
  if (!c || g_cfg.nx != nx || g_cfg.ny != ny || g_cfg.nz != nz || g_cfg.n_ghost != n_ghost)
    c = context(nx, ny, nz, n_ghost, n_fields, 1.0, 1.0, 1.0, 0.0);
  int const rc = chl_apply_temperature_floor(c, dev_conserved, U_floor, nullptr);
  if (rc != CHL_OK) die("chl_apply_temperature_floor", rc);
}

// src/grid/cuda_boundaries.h
void SetGhostCells(Real *c_head, int nx, int ny, int nz, int n_fields, int n_cells, int n_ghost, int flags[], int isize,
                   int jsize, int ksize, int imin, int jmin, int kmin, int dir)
{
  int const rc = chl_set_ghost_cells_ext(c_head, nx, ny, nz, n_fields, n_cells, n_ghost, flags, isize, jsize, ksize,
                                         imin, jmin, kmin, dir, PRECISION, kMhd, nullptr);
  if (rc != CHL_OK) die("chl_set_ghost_cells_ext", rc);
}

void PackBuffers3D(Real *buffer, Real *c_head, int nx, int ny, int n_fields, int n_cells, int idxoffset, int isize,
                   int jsize, int ksize)
{
  int const rc = chl_pack_buffer(buffer, c_head, nx, ny, n_fields, n_cells, idxoffset, isize, jsize, ksize, PRECISION,
                                 nullptr);
  if (rc != CHL_OK) die("chl_pack_buffer", rc);
}

void UnpackBuffers3D(Real *buffer, Real *c_head, int nx, int ny, int n_fields, int n_cells, int idxoffset, int isize,
                     int jsize, int ksize)
{
  int const rc = chl_unpack_buffer(buffer, c_head, nx, ny, n_fields, n_cells, idxoffset, isize, jsize, ksize, PRECISION,
                                   nullptr);
  if (rc != CHL_OK) die("chl_unpack_buffer", rc);
}
