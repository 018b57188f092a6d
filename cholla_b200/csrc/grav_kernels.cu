// grav_kernels.cu -- static gravity (-DSTATIC_GRAV) for the Van Leer path.
//
// The reference evaluates an analytic acceleration inside its corrector update kernels, per cell and per step
// (src/gravity/static_grav.h:16-210 calc_g_{1,2,3}D, called from Update_Conserved_Variables_{1,2,3}D,
// src/hydro/hydro_cuda.cu:59-68, 141-153, 301-317), and adds the source terms time-centred between t^n and t^{n+1}.  The
// field does not change during a run, so here it is evaluated ONCE into three arrays in HBM (launch_static_gravity_field;
// a caller may also fill them with a field of its own), and a step costs one copy of (rho, momentum)^n before the corrector
// and one elementwise pass after it (launch_gravity_source) -- no transcendental arithmetic in the time loop.
//
// Built -fmad=false for every arithmetic mode: the expressions below keep the reference's order, and log / pow / atan2 /
// cos / sin are the same device library the reference's build links, so the reference's bits come out.
#include <math.h>

#include "chl_aux.h"

namespace chl
{
namespace
{
constexpr double kGN = 4.49451e-18;        // global.h:29 GN, kpc^3 / M_sun / kyr^2
constexpr double kMsunCgs = 1.98847e33;    // global.h:35 MSUN_CGS

template <typename T>
__device__ __forceinline__ T sq(T x) { return x * x; }

// NFW halo + Miyamoto-Nagai disk of calc_g_3D cases 1 (Milky Way) and 2 (M82), static_grav.h:142-199
template <typename T>
__device__ __forceinline__ void disk_halo_3d(T x_pos, T y_pos, T z_pos, T M_vir, T M_d, T R_d, T z_d, T R_vir, T c_vir,
                                             T &gx, T &gy, T &gz)
{
  T const r_disk = sqrt(x_pos * x_pos + y_pos * y_pos);
  T const r_halo = sqrt(x_pos * x_pos + y_pos * y_pos + z_pos * z_pos);
  T const M_h = M_vir - M_d, R_h = R_vir / c_vir;
  T const phi_0_h = (T)kGN * M_h / (log((T)1.0 + c_vir) - c_vir / ((T)1.0 + c_vir));
  T const x = r_halo / R_h;
  T const a_halo   = -phi_0_h * (log(1 + x) - x / (1 + x)) / (r_halo * r_halo);
  T const a_halo_r = a_halo * (r_disk / r_halo);
  T const a_halo_z = a_halo * (z_pos / r_halo);
  T const a_disk_r = -(T)kGN * M_d * r_disk * pow(r_disk * r_disk + sq(R_d + sqrt(z_pos * z_pos + z_d * z_d)), (T)-1.5);
  T const a_disk_z = -(T)kGN * M_d * z_pos * (R_d + sqrt(z_pos * z_pos + z_d * z_d)) /
                     (pow(r_disk * r_disk + sq(R_d + sqrt(z_pos * z_pos + z_d * z_d)), (T)1.5) *
                      sqrt(z_pos * z_pos + z_d * z_d));
  gx = (x_pos / r_disk) * (a_disk_r + a_halo_r);
  gy = (y_pos / r_disk) * (a_disk_r + a_halo_r);
  gz = a_disk_z + a_halo_z;
}

struct GravGeom {
  int nx, ny, nz, g, custom_grav, off[3];
  double d[3], bound[3];
};

// calc_g_1D / calc_g_2D / calc_g_3D for every cell of the local array (the update only reads the real ones)
template <typename T, int NDIM>
__global__ void __launch_bounds__(256) static_gravity_kernel(T *__restrict__ out, GravGeom G)
{
  long const N = (long)G.nx * G.ny * G.nz;
  for (long id = blockIdx.x * (long)blockDim.x + threadIdx.x; id < N; id += (long)gridDim.x * blockDim.x) {
    int const xid = (int)(id % G.nx), yid = (int)((id / G.nx) % G.ny), zid = (int)(id / ((long)G.nx * G.ny));
    T const dx = (T)G.d[0], dy = (T)G.d[1], dz = (T)G.d[2];
    T gx = 0, gy = 0, gz = 0;
    T const x_pos = (G.off[0] + xid - G.g + (T)0.5) * dx + (T)G.bound[0];
    if (NDIM == 1) {
      if (G.custom_grav == 1) {  // 1-D NFW halo + Miyamoto-Nagai disk column, static_grav.h:22-55
        T const r_disk = (T)13.9211647546;
        T const r_halo = sqrt(x_pos * x_pos + r_disk * r_disk);
        T const M_vir = (T)1.0e12, M_d = (T)6.5e10, M_h = M_vir - M_d, R_vir = 261, c_vir = (T)20.0, R_h = R_vir / c_vir;
        T const R_d = (T)3.5, z_d = (T)3.5 / (T)5.0;
        T const phi_0_h = (T)kGN * M_h / (log((T)1.0 + c_vir) - c_vir / ((T)1.0 + c_vir));
        T const x = r_halo / R_h;
        T const a_halo = -phi_0_h * (log(1 + x) - x / (1 + x)) / (r_halo * r_halo);
        T const a_disk_z = -(T)kGN * M_d * x_pos * (R_d + sqrt(x_pos * x_pos + z_d * z_d)) /
                           (pow(r_disk * r_disk + sq(R_d + sqrt(x_pos * x_pos + z_d * z_d)), (T)1.5) *
                            sqrt(x_pos * x_pos + z_d * z_d));
        gx = (x_pos / r_halo) * a_halo + a_disk_z;
      }
    } else if (NDIM == 2) {
      T const y_pos = (G.off[1] + yid - G.g + (T)0.5) * dy + (T)G.bound[1];
      T const r = sqrt(x_pos * x_pos + y_pos * y_pos), phi = atan2(y_pos, x_pos);
      switch (G.custom_grav) {
        case 1:  // Gresho vortex: balances v_phi
          if (r < (T)0.2) {
            gx = -cos(phi) * (T)25.0 * r;
            gy = -sin(phi) * (T)25.0 * r;
          } else if (r >= (T)0.2 && r < (T)0.4) {
            gx = -cos(phi) * ((T)4.0 - (T)20.0 * r + (T)25.0 * r * r) / r;
            gy = -sin(phi) * ((T)4.0 - (T)20.0 * r + (T)25.0 * r * r) / r;
          }
          break;
        case 2:  // Rayleigh-Taylor
          gx = 0;
          gy = -1;
          break;
        case 3: {  // Keplerian disk around one solar mass
          T const M = 1 * (T)kMsunCgs;
          gx = -cos(phi) * (T)kGN * M / (r * r);
          gy = -sin(phi) * (T)kGN * M / (r * r);
          break;
        }
        case 4: {  // Kuzmin disk + NFW halo
          T const M_vir = (T)1.0e12, M_d = (T)6.5e10, M_h = M_vir - M_d, R_vir = 261, c_vir = 20, R_s = R_vir / c_vir;
          T const R_d = (T)3.5;
          T const x   = r / R_s;
          T const a_d = (T)kGN * M_d * r * pow(r * r + R_d * R_d, (T)-1.5);
          T const a_h = (T)kGN * M_h * (log(1 + x) - x / (1 + x)) / ((log(1 + c_vir) - c_vir / (1 + c_vir)) * r * r);
          T const a   = a_d + a_h;
          gx = -cos(phi) * a;
          gy = -sin(phi) * a;
          break;
        }
        default: break;
      }
    } else {
      T const y_pos = (G.off[1] + yid - G.g + (T)0.5) * dy + (T)G.bound[1];
      T const z_pos = (G.off[2] + zid - G.g + (T)0.5) * dz + (T)G.bound[2];
      if (G.custom_grav == 1)  // Milky Way
        disk_halo_3d<T>(x_pos, y_pos, z_pos, (T)1.0e12, (T)6.5e10, (T)3.5, (T)3.5 / (T)5.0, (T)261., (T)20.0, gx, gy, gz);
      else if (G.custom_grav == 2)  // M82
        disk_halo_3d<T>(x_pos, y_pos, z_pos, (T)5.0e10, (T)1.0e10, (T)0.8, (T)0.15, (T)0.8 / (T)0.015, (T)10.0, gx, gy, gz);
    }
    out[id] = gx;
    if (NDIM > 1) out[N + id] = gy;
    if (NDIM > 2) out[2 * N + id] = gz;
  }
}

// The source terms of Update_Conserved_Variables_{1,2,3}D, real cells: `old` holds (rho, momentum)^n, U the state after the
// flux update (and the density floor), g the acceleration.
template <typename T, int NDIM>
__global__ void __launch_bounds__(256) gravity_source_kernel(T *__restrict__ U, const T *__restrict__ old,
                                                             const T *__restrict__ g, int nx, int ny, int nz, int ng, T dt)
{
  long const N = (long)nx * ny * nz;
  int const rx = nx - 2 * ng, ry = NDIM > 1 ? ny - 2 * ng : 1, rz = NDIM > 2 ? nz - 2 * ng : 1;
  long const n = (long)rx * ry * rz;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
    int const x = (int)(t % rx) + ng;
    long const r = t / rx;
    int const y = NDIM > 1 ? (int)(r % ry) + ng : 0, z = NDIM > 2 ? (int)(r / ry) + ng : 0;
    long const id = x + (long)nx * (y + (long)ny * z);
    T const d = old[id], d_inv = (T)1.0 / d;
    T const d_n = U[id], d_inv_n = (T)1.0 / d_n;
    T const gx = g[id];
    T const vx = old[N + id] * d_inv, vx_n = U[N + id] * d_inv_n;
    T dE = (T)0.25 * dt * gx * (d + d_n) * (vx + vx_n);
    U[N + id] += (T)0.5 * dt * gx * (d + d_n);
    if (NDIM > 1) {
      T const gy = g[N + id];
      T const vy = old[2 * N + id] * d_inv, vy_n = U[2 * N + id] * d_inv_n;
      dE = dE + (T)0.25 * dt * gy * (d + d_n) * (vy + vy_n);
      U[2 * N + id] += (T)0.5 * dt * gy * (d + d_n);
    }
    if (NDIM > 2) {
      T const gz = g[2 * N + id];
      T const vz = old[3 * N + id] * d_inv, vz_n = U[3 * N + id] * d_inv_n;
      dE = dE + (T)0.25 * dt * gz * (d + d_n) * (vz + vz_n);
      U[3 * N + id] += (T)0.5 * dt * gz * (d + d_n);
    }
    U[4 * N + id] += dE;
  }
}

// The -DGRAVITY branch of Update_Conserved_Variables_3D (hydro_cuda.cu:317-379): acceleration from the centred gradient of
// the potential the caller's Poisson solver left in device memory (3-point, or -DGRAVITY_5_POINTS_GRADIENT -- whose y and
// z components divide by 12 dx like the x one, reproduced here), momentum sources 0.5 dt g (rho^n + rho^{n+1}) and the work
// 0.5 dt g . (rho v^n + rho v^{n+1}).  3-D only, like the reference.
template <typename T>
__global__ void __launch_bounds__(256) gravity_potential_source_kernel(T *__restrict__ U, const T *__restrict__ old,
                                                                       const T *__restrict__ pot, int nx, int ny, int nz,
                                                                       int ng, T dt, T dx, T dy, T dz, int five_point)
{
  long const N = (long)nx * ny * nz;
  int const rx = nx - 2 * ng, ry = ny - 2 * ng, rz = nz - 2 * ng;
  long const n = (long)rx * ry * rz, sy = nx, sz = (long)nx * ny;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
    int const x = (int)(t % rx) + ng;
    long const r = t / rx;
    int const y = (int)(r % ry) + ng, z = (int)(r / ry) + ng;
    long const id = x + (long)nx * (y + (long)ny * z);
    T const d = old[id], d_inv = (T)1.0 / d;
    T const vx = old[N + id] * d_inv, vy = old[2 * N + id] * d_inv, vz = old[3 * N + id] * d_inv;
    T const d_n = U[id], d_inv_n = (T)1.0 / d_n;
    T const vx_n = U[N + id] * d_inv_n, vy_n = U[2 * N + id] * d_inv_n, vz_n = U[3 * N + id] * d_inv_n;
    T gx, gy, gz;
    if (five_point) {
      gx = -1 * (-pot[id + 2] + 8 * pot[id + 1] - 8 * pot[id - 1] + pot[id - 2]) / (12 * dx);
      gy = -1 * (-pot[id + 2 * sy] + 8 * pot[id + sy] - 8 * pot[id - sy] + pot[id - 2 * sy]) / (12 * dx);
      gz = -1 * (-pot[id + 2 * sz] + 8 * pot[id + sz] - 8 * pot[id - sz] + pot[id - 2 * sz]) / (12 * dx);
    } else {
      gx = (T)-0.5 * (pot[id + 1] - pot[id - 1]) / dx;
      gy = (T)-0.5 * (pot[id + sy] - pot[id - sy]) / dy;
      gz = (T)-0.5 * (pot[id + sz] - pot[id - sz]) / dz;
    }
    U[N + id] += (T)0.5 * dt * gx * (d + d_n);
    U[2 * N + id] += (T)0.5 * dt * gy * (d + d_n);
    U[3 * N + id] += (T)0.5 * dt * gz * (d + d_n);
    U[4 * N + id] += (T)0.5 * dt * (gx * (d * vx + d_n * vx_n) + gy * (d * vy + d_n * vy_n) + gz * (d * vz + d_n * vz_n));
  }
}

int grid_for(long n)
{
  long b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  return (int)(b < 1 ? 1 : b);
}
}  // namespace

cudaError_t launch_static_gravity_field(void *g_out, int nx, int ny, int nz, int n_ghost, int custom_grav, const int off[3],
                                        const double d[3], const double bound[3], int real_bytes, cudaStream_t s)
{
  GravGeom G;
  G.nx = nx; G.ny = ny; G.nz = nz; G.g = n_ghost; G.custom_grav = custom_grav;
  for (int i = 0; i < 3; i++) { G.off[i] = off[i]; G.d[i] = d[i]; G.bound[i] = bound[i]; }
  int const ndim = 1 + (ny > 1) + (nz > 1);
  int const b = grid_for((long)nx * ny * nz);
#define CHL_G(T, ND) static_gravity_kernel<T, ND><<<b, 256, 0, s>>>((T *)g_out, G)
  if (real_bytes == 8) {
    if (ndim == 3) CHL_G(double, 3); else if (ndim == 2) CHL_G(double, 2); else CHL_G(double, 1);
  } else {
    if (ndim == 3) CHL_G(float, 3); else if (ndim == 2) CHL_G(float, 2); else CHL_G(float, 1);
  }
#undef CHL_G
  return cudaGetLastError();
}

cudaError_t launch_gravity_source(void *U, const void *old, const void *g, int nx, int ny, int nz, int n_ghost, double dt,
                                  int real_bytes, cudaStream_t s)
{
  int const ndim = 1 + (ny > 1) + (nz > 1);
  long const n = (long)(nx - 2 * n_ghost) * (ndim > 1 ? ny - 2 * n_ghost : 1) * (ndim > 2 ? nz - 2 * n_ghost : 1);
  int const b = grid_for(n);
#define CHL_S(T, ND) \
  gravity_source_kernel<T, ND><<<b, 256, 0, s>>>((T *)U, (const T *)old, (const T *)g, nx, ny, nz, n_ghost, (T)dt)
  if (real_bytes == 8) {
    if (ndim == 3) CHL_S(double, 3); else if (ndim == 2) CHL_S(double, 2); else CHL_S(double, 1);
  } else {
    if (ndim == 3) CHL_S(float, 3); else if (ndim == 2) CHL_S(float, 2); else CHL_S(float, 1);
  }
#undef CHL_S
  return cudaGetLastError();
}
cudaError_t launch_gravity_potential_source(void *U, const void *old, const void *pot, int nx, int ny, int nz, int n_ghost,
                                            double dt, double dx, double dy, double dz, int five_point, int real_bytes,
                                            cudaStream_t s)
{
  long const n = (long)(nx - 2 * n_ghost) * (ny - 2 * n_ghost) * (nz - 2 * n_ghost);
  int const b = grid_for(n);
  if (real_bytes == 8)
    gravity_potential_source_kernel<double><<<b, 256, 0, s>>>((double *)U, (const double *)old, (const double *)pot, nx, ny,
                                                             nz, n_ghost, dt, dx, dy, dz, five_point);
  else
    gravity_potential_source_kernel<float><<<b, 256, 0, s>>>((float *)U, (const float *)old, (const float *)pot, nx, ny, nz,
                                                            n_ghost, (float)dt, (float)dx, (float)dy, (float)dz, five_point);
  return cudaGetLastError();
}
}  // namespace chl
