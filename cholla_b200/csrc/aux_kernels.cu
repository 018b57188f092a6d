// aux_kernels.cu -- ghost-cell fill and halo pack/unpack: pure copies, HBM-bound, element type agnostic.
// Behaviour follows src/grid/cuda_boundaries.cu:11-303 of the reference (SetGhostCellsKernel,
// SetBoundaryMapping, FindIndex, Pack/UnpackBuffers3DKernel); the code is new.
#include "chl_internal.h"

namespace chl
{
namespace
{
struct GhostArgs {
  int nx, ny, nz, n_fields, g;
  long n_cells;
  int flags[6];
  int isize, jsize, ksize, imin, jmin, kmin, dir;
  int diode;  // hydro builds only (cuda_boundaries.cu:121-147 is inside #ifndef MHD)
};

// index of the source cell along one axis for ghost index `c` (FindIndex); sets refl when mirrored
__device__ __forceinline__ int map_axis(int c, int n, int g, int flag_lo, int flag_hi, bool &refl)
{
  if (c < g) {
    switch (flag_lo) {
      case 2: refl = true; return 2 * g - c - 1;
      case 3: return g;
      case 4: return -1;
      case 5: return c;
      default: return c + n - 2 * g;
    }
  }
  if (c >= n - g) {
    switch (flag_hi) {
      case 2: refl = true; return 2 * (n - g) - c - 1;
      case 3: return n - g - 1;
      case 4: return -1;
      case 5: return c;
      default: return c - n + 2 * g;
    }
  }
  return c;
}

template <typename T>
__global__ void __launch_bounds__(256) ghost_kernel(T *__restrict__ c, GhostArgs a)
{
  long const total = (long)a.isize * a.jsize * a.ksize;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    int const i = (int)(t % a.isize) + a.imin;
    long const r = t / a.isize;
    int const j = (int)(r % a.jsize) + a.jmin, k = (int)(r / a.jsize) + a.kmin;
    long const gidx = i + (long)a.nx * (j + (long)a.ny * k);
    bool rx = false, ry = false, rz = false;
    long idx = 0;
    if (a.nx > 1) {
      int const s = map_axis(i, a.nx, a.g, a.flags[0], a.flags[1], rx);
      if (s < 0) continue;
      idx += s;
    }
    if (a.ny > 1) {
      int const s = map_axis(j, a.ny, a.g, a.flags[2], a.flags[3], ry);
      if (s < 0) continue;
      idx += (long)a.nx * s;
    }
    if (a.nz > 1) {
      int const s = map_axis(k, a.nz, a.g, a.flags[4], a.flags[5], rz);
      if (s < 0) continue;
      idx += (long)a.nx * a.ny * s;
    }
    T d = c[idx], m1 = c[idx + a.n_cells], m2 = c[idx + 2 * a.n_cells], m3 = c[idx + 3 * a.n_cells],
      e = c[idx + 4 * a.n_cells];
    if (rx && (a.flags[0] == 2 || a.flags[1] == 2)) m1 = -m1;
    if (ry && (a.flags[2] == 2 || a.flags[3] == 2)) m2 = -m2;
    if (rz && (a.flags[4] == 2 || a.flags[5] == 2)) m3 = -m3;
    if (a.flags[a.dir] == 3 && a.diode) {  // transmissive "diode": no inflow through this face
      int const ax = a.dir / 2;
      T &mn = ax == 0 ? m1 : (ax == 1 ? m2 : m3);
      bool const inflow = (a.dir % 2 == 0) ? (mn > (T)0.0) : (mn < (T)0.0);
      if (inflow) {
        e -= (T)0.5 * (mn * mn) / d;
        mn = (T)0.0;
      }
    }
    c[gidx] = d; c[gidx + a.n_cells] = m1; c[gidx + 2 * a.n_cells] = m2; c[gidx + 3 * a.n_cells] = m3;
    c[gidx + 4 * a.n_cells] = e;
    for (int f = 5; f < a.n_fields; f++) c[gidx + f * a.n_cells] = c[idx + f * a.n_cells];
  }
}

template <typename T, int UNPACK>
__global__ void __launch_bounds__(256) pack_kernel(T *__restrict__ buf, T *__restrict__ c, int nx, int ny, int n_fields,
                                                   long n_cells, long idxoffset, int isize, int jsize, int ksize)
{
  long const nb = (long)isize * jsize * ksize;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < nb; t += (long)gridDim.x * blockDim.x) {
    int const i = (int)(t % isize);
    long const r = t / isize;
    int const j = (int)(r % jsize), k = (int)(r / jsize);
    long const idx = i + ((long)j + (long)k * ny) * nx + idxoffset;
    for (int f = 0; f < n_fields; f++) {
      if (UNPACK) c[idx + f * n_cells] = buf[t + f * nb];
      else buf[t + f * nb] = c[idx + f * n_cells];
    }
  }
}

inline int grid_for(long n)
{
  long b = (n + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}
}  // namespace

cudaError_t launch_set_ghost_cells(void *c, int nx, int ny, int nz, int n_fields, long n_cells, int n_ghost,
                                   const int flags[6], int isize, int jsize, int ksize, int imin, int jmin, int kmin,
                                   int dir, int diode, size_t real_bytes, cudaStream_t s)
{
  GhostArgs a;
  a.diode = diode;
  a.nx = nx; a.ny = ny; a.nz = nz; a.n_fields = n_fields; a.g = n_ghost; a.n_cells = n_cells;
  for (int i = 0; i < 6; i++) a.flags[i] = flags[i];
  a.isize = isize; a.jsize = jsize; a.ksize = ksize; a.imin = imin; a.jmin = jmin; a.kmin = kmin; a.dir = dir;
  long const total = (long)isize * jsize * ksize;
  if (total <= 0) return cudaSuccess;
  if (real_bytes == 8) ghost_kernel<double><<<grid_for(total), 256, 0, s>>>((double *)c, a);
  else ghost_kernel<float><<<grid_for(total), 256, 0, s>>>((float *)c, a);
  return cudaGetLastError();
}

cudaError_t launch_pack(void *buf, const void *c, int nx, int ny, int n_fields, long n_cells, long idxoffset,
                        int isize, int jsize, int ksize, int unpack, size_t real_bytes, cudaStream_t s)
{
  long const nb = (long)isize * jsize * ksize;
  if (nb <= 0) return cudaSuccess;
  int const g = grid_for(nb);
  if (real_bytes == 8) {
    if (unpack) pack_kernel<double, 1><<<g, 256, 0, s>>>((double *)buf, (double *)c, nx, ny, n_fields, n_cells, idxoffset, isize, jsize, ksize);
    else pack_kernel<double, 0><<<g, 256, 0, s>>>((double *)buf, (double *)c, nx, ny, n_fields, n_cells, idxoffset, isize, jsize, ksize);
  } else {
    if (unpack) pack_kernel<float, 1><<<g, 256, 0, s>>>((float *)buf, (float *)c, nx, ny, n_fields, n_cells, idxoffset, isize, jsize, ksize);
    else pack_kernel<float, 0><<<g, 256, 0, s>>>((float *)buf, (float *)c, nx, ny, n_fields, n_cells, idxoffset, isize, jsize, ksize);
  }
  return cudaGetLastError();
}
}  // namespace chl
