// aux_kernels.cu -- ghost-cell fill and halo pack/unpack: pure copies, HBM-bound, element type agnostic.
// Behaviour follows src/grid/cuda_boundaries.cu:11-303 of the reference (SetGhostCellsKernel,
// SetBoundaryMapping, FindIndex, Pack/UnpackBuffers3DKernel); the code is new.
#include "chl_aux.h"

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

// mhd::calculateMagneticDivergence (src/mhd/magnetic_divergence.cu:33-76): max over cells with x, y, z > 1 (ghost cells
// whose low faces exist included) of |(Bx - Bx[i-1]) / dx + (By - By[j-1]) / dy + (Bz - Bz[k-1]) / dz|; the reduction is a
// warp shuffle + one atomicMax per CTA on the bit pattern (non-negative doubles order like their bits)
template <typename T>
__global__ void __launch_bounds__(256) divb_kernel(const T *__restrict__ c, int nx, int ny, int nz, long n_cells, T dx, T dy,
                                                   T dz, unsigned long long *out)
{
  double m = 0.0;
  for (long id = blockIdx.x * (long)blockDim.x + threadIdx.x; id < n_cells; id += (long)gridDim.x * blockDim.x) {
    int const x = (int)(id % nx);
    long const r = id / nx;
    int const y = (int)(r % ny), z = (int)(r / ny);
    if (x > 1 && y > 1 && z > 1) {
      T const d = ((c[id + 5 * n_cells] - c[id - 1 + 5 * n_cells]) / dx) + ((c[id + 6 * n_cells] - c[id - nx + 6 * n_cells]) / dy) +
                  ((c[id + 7 * n_cells] - c[id - (long)nx * ny + 7 * n_cells]) / dz);
      m = fmax(m, fabs((double)d));
    }
  }
  __shared__ double sred[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) m = fmax(m, sred[w]);
    atomicMax(out, (unsigned long long)__double_as_longlong(m));  // NaN sorts above every number: it is reported
  }
}

inline int grid_for(long n)
{
  long b = (n + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

// Snapshot packing (the reference's CopyReal3D_GPU_Kernel / CopyReal2D_GPU_Kernel, src/io/io_gpu.cu:21-95): the real
// cells of one field, x fastest with ghosts, to the HDF5 buffer order out[k + n_k (j + n_j i)] -- x slowest.  A
// transposition of the fastest axis, so a 32 x 32 shared-memory tile keeps both the loads (along i) and the stores
// (along k) coalesced; the reference's kernel stores with a stride of n_j n_k elements.
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256) snapshot_pack_kernel(const Tin *__restrict__ in, Tout *__restrict__ out, int n_i,
                                                            int n_j, int n_k, long stride_j, long stride_k)
{
  __shared__ Tin tile[32][33];
  int const j = blockIdx.z, i0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    int const i = i0 + threadIdx.x, k = k0 + r;
    if (i < n_i && k < n_k) tile[r][threadIdx.x] = in[i + stride_j * j + stride_k * k];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    int const i = i0 + r, k = k0 + threadIdx.x;
    if (i < n_i && k < n_k) out[k + (long)n_k * (j + (long)n_j * i)] = (Tout)tile[threadIdx.x][r];
  }
}
// Scalar_Floor_Kernel (hydro_cuda.cu:1284-1307): one field, real cells of a 3-D grid
template <typename T>
__global__ void __launch_bounds__(256) scalar_floor_kernel(T *field, int nx, int ny, int nz, int g, T floor_value)
{
  int const rx = nx - 2 * g, ry = ny - 2 * g, rz = nz - 2 * g;
  long const n = (long)rx * ry * rz;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
    int const x = (int)(t % rx) + g;
    long const r = t / rx;
    int const y = (int)(r % ry) + g, z = (int)(r / ry) + g;
    long const id = x + (long)nx * (y + (long)ny * z);
    if (field[id] < floor_value) field[id] = floor_value;
  }
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

cudaError_t launch_scalar_floor(void *c, int nx, int ny, int nz, int n_ghost, int field_num, double floor_value,
                                int real_bytes, cudaStream_t s)
{
  if (nx <= 2 * n_ghost || ny <= 2 * n_ghost || nz <= 2 * n_ghost) return cudaSuccess;  // the kernel's guard
  long const n = (long)(nx - 2 * n_ghost) * (ny - 2 * n_ghost) * (nz - 2 * n_ghost);
  long const N = (long)nx * ny * nz;
  int const g = grid_for(n) > 148 * 8 ? 148 * 8 : grid_for(n);
  if (real_bytes == 8) scalar_floor_kernel<double><<<g, 256, 0, s>>>((double *)c + field_num * N, nx, ny, nz, n_ghost, floor_value);
  else scalar_floor_kernel<float><<<g, 256, 0, s>>>((float *)c + field_num * N, nx, ny, nz, n_ghost, (float)floor_value);
  return cudaGetLastError();
}
cudaError_t launch_divb(const void *c, int nx, int ny, int nz, double dx, double dy, double dz, unsigned long long *out,
                        size_t real_bytes, cudaStream_t s)
{
  long const n = (long)nx * ny * nz;
  int const g = grid_for(n) > 148 * 8 ? 148 * 8 : grid_for(n);
  if (real_bytes == 8) divb_kernel<double><<<g, 256, 0, s>>>((const double *)c, nx, ny, nz, n, dx, dy, dz, out);
  else divb_kernel<float><<<g, 256, 0, s>>>((const float *)c, nx, ny, nz, n, (float)dx, (float)dy, (float)dz, out);
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

cudaError_t launch_snapshot_pack(void *out, const void *field, int nx, int ny, int nz, int n_ghost, int nx_out,
                                 int ny_out, int nz_out, int mhd_direction, size_t in_bytes, size_t out_bytes,
                                 cudaStream_t s)
{
  // first source cell of the output (io_gpu.cu:62-66: the face-centred fields start one face below the real cells)
  long const gx = n_ghost - (mhd_direction == 0), gy = ny > 1 ? n_ghost - (mhd_direction == 1) : 0,
             gz = nz > 1 ? n_ghost - (mhd_direction == 2) : 0;
  long const base = gx + (long)nx * (gy + (long)ny * gz);
  // the fastest output axis is z in 3-D, y in 2-D; a 1-D grid is a straight copy (n_k = 1)
  int n_i = nx_out, n_j, n_k;
  long sj, sk;
  if (nz > 1) { n_j = ny_out; n_k = nz_out; sj = nx; sk = (long)nx * ny; }
  else { n_j = 1; n_k = ny > 1 ? ny_out : 1; sj = 0; sk = nx; }
  if (n_i <= 0 || n_j <= 0 || n_k <= 0) return cudaSuccess;
  if (n_j > 65535) return cudaErrorInvalidValue;
  dim3 const grid((n_i + 31) / 32, (n_k + 31) / 32, n_j), block(32, 8);
  if (in_bytes == 8 && out_bytes == 8)
    snapshot_pack_kernel<double, double><<<grid, block, 0, s>>>((const double *)field + base, (double *)out, n_i, n_j, n_k, sj, sk);
  else if (in_bytes == 8)
    snapshot_pack_kernel<double, float><<<grid, block, 0, s>>>((const double *)field + base, (float *)out, n_i, n_j, n_k, sj, sk);
  else if (out_bytes == 8)
    snapshot_pack_kernel<float, double><<<grid, block, 0, s>>>((const float *)field + base, (double *)out, n_i, n_j, n_k, sj, sk);
  else
    snapshot_pack_kernel<float, float><<<grid, block, 0, s>>>((const float *)field + base, (float *)out, n_i, n_j, n_k, sj, sk);
  return cudaGetLastError();
}
}  // namespace chl
