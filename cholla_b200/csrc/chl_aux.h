// chl_aux.h -- launchers of the element-type-agnostic kernels (ghost cells, pack / unpack, snapshot transposition, halo
// exchange, arithmetic diagnostics) for the C-ABI TU.  Kept apart from chl_internal.h so that the large per-mode sweep TUs do
// not depend on it.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace chl
{
// element-type-templated copy kernels (ghost cells, pack/unpack), aux_kernels.cu
cudaError_t launch_set_ghost_cells(void *c, int nx, int ny, int nz, int n_fields, long n_cells, int n_ghost,
                                   const int flags[6], int isize, int jsize, int ksize, int imin, int jmin, int kmin,
                                   int dir, int diode, size_t real_bytes, cudaStream_t s);
cudaError_t launch_divb(const void *c, int nx, int ny, int nz, double dx, double dy, double dz, unsigned long long *out,
                        size_t real_bytes, cudaStream_t s);
// grav_kernels.cu: static gravity (-DSTATIC_GRAV): the acceleration field, once; the corrector's source terms, per step
cudaError_t launch_static_gravity_field(void *g_out, int nx, int ny, int nz, int n_ghost, int custom_grav, const int off[3],
                                        const double d[3], const double bound[3], int real_bytes, cudaStream_t s);
cudaError_t launch_gravity_potential_source(void *U, const void *old, const void *pot, int nx, int ny, int nz, int n_ghost,
                                            double dt, double dx, double dy, double dz, int five_point, int real_bytes,
                                            cudaStream_t s);
cudaError_t launch_gravity_source(void *U, const void *old, const void *g, int nx, int ny, int nz, int n_ghost, double dt,
                                  int real_bytes, cudaStream_t s);
// cooling_kernels.cu: Cooling_Update with the analytic CIE curve (-DCOOLING_GPU without -DCLOUDY_COOL)
cudaError_t launch_cooling_update(void *U, int nx, int ny, int nz, int n_ghost, int n_fields, int dual_energy, double dt,
                                  double gamma, int real_bytes, cudaStream_t s);
cudaError_t launch_scalar_floor(void *c, int nx, int ny, int nz, int n_ghost, int field_num, double floor_value,
                                int real_bytes, cudaStream_t s);
cudaError_t launch_pack(void *buf, const void *c, int nx, int ny, int n_fields, long n_cells, long idxoffset,
                        int isize, int jsize, int ksize, int unpack, size_t real_bytes, cudaStream_t s);
cudaError_t launch_snapshot_pack(void *out, const void *field, int nx, int ny, int nz, int n_ghost, int nx_out,
                                 int ny_out, int nz_out, int mhd_direction, size_t in_bytes, size_t out_bytes,
                                 cudaStream_t s);
// halo_p2p.cu: slab -> the neighbour's peer-mapped receive buffer (+ arrival word), and wait + buffer -> ghost slab
cudaError_t launch_halo_send(void *remote, const void *c, int nx, int ny, int n_fields, long n_cells, long idxoffset,
                             int isize, int jsize, int ksize, unsigned *counter, unsigned long long *remote_flag,
                             unsigned long long seq, size_t real_bytes, cudaStream_t s);
cudaError_t launch_halo_recv(const void *buf, void *c, int nx, int ny, int n_fields, long n_cells, long idxoffset, int isize,
                             int jsize, int ksize, const unsigned long long *flag, unsigned long long seq, unsigned *err,
                             size_t real_bytes, cudaStream_t s);
// exact_div_test.cu: shared-reciprocal division vs the compiler's a / b
cudaError_t launch_exact_div_test(const double *a, const double *b, double *q_shared, double *q_plain, long n,
                                  cudaStream_t s);
// the branch-free form: the raw quotient and sqrt plus the sticky flag each raised (flag bit 0: quotient, bit 1: sqrt of a)
cudaError_t launch_exact_div_flag_test(const double *a, const double *b, double *q, int *flag, int zero_ok, long n,
                                       cudaStream_t s);
cudaError_t launch_exact_sqrt_test(const double *x, double *r_own, double *r_plain, long n, cudaStream_t s);
}  // namespace chl
