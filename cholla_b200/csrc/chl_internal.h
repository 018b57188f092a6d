// chl_internal.h -- glue between the mode-agnostic C-ABI TU (chl_api.cu) and the per-mode kernel TUs
// (vl_sweep_inst.cu compiled once per arithmetic mode: f64 strict, f64 fast, f32 fast).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace chl
{
struct SweepArgs {  // everything in double; each mode TU narrows to its Real
  const void *in;
  void *out;
  void *acc;  // direction-split sweeps: running sum of the flux differences, 5 * n_cells reals
  int nx, ny, nz;
  int lo[3], hi[3];
  int has_y, has_z;
  int zchunk;
  int dir_chunk;  // direction-split sweeps: cells per thread along the marching direction (0 = choose)
  int simple;     // 1: SIMPLE integrator -- stage 1 reconstructs with the -DSIMPLE forms (PLMC gains its tracing)
  double dtodx, dtody, dtodz, gamma, density_floor;
  int apply_tfloor;
  double U_floor;
  int do_dti;
  double dx, dy, dz;
  unsigned long long *dti_key;
};

struct ModeOps {
  // stage 0 = predictor (U -> W half), stage 1 = corrector (W half, U -> U)
  cudaError_t (*sweep)(int stage, int recon, int solver, SweepArgs const &a, cudaStream_t s);        // fused
  cudaError_t (*sweep_split)(int stage, int recon, int solver, SweepArgs const &a, cudaStream_t s);  // per direction
  cudaError_t (*calc_dti)(const void *U, int nx, int ny, int nz, int n_ghost, double dx, double dy, double dz,
                          double gamma, unsigned long long *key, cudaStream_t s);
  cudaError_t (*tfloor)(void *U, int nx, int ny, int nz, int n_ghost, double U_floor, cudaStream_t s);
  // SIMPLE integrator: primitives of U for every cell -> W, then sweep_split(1, ...) with simple = 1 in SweepArgs
  cudaError_t (*prim_from_cons)(const void *U, void *W, size_t n_cells, double gamma, int recon, cudaStream_t s);
  size_t real_bytes;
  int half_fields;  // fields of the half-step scratch array (5, or 7 in strict mode)
  int tile_y;       // rows of cells per CTA column
};

// ---- MHD (mhd_inst.cu) ----
struct MhdArgs {
  void *U;      // [8][N] conserved state, updated in place
  void *half;   // [8][N] half-step conserved state
  void *prim;   // [prim_fields][N] half-step primitives + cell-centred fields
  void *acc;    // [5][N]
  void *ctf;    // [3 directions][3][N] fluxes kept for constrained transport
  void *E;      // [6][N]: three edge EMFs, then three cell-centred reference EMFs
  int nx, ny, nz, n_ghost, recon, dir_chunk;
  double dx, dy, dz, dt, gamma, density_floor;
};
struct MhdOps {
  cudaError_t (*step)(MhdArgs const &a, cudaStream_t s);
  cudaError_t (*calc_dti)(const void *U, int nx, int ny, int nz, int n_ghost, double dx, double dy, double dz,
                          double gamma, unsigned long long *key, cudaStream_t s);
  size_t real_bytes;
  int prim_fields;
};
MhdOps const *mhd_ops_f64_strict();
MhdOps const *mhd_ops_f64_fast();

ModeOps const *mode_ops_f64_strict();
ModeOps const *mode_ops_f64_fast();
ModeOps const *mode_ops_f32_fast();  // nullptr when the fp32 TU is not linked

// element-type-templated copy kernels (ghost cells, pack/unpack), aux_kernels.cu
cudaError_t launch_set_ghost_cells(void *c, int nx, int ny, int nz, int n_fields, long n_cells, int n_ghost,
                                   const int flags[6], int isize, int jsize, int ksize, int imin, int jmin, int kmin,
                                   int dir, int diode, size_t real_bytes, cudaStream_t s);
cudaError_t launch_pack(void *buf, const void *c, int nx, int ny, int n_fields, long n_cells, long idxoffset,
                        int isize, int jsize, int ksize, int unpack, size_t real_bytes, cudaStream_t s);
cudaError_t launch_snapshot_pack(void *out, const void *field, int nx, int ny, int nz, int n_ghost, int nx_out,
                                 int ny_out, int nz_out, int mhd_direction, size_t in_bytes, size_t out_bytes,
                                 cudaStream_t s);
// exact_div_test.cu: shared-reciprocal division vs the compiler's a / b
cudaError_t launch_exact_div_test(const double *a, const double *b, double *q_shared, double *q_plain, long n,
                                  cudaStream_t s);
cudaError_t launch_exact_sqrt_test(const double *x, double *r_own, double *r_plain, long n, cudaStream_t s);
}  // namespace chl
