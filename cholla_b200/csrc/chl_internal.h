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
  int dir_chunk;  // cells per thread along the marching direction (0 = choose)
  int part;       // 0 = the whole stage; 1 = all but the interior box of the last pass; 2 = that interior box
  int shell;      // thickness of the boundary shell completed first when part != 0 (the ghost depth)
  int fused_predictor;  // 1: the predictor of a 3-D grid as ONE launch (vl_predictor.cuh) instead of three direction sweeps
  int simple;     // 1: SIMPLE integrator -- stage 1 reconstructs with the -DSIMPLE forms (PLMC gains its tracing)
  double dtodx, dtody, dtodz, gamma, density_floor;
  int apply_tfloor;
  double U_floor;
  int do_dti;
  double dx, dy, dz;
  unsigned long long *dti_key;
  unsigned long long *bad_cells;  // corrector: count of cells whose density came out <= 0
  const void *de_term;            // -DDE builds: n_cells reals from ModeOps::de_term (see DirParams::de_term)
};

struct ModeOps {
  // stage 0 = predictor (U -> W half), stage 1 = corrector (W half, U -> U); one launch per direction
  cudaError_t (*sweep_split)(int stage, int recon, int solver, SweepArgs const &a, cudaStream_t s);
  cudaError_t (*calc_dti)(const void *U, int nx, int ny, int nz, int n_ghost, double dx, double dy, double dz,
                          double gamma, unsigned long long *key, cudaStream_t s);
  cudaError_t (*tfloor)(void *U, int nx, int ny, int nz, int n_ghost, double U_floor, cudaStream_t s);
  // SIMPLE integrator: primitives of U for every cell -> W, then sweep_split(1, ...) with simple = 1 in SweepArgs
  cudaError_t (*prim_from_cons)(const void *U, void *W, size_t n_cells, double gamma, int recon, cudaStream_t s);
  size_t real_bytes;
  int half_fields;  // hydro fields of the half-step scratch array (5, or 7 in strict mode); the passive ones follow
  // -DSCALAR / -DDE builds (TUs compiled with CHL_NP / CHL_DE): passive fields riding with the five hydro fields
  // (GasEnergy last when dual_energy), and the dual-energy passes around the sweeps (nullptr without DE)
  int n_passive, dual_energy;
  cudaError_t (*de_term)(const void *U, void *out, int nx, int ny, int nz, const int lo[3], const int hi[3], double dtodx,
                         double dtody, double dtodz, double gamma, int mode, cudaStream_t s);
  cudaError_t (*de_finish)(void *U, int nx, int ny, int nz, int n_ghost, double gamma, int apply_tfloor, double U_floor,
                           int do_dti, double dx, double dy, double dz, unsigned long long *key, cudaStream_t s);
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
// fp64 TUs with passive fields: _s1 one scalar, _de dual energy, _de_s1 both (Makefile: PASSIVE_MODES)
ModeOps const *mode_ops_f64_strict_s1();
ModeOps const *mode_ops_f64_strict_de();
ModeOps const *mode_ops_f64_strict_de_s1();
ModeOps const *mode_ops_f64_fast_s1();
ModeOps const *mode_ops_f64_fast_de();
ModeOps const *mode_ops_f64_fast_de_s1();

}  // namespace chl
