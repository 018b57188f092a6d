// chl_api.cu -- the C ABI declared in include/cholla_b200.h.  Mode-agnostic: dispatches to the per-mode
// kernel TUs through ModeOps.  There is no CPU path: without a CUDA device every compute entry point fails
// with CHL_ERR_NO_DEVICE / CHL_ERR_CUDA.
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/cholla_b200.h"
#include "chl_internal.h"
#include "chl_aux.h"

namespace
{
thread_local std::string g_err;

int fail(int code, const char *what, cudaError_t e = cudaSuccess)
{
  g_err = what;
  if (e != cudaSuccess) {
    g_err += ": ";
    g_err += cudaGetErrorString(e);
  }
  return code;
}
#define CHL_CUDA(call)                                           \
  do {                                                           \
    cudaError_t e_ = (call);                                     \
    if (e_ != cudaSuccess) return fail(CHL_ERR_CUDA, #call, e_); \
  } while (0)

double host_decode(unsigned long long k)
{
  if (k == 0ull) return -DBL_MAX;  // accumulator still at its reset value: no cell contributed
  unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  double x;
  memcpy(&x, &u, 8);
  return x;
}
}  // namespace

struct chl_ctx {
  chl_config cfg;
  chl::ModeOps const *ops;
  void *d_half;               // W at t + dt/2 (primitives), half_fields * n_cells reals
  void *d_acc;                // direction-split sweeps: running sum of flux differences, (5 + passive) * n_cells reals
  void *d_grav;               // static gravity: [ndim][n_cells] accelerations (nullptr = off), chl_set_static_gravity
  const void *d_pot;          // -DGRAVITY: the caller's potential (nullptr = off), chl_set_gravity_potential
  int pot_five_point;
  void *d_gold;               // gravity: (rho, momentum)^n, 4 * n_cells reals, copied before the corrector
  void *d_de;                 // -DDE contexts: n_cells reals, the p div v term / partially updated GasEnergy (de_term)
  chl::MhdOps const *mhd;     // non-null in an MHD context
  void *d_prim;               // MHD: half-step primitives + cell-centred fields
  void *d_ctf;                // MHD: 9 * n_cells flux components kept for constrained transport
  void *d_E;                  // MHD: 3 * n_cells edge EMFs
  int dir_chunk;              // marching length override (0 = auto), chl_set_tuning
  int fused_predictor;        // chl_set_tuning: 1 = the single-launch predictor (vl_predictor.cuh) instead of three sweeps
  unsigned long long *d_key;  // [0] CFL accumulator (ordered key), [1] count of cells with density <= 0, [2] max |div B|
  unsigned long long *h_key;  // pinned, 3 words
  size_t n_cells;
  int ndim;
  int profiling;                       // record events around each stage
  std::vector<cudaEvent_t> ev;         // 3 per profiled step: before predictor, between, after corrector
  int ev_steps;                        // steps recorded since the last chl_stage_times
};
static const int kMaxProfiledSteps = 256;

extern "C" {

int chl_abi_version(void) { return CHL_ABI_VERSION; }
const char *chl_last_error(void) { return g_err.c_str(); }

int chl_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int chl_has_precision(int precision)
{
  if (precision == 2) return 1;
  if (precision == 1) return chl::mode_ops_f32_fast() != nullptr;
  return 0;
}

int chl_create(const chl_config *cfg, chl_ctx **out)
{
  if (!cfg || !out) return fail(CHL_ERR_INVALID, "chl_create: null argument");
  *out = nullptr;
  if (cfg->nx < 1 || cfg->ny < 1 || cfg->nz < 1) return fail(CHL_ERR_INVALID, "chl_create: bad grid size");
  if (cfg->ny == 1 && cfg->nz > 1) return fail(CHL_ERR_INVALID, "chl_create: nz > 1 requires ny > 1");
  if (cfg->recon < CHL_RECON_PCM || cfg->recon > CHL_RECON_PPMP) return fail(CHL_ERR_INVALID, "chl_create: recon");
  if (cfg->mhd && cfg->recon >= CHL_RECON_PLMP) return fail(CHL_ERR_INVALID, "chl_create: MHD supports PCM, PLMC and PPMC");
  if (cfg->solver < CHL_SOLVER_HLLC || cfg->solver > CHL_SOLVER_EXACT)
    return fail(CHL_ERR_INVALID, "chl_create: solver");
  if (cfg->mhd) {  // the reference's own constraints, src/utils/error_handling.cpp:83-118
    if (cfg->n_fields != 8) return fail(CHL_ERR_UNSUPPORTED, "chl_create: an MHD grid has n_fields = 8");
    if (cfg->solver != CHL_SOLVER_HLLD) return fail(CHL_ERR_INVALID, "chl_create: MHD only supports the HLLD solver");
    if (cfg->ny == 1 || cfg->nz == 1) return fail(CHL_ERR_INVALID, "chl_create: MHD runs must be 3D");
    if (cfg->n_scalars || cfg->dual_energy)
      return fail(CHL_ERR_UNSUPPORTED, "chl_create: scalars / dual energy are built for hydro contexts only");
  } else {
    if (cfg->n_scalars < 0 || (cfg->dual_energy != 0 && cfg->dual_energy != 1))
      return fail(CHL_ERR_INVALID, "chl_create: n_scalars / dual_energy");
    if (cfg->n_fields != 5 + cfg->n_scalars + cfg->dual_energy)
      return fail(CHL_ERR_INVALID, "chl_create: a hydro grid has n_fields = 5 + n_scalars + dual_energy (grid_enum.h:21-78)");
    if (cfg->solver == CHL_SOLVER_HLLD) return fail(CHL_ERR_INVALID, "chl_create: HLLD needs mhd = 1");
  }
  int const need_ghost =
      ((cfg->recon == CHL_RECON_PPMC || cfg->recon == CHL_RECON_PPMP) ? 4 : (cfg->recon == CHL_RECON_PCM ? 2 : 3)) +
      (cfg->mhd ? 1 : 0);
  if (cfg->n_ghost < need_ghost)
    return fail(CHL_ERR_INVALID, "chl_create: n_ghost too small for the reconstruction");
  if (cfg->nx <= 2 * cfg->n_ghost || (cfg->ny > 1 && cfg->ny <= 2 * cfg->n_ghost) ||
      (cfg->nz > 1 && cfg->nz <= 2 * cfg->n_ghost))
    return fail(CHL_ERR_INVALID, "chl_create: grid smaller than its ghost layers");
  chl::ModeOps const *ops = nullptr;
  bool const passive = !cfg->mhd && (cfg->n_scalars > 0 || cfg->dual_energy);
  if (passive) {  // the reference compiles NSCALARS / DE in; here each built combination is its own kernel TU
    if (cfg->precision == 1) return fail(CHL_ERR_UNSUPPORTED, "chl_create: scalars / dual energy are built in fp64 only");
    bool const pxmp = cfg->recon == CHL_RECON_PLMP || cfg->recon == CHL_RECON_PPMP;
    if ((cfg->solver != CHL_SOLVER_HLLC && cfg->solver != CHL_SOLVER_ROE && cfg->solver != CHL_SOLVER_HLL) ||
        (pxmp && cfg->solver != CHL_SOLVER_HLLC))
      return fail(CHL_ERR_UNSUPPORTED,
                  "chl_create: scalars / dual energy are built for PCM, PLMC, PPMC with HLLC, Roe, HLL and PLMP, PPMP with HLLC");
    if (cfg->n_scalars == 1 && !cfg->dual_energy) ops = cfg->strict ? chl::mode_ops_f64_strict_s1() : chl::mode_ops_f64_fast_s1();
    else if (cfg->n_scalars == 0) ops = cfg->strict ? chl::mode_ops_f64_strict_de() : chl::mode_ops_f64_fast_de();
    else if (cfg->n_scalars == 1) ops = cfg->strict ? chl::mode_ops_f64_strict_de_s1() : chl::mode_ops_f64_fast_de_s1();
    if (!ops) return fail(CHL_ERR_UNSUPPORTED, "chl_create: built passive-field combinations: 1 scalar, dual energy, dual energy + 1 scalar");
  } else if (cfg->precision == 2 || cfg->precision == 0) {
    ops = cfg->strict ? chl::mode_ops_f64_strict() : chl::mode_ops_f64_fast();
  } else if (cfg->precision == 1) {
    if (cfg->strict)
      return fail(CHL_ERR_UNSUPPORTED,
                  "chl_create: strict mode exists in fp64 only (the reference has no fp32 build)");
    ops = chl::mode_ops_f32_fast();
  }
  if (!ops) return fail(CHL_ERR_UNSUPPORTED, "chl_create: precision not built");
  chl::MhdOps const *mhd = nullptr;
  if (cfg->mhd) {
    if (cfg->precision == 1) return fail(CHL_ERR_UNSUPPORTED, "chl_create: the MHD kernels are built in fp64 only");
    mhd = cfg->strict ? chl::mhd_ops_f64_strict() : chl::mhd_ops_f64_fast();
  }
  if (chl_device_count() < 1)
    return fail(CHL_ERR_NO_DEVICE, "chl_create: no CUDA device (there is no CPU fallback)");

  chl_ctx *c = new chl_ctx();
  c->cfg = *cfg;
  c->cfg.precision = cfg->precision == 1 ? 1 : 2;
  c->ops = ops;
  c->n_cells = (size_t)cfg->nx * cfg->ny * cfg->nz;
  c->ndim = 1 + (cfg->ny > 1) + (cfg->nz > 1);
  c->d_half = nullptr;
  c->d_acc = nullptr;
  c->d_de = nullptr;
  c->d_grav = c->d_gold = nullptr;
  c->d_pot = nullptr;
  c->pot_five_point = 0;
  c->mhd = mhd;
  c->d_prim = c->d_ctf = c->d_E = nullptr;
  c->dir_chunk = 0;
  c->fused_predictor = 0;
  c->d_key = nullptr;
  c->h_key = nullptr;
  c->profiling = 0;
  c->ev_steps = 0;
  size_t const bytes = c->n_cells * (mhd ? 8 : ops->half_fields + ops->n_passive) * ops->real_bytes;
  cudaError_t e = cudaMalloc(&c->d_half, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&c->d_acc, c->n_cells * (5 + (mhd ? 0 : ops->n_passive)) * ops->real_bytes);
  if (e == cudaSuccess && !mhd && ops->dual_energy) {
    e = cudaMalloc(&c->d_de, c->n_cells * ops->real_bytes);
    if (e == cudaSuccess) e = cudaMemset(c->d_de, 0, c->n_cells * ops->real_bytes);
  }
  if (mhd) {
    size_t const w = c->n_cells * ops->real_bytes;
    if (e == cudaSuccess) e = cudaMalloc(&c->d_prim, mhd->prim_fields * w);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_ctf, 9 * w);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_E, 6 * w);  // three edge EMFs + three cell-centred reference EMFs
    if (e == cudaSuccess) e = cudaMemset(c->d_prim, 0, mhd->prim_fields * w);
    if (e == cudaSuccess) e = cudaMemset(c->d_ctf, 0, 9 * w);
    if (e == cudaSuccess) e = cudaMemset(c->d_E, 0, 6 * w);
  }
  if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_key, 3 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMallocHost((void **)&c->h_key, 3 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(c->d_half, 0, bytes);
  if (e == cudaSuccess) e = cudaMemset(c->d_key, 0, 3 * sizeof(unsigned long long));
  if (e != cudaSuccess) {
    chl_destroy(c);
    return fail(CHL_ERR_CUDA, "chl_create: allocation", e);
  }
  *out = c;
  return CHL_OK;
}

int chl_destroy(chl_ctx *c)
{
  if (!c) return CHL_OK;
  if (c->d_half) cudaFree(c->d_half);
  if (c->d_acc) cudaFree(c->d_acc);
  if (c->d_de) cudaFree(c->d_de);
  if (c->d_grav) cudaFree(c->d_grav);
  if (c->d_gold) cudaFree(c->d_gold);
  if (c->d_prim) cudaFree(c->d_prim);
  if (c->d_ctf) cudaFree(c->d_ctf);
  if (c->d_E) cudaFree(c->d_E);
  if (c->d_key) cudaFree(c->d_key);
  if (c->h_key) cudaFreeHost(c->h_key);
  for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
  delete c;
  return CHL_OK;
}

size_t chl_scratch_bytes(const chl_ctx *c)
{
  if (!c) return 0;
  size_t fields = c->mhd ? 8 + 5 + c->mhd->prim_fields + 9 + 6
                         : c->ops->half_fields + 5 + 2 * c->ops->n_passive + c->ops->dual_energy;
  return c->n_cells * fields * c->ops->real_bytes + 2 * sizeof(unsigned long long);
}

struct chl_halo;
static int halo_exchange_impl(chl_halo *h, void *U, cudaStream_t s);
static int halo_fork(chl_halo *h, cudaStream_t main, cudaStream_t *side);
static int halo_join(chl_halo *h, cudaStream_t main);

// halo != nullptr: the last corrector pass completes the boundary shell first, the ghost-cell exchange then runs on the
// halo's own stream while the interior of that pass runs on `s`; both are joined on `s` before returning.
static int step_impl(chl_ctx *c, void *U, double dt, int fused, cudaStream_t s, chl_halo *halo = nullptr)
{
  if (!c || !U) return fail(CHL_ERR_INVALID, "chl_vl_step: null argument");
  chl_config const &g = c->cfg;
  if (c->mhd) {
    chl::MhdArgs m;
    m.U = U; m.half = c->d_half; m.prim = c->d_prim; m.acc = c->d_acc; m.ctf = c->d_ctf; m.E = c->d_E;
    m.nx = g.nx; m.ny = g.ny; m.nz = g.nz; m.n_ghost = g.n_ghost; m.recon = g.recon; m.dir_chunk = c->dir_chunk;
    m.dx = g.dx; m.dy = g.dy; m.dz = g.dz; m.dt = dt; m.gamma = g.gamma; m.density_floor = g.density_floor;
    bool const prof = c->profiling && c->ev_steps < kMaxProfiledSteps;
    if (prof) {
      while ((int)c->ev.size() < 3 * (c->ev_steps + 1)) {
        cudaEvent_t e;
        CHL_CUDA(cudaEventCreate(&e));
        c->ev.push_back(e);
      }
      CHL_CUDA(cudaEventRecord(c->ev[3 * c->ev_steps], s));
      CHL_CUDA(cudaEventRecord(c->ev[3 * c->ev_steps + 1], s));
    }
    CHL_CUDA(c->mhd->step(m, s));
    if (prof) {
      CHL_CUDA(cudaEventRecord(c->ev[3 * c->ev_steps + 2], s));
      c->ev_steps++;
    }
    if (fused) {  // the same epilogue as the hydro path, as separate launches: temperature floor, then the CFL max
      if (g.temperature_floor_U >= 0.0)
        CHL_CUDA(c->ops->tfloor(U, g.nx, g.ny, g.nz, g.n_ghost, g.temperature_floor_U, s));
      CHL_CUDA(cudaMemsetAsync(c->d_key, 0, sizeof(unsigned long long), s));
      CHL_CUDA(c->mhd->calc_dti(U, g.nx, g.ny, g.nz, g.n_ghost, g.dx, g.dy, g.dz, g.gamma, c->d_key, s));
    }
    return CHL_OK;
  }
  chl::SweepArgs a;
  a.nx = g.nx;
  a.ny = g.ny;
  a.nz = g.nz;
  a.has_y = g.ny > 1;
  a.has_z = g.nz > 1;
  a.gamma = g.gamma;
  a.density_floor = g.density_floor;
  a.dx = g.dx;
  a.dy = g.dy;
  a.dz = g.dz;
  a.dti_key = c->d_key;
  a.bad_cells = c->d_key + 1;
  a.acc = c->d_acc;
  a.dir_chunk = c->dir_chunk;
  a.simple = 0;
  a.part = 0;
  a.shell = 0;
  a.fused_predictor = c->fused_predictor;
  a.de_term = c->d_de;
  bool const de = c->ops->dual_energy != 0;
  auto const sweep = c->ops->sweep_split;
  int const n[3] = {g.nx, g.ny, g.nz};
  // ---- predictor: every cell but the outermost ring (VL_3D_cuda.cu:408) ----
  for (int d = 0; d < 3; d++) {
    bool const on = d == 0 || (d == 1 && a.has_y) || (d == 2 && a.has_z);
    a.lo[d] = on ? 1 : 0;
    a.hi[d] = on ? n[d] - 1 : 1;
  }
  a.in = U;
  a.out = c->d_half;
  double const h = 0.5 * dt;
  a.dtodx = h / g.dx;
  a.dtody = h / g.dy;
  a.dtodz = h / g.dz;
  a.apply_tfloor = 0;
  a.U_floor = 0.0;
  a.do_dti = 0;
  bool const prof = c->profiling && c->ev_steps < kMaxProfiledSteps;
  if (prof) {
    while ((int)c->ev.size() < 3 * (c->ev_steps + 1)) {
      cudaEvent_t e;
      CHL_CUDA(cudaEventCreate(&e));
      c->ev.push_back(e);
    }
    CHL_CUDA(cudaEventRecord(c->ev[3 * c->ev_steps], s));
  }
  if (g.recon == CHL_RECON_PCM && c->ndim == 2) {
    // VL_Algorithm_2D_CUDA has no -DPCM branch in its second reconstruction (VL_2D_cuda.cu:95-120; the 1-D and 3-D files
    // do: VL_1D_cuda.cu:90, VL_3D_cuda.cu:226), so a -DPCM 2-D build solves the corrector's Riemann problems on the
    // interface arrays of step 1 -- copies of U^n.  Reproduced: the corrector reads U^n instead of the half-step state.
    CHL_CUDA(c->ops->prim_from_cons(U, c->d_half, c->n_cells, g.gamma, g.recon, s));
  } else {
    if (de) CHL_CUDA(c->ops->de_term(U, c->d_de, g.nx, g.ny, g.nz, a.lo, a.hi, a.dtodx, a.dtody, a.dtodz, g.gamma, 0, s));
    CHL_CUDA(sweep(0, g.recon, g.solver, a, s));
  }
  if (prof) CHL_CUDA(cudaEventRecord(c->ev[3 * c->ev_steps + 1], s));
  // ---- corrector: real cells (hydro_cuda.cu:225) ----
  for (int d = 0; d < 3; d++) {
    bool const on = d == 0 || (d == 1 && a.has_y) || (d == 2 && a.has_z);
    a.lo[d] = on ? g.n_ghost : 0;
    a.hi[d] = on ? n[d] - g.n_ghost : 1;
  }
  a.in = c->d_half;
  a.out = U;
  a.dtodx = dt / g.dx;
  a.dtody = dt / g.dy;
  a.dtodz = dt / g.dz;
  // -DTEMPERATURE_FLOOR builds floor unconditionally (grid3D.cpp:493-503): with the default threshold 0 the pass resets a
  // negative internal energy to zero; a negative temperature_floor_U stands for a build without the macro
  int const want_tfloor = (fused && g.temperature_floor_U >= 0.0 && c->ndim == 3) ? 1 : 0;
  bool const grav = c->d_grav != nullptr || (c->d_pot != nullptr && c->ndim == 3);
  if (fused) {
    // a -DDE context floors and reduces after Select_Internal_Energy / Sync_Energies (ModeOps::de_finish), a context with
    // static gravity after the source terms -- not in the sweep
    a.apply_tfloor = (de || grav) ? 0 : want_tfloor;
    a.U_floor = g.temperature_floor_U;
    a.do_dti = (de || grav) ? 0 : 1;
    CHL_CUDA(cudaMemsetAsync(c->d_key, 0, sizeof(unsigned long long), s));
  }
  if (de || grav) {
    // Partial_Update_Advected_Internal_Energy (from U^n, before anything is updated) and / or the copy of (rho, momentum)^n
    // the gravity terms average with; the corrector; the gravity source terms (inside the reference's update kernel:
    // before everything that follows it); Select_Internal_Energy + Sync_Energies over the completed cells.  The exchange
    // needs their result: no overlap here.
    size_t const rb = c->ops->real_bytes;
    if (de) CHL_CUDA(c->ops->de_term(U, c->d_de, g.nx, g.ny, g.nz, a.lo, a.hi, a.dtodx, a.dtody, a.dtodz, g.gamma, 1, s));
    if (grav) CHL_CUDA(cudaMemcpyAsync(c->d_gold, U, 4 * c->n_cells * rb, cudaMemcpyDeviceToDevice, s));
    CHL_CUDA(sweep(1, g.recon, g.solver, a, s));
    if (c->d_grav) CHL_CUDA(chl::launch_gravity_source(U, c->d_gold, c->d_grav, g.nx, g.ny, g.nz, g.n_ghost, dt, (int)rb, s));
    if (c->d_pot && c->ndim == 3)
      CHL_CUDA(chl::launch_gravity_potential_source(U, c->d_gold, c->d_pot, g.nx, g.ny, g.nz, g.n_ghost, dt, g.dx, g.dy, g.dz,
                                                    c->pot_five_point, (int)rb, s));
    if (de) {
      CHL_CUDA(c->ops->de_finish(U, g.nx, g.ny, g.nz, g.n_ghost, g.gamma, want_tfloor, g.temperature_floor_U, fused ? 1 : 0,
                                 g.dx, g.dy, g.dz, c->d_key, s));
    } else if (fused) {
      if (want_tfloor) CHL_CUDA(c->ops->tfloor(U, g.nx, g.ny, g.nz, g.n_ghost, g.temperature_floor_U, s));
      CHL_CUDA(c->ops->calc_dti(U, g.nx, g.ny, g.nz, g.n_ghost, g.dx, g.dy, g.dz, g.gamma, c->d_key, s));
    }
    if (halo) {
      int const r = halo_exchange_impl(halo, U, s);
      if (r != CHL_OK) return r;
    }
  } else if (halo && c->ndim == 3) {
    a.part  = 1;
    a.shell = g.n_ghost;
    CHL_CUDA(sweep(1, g.recon, g.solver, a, s));  // x and y passes, then the boundary shell of the z pass
    cudaStream_t side;
    int r = halo_fork(halo, s, &side);
    if (r != CHL_OK) return r;
    r = halo_exchange_impl(halo, U, side);        // reads shell cells, writes ghost cells
    if (r != CHL_OK) return r;
    a.part = 2;
    CHL_CUDA(sweep(1, g.recon, g.solver, a, s));  // interior of the z pass: touches neither
    r = halo_join(halo, s);
    if (r != CHL_OK) return r;
  } else {
    CHL_CUDA(sweep(1, g.recon, g.solver, a, s));
    if (halo) {
      int const r = halo_exchange_impl(halo, U, s);
      if (r != CHL_OK) return r;
    }
  }
  if (prof) {
    CHL_CUDA(cudaEventRecord(c->ev[3 * c->ev_steps + 2], s));
    c->ev_steps++;
  }
  return CHL_OK;
}

// Simple_Algorithm_{1,2,3}D_CUDA (src/integrators/simple_*_cuda.cu): reconstruct on U, Riemann, update over dt.
static int simple_impl(chl_ctx *c, void *U, double dt, int fused, cudaStream_t s)
{
  if (!c || !U) return fail(CHL_ERR_INVALID, "chl_simple_step: null argument");
  chl_config const &g = c->cfg;
  if (c->mhd) return fail(CHL_ERR_UNSUPPORTED, "chl_simple_step: MHD requires the Van Leer integrator (error_handling.cpp:89)");
  if (c->d_grav || c->d_pot)
    return fail(CHL_ERR_UNSUPPORTED, "chl_simple_step: the gravity source terms are built for the Van Leer integrator");
  if (c->ops->n_passive)
    return fail(CHL_ERR_UNSUPPORTED, "chl_simple_step: scalars / dual energy are built for the Van Leer integrator");
  if (g.recon != CHL_RECON_PCM && g.recon != CHL_RECON_PLMC)
    return fail(CHL_ERR_UNSUPPORTED, "chl_simple_step: PCM and PLMC only (the CTU forms of PLMP / PPMP / PPMC are not built)");
  chl::SweepArgs a;
  a.nx = g.nx; a.ny = g.ny; a.nz = g.nz;
  a.has_y = g.ny > 1;
  a.has_z = g.nz > 1;
  a.gamma = g.gamma;
  a.density_floor = g.density_floor;
  a.dx = g.dx; a.dy = g.dy; a.dz = g.dz;
  a.dti_key = c->d_key;
  a.bad_cells = c->d_key + 1;
  a.acc = c->d_acc;
  a.dir_chunk = c->dir_chunk;
  a.simple = 1;
  a.part = 0;
  a.shell = 0;
  a.fused_predictor = 0;
  a.de_term = nullptr;
  int const n[3] = {g.nx, g.ny, g.nz};
  for (int d = 0; d < 3; d++) {
    bool const on = d == 0 || (d == 1 && a.has_y) || (d == 2 && a.has_z);
    a.lo[d] = on ? g.n_ghost : 0;
    a.hi[d] = on ? n[d] - g.n_ghost : 1;
  }
  a.in = c->d_half;
  a.out = U;
  a.dtodx = dt / g.dx;
  a.dtody = dt / g.dy;
  a.dtodz = dt / g.dz;
  a.apply_tfloor = 0;
  a.U_floor = 0.0;
  a.do_dti = 0;
  if (fused) {
    a.apply_tfloor = (g.temperature_floor_U >= 0.0 && c->ndim == 3) ? 1 : 0;
    a.U_floor = g.temperature_floor_U;
    a.do_dti = 1;
    CHL_CUDA(cudaMemsetAsync(c->d_key, 0, sizeof(unsigned long long), s));
  }
  bool const prof = c->profiling && c->ev_steps < kMaxProfiledSteps;  // "predictor" = the primitive pass here
  if (prof) {
    while ((int)c->ev.size() < 3 * (c->ev_steps + 1)) {
      cudaEvent_t e;
      CHL_CUDA(cudaEventCreate(&e));
      c->ev.push_back(e);
    }
    CHL_CUDA(cudaEventRecord(c->ev[3 * c->ev_steps], s));
  }
  CHL_CUDA(c->ops->prim_from_cons(U, c->d_half, c->n_cells, g.gamma, g.recon, s));
  if (prof) CHL_CUDA(cudaEventRecord(c->ev[3 * c->ev_steps + 1], s));
  CHL_CUDA(c->ops->sweep_split(1, g.recon, g.solver, a, s));
  if (prof) {
    CHL_CUDA(cudaEventRecord(c->ev[3 * c->ev_steps + 2], s));
    c->ev_steps++;
  }
  return CHL_OK;
}

int chl_simple_step(chl_ctx *c, void *U, double dt, void *stream) { return simple_impl(c, U, dt, 0, (cudaStream_t)stream); }
int chl_simple_step_fused(chl_ctx *c, void *U, double dt, void *stream)
{
  return simple_impl(c, U, dt, 1, (cudaStream_t)stream);
}

int chl_set_profiling(chl_ctx *c, int enable)
{
  if (!c) return fail(CHL_ERR_INVALID, "chl_set_profiling: null context");
  c->profiling = enable ? 1 : 0;
  c->ev_steps = 0;
  return CHL_OK;
}

int chl_stage_times(chl_ctx *c, double ms[2], int *n_steps)
{
  if (!c || !ms || !n_steps) return fail(CHL_ERR_INVALID, "chl_stage_times: null argument");
  ms[0] = ms[1] = 0.0;
  *n_steps = c->ev_steps;
  if (c->ev_steps == 0) return CHL_OK;
  CHL_CUDA(cudaEventSynchronize(c->ev[3 * c->ev_steps - 1]));
  for (int i = 0; i < c->ev_steps; i++) {
    float a = 0.f, b = 0.f;
    CHL_CUDA(cudaEventElapsedTime(&a, c->ev[3 * i], c->ev[3 * i + 1]));
    CHL_CUDA(cudaEventElapsedTime(&b, c->ev[3 * i + 1], c->ev[3 * i + 2]));
    ms[0] += a;
    ms[1] += b;
  }
  c->ev_steps = 0;
  return CHL_OK;
}

int chl_vl_step(chl_ctx *c, void *U, double dt, void *stream)
{
  return step_impl(c, U, dt, 0, (cudaStream_t)stream);
}
int chl_vl_step_fused(chl_ctx *c, void *U, double dt, void *stream)
{
  return step_impl(c, U, dt, 1, (cudaStream_t)stream);
}

int chl_dti_result(chl_ctx *c, double *h_dti, void *stream)
{
  if (!c || !h_dti) return fail(CHL_ERR_INVALID, "chl_dti_result: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  CHL_CUDA(cudaMemcpyAsync(c->h_key, c->d_key, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  CHL_CUDA(cudaStreamSynchronize(s));
  *h_dti = host_decode(*c->h_key);
  return CHL_OK;
}

int chl_bad_cells(chl_ctx *c, unsigned long long *h_count, void *stream)
{
  if (!c || !h_count) return fail(CHL_ERR_INVALID, "chl_bad_cells: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  CHL_CUDA(cudaMemcpyAsync(c->h_key + 1, c->d_key + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  CHL_CUDA(cudaMemsetAsync(c->d_key + 1, 0, sizeof(unsigned long long), s));
  CHL_CUDA(cudaStreamSynchronize(s));
  *h_count = c->h_key[1];
  return CHL_OK;
}

int chl_mhd_max_divergence(chl_ctx *c, const void *U, double *h_max, void *stream)
{
  if (!c || !U || !h_max) return fail(CHL_ERR_INVALID, "chl_mhd_max_divergence: null argument");
  if (!c->mhd) return fail(CHL_ERR_INVALID, "chl_mhd_max_divergence: not an MHD context");
  cudaStream_t s = (cudaStream_t)stream;
  chl_config const &g = c->cfg;
  CHL_CUDA(cudaMemsetAsync(c->d_key + 2, 0, sizeof(unsigned long long), s));
  CHL_CUDA(chl::launch_divb(U, g.nx, g.ny, g.nz, g.dx, g.dy, g.dz, c->d_key + 2, c->ops->real_bytes, s));
  CHL_CUDA(cudaMemcpyAsync(c->h_key + 2, c->d_key + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  CHL_CUDA(cudaStreamSynchronize(s));
  memcpy(h_max, c->h_key + 2, sizeof(double));
  return CHL_OK;
}

int chl_set_tuning(chl_ctx *c, int dir_chunk, int fused_predictor)
{
  if (!c || dir_chunk < 0) return fail(CHL_ERR_INVALID, "chl_set_tuning: bad argument");
  c->dir_chunk = dir_chunk;
  c->fused_predictor = fused_predictor ? 1 : 0;
  return CHL_OK;
}

int chl_dti_device_ptr(chl_ctx *c, void **d_key)
{
  if (!c || !d_key) return fail(CHL_ERR_INVALID, "chl_dti_device_ptr: null argument");
  *d_key = c->d_key;
  return CHL_OK;
}

double chl_dti_decode(unsigned long long key) { return host_decode(key); }

int chl_calc_dti(chl_ctx *c, const void *U, double *h_dti, void *stream)
{
  if (!c || !U || !h_dti) return fail(CHL_ERR_INVALID, "chl_calc_dti: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  chl_config const &g = c->cfg;
  CHL_CUDA(cudaMemsetAsync(c->d_key, 0, sizeof(unsigned long long), s));
  if (c->mhd) CHL_CUDA(c->mhd->calc_dti(U, g.nx, g.ny, g.nz, g.n_ghost, g.dx, g.dy, g.dz, g.gamma, c->d_key, s));
  else CHL_CUDA(c->ops->calc_dti(U, g.nx, g.ny, g.nz, g.n_ghost, g.dx, g.dy, g.dz, g.gamma, c->d_key, s));
  return chl_dti_result(c, h_dti, stream);
}

int chl_apply_temperature_floor(chl_ctx *c, void *U, double U_floor, void *stream)
{
  if (!c || !U) return fail(CHL_ERR_INVALID, "chl_apply_temperature_floor: null argument");
  chl_config const &g = c->cfg;
  CHL_CUDA(c->ops->tfloor(U, g.nx, g.ny, g.nz, g.n_ghost, U_floor, (cudaStream_t)stream));
  return CHL_OK;
}

static int gravity_alloc(chl_ctx *c, bool field)
{
  size_t const w = c->n_cells * c->ops->real_bytes;
  if (field && !c->d_grav) CHL_CUDA(cudaMalloc(&c->d_grav, (size_t)c->ndim * w));
  if (!c->d_gold) CHL_CUDA(cudaMalloc(&c->d_gold, 4 * w));
  return CHL_OK;
}

int chl_set_gravity_potential(chl_ctx *c, const void *d_potential, int five_point)
{
  if (!c) return fail(CHL_ERR_INVALID, "chl_set_gravity_potential: null context");
  if (c->mhd) return fail(CHL_ERR_UNSUPPORTED, "chl_set_gravity_potential: hydro contexts only");
  if (d_potential && c->ndim != 3)
    return fail(CHL_ERR_INVALID, "chl_set_gravity_potential: the reference's -DGRAVITY source terms exist in 3-D only");
  if (d_potential && five_point && c->cfg.n_ghost < 2)
    return fail(CHL_ERR_INVALID, "chl_set_gravity_potential: the five-point gradient needs two ghost layers");
  if (d_potential) {
    int const r = gravity_alloc(c, false);
    if (r != CHL_OK) return r;
  }
  c->d_pot = d_potential;
  c->pot_five_point = five_point ? 1 : 0;
  return CHL_OK;
}

int chl_set_static_gravity(chl_ctx *c, int custom_grav, int x_off, int y_off, int z_off, double xbound, double ybound,
                           double zbound, void *stream)
{
  if (!c) return fail(CHL_ERR_INVALID, "chl_set_static_gravity: null context");
  if (c->mhd) return fail(CHL_ERR_UNSUPPORTED, "chl_set_static_gravity: hydro contexts only");
  if (custom_grav < 0) {
    CHL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (c->d_grav) cudaFree(c->d_grav);
    c->d_grav = nullptr;
    if (c->d_gold && !c->d_pot) {
      cudaFree(c->d_gold);
      c->d_gold = nullptr;
    }
    return CHL_OK;
  }
  int const r = gravity_alloc(c, true);
  if (r != CHL_OK) return r;
  chl_config const &g = c->cfg;
  int const off[3] = {x_off, y_off, z_off};
  double const d[3] = {g.dx, g.dy, g.dz}, bound[3] = {xbound, ybound, zbound};
  CHL_CUDA(chl::launch_static_gravity_field(c->d_grav, g.nx, g.ny, g.nz, g.n_ghost, custom_grav, off, d, bound,
                                            (int)c->ops->real_bytes, (cudaStream_t)stream));
  return CHL_OK;
}

int chl_set_gravity_field(chl_ctx *c, const void *d_g, void *stream)
{
  if (!c || !d_g) return fail(CHL_ERR_INVALID, "chl_set_gravity_field: null argument");
  if (c->mhd) return fail(CHL_ERR_UNSUPPORTED, "chl_set_gravity_field: hydro contexts only");
  int const r = gravity_alloc(c, true);
  if (r != CHL_OK) return r;
  CHL_CUDA(cudaMemcpyAsync(c->d_grav, d_g, (size_t)c->ndim * c->n_cells * c->ops->real_bytes, cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
  return CHL_OK;
}

int chl_gravity_field_ptr(chl_ctx *c, void **d_g)
{
  if (!c || !d_g) return fail(CHL_ERR_INVALID, "chl_gravity_field_ptr: null argument");
  *d_g = c->d_grav;
  return CHL_OK;
}

int chl_copy_gravity_field(chl_ctx *c, void *d_dst, void *stream)
{
  if (!c || !d_dst) return fail(CHL_ERR_INVALID, "chl_copy_gravity_field: null argument");
  if (!c->d_grav) return fail(CHL_ERR_INVALID, "chl_copy_gravity_field: gravity is off");
  CHL_CUDA(cudaMemcpyAsync(d_dst, c->d_grav, (size_t)c->ndim * c->n_cells * c->ops->real_bytes, cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
  return CHL_OK;
}

int chl_cooling_update(chl_ctx *c, void *U, double dt, void *stream)
{
  if (!c || !U) return fail(CHL_ERR_INVALID, "chl_cooling_update: null argument");
  if (c->mhd) return fail(CHL_ERR_UNSUPPORTED, "chl_cooling_update: hydro contexts only");
  chl_config const &g = c->cfg;
  CHL_CUDA(chl::launch_cooling_update(U, g.nx, g.ny, g.nz, g.n_ghost, g.n_fields, c->ops->dual_energy, dt, g.gamma,
                                      (int)c->ops->real_bytes, (cudaStream_t)stream));
  return CHL_OK;
}

int chl_apply_scalar_floor(chl_ctx *c, void *U, int field_num, double scalar_floor, void *stream)
{
  if (!c || !U) return fail(CHL_ERR_INVALID, "chl_apply_scalar_floor: null argument");
  chl_config const &g = c->cfg;
  if (field_num < 0 || field_num >= g.n_fields) return fail(CHL_ERR_INVALID, "chl_apply_scalar_floor: field_num");
  CHL_CUDA(chl::launch_scalar_floor(U, g.nx, g.ny, g.nz, g.n_ghost, field_num, scalar_floor, (int)c->ops->real_bytes,
                                    (cudaStream_t)stream));
  return CHL_OK;
}

int chl_boundary_extents(int nx, int ny, int nz, int g, int dir, int imin[3], int imax[3])
{
  if (dir < 0 || dir > 5 || !imin || !imax) return fail(CHL_ERR_INVALID, "chl_boundary_extents: bad argument");
  int const n[3] = {nx, ny, nz};
  int lo[3] = {0, 0, 0}, hi[3] = {nx, ny, nz};
  // a face spans the real cells of the axes after its own and everything in the axes before it, so edges and
  // corners are filled by the x -> y -> z ordering (boundary_conditions.cpp:419-)
  if (ny > 1) {
    lo[1] = g;
    hi[1] = ny - g;
  }
  if (nz > 1) {
    lo[2] = g;
    hi[2] = nz - g;
  }
  int const ax = dir / 2;
  for (int d = 0; d < ax; d++) {
    lo[d] = 0;
    hi[d] = n[d];
  }
  if (dir % 2 == 0) {
    lo[ax] = 0;
    hi[ax] = g;
  } else {
    lo[ax] = n[ax] - g;
    hi[ax] = n[ax];
  }
  for (int d = 0; d < 3; d++) {
    imin[d] = lo[d];
    imax[d] = hi[d];
  }
  return CHL_OK;
}

int chl_set_ghost_cells_ext(void *U, int nx, int ny, int nz, int n_fields, int n_cells, int n_ghost,
                            const int flags[6], int isize, int jsize, int ksize, int imin, int jmin, int kmin,
                            int dir, int precision, int mhd, void *stream)
{
  if (!U || !flags || dir < 0 || dir > 5) return fail(CHL_ERR_INVALID, "chl_set_ghost_cells_ext: bad argument");
  if (n_fields < 5) return fail(CHL_ERR_INVALID, "chl_set_ghost_cells_ext: n_fields < 5");
  if (mhd && n_fields < 8) return fail(CHL_ERR_INVALID, "chl_set_ghost_cells_ext: an MHD grid has 8 fields");
  CHL_CUDA(chl::launch_set_ghost_cells(U, nx, ny, nz, n_fields, n_cells, n_ghost, flags, isize, jsize, ksize, imin,
                                       jmin, kmin, dir, mhd ? 0 : 1, precision == 1 ? 4 : 8, (cudaStream_t)stream));
  return CHL_OK;
}

int chl_set_ghost_cells(chl_ctx *c, void *U, const int flags[6], int dir, void *stream)
{
  if (!c || !U || !flags || dir < 0 || dir > 5) return fail(CHL_ERR_INVALID, "chl_set_ghost_cells: bad argument");
  if (flags[dir] == CHL_BC_CUSTOM || flags[dir] == CHL_BC_MPI) return CHL_OK;  // Grid3D::Set_Boundaries early-outs
  chl_config const &g = c->cfg;
  if (dir / 2 >= c->ndim) return CHL_OK;
  int lo[3], hi[3];
  chl_boundary_extents(g.nx, g.ny, g.nz, g.n_ghost, dir, lo, hi);
  return chl_set_ghost_cells_ext(U, g.nx, g.ny, g.nz, g.n_fields, (int)c->n_cells, g.n_ghost, flags, hi[0] - lo[0],
                                 hi[1] - lo[1], hi[2] - lo[2], lo[0], lo[1], lo[2], dir, g.precision, g.mhd, stream);
}

int chl_set_boundaries(chl_ctx *c, void *U, const int flags[6], void *stream)
{
  if (!c) return fail(CHL_ERR_INVALID, "chl_set_boundaries: null context");
  for (int dir = 0; dir < 2 * c->ndim; dir++) {
    int const r = chl_set_ghost_cells(c, U, flags, dir, stream);
    if (r != CHL_OK) return r;
  }
  return CHL_OK;
}

int chl_pack_buffer(void *buf, const void *U, int nx, int ny, int n_fields, long n_cells, long idxoffset, int isize,
                    int jsize, int ksize, int precision, void *stream)
{
  if (!buf || !U) return fail(CHL_ERR_INVALID, "chl_pack_buffer: null argument");
  CHL_CUDA(chl::launch_pack(buf, U, nx, ny, n_fields, n_cells, idxoffset, isize, jsize, ksize, 0,
                            precision == 1 ? 4 : 8, (cudaStream_t)stream));
  return CHL_OK;
}

int chl_unpack_buffer(const void *buf, void *U, int nx, int ny, int n_fields, long n_cells, long idxoffset, int isize,
                      int jsize, int ksize, int precision, void *stream)
{
  if (!buf || !U) return fail(CHL_ERR_INVALID, "chl_unpack_buffer: null argument");
  CHL_CUDA(chl::launch_pack((void *)buf, U, nx, ny, n_fields, n_cells, idxoffset, isize, jsize, ksize, 1,
                            precision == 1 ? 4 : 8, (cudaStream_t)stream));
  return CHL_OK;
}

int chl_pack_snapshot_field(void *d_out, const void *d_field, int nx, int ny, int nz, int n_ghost, int nx_out, int ny_out,
                            int nz_out, int mhd_direction, int precision, int out_precision, void *stream)
{
  if (!d_out || !d_field) return fail(CHL_ERR_INVALID, "chl_pack_snapshot_field: null argument");
  if (mhd_direction < -1 || mhd_direction > 2 || nx < 1 || ny < 1 || nz < 1 || n_ghost < 1 ||
      nx_out + n_ghost - (mhd_direction == 0) > nx || (ny > 1 && ny_out + n_ghost - (mhd_direction == 1) > ny) ||
      (nz > 1 && nz_out + n_ghost - (mhd_direction == 2) > nz))
    return fail(CHL_ERR_INVALID, "chl_pack_snapshot_field: extents outside the grid");
  CHL_CUDA(chl::launch_snapshot_pack(d_out, d_field, nx, ny, nz, n_ghost, nx_out, ny_out, nz_out, mhd_direction,
                                     precision == 1 ? 4 : 8, out_precision == 1 ? 4 : 8, (cudaStream_t)stream));
  return CHL_OK;
}

}  // extern "C"

// ---- native ghost-cell exchange between the ranks of one box (halo_p2p.cu) -----------------------------------------
namespace
{
struct HaloLayout {  // this rank's receive block, as its neighbours need to know it
  unsigned long long buf_off[6][2];  // byte offset of the receive buffer of face f, exchange parity p
  unsigned long long buf_bytes[6];
  unsigned long long flag_off;       // 6 arrival words (unsigned long long)
  unsigned long long block_bytes;
};
struct HaloBlob {  // what a rank publishes to its neighbours (chl_halo_export)
  cudaIpcMemHandle_t mem;
  HaloLayout lay;
  int rank, device;
  unsigned long long magic;
};
constexpr unsigned long long kHaloMagic = 0x63686c68616c6f31ull;  // "chlhalo1"

struct Slab {
  long idxoffset;
  int isize, jsize, ksize;
};
// Load_/Unload_Hydro_DeviceBuffer_{X,Y,Z}{0,1} geometry (mpi_boundaries.cpp:113-298): g layers along the face's axis, the
// full extent (ghosts included) of the axes before it, the real cells of the axes after it
Slab halo_slab(int nx, int ny, int nz, int g, int face, bool recv)
{
  int const n[3] = {nx, ny, nz};
  int const ax = face / 2, side = face % 2;
  int lo[3] = {0, 0, 0}, size[3] = {1, 1, 1};
  for (int d = 0; d < 3; d++) {
    if (n[d] == 1) continue;
    if (d < ax) {
      lo[d]   = 0;
      size[d] = n[d];
    } else if (d > ax) {
      lo[d]   = g;
      size[d] = n[d] - 2 * g;
    } else {
      size[d] = g;
      lo[d]   = recv ? (side == 0 ? 0 : n[d] - g) : (side == 0 ? g : n[d] - 2 * g);
    }
  }
  return Slab{lo[0] + (long)nx * (lo[1] + (long)ny * lo[2]), size[0], size[1], size[2]};
}
}  // namespace

struct chl_halo {
  chl_ctx *ctx;
  int flags[6];
  int rank;
  char *d_block;                 // this rank's receive buffers + arrival words + send counters + error word
  HaloLayout lay;
  unsigned *d_counter;           // 6 last-block counters of this rank's send launches
  unsigned *d_err;
  unsigned *h_err;               // pinned
  char *peer_base[6];            // neighbour behind face f: its block mapped here
  HaloLayout peer_lay[6];
  std::vector<std::pair<int, char *>> opened;  // (peer rank, mapped base): one mapping per distinct peer
  unsigned long long seq;        // exchanges done
  cudaStream_t side;
  cudaEvent_t ev_fork, ev_join;
};

static int halo_fork(chl_halo *h, cudaStream_t main, cudaStream_t *side)
{
  CHL_CUDA(cudaEventRecord(h->ev_fork, main));
  CHL_CUDA(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
  *side = h->side;
  return CHL_OK;
}
static int halo_join(chl_halo *h, cudaStream_t main)
{
  CHL_CUDA(cudaEventRecord(h->ev_join, h->side));
  CHL_CUDA(cudaStreamWaitEvent(main, h->ev_join, 0));
  return CHL_OK;
}

// Grid3D::Set_Boundaries_MPI_BLOCK (mpi_boundaries.cpp:29-111): per axis -- send both slabs, fill the physical faces of
// the axis, receive both slabs
static int halo_exchange_impl(chl_halo *h, void *U, cudaStream_t s)
{
  chl_ctx *c = h->ctx;
  chl_config const &g = c->cfg;
  size_t const rb = c->ops->real_bytes;
  unsigned long long const seq = ++h->seq;
  int const par = (int)(seq & 1ull);
  for (int ax = 0; ax < c->ndim; ax++) {
    for (int face = 2 * ax; face < 2 * ax + 2; face++) {
      if (h->flags[face] != CHL_BC_MPI) continue;
      if (!h->peer_base[face]) return fail(CHL_ERR_INVALID, "chl_halo_exchange: a neighbour is not connected");
      Slab const sl = halo_slab(g.nx, g.ny, g.nz, g.n_ghost, face, false);
      int const pf = face ^ 1;  // my low slab fills the neighbour's HIGH ghosts and vice versa
      HaloLayout const &pl = h->peer_lay[face];
      size_t const bytes = (size_t)sl.isize * sl.jsize * sl.ksize * g.n_fields * rb;
      if (bytes != pl.buf_bytes[pf]) return fail(CHL_ERR_INVALID, "chl_halo_exchange: slab size differs from the neighbour's");
      CHL_CUDA(chl::launch_halo_send(h->peer_base[face] + pl.buf_off[pf][par], U, g.nx, g.ny, g.n_fields, (long)c->n_cells,
                                     sl.idxoffset, sl.isize, sl.jsize, sl.ksize, h->d_counter + face,
                                     (unsigned long long *)(h->peer_base[face] + pl.flag_off) + pf, seq, rb, s));
    }
    for (int face = 2 * ax; face < 2 * ax + 2; face++) {
      int const r = chl_set_ghost_cells(c, U, h->flags, face, (void *)s);  // skips exchange and custom faces
      if (r != CHL_OK) return r;
    }
    for (int face = 2 * ax; face < 2 * ax + 2; face++) {
      if (h->flags[face] != CHL_BC_MPI) continue;
      Slab const sl = halo_slab(g.nx, g.ny, g.nz, g.n_ghost, face, true);
      CHL_CUDA(chl::launch_halo_recv(h->d_block + h->lay.buf_off[face][par], U, g.nx, g.ny, g.n_fields, (long)c->n_cells,
                                     sl.idxoffset, sl.isize, sl.jsize, sl.ksize,
                                     (const unsigned long long *)(h->d_block + h->lay.flag_off) + face, seq, h->d_err, rb, s));
    }
  }
  return CHL_OK;
}

extern "C" {

size_t chl_halo_blob_bytes(void) { return sizeof(HaloBlob); }

int chl_halo_create(chl_ctx *c, const int flags[6], int rank, chl_halo **out)
{
  if (!c || !flags || !out) return fail(CHL_ERR_INVALID, "chl_halo_create: null argument");
  *out = nullptr;
  chl_halo *h = new chl_halo();
  h->ctx = c;
  h->rank = rank;
  h->d_block = nullptr;
  h->d_counter = h->d_err = h->h_err = nullptr;
  h->seq = 0;
  h->side = nullptr;
  h->ev_fork = h->ev_join = nullptr;
  chl_config const &g = c->cfg;
  size_t off = 0;
  for (int f = 0; f < 6; f++) {
    h->flags[f] = flags[f];
    h->peer_base[f] = nullptr;
    size_t bytes = 0;
    if (flags[f] == CHL_BC_MPI && f / 2 < c->ndim) {
      Slab const sl = halo_slab(g.nx, g.ny, g.nz, g.n_ghost, f, true);
      bytes = (size_t)sl.isize * sl.jsize * sl.ksize * g.n_fields * c->ops->real_bytes;
    }
    h->lay.buf_bytes[f] = bytes;
    for (int p = 0; p < 2; p++) {
      h->lay.buf_off[f][p] = off;
      off += (bytes + 255) / 256 * 256;
    }
  }
  h->lay.flag_off = off;
  off += 256;
  h->lay.block_bytes = off;
  cudaError_t e = cudaMalloc((void **)&h->d_block, off);
  if (e == cudaSuccess) e = cudaMemset(h->d_block + h->lay.flag_off, 0, 256);
  if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_counter, 8 * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMemset(h->d_counter, 0, 8 * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMallocHost((void **)&h->h_err, sizeof(unsigned));
  if (e == cudaSuccess) {
    h->d_err = h->d_counter + 6;
    int lo = 0, hi = 0;
    e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
    // highest priority: the exchange's few CTAs go ahead of the interior update's queue of CTAs
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    chl_halo_destroy(h);
    return fail(CHL_ERR_CUDA, "chl_halo_create", e);
  }
  *out = h;
  return CHL_OK;
}

int chl_halo_destroy(chl_halo *h)
{
  if (!h) return CHL_OK;
  cudaDeviceSynchronize();
  for (auto &p : h->opened) cudaIpcCloseMemHandle(p.second);
  if (h->d_block) cudaFree(h->d_block);
  if (h->d_counter) cudaFree(h->d_counter);
  if (h->h_err) cudaFreeHost(h->h_err);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  delete h;
  return CHL_OK;
}

int chl_halo_export(chl_halo *h, void *blob)
{
  if (!h || !blob) return fail(CHL_ERR_INVALID, "chl_halo_export: null argument");
  HaloBlob b;
  memset(&b, 0, sizeof b);
  CHL_CUDA(cudaIpcGetMemHandle(&b.mem, h->d_block));
  b.lay = h->lay;
  b.rank = h->rank;
  CHL_CUDA(cudaGetDevice(&b.device));
  b.magic = kHaloMagic;
  memcpy(blob, &b, sizeof b);
  return CHL_OK;
}

int chl_halo_connect(chl_halo *h, int face, const void *peer_blob)
{
  if (!h || !peer_blob || face < 0 || face > 5) return fail(CHL_ERR_INVALID, "chl_halo_connect: bad argument");
  HaloBlob b;
  memcpy(&b, peer_blob, sizeof b);
  if (b.magic != kHaloMagic) return fail(CHL_ERR_INVALID, "chl_halo_connect: not a chl_halo_export blob");
  if (b.rank == h->rank) return fail(CHL_ERR_INVALID, "chl_halo_connect: a rank cannot be its own neighbour (use a periodic face)");
  char *base = nullptr;
  for (auto &p : h->opened)
    if (p.first == b.rank) base = p.second;
  if (!base) {
    CHL_CUDA(cudaIpcOpenMemHandle((void **)&base, b.mem, cudaIpcMemLazyEnablePeerAccess));
    h->opened.push_back(std::make_pair(b.rank, base));
  }
  h->peer_base[face] = base;
  h->peer_lay[face]  = b.lay;
  return CHL_OK;
}

int chl_halo_exchange(chl_halo *h, void *U, void *stream)
{
  if (!h || !U) return fail(CHL_ERR_INVALID, "chl_halo_exchange: null argument");
  return halo_exchange_impl(h, U, (cudaStream_t)stream);
}

int chl_halo_error(chl_halo *h, void *stream)
{
  if (!h) return fail(CHL_ERR_INVALID, "chl_halo_error: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  CHL_CUDA(cudaMemcpyAsync(h->h_err, h->d_err, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
  CHL_CUDA(cudaStreamSynchronize(s));
  if (*h->h_err) return fail(CHL_ERR_CUDA, "chl_halo: a receive waited 20 s for its neighbour's slab and gave up");
  return CHL_OK;
}

int chl_vl_step_exchange(chl_ctx *c, chl_halo *h, void *U, double dt, void *stream)
{
  if (!h || h->ctx != c) return fail(CHL_ERR_INVALID, "chl_vl_step_exchange: the halo belongs to another context");
  if (c->mhd) {  // the MHD step is one block of launches: exchange after it, same stream
    int const r = step_impl(c, U, dt, 1, (cudaStream_t)stream);
    return r != CHL_OK ? r : halo_exchange_impl(h, U, (cudaStream_t)stream);
  }
  return step_impl(c, U, dt, 1, (cudaStream_t)stream, h);
}

int chl_dti_copy_key(chl_ctx *c, void *d_dst, void *stream)
{
  if (!c || !d_dst) return fail(CHL_ERR_INVALID, "chl_dti_copy_key: null argument");
  CHL_CUDA(cudaMemcpyAsync(d_dst, c->d_key, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return CHL_OK;
}

}  // extern "C"

// ---- host-only mirrors of the reference's decomposition logic (src/mpi/mpi_routines.cpp) ----
extern "C" {
static int greatest_prime_factor(int n)
{
  int best = 1;
  for (int p = 2; (long)p * p <= n; p++) {
    while (n % p == 0) {
      best = p;
      n /= p;
    }
  }
  if (n > 1) best = n;
  return best;
}

int chl_tile_decomposition(int nproc, int nxg, int nyg, int nzg, int np[3])
{
  if (nproc < 1 || !np) return fail(CHL_ERR_INVALID, "chl_tile_decomposition: bad argument");
  np[0] = np[1] = np[2] = 1;
  if (nzg == 1 && nyg == 1) {
    np[0] = nproc;
    return CHL_OK;
  }
  if (nzg == 1) {
    int const a = greatest_prime_factor(nproc), b = nproc / a;
    np[0] = a > b ? a : b;
    np[1] = a > b ? b : a;
    return CHL_OK;
  }
  int dims[3] = {1, 1, 1}, number = nproc, index = 0;
  while (number > 1) {
    int const gpf = greatest_prime_factor(number);
    number /= gpf;
    dims[index % 3] *= gpf;
    index++;
  }
  for (int i = 0; i < 3; i++)  // np_x >= np_y >= np_z
    for (int j = 0; j + 1 < 3; j++)
      if (dims[j + 1] > dims[j]) {
        int const t = dims[j];
        dims[j] = dims[j + 1];
        dims[j + 1] = t;
      }
  np[0] = dims[0];
  np[1] = dims[1];
  np[2] = dims[2];
  if (nzg > nxg) {  // mpi_routines.cpp:299-304
    int const t = np[0];
    np[0] = np[2];
    np[2] = t;
  }
  return CHL_OK;
}

int chl_local_extent(int n_global, int np, int coord, int *n_local, int *offset)
{
  if (np < 1 || coord < 0 || coord >= np || !n_local || !offset)
    return fail(CHL_ERR_INVALID, "chl_local_extent: bad argument");
  int const rem = n_global % np, base = n_global / np;
  if (rem == 0) {
    *n_local = base;
    *offset = coord * base;
  } else if (coord < rem) {
    *n_local = base + 1;
    *offset = coord * (base + 1);
  } else {
    *n_local = base;
    *offset = rem * (base + 1) + (coord - rem) * base;
  }
  return CHL_OK;
}

int chl_debug_exact_div(const double *d_a, const double *d_b, double *d_q_shared, double *d_q_plain, long n,
                        void *stream)
{
  if (!d_a || !d_b || !d_q_shared || !d_q_plain || n < 0) return fail(CHL_ERR_INVALID, "chl_debug_exact_div: bad argument");
  cudaError_t const e = chl::launch_exact_div_test(d_a, d_b, d_q_shared, d_q_plain, n, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(CHL_ERR_CUDA, "chl_debug_exact_div", e);
  return CHL_OK;
}

int chl_debug_exact_div_flagged(const double *d_a, const double *d_b, double *d_q, int *d_flag, int zero_ok, long n,
                                void *stream)
{
  if (!d_a || !d_b || !d_q || !d_flag || n < 0) return fail(CHL_ERR_INVALID, "chl_debug_exact_div_flagged: bad argument");
  cudaError_t const e = chl::launch_exact_div_flag_test(d_a, d_b, d_q, d_flag, zero_ok, n, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(CHL_ERR_CUDA, "chl_debug_exact_div_flagged", e);
  return CHL_OK;
}

int chl_debug_exact_sqrt(const double *d_x, double *d_r_own, double *d_r_plain, long n, void *stream)
{
  if (!d_x || !d_r_own || !d_r_plain || n < 0) return fail(CHL_ERR_INVALID, "chl_debug_exact_sqrt: bad argument");
  cudaError_t e = chl::launch_exact_sqrt_test(d_x, d_r_own, d_r_plain, n, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(CHL_ERR_CUDA, "chl_debug_exact_sqrt", e);
  return CHL_OK;
}

}  // extern "C"
