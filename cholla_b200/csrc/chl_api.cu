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
  void *d_acc;                // direction-split sweeps: running sum of flux differences, 5 * n_cells reals
  chl::MhdOps const *mhd;     // non-null in an MHD context
  void *d_prim;               // MHD: half-step primitives + cell-centred fields
  void *d_ctf;                // MHD: 9 * n_cells flux components kept for constrained transport
  void *d_E;                  // MHD: 3 * n_cells edge EMFs
  int split;                  // 1 = one launch per direction (default), 0 = single fused launch per stage
  int dir_chunk;              // CHL_DIR_CHUNK override (0 = auto)
  unsigned long long *d_key;  // CFL accumulator (ordered key)
  unsigned long long *h_key;  // pinned
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
  } else {
    if (cfg->n_fields != 5) return fail(CHL_ERR_UNSUPPORTED, "chl_create: a hydro grid has n_fields = 5");
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
  if (cfg->precision == 2 || cfg->precision == 0) {
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
  c->mhd = mhd;
  c->d_prim = c->d_ctf = c->d_E = nullptr;
  {
    // development knobs: CHL_SWEEP_IMPL=fused|split selects the kernel family, CHL_DIR_CHUNK the marching length
    const char *impl = getenv("CHL_SWEEP_IMPL");
    c->split = !(impl && strcmp(impl, "fused") == 0) || cfg->solver >= CHL_SOLVER_HLL ||
               cfg->recon >= CHL_RECON_PLMP;  // HLL, exact, PLMP, PPMP: split sweeps only
    const char *ch = getenv("CHL_DIR_CHUNK");
    c->dir_chunk = ch ? atoi(ch) : 0;
  }
  c->d_key = nullptr;
  c->h_key = nullptr;
  c->profiling = 0;
  c->ev_steps = 0;
  size_t const bytes = c->n_cells * (mhd ? 8 : ops->half_fields) * ops->real_bytes;
  cudaError_t e = cudaMalloc(&c->d_half, bytes);
  if (e == cudaSuccess && (c->split || mhd)) e = cudaMalloc(&c->d_acc, c->n_cells * 5 * ops->real_bytes);
  if (mhd) {
    size_t const w = c->n_cells * ops->real_bytes;
    if (e == cudaSuccess) e = cudaMalloc(&c->d_prim, mhd->prim_fields * w);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_ctf, 9 * w);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_E, 6 * w);  // three edge EMFs + three cell-centred reference EMFs
    if (e == cudaSuccess) e = cudaMemset(c->d_prim, 0, mhd->prim_fields * w);
    if (e == cudaSuccess) e = cudaMemset(c->d_ctf, 0, 9 * w);
    if (e == cudaSuccess) e = cudaMemset(c->d_E, 0, 6 * w);
  }
  if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_key, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMallocHost((void **)&c->h_key, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(c->d_half, 0, bytes);
  if (e == cudaSuccess) e = cudaMemset(c->d_key, 0, sizeof(unsigned long long));
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
  size_t fields = c->mhd ? 8 + 5 + c->mhd->prim_fields + 9 + 6 : c->ops->half_fields + (c->split ? 5 : 0);
  return c->n_cells * fields * c->ops->real_bytes + sizeof(unsigned long long);
}

static int step_impl(chl_ctx *c, void *U, double dt, int fused, cudaStream_t s)
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
      if (g.temperature_floor_U > 0.0)
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
  a.acc = c->d_acc;
  a.dir_chunk = c->dir_chunk;
  a.simple = 0;
  auto const sweep = c->split ? c->ops->sweep_split : c->ops->sweep;
  int const n[3] = {g.nx, g.ny, g.nz};
  auto choose_chunk = [&](int const *lo, int const *hi) {
    int const ty = c->ops->tile_y;
    long const bxy = (long)((hi[0] - lo[0] + 31) / 32) * ((hi[1] - lo[1] + ty - 1) / ty);
    int const oz = hi[2] - lo[2];
    long chunks = (2368 + bxy - 1) / bxy;  // aim at >= 8 waves of 2 CTAs on 148 SMs
    if (chunks > oz / 8) chunks = oz / 8;
    if (chunks < 1) chunks = 1;
    return (int)((oz + chunks - 1) / chunks);
  };
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
  a.zchunk = a.has_z ? choose_chunk(a.lo, a.hi) : 1;
  bool const prof = c->profiling && c->ev_steps < kMaxProfiledSteps;
  if (prof) {
    while ((int)c->ev.size() < 3 * (c->ev_steps + 1)) {
      cudaEvent_t e;
      CHL_CUDA(cudaEventCreate(&e));
      c->ev.push_back(e);
    }
    CHL_CUDA(cudaEventRecord(c->ev[3 * c->ev_steps], s));
  }
  CHL_CUDA(sweep(0, g.recon, g.solver, a, s));
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
  if (fused) {
    a.apply_tfloor = (g.temperature_floor_U > 0.0 && c->ndim == 3) ? 1 : 0;
    a.U_floor = g.temperature_floor_U;
    a.do_dti = 1;
    CHL_CUDA(cudaMemsetAsync(c->d_key, 0, sizeof(unsigned long long), s));
  }
  a.zchunk = a.has_z ? choose_chunk(a.lo, a.hi) : 1;
  CHL_CUDA(sweep(1, g.recon, g.solver, a, s));
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
  if (g.recon != CHL_RECON_PCM && g.recon != CHL_RECON_PLMC)
    return fail(CHL_ERR_UNSUPPORTED, "chl_simple_step: PCM and PLMC only (the CTU forms of PLMP / PPMP / PPMC are not built)");
  if (!c->d_acc) return fail(CHL_ERR_UNSUPPORTED, "chl_simple_step: needs the direction-split sweeps (CHL_SWEEP_IMPL=split)");
  chl::SweepArgs a;
  a.nx = g.nx; a.ny = g.ny; a.nz = g.nz;
  a.has_y = g.ny > 1;
  a.has_z = g.nz > 1;
  a.gamma = g.gamma;
  a.density_floor = g.density_floor;
  a.dx = g.dx; a.dy = g.dy; a.dz = g.dz;
  a.dti_key = c->d_key;
  a.acc = c->d_acc;
  a.dir_chunk = c->dir_chunk;
  a.simple = 1;
  a.zchunk = 1;
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
    a.apply_tfloor = (g.temperature_floor_U > 0.0 && c->ndim == 3) ? 1 : 0;
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

// ---- host-only mirrors of the reference's decomposition logic (src/mpi/mpi_routines.cpp) ----
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

int chl_debug_exact_sqrt(const double *d_x, double *d_r_own, double *d_r_plain, long n, void *stream)
{
  if (!d_x || !d_r_own || !d_r_plain || n < 0) return fail(CHL_ERR_INVALID, "chl_debug_exact_sqrt: bad argument");
  cudaError_t e = chl::launch_exact_sqrt_test(d_x, d_r_own, d_r_plain, n, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(CHL_ERR_CUDA, "chl_debug_exact_sqrt", e);
  return CHL_OK;
}

}  // extern "C"
