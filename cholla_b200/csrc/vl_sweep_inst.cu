// vl_sweep_inst.cu -- instantiates the hydro sweep kernels for ONE arithmetic mode.  The build compiles this
// file three times:   -DCHL_STRICT=1 -DCHL_REAL=double -fmad=false   (mode f64_strict)
//                     -DCHL_STRICT=0 -DCHL_REAL=double               (mode f64_fast)
//                     -DCHL_STRICT=0 -DCHL_REAL=float                (mode f32_fast)
// CHL_MODE names the exported factory.
#include <float.h>
#include <string.h>

#include "chl_internal.h"
#include "vl_dirsweep.cuh"
#include "vl_predictor.cuh"

#ifndef CHL_MODE
#error "CHL_MODE must be defined (f64_strict | f64_fast | f32_fast)"
#endif
#define CHL_CAT2(a, b) a##b
#define CHL_CAT(a, b) CHL_CAT2(a, b)

namespace CHL_NS
{
using chl::ModeOps;
using chl::SweepArgs;
namespace
{
// ---- direction-split sweeps (vl_dirsweep.cuh) ----
template <int RECON, int SOLVER, int STAGE, int DIR, bool FIRST, bool LAST>
cudaError_t launch_dir(DirParams const &p, cudaStream_t s)
{
  if (p.nbox > 0 && DIR != 0) {  // several boxes, one launch: the table was filled by the caller
    DirParams::Box const &last = p.box[p.nbox - 1];
    int const total = last.cta0 + last.gx * last.gy * ((last.hi[DIR] - last.lo[DIR] + last.chunk - 1) / last.chunk);
    vl_dir_kernel<RECON, SOLVER, STAGE, DIR, FIRST, LAST><<<total, kDirThreads, 0, s>>>(p);
    return cudaGetLastError();
  }
  int const ox = p.hi[0] - p.lo[0], oy = p.hi[1] - p.lo[1], oz = p.hi[2] - p.lo[2];
  if (ox <= 0 || oy <= 0 || oz <= 0) return cudaSuccess;
  dim3 grid;
  DirParams q = p;
  if (DIR == 0) grid = dim3((oy + kRowsPerCta - 1) / kRowsPerCta, oz, 1);
  else {
    int const om = DIR == 1 ? oy : oz, oo = DIR == 1 ? oz : oy;
    q.xw_log2 = pick_xw_log2(ox, oo);
    int const xw = 1 << q.xw_log2, ow = kDirThreads >> q.xw_log2;
    grid = dim3((ox + xw - 1) / xw, (oo + ow - 1) / ow, (om + p.chunk - 1) / p.chunk);
  }
  vl_dir_kernel<RECON, SOLVER, STAGE, DIR, FIRST, LAST><<<grid, kDirThreads, 0, s>>>(q);
  return cudaGetLastError();
}

template <int RECON, int SOLVER, int STAGE>
cudaError_t run_dirs(DirParams p, SweepArgs const &a, cudaStream_t s)
{
  Real const dto[3] = {(Real)a.dtodx, (Real)a.dtody, (Real)a.dtodz};
  auto chunk_for = [&](int d) {
    int const o = p.hi[d] - p.lo[d];
    if (a.dir_chunk > 0) return a.dir_chunk < o ? a.dir_chunk : o;  // caller override (tests, tuning)
    if (o >= 384) return 64;  // 512^3: 64 vs 32 vs 128 cells per thread are within 1 % of each other
    if (o >= 64) return 32;   // smaller grids need the extra CTAs to fill 148 SMs x 8 CTAs
    int const c = (o + 1) / 2;
    return c < 8 ? (o < 8 ? o : 8) : c;
  };
  cudaError_t e;
  // a.part (3-D only; 1-D / 2-D grids run whole in part 1): 0 = the whole stage; 1 = every pass except the interior box
  // of the last one -- the last pass then completes the boundary shell (the `shell` outermost layers of the output
  // region, the cells a neighbouring rank's ghost layers mirror) as six thin boxes; 2 = the interior box of the last
  // pass.  The caller starts the halo exchange between part 1 and part 2 (chl_api.cu, overlapped step).
  bool const split = a.part != 0 && p.ndim == 3 && a.shell > 0 && p.hi[0] - p.lo[0] > 2 * a.shell &&
                     p.hi[1] - p.lo[1] > 2 * a.shell && p.hi[2] - p.lo[2] > 2 * a.shell;
  if (a.part == 2 && !split) return cudaSuccess;
  if (a.part != 2) {
    if (p.ndim == 1) {
      p.dto = dto[0];
      return launch_dir<RECON, SOLVER, STAGE, 0, true, true>(p, s);
    }
    p.dto = dto[0];
    if ((e = launch_dir<RECON, SOLVER, STAGE, 0, true, false>(p, s)) != cudaSuccess) return e;
    p.dto   = dto[1];
    p.chunk = chunk_for(1);
    if (p.ndim == 2) return launch_dir<RECON, SOLVER, STAGE, 1, false, true>(p, s);
    if ((e = launch_dir<RECON, SOLVER, STAGE, 1, false, false>(p, s)) != cudaSuccess) return e;
  }
  p.dto = dto[2];
  if (!split) {
    p.chunk = chunk_for(2);
    return launch_dir<RECON, SOLVER, STAGE, 2, false, true>(p, s);
  }
  int const lo[3] = {p.lo[0], p.lo[1], p.lo[2]}, hi[3] = {p.hi[0], p.hi[1], p.hi[2]}, sh = a.shell;
  if (a.part == 2) {
    p.lo[0] = lo[0] + sh; p.hi[0] = hi[0] - sh; p.lo[1] = lo[1] + sh; p.hi[1] = hi[1] - sh;
    p.lo[2] = lo[2] + sh; p.hi[2] = hi[2] - sh;
    p.chunk = chunk_for(2);
    return launch_dir<RECON, SOLVER, STAGE, 2, false, true>(p, s);
  }
  // the six thin boxes of the shell in ONE launch (DirParams::box)
  int cta = 0;
  p.nbox  = 0;
  auto box = [&](int x0, int x1, int y0, int y1, int z0, int z1, int xw_log2) {
    DirParams::Box &B = p.box[p.nbox++];
    B.lo[0] = x0; B.hi[0] = x1; B.lo[1] = y0; B.hi[1] = y1; B.lo[2] = z0; B.hi[2] = z1;
    B.xw_log2 = xw_log2;
    int const xw = 1 << xw_log2, ow = kDirThreads >> xw_log2;
    B.gx = (x1 - x0 + xw - 1) / xw;
    B.gy = (y1 - y0 + ow - 1) / ow;
    // few columns: march short chunks so that the launch fills 148 SMs x 7-8 CTAs (a chunk of c cells reconstructs c + 2,
    // so 8 is the floor; the z slabs are `sh` thick and take what they are)
    B.chunk = z1 - z0 < 8 ? z1 - z0 : 8;
    B.cta0  = cta;
    cta += B.gx * B.gy * ((z1 - z0 + B.chunk - 1) / B.chunk);
  };
  int xs = 0;
  while ((1 << xs) < sh) xs++;  // the x shells are `sh` cells wide: 2^xs x (128 >> xs) threads per CTA
  box(lo[0], hi[0], lo[1], hi[1], lo[2], lo[2] + sh, 7);                        // -z
  box(lo[0], hi[0], lo[1], hi[1], hi[2] - sh, hi[2], 7);                        // +z
  box(lo[0], hi[0], lo[1], lo[1] + sh, lo[2] + sh, hi[2] - sh, 7);              // -y
  box(lo[0], hi[0], hi[1] - sh, hi[1], lo[2] + sh, hi[2] - sh, 7);              // +y
  box(lo[0], lo[0] + sh, lo[1] + sh, hi[1] - sh, lo[2] + sh, hi[2] - sh, xs);   // -x
  box(hi[0] - sh, hi[0], lo[1] + sh, hi[1] - sh, lo[2] + sh, hi[2] - sh, xs);   // +x
  return launch_dir<RECON, SOLVER, STAGE, 2, false, true>(p, s);
}

// ---- fused predictor (vl_predictor.cuh): 3-D grids, one launch ----
template <int SOLVER>
cudaError_t launch_predictor(SweepArgs const &a, int plmp, cudaStream_t s)
{
  PredParams p;
  p.U = (const Real *)a.in;
  p.W = (Real *)a.out;
  p.nx = a.nx; p.ny = a.ny; p.nz = a.nz;
  for (int d = 0; d < 3; d++) { p.lo[d] = a.lo[d]; p.hi[d] = a.hi[d]; }
  p.dto[0] = (Real)a.dtodx; p.dto[1] = (Real)a.dtody; p.dto[2] = (Real)a.dtodz;
  p.gamma = (Real)a.gamma; p.density_floor = (Real)a.density_floor;
  p.plmp = plmp;
  int const ox = p.hi[0] - p.lo[0], oy = p.hi[1] - p.lo[1], oz = p.hi[2] - p.lo[2];
  if (ox <= 0 || oy <= 0 || oz <= 0) return cudaSuccess;
  int const bx = (ox + kPTX - 1) / kPTX, by = (oy + kPTY - 1) / kPTY;
  // planes per CTA: long marches amortise the two extra planes a chunk converts, but the grid must still fill
  // 148 SMs x 2 CTAs several times over
  int zchunk = a.dir_chunk > 0 ? a.dir_chunk : 64;
  while (zchunk > 8 && (long)bx * by * ((oz + zchunk - 1) / zchunk) < 148 * 2 * 6) zchunk /= 2;
  if (zchunk > oz) zchunk = oz;
  p.zchunk = zchunk;
  static bool configured = false;  // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(vl_predictor_kernel<SOLVER>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kPSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 const grid(bx, by, (oz + zchunk - 1) / zchunk);
  vl_predictor_kernel<SOLVER><<<grid, kPThreads, kPSmemBytes, s>>>(p);
  return cudaGetLastError();
}

cudaError_t sweep_split(int stage, int recon, int solver, SweepArgs const &a, cudaStream_t s)
{
#if CHL_NP == 0
  if (stage == 0 && a.has_y && a.has_z && a.fused_predictor) {
    int const plmp = half_kind(recon);
#ifdef CHL_DEV_SUBSET
    if (solver == kHLLC) return launch_predictor<kHLLC>(a, plmp, s);
#else
    if (solver == kROE) return launch_predictor<kROE>(a, plmp, s);
    if (solver == kHLL) return launch_predictor<kHLL>(a, plmp, s);
    if (solver == kEXACT) return launch_predictor<kEXACT>(a, plmp, s);
    return launch_predictor<kHLLC>(a, plmp, s);
#endif
  }
#endif
  DirParams p;
  p.in  = (const Real *)a.in;
  p.U   = (const Real *)a.in;
  p.acc = (Real *)a.acc;
  p.out = (Real *)a.out;
  p.nx = a.nx; p.ny = a.ny; p.nz = a.nz;
  for (int d = 0; d < 3; d++) { p.lo[d] = a.lo[d]; p.hi[d] = a.hi[d]; }
  p.chunk = 1;
  p.ndim  = 1 + (a.has_y ? 1 : 0) + (a.has_z ? 1 : 0);
  p.dto   = 0;
  p.gamma = (Real)a.gamma; p.density_floor = (Real)a.density_floor;
  p.apply_tfloor = a.apply_tfloor; p.U_floor = (Real)a.U_floor; p.do_dti = a.do_dti;
  p.dx = (Real)a.dx; p.dy = (Real)a.dy; p.dz = (Real)a.dz;
  p.idx = (Real)(1.0 / a.dx); p.idy = (Real)(1.0 / a.dy); p.idz = (Real)(1.0 / a.dz);
  p.dti_key   = a.dti_key;
  p.bad_cells = a.bad_cells;
  p.xw_log2   = 7;
  p.nbox      = 0;
  p.plmp      = half_kind(recon);
  p.de_term   = (const Real *)a.de_term;
#if CHL_NP > 0
  // -DSCALAR / -DDE builds: the Van Leer integrator with PCM / PLMC / PPMC and HLLC / Roe / HLL, PLMP / PPMP with HLLC
  if (a.simple || (solver != kHLLC && solver != kROE && solver != kHLL)) return cudaErrorNotSupported;
  if (CHL_DE && !p.de_term) return cudaErrorInvalidValue;
  if (stage == 0) {
    if (solver == kROE) return run_dirs<kPCM, kROE, 0>(p, a, s);
    if (solver == kHLL) return run_dirs<kPCM, kHLL, 0>(p, a, s);
    return run_dirs<kPCM, kHLLC, 0>(p, a, s);
  }
  switch (recon * 4 + solver) {
    case kPLMP * 4 + kHLLC: return run_dirs<kPLMP, kHLLC, 1>(p, a, s);  // builds/make.type.cooling: PPMP + HLLC + DE
    case kPPMP * 4 + kHLLC: return run_dirs<kPPMP, kHLLC, 1>(p, a, s);
    case kPCM * 4 + kHLLC: return run_dirs<kPCM, kHLLC, 1>(p, a, s);
    case kPCM * 4 + kROE: return run_dirs<kPCM, kROE, 1>(p, a, s);
    case kPCM * 4 + kHLL: return run_dirs<kPCM, kHLL, 1>(p, a, s);
    case kPLMC * 4 + kHLLC: return run_dirs<kPLMC, kHLLC, 1>(p, a, s);
    case kPLMC * 4 + kROE: return run_dirs<kPLMC, kROE, 1>(p, a, s);
    case kPLMC * 4 + kHLL: return run_dirs<kPLMC, kHLL, 1>(p, a, s);
    case kPPMC * 4 + kHLLC: return run_dirs<kPPMC, kHLLC, 1>(p, a, s);
    case kPPMC * 4 + kROE: return run_dirs<kPPMC, kROE, 1>(p, a, s);
    case kPPMC * 4 + kHLL: return run_dirs<kPPMC, kHLL, 1>(p, a, s);
  }
  return cudaErrorNotSupported;
#elif defined(CHL_DEV_SUBSET)  // development builds (make VARIANT=... EXTRA=-DCHL_DEV_SUBSET): the headline combination only
  if (solver != kHLLC || recon != kPPMC) return cudaErrorNotSupported;
  if (stage == 0) return run_dirs<kPCM, kHLLC, 0>(p, a, s);
  return run_dirs<kPPMC, kHLLC, 1>(p, a, s);
#else
  if (stage == 0) {
    if (solver == kROE) return run_dirs<kPCM, kROE, 0>(p, a, s);
    if (solver == kHLL) return run_dirs<kPCM, kHLL, 0>(p, a, s);
    if (solver == kEXACT) return run_dirs<kPCM, kEXACT, 0>(p, a, s);
    return run_dirs<kPCM, kHLLC, 0>(p, a, s);
  }
  if (solver == kEXACT) {  // -DEXACT: PCM, PLMC and PPMC under the Van Leer integrator, PCM and PLMC under SIMPLE
    if (recon == kPLMP || recon == kPPMP || (a.simple && recon == kPPMC)) return cudaErrorNotSupported;
    if (a.simple && recon == kPLMC) return run_dirs<kPLMCT, kEXACT, 1>(p, a, s);
    if (recon == kPCM) return run_dirs<kPCM, kEXACT, 1>(p, a, s);
    if (recon == kPLMC) return run_dirs<kPLMC, kEXACT, 1>(p, a, s);
    return run_dirs<kPPMC, kEXACT, 1>(p, a, s);
  }
  if (a.simple) {  // Simple_Algorithm_*_CUDA: PCM as is, PLMC with characteristic tracing; the CTU forms are not built
    if (recon == kPLMC) {
      if (solver == kROE) return run_dirs<kPLMCT, kROE, 1>(p, a, s);
      if (solver == kHLL) return run_dirs<kPLMCT, kHLL, 1>(p, a, s);
      return run_dirs<kPLMCT, kHLLC, 1>(p, a, s);
    }
    if (recon != kPCM) return cudaErrorNotSupported;
  }
  if (recon == kPPMP) {
    if (solver == kROE) return run_dirs<kPPMP, kROE, 1>(p, a, s);
    if (solver == kHLL) return run_dirs<kPPMP, kHLL, 1>(p, a, s);
    return run_dirs<kPPMP, kHLLC, 1>(p, a, s);
  }
  if (recon == kPLMP) {
    if (solver == kROE) return run_dirs<kPLMP, kROE, 1>(p, a, s);
    if (solver == kHLL) return run_dirs<kPLMP, kHLL, 1>(p, a, s);
    return run_dirs<kPLMP, kHLLC, 1>(p, a, s);
  }
  if (solver == kHLL) {
    if (recon == kPCM) return run_dirs<kPCM, kHLL, 1>(p, a, s);
    if (recon == kPLMC) return run_dirs<kPLMC, kHLL, 1>(p, a, s);
    return run_dirs<kPPMC, kHLL, 1>(p, a, s);
  }
  switch (recon * 2 + (solver == kROE ? 1 : 0)) {
    case kPCM * 2 + 0: return run_dirs<kPCM, kHLLC, 1>(p, a, s);
    case kPCM * 2 + 1: return run_dirs<kPCM, kROE, 1>(p, a, s);
    case kPLMC * 2 + 0: return run_dirs<kPLMC, kHLLC, 1>(p, a, s);
    case kPLMC * 2 + 1: return run_dirs<kPLMC, kROE, 1>(p, a, s);
    case kPPMC * 2 + 0: return run_dirs<kPPMC, kHLLC, 1>(p, a, s);
    case kPPMC * 2 + 1: return run_dirs<kPPMC, kROE, 1>(p, a, s);
  }
  return cudaErrorInvalidValue;
#endif
}

// Calc_dt_{1,2,3}D (hydro_cuda.cu:437-560): grid-stride max over real cells, warp shuffle, one atomic per CTA
template <int NDIM>
__global__ void __launch_bounds__(256) calc_dti_kernel(const Real *__restrict__ U, int nx, int ny, int nz, int g,
                                                       Real dx, Real dy, Real dz, Real gamma,
                                                       unsigned long long *key)
{
  __shared__ double sred[8];
  size_t const N = (size_t)nx * ny * nz;
  int const rx = nx - 2 * g, ry = NDIM > 1 ? ny - 2 * g : 1, rz = NDIM > 2 ? nz - 2 * g : 1;
  size_t const nreal = (size_t)rx * ry * rz;
  Real const idx = RC(1.0) / dx, idy = RC(1.0) / dy, idz = RC(1.0) / dz;
  KC const kc = make_kc(gamma);
  double m = -DBL_MAX;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < nreal; t += (size_t)gridDim.x * blockDim.x) {
    int const x = (int)(t % rx) + g;
    size_t const r = t / rx;
    int const y = NDIM > 1 ? (int)(r % ry) + g : 0;
    int const z = NDIM > 2 ? (int)(r / ry) + g : 0;
    size_t const id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
    m = fmax(m, (double)inverse_crossing_time<NDIM>(__ldg(U + id), __ldg(U + N + id), __ldg(U + 2 * N + id),
                                                    __ldg(U + 3 * N + id), __ldg(U + 4 * N + id), idx, idy, idz, dx,
                                                    dy, dz, kc));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < 8 ? sred[threadIdx.x] : -DBL_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) atomicMax(key, dti_encode(m));
  }
}

cudaError_t calc_dti(const void *U, int nx, int ny, int nz, int g, double dx, double dy, double dz, double gamma,
                     unsigned long long *key, cudaStream_t s)
{
  int ndim = 1 + (ny > 1) + (nz > 1);
  if (nx == 1 && ny == 1 && nz == 1) ndim = 3;  // the reference's unit test drives Calc_dt_3D on a single cell
  size_t nreal = (size_t)(nx - 2 * g) * (ndim > 1 ? ny - 2 * g : 1) * (ndim > 2 ? nz - 2 * g : 1);
  int blocks = (int)((nreal + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  const Real *u = (const Real *)U;
  if (ndim == 3) calc_dti_kernel<3><<<blocks, 256, 0, s>>>(u, nx, ny, nz, g, (Real)dx, (Real)dy, (Real)dz, (Real)gamma, key);
  else if (ndim == 2) calc_dti_kernel<2><<<blocks, 256, 0, s>>>(u, nx, ny, nz, g, (Real)dx, (Real)dy, (Real)dz, (Real)gamma, key);
  else calc_dti_kernel<1><<<blocks, 256, 0, s>>>(u, nx, ny, nz, g, (Real)dx, (Real)dy, (Real)dz, (Real)gamma, key);
  return cudaGetLastError();
}

// Temperature_Floor_Kernel (hydro_cuda.cu:1117-1153); its guard makes it a no-op unless nz > 2*n_ghost
__global__ void __launch_bounds__(256) tfloor_kernel(Real *U, int nx, int ny, int nz, int g, Real U_floor)
{
  size_t const N = (size_t)nx * ny * nz;
  int const rx = nx - 2 * g, ry = ny - 2 * g, rz = nz - 2 * g;
  size_t const nreal = (size_t)rx * ry * rz;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < nreal; t += (size_t)gridDim.x * blockDim.x) {
    int const x = (int)(t % rx) + g;
    size_t const r = t / rx;
    int const y = (int)(r % ry) + g, z = (int)(r / ry) + g;
    size_t const id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
    Real const d = U[id], d_inv = RC(1.0) / d;
    Real const vx = U[N + id] * d_inv, vy = U[2 * N + id] * d_inv, vz = U[3 * N + id] * d_inv;
    Real const E = U[4 * N + id];
    Real const Ekin = RC(0.5) * d * (vx * vx + vy * vy + vz * vz);
    if ((E - Ekin) / d < U_floor) U[4 * N + id] = Ekin + d * U_floor;
#if CHL_DE  // hydro_cuda.cu:1146-1151
    size_t const ge = (size_t)(5 + kNP - 1) * N + id;
    if (U[ge] / d < U_floor) U[ge] = d * U_floor;
#endif
  }
}

cudaError_t tfloor(void *U, int nx, int ny, int nz, int g, double U_floor, cudaStream_t s)
{
  if (nx <= 2 * g || ny <= 2 * g || nz <= 2 * g) return cudaSuccess;
  size_t nreal = (size_t)(nx - 2 * g) * (ny - 2 * g) * (nz - 2 * g);
  int blocks = (int)((nreal + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  tfloor_kernel<<<blocks, 256, 0, s>>>((Real *)U, nx, ny, nz, g, (Real)U_floor);
  return cudaGetLastError();
}

// SIMPLE integrator (simple_3D_cuda.cu): there is no predictor, the reconstruction reads U itself.  The sweeps above
// read primitives, so one elementwise pass derives them for every cell, ghosts included (what Load_Data would do per
// stencil access): 40 B read + 40 (56 strict) B written per cell.
__global__ void __launch_bounds__(256) prim_from_cons_kernel(const Real *__restrict__ U, Real *__restrict__ W, size_t N,
                                                             Real gamma, int plmp)
{
  KH const kc = make_kct<KH>(gamma);
  for (size_t id = blockIdx.x * (size_t)blockDim.x + threadIdx.x; id < N; id += (size_t)gridDim.x * blockDim.x) {
    U5 c{{__ldg(U + id), __ldg(U + N + id), __ldg(U + 2 * N + id), __ldg(U + 3 * N + id), __ldg(U + 4 * N + id)}};
#if CHL_NP > 0
#pragma unroll
    for (int i = 0; i < kNP; i++) c.q[i] = __ldg(U + (size_t)(5 + i) * N + id);
#endif
    store_half(W, N, id, half_state_checked(c, kc, plmp));
  }
}

cudaError_t prim_from_cons(const void *U, void *W, size_t n_cells, double gamma, int recon, cudaStream_t s)
{
  size_t blocks = (n_cells + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  prim_from_cons_kernel<<<(unsigned)blocks, 256, 0, s>>>((const Real *)U, (Real *)W, n_cells, (Real)gamma,
                                                        half_kind(recon));
  return cudaGetLastError();
}

#if CHL_DE
// ---- dual energy (-DDE): the three small passes around the sweeps.  Plain IEEE divisions (cold code; the strict TU is
// built -fmad=false, so every operation rounds like the reference's). ----
__device__ __forceinline__ Real de_pressure_ieee(Real E, Real U_total, Real U_advected, Real gamma)
{  // Get_Pressure_From_DE, hydro_utilities.h:136-151
  Real const U = (U_total / E > kDeEta1) ? U_total : U_advected;
  return fmax(U * (gamma - RC(1.0)), kTiny);
}
// mode 0, before the predictor: out = 0.5 P (dt/2dx (vx_imo - vx_ipo) + ...), the last term of the half-step GasEnergy
//   (VL_3D_cuda.cu:408-426, 471-477; the 1-D / 2-D kernels take the plain pressure: VL_2D_cuda.cu:213, VL_1D_cuda.cu:183);
// mode 1, before the corrector: out = GasEnergy + 0.5 P (dt/dx (vx_imo - vx_ipo) + ...), i.e. GasEnergy after
//   Partial_Update_Advected_Internal_Energy_{1,2,3}D (hydro_cuda.cu:662-833).  Both read U^n only.
template <int NDIM>
__global__ void __launch_bounds__(256) de_term_kernel(const Real *__restrict__ U, Real *__restrict__ out, int nx, int ny,
                                                      int nz, int lo0, int lo1, int lo2, int hi0, int hi1, int hi2,
                                                      Real dtodx, Real dtody, Real dtodz, Real gamma, int mode)
{
  size_t const N = (size_t)nx * ny * nz;
  int const rx = hi0 - lo0, ry = hi1 - lo1, rz = hi2 - lo2;
  size_t const n = (size_t)rx * ry * rz;
  size_t const sy = nx, sz = (size_t)nx * ny, GE = (size_t)(5 + kNP - 1) * N;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    int const x = (int)(t % rx) + lo0;
    size_t const r = t / rx;
    int const y = (int)(r % ry) + lo1, z = (int)(r / ry) + lo2;
    size_t const id = (size_t)x + sy * y + sz * z;
    Real const d = U[id], d_inv = RC(1.0) / d;
    Real const vx = U[N + id] * d_inv, vy = U[2 * N + id] * d_inv, vz = U[3 * N + id] * d_inv;
    Real const E = U[4 * N + id], ge = U[GE + id];
    Real P;
    if (mode == 0 && NDIM < 3) {
      P = (E - RC(0.5) * d * (vx * vx + vy * vy + vz * vz)) * (gamma - RC(1.0));
    } else {
      Real const E_kin = NDIM == 3 ? RC(0.5) * d * sq3(vx, vy, vz) : RC(0.5) * d * (vx * vx + vy * vy + vz * vz);
      P = fmax(de_pressure_ieee(E, E - E_kin, ge, gamma), kTiny);
    }
    Real div = dtodx * (U[N + id - 1] / U[id - 1] - U[N + id + 1] / U[id + 1]);
    if (NDIM > 1) div = div + dtody * (U[2 * N + id - sy] / U[id - sy] - U[2 * N + id + sy] / U[id + sy]);
    if (NDIM > 2) div = div + dtodz * (U[3 * N + id - sz] / U[id - sz] - U[3 * N + id + sz] / U[id + sz]);
    Real const term = RC(0.5) * P * (div);
    out[id] = mode == 0 ? term : ge + term;
  }
}

cudaError_t de_term(const void *U, void *out, int nx, int ny, int nz, const int lo[3], const int hi[3], double dtodx,
                    double dtody, double dtodz, double gamma, int mode, cudaStream_t s)
{
  int const ndim = 1 + (ny > 1) + (nz > 1);
  size_t const n = (size_t)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
  if (n == 0) return cudaSuccess;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
#define CHL_DE_TERM(ND)                                                                                              \
  de_term_kernel<ND><<<(unsigned)blocks, 256, 0, s>>>((const Real *)U, (Real *)out, nx, ny, nz, lo[0], lo[1], lo[2], \
                                                      hi[0], hi[1], hi[2], (Real)dtodx, (Real)dtody, (Real)dtodz,    \
                                                      (Real)gamma, mode)
  if (ndim == 3) CHL_DE_TERM(3);
  else if (ndim == 2) CHL_DE_TERM(2);
  else CHL_DE_TERM(1);
#undef CHL_DE_TERM
  return cudaGetLastError();
}

// Select_Internal_Energy_{1,2,3}D (hydro_cuda.cu:835-1010): which internal energy the cell keeps -- decided from the
// total energies of the cell and its neighbours (clamped to the real cells) BEFORE any of them is re-synchronised, so it
// is a launch of its own; writes GasEnergy only.
template <int NDIM>
__global__ void __launch_bounds__(256) de_select_kernel(Real *__restrict__ U, int nx, int ny, int nz, int g)
{
  size_t const N = (size_t)nx * ny * nz;
  int const rx = nx - 2 * g, ry = NDIM > 1 ? ny - 2 * g : 1, rz = NDIM > 2 ? nz - 2 * g : 1;
  size_t const n = (size_t)rx * ry * rz;
  size_t const sy = nx, sz = (size_t)nx * ny, GE = (size_t)(5 + kNP - 1) * N;
  const Real *__restrict__ Et = U + 4 * N;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    int const x = (int)(t % rx) + g;
    size_t const r = t / rx;
    int const y = NDIM > 1 ? (int)(r % ry) + g : 0, z = NDIM > 2 ? (int)(r / ry) + g : 0;
    size_t const id = (size_t)x + sy * y + sz * z;
    Real const d = U[id], d_inv = RC(1.0) / d;
    Real const vx = U[N + id] * d_inv, vy = U[2 * N + id] * d_inv, vz = U[3 * N + id] * d_inv;
    Real const E = Et[id], U_advected = U[GE + id];
    Real const U_total = E - RC(0.5) * d * (vx * vx + vy * vy + vz * vz);
    Real Emax = fmax(Et[id - (x - 1 >= g ? 1 : 0)], E);
    Emax      = fmax(Emax, Et[id + (x + 1 <= nx - g - 1 ? 1 : 0)]);
    if (NDIM > 1) {
      Emax = fmax(Emax, Et[id - (y - 1 >= g ? sy : 0)]);
      Emax = fmax(Emax, Et[id + (y + 1 <= ny - g - 1 ? sy : 0)]);
    }
    if (NDIM > 2) {
      Emax = fmax(Emax, Et[id - (z - 1 >= g ? sz : 0)]);
      Emax = fmax(Emax, Et[id + (z + 1 <= nz - g - 1 ? sz : 0)]);
    }
    Real const Us = ((U_total / E > kDeEta1) || (U_total / Emax > kDeEta2)) ? U_total : U_advected;
    U[GE + id] = fmax(Us, kTiny);
  }
}

// Sync_Energies_{1,2,3}D (hydro_cuda.cu:1012-1099) with what follows it in Grid3D::Update_Hydro_Grid folded in, like the
// last corrector pass of a build without DE: the temperature floor of a -DDE build (hydro_cuda.cu:1117-1153; 3-D only) and
// the cell's inverse crossing time for Calc_dt_GPU.
template <int NDIM>
__global__ void __launch_bounds__(256) de_sync_kernel(Real *__restrict__ U, int nx, int ny, int nz, int g, Real gamma,
                                                      int apply_tfloor, Real U_floor, int do_dti, Real dx, Real dy,
                                                      Real dz, unsigned long long *key)
{
  __shared__ double sred[8];
  size_t const N = (size_t)nx * ny * nz;
  int const rx = nx - 2 * g, ry = NDIM > 1 ? ny - 2 * g : 1, rz = NDIM > 2 ? nz - 2 * g : 1;
  size_t const n = (size_t)rx * ry * rz;
  size_t const sy = nx, sz = (size_t)nx * ny, GE = (size_t)(5 + kNP - 1) * N;
  Real const idx = RC(1.0) / dx, idy = RC(1.0) / dy, idz = RC(1.0) / dz;
  KCS const kc = make_kct<KCS>(gamma);
  double m = -DBL_MAX;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    int const x = (int)(t % rx) + g;
    size_t const r = t / rx;
    int const y = NDIM > 1 ? (int)(r % ry) + g : 0, z = NDIM > 2 ? (int)(r / ry) + g : 0;
    size_t const id = (size_t)x + sy * y + sz * z;
    Real const d = U[id], d_inv = RC(1.0) / d;
    Real const mx = U[N + id], my = U[2 * N + id], mz = U[3 * N + id];
    Real const vx = mx * d_inv, vy = my * d_inv, vz = mz * d_inv;
    Real ge = U[GE + id];
    Real const Ekin = RC(0.5) * d * (vx * vx + vy * vy + vz * vz);
    Real E = Ekin + ge;
    if (apply_tfloor && NDIM == 3) {
      if ((E - Ekin) / d < U_floor) E = Ekin + d * U_floor;
      if (ge / d < U_floor) {
        ge         = d * U_floor;
        U[GE + id] = ge;
      }
    }
    U[4 * N + id] = E;
    if (do_dti) m = fmax(m, (double)inverse_crossing_time<NDIM>(d, mx, my, mz, E, idx, idy, idz, dx, dy, dz, kc));
  }
  if (!do_dti) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < 8 ? sred[threadIdx.x] : -DBL_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) atomicMax(key, dti_encode(m));
  }
}

cudaError_t de_finish(void *U, int nx, int ny, int nz, int g, double gamma, int apply_tfloor, double U_floor, int do_dti,
                      double dx, double dy, double dz, unsigned long long *key, cudaStream_t s)
{
  int const ndim = 1 + (ny > 1) + (nz > 1);
  size_t const n = (size_t)(nx - 2 * g) * (ndim > 1 ? ny - 2 * g : 1) * (ndim > 2 ? nz - 2 * g : 1);
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  Real *u = (Real *)U;
#define CHL_DE_FIN(ND)                                                                                                \
  de_select_kernel<ND><<<(unsigned)blocks, 256, 0, s>>>(u, nx, ny, nz, g);                                             \
  de_sync_kernel<ND><<<(unsigned)blocks, 256, 0, s>>>(u, nx, ny, nz, g, (Real)gamma, apply_tfloor, (Real)U_floor, do_dti, \
                                                      (Real)dx, (Real)dy, (Real)dz, key)
  if (ndim == 3) { CHL_DE_FIN(3); }
  else if (ndim == 2) { CHL_DE_FIN(2); }
  else { CHL_DE_FIN(1); }
#undef CHL_DE_FIN
  return cudaGetLastError();
}
#else
cudaError_t (*const de_term)(const void *, void *, int, int, int, const int[3], const int[3], double, double, double, double,
                             int, cudaStream_t) = nullptr;
cudaError_t (*const de_finish)(void *, int, int, int, int, double, int, double, int, double, double, double,
                               unsigned long long *, cudaStream_t) = nullptr;
#endif

ModeOps const ops = {sweep_split, calc_dti, tfloor, prim_from_cons, sizeof(Real), kHalfFields, kNP, CHL_DE, de_term, de_finish};
}  // namespace
ModeOps const *get_ops() { return &ops; }
}  // namespace CHL_NS

namespace chl
{
ModeOps const *CHL_CAT(mode_ops_, CHL_MODE)() { return CHL_NS::get_ops(); }
}  // namespace chl
