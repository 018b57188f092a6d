// mhd_inst.cu -- instantiates the MHD kernels (mhd_sweep.cuh) for ONE arithmetic mode and exposes them through
// chl::MhdOps.  Compiled once per mode like vl_sweep_inst.cu (strict: -fmad=false).
#include <float.h>

#include "chl_internal.h"
#include "mhd_sweep.cuh"

#define CHL_CAT2(a, b) a##b
#define CHL_CAT(a, b) CHL_CAT2(a, b)

namespace CHL_NS
{
using chl::MhdArgs;
using chl::MhdOps;
namespace
{
template <int RECON, int STAGE, int DIR, bool FIRST, bool LAST>
cudaError_t launch_mhd_dir(MhdDirParams const &p, cudaStream_t s)
{
  int const ox = p.fhi[0] - p.flo[0], oy = p.fhi[1] - p.flo[1], oz = p.fhi[2] - p.flo[2];
  if (ox <= 0 || oy <= 0 || oz <= 0) return cudaSuccess;
  dim3 grid;
  if (DIR == 0) grid = dim3((oy + kRowsPerCta - 1) / kRowsPerCta, oz, 1);
  MhdDirParams q = p;
  q.xw_log2 = 7;
  if (DIR != 0) {
    int const om = DIR == 1 ? oy : oz, oo = DIR == 1 ? oz : oy;
    q.xw_log2 = pick_xw_log2(ox, oo);
    int const xw = 1 << q.xw_log2, ow = kDirThreads >> q.xw_log2;
    grid = dim3((ox + xw - 1) / xw, (oo + ow - 1) / ow, (om + p.chunk - 1) / p.chunk);
  }
  mhd_dir_kernel<RECON, STAGE, DIR, FIRST, LAST><<<grid, kDirThreads, 0, s>>>(q);
  return cudaGetLastError();
}

template <int RECON, int STAGE>
cudaError_t run_mhd_dirs(MhdDirParams p, Real *ctf, size_t N, double const dto[3], int dir_chunk, cudaStream_t s)
{
  auto chunk_for = [&](int d) {
    int const o = p.fhi[d] - p.flo[d];
    if (dir_chunk > 0) return dir_chunk < o ? dir_chunk : o;
    if (o >= 384) return 64;
    if (o >= 64) return 32;
    int const c = (o + 1) / 2;
    return c < 8 ? (o < 8 ? o : 8) : c;
  };
  cudaError_t e;
  p.dto = (Real)dto[0];
  p.ctf = ctf;
  if ((e = launch_mhd_dir<RECON, STAGE, 0, true, false>(p, s)) != cudaSuccess) return e;
  p.dto   = (Real)dto[1];
  p.ctf   = ctf + 3 * N;
  p.chunk = chunk_for(1);
  if ((e = launch_mhd_dir<RECON, STAGE, 1, false, false>(p, s)) != cudaSuccess) return e;
  p.dto   = (Real)dto[2];
  p.ctf   = ctf + 6 * N;
  p.chunk = chunk_for(2);
  return launch_mhd_dir<RECON, STAGE, 2, false, true>(p, s);
}

cudaError_t ct_and_b(MhdArgs const &a, const Real *ecen, const Real *bsrc, Real *bdst, int const elo[3],
                     int const ehi[3], int const blo[3], int const bhi[3], int const clo[3], int const chi[3], double dt,
                     cudaStream_t s)
{
  size_t const N = (size_t)a.nx * a.ny * a.nz;
  MhdCtParams c;
  for (int d = 0; d < 3; d++) c.ctf[d] = (const Real *)a.ctf + (size_t)3 * d * N;
  c.ecen = ecen;
  c.E = (Real *)a.E;
  c.nx = a.nx; c.ny = a.ny; c.nz = a.nz;
  for (int d = 0; d < 3; d++) { c.lo[d] = elo[d]; c.hi[d] = ehi[d]; }
  dim3 g1((ehi[0] - elo[0] + 127) / 128, ehi[1] - elo[1], ehi[2] - elo[2]);
  mhd_ct_kernel<<<g1, 128, 0, s>>>(c);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  MhdBParams b;
  b.src = bsrc; b.dst = bdst; b.E = (const Real *)a.E;
  b.nx = a.nx; b.ny = a.ny; b.nz = a.nz;
  for (int d = 0; d < 3; d++) { b.lo[d] = blo[d]; b.hi[d] = bhi[d]; b.clo[d] = clo[d]; b.chi[d] = chi[d]; }
  b.dtodx = (Real)(dt / a.dx); b.dtody = (Real)(dt / a.dy); b.dtodz = (Real)(dt / a.dz);
  dim3 g2((bhi[0] - blo[0] + 127) / 128, bhi[1] - blo[1], bhi[2] - blo[2]);
  mhd_bupdate_kernel<<<g2, 128, 0, s>>>(b);
  return cudaGetLastError();
}

// VL_Algorithm_3D_CUDA, MHD branch
cudaError_t mhd_step(MhdArgs const &a, cudaStream_t s)
{
  size_t const N = (size_t)a.nx * a.ny * a.nz;
  int const n[3] = {a.nx, a.ny, a.nz};
  int const g    = a.n_ghost;
  Real *U = (Real *)a.U, *H = (Real *)a.half, *W = (Real *)a.prim;
  Real *ecen = (Real *)a.E + 3 * N;  // cell-centred EMFs live behind the three edge EMFs
  MhdDirParams p;
  p.acc = (Real *)a.acc;
  p.nx = a.nx; p.ny = a.ny; p.nz = a.nz;
  p.gamma = (Real)a.gamma; p.density_floor = (Real)a.density_floor;
  p.chunk = 1;
  p.ecen  = ecen;
  cudaError_t e;
  // ---- predictor ----
  {
    double const h = 0.5 * a.dt;
    double const dto[3] = {h / a.dx, h / a.dy, h / a.dz};
    p.in = U; p.Bsrc = U; p.out = H;
    for (int d = 0; d < 3; d++) { p.flo[d] = 0; p.fhi[d] = n[d] - 1; p.ulo[d] = 1; p.uhi[d] = n[d] - 1; }
    // transverse cells: every cell, [0, n); along the sweep the +faces of [0, n-1) -- the kernel reads flo/fhi of the
    // sweep axis as the face range and of the other axes as the cell range, so widen the transverse ones per launch
    MhdDirParams q = p;
    auto launch_stage0 = [&](void) -> cudaError_t {
      // x, y, z with the transverse extents opened to the full grid
      cudaError_t r;
      double const *t = dto;
      MhdDirParams x = q; x.fhi[1] = n[1]; x.fhi[2] = n[2]; x.dto = (Real)t[0]; x.ctf = (Real *)a.ctf;
      if ((r = launch_mhd_dir<kPCM, 0, 0, true, false>(x, s)) != cudaSuccess) return r;
      MhdDirParams y = q; y.fhi[0] = n[0]; y.fhi[2] = n[2]; y.dto = (Real)t[1]; y.ctf = (Real *)a.ctf + 3 * N;
      y.chunk = a.dir_chunk > 0 ? a.dir_chunk : (n[1] >= 384 ? 64 : 32);
      if ((r = launch_mhd_dir<kPCM, 0, 1, false, false>(y, s)) != cudaSuccess) return r;
      MhdDirParams z = q; z.fhi[0] = n[0]; z.fhi[1] = n[1]; z.dto = (Real)t[2]; z.ctf = (Real *)a.ctf + 6 * N;
      z.chunk = a.dir_chunk > 0 ? a.dir_chunk : (n[2] >= 384 ? 64 : 32);
      return launch_mhd_dir<kPCM, 0, 2, false, true>(z, s);
    };
    if ((e = launch_stage0()) != cudaSuccess) return e;
    int const elo[3] = {1, 1, 1}, ehi[3] = {n[0], n[1], n[2]};
    int const blo[3] = {1, 1, 1}, bhi[3] = {n[0] - 1, n[1] - 1, n[2] - 1};
    if ((e = ct_and_b(a, ecen, U, H, elo, ehi, blo, bhi, blo, bhi, h, s)) != cudaSuccess) return e;
    dim3 gp((n[0] - 2 + 127) / 128, n[1] - 2, n[2] - 2);
    mhd_prim_kernel<<<gp, 128, 0, s>>>(H, W, ecen, a.nx, a.ny, a.nz, 1, (Real)a.gamma);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  // ---- corrector ----
  {
    double const dto[3] = {a.dt / a.dx, a.dt / a.dy, a.dt / a.dz};
    p.in = W; p.Bsrc = H; p.out = U;
    for (int d = 0; d < 3; d++) { p.flo[d] = g - 1; p.fhi[d] = n[d] - g + 1; p.ulo[d] = g; p.uhi[d] = n[d] - g; }
    switch (a.recon) {
      case kPCM: e = run_mhd_dirs<kPCM, 1>(p, (Real *)a.ctf, N, dto, a.dir_chunk, s); break;
      case kPLMC: e = run_mhd_dirs<kPLMC, 1>(p, (Real *)a.ctf, N, dto, a.dir_chunk, s); break;
      default: e = run_mhd_dirs<kPPMC, 1>(p, (Real *)a.ctf, N, dto, a.dir_chunk, s); break;
    }
    if (e != cudaSuccess) return e;
    int const elo[3] = {g, g, g}, ehi[3] = {n[0] - g + 1, n[1] - g + 1, n[2] - g + 1};
    int const blo[3] = {g - 1, g - 1, g - 1}, bhi[3] = {n[0] - g, n[1] - g, n[2] - g};
    int const clo[3] = {g, g, g};
    if ((e = ct_and_b(a, ecen, U, U, elo, ehi, blo, bhi, clo, bhi, a.dt, s)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t mhd_calc_dti(const void *U, int nx, int ny, int nz, int g, double dx, double dy, double dz, double gamma,
                         unsigned long long *key, cudaStream_t s)
{
  size_t const nreal = (size_t)(nx - 2 * g) * (ny - 2 * g) * (nz - 2 * g);
  int blocks = (int)((nreal + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  mhd_calc_dti_kernel<<<blocks, 256, 0, s>>>((const Real *)U, nx, ny, nz, g, (Real)dx, (Real)dy, (Real)dz, (Real)gamma,
                                             key);
  return cudaGetLastError();
}

MhdOps const ops = {mhd_step, mhd_calc_dti, sizeof(Real), kMhdPrimFields};
}  // namespace
MhdOps const *get_mhd_ops() { return &ops; }
}  // namespace CHL_NS

namespace chl
{
MhdOps const *CHL_CAT(mhd_ops_, CHL_MODE)() { return CHL_NS::get_mhd_ops(); }
}  // namespace chl
