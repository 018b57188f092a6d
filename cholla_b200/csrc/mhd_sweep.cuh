// mhd_sweep.cuh -- the MHD branch of VL_Algorithm_3D_CUDA (VL_3D_cuda.cu:38-360 with -DMHD) as direction-split
// marching sweeps plus three small constrained-transport kernels.  Per stage (predictor / corrector):
//
//   mhd_dir_kernel x3   reconstruct (PCM | PLMC | PPMC on the 7-wave eigen-system) + HLLD along one direction, marching
//                       like vl_dir_kernel (vl_dirsweep.cuh).  The five hydro components of the flux difference go to
//                       the running sum / the cell update; the three flux components constrained transport needs
//                       (mass flux for upwinding, the two induction fluxes) are stored: 9 words per cell for the three
//                       directions instead of the reference's three 7-field flux arrays and six 7-field interface arrays.
//   mhd_ct_kernel       Calculate_CT_Electric_Fields (ct_electric_fields.cu:20-282): edge-centred EMFs.
//   mhd_bupdate_kernel  Update_Magnetic_Field_3D (magnetic_update.cu:21-80): face-centred induction update.
//   mhd_prim_kernel     (after the predictor) half-step primitives + cell-centred fields, once per cell, for the
//                       corrector's stencils -- what Load_Data (reconstruction.h:139-212) recomputes 15x per cell.
//
// Regions (g = n_ghost, n = cells per axis incl. ghosts).  Predictor: fluxes through the +faces of cells [0, n-2] along
// the sweep for every transverse cell; hydro update and B update on [1, n-2] (VL_3D_cuda.cu:408, magnetic_update.cu:32).
// Corrector: fluxes through the +faces of cells [g-1, n-g] with transverse cells in [g-1, n-g]; EMFs on [g, n-g]; hydro
// update on the real cells; B_d on d-index [g-1, n-g) (the low face of the first real cell feeds the CFL reduction's
// cell-centred field) and real transverse cells.  Everything the reference computes outside these regions is
// overwritten by the boundary routine before it is read.
#pragma once
#include "mhd_physics.cuh"
#include "vl_dirsweep.cuh"

namespace CHL_NS
{
#ifndef CHL_MHD_MINB
#define CHL_MHD_MINB 4
#endif
#ifndef CHL_MHD_MINB0  // the predictor passes (PCM + HLLD) may want a different register budget
#define CHL_MHD_MINB0 CHL_MHD_MINB
#endif
#if CHL_STRICT
constexpr int kMhdPrimFields = 10;  // rho, vx, vy, vz, P_x, Bx, By, Bz, P_y, P_z
#else
constexpr int kMhdPrimFields = 8;
#endif

struct MhdDirParams {
  const Real *in;    // stage 0: U [8][N];  stage 1: W [kMhdPrimFields][N]
  const Real *Bsrc;  // the state whose face-centred field this stage uses: U (stage 0) or U-half (stage 1), [8][N]
  Real *acc;         // [5][N]
  Real *out;         // LAST: stage 0 -> U-half (hydro fields), stage 1 -> U
  Real *ctf;         // [3][N]: mass flux, induction flux slot 5, slot 6 through the +face of each cell, this direction
  Real *ecen;        // predictor x sweep only: [3][N] cell-centred reference EMFs of the state, for constrained transport
  int nx, ny, nz;
  int flo[3], fhi[3];  // cells whose +face flux is produced: [flo, fhi) along the sweep; transverse cells [flo, fhi)
  int ulo[3], uhi[3];  // cells whose hydro fields are updated
  int chunk;
  int xw_log2;  // y / z sweeps: log2 of the CTA's extent along x (pick_xw_log2, vl_dirsweep.cuh)
  Real dto, gamma, density_floor;
};

template <int DIR>
__device__ __forceinline__ MOff mhd_offsets(size_t N, int nx, int ny)
{
  MOff o;
  o.v1 = (size_t)(1 + DIR) * N;
  o.v2 = (size_t)(1 + (DIR + 1) % 3) * N;
  o.v3 = (size_t)(1 + (DIR + 2) % 3) * N;
  o.p  = (CHL_STRICT && DIR > 0) ? (size_t)(7 + DIR) * N : (size_t)4 * N;
  o.b1 = (size_t)(5 + DIR) * N;
  o.b2 = (size_t)(5 + (DIR + 1) % 3) * N;
  o.b3 = (size_t)(5 + (DIR + 2) % 3) * N;
  o.s  = DIR == 0 ? 1 : (DIR == 1 ? (long)nx : (long)nx * ny);
  return o;
}

// cellCenteredMagneticFields (mhd_utilities.h:205-236) of cell (x, y, z) of a conserved array
__device__ __forceinline__ void cell_centred_b(const Real *__restrict__ U, size_t N, size_t id, int x, int y, int z,
                                               int nx, int ny, Real &bx, Real &by, Real &bz)
{
  bx = (x > 0) ? RC(0.5) * (__ldg(U + 5 * N + id) + __ldg(U + 5 * N + id - 1)) : __ldg(U + 5 * N + id);
  by = (y > 0) ? RC(0.5) * (__ldg(U + 6 * N + id) + __ldg(U + 6 * N + id - nx)) : __ldg(U + 6 * N + id);
  bz = (z > 0) ? RC(0.5) * (__ldg(U + 7 * N + id) + __ldg(U + 7 * N + id - (size_t)nx * ny)) : __ldg(U + 7 * N + id);
}

// The cell-centred reference EMFs _ctSlope subtracts from the face EMFs (ct_electric_fields.h:98-113):
// E_c = (M3 * B2 - M2 * B3) / d with (2, 3) the two directions after c in cyclic order.  The reference recomputes each
// of them in every one of the 24 slopes an edge triple needs (~8 times per cell, a division and seven loads each);
// here they are computed once per cell by a kernel that visits the cell anyway -- the predictor's x sweep for the
// state U, the half-step primitive kernel for U-half -- with the same expression, hence the same bits.
__device__ __forceinline__ void cell_emf(Real *__restrict__ ecen, size_t N, size_t id, Real d, Real mx, Real my, Real mz,
                                         Real bx, Real by, Real bz)
{
  KC kb;  // once per cell and stage: the per-quotient fallback branch (kDivBranchy) is fine here
  kb.bad           = false;
  RDen const dd    = rden(d, kb);
  ecen[id]         = rdiv(mz * by - my * bz, dd, kb);
  ecen[N + id]     = rdiv(mx * bz - mz * bx, dd, kb);
  ecen[2 * N + id] = rdiv(my * bx - mx * by, dd, kb);
}

// ---- the two blocks of a sweep iteration: branch-free arithmetic, one flag test, out-of-line IEEE redo (as in
// sweep_common.cuh) ------------------------------------------------------------------------------------------------
struct M7x2 {
  M7 L, R;
};
template <int RECON, class K>
__device__ __forceinline__ void mhd_recon(const Real *__restrict__ c, MOff const &o, Real bhi, Real blo, K const &k, M7 &L,
                                          M7 &R)
{
  if (RECON == kPCM) {
    MPrim const w = load_mprim(c, o, 0);  // PCM corrector: interface = cell, energy from the cell state
    L             = mhd_interface(w, w.bx, k);
    // PCM copies the conserved energy; rebuilding it from primitives with the centred field is the same sum
    R = L;
    return;
  }
  MPrim l, r;
  if (RECON == kPPMC) mhd_ppmc(c, o, k, l, r);
  else mhd_plmc(c, o, k, l, r);
  L = mhd_interface(l, bhi, k);
  R = mhd_interface(r, blo, k);
}
#if CHL_STRICT
template <int RECON>
static __device__ __noinline__ M7x2 mhd_recon_ieee(const Real *c, MOff o, Real bhi, Real blo, Real gamma)
{
  KCS const k = make_kct<KCS>(gamma);
  M7x2 out;
  mhd_recon<RECON>(c, o, bhi, blo, k, out.L, out.R);
  return out;
}
static __device__ __noinline__ M7 hlld_ieee(M7 L, M7 R, Real bface, Real gamma)
{
  KCS const k = make_kct<KCS>(gamma);
  M7 f;
  hlld_flux(L, R, bface, k, f);
  return f;
}
#endif
template <int RECON>
__device__ __forceinline__ void mhd_recon_checked(const Real *__restrict__ c, MOff const &o, Real bhi, Real blo, KH const &k,
                                                  M7 &L, M7 &R)
{
  mhd_recon<RECON>(c, o, bhi, blo, k, L, R);
#if CHL_STRICT
  if (k.bad) {
    M7x2 const out = mhd_recon_ieee<RECON>(c, o, bhi, blo, k.gamma);
    L     = out.L;
    R     = out.R;
    k.bad = false;
  }
#endif
}
__device__ __forceinline__ void hlld_checked(M7 const &L, M7 const &R, Real bface, KH const &k, M7 &F)
{
  hlld_flux(L, R, bface, k, F);
#if CHL_STRICT
  if (k.bad) {
    F     = hlld_ieee(L, R, bface, k.gamma);
    k.bad = false;
  }
#endif
}

template <int RECON, int STAGE, int DIR, bool FIRST, bool LAST>
__global__ void __launch_bounds__(kDirThreads, STAGE == 0 ? CHL_MHD_MINB0 : CHL_MHD_MINB) mhd_dir_kernel(MhdDirParams const P)
{
  int const nx = P.nx, ny = P.ny;
  size_t const N = (size_t)nx * ny * P.nz;
  int const lane = threadIdx.x & 31;
  Real const gamma = P.gamma;
  MOff const o     = mhd_offsets<DIR>(N, nx, ny);
  constexpr int T1 = (DIR + 1) % 3, T2 = (DIR + 2) % 3;

  // coordinates: m = marching coordinate, (t1, t2) the transverse ones in cyclic order
  bool live;
  int xyz[3];
  int c_first, c_last;
  if (DIR == 0) {
    xyz[1]  = P.flo[1] + blockIdx.x * kRowsPerCta + (threadIdx.x >> 5);
    xyz[2]  = P.flo[2] + blockIdx.y;
    live    = xyz[1] < P.fhi[1];
    c_first = P.flo[0];
    c_last  = P.fhi[0];
  } else {
    int const xw = 1 << P.xw_log2;
    xyz[0]      = P.flo[0] + blockIdx.x * xw + (threadIdx.x & (xw - 1));
    int const q = DIR == 1 ? 2 : 1;
    xyz[q]      = P.flo[q] + blockIdx.y * (kDirThreads >> P.xw_log2) + (threadIdx.x >> P.xw_log2);
    live        = xyz[0] < P.fhi[0] && xyz[q] < P.fhi[q];
    int const b = P.flo[DIR] + blockIdx.z * P.chunk;  // this chunk completes cells [b, b + chunk)
    c_first     = max(b - 1, P.flo[DIR]);
    c_last      = min(b + P.chunk, P.fhi[DIR]);
  }
  if (!live) return;
  // hydro update allowed for this thread's transverse position?
  bool upd_t;
  if (DIR == 0) upd_t = xyz[1] >= P.ulo[1] && xyz[1] < P.uhi[1] && xyz[2] >= P.ulo[2] && xyz[2] < P.uhi[2];
  else if (DIR == 1) upd_t = xyz[2] >= P.ulo[2] && xyz[2] < P.uhi[2];
  else upd_t = xyz[1] >= P.ulo[1] && xyz[1] < P.uhi[1];

  // carried values parked in shared memory and update operands staged by cp.async, as in vl_dir_kernel
  typedef StageGeom<STAGE, FIRST, LAST> SG;
  constexpr bool kStage = CHL_DIR_STAGE && SG::n > 0;
  constexpr int kParkN  = CHL_DIR_PARK ? (DIR == 0 ? 12 * kRowsPerCta : 12 * kDirThreads) : 1;
  __shared__ Real stage[kStage ? SG::n * kDirThreads : 1];
  __shared__ Real parked[kParkN];
  Real *const park   = parked + (DIR == 0 ? (threadIdx.x >> 5) : threadIdx.x);
  int const pstride  = DIR == 0 ? kRowsPerCta : kDirThreads;
  Real *const staged = stage + threadIdx.x;
  constexpr int field_of_comp[5] = {0, 1 + DIR, 1 + (DIR + 1) % 3, 1 + (DIR + 2) % 3, 4};
  KH const kc = make_kct<KH>(gamma);

  Real Lc[7], Fc[5];
#pragma unroll
  for (int k = 0; k < 7; k++) Lc[k] = RC(0.0);
#pragma unroll
  for (int k = 0; k < 5; k++) Fc[k] = RC(0.0);
  int const step = DIR == 0 ? 32 : 1;
  for (int cb = c_first; cb <= c_last; cb += step) {
    int const c      = DIR == 0 ? cb + lane : cb;
    bool const valid = c <= c_last;
    xyz[DIR]         = c;
    size_t const id  = (size_t)xyz[0] + (size_t)nx * ((size_t)xyz[1] + (size_t)ny * xyz[2]);
    bool upd_x = true;  // y/z sweeps: the thread's own x must be an updated column
    if (DIR != 0) upd_x = xyz[0] >= P.ulo[0] && xyz[0] < P.uhi[0];
    bool const face = valid && c > c_first;  // the +face of cell c-1
    int const u     = c - 1;                 // the cell both of whose faces are known after this iteration
    bool const fin  = face && c > c_first + 1 && upd_t && upd_x && u >= P.ulo[DIR] && u < P.uhi[DIR];
    if (kStage) {
      if (fin) {
        size_t const idu = id - o.s;
#pragma unroll
        for (int k = 0; k < SG::n_sum; k++) {
          const Real *src = (STAGE == 0 && FIRST) ? P.in : P.acc;
          cp_async_real(staged + k * kDirThreads, src + (size_t)field_of_comp[k] * N + idu);
        }
#pragma unroll
        for (int f = 0; f < SG::n_out; f++) cp_async_real(staged + (SG::n_sum + f) * kDirThreads, P.out + (size_t)f * N + idu);
      }
      cp_async_commit();
    }
    M7 L, R;
    if (valid) {
      if (STAGE == 0) {  // PCM_Reconstruction_3D (pcm_cuda.cu:256-424): both sides are the cell itself
        Real b[3];
        cell_centred_b(P.in, N, id, xyz[0], xyz[1], xyz[2], nx, ny, b[0], b[1], b[2]);
        L.v[0] = __ldg(P.in + id);
        L.v[1] = __ldg(P.in + o.v1 + id);
        L.v[2] = __ldg(P.in + o.v2 + id);
        L.v[3] = __ldg(P.in + o.v3 + id);
        L.v[4] = __ldg(P.in + 4 * N + id);
        L.v[5] = b[T1];
        L.v[6] = b[T2];
        R      = L;
        if (DIR == 0) cell_emf(P.ecen, N, id, L.v[0], L.v[1], L.v[2], L.v[3], b[0], b[1], b[2]);
      } else {
        Real const bhi = __ldg(P.Bsrc + (size_t)(5 + DIR) * N + id);            // +face of cell c
        Real const blo = c > 0 ? __ldg(P.Bsrc + (size_t)(5 + DIR) * N + id - o.s) : bhi;  // -face of cell c
        mhd_recon_checked<RECON>(P.in + id, o, bhi, blo, kc, L, R);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 7; k++) L.v[k] = R.v[k] = RC(1.0);
    }
    M7 Lp;
    from_low_side_n<DIR, 7>(L.v, Lc, Lp.v, park, pstride, lane);
    M7 F;
    if (face) {
      Real const bface = __ldg(P.Bsrc + (size_t)(5 + DIR) * N + id - o.s);
      hlld_checked(Lp, R, bface, kc, F);
      size_t const idm = id - o.s;
      P.ctf[idm]         = F.v[0];
      P.ctf[N + idm]     = F.v[5];
      P.ctf[2 * N + idm] = F.v[6];
    } else {
#pragma unroll
      for (int k = 0; k < 7; k++) F.v[k] = RC(0.0);
    }
    Real F5[5], Fp[5];
#pragma unroll
    for (int k = 0; k < 5; k++) F5[k] = F.v[k];
    from_low_side_n<DIR, 5>(F5, Fc, Fp, park + 7 * pstride, pstride, lane);
    if (kStage) cp_async_wait_all();
    if (fin) {
      size_t const idu = id - o.s;
      Real v[5];
#pragma unroll
      for (int k = 0; k < 5; k++) {
        int const f  = field_of_comp[k];
        Real const t = P.dto * (Fp[k] - F.v[k]);
        Real a;
        if (kStage && SG::n_sum) a = staged[k * kDirThreads] + t;
        else if (STAGE == 0) a = (FIRST ? __ldg(P.in + f * N + idu) : P.acc[f * N + idu]) + t;
        else a = FIRST ? t : P.acc[f * N + idu] + t;
        if (!LAST) P.acc[f * N + idu] = a;
        v[f] = a;
      }
      if (LAST) {
        if (STAGE == 1) {
#pragma unroll
          for (int f = 0; f < 5; f++) v[f] = (kStage ? staged[(SG::n_sum + f) * kDirThreads] : P.out[f * N + idu]) + v[f];
        }
        bool const fl = STAGE == 0 ? (v[0] < P.density_floor) : (v[0] < P.density_floor && v[0] > RC(0.));
        if (fl) {  // VL_3D_cuda.cu:479-493 / hydro_cuda.cu:279-291
          Real const d0 = v[0];
          v[0]          = P.density_floor;
#pragma unroll
          for (int f = 1; f < 5; f++) v[f] *= (P.density_floor / d0);
        }
#pragma unroll
        for (int f = 0; f < 5; f++) P.out[f * N + idu] = v[f];
      }
    }
  }
}

// ---- constrained transport ----------------------------------------------------------------------------------------
struct MhdCtParams {
  const Real *ctf[3];  // per direction [3][N]
  const Real *ecen;    // [3][N] cell-centred reference EMFs of the state (cell_emf)
  Real *E;             // [3][N]
  int nx, ny, nz;
  int lo[3], hi[3];  // EMFs are produced on [lo, hi)
};

// _ctSlope (ct_electric_fields.h:66-116): face EMF minus the cell-centred reference EMF
__device__ __forceinline__ Real ct_slope(const Real *__restrict__ ctf, const Real *__restrict__ ecen, Real sign, int ctdir,
                                         int cq1, int cq2, int fq1, int fq2, int x, int y, int z, int nx, int ny,
                                         size_t N)
{
  int const xc = x - (cq1 == 0) - (cq2 == 0), yc = y - (cq1 == 1) - (cq2 == 1), zc = z - (cq1 == 2) - (cq2 == 2);
  size_t const ic  = (size_t)xc + (size_t)nx * ((size_t)yc + (size_t)ny * zc);
  size_t const ifl = (size_t)(x - (fq1 == 0) - (fq2 == 0)) +
                     (size_t)nx * ((size_t)(y - (fq1 == 1) - (fq2 == 1)) + (size_t)ny * (z - (fq1 == 2) - (fq2 == 2)));
  Real const ec = __ldg(ecen + (size_t)ctdir * N + ic);
  Real const ef = sign * __ldg(ctf + ifl + (size_t)(sign == RC(1.0) ? 2 : 1) * N);
  return ef - ec;
}

__device__ __forceinline__ Real ct_upwind(Real su, Real a, Real b)
{
  if (su > RC(0.0)) return a;
  if (su < RC(0.0)) return b;
  return RC(0.5) * (a + b);
}

__global__ void __launch_bounds__(128) mhd_ct_kernel(MhdCtParams const P)
{
  int const nx = P.nx, ny = P.ny;
  size_t const N = (size_t)nx * ny * P.nz;
  int const x = P.lo[0] + blockIdx.x * 128 + threadIdx.x, y = P.lo[1] + blockIdx.y, z = P.lo[2] + blockIdx.z;
  if (x >= P.hi[0]) return;
  const Real *Fx = P.ctf[0], *Fy = P.ctf[1], *Fz = P.ctf[2];
  const Real *U  = P.ecen;
#define CIDX(a, b, c) ((size_t)(a) + (size_t)nx * ((size_t)(b) + (size_t)ny * (size_t)(c)))
#define SLP(fl, sg, cd, c1, c2, f1, f2) ct_slope(fl, U, RC(sg), cd, c1, c2, f1, f2, x, y, z, nx, ny, N)
  size_t const id = CIDX(x, y, z);
  Real su, syp, syn, szp, szn, sxp, sxn;
  // X electric field (ct_electric_fields.cu:65-129)
  su  = __ldg(Fz + CIDX(x, y, z - 1));
  syp = ct_upwind(su, SLP(Fy, -1, 0, 2, -1, 1, 2), SLP(Fy, -1, 0, -1, -1, 1, -1));
  su  = __ldg(Fz + CIDX(x, y - 1, z - 1));
  syn = ct_upwind(su, SLP(Fy, -1, 0, 1, 2, 1, 2), SLP(Fy, -1, 0, 1, -1, 1, -1));
  su  = __ldg(Fy + CIDX(x, y - 1, z));
  szp = ct_upwind(su, SLP(Fz, 1, 0, 1, -1, 1, 2), SLP(Fz, 1, 0, -1, -1, 2, -1));
  su  = __ldg(Fy + CIDX(x, y - 1, z - 1));
  szn = ct_upwind(su, SLP(Fz, 1, 0, 1, 2, 1, 2), SLP(Fz, 1, 0, 2, -1, -1, 2));
  {
    Real const fyp = +__ldg(Fz + CIDX(x, y, z - 1) + 2 * N), fyn = +__ldg(Fz + CIDX(x, y - 1, z - 1) + 2 * N);
    Real const fzp = -__ldg(Fy + CIDX(x, y - 1, z) + N), fzn = -__ldg(Fy + CIDX(x, y - 1, z - 1) + N);
    P.E[id] = RC(0.25) * (+fyp + fyn + fzp + fzn + syp + syn + szp + szn);
  }
  // Y electric field (:131-194)
  su  = __ldg(Fz + CIDX(x, y, z - 1));
  sxp = ct_upwind(su, SLP(Fx, 1, 1, 2, -1, 0, 2), SLP(Fx, 1, 1, -1, -1, 0, -1));
  su  = __ldg(Fz + CIDX(x - 1, y, z - 1));
  sxn = ct_upwind(su, SLP(Fx, 1, 1, 0, 2, 0, 2), SLP(Fx, 1, 1, 0, -1, 0, -1));
  su  = __ldg(Fx + CIDX(x - 1, y, z));
  szp = ct_upwind(su, SLP(Fz, -1, 1, 0, -1, 0, 2), SLP(Fz, -1, 1, -1, -1, 2, -1));
  su  = __ldg(Fx + CIDX(x - 1, y, z - 1));
  szn = ct_upwind(su, SLP(Fz, -1, 1, 0, 2, 0, 2), SLP(Fz, -1, 1, 2, -1, 2, -1));
  {
    Real const fxp = -__ldg(Fz + CIDX(x, y, z - 1) + N), fxn = -__ldg(Fz + CIDX(x - 1, y, z - 1) + N);
    Real const fzp = +__ldg(Fx + CIDX(x - 1, y, z) + 2 * N), fzn = +__ldg(Fx + CIDX(x - 1, y, z - 1) + 2 * N);
    P.E[N + id] = RC(0.25) * (+fxp + fxn + fzp + fzn + sxp + sxn + szp + szn);
  }
  // Z electric field (:196-260)
  su  = __ldg(Fx + CIDX(x - 1, y, z));
  syp = ct_upwind(su, SLP(Fy, 1, 2, 0, -1, 0, 1), SLP(Fy, 1, 2, -1, -1, 1, -1));
  su  = __ldg(Fx + CIDX(x - 1, y - 1, z));
  syn = ct_upwind(su, SLP(Fy, 1, 2, 0, 1, 0, 1), SLP(Fy, 1, 2, 1, -1, 1, -1));
  su  = __ldg(Fy + CIDX(x, y - 1, z));
  sxp = ct_upwind(su, SLP(Fx, -1, 2, 1, -1, 0, 1), SLP(Fx, -1, 2, -1, -1, 0, -1));
  su  = __ldg(Fy + CIDX(x - 1, y - 1, z));
  sxn = ct_upwind(su, SLP(Fx, -1, 2, 0, 1, 0, 1), SLP(Fx, -1, 2, 0, -1, 0, -1));
  {
    Real const fxp = +__ldg(Fy + CIDX(x, y - 1, z) + 2 * N), fxn = +__ldg(Fy + CIDX(x - 1, y - 1, z) + 2 * N);
    Real const fyp = -__ldg(Fx + CIDX(x - 1, y, z) + N), fyn = -__ldg(Fx + CIDX(x - 1, y - 1, z) + N);
    P.E[2 * N + id] = RC(0.25) * (+fxp + fxn + fyp + fyn + sxp + sxn + syp + syn);
  }
#undef SLP
}

struct MhdBParams {
  const Real *src;  // [8][N]
  Real *dst;        // [8][N] (may alias src)
  const Real *E;    // [3][N]
  int nx, ny, nz;
  int lo[3], hi[3];    // box of cells visited
  int clo[3], chi[3];  // component d is updated only where the two coordinates other than d lie in [clo, chi)
  Real dtodx, dtody, dtodz;
};

__global__ void __launch_bounds__(128) mhd_bupdate_kernel(MhdBParams const P)
{
  int const nx = P.nx, ny = P.ny;
  size_t const N = (size_t)nx * ny * P.nz;
  int const x = P.lo[0] + blockIdx.x * 128 + threadIdx.x, y = P.lo[1] + blockIdx.y, z = P.lo[2] + blockIdx.z;
  if (x >= P.hi[0]) return;
  size_t const id = CIDX(x, y, z);
  bool const ix = x >= P.clo[0] && x < P.chi[0], iy = y >= P.clo[1] && y < P.chi[1], iz = z >= P.clo[2] && z < P.chi[2];
  const Real *E = P.E;
  if (iy && iz) {  // S&G 2009 eq. 10
    Real const ey1 = __ldg(E + CIDX(x + 1, y, z) + N), ey3 = __ldg(E + CIDX(x + 1, y, z + 1) + N);
    Real const ez1 = __ldg(E + CIDX(x + 1, y, z) + 2 * N), ez3 = __ldg(E + CIDX(x + 1, y + 1, z) + 2 * N);
    P.dst[id + 5 * N] = P.src[id + 5 * N] + P.dtodz * (ey3 - ey1) + P.dtody * (ez1 - ez3);
  }
  if (ix && iz) {  // eq. 11
    Real const ez2 = __ldg(E + CIDX(x, y + 1, z) + 2 * N), ez3 = __ldg(E + CIDX(x + 1, y + 1, z) + 2 * N);
    Real const ex1 = __ldg(E + CIDX(x, y + 1, z)), ex3 = __ldg(E + CIDX(x, y + 1, z + 1));
    P.dst[id + 6 * N] = P.src[id + 6 * N] + P.dtodx * (ez3 - ez2) + P.dtodz * (ex1 - ex3);
  }
  if (ix && iy) {  // eq. 12
    Real const ex2 = __ldg(E + CIDX(x, y, z + 1)), ex3 = __ldg(E + CIDX(x, y + 1, z + 1));
    Real const ey2 = __ldg(E + CIDX(x, y, z + 1) + N), ey3 = __ldg(E + CIDX(x + 1, y, z + 1) + N);
    P.dst[id + 7 * N] = P.src[id + 7 * N] + P.dtody * (ex3 - ex2) + P.dtodx * (ey2 - ey3);
  }
}

// half-step primitives: Load_Data (reconstruction.h:139-212, MHD) once per cell
__global__ void __launch_bounds__(128) mhd_prim_kernel(const Real *__restrict__ H, Real *__restrict__ W,
                                                       Real *__restrict__ ecen, int nx, int ny, int nz, int lo,
                                                       Real gamma)
{
  size_t const N = (size_t)nx * ny * nz;
  int const x = lo + blockIdx.x * 128 + threadIdx.x, y = lo + blockIdx.y, z = lo + blockIdx.z;
  if (x >= nx - lo) return;
  size_t const id = CIDX(x, y, z);
  Real const d  = H[id];
  KC kb;  // one pass per cell and step: the per-quotient fallback branch is fine here
  kb.bad        = false;
  RDen const dd = rden(d, kb);
  Real const vx = rdiv(H[N + id], dd, kb), vy = rdiv(H[2 * N + id], dd, kb), vz = rdiv(H[3 * N + id], dd, kb), E = H[4 * N + id];
  Real bx, by, bz;
  cell_centred_b(H, N, id, x, y, z, nx, ny, bx, by, bz);
  cell_emf(ecen, N, id, d, H[N + id], H[2 * N + id], H[3 * N + id], bx, by, bz);
  W[id]         = d;
  W[N + id]     = vx;
  W[2 * N + id] = vy;
  W[3 * N + id] = vz;
  W[4 * N + id] = mhd_pressure(E, d, vx, vy, vz, gamma, bx, by, bz);
  W[5 * N + id] = bx;
  W[6 * N + id] = by;
  W[7 * N + id] = bz;
#if CHL_STRICT
  W[8 * N + id] = mhd_pressure(E, d, vy, vz, vx, gamma, by, bz, bx);  // the y sweep's summation order
  W[9 * N + id] = mhd_pressure(E, d, vz, vx, vy, gamma, bz, bx, by);  // the z sweep's
#endif
}

// Calc_dt_3D, MHD branch (hydro_cuda.cu:514-560)
__global__ void __launch_bounds__(256) mhd_calc_dti_kernel(const Real *__restrict__ U, int nx, int ny, int nz, int g,
                                                           Real dx, Real dy, Real dz, Real gamma,
                                                           unsigned long long *key)
{
  __shared__ double sred[8];
  size_t const N = (size_t)nx * ny * nz;
  int const rx = nx - 2 * g, ry = ny - 2 * g, rz = nz - 2 * g;
  size_t const nreal = (size_t)rx * ry * rz;
  double m = -1.7976931348623157e308;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < nreal; t += (size_t)gridDim.x * blockDim.x) {
    int const x = (int)(t % rx) + g;
    size_t const r = t / rx;
    int const y = (int)(r % ry) + g, z = (int)(r / ry) + g;
    size_t const id = CIDX(x, y, z);
    Real const d = __ldg(U + id), d_inv = RC(1.0) / d;
    Real const vx = __ldg(U + N + id) * d_inv, vy = __ldg(U + 2 * N + id) * d_inv, vz = __ldg(U + 3 * N + id) * d_inv;
    Real bx, by, bz;
    cell_centred_b(U, N, id, x, y, z, nx, ny, bx, by, bz);
    m = fmax(m, (double)mhd_inverse_crossing_time(__ldg(U + 4 * N + id), d, vx, vy, vz, bx, by, bz, dx, dy, dz, gamma));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < 8 ? sred[threadIdx.x] : -1.7976931348623157e308;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) atomicMax(key, dti_encode(m));
  }
}
#undef CIDX

}  // namespace CHL_NS
