// vl_sweep.cuh -- the fused Van Leer sweep kernel (one launch per predictor / corrector stage).
//
// Replaces, per stage, the reference's chain of one-thread-per-cell kernels that round-trip every
// intermediate through HBM (VL_3D_cuda.cu:141-339: PCM -> 3x Riemann -> half update, then
// 3x reconstruct -> 3x Riemann -> full update, ~1.84 kB/cell-update of DRAM traffic):
//
//   STAGE 0 (predictor)  reads U (conserved, 1-cell stencil), first-order fluxes, writes the half-step
//                        state as PRIMITIVES W = (rho, vx, vy, vz, P) -- exactly what the reference's
//                        Load_Data (reconstruction.h:139-212) would derive from U-half.
//   STAGE 1 (corrector)  reads W (2- or 3-cell stencil), PLMC/PPMC + HLLC/Roe in the three directions,
//                        reads U once, writes U in place; optionally applies the temperature floor and
//                        folds the CFL max (Calc_dt_3D) into the same pass.
// HBM traffic per cell-update: 40 B read + 40 B write (stage 0), 80 B read + 40 B write (stage 1) = 200 B
// in fp64 -- SURVEY 8(d)'s algorithmic figure.  No interface or flux array ever exists in global memory.
//
// Work decomposition ("2.5-D"): a CTA owns a TX x TY column of cells and marches along z.  z-direction
// interface states and fluxes are carried in registers from plane to plane (zero redundancy, no barrier);
// x/y interface states and fluxes are exchanged through shared memory (R state published by the cell on
// the high side of a face, flux published by the cell on the low side).  The cells just outside the
// column whose interface states are needed (one per row/column end) are reconstructed by "halo tasks"
// packed into the first three warps.  All stencil loads go through L1/L2 (read-only path); the kernel is
// bound by the fp64 pipe, not by LSU or HBM, so no TMA staging is used -- see DESIGN.md for the numbers.
#pragma once
#include "hydro_physics.cuh"

namespace CHL_NS
{
constexpr int TX = 32;
constexpr int TY = 8;
constexpr int kThreads = TX * TY;
constexpr int kHaloYBase = 32;  // first thread of the y-halo tasks (x-halo tasks are threads [0, 2*TY))
static_assert(2 * TY <= 32 && kHaloYBase + 2 * TX <= kThreads, "halo tasks must fit in the CTA");

struct SweepParams {
  const Real *in;  // stage 0: U conserved [5][N];  stage 1: W primitives [5][N]
  Real *out;       // stage 0: W;                   stage 1: U (updated in place)
  int nx, ny, nz;
  int lo[3], hi[3];  // output region [lo, hi) per dimension
  int has_y, has_z;
  int zchunk;  // planes per CTA along z
  Real dtodx, dtody, dtodz;
  Real gamma;
  Real density_floor;
  // corrector epilogue
  int apply_tfloor;
  Real U_floor;
  int do_dti;
  Real dx, dy, dz, idx, idy, idz;
  unsigned long long *dti_key;  // ordered-bits max accumulator
};

struct S5 {
  Real v[5];
};  // strict: sweep-frame conserved; fast: sweep-frame primitives

// ---- ordered-bits encoding so that unsigned atomicMax orders doubles/floats ----------------------------------
__host__ __device__ inline unsigned long long dti_encode(double x)
{
  unsigned long long u;
#ifdef __CUDA_ARCH__
  u = (unsigned long long)__double_as_longlong(x);
#else
  memcpy(&u, &x, 8);
#endif
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ inline double dti_decode(unsigned long long k)
{
  unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  double x;
#ifdef __CUDA_ARCH__
  x = __longlong_as_double((long long)u);
#else
  memcpy(&x, &u, 8);
#endif
  return x;
}

// ---- frame rotations (compile-time) -----------------------------------------------------------------------
template <int DIR>
__device__ __forceinline__ S5 rot_in(Real d, Real a1, Real a2, Real a3, Real e)
{  // physical (x,y,z) vector components -> sweep frame (normal first): plmc_cuda.cu:41-58
  S5 s;
  s.v[0] = d;
  s.v[4] = e;
  if (DIR == 0) { s.v[1] = a1; s.v[2] = a2; s.v[3] = a3; }
  if (DIR == 1) { s.v[1] = a2; s.v[2] = a3; s.v[3] = a1; }
  if (DIR == 2) { s.v[1] = a3; s.v[2] = a1; s.v[3] = a2; }
  return s;
}
template <int DIR>
__device__ __forceinline__ void unrot_flux(Real const (&f)[5], Real (&o)[5])
{
  o[0] = f[0];
  o[4] = f[4];
  if (DIR == 0) { o[1] = f[1]; o[2] = f[2]; o[3] = f[3]; }
  if (DIR == 1) { o[2] = f[1]; o[3] = f[2]; o[1] = f[3]; }
  if (DIR == 2) { o[3] = f[1]; o[1] = f[2]; o[2] = f[3]; }
}

__device__ __forceinline__ Prim as_prim(S5 const &s) { return Prim{s.v[0], s.v[1], s.v[2], s.v[3], s.v[4]}; }
__device__ __forceinline__ Cons as_cons(S5 const &s) { return Cons{s.v[0], s.v[1], s.v[2], s.v[3], s.v[4]}; }
__device__ __forceinline__ S5 pack5(Prim const &q) { return S5{{q.d, q.vx, q.vy, q.vz, q.p}}; }
__device__ __forceinline__ S5 pack5(Cons const &q) { return S5{{q.d, q.mx, q.my, q.mz, q.E}}; }

__device__ __forceinline__ Face to_face(S5 const &s, Real gamma)
{
#if CHL_STRICT
  return face_from_cons(as_cons(s), gamma);
#else
  return face_from_prim(as_prim(s), gamma);
#endif
}
__device__ __forceinline__ S5 iface_state(Prim const &q, Real gamma)
{
#if CHL_STRICT
  return pack5(cons_from_prim(q, gamma));  // Write_Data (reconstruction.h:918-968)
#else
  (void)gamma;
  return pack5(q);
#endif
}

template <int DIR>
__device__ __forceinline__ S5 load_rot(const Real *__restrict__ a, size_t N, size_t id)
{
  return rot_in<DIR>(__ldg(a + id), __ldg(a + N + id), __ldg(a + 2 * N + id), __ldg(a + 3 * N + id),
                     __ldg(a + 4 * N + id));
}

// Half-step primitives of one cell, rotated into the sweep frame.  Product mode stores (rho, v, P).  Strict mode
// stores (rho, v, E) and derives P here, because the reference's Load_Data sums |v|^2 in the ROTATED component
// order (reconstruction.h:151-201 with o1,o2,o3), so the rounding of P differs between the three sweeps.
template <int DIR>
__device__ __forceinline__ Prim load_half(const Real *__restrict__ a, size_t N, size_t id, Real gamma)
{
  S5 const s = load_rot<DIR>(a, N, id);
#if CHL_STRICT
  Real const e = s.v[4] - RC(0.5) * s.v[0] * sq3(s.v[1], s.v[2], s.v[3]);
  return Prim{s.v[0], s.v[1], s.v[2], s.v[3], fmax((gamma - RC(1.)) * e, kTiny)};
#else
  (void)gamma;
  return Prim{s.v[0], s.v[1], s.v[2], s.v[3], s.v[4]};
#endif
}

// Interface states of one cell along DIR: L = low side of its upper face, R = high side of its lower face.
template <int RECON, int STAGE, int DIR>
__device__ __forceinline__ void recon_cell(const Real *__restrict__ in, size_t N, size_t id, long stride, Real gamma,
                                           S5 &L, S5 &R)
{
  if (STAGE == 0) {  // first-order: both sides are the cell itself (pcm_cuda.cu)
    S5 const u = load_rot<DIR>(in, N, id);
#if CHL_STRICT
    L = u;
#else
    L = pack5(prim_from_cons(as_cons(u), gamma));
#endif
    R = L;
    return;
  }
  Prim const wi = load_half<DIR>(in, N, id, gamma);
  if (RECON == kPCM) {
    L = iface_state(wi, gamma);
    R = L;
    return;
  }
  Prim const wm1 = load_half<DIR>(in, N, id - stride, gamma);
  Prim const wp1 = load_half<DIR>(in, N, id + stride, gamma);
  Prim l, r;
  if (RECON == kPPMC) {
    Prim const wm2 = load_half<DIR>(in, N, id - 2 * stride, gamma);
    Prim const wp2 = load_half<DIR>(in, N, id + 2 * stride, gamma);
    ppmc_cell(wm2, wm1, wi, wp1, wp2, gamma, l, r);
  } else {
    plmc_cell(wm1, wi, wp1, gamma, l, r);
  }
  L = iface_state(l, gamma);
  R = iface_state(r, gamma);
}

template <int SOLVER>
__device__ __forceinline__ void face_flux(S5 const &L, S5 const &R, Real gamma, Real (&F)[5])
{
  Cons const f = riemann_flux<SOLVER>(to_face(L, gamma), to_face(R, gamma), gamma);
  F[0] = f.d; F[1] = f.mx; F[2] = f.my; F[3] = f.mz; F[4] = f.E;
}

// One in-plane direction (DIR 0 = x, 1 = y) for the CTA's column at plane k.
// sb: shared buffer of 5 * NSLOT reals, NSLOT = TY*(TX+1) (x) or (TY+1)*TX (y); slot(face, row).
// Returns dF = F(lower face) - F(upper face), in PHYSICAL component order, for threads with do_flux.
template <int RECON, int SOLVER, int STAGE, int DIR>
__device__ __forceinline__ void sweep_inplane(const Real *__restrict__ in, size_t N, int nx, int ny, Real gamma,
                                              Real *sb, size_t id, bool do_recon, bool do_flux, int gx0, int gy0,
                                              int gk, int const (&hi)[3], Real (&dF)[5])
{
  constexpr int NSLOT = DIR == 0 ? TY * (TX + 1) : (TY + 1) * TX;
  int const tx = threadIdx.x, ty = threadIdx.y, t = ty * TX + tx;
  long const stride = DIR == 0 ? 1 : nx;
  // slot of the LOWER face of local cell (i, j)
  auto slot = [](int i, int j) -> int { return DIR == 0 ? j * (TX + 1) + i : j * TX + i; };
  int const my_lo = slot(tx, ty);
  int const my_hi = DIR == 0 ? slot(tx + 1, ty) : slot(tx, ty + 1);

  S5 L, R;
  if (do_recon) {
    recon_cell<RECON, STAGE, DIR>(in, N, id, stride, gamma, L, R);
#pragma unroll
    for (int c = 0; c < 5; c++) sb[c * NSLOT + my_lo] = R.v[c];
  }

  // halo tasks: the cell below the column (side 0: its L feeds face 0) and above it (side 1: its R feeds the
  // last face), one pair per row (x) or per column (y)
  bool h_flux = false;
  int h_slot  = 0;
  S5 Lh;
  {
    int const base = DIR == 0 ? 0 : kHaloYBase;
    int const cnt  = DIR == 0 ? TY : TX;
    int const q    = t - base;
    if (q >= 0 && q < 2 * cnt) {
      int const side = q / cnt, r = q - side * cnt;
      int hx, hy;
      if (DIR == 0) {
        hx = gx0 + (side ? TX : -1);
        hy = gy0 + r;
      } else {
        hx = gx0 + r;
        hy = gy0 + (side ? TY : -1);
      }
      bool valid;
      if (DIR == 0) valid = hy < hi[1] && (side == 0 || hx <= hi[0]);
      else valid = hx < hi[0] && (side == 0 || hy <= hi[1]);
      if (valid) {
        size_t const hid = (size_t)hx + (size_t)nx * ((size_t)hy + (size_t)ny * gk);
        S5 Rh;
        recon_cell<RECON, STAGE, DIR>(in, N, hid, stride, gamma, Lh, Rh);
        if (side == 1) {
          int const s = DIR == 0 ? slot(TX, r) : slot(r, TY);
#pragma unroll
          for (int c = 0; c < 5; c++) sb[c * NSLOT + s] = Rh.v[c];
        } else {
          h_flux = true;
          h_slot = DIR == 0 ? slot(0, r) : slot(r, 0);
        }
      }
    }
  }
  __syncthreads();

  Real F[5];
  if (do_flux) {
    S5 Rn;
#pragma unroll
    for (int c = 0; c < 5; c++) Rn.v[c] = sb[c * NSLOT + my_hi];
    face_flux<SOLVER>(L, Rn, gamma, F);
#pragma unroll
    for (int c = 0; c < 5; c++) sb[c * NSLOT + my_hi] = F[c];
  }
  if (h_flux) {
    S5 Rn;
    Real Fh[5];
#pragma unroll
    for (int c = 0; c < 5; c++) Rn.v[c] = sb[c * NSLOT + h_slot];
    face_flux<SOLVER>(Lh, Rn, gamma, Fh);
#pragma unroll
    for (int c = 0; c < 5; c++) sb[c * NSLOT + h_slot] = Fh[c];
  }
  __syncthreads();

  if (do_flux) {
    Real d[5];
#pragma unroll
    for (int c = 0; c < 5; c++) d[c] = sb[c * NSLOT + my_lo] - F[c];
    unrot_flux<DIR>(d, dF);
  }
}

template <int RECON, int SOLVER, int STAGE>
__global__ void __launch_bounds__(kThreads, 2) vl_sweep_kernel(SweepParams const P)
{
  __shared__ Real sbx[5 * TY * (TX + 1)];
  __shared__ Real sby[5 * (TY + 1) * TX];
  __shared__ double sred[kThreads / 32];

  int const nx = P.nx, ny = P.ny;
  size_t const N     = (size_t)nx * ny * P.nz;
  size_t const plane = (size_t)nx * ny;
  Real const gamma   = P.gamma;
  int const gx0 = P.lo[0] + blockIdx.x * TX, gy0 = P.lo[1] + blockIdx.y * TY;
  int const gx = gx0 + threadIdx.x, gy = gy0 + threadIdx.y;
  bool const active  = gx < P.hi[0] && gy < P.hi[1];
  bool const recon_x = gx <= P.hi[0] && gy < P.hi[1];
  bool const recon_y = gx < P.hi[0] && gy <= P.hi[1];

  int kz0 = P.lo[2], kz1 = P.hi[2];
  if (P.has_z) {
    kz0 = P.lo[2] + blockIdx.z * P.zchunk;
    kz1 = min(kz0 + P.zchunk, P.hi[2]);
  }
  int const kbeg = P.has_z ? kz0 - 1 : kz0, kend = P.has_z ? kz1 : kz1 - 1;

  S5 Lz;            // low-side state of face k-1/2 (from cell k-1)
  Real Fz_lo[5];    // flux through the lower z face of the cell awaiting completion (physical order)
  Real acc[5];      // x,y part of the update of cell k-1, waiting for its upper z flux
  Real dti = -1.7976931348623157e308;
#pragma unroll
  for (int c = 0; c < 5; c++) { Lz.v[c] = RC(1.0); Fz_lo[c] = RC(0.0); acc[c] = RC(0.0); }

  for (int k = kbeg; k <= kend; k++) {
    size_t const id = (size_t)gx + (size_t)nx * ((size_t)gy + (size_t)ny * k);
    bool const full = k >= kz0 && k < kz1;

    // ---------------- z: registers only ----------------
    if (P.has_z && active) {
      S5 Lz_new, Rz;
      recon_cell<RECON, STAGE, 2>(P.in, N, id, (long)plane, gamma, Lz_new, Rz);
      if (k >= kz0) {
        Real Fs[5], Fz[5];
        face_flux<SOLVER>(Lz, Rz, gamma, Fs);
        unrot_flux<2>(Fs, Fz);
        if (k > kz0) {  // complete cell k-1
          size_t const idm = id - plane;
          Real u[5];
          if (STAGE == 0) {
#pragma unroll
            for (int c = 0; c < 5; c++) u[c] = acc[c] + P.dtodz * (Fz_lo[c] - Fz[c]);
          } else {
#pragma unroll
            for (int c = 0; c < 5; c++) u[c] = P.out[c * N + idm] + (acc[c] + P.dtodz * (Fz_lo[c] - Fz[c]));
          }
          if (STAGE == 0) {
            if (u[0] < P.density_floor) {  // VL_3D_cuda.cu:479-493
              Real const d0 = u[0];
              u[0]          = P.density_floor;
#pragma unroll
              for (int c = 1; c < 5; c++) u[c] *= (P.density_floor / d0);
            }
            Prim const w = prim_from_cons(Cons{u[0], u[1], u[2], u[3], u[4]}, gamma);
            P.out[idm] = w.d; P.out[N + idm] = w.vx; P.out[2 * N + idm] = w.vy; P.out[3 * N + idm] = w.vz;
            P.out[4 * N + idm] = CHL_STRICT ? u[4] : w.p;
          } else {
            if (u[0] < P.density_floor && u[0] > RC(0.)) {  // hydro_cuda.cu:279-291
              Real const d0 = u[0];
              u[0]          = P.density_floor;
#pragma unroll
              for (int c = 1; c < 5; c++) u[c] *= (P.density_floor / d0);
            }
            if (P.apply_tfloor) {  // Temperature_Floor_Kernel, hydro_cuda.cu:1117-1153
              Real const d_inv = RC(1.0) / u[0];
              Real const vx = u[1] * d_inv, vy = u[2] * d_inv, vz = u[3] * d_inv;
              Real const Ekin = RC(0.5) * u[0] * (vx * vx + vy * vy + vz * vz);
              if ((u[4] - Ekin) / u[0] < P.U_floor) u[4] = Ekin + u[0] * P.U_floor;
            }
#pragma unroll
            for (int c = 0; c < 5; c++) P.out[c * N + idm] = u[c];
            if (P.do_dti)
              dti = fmax(dti, (double)inverse_crossing_time<3>(u[0], u[1], u[2], u[3], u[4], P.idx, P.idy, P.idz,
                                                               P.dx, P.dy, P.dz, gamma));
          }
        }
#pragma unroll
        for (int c = 0; c < 5; c++) Fz_lo[c] = Fz[c];
      }
      Lz = Lz_new;
    }
    if (!full) continue;  // uniform across the CTA

    // ---------------- x, then y: shared-memory exchange ----------------
    Real dF[5];
    sweep_inplane<RECON, SOLVER, STAGE, 0>(P.in, N, nx, ny, gamma, sbx, id, recon_x, active, gx0, gy0, k, P.hi, dF);
    if (active) {
      if (STAGE == 0) {
#pragma unroll
        for (int c = 0; c < 5; c++) acc[c] = __ldg(P.in + c * N + id) + P.dtodx * dF[c];
      } else {
#pragma unroll
        for (int c = 0; c < 5; c++) acc[c] = P.dtodx * dF[c];
      }
    }
    if (P.has_y) {
      sweep_inplane<RECON, SOLVER, STAGE, 1>(P.in, N, nx, ny, gamma, sby, id, recon_y, active, gx0, gy0, k, P.hi, dF);
      if (active) {
#pragma unroll
        for (int c = 0; c < 5; c++) acc[c] = acc[c] + P.dtody * dF[c];
      }
    } else {
      __syncthreads();  // 1-D: keep successive planes (there is only one) from racing on sbx
    }

    if (!P.has_z && active) {  // 1-D / 2-D: the cell is complete now (no density or temperature floor there)
      if (STAGE == 0) {
        Prim const w = prim_from_cons(Cons{acc[0], acc[1], acc[2], acc[3], acc[4]}, gamma);
        P.out[id] = w.d; P.out[N + id] = w.vx; P.out[2 * N + id] = w.vy; P.out[3 * N + id] = w.vz;
        P.out[4 * N + id] = CHL_STRICT ? acc[4] : w.p;
      } else {
        Real u[5];
#pragma unroll
        for (int c = 0; c < 5; c++) {
          u[c]              = P.out[c * N + id] + acc[c];
          P.out[c * N + id] = u[c];
        }
        if (P.do_dti) {
          Real v;
          if (P.has_y) v = inverse_crossing_time<2>(u[0], u[1], u[2], u[3], u[4], P.idx, P.idy, P.idz, P.dx, P.dy, P.dz, gamma);
          else v = inverse_crossing_time<1>(u[0], u[1], u[2], u[3], u[4], P.idx, P.idy, P.idz, P.dx, P.dy, P.dz, gamma);
          dti = fmax(dti, (double)v);
        }
      }
    }
  }

  if (STAGE == 1 && P.do_dti) {  // warp shuffle -> one value per warp -> one atomic per CTA
    double m = dti;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
    int const t = threadIdx.y * TX + threadIdx.x;
    if ((t & 31) == 0) sred[t >> 5] = m;
    __syncthreads();
    if (t < 32) {
      m = t < kThreads / 32 ? sred[t] : -1.7976931348623157e308;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
      if (t == 0) atomicMax(P.dti_key, dti_encode(m));
    }
  }
}

}  // namespace CHL_NS
