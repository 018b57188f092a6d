// vl_sweep.cuh -- the fused Van Leer sweep kernel (one launch per predictor / corrector stage).
//
// Replaces, per stage, the reference's chain of one-thread-per-cell kernels that round-trip every
// intermediate through HBM (VL_3D_cuda.cu:141-339: PCM -> 3x Riemann -> half update, then
// 3x reconstruct -> 3x Riemann -> full update, ~1.84 kB/cell-update of DRAM traffic):
//
//   STAGE 0 (predictor)  reads U (conserved, 1-cell stencil), first-order fluxes, writes the half-step
//                        state as PRIMITIVES W = (rho, vx, vy, vz, P) -- what the reference's Load_Data
//                        (reconstruction.h:139-212) would derive from U-half.  Strict mode stores three
//                        pressures (one per sweep direction): Load_Data sums |v|^2 in the rotated component
//                        order, so the rounding of P differs between the x, y and z sweeps.
//   STAGE 1 (corrector)  reads W (2- or 3-cell stencil), PLMC/PPMC + HLLC/Roe in the three directions,
//                        reads U once, writes U in place; optionally applies the temperature floor and
//                        folds the CFL max (Calc_dt_3D) into the same pass.
// HBM traffic per cell-update (product mode): 40 B read + 40 B write (stage 0), 80 B read + 40 B write
// (stage 1) = 200 B in fp64 -- SURVEY 8(d)'s algorithmic figure.  No interface or flux array ever exists in
// global memory.
//
// Work decomposition ("2.5-D"): a CTA owns a 32 x TY column of cells and marches along z, one plane per
// iteration, two barriers per plane:
//   phase 1  every cell reconstructs its interface states in x, y and z -- ONE inlined copy of the
//            reconstruction, looped over the direction.  The high-side state of each lower x/y face goes to
//            the shared exchange buffer of that direction; low-side states stay with the thread.  z states are
//            carried from plane to plane (no redundancy, no exchange).
//   phase 2  every cell solves the Riemann problem on its upper x and y face and its lower z face -- ONE
//            inlined copy of the solver -- and publishes the x/y fluxes.
//   phase 3  every cell differences the fluxes, completes cell k-1 (whose upper z flux is now known) and
//            starts the accumulation for cell k.
// The cells just outside the column whose states are needed (one per row end, one per column end) are handled
// by three dedicated "halo" warps running the same phases on those cells only, so the TY main warps carry
// identical work and the barriers are not skewed.  Per-thread state that lives across barriers or planes sits
// in shared memory, which leaves the register file to the arithmetic.  Stencil loads go through L1/L2 on the
// read-only path: the kernel is bound by the fp64 pipe, not by LSU or HBM (DESIGN.md has the ncu numbers).
#pragma once
#include "hydro_physics.cuh"

namespace CHL_NS
{
constexpr int TX = 32;
#ifndef CHL_TY
#define CHL_TY 8
#endif
constexpr int TY           = CHL_TY;
constexpr int kMainThreads = TX * TY;
constexpr int kThreads     = TX * (TY + 3);  // TY main warps + one x-halo warp + two y-halo warps
constexpr int kNSX         = TY * (TX + 1);  // x faces: slot(face i, row j) = j*(TX+1) + i
constexpr int kNSY         = (TY + 1) * TX;  // y faces: slot(col i, face j) = j*TX + i
static_assert(2 * TY <= 32, "x-halo tasks must fit in one warp");
#if CHL_STRICT
constexpr int kHalfFields = 7;  // rho, vx, vy, vz, P_x, P_y, P_z
#else
constexpr int kHalfFields = 5;  // rho, vx, vy, vz, P
#endif
// per-main-thread private shared slots (slot-major, thread-minor), in Reals
constexpr int kSlotB   = 0;   // [5] L state in y -> own y flux
constexpr int kSlotC   = 5;   // [5] R state in z at face k-1/2 -> z flux through that face
constexpr int kSlotLz  = 10;  // [2][5] L state in z at face k+1/2 from cell k, ping-pong by plane parity
constexpr int kSlotFz  = 20;  // [5] flux through the lower z face of the cell awaiting completion
constexpr int kSlotAcc = 25;  // [5] x,y part of the update of the cell awaiting completion (field order)
constexpr int kPrivMain = 30;
// slot A ([5], every thread incl. halo): L state of the thread's x (or only) task -> its own flux
constexpr size_t kSmemReals = 5 * kNSX + 5 * kNSY + 5 * kThreads + (size_t)kPrivMain * kMainThreads;
constexpr size_t kSmemBytes = kSmemReals * sizeof(Real) + 32 * sizeof(double);

struct SweepParams {
  const Real *in;  // stage 0: U conserved [5][N];  stage 1: W primitives [kHalfFields][N]
  Real *out;       // stage 0: W;                   stage 1: U (updated in place)
  int nx, ny, nz;
  int lo[3], hi[3];  // output region [lo, hi) per dimension
  int has_y, has_z;
  int zchunk;  // planes per CTA along z
  Real dto[3];
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
};  // strict: sweep-frame conserved; product mode: sweep-frame primitives

// ordered-bits encoding so that an unsigned atomicMax orders doubles
__device__ __forceinline__ unsigned long long dti_encode(double x)
{
  unsigned long long const u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

__device__ __forceinline__ Face to_face(S5 const &s, KC const &k)
{
#if CHL_STRICT
  return face_from_cons(Cons{s.v[0], s.v[1], s.v[2], s.v[3], s.v[4]}, k.gamma);
#else
  return face_from_prim(Prim{s.v[0], s.v[1], s.v[2], s.v[3], s.v[4]}, k);
#endif
}
__device__ __forceinline__ S5 iface_state(Prim const &q, KC const &k)
{
#if CHL_STRICT
  Cons const c = cons_from_prim(q, k);  // Write_Data (reconstruction.h:918-968)
  return S5{{c.d, c.mx, c.my, c.mz, c.E}};
#else
  (void)k;
  return S5{{q.d, q.vx, q.vy, q.vz, q.p}};
#endif
}

// where the five sweep-frame components of direction `dir` live in a [field][N] array
struct DirMap {
  size_t o1, o2, o3, op;  // element offsets: normal / transverse-1 / transverse-2 vector component, fifth scalar
  long stride;            // cell stride along dir
  int f1, f2, f3;         // the vector components as field numbers (to un-rotate fluxes)
};
__device__ __forceinline__ DirMap dir_map(int dir, size_t N, int nx, int ny, int fifth_field)
{
  DirMap m;
  m.f1 = 1 + dir;  // (x,y,z), (y,z,x), (z,x,y): plmc_cuda.cu:41-58
  m.f2 = 1 + (dir + 1) % 3;
  m.f3 = 1 + (dir + 2) % 3;
  m.o1 = (size_t)m.f1 * N;
  m.o2 = (size_t)m.f2 * N;
  m.o3 = (size_t)m.f3 * N;
  m.op = (size_t)fifth_field * N;
  m.stride = dir == 0 ? 1 : (dir == 1 ? (long)nx : (long)nx * ny);
  return m;
}

// predictor (PCM on the conserved state): the interface state both faces of the cell see
__device__ __forceinline__ void pcm_from_cons(Cons const &u, KC const &k, S5 &L)
{
#if CHL_STRICT
  (void)k;
  L = S5{{u.d, u.mx, u.my, u.mz, u.E}};
#else
  Prim const w = prim_from_cons(u, k.gamma);
  L = S5{{w.d, w.vx, w.vy, w.vz, w.p}};
#endif
}

// Interface states of cell `id` along the direction described by m:
// L = low side of its upper face, R = high side of its lower face.
template <int RECON, int STAGE>
__device__ __forceinline__ void recon_cell(const Real *__restrict__ in, size_t id, DirMap const &m, KC const &k, S5 &L,
                                           S5 &R)
{
  Real const gamma = k.gamma;
  (void)gamma;
  const Real *__restrict__ c = in + id;
  if (STAGE == 0) {  // first order: both sides are the cell itself (pcm_cuda.cu); the input is conserved
    pcm_from_cons(Cons{__ldg(c), __ldg(c + m.o1), __ldg(c + m.o2), __ldg(c + m.o3), __ldg(c + m.op)}, k, L);
    R = L;
    return;
  }
  if (RECON == kPCM) {
    Prim const wi{__ldg(c), __ldg(c + m.o1), __ldg(c + m.o2), __ldg(c + m.o3), __ldg(c + m.op)};
    L = iface_state(wi, k);
    R = L;
    return;
  }
  GlobalStencil const ld{c, {0, m.o1, m.o2, m.o3, m.op}, m.stride};
  Prim l, r;
  if (RECON == kPLMP || RECON == kPPMP) {
    if (RECON == kPPMP) ppmp_stencil(ld, k, l, r);
    else plmp_stencil(ld, k, l, r);
#if CHL_STRICT
    Cons const cl = cons_from_prim_plmp(l, k), cr = cons_from_prim_plmp(r, k);
    L = S5{{cl.d, cl.mx, cl.my, cl.mz, cl.E}};
    R = S5{{cr.d, cr.mx, cr.my, cr.mz, cr.E}};
#else
    L = iface_state(l, k);
    R = iface_state(r, k);
#endif
    return;
  }
  if (RECON == kPPMC) ppmc_stencil(ld, k, l, r);
  else plmc_stencil<RECON == kPLMCT>(ld, k, l, r);
  L = iface_state(l, k);
  R = iface_state(r, k);
}

// the same for a PLMC / PPMC stencil served by any accessor (vl_dirsweep.cuh: the shared-memory row ring)
template <int RECON, class A>
__device__ __forceinline__ void recon_from(A const &ld, KC const &k, S5 &L, S5 &R)
{
  Prim l, r;
  if (RECON == kPPMC) ppmc_stencil(ld, k, l, r);
  else plmc_stencil(ld, k, l, r);
  L = iface_state(l, k);
  R = iface_state(r, k);
}

// Half-step state of one cell, from its conserved half-step values: what Load_Data (reconstruction.h:139-212)
// derives, stored once instead of being recomputed by every stencil that touches the cell.
__device__ __forceinline__ void store_half(Real *__restrict__ out, size_t N, size_t id, Real const (&u)[5], Real gamma,
                                           int plmp)
{
  Prim const w = prim_from_cons(Cons{u[0], u[1], u[2], u[3], u[4]}, gamma);  // P with the x-sweep association
  out[id]         = w.d;
  out[N + id]     = w.vx;
  out[2 * N + id] = w.vy;
  out[3 * N + id] = w.vz;
#if CHL_STRICT
  if (plmp) {
    // PLMP and PPMP derive the stencil pressures with the plain left-to-right sum of the rotated velocities and
    // multiply by (gamma - 1) on the right (plmp_cuda.cu:117-118, ppmp_cuda.cu:168-170)
    Real const ex = u[4] - RC(0.5) * w.d * (w.vx * w.vx + w.vy * w.vy + w.vz * w.vz);
    Real const ey = u[4] - RC(0.5) * w.d * (w.vy * w.vy + w.vz * w.vz + w.vx * w.vx);
    Real const ez = u[4] - RC(0.5) * w.d * (w.vz * w.vz + w.vx * w.vx + w.vy * w.vy);
    out[4 * N + id] = max_s(ex * (gamma - RC(1.)), kTiny);
    out[5 * N + id] = max_s(ey * (gamma - RC(1.)), kTiny);
    out[6 * N + id] = max_s(ez * (gamma - RC(1.)), kTiny);
    return;
  }
  out[4 * N + id] = w.p;
  // y sweep sums vy^2 + (vz^2 + vx^2), z sweep vz^2 + (vx^2 + vy^2)
  Real const ey = u[4] - RC(0.5) * w.d * sq3(w.vy, w.vz, w.vx);
  Real const ez = u[4] - RC(0.5) * w.d * sq3(w.vz, w.vx, w.vy);
  out[5 * N + id] = max_s((gamma - RC(1.)) * ey, kTiny);
  out[6 * N + id] = max_s((gamma - RC(1.)) * ez, kTiny);
#else
  (void)plmp;
  out[4 * N + id] = w.p;
#endif
}

template <int RECON, int SOLVER, int STAGE>
__global__ void __launch_bounds__(kThreads, (TY <= 8 ? 2 : 1)) vl_sweep_kernel(SweepParams const P)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Real *const sbx    = reinterpret_cast<Real *>(smem_raw);  // [5][kNSX]
  Real *const sby    = sbx + 5 * kNSX;                      // [5][kNSY]
  Real *const pvA    = sby + 5 * kNSY;                      // [5][kThreads]
  Real *const pvM    = pvA + 5 * kThreads;                  // [kPrivMain][kMainThreads]
  double *const sred = reinterpret_cast<double *>(pvM + (size_t)kPrivMain * kMainThreads);

  int const nx = P.nx, ny = P.ny;
  size_t const N     = (size_t)nx * ny * P.nz;
  size_t const plane = (size_t)nx * ny;
  Real const gamma   = P.gamma;
  KC const kc        = make_kc(gamma);
  int const t = threadIdx.x, lane = t & 31, warp = t >> 5;
  int const gx0 = P.lo[0] + blockIdx.x * TX, gy0 = P.lo[1] + blockIdx.y * TY;
  int const ndir = 1 + (P.has_y ? 1 : 0) + (P.has_z ? 1 : 0);

  // ---- role of this thread ----
  // main (warp < TY): cell (gx0+lane, gy0+warp), one task in every direction.
  // halo: a single task in a single direction -- x: warp TY, lanes [0,TY) the cells left of each row, lanes
  // [TY,2TY) the cells right of each row;  y: warp TY+1 the cells below each column, warp TY+2 the cells above.
  bool const is_main = warp < TY;
  int hdir = -1, hside = 0, cx, cy;
  if (is_main) {
    cx = gx0 + lane;
    cy = gy0 + warp;
  } else if (warp == TY) {
    hdir  = lane < 2 * TY ? 0 : -1;
    hside = lane >= TY ? 1 : 0;
    cx    = gx0 + (hside ? TX : -1);
    cy    = gy0 + (lane - hside * TY);
  } else {
    hdir  = 1;
    hside = warp - (TY + 1);
    cx    = gx0 + lane;
    cy    = gy0 + (hside ? TY : -1);
  }
  if (hdir == 1 && !P.has_y) hdir = -1;
  bool const active = is_main && cx < P.hi[0] && cy < P.hi[1];  // this thread updates cell (cx, cy, k)
  int const li = cx - gx0, lj = cy - gy0;
  // does this thread reconstruct / solve in x and in y?  (a cell one past the end of a ragged column still has to
  // publish its R state)
  bool rec[2], flx[2];
  int slot_lo[2], slot_hi[2];  // exchange slots of the task's lower / upper face
  slot_lo[0] = lj * (TX + 1) + li;
  slot_hi[0] = slot_lo[0] + 1;
  slot_lo[1] = lj * TX + li;
  slot_hi[1] = slot_lo[1] + TX;
  if (is_main) {
    rec[0] = cx <= P.hi[0] && cy < P.hi[1];
    rec[1] = cx < P.hi[0] && cy <= P.hi[1] && P.has_y;
    flx[0] = active;
    flx[1] = active && P.has_y;
  } else {
    bool const vx_ = hdir == 0 && cy < P.hi[1] && (hside == 0 || cx <= P.hi[0]);
    bool const vy_ = hdir == 1 && cx < P.hi[0] && (hside == 0 || cy <= P.hi[1]);
    rec[0] = vx_;
    rec[1] = vy_;
    flx[0] = vx_ && hside == 0;  // the cell below the column owns face 0; the cell above only feeds its R state
    flx[1] = vy_ && hside == 0;
  }
  Real *const myA = pvA + t;                      // + c*kThreads
  Real *const myM = pvM + (is_main ? t : 0);      // + slot*kMainThreads (main threads only)

  int kz0 = P.lo[2], kz1 = P.hi[2];
  if (P.has_z) {
    kz0 = P.lo[2] + blockIdx.z * P.zchunk;
    kz1 = min(kz0 + P.zchunk, P.hi[2]);
  }
  int const kbeg = P.has_z ? kz0 - 1 : kz0, kend = P.has_z ? kz1 : kz1 - 1;
  double dti = -1.7976931348623157e308;

  for (int k = kbeg; k <= kend; k++) {
    bool const full = k >= kz0 && k < kz1;  // plane whose cells get x/y work (uniform over the CTA)
    size_t const id = (size_t)cx + (size_t)nx * ((size_t)cy + (size_t)ny * k);

    // ======================= phase 1: reconstruction =======================
    for (int dir = 0; dir < ndir; dir++) {
      bool const zdir = dir == 2;
      if (zdir ? !active : !(full && rec[dir])) continue;
      DirMap const m = dir_map(dir, N, nx, ny, (STAGE == 1 && CHL_STRICT) ? 4 + dir : 4);
      S5 L, R;
      recon_cell<RECON, STAGE>(P.in, id, m, kc, L, R);
      if (zdir) {
        Real *const lz = myM + (size_t)(kSlotLz + 5 * (k & 1)) * kMainThreads;
#pragma unroll
        for (int c = 0; c < 5; c++) {
          myM[(size_t)(kSlotC + c) * kMainThreads] = R.v[c];
          lz[(size_t)c * kMainThreads]            = L.v[c];
        }
      } else {
        Real *const sb  = dir == 0 ? sbx : sby;
        int const ns    = dir == 0 ? kNSX : kNSY;
        Real *const myL = (dir == 0 || !is_main) ? myA : myM + (size_t)kSlotB * kMainThreads;
        int const st    = (dir == 0 || !is_main) ? kThreads : kMainThreads;
        bool const pub  = is_main || hside == 1;
        // a halo cell above the column publishes into the last face of its row / column
        int const ps = is_main ? slot_lo[dir] : (dir == 0 ? lj * (TX + 1) + TX : TY * TX + li);
#pragma unroll
        for (int c = 0; c < 5; c++) {
          if (pub) sb[c * ns + ps] = R.v[c];
          myL[(size_t)c * st] = L.v[c];
        }
      }
    }
    if (full) __syncthreads();

    // ======================= phase 2: Riemann problems =======================
    for (int dir = 0; dir < ndir; dir++) {
      bool const zdir = dir == 2;
      if (zdir ? !(active && k >= kz0) : !(full && flx[dir])) continue;
      S5 L, R;
      Real *dst, *pub = nullptr;  // where the flux goes: the thread's own slot, and (x/y) the exchange buffer
      int dst_stride, pub_stride = 0;
      if (zdir) {  // face k-1/2: L from cell k-1 (previous plane), R from cell k (this plane)
        Real const *const lz = myM + (size_t)(kSlotLz + 5 * ((k - 1) & 1)) * kMainThreads;
        dst        = myM + (size_t)kSlotC * kMainThreads;
        dst_stride = kMainThreads;
#pragma unroll
        for (int c = 0; c < 5; c++) {
          L.v[c] = lz[(size_t)c * kMainThreads];
          R.v[c] = dst[(size_t)c * dst_stride];
        }
      } else {
        Real *const sb = dir == 0 ? sbx : sby;
        pub_stride     = dir == 0 ? kNSX : kNSY;
        dst            = (dir == 0 || !is_main) ? myA : myM + (size_t)kSlotB * kMainThreads;
        dst_stride     = (dir == 0 || !is_main) ? kThreads : kMainThreads;
        // main: upper face of the cell; halo cell below the column: face 0 of its row / column
        pub = sb + (is_main ? slot_hi[dir] : (dir == 0 ? lj * (TX + 1) : li));
#pragma unroll
        for (int c = 0; c < 5; c++) {
          L.v[c] = dst[(size_t)c * dst_stride];
          R.v[c] = pub[c * pub_stride];
        }
      }
      Cons const f    = riemann_flux<SOLVER>(to_face(L, kc), to_face(R, kc), gamma);
      Real const F[5] = {f.d, f.mx, f.my, f.mz, f.E};
#pragma unroll
      for (int c = 0; c < 5; c++) {
        dst[(size_t)c * dst_stride] = F[c];
        if (pub) pub[c * pub_stride] = F[c];
      }
    }
    if (full) __syncthreads();

    // ======================= phase 3: updates =======================
    if (active) {
      Real *const acc = myM + (size_t)kSlotAcc * kMainThreads;  // acc[f * kMainThreads]
      Real *const fzl = myM + (size_t)kSlotFz * kMainThreads;
      Real *const fzn = myM + (size_t)kSlotC * kMainThreads;
      if (P.has_z && k >= kz0) {
        if (k > kz0) {  // complete cell k-1: its upper z flux (face k-1/2) is in fzn, its lower one in fzl
          size_t const idm = id - plane;
          // z sweep frame (mz, mx, my): component 1 -> field 3, 2 -> field 1, 3 -> field 2
          int const comp_of_field[5] = {0, 2, 3, 1, 4};
          Real u[5];
#pragma unroll
          for (int f = 0; f < 5; f++) {
            int const c    = comp_of_field[f];
            Real const dFz = fzl[(size_t)c * kMainThreads] - fzn[(size_t)c * kMainThreads];
            Real const a   = acc[(size_t)f * kMainThreads];
            if (STAGE == 0) u[f] = a + P.dto[2] * dFz;
            else u[f] = P.out[f * N + idm] + (a + P.dto[2] * dFz);
          }
          if (STAGE == 0) {
            if (u[0] < P.density_floor) {  // VL_3D_cuda.cu:479-493
              Real const d0 = u[0];
              u[0]          = P.density_floor;
#pragma unroll
              for (int f = 1; f < 5; f++) u[f] *= (P.density_floor / d0);
            }
            store_half(P.out, N, idm, u, gamma, 0);
          } else {
            if (u[0] < P.density_floor && u[0] > RC(0.)) {  // hydro_cuda.cu:279-291
              Real const d0 = u[0];
              u[0]          = P.density_floor;
#pragma unroll
              for (int f = 1; f < 5; f++) u[f] *= (P.density_floor / d0);
            }
            if (P.apply_tfloor) {  // Temperature_Floor_Kernel, hydro_cuda.cu:1117-1153
              Real const d_inv = RC(1.0) / u[0];
              Real const vx = u[1] * d_inv, vy = u[2] * d_inv, vz = u[3] * d_inv;
              Real const Ekin = RC(0.5) * u[0] * (vx * vx + vy * vy + vz * vz);
              if ((u[4] - Ekin) / u[0] < P.U_floor) u[4] = Ekin + u[0] * P.U_floor;
            }
#pragma unroll
            for (int f = 0; f < 5; f++) P.out[f * N + idm] = u[f];
            if (P.do_dti)
              dti = fmax(dti, (double)inverse_crossing_time<3>(u[0], u[1], u[2], u[3], u[4], P.idx, P.idy, P.idz,
                                                               P.dx, P.dy, P.dz, gamma));
          }
        }
#pragma unroll
        for (int c = 0; c < 5; c++) fzl[(size_t)c * kMainThreads] = fzn[(size_t)c * kMainThreads];
      }
      if (full) {
        // x: acc = U + dt/dx (F_{i-1/2} - F_{i+1/2})   [predictor]   or   dt/dx (...)   [corrector]
        {
          Real const *const own = myA;
#pragma unroll
          for (int c = 0; c < 5; c++) {
            Real const dF = sbx[c * kNSX + slot_lo[0]] - own[(size_t)c * kThreads];
            // x sweep frame = field order
            if (STAGE == 0) acc[(size_t)c * kMainThreads] = __ldg(P.in + c * N + id) + P.dto[0] * dF;
            else acc[(size_t)c * kMainThreads] = P.dto[0] * dF;
          }
        }
        if (P.has_y) {
          Real const *const own = myM + (size_t)kSlotB * kMainThreads;
          // y sweep frame (my, mz, mx): component 1 -> field 2, 2 -> field 3, 3 -> field 1
          int const field_of_comp[5] = {0, 2, 3, 1, 4};
#pragma unroll
          for (int c = 0; c < 5; c++) {
            Real const dF = sby[c * kNSY + slot_lo[1]] - own[(size_t)c * kMainThreads];
            int const f   = field_of_comp[c];
            acc[(size_t)f * kMainThreads] = acc[(size_t)f * kMainThreads] + P.dto[1] * dF;
          }
        }
        if (!P.has_z) {  // 1-D / 2-D: the cell is complete (no density or temperature floor in those builds)
          Real u[5];
          if (STAGE == 0) {
#pragma unroll
            for (int f = 0; f < 5; f++) u[f] = acc[(size_t)f * kMainThreads];
            store_half(P.out, N, id, u, gamma, 0);
          } else {
#pragma unroll
            for (int f = 0; f < 5; f++) {
              u[f]              = P.out[f * N + id] + acc[(size_t)f * kMainThreads];
              P.out[f * N + id] = u[f];
            }
            if (P.do_dti) {
              Real v;
              if (P.has_y)
                v = inverse_crossing_time<2>(u[0], u[1], u[2], u[3], u[4], P.idx, P.idy, P.idz, P.dx, P.dy, P.dz, gamma);
              else
                v = inverse_crossing_time<1>(u[0], u[1], u[2], u[3], u[4], P.idx, P.idy, P.idz, P.dx, P.dy, P.dz, gamma);
              dti = fmax(dti, (double)v);
            }
          }
        }
      }
    }
    // no third barrier: the exchange slot a thread reads in phase 3 is the one it overwrites itself in the next
    // phase 1, and every other reuse is separated by one of the two barriers above
  }

  if (STAGE == 1 && P.do_dti) {  // warp shuffle -> one value per warp -> one atomic per CTA
    double m = dti;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
    if (lane == 0) sred[warp] = m;
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
