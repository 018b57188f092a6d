// vl_predictor.cuh -- the Van Leer predictor of a 3-D grid as ONE launch (option: chl_set_tuning(ctx, 0, 1)).
//
// Measured on B200 at 512^3, strict arithmetic (profiles/r3_predictor_fused.md): DRAM traffic 14.2 GB instead of 50.4 GB,
// warp instructions 8.9 G instead of 10.0 G -- and 16.4 ms instead of 13.4 ms: 80 registers at 2 x 352 threads per SM
// spill (660 M local requests), the two barriers per plane cost 21 % of the stall samples, and 22 resident warps do not
// hide the rest.  The predictor is issue-bound, not HBM-bound, on this part, so the three sweeps stay the default; this
// form is kept, bit-identical and tested, for parts with less fp64 throughput per byte of HBM bandwidth.
//
// Reference chain (VL_3D_cuda.cu:141-212): PCM_Reconstruction_3D -> Calculate_*_Fluxes_CUDA x3 ->
// Update_Conserved_Variables_3D_half, every intermediate through HBM.  The direction-split form (vl_dirsweep.cuh, three
// launches) already keeps interface states and fluxes on chip but still reads U three times and round-trips the
// running sum twice: 50 GB per step at 512^3 (profiles/r3_prof_strict_512.md) and -- what costs the time, the passes
// being issue-bound -- converts every cell's state into solver form SIX times (as the low and as the high side of a
// face, in three directions).  Here a CTA owns a 32 x 8 column of cells and marches along z:
//   phase A  every cell of plane p (the column, plus the one-cell rim of it the x / y faces need: three halo warps) is
//            loaded once and converted once -- reciprocal of the density, the three velocities, the three pressures
//            (the reference's solvers sum v^2 in the rotated component order, so the rounding of P differs per
//            direction: strict mode keeps all three) -- into a two-plane ring in shared memory;
//   phase B  the z face below plane p (ring slots p-1 and p of the thread's own cell) and the +x / +y faces of plane
//            p-1 (its neighbours' ring entries) are solved; x / y fluxes go to a shared flux buffer;
//   phase C  cell p-1 = ((U + dt/dx dFx) + dt/dy dFy) + dt/dz dFz in the reference's association (VL_3D_cuda.cu:
//            444-461), density floor, half-step primitives -> W.
// Two barriers per plane; the carried state is five doubles (the lower z flux).  HBM traffic: U once (40 B) + W once
// (56 B strict, 40 B product) per cell instead of 367 B.  The arithmetic per face is riemann_flux() on the same Face
// values the direction-split kernels build, so the bits are the same (tests/test_gpu_parity.py, test_gpu_reference.py).
#pragma once
#include "sweep_common.cuh"

namespace CHL_NS
{
constexpr int kPTX = 32, kPTY = 8;                       // cells per CTA column
constexpr int kPMain    = kPTX * kPTY;                   // main threads: one cell each
constexpr int kPThreads = kPMain + 3 * 32;               // + x-halo warp, y-low warp, y-high warp
constexpr int kPSlots   = kPMain + 2 * kPTY + 2 * kPTX;  // ring slots per plane: main | x-low | x-high | y-low | y-high
constexpr int kPFields  = 12;                            // d, 1/d, vx, vy, vz, mx, my, mz, E, P_x, P_y, P_z
constexpr int kPFx      = (kPTX + 1) * kPTY;             // x faces per plane: face f of row j at j * 33 + f
constexpr int kPFy      = kPTX * (kPTY + 1);             // y faces per plane: face f of column i at f * 32 + i
constexpr size_t kPSmemBytes = (size_t)(2 * kPFields * kPSlots + 5 * (kPFx + kPFy)) * sizeof(Real);
#ifndef CHL_PRED_MINB
#define CHL_PRED_MINB 2
#endif

struct PredParams {
  const Real *U;  // conserved state [5][N]
  Real *W;        // half-step primitives [kHalfFields][N]
  int nx, ny, nz;
  int lo[3], hi[3];  // output region [lo, hi): every cell but the outermost ring (VL_3D_cuda.cu:408)
  int zchunk;        // output planes per CTA
  Real dto[3];       // dt/2 / dx, dy, dz
  Real gamma, density_floor;
  int plmp;
};

// the pressure a solver derives from a Q_L / Q_R entry whose velocity components arrive as (a, b, c): hllc_cuda.cu:103-131
__device__ __forceinline__ Real face_pressure(Real d, Real E, Real a, Real b, Real c, Real gamma)
{
  Real const p = (E - RC(0.5) * d * (a * a + b * b + c * c)) * (gamma - RC(1.0));
  return max_s(p, kTiny);
}

struct CellForm {  // one cell in solver form, field order
  Real d, y, v[3], m[3], E, p[3];
};
template <class K>
__device__ __forceinline__ CellForm cell_form(Real d, Real mx, Real my, Real mz, Real E, K const &k)
{
  CellForm c;
#if CHL_STRICT
  // face_from_cons() of the three rotations at once: they share the reciprocal and the three quotients
  Den const dd = den(d, k);
  c.d = d; c.y = dd.y;
  c.m[0] = mx; c.m[1] = my; c.m[2] = mz;
  c.E = E;
  c.v[0] = qdiv(mx, dd, k);
  c.v[1] = qdiv(my, dd, k);
  c.v[2] = qdiv(mz, dd, k);
  c.p[0] = face_pressure(d, E, c.v[0], c.v[1], c.v[2], k.gamma);
  c.p[1] = face_pressure(d, E, c.v[1], c.v[2], c.v[0], k.gamma);
  c.p[2] = face_pressure(d, E, c.v[2], c.v[0], c.v[1], k.gamma);
#else
  // product arithmetic: primitives once, then what face_from_prim() builds from them
  Prim const w = prim_from_cons(Cons{d, mx, my, mz, E}, k);
  c.d = w.d; c.y = RC(0.);
  c.v[0] = w.vx; c.v[1] = w.vy; c.v[2] = w.vz;
  c.p[0] = c.p[1] = c.p[2] = max_s(w.p, kTiny);
  c.m[0] = w.d * w.vx; c.m[1] = w.d * w.vy; c.m[2] = w.d * w.vz;
  c.E = c.p[0] * k.ig1 + RC(0.5) * w.d * (w.vx * w.vx + w.vy * w.vy + w.vz * w.vz);
#endif
  return c;
}
#if CHL_STRICT
static __device__ __noinline__ CellForm cell_form_ieee(Real d, Real mx, Real my, Real mz, Real E, Real gamma)
{
  KCS const k = make_kct<KCS>(gamma);
  return cell_form(d, mx, my, mz, E, k);
}
#endif

// ring entry `s` of one plane -> the Face the solver sees when sweeping along DIR
template <int DIR>
__device__ __forceinline__ Face face_of(const Real *__restrict__ ring, int s)
{
  constexpr int a = DIR, b = (DIR + 1) % 3, c = (DIR + 2) % 3;
  Face f;
  f.d  = ring[0 * kPSlots + s];
  f.vx = ring[(2 + a) * kPSlots + s];
  f.vy = ring[(2 + b) * kPSlots + s];
  f.vz = ring[(2 + c) * kPSlots + s];
  f.mx = ring[(5 + a) * kPSlots + s];
  f.my = ring[(5 + b) * kPSlots + s];
  f.mz = ring[(5 + c) * kPSlots + s];
  f.E  = ring[8 * kPSlots + s];
  f.p  = ring[(9 + DIR) * kPSlots + s];
#if CHL_STRICT
  f.dd = Den{f.d, ring[1 * kPSlots + s]};
#endif
  return f;
}

// flux through the face between ring entries sl (low side) and sr (high side), un-rotated to field order
template <int SOLVER, int DIR>
__device__ __forceinline__ void face_flux(const Real *__restrict__ ring_l, int sl, const Real *__restrict__ ring_r, int sr,
                                          KH const &k, Real (&F)[5])
{
  constexpr int a = DIR, b = (DIR + 1) % 3, c = (DIR + 2) % 3;
  Face const l = face_of<DIR>(ring_l, sl), r = face_of<DIR>(ring_r, sr);
  Cons f = riemann_flux<SOLVER>(l, r, k);
#if CHL_STRICT
  if (k.bad) {  // a quotient left the fast-path window: redo from the conserved states with IEEE divisions
    f     = riemann_ieee<SOLVER>(S5{{l.d, l.mx, l.my, l.mz, l.E}}, S5{{r.d, r.mx, r.my, r.mz, r.E}}, k.gamma);
    k.bad = false;
  }
#endif
  F[0]     = f.d;
  F[1 + a] = f.mx;
  F[1 + b] = f.my;
  F[1 + c] = f.mz;
  F[4]     = f.E;
}

template <int SOLVER>
__global__ void __launch_bounds__(kPThreads, CHL_PRED_MINB) vl_predictor_kernel(PredParams const P)
{
  extern __shared__ __align__(16) unsigned char pred_smem[];
  Real *const ring = reinterpret_cast<Real *>(pred_smem);      // [2][kPFields][kPSlots]
  Real *const fxb  = ring + 2 * kPFields * kPSlots;            // [5][kPFx]
  Real *const fyb  = fxb + 5 * kPFx;                           // [5][kPFy]
  int const nx = P.nx, ny = P.ny, nz = P.nz;
  size_t const N = (size_t)nx * ny * nz, plane = (size_t)nx * ny;
  int const t = threadIdx.x, lane = t & 31, warp = t >> 5;
  int const gx0 = P.lo[0] + blockIdx.x * kPTX, gy0 = P.lo[1] + blockIdx.y * kPTY;
  int const kz0 = P.lo[2] + blockIdx.z * P.zchunk, kz1 = min(kz0 + P.zchunk, P.hi[2]);
  KH const kc = make_kct<KH>(P.gamma);

  // ---- role: which cell this thread converts, where it lives in the ring, which faces it solves ----
  // main (warp < 8): cell (gx0 + lane, gy0 + warp): +x face, +y face, lower z face, and the update of the cell.
  // warp 8: lanes 0-7 the cells left of each row (+x face = face 0 of the row), lanes 8-15 the cells right of each row.
  // warp 9: the cells below each column (+y face = face 0 of the column); warp 10: the cells above each column.
  bool const is_main = warp < kPTY;
  int cx, cy, slot;
  int xn = -1, yn = -1;    // ring slot of the +x / +y neighbour (-1: this thread solves no such face)
  int fxs = -1, fys = -1;  // flux-buffer slot of the +x / +y face
  if (is_main) {
    cx = gx0 + lane; cy = gy0 + warp;
    slot = warp * kPTX + lane;
    xn  = lane < kPTX - 1 ? slot + 1 : kPMain + kPTY + warp;
    yn  = warp < kPTY - 1 ? slot + kPTX : kPMain + 2 * kPTY + kPTX + lane;
    fxs = warp * (kPTX + 1) + lane + 1;
    fys = (warp + 1) * kPTX + lane;
  } else if (warp == kPTY) {
    int const r = lane & (kPTY - 1), hi_side = (lane >> 3) & 1;
    cx = hi_side ? gx0 + kPTX : gx0 - 1; cy = gy0 + r;
    slot = kPMain + hi_side * kPTY + r;
    if (lane >= 2 * kPTY) cx = -1;  // idle lanes
    if (!hi_side) { xn = r * kPTX; fxs = r * (kPTX + 1); }
  } else {
    int const hi_side = warp - (kPTY + 1);
    cx = gx0 + lane; cy = hi_side ? gy0 + kPTY : gy0 - 1;
    slot = kPMain + 2 * kPTY + hi_side * kPTX + lane;
    if (!hi_side) { yn = lane; fys = lane; }
  }
  bool const in_array = cx >= 0 && cx < nx && cy >= 0 && cy < ny;
  // a cell is converted for the x / y faces only if it is inside the array and in (or next to) the output region
  bool const near_out = in_array && cx >= P.lo[0] - 1 && cx <= P.hi[0] && cy >= P.lo[1] - 1 && cy <= P.hi[1];
  bool const active   = is_main && cx < P.hi[0] && cy < P.hi[1];  // this thread updates its cell
  // the +x face of cell cx is wanted when its row is an output row and the face bounds an output cell (cx or cx + 1);
  // both cells are then inside the array and converted (near_out).  Likewise +y.
  bool const want_x = xn >= 0 && in_array && cy >= P.lo[1] && cy < P.hi[1] && cx >= P.lo[0] - 1 && cx < P.hi[0];
  bool const want_y = yn >= 0 && in_array && cx >= P.lo[0] && cx < P.hi[0] && cy >= P.lo[1] - 1 && cy < P.hi[1];
  size_t const col = (size_t)(in_array ? cx : 0) + (size_t)nx * (size_t)(in_array ? cy : 0);

  Real Fz_lo[5] = {RC(0.), RC(0.), RC(0.), RC(0.), RC(0.)};
  for (int p = kz0 - 1; p <= kz1; p++) {
    Real *const cur        = ring + (size_t)(p & 1) * kPFields * kPSlots;
    const Real *const prev = ring + (size_t)((p - 1) & 1) * kPFields * kPSlots;
    // ======================= phase A: plane p into the ring =======================
    // planes kz0-1 and kz1 only feed z faces: the column's own cells suffice there
    bool const xy_plane = p >= kz0 && p < kz1;
    if (near_out && (is_main ? (xy_plane || active) : xy_plane)) {
      size_t const id = col + plane * (size_t)p;
      Real const d = __ldg(P.U + id), mx = __ldg(P.U + N + id), my = __ldg(P.U + 2 * N + id),
                 mz = __ldg(P.U + 3 * N + id), E = __ldg(P.U + 4 * N + id);
      CellForm c = cell_form(d, mx, my, mz, E, kc);
#if CHL_STRICT
      if (kc.bad) {
        c      = cell_form_ieee(d, mx, my, mz, E, kc.gamma);
        kc.bad = false;
      }
#endif
      cur[0 * kPSlots + slot] = c.d;
      cur[1 * kPSlots + slot] = c.y;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        cur[(2 + a) * kPSlots + slot] = c.v[a];
        cur[(5 + a) * kPSlots + slot] = c.m[a];
        cur[(9 + a) * kPSlots + slot] = c.p[a];
      }
      cur[8 * kPSlots + slot] = c.E;
    }
    __syncthreads();
    // ======================= phase B: faces =======================
    int const q = p - 1;  // the plane whose x / y faces are solved and whose cells are completed in this iteration
    bool const do_xy = q >= kz0;
    Real Fz_hi[5];
    if (active && p >= kz0) face_flux<SOLVER, 2>(prev, slot, cur, slot, kc, Fz_hi);  // face p - 1/2
    if (do_xy) {
      Real F[5];
      if (want_x) {
        face_flux<SOLVER, 0>(prev, slot, prev, xn, kc, F);
#pragma unroll
        for (int f = 0; f < 5; f++) fxb[f * kPFx + fxs] = F[f];
      }
      if (want_y) {
        face_flux<SOLVER, 1>(prev, slot, prev, yn, kc, F);
#pragma unroll
        for (int f = 0; f < 5; f++) fyb[f * kPFy + fys] = F[f];
      }
    }
    __syncthreads();
    // ======================= phase C: complete cell (cx, cy, q) =======================
    if (active && do_xy) {
      size_t const id = col + plane * (size_t)q;
      U5 c;  // five-field builds only (CHL_NP = 0): the dispatcher never selects this kernel for passive fields
      Real(&u)[5] = c.u;
#pragma unroll
      for (int f = 0; f < 5; f++) {
        Real v = __ldg(P.U + (size_t)f * N + id) + P.dto[0] * (fxb[f * kPFx + fxs - 1] - fxb[f * kPFx + fxs]);
        v      = v + P.dto[1] * (fyb[f * kPFy + fys - kPTX] - fyb[f * kPFy + fys]);
        u[f]   = v + P.dto[2] * (Fz_lo[f] - Fz_hi[f]);
      }
      if (u[0] < P.density_floor) {  // VL_3D_cuda.cu:479-493
        Real const d0 = u[0];
        u[0]          = P.density_floor;
#pragma unroll
        for (int f = 1; f < 5; f++) u[f] *= (P.density_floor / d0);
      }
      store_half(P.W, N, id, half_state_checked(c, kc, P.plmp));
    }
    if (active && p >= kz0) {
#pragma unroll
      for (int f = 0; f < 5; f++) Fz_lo[f] = Fz_hi[f];
    }
    // no third barrier: the ring plane written in the next phase A is the one whose last readers (phase B above) are
    // behind the second barrier, and the flux buffers are rewritten only after the next first barrier
  }
}

}  // namespace CHL_NS
