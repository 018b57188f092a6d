// vl_dirsweep.cuh -- direction-split Van Leer sweeps: one launch per (stage, direction).
//
// Why: a single launch per stage that exchanges every interface state through shared memory (round 1's fused sweep,
// removed) moves the algorithmic minimum of HBM bytes but was nowhere near HBM-bound -- fp64 issue latency at 21
// resident warps/SM (fp64 pipe 37 % busy; top stalls: fixed-latency wait, barrier, local-memory spills, instruction
// cache misses; profiles/r1_*).  With most of the HBM bandwidth idle, the cheaper trade on B200 for the CORRECTOR is
// the opposite one: spend a few more bytes per cell and make every launch a small, barrier-free loop that keeps the
// fp64 pipe fed.  (The same holds for the predictor: its single-launch form, vl_predictor.cuh, moves 14 GB instead of
// 50 GB per step at 512^3 but takes 16.4 ms against 13.4 ms for the three sweeps here -- round-2 measurement,
// profiles/r3_predictor_fused.md -- and is kept as a tested option, chl_set_tuning.)
//
// Each launch sweeps ONE direction.  A thread (y/z sweeps) or a warp (x sweep) MARCHES along that direction and
// carries the low-side interface state and the lower-face flux from cell to cell, so every cell is reconstructed
// once and every face is solved once -- no barriers, no recomputation apart from two cells at the ends of a chunk:
//     for c = first-1 .. last+1:   cp.async: update operands of cell c-1 -> shared memory (StageGeom)
//                                  (L_c, R_c) = reconstruct(c)            L: low side of face c+1/2, R: high side of c-1/2
//                                  F_{c-1/2}  = riemann(L_{c-1}, R_c)
//                                  cell c-1  += dt/dx (F_{c-3/2} - F_{c-1/2})
// x sweep: lanes are 32 consecutive cells, "c-1" is the lane to the left (__shfl_up) and lane 0 takes the value
// lane 31 parked for it after the previous 32 cells.  y/z sweeps: lanes are 32 consecutive x (coalesced), "c-1" is
// the thread's own previous iteration, parked in a thread-private shared-memory column (CHL_DIR_PARK) so that it
// does not hold 20 registers through the reconstruction.
//
// The flux differences of the directions are summed in the reference's order through one accumulator array:
//   predictor  ((U + x) + y) + z -> half-step primitives W          (VL_3D_cuda.cu:444-461)
//   corrector  U += (x + y) + z, floors, CFL max                    (hydro_cuda.cu:236-249, 279-291, 1117-1153)
// FIRST = this pass starts the sum, LAST = this pass completes the cell.  HBM traffic per cell-update in 3-D fp64
// (product mode): predictor 40+40 | 40+40+40 | 40+40+40 B, corrector 40+40 | 40+40+40 | 40+40+40+40 B = 680 B by
// design (734 B measured, profiles/roofline_traffic.json): 3.4x the fused kernel's 200 B, 0.4x the reference's 1.84 kB.
#pragma once
#include "sweep_common.cuh"

namespace CHL_NS
{
#ifndef CHL_DIR_MINB
#define CHL_DIR_MINB 4
#endif
#ifndef CHL_DIR_MINB0  // the predictor passes (PCM + Riemann) may want a different register budget
#define CHL_DIR_MINB0 CHL_DIR_MINB
#endif
// the last corrector pass (old U staged too, floors + CFL epilogue): strict builds give it 80 registers / 6 CTAs -- 512^3
// corrector 29.21 -> 28.99 / 29.02 ms (gpurun_out/r27_run27.txt, r28_run28.txt); 96 registers: 30.3 ms; 80 registers for the
// x or y pass as well: 29.3 / 29.5 ms, 64 for the x pass: 29.2 ms
#ifndef CHL_DIR_MINB_LAST
#define CHL_DIR_MINB_LAST CHL_DIR_MINB
#endif
#ifndef CHL_DIR_THREADS
#define CHL_DIR_THREADS 128
#endif
constexpr int kDirThreads = CHL_DIR_THREADS;
constexpr int kRowsPerCta = kDirThreads / 32;  // x sweep: one row of cells per warp

struct DirParams {
  const Real *in;  // stage 0: U conserved [5][N];  stage 1: W half-step primitives [kHalfFields][N]
  const Real *U;   // stage 0, FIRST: the state the predictor starts from
  Real *acc;       // [5][N] running sum
  Real *out;       // LAST: stage 0 -> W, stage 1 -> U (updated in place)
  int nx, ny, nz;
  int lo[3], hi[3];  // output region [lo, hi)
  int chunk;         // cells per thread along the marching direction (y and z sweeps)
  int ndim;
  Real dto;  // dt / dx of this direction
  Real gamma, density_floor;
  int apply_tfloor;
  Real U_floor;
  int do_dti;
  int plmp;  // predictor, strict mode: store the half-step pressures with PLMP's / PPMP's summation order (store_half)
  Real dx, dy, dz, idx, idy, idz;
  unsigned long long *dti_key;
  unsigned long long *bad_cells;  // last corrector pass: running count of cells whose density came out <= 0
  // -DDE builds (CHL_DE), [N] reals filled by de_term_kernel before the stage: last predictor pass -- the p div v term
  // of the half-step GasEnergy (VL_3D_cuda.cu:471-477); last corrector pass -- GasEnergy after
  // Partial_Update_Advected_Internal_Energy (hydro_cuda.cu:754-833), taken instead of the old GasEnergy
  const Real *de_term;
  int xw_log2;                    // y / z sweeps: log2 of the CTA's extent along x (7 = 128 x 1 threads, 2 = 4 x 32)
  // several boxes in ONE launch (the six thin boxes of the boundary shell: no drained-GPU tail between them): a 1-D grid,
  // CTAs [cta0[b], cta0[b + 1]) belong to box b and are laid out gx x gy x (chunks) inside it.  nbox = 0: lo / hi above.
  int nbox;
  struct Box {
    int lo[3], hi[3], chunk, xw_log2, cta0, gx, gy;
  } box[6];
};

// y / z sweeps: the CTA's extent along x (128, 64 or 32 threads; the rest of its 128 threads stack along the other fixed
// axis) that leaves the fewest idle thread slots.  A CTA keeps its registers and shared memory until its last warp is
// done, so a column of CTAs with a handful of live lanes (518 = 4 x 128 + 6 cells in the 512^3 predictor, 258 = 2 x 128 + 2
// in the 256^3 MHD corrector) costs a fifth to a third of the resident warps of these latency-bound kernels.
inline int pick_xw_log2(int ox, int oo)
{
  int best = 7;
  long best_slots = -1;
  for (int l = 7; l >= 5; l--) {
    int const xw = 1 << l, ow = kDirThreads >> l;
    long const slots = (long)((ox + xw - 1) / xw) * xw * ((oo + ow - 1) / ow) * ow;
    if (best_slots < 0 || slots < best_slots) {
      best_slots = slots;
      best       = l;
    }
  }
  return best;
}

__device__ __forceinline__ void cp_async_real(Real *smem_dst, const Real *gsrc)
{
  unsigned const sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (sizeof(Real) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- staging of the update's read-modify-write operands ----------------------------------------------------
// Completing a cell needs the running sum (and, in the last corrector pass, the old U) of that cell: five or ten
// values the thread would otherwise load only after the Riemann solve, one exposed L2/HBM latency per cell (ncu at
// 512^3: half of the long-scoreboard samples of the y/z passes, 90 % in the predictor).  They are fetched with
// cp.async into a thread-private shared-memory column at the top of the iteration instead, costing no registers.
#ifndef CHL_DIR_STAGE
#define CHL_DIR_STAGE 1
#endif
// Carried values (low-side interface state, lower-face flux) parked in shared memory between iterations instead of
// occupying 20 registers through the reconstruction (the kernels spill at 64-72 registers).
#ifndef CHL_DIR_PARK
#define CHL_DIR_PARK 1
#endif
// (Fetching the predictor's NEXT cell the same way, one iteration ahead into a two-slot buffer, was measured and
// dropped: the extra 10 KB of shared memory per CTA shrinks L1 and the three predictor passes went 9.9 -> 13.3 ms.
// The register form -- the next cell's five values loaded one iteration ahead -- was measured in the strict build after
// the side cache: predictor 12.04 -> 13.06 ms at 72 registers, 12.83 ms at 80; gpurun_out/r27_run27.txt.  Dropped too.)
// First-order predictor, strict HLLC / Roe / HLL: a cell's state is the high side of face c-1/2 AND the low side of face c+1/2, and
// the solver's conversion of a Q_L / Q_R entry (three exact quotients by the density, the pressure) plus its per-side
// terms (sqrt(d), the enthalpy, the sound speed) are a quarter of the pass's instructions.  With CHL_SIDE_CACHE the
// converted state is what travels to the next cell (14 values + a "left the fast path" mark instead of 5), so each cell
// is converted once.  Same expressions, same bits.
#ifndef CHL_SIDE_CACHE
#define CHL_SIDE_CACHE 1
#endif
constexpr int kSideVals = 14;  // d, mx, my, mz, E, vx, vy, vz, p, 1/d, sqrt(d), H, cf, mark

template <int STAGE, bool FIRST, bool LAST>
struct StageGeom {
  static constexpr int n_sum = (STAGE == 1 && FIRST) ? 0 : 5;  // U (predictor, first pass) or the running sum
  static constexpr int n_out = (STAGE == 1 && LAST) ? 5 : 0;   // old U, last corrector pass
  static constexpr int n     = n_sum + n_out;
};

// value held by the cell on the low side: previous lane (x sweep) or previous iteration (y/z sweeps)
template <int DIR>
__device__ __forceinline__ Real from_low_side(Real v, Real &carry, int lane)
{
  if (DIR == 0) {
    Real t = __shfl_up_sync(0xffffffffu, v, 1);
    if (lane == 0) t = carry;
    carry = __shfl_sync(0xffffffffu, v, 31);
    return t;
  }
  Real const t = carry;
  carry        = v;
  return t;
}

// NV values at once; `park` = this thread's (y/z sweeps) or this warp's (x sweep) shared-memory slots, stride pstride
template <int DIR, int NV>
__device__ __forceinline__ void from_low_side_n(Real const (&v)[NV], Real (&carry)[NV], Real (&t)[NV], Real *park,
                                                int pstride, int lane)
{
#if CHL_DIR_PARK
  if (DIR == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) {
      t[k] = __shfl_up_sync(0xffffffffu, v[k], 1);
      if (lane == 0) t[k] = park[k * pstride];
    }
    __syncwarp();
    if (lane == 31) {
#pragma unroll
      for (int k = 0; k < NV; k++) park[k * pstride] = v[k];
    }
    __syncwarp();
  } else {
#pragma unroll
    for (int k = 0; k < NV; k++) {
      t[k]              = park[k * pstride];
      park[k * pstride] = v[k];
    }
  }
#else
#pragma unroll
  for (int k = 0; k < NV; k++) t[k] = from_low_side<DIR>(v[k], carry[k], lane);
#endif
}
template <int DIR>
__device__ __forceinline__ void from_low_side5(Real const (&v)[5], Real (&carry)[5], Real (&t)[5], Real *park,
                                               int pstride, int lane)
{
  from_low_side_n<DIR, 5>(v, carry, t, park, pstride, lane);
}

template <int STAGE, int DIR, bool FIRST, bool LAST, bool STAGED>
__device__ __forceinline__ void finish_cell(DirParams const &P, KH const &k, size_t N, size_t id, Cons const &Flo,
                                            Cons const &Fhi, double &dti, int &n_bad, const Real *staged,
                                            const Real *cell_dens = nullptr)
{
  // staged (STAGED only): this thread's shared-memory column, component k at staged[k * kDirThreads]: the five
  // running-sum / U operands in sweep-frame order, then (last corrector pass) the five old-U values in field order.
  // Passive fields (CHL_NP) are read straight from global memory.
  typedef StageGeom<STAGE, FIRST, LAST> SG;
  // sweep frame (normal, t1, t2) -> fields: (x,y,z), (y,z,x), (z,x,y)
  constexpr int field_of_comp[5] = {0, 1 + DIR, 1 + (DIR + 1) % 3, 1 + (DIR + 2) % 3, 4};
  Real const flo[5] = {Flo.d, Flo.mx, Flo.my, Flo.mz, Flo.E}, fhi[5] = {Fhi.d, Fhi.mx, Fhi.my, Fhi.mz, Fhi.E};
  U5 c;
  Real(&u)[5] = c.u;
#pragma unroll
  for (int cc = 0; cc < 5; cc++) {
    int const f  = field_of_comp[cc];
    Real const t = P.dto * (flo[cc] - fhi[cc]);
    Real v;
    if (STAGED && SG::n_sum) v = staged[cc * kDirThreads] + t;
    else if (STAGE == 0) v = (FIRST ? __ldg(P.U + f * N + id) : P.acc[f * N + id]) + t;
    else v = FIRST ? t : P.acc[f * N + id] + t;
    if (!LAST) P.acc[f * N + id] = v;
    u[f] = v;
  }
#if CHL_NP > 0
#pragma unroll
  for (int i = 0; i < kNP; i++) {
    size_t const o = (size_t)(5 + i) * N + id;
    Real const t   = P.dto * (Flo.q[i] - Fhi.q[i]);
    Real v;
    if (STAGE == 0) v = (FIRST ? __ldg(P.U + o) : P.acc[o]) + t;
    else v = FIRST ? t : P.acc[o] + t;
    if (!LAST) P.acc[o] = v;
    c.q[i] = v;
  }
#endif
  if (!LAST) return;
  bool const floors = P.ndim == 3;  // the reference's 1-D / 2-D update kernels have no floors (hydro_cuda.cu:20-170)
  if (STAGE == 0) {
#if CHL_DE
    c.q[kNP - 1] = c.q[kNP - 1] + __ldg(P.de_term + id);  // + 0.5 P (dt/dx dvx + ...), VL_3D_cuda.cu:471-477
#endif
    if (floors && u[0] < P.density_floor) {  // VL_3D_cuda.cu:479-493: momenta, Energy and GasEnergy; not the scalars
      Real const d0 = u[0];
      u[0]          = P.density_floor;
#pragma unroll
      for (int f = 1; f < 5; f++) u[f] *= (P.density_floor / d0);
#if CHL_DE
      c.q[kNP - 1] *= (P.density_floor / d0);
#endif
    }
    store_half(P.out, N, id, half_state_checked(c, k, P.plmp));
    return;
  }
#pragma unroll
  for (int f = 0; f < 5; f++) u[f] = (STAGED ? staged[(SG::n_sum + f) * kDirThreads] : P.out[f * N + id]) + u[f];
#if CHL_NP > 0
#pragma unroll
  for (int i = 0; i < kNP; i++) {
    Real old = P.out[(size_t)(5 + i) * N + id];
#if CHL_DE
    if (i == kNP - 1) old = __ldg(P.de_term + id);
#endif
    c.q[i] = old + c.q[i];
  }
#endif
  if (floors) {
    if (u[0] < P.density_floor && u[0] > RC(0.)) {  // hydro_cuda.cu:279-291
      Real const d0 = u[0];
      u[0]          = P.density_floor;
#pragma unroll
      for (int f = 1; f < 5; f++) u[f] *= (P.density_floor / d0);
#if CHL_DE
      c.q[kNP - 1] *= (P.density_floor / d0);
#endif
    }
  }
  // a cell whose density came out non-positive is counted (chl_bad_cells): the reference averages its neighbours inside
  // the in-place update, a race with no bits to match (hydro_cuda.cu:292-297), and leaves the decision to the host
  if (!(u[0] > RC(0.))) n_bad++;
  // (-DDE builds: the temperature floor and the CFL maximum follow Select_Internal_Energy / Sync_Energies, de_sync_kernel)
  Real const v = cell_epilogue_checked<CHL_STRICT != 0>(u, k, floors && P.apply_tfloor, P.U_floor, P.do_dti, P.ndim, P.idx,
                                                        P.idy, P.idz, P.dx, P.dy, P.dz, cell_dens);
#pragma unroll
  for (int f = 0; f < 5; f++) P.out[f * N + id] = u[f];
#if CHL_NP > 0
#pragma unroll
  for (int i = 0; i < kNP; i++) P.out[(size_t)(5 + i) * N + id] = c.q[i];
#endif
  if (P.do_dti) dti = fmax(dti, (double)v);
}

template <int RECON, int SOLVER, int STAGE, int DIR, bool FIRST, bool LAST>
__global__ void __launch_bounds__(kDirThreads, STAGE == 0 ? CHL_DIR_MINB0 : (LAST ? CHL_DIR_MINB_LAST : CHL_DIR_MINB)) vl_dir_kernel(DirParams const P)
{
  int const nx = P.nx, ny = P.ny;
  size_t const N = (size_t)nx * ny * P.nz;
  int const lane = threadIdx.x & 31;
  KH kc    = make_kct<KH>(P.gamma);
  kc.dtodx = P.dto;  // read by PLMC's characteristic tracing only (SIMPLE integrator)
  DirMap const m = dir_map(DIR, N, nx, ny, (STAGE == 1 && CHL_STRICT) ? 4 + DIR : 4);

  // the cells this thread visits: c = c_first .. c_last along DIR at fixed `base`; it completes cells
  // [c_first + 1, c_last - 1]
  bool live;
  size_t base;
  int c_first, c_last;
  if (DIR == 0) {
    int const y = P.lo[1] + blockIdx.x * kRowsPerCta + (threadIdx.x >> 5), z = P.lo[2] + blockIdx.y;
    live    = y < P.hi[1];
    base    = (size_t)nx * ((size_t)y + (size_t)ny * z);
    c_first = P.lo[0] - 1;
    c_last  = P.hi[0];
  } else {
    int lo[3] = {P.lo[0], P.lo[1], P.lo[2]}, hi[3] = {P.hi[0], P.hi[1], P.hi[2]};
    int chunk = P.chunk, xw_log2 = P.xw_log2, bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z;
    if (P.nbox > 0) {  // which box does this CTA belong to (uniform over the CTA)
      int b = 0;
      while (b + 1 < P.nbox && (int)blockIdx.x >= P.box[b + 1].cta0) b++;
      DirParams::Box const &B = P.box[b];
#pragma unroll
      for (int d = 0; d < 3; d++) {
        lo[d] = B.lo[d];
        hi[d] = B.hi[d];
      }
      chunk   = B.chunk;
      xw_log2 = B.xw_log2;
      int const l = blockIdx.x - B.cta0;
      bx = l % B.gx;
      by = (l / B.gx) % B.gy;
      bz = l / (B.gx * B.gy);
    }
    // a CTA covers 2^xw_log2 cells along x and 128 / 2^xw_log2 along the other fixed axis: 128 x 1 for ordinary boxes,
    // 4 x 32 for the four-cell-wide x shells of the boundary-first last pass (same sectors per warp either way)
    int const tx = threadIdx.x & ((1 << xw_log2) - 1), to = threadIdx.x >> xw_log2;
    int const x  = lo[0] + (bx << xw_log2) + tx;
    int const od = DIR == 1 ? 2 : 1;  // the other fixed axis
    int const o  = lo[od] + by * (kDirThreads >> xw_log2) + to;
    live         = x < hi[0] && o < hi[od];
    base         = DIR == 1 ? (size_t)x + (size_t)nx * ny * o : (size_t)x + (size_t)nx * o;
    int const b  = lo[DIR] + bz * chunk;
    c_first      = b - 1;
    c_last       = min(b + chunk, hi[DIR]);
  }
  double dti = -1.7976931348623157e308;
  int n_bad  = 0;

  typedef StageGeom<STAGE, FIRST, LAST> SG;
  constexpr bool kStage = CHL_DIR_STAGE && SG::n > 0;
  constexpr bool kSide  = CHL_SIDE_CACHE && CHL_STRICT && kNP == 0 && STAGE == 0 &&
                          (SOLVER == kHLLC || SOLVER == kROE || SOLVER == kHLL);
  constexpr int kStateV = kSide ? kSideVals : 5;
  constexpr int kParkV  = kStateV + 5 + 2 * kNP;  // parked per thread / warp: low-side state and lower-face flux
  constexpr int kParkN  = CHL_DIR_PARK ? (DIR == 0 ? kParkV * kRowsPerCta : kParkV * kDirThreads) : 1;
  __shared__ Real stage[kStage ? SG::n * kDirThreads : 1];
  __shared__ Real parked[kParkN];
  // last corrector pass, strict mode: the shared reciprocals of dx, dy, dz for the CFL term, one set per warp
  constexpr bool kCellDens = CHL_STRICT && STAGE == 1 && LAST;
  __shared__ Real s_cell_dens[kCellDens ? (kDirThreads / 32) * 7 : 1];
  const Real *cell_dens = nullptr;
#if CHL_STRICT
  if (kCellDens && P.do_dti) {
    Real *const w = s_cell_dens + (threadIdx.x >> 5) * 7;
    if (lane == 0) {
      Den const a = den(P.dx, kc), b = den(P.dy, kc), c = den(P.dz, kc);
      w[0] = a.b; w[1] = a.y; w[2] = b.b; w[3] = b.y; w[4] = c.b; w[5] = c.y;
      // a cell size outside the fast-path window (never, in practice): the mark sends every epilogue to its IEEE redo
      w[6] = kc.bad ? RC(1.0) : RC(0.0);
    }
    kc.bad = false;
    __syncwarp();
    cell_dens = w;
  }
#endif

  if (live) {
    Real Lc[5], Fc[5];  // carried: low-side state of the face below / flux through the face below the previous cell
#pragma unroll
    for (int c = 0; c < 5; c++) Lc[c] = Fc[c] = RC(0.0);
#if CHL_NP > 0
    Real Lqc[kNP], Fqc[kNP];
#pragma unroll
    for (int c = 0; c < kNP; c++) Lqc[c] = Fqc[c] = RC(0.0);
#endif
    int const step = DIR == 0 ? 32 : 1;
    constexpr int field_of_comp[5] = {0, 1 + DIR, 1 + (DIR + 1) % 3, 1 + (DIR + 2) % 3, 4};
    Real *const park   = parked + (DIR == 0 ? (threadIdx.x >> 5) : threadIdx.x);
    int const pstride  = DIR == 0 ? kRowsPerCta : kDirThreads;
    Real *const staged = stage + threadIdx.x;
    for (int cb = c_first; cb <= c_last; cb += step) {
      int const c      = DIR == 0 ? cb + lane : cb;
      bool const valid = c <= c_last;
      bool const face  = valid && c > c_first;  // face c-1/2 lies between two reconstructed cells
      bool const fin   = face && c > c_first + 1;  // cell c-1 is completed by this iteration
      size_t const id  = base + (size_t)c * (size_t)m.stride;
      if (kStage) {
        if (fin) {
          size_t const idp = id - (size_t)m.stride;
#pragma unroll
          for (int k = 0; k < SG::n_sum; k++) {
            const Real *src = (STAGE == 0 && FIRST) ? P.U : P.acc;
            cp_async_real(staged + k * kDirThreads, src + (size_t)field_of_comp[k] * N + idp);
          }
#pragma unroll
          for (int f = 0; f < SG::n_out; f++) cp_async_real(staged + (SG::n_sum + f) * kDirThreads, P.out + (size_t)f * N + idp);
        }
        cp_async_commit();
      }
      S5 L, R;
      if (valid) {
        recon_cell_checked<RECON, STAGE>(P.in, id, m, kc, L, R);
      } else {
#pragma unroll
        for (int k = 0; k < 5; k++) L.v[k] = R.v[k] = RC(1.0);
#if CHL_NP > 0
#pragma unroll
        for (int k = 0; k < kNP; k++) L.q[k] = R.q[k] = RC(1.0);
#endif
      }
      S5 Lp;
#if CHL_STRICT
      Face fL, fR;
      bool side_bad = false;
      if (kSide) {
        fR = to_face(R, kc);
        face_side_cache(fR, kc);
        Real const mark = kc.bad ? RC(1.0) : RC(0.0);  // this state's conversion left the fast-path window
        kc.bad          = false;
        Real const cur[kSideVals] = {fR.d, fR.mx, fR.my, fR.mz, fR.E, fR.vx, fR.vy, fR.vz, fR.p, fR.dd.y, fR.sd, fR.H, fR.cf, mark};
        Real low[kSideVals], unused[kSideVals];
        from_low_side_n<DIR, kSideVals>(cur, unused, low, park, pstride, lane);
        fL.d = low[0]; fL.mx = low[1]; fL.my = low[2]; fL.mz = low[3]; fL.E = low[4];
        fL.vx = low[5]; fL.vy = low[6]; fL.vz = low[7]; fL.p = low[8];
        fL.dd = Den{low[0], low[9]};
        fL.sd = low[10]; fL.H = low[11]; fL.cf = low[12];
#pragma unroll
        for (int k = 0; k < 5; k++) Lp.v[k] = low[k];
        side_bad = (mark != RC(0.0)) | (low[13] != RC(0.0));
      } else
#endif
      {
        from_low_side5<DIR>(L.v, Lc, Lp.v, park, pstride, lane);
      }
#if CHL_NP > 0
      from_low_side_n<DIR, kNP>(L.q, Lqc, Lp.q, park + 10 * pstride, pstride, lane);
#endif
      Real F[5];
      Cons Fhi;
      if (face) {
#if CHL_STRICT
        if (kSide) {
          kc.bad = side_bad;
          Fhi    = riemann_flux<SOLVER, true>(fL, fR, kc);
          if (kc.bad) {
            Fhi    = riemann_ieee<SOLVER>(Lp, R, kc.gamma);
            kc.bad = false;
          }
        } else
#endif
        {
          Fhi = riemann_checked<SOLVER>(Lp, R, kc);
        }
      } else {
        Fhi.d = Fhi.mx = Fhi.my = Fhi.mz = Fhi.E = RC(0.0);
#if CHL_NP > 0
#pragma unroll
        for (int k = 0; k < kNP; k++) Fhi.q[k] = RC(0.0);
#endif
      }
      F[0] = Fhi.d; F[1] = Fhi.mx; F[2] = Fhi.my; F[3] = Fhi.mz; F[4] = Fhi.E;
      Real Fp[5];
      from_low_side5<DIR>(F, Fc, Fp, park + kStateV * pstride, pstride, lane);
      Cons Flo{Fp[0], Fp[1], Fp[2], Fp[3], Fp[4]};
#if CHL_NP > 0
      from_low_side_n<DIR, kNP>(Fhi.q, Fqc, Flo.q, park + (10 + kNP) * pstride, pstride, lane);
#endif
      if (kStage) cp_async_wait_all();
      if (fin)
        finish_cell<STAGE, DIR, FIRST, LAST, kStage>(P, kc, N, id - (size_t)m.stride, Flo, Fhi, dti, n_bad, staged, cell_dens);
    }
  }

  if (STAGE == 1 && LAST) {  // warp shuffle -> one value per warp -> one atomic per CTA
    __shared__ double sred[kDirThreads / 32];
    __shared__ int sbad[kDirThreads / 32];
    double mx = dti;
    int nb    = n_bad;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
      nb += __shfl_down_sync(0xffffffffu, nb, o);
    }
    if (lane == 0) {
      sred[threadIdx.x >> 5] = mx;
      sbad[threadIdx.x >> 5] = nb;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int w = 1; w < kDirThreads / 32; w++) {
        mx = fmax(mx, sred[w]);
        nb += sbad[w];
      }
      if (P.do_dti) atomicMax(P.dti_key, dti_encode(mx));
      if (nb) atomicAdd(P.bad_cells, (unsigned long long)nb);
    }
  }
}

}  // namespace CHL_NS
