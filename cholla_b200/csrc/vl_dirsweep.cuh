// vl_dirsweep.cuh -- direction-split Van Leer sweeps: one launch per (stage, direction).
//
// Why: the fused single-launch sweep (vl_sweep.cuh) moves the algorithmic minimum of HBM bytes but ncu shows it
// is nowhere near HBM-bound -- it is bound by fp64 issue latency at 21 resident warps/SM (fp64 pipe 37 % busy,
// top stalls: fixed-latency wait, barrier, local-memory spills, instruction-cache misses; profiles/r1_*).  With
// >90 % of HBM bandwidth idle, the cheaper trade on B200 is the opposite one: spend a few more bytes per cell and
// make every launch a small, barrier-free, spill-free loop that keeps the fp64 pipe fed.
//
// Each launch sweeps ONE direction.  A thread (y/z sweeps) or a warp (x sweep) MARCHES along that direction and
// carries the low-side interface state and the lower-face flux from cell to cell, so every cell is reconstructed
// once and every face is solved once -- no shared memory, no barriers, no recomputation apart from two cells at
// the ends of a chunk:
//     for c = first-1 .. last+1:   (L_c, R_c) = reconstruct(c)            L: low side of face c+1/2, R: high side of c-1/2
//                                  F_{c-1/2}  = riemann(L_{c-1}, R_c)
//                                  cell c-1  += dt/dx (F_{c-3/2} - F_{c-1/2})
// x sweep: lanes are 32 consecutive cells, "c-1" is the lane to the left (__shfl_up) and lane 0 takes the value
// lane 31 carried over from the previous 32 cells.  y/z sweeps: lanes are 32 consecutive x (coalesced), "c-1" is
// the thread's own previous iteration.
//
// The flux differences of the directions are summed in the reference's order through one accumulator array:
//   predictor  ((U + x) + y) + z -> half-step primitives W          (VL_3D_cuda.cu:444-461)
//   corrector  U += (x + y) + z, floors, CFL max                    (hydro_cuda.cu:236-249, 279-291, 1117-1153)
// FIRST = this pass starts the sum, LAST = this pass completes the cell.  HBM traffic per cell-update in 3-D fp64
// (product mode): predictor 40+40 | 40+40+40 | 40+40+40 B, corrector 40+40 | 40+40+40 | 40+40+40+40 B = 680 B,
// i.e. 3.4x the fused kernel's 200 B and 0.37x the reference's 1.84 kB.
#pragma once
#include "vl_sweep.cuh"

namespace CHL_NS
{
#ifndef CHL_DIR_MINB
#define CHL_DIR_MINB 4
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
  Real dx, dy, dz, idx, idy, idz;
  unsigned long long *dti_key;
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

template <int STAGE, int DIR, bool FIRST, bool LAST>
__device__ __forceinline__ void finish_cell(DirParams const &P, size_t N, size_t id, Real const (&Flo)[5],
                                            Real const (&Fhi)[5], double &dti)
{
  // sweep frame (normal, t1, t2) -> fields: (x,y,z), (y,z,x), (z,x,y)
  constexpr int field_of_comp[5] = {0, 1 + DIR, 1 + (DIR + 1) % 3, 1 + (DIR + 2) % 3, 4};
  Real u[5];
#pragma unroll
  for (int c = 0; c < 5; c++) {
    int const f  = field_of_comp[c];
    Real const t = P.dto * (Flo[c] - Fhi[c]);
    Real v;
    if (STAGE == 0) v = (FIRST ? __ldg(P.U + f * N + id) : P.acc[f * N + id]) + t;
    else v = FIRST ? t : P.acc[f * N + id] + t;
    if (!LAST) P.acc[f * N + id] = v;
    u[f] = v;
  }
  if (!LAST) return;
  bool const floors = P.ndim == 3;  // the reference's 1-D / 2-D update kernels have no floors (hydro_cuda.cu:20-170)
  if (STAGE == 0) {
    if (floors && u[0] < P.density_floor) {  // VL_3D_cuda.cu:479-493
      Real const d0 = u[0];
      u[0]          = P.density_floor;
#pragma unroll
      for (int f = 1; f < 5; f++) u[f] *= (P.density_floor / d0);
    }
    store_half(P.out, N, id, u, P.gamma);
    return;
  }
#pragma unroll
  for (int f = 0; f < 5; f++) u[f] = P.out[f * N + id] + u[f];
  if (floors) {
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
  }
#pragma unroll
  for (int f = 0; f < 5; f++) P.out[f * N + id] = u[f];
  if (P.do_dti) {
    Real v;
    if (P.ndim == 3)
      v = inverse_crossing_time<3>(u[0], u[1], u[2], u[3], u[4], P.idx, P.idy, P.idz, P.dx, P.dy, P.dz, P.gamma);
    else if (P.ndim == 2)
      v = inverse_crossing_time<2>(u[0], u[1], u[2], u[3], u[4], P.idx, P.idy, P.idz, P.dx, P.dy, P.dz, P.gamma);
    else
      v = inverse_crossing_time<1>(u[0], u[1], u[2], u[3], u[4], P.idx, P.idy, P.idz, P.dx, P.dy, P.dz, P.gamma);
    dti = fmax(dti, (double)v);
  }
}

template <int RECON, int SOLVER, int STAGE, int DIR, bool FIRST, bool LAST>
__global__ void __launch_bounds__(kDirThreads, CHL_DIR_MINB) vl_dir_kernel(DirParams const P)
{
  int const nx = P.nx, ny = P.ny;
  size_t const N = (size_t)nx * ny * P.nz;
  int const lane = threadIdx.x & 31;
  Real const gamma = P.gamma;
  KC const kc      = make_kc(gamma);
  DirMap const m   = dir_map(DIR, N, nx, ny, (STAGE == 1 && CHL_STRICT) ? 4 + DIR : 4);

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
    int const x = P.lo[0] + blockIdx.x * kDirThreads + threadIdx.x;
    int const o = DIR == 1 ? P.lo[2] + blockIdx.y : P.lo[1] + blockIdx.y;  // the other fixed coordinate
    live        = x < P.hi[0];
    base        = DIR == 1 ? (size_t)x + (size_t)nx * ny * o : (size_t)x + (size_t)nx * o;
    int const b = P.lo[DIR] + blockIdx.z * P.chunk;
    c_first     = b - 1;
    c_last      = min(b + P.chunk, P.hi[DIR]);
  }
  double dti = -1.7976931348623157e308;

  if (live) {
    Real Lc[5], Fc[5];  // carried: low-side state of the face below / flux through the face below the previous cell
#pragma unroll
    for (int c = 0; c < 5; c++) Lc[c] = Fc[c] = RC(0.0);
    int const step = DIR == 0 ? 32 : 1;
    for (int cb = c_first; cb <= c_last; cb += step) {
      int const c      = DIR == 0 ? cb + lane : cb;
      bool const valid = c <= c_last;
      size_t const id  = base + (size_t)c * (size_t)m.stride;
      S5 L, R;
      if (valid) {
        recon_cell<RECON, STAGE>(P.in, id, m, kc, L, R);
      } else {
#pragma unroll
        for (int k = 0; k < 5; k++) L.v[k] = R.v[k] = RC(1.0);
      }
      S5 Lp;
#pragma unroll
      for (int k = 0; k < 5; k++) Lp.v[k] = from_low_side<DIR>(L.v[k], Lc[k], lane);
      bool const face = valid && c > c_first;  // face c-1/2 lies between two reconstructed cells
      Real F[5];
      if (face) {
        Cons const f = riemann_flux<SOLVER>(to_face(Lp, kc), to_face(R, kc), gamma);
        F[0] = f.d; F[1] = f.mx; F[2] = f.my; F[3] = f.mz; F[4] = f.E;
      } else {
#pragma unroll
        for (int k = 0; k < 5; k++) F[k] = RC(0.0);
      }
      Real Fp[5];
#pragma unroll
      for (int k = 0; k < 5; k++) Fp[k] = from_low_side<DIR>(F[k], Fc[k], lane);
      if (face && c > c_first + 1) finish_cell<STAGE, DIR, FIRST, LAST>(P, N, id - (size_t)m.stride, Fp, F, dti);
    }
  }

  if (STAGE == 1 && LAST && P.do_dti) {  // warp shuffle -> one value per warp -> one atomic per CTA
    __shared__ double sred[kDirThreads / 32];
    double mx = dti;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    if (lane == 0) sred[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int w = 1; w < kDirThreads / 32; w++) mx = fmax(mx, sred[w]);
      atomicMax(P.dti_key, dti_encode(mx));
    }
  }
}

}  // namespace CHL_NS
