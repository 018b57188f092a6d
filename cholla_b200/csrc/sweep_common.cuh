// sweep_common.cuh -- what every sweep kernel of the hydro path shares: the interface-state container, the rotation of
// the sweep frame onto the [field][N] arrays, the per-cell reconstruction dispatch, the half-step store, and the
// out-of-line IEEE redo of a block whose branch-free quotients raised the sticky flag (exact_div.cuh, kDivFlag).
#pragma once
#include "hydro_physics.cuh"

namespace CHL_NS
{
#if CHL_STRICT
constexpr int kHalfFields = 7;  // rho, vx, vy, vz, P_x, P_y, P_z
typedef KCF KH;                 // the hydro sweeps' constants: branch-free exact quotients + sticky flag
#else
constexpr int kHalfFields = 5;  // rho, vx, vy, vz, P
typedef KC KH;
#endif

struct S5 {
  Real v[5];
#if CHL_NP > 0
  Real q[CHL_NP];  // passive fields: conserved (strict) / per unit mass (product mode)
#endif
};  // strict: sweep-frame conserved; product mode: sweep-frame primitives
struct S5x2 {
  S5 L, R;
};
struct Half {
  Real w[kHalfFields];
#if CHL_NP > 0
  Real q[CHL_NP];  // passive fields per unit mass, stored after the kHalfFields hydro entries
#endif
};
// the five hydro values of a cell plus its passive ones, as the update / half-state routines pass them around
struct U5 {
  Real u[5];
#if CHL_NP > 0
  Real q[CHL_NP];
#endif
};

// ordered-bits encoding so that an unsigned atomicMax orders doubles
__device__ __forceinline__ unsigned long long dti_encode(double x)
{
  unsigned long long const u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

template <class K>
__device__ __forceinline__ Face to_face(S5 const &s, K const &k)
{
#if CHL_STRICT
  Cons c{s.v[0], s.v[1], s.v[2], s.v[3], s.v[4]};
#else
  Prim c{s.v[0], s.v[1], s.v[2], s.v[3], s.v[4]};
#endif
#if CHL_NP > 0
#pragma unroll
  for (int i = 0; i < kNP; i++) c.q[i] = s.q[i];
#endif
#if CHL_STRICT
  return face_from_cons(c, k);
#else
  return face_from_prim(c, k);
#endif
}
template <class K>
__device__ __forceinline__ S5 iface_state(Prim const &q, K const &k)
{
#if CHL_STRICT
  Cons const c = cons_from_prim(q, k);  // Write_Data (reconstruction.h:918-968)
  S5 o{{c.d, c.mx, c.my, c.mz, c.E}};
#else
  (void)k;
  Prim const &c = q;
  S5 o{{q.d, q.vx, q.vy, q.vz, q.p}};
#endif
#if CHL_NP > 0
#pragma unroll
  for (int i = 0; i < kNP; i++) o.q[i] = c.q[i];
#endif
  return o;
}

// where the five sweep-frame components of direction `dir` live in a [field][N] array
struct DirMap {
  size_t o1, o2, o3, op;  // element offsets: normal / transverse-1 / transverse-2 vector component, fifth scalar
  long stride;            // cell stride along dir
  int f1, f2, f3;         // the vector components as field numbers (to un-rotate fluxes)
  size_t N;               // field stride
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
  m.N  = N;
  m.stride = dir == 0 ? 1 : (dir == 1 ? (long)nx : (long)nx * ny);
  return m;
}

// predictor (PCM on the conserved state): the interface state both faces of the cell see
template <class K>
__device__ __forceinline__ void pcm_from_cons(Cons const &u, K const &k, S5 &L)
{
#if CHL_STRICT
  (void)k;
  Cons const &w = u;
  L = S5{{u.d, u.mx, u.my, u.mz, u.E}};
#else
  Prim const w = prim_from_cons(u, k);
  L = S5{{w.d, w.vx, w.vy, w.vz, w.p}};
#endif
#if CHL_NP > 0
#pragma unroll
  for (int i = 0; i < kNP; i++) L.q[i] = w.q[i];
#endif
}

// Interface states of cell `id` along the direction described by m:
// L = low side of its upper face, R = high side of its lower face.
template <int RECON, int STAGE, class K>
__device__ __forceinline__ void recon_cell(const Real *__restrict__ in, size_t id, DirMap const &m, K const &k, S5 &L,
                                           S5 &R)
{
  const Real *__restrict__ c = in + id;
  if (STAGE == 0) {  // first order: both sides are the cell itself (pcm_cuda.cu); the input is conserved
    Cons u{__ldg(c), __ldg(c + m.o1), __ldg(c + m.o2), __ldg(c + m.o3), __ldg(c + m.op)};
#if CHL_NP > 0
#pragma unroll
    for (int i = 0; i < kNP; i++) u.q[i] = __ldg(c + (size_t)(5 + i) * m.N);  // conserved input: fields 5 ..
#endif
    pcm_from_cons(u, k, L);
    R = L;
    return;
  }
  if (RECON == kPCM) {
#if CHL_STRICT
    // PCM_Reconstruction_{1,2,3}D copies the conserved values (pcm_cuda.cu:13-411): for a PCM corrector the predictor
    // (or the SIMPLE integrator's elementwise pass) stored the half-step state CONSERVED (half_state, kHalfConserved),
    // the energy in all three fifth-field slots, the passive fields as they are -- nothing is converted, same bits
    L = S5{{__ldg(c), __ldg(c + m.o1), __ldg(c + m.o2), __ldg(c + m.o3), __ldg(c + m.op)}};
#if CHL_NP > 0
#pragma unroll
    for (int i = 0; i < kNP; i++) L.q[i] = __ldg(c + (size_t)(kHalfFields + i) * m.N);
#endif
#else
    Prim wi{__ldg(c), __ldg(c + m.o1), __ldg(c + m.o2), __ldg(c + m.o3), __ldg(c + m.op)};
#if CHL_NP > 0
#pragma unroll
    for (int i = 0; i < kNP; i++) wi.q[i] = __ldg(c + (size_t)(kHalfFields + i) * m.N);
#endif
    L = iface_state(wi, k);
#endif
    R = L;
    return;
  }
#if CHL_NP > 0
  GlobalStencil const ld{c, {0, m.o1, m.o2, m.o3, m.op}, m.stride, (size_t)kHalfFields * m.N, m.N};
#else
  GlobalStencil const ld{c, {0, m.o1, m.o2, m.o3, m.op}, m.stride};
#endif
  Prim l, r;
  if (RECON == kPLMP || RECON == kPPMP) {
    if (RECON == kPPMP) ppmp_stencil(ld, k, l, r);
    else plmp_stencil(ld, k, l, r);
#if CHL_STRICT
    Cons const cl = cons_from_prim_plmp(l, k), cr = cons_from_prim_plmp(r, k);
    L = S5{{cl.d, cl.mx, cl.my, cl.mz, cl.E}};
    R = S5{{cr.d, cr.mx, cr.my, cr.mz, cr.E}};
#if CHL_NP > 0
#pragma unroll
    for (int i = 0; i < kNP; i++) {
      L.q[i] = cl.q[i];
      R.q[i] = cr.q[i];
    }
#endif
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

// `plmp` argument of half_state: 0 = PLMC / PPMC, 1 = PLMP / PPMP (their summation order of the pressures), 2 = PCM
constexpr int kHalfConserved = 2;
__host__ __device__ inline int half_kind(int recon) { return recon == kPCM ? kHalfConserved : (recon == kPLMP || recon == kPPMP) ? 1 : 0; }
// Half-step state of one cell, from its conserved half-step values: what Load_Data (reconstruction.h:139-212)
// derives, stored once instead of being recomputed by every stencil that touches the cell.
template <class K>
__device__ __forceinline__ Half half_state(U5 const &c, K const &k, int plmp)
{
  Real const gamma = k.gamma;
  Real const(&u)[5] = c.u;
  Cons uc{u[0], u[1], u[2], u[3], u[4]};
#if CHL_NP > 0
#pragma unroll
  for (int i = 0; i < kNP; i++) uc.q[i] = c.q[i];
#endif
  Half h;
#if CHL_STRICT
  if (plmp == kHalfConserved) {  // a PCM corrector reads the conserved values themselves (recon_cell)
#pragma unroll
    for (int f = 0; f < kHalfFields; f++) h.w[f] = u[f < 4 ? f : 4];
#if CHL_NP > 0
#pragma unroll
    for (int i = 0; i < kNP; i++) h.q[i] = c.q[i];
#endif
    return h;
  }
#endif
  Prim const w = prim_from_cons(uc, k);  // P with the x-sweep association
#if CHL_NP > 0
#pragma unroll
  for (int i = 0; i < kNP; i++) h.q[i] = w.q[i];
#endif
  h.w[0] = w.d;
  h.w[1] = w.vx;
  h.w[2] = w.vy;
  h.w[3] = w.vz;
  h.w[4] = w.p;
#if CHL_STRICT
  if (plmp) {
    // PLMP and PPMP derive the stencil pressures with the plain left-to-right sum of the rotated velocities and
    // multiply by (gamma - 1) on the right (plmp_cuda.cu:117-118, ppmp_cuda.cu:168-170)
    Real const ex = u[4] - RC(0.5) * w.d * (w.vx * w.vx + w.vy * w.vy + w.vz * w.vz);
    Real const ey = u[4] - RC(0.5) * w.d * (w.vy * w.vy + w.vz * w.vz + w.vx * w.vx);
    Real const ez = u[4] - RC(0.5) * w.d * (w.vz * w.vz + w.vx * w.vx + w.vy * w.vy);
#if CHL_DE  // plmp_cuda.cu:116-124, ppmp_cuda.cu:163-170: Get_Pressure_From_DE on the same remainder
    h.w[4] = de_pressure(u[4], ex, c.q[kNP - 1], k);
    h.w[5] = de_pressure(u[4], ey, c.q[kNP - 1], k);
    h.w[6] = de_pressure(u[4], ez, c.q[kNP - 1], k);
#else
    h.w[4] = max_s(ex * (gamma - RC(1.)), kTiny);
    h.w[5] = max_s(ey * (gamma - RC(1.)), kTiny);
    h.w[6] = max_s(ez * (gamma - RC(1.)), kTiny);
#endif
    return h;
  }
  // y sweep sums vy^2 + (vz^2 + vx^2), z sweep vz^2 + (vx^2 + vy^2)
  Real const ey = u[4] - RC(0.5) * w.d * sq3(w.vy, w.vz, w.vx);
  Real const ez = u[4] - RC(0.5) * w.d * sq3(w.vz, w.vx, w.vy);
#if CHL_DE
  h.w[5] = de_pressure(u[4], ey, c.q[kNP - 1], k);
  h.w[6] = de_pressure(u[4], ez, c.q[kNP - 1], k);
#else
  h.w[5] = max_s((gamma - RC(1.)) * ey, kTiny);
  h.w[6] = max_s((gamma - RC(1.)) * ez, kTiny);
#endif
#else
  (void)plmp;
  (void)gamma;
#endif
  return h;
}

// The corrector's per-cell epilogue after the flux differences are in: temperature floor
// (Temperature_Floor_Kernel, hydro_cuda.cu:1117-1153) and the cell's inverse crossing time.  u is updated in place.
// SHARED (strict sweeps, last corrector pass): 1 / d is formed once for the temperature floor and the CFL term (the same
// expression on the same d: hydro_cuda.cu:1133, 534) and the cell sizes' shared reciprocals come from cell_dens.
template <bool SHARED = false, class K>
__device__ __forceinline__ Real cell_epilogue(Real (&u)[5], K const &k, int apply_tfloor, Real U_floor, int do_dti,
                                              int ndim, Real idx, Real idy, Real idz, Real dx, Real dy, Real dz,
                                              const Real *cell_dens = nullptr)
{
#if CHL_STRICT
  Den dd     = Den{RC(0.), RC(0.)};
  Real d_inv = RC(0.);
  if (SHARED) {
    dd    = den(u[0], k);
    d_inv = qdiv<false>(RC(1.0), dd, k);
  }
#endif
  if (apply_tfloor) {
#if CHL_STRICT
    if (!SHARED) {
      dd    = den(u[0], k);
      d_inv = qdiv<false>(RC(1.0), dd, k);
    }
#else
    Real const d_inv = RC(1.0) / u[0];
#endif
    Real const vx = u[1] * d_inv, vy = u[2] * d_inv, vz = u[3] * d_inv;
    Real const Ekin = RC(0.5) * u[0] * (vx * vx + vy * vy + vz * vz);
#if CHL_STRICT
    if (qdiv<false>(u[4] - Ekin, dd, k) < U_floor) u[4] = Ekin + u[0] * U_floor;
#else
    if ((u[4] - Ekin) / u[0] < U_floor) u[4] = Ekin + u[0] * U_floor;
#endif
  }
  Real v = RC(0.0);
  if (do_dti) {
#if CHL_STRICT
    const Real *const di = &d_inv;
#else
    const Real *const di = nullptr;
#endif
    if (ndim == 3) v = inverse_crossing_time<3, SHARED>(u[0], u[1], u[2], u[3], u[4], idx, idy, idz, dx, dy, dz, k, di, cell_dens);
    else if (ndim == 2) v = inverse_crossing_time<2, SHARED>(u[0], u[1], u[2], u[3], u[4], idx, idy, idz, dx, dy, dz, k, di, cell_dens);
    else v = inverse_crossing_time<1, SHARED>(u[0], u[1], u[2], u[3], u[4], idx, idy, idz, dx, dy, dz, k, di, cell_dens);
  }
  return v;
}

// ---- IEEE redo of a flagged block (strict mode): out of line, reached only when a quotient or square root of the
// block left the fast-path window (denormal / huge operands, a zero where none is expected, a NaN) ---------------------
#if CHL_STRICT
template <int RECON, int STAGE>
static __device__ __noinline__ S5x2 recon_cell_ieee(const Real *in, size_t id, DirMap m, Real gamma, Real dtodx)
{
  KCS k   = make_kct<KCS>(gamma);
  k.dtodx = dtodx;
  S5x2 o;
  recon_cell<RECON, STAGE>(in, id, m, k, o.L, o.R);
  return o;
}
template <int SOLVER>
static __device__ __noinline__ Cons riemann_ieee(S5 L, S5 R, Real gamma)
{
  KCS const k = make_kct<KCS>(gamma);
  return riemann_flux<SOLVER>(to_face(L, k), to_face(R, k), k);
}
static __device__ __noinline__ Half half_state_ieee(U5 c, Real gamma, int plmp)
{
  KCS const k = make_kct<KCS>(gamma);
  return half_state(c, k, plmp);
}
struct Epi {
  Real E, dti;
};
static __device__ __noinline__ Epi cell_epilogue_ieee(Real u0, Real u1, Real u2, Real u3, Real u4, Real gamma,
                                                      int apply_tfloor, Real U_floor, int do_dti, int ndim, Real dx,
                                                      Real dy, Real dz)
{
  KCS const k = make_kct<KCS>(gamma);
  Real u[5]   = {u0, u1, u2, u3, u4};
  Epi e;
  e.dti = cell_epilogue(u, k, apply_tfloor, U_floor, do_dti, ndim, RC(0.), RC(0.), RC(0.), dx, dy, dz);
  e.E   = u[4];
  return e;
}
#endif

// ---- the blocks as the kernels call them: branch-free arithmetic, one flag test, IEEE redo ---------------------------
template <int RECON, int STAGE>
__device__ __forceinline__ void recon_cell_checked(const Real *__restrict__ in, size_t id, DirMap const &m, KH const &k,
                                                   S5 &L, S5 &R)
{
  recon_cell<RECON, STAGE>(in, id, m, k, L, R);
#if CHL_STRICT
  if (k.bad) {
    S5x2 const o = recon_cell_ieee<RECON, STAGE>(in, id, m, k.gamma, k.dtodx);
    L     = o.L;
    R     = o.R;
    k.bad = false;
  }
#endif
}
template <int SOLVER>
__device__ __forceinline__ Cons riemann_checked(S5 const &L, S5 const &R, KH const &k)
{
  Cons f = riemann_flux<SOLVER>(to_face(L, k), to_face(R, k), k);
#if CHL_STRICT
  if (k.bad) {
    f     = riemann_ieee<SOLVER>(L, R, k.gamma);
    k.bad = false;
  }
#endif
  return f;
}
__device__ __forceinline__ Half half_state_checked(U5 const &c, KH const &k, int plmp)
{
  Half h = half_state(c, k, plmp);
#if CHL_STRICT
  if (k.bad) {
    h     = half_state_ieee(c, k.gamma, plmp);
    k.bad = false;
  }
#endif
  return h;
}
template <bool SHARED = false>
__device__ __forceinline__ Real cell_epilogue_checked(Real (&u)[5], KH const &k, int apply_tfloor, Real U_floor,
                                                      int do_dti, int ndim, Real idx, Real idy, Real idz, Real dx,
                                                      Real dy, Real dz, const Real *cell_dens = nullptr)
{
#if CHL_STRICT
  Real const E0 = u[4];
#endif
  Real v = cell_epilogue<SHARED>(u, k, apply_tfloor, U_floor, do_dti, ndim, idx, idy, idz, dx, dy, dz, cell_dens);
#if CHL_STRICT
  if (k.bad) {
    Epi const e = cell_epilogue_ieee(u[0], u[1], u[2], u[3], E0, k.gamma, apply_tfloor, U_floor, do_dti, ndim, dx, dy, dz);
    u[4]  = e.E;
    v     = e.dti;
    k.bad = false;
  }
#endif
  return v;
}

__device__ __forceinline__ void store_half(Real *__restrict__ out, size_t N, size_t id, Half const &h)
{
#pragma unroll
  for (int f = 0; f < kHalfFields; f++) out[(size_t)f * N + id] = h.w[f];
#if CHL_NP > 0
#pragma unroll
  for (int i = 0; i < kNP; i++) out[(size_t)(kHalfFields + i) * N + id] = h.q[i];
#endif
}

}  // namespace CHL_NS
