// mhd_physics.cuh -- per-cell / per-face device arithmetic of the Van Leer MHD path (HLLD + constrained transport).
//
// Written from scratch for sm_100a; the *algorithms* are the reference's (cholla @ 8478c9e):
//   mhd_utilities.h:48-236          magnetosonic / Alfven speeds, magnetic energy, total pressure, cell-centred B
//   reconstruction.h:310-361        Compute_Eigenvectors            -> Eigen::init()
//   reconstruction.h:378-421        Primitive_To_Characteristic     -> Eigen::project()
//   reconstruction.h:450-481        Characteristic_To_Primitive     -> Eigen::unproject()
//   plmc_cuda.cu:60-128,287-300     PLMC (MHD)                      -> mhd_plmc()
//   ppmc_cuda.cu:576-684            PPMC_VL (MHD)                   -> mhd_ppmc()
//   hlld_cuda.cu:31-509             HLLD (Miyoshi & Kusano 2005)    -> hlld_flux()
//   ct_electric_fields.h:66-116     _ctSlope                        -> ct_slope()
// Expressions keep the reference's association in BOTH arithmetic modes; the strict build (-fmad=false) therefore
// reproduces the reference's bits (it calls the same CUDA math functions: sqrt, rsqrt, rhypot, copysign, fmin, fmax),
// the product build differs by FMA contraction and the straight-line square roots / shared PPM of hydro_physics.cuh.
#pragma once
#include "hydro_physics.cuh"

namespace CHL_NS
{
constexpr Real kHlldSmall = RC(1.0e-8);  // hlld_cuda.h:35

// A denominator several quotients share.  Strict: exact_div.cuh's Den / qdiv (the bits of a / b).  Product mode: one
// straight-line reciprocal and a multiplication per quotient.
// The policy of exact_div.cuh applies (K::kDiv): the sweeps evaluate their quotients branch-free with a sticky flag and redo a
// flagged block with IEEE divisions; msq() is the square root with exact zeros (a vanishing slow speed) on the fast path.
#if CHL_STRICT
typedef Den RDen;
template <class K>
__device__ __forceinline__ RDen rden(Real b, K const &k) { return den(b, k); }
// NZ: the numerator cannot be an exact zero (a pressure, a density, a constant): exact_div.cuh's ZERO_OK = false hint
template <bool NZ = false, class K>
__device__ __forceinline__ Real rdiv(Real a, RDen const &d, K const &k) { return qdiv<!NZ>(a, d, k); }
template <class K>
__device__ __forceinline__ Real msq(Real x, K const &k) { return sqrt_r<true>(x, k); }
#else
struct RDen {
  Real y;
};
template <class K>
__device__ __forceinline__ RDen rden(Real b, K const &k) { (void)k; return RDen{rcp_r(b)}; }
template <bool NZ = false, class K>
__device__ __forceinline__ Real rdiv(Real a, RDen const &d, K const &k) { (void)k; return a * d.y; }
template <class K>
__device__ __forceinline__ Real msq(Real x, K const &k) { (void)k; return sqrt_r(x); }
#endif

__device__ __forceinline__ Real dot3(Real a1, Real a2, Real a3, Real b1, Real b2, Real b3)
{
  return a1 * b1 + ((a2 * b2) + (a3 * b3));  // math_utilities.h:78-82
}
__device__ __forceinline__ Real mag_energy(Real bx, Real by, Real bz) { return RC(0.5) * dot3(bx, by, bz, bx, by, bz); }
__device__ __forceinline__ Real total_pressure(Real p, Real bx, Real by, Real bz)
{
  return max_s(p + mag_energy(bx, by, bz), kTiny);
}
template <class K>
__device__ __forceinline__ Real msonic_disc(Real p, Real bx, Real by, Real bz, K const &k, Real &t1)
{
  Real const gamma = k.gamma;
  Real const bx2 = bx * bx;
  Real const b2  = bx2 + ((by * by) + (bz * bz));
  t1             = gamma * p + b2;
  Real t2        = (t1 * t1) - RC(4.) * gamma * p * bx2;
#if CHL_STRICT
  t2 = msq(t2, k);
#else
  t2 = t2 > RC(0.0) ? sqrt_r(t2) : RC(0.0);  // (gamma p - bx^2)^2 when the transverse field vanishes: can be an exact zero
#endif
  return t2;
}
// _magnetosonicSpeed (mhd_utilities.h:48-73), wave = +1 fast / -1 slow
template <class K>
__device__ __forceinline__ Real magnetosonic(Real d, Real p, Real bx, Real by, Real bz, K const &k, Real wave)
{
  Real t1;
  Real const t2 = msonic_disc(p, bx, by, bz, k, t1);
  return msq(rdiv(t1 + wave * t2, rden(RC(2.0) * max_s(d, kTiny), k), k), k);
}
// fast and slow speed of one state: the discriminant and the division by 2 rho are shared
template <class K>
__device__ __forceinline__ void magnetosonic_pair(Real d, Real p, Real bx, Real by, Real bz, K const &k, Real &cf, Real &cs)
{
  Real t1;
  Real const t2 = msonic_disc(p, bx, by, bz, k, t1);
  RDen const dd = rden(RC(2.0) * max_s(d, kTiny), k);
  cf = msq(rdiv<true>(t1 + RC(1.0) * t2, dd, k), k);
  cs = msq(rdiv(t1 + RC(-1.0) * t2, dd, k), k);
}
template <class K>
__device__ __forceinline__ Real alfven_speed(Real bx, Real d, K const &k)
{
  return rdiv(fabs(bx), rden(msq(max_s(d, kTiny), k), k), k);
}
// hydro_utilities.h:34-45 / 106-118 (MHD)
__device__ __forceinline__ Real mhd_pressure(Real E, Real d, Real vx, Real vy, Real vz, Real gamma, Real bx, Real by,
                                             Real bz)
{
  Real p = E - RC(0.5) * d * dot3(vx, vy, vz, vx, vy, vz);
  p -= mag_energy(bx, by, bz);
  return max_s((gamma - RC(1.)) * p, kTiny);
}
template <class K>
__device__ __forceinline__ Real mhd_energy(Real P, Real d, Real vx, Real vy, Real vz, K const &k, Real bx, Real by,
                                           Real bz)
{
#if CHL_STRICT
  Real e = qdiv<false>(max_s(P, kTiny), k.g1, k) + RC(0.5) * d * dot3(vx, vy, vz, vx, vy, vz);
#else
  Real e = (max_s(P, kTiny) * k.ig1) + RC(0.5) * d * dot3(vx, vy, vz, vx, vy, vz);
#endif
  e += mag_energy(bx, by, bz);
  return e;
}

struct MPrim {  // sweep frame; bx is the cell-centred normal field (eigenvectors only), by/bz the transverse ones
  Real d, vx, vy, vz, p, bx, by, bz;
};
struct M7 {  // what travels between reconstruction and solver: (d, m1, m2, m3, E, By, Bz), sweep frame
  Real v[7];
};

// ---- eigen-system of one cell (Stone et al. 2008, appendix A) ---------------------------------------------------
struct Eigen {
  Real cf, cs, cf2, cs2, alpha_f, alpha_s, beta_y, beta_z, n_fs, sign, q_f, q_s, a_f, a_s, qp_f, qp_s, ap_f, ap_s;
  Real d, a2, isd, sqd;
  RDen dd, da2;  // the two denominators of Primitive_To_Characteristic: rho and a^2
  template <class K>
  __device__ __forceinline__ void init(MPrim const &w, K const &k)
  {
    Real const gamma = k.gamma;
    dd           = rden(w.d, k);
    Real const a = msq(rdiv<true>(gamma * w.p, dd, k), k);  // hydro_utilities::Calc_Sound_Speed
    a2           = a * a;
    da2          = rden(a2, k);
    d            = w.d;
    magnetosonic_pair(w.d, w.p, w.bx, w.by, w.bz, k, cf, cs);
    cf2 = cf * cf;
    cs2 = cs * cs;
    Real const denom = cf2 - cs2, num2 = cf2 - a2;
    if (denom <= RC(0.0) || num2 <= RC(0.0)) {
      alpha_f = RC(1.0);
      alpha_s = RC(0.0);
    } else {
      Real const num1 = a2 - cs2;
      if (num1 <= RC(0.0)) {
        alpha_f = RC(0.0);
        alpha_s = RC(1.0);
      } else {
        RDen const dn = rden(denom, k);
        alpha_f       = msq(rdiv(num1, dn, k), k);
        alpha_s       = msq(rdiv(num2, dn, k), k);
      }
    }
    Real const bd = rhypot(w.by, w.bz);
    beta_y        = isfinite(bd) ? w.by * bd : RC(1.0);
    beta_z        = isfinite(bd) ? w.bz * bd : RC(0.0);
    sign          = copysign(RC(1.0), w.bx);
    n_fs          = rdiv<true>(RC(0.5), da2, k);
    qp_f          = sign * n_fs * alpha_f * cf;
    qp_s          = sign * n_fs * alpha_s * cs;
    q_f           = sign * alpha_f * cf;
    q_s           = sign * alpha_s * cs;
    sqd           = msq(w.d, k);
    a_f           = alpha_f * a * sqd;
    a_s           = alpha_s * a * sqd;
    RDen const das = rden(a * sqd, k);
    ap_f           = rdiv(RC(0.5) * alpha_f, das, k);
    ap_s           = rdiv(RC(0.5) * alpha_s, das, k);
    isd           = rsqrt(w.d);
  }
  // s = (d, vx, vy, vz, p, by, bz) of a state or of a difference
  template <class K>
  __device__ __forceinline__ void project(Real sd, Real svx, Real svy, Real svz, Real sp, Real sby, Real sbz,
                                          Real (&o)[7], K const &k) const
  {
    Real const spd = rdiv(sp, dd, k);
    o[0] = n_fs * alpha_f * (spd - cf * svx) + qp_s * (beta_y * svy + beta_z * svz) +
           ap_s * (beta_y * sby + beta_z * sbz);
    o[1] = RC(0.5) * (beta_y * (sbz * sign * isd + svz) - beta_z * (sby * sign * isd + svy));
    o[2] = n_fs * alpha_s * (spd - cs * svx) - qp_f * (beta_y * svy + beta_z * svz) -
           ap_f * (beta_y * sby + beta_z * sbz);
    o[3] = sd - rdiv(sp, da2, k);
    o[4] = n_fs * alpha_s * (spd + cs * svx) + qp_f * (beta_y * svy + beta_z * svz) -
           ap_f * (beta_y * sby + beta_z * sbz);
    o[5] = RC(0.5) * (beta_y * (sbz * sign * isd - svz) - beta_z * (sby * sign * isd - svy));
    o[6] = n_fs * alpha_f * (spd + cf * svx) - qp_s * (beta_y * svy + beta_z * svz) +
           ap_s * (beta_y * sby + beta_z * sbz);
  }
  __device__ __forceinline__ MPrim unproject(Real const (&a)[7]) const
  {
    MPrim o;
    o.d  = d * (alpha_f * (a[0] + a[6]) + alpha_s * (a[2] + a[4])) + a[3];
    o.vx = cf * alpha_f * (a[6] - a[0]) + cs * alpha_s * (a[4] - a[2]);
    o.vy = beta_y * (q_s * (a[0] - a[6]) + q_f * (a[4] - a[2])) + beta_z * (a[5] - a[1]);
    o.vz = beta_z * (q_s * (a[0] - a[6]) + q_f * (a[4] - a[2])) + beta_y * (a[1] - a[5]);
    o.p  = d * a2 * (alpha_f * (a[0] + a[6]) + alpha_s * (a[2] + a[4]));
    o.by = beta_y * (a_s * (a[0] + a[6]) - a_f * (a[2] + a[4])) - beta_z * sign * sqd * (a[5] + a[1]);
    o.bz = beta_z * (a_s * (a[0] + a[6]) - a_f * (a[2] + a[4])) + beta_y * sign * sqd * (a[5] + a[1]);
    o.bx = RC(0.0);
    return o;
  }
};

// ---- reconstructions on the stencil of half-step primitives -----------------------------------------------------
// W holds per cell (rho, vx, vy, vz, P, Bcc_x, Bcc_y, Bcc_z [, P_y, P_z in strict mode]); the caller passes the
// element offsets of the sweep-frame components (normal velocity, two transverse, pressure of this sweep, the three
// cell-centred field components) and the cell stride along the sweep.
struct MOff {
  size_t v1, v2, v3, p, b1, b2, b3;
  long s;
};
__device__ __forceinline__ MPrim load_mprim(const Real *__restrict__ c, MOff const &o, long k)
{
  const Real *q = c + k * o.s;
  MPrim w;
  w.d  = __ldg(q);
  w.vx = __ldg(q + o.v1);
  w.vy = __ldg(q + o.v2);
  w.vz = __ldg(q + o.v3);
  w.p  = __ldg(q + o.p);
  w.bx = __ldg(q + o.b1);
  w.by = __ldg(q + o.b2);
  w.bz = __ldg(q + o.b3);
  return w;
}

__device__ __forceinline__ void mhd_floor(MPrim &l, MPrim &r)
{
  r.d = max_s(r.d, kTiny);
  l.d = max_s(l.d, kTiny);
  r.p = max_s(r.p, kTiny);
  l.p = max_s(l.p, kTiny);
}

// PPMC_VL (MHD).  The reference projects the five cells on all seven characteristics first (35 values) and then limits
// them; with the 22-double eigen-system that is > 60 live doubles.  Here the characteristics are produced one wave
// family at a time -- the four magneto-sonic ones, the two Alfven ones, the entropy one -- each family reloading the few
// primitives it needs (L1) and handing its 5-cell stencils straight to ppm_single, so the projection-only part of the
// eigen-system dies after the first family.  Every expression keeps Eigen::project()'s association (same bits).
template <class K>
__device__ __forceinline__ void mhd_ppmc(const Real *__restrict__ c, MOff const &o, K const &k, MPrim &L, MPrim &R)
{
  Eigen e;
  e.init(load_mprim(c, o, 0), k);
  Real cl[7], cr[7];
  {  // fast / slow magneto-sonic pairs: characteristics 0, 6 (fast) and 2, 4 (slow)
    Real const nf = e.n_fs * e.alpha_f, ns = e.n_fs * e.alpha_s;
    Real t0[5], t6[5], t2[5], t4[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {
      const Real *q  = c + (long)(i - 2) * o.s;
      Real const svx = __ldg(q + o.v1), svy = __ldg(q + o.v2), svz = __ldg(q + o.v3);
      Real const sby = __ldg(q + o.b2), sbz = __ldg(q + o.b3);
      Real const spd = rdiv<true>(__ldg(q + o.p), e.dd, k);
      Real const A = e.beta_y * svy + e.beta_z * svz, B = e.beta_y * sby + e.beta_z * sbz;
      Real const qsA = e.qp_s * A, qfA = e.qp_f * A, asB = e.ap_s * B, afB = e.ap_f * B;
      Real const fx = e.cf * svx, sx = e.cs * svx;
      t0[i] = nf * (spd - fx) + qsA + asB;
      t2[i] = ns * (spd - sx) - qfA - afB;
      t4[i] = ns * (spd + sx) + qfA - afB;
      t6[i] = nf * (spd + fx) - qsA + asB;
    }
    // (inlined on purpose: the out-of-line form that pays off in the hydro sweeps -- CHL_PPM_CALL, hydro_physics.cuh -- costs
    // 17.2 -> 17.7 ms per Orszag-Tang 256^3 step here, gpurun_out/r16_run16.txt: at 128 registers the calls spill)
    ppm_single(t0[0], t0[1], t0[2], t0[3], t0[4], cl[0], cr[0], k);
    ppm_single(t6[0], t6[1], t6[2], t6[3], t6[4], cl[6], cr[6], k);
    ppm_single(t2[0], t2[1], t2[2], t2[3], t2[4], cl[2], cr[2], k);
    ppm_single(t4[0], t4[1], t4[2], t4[3], t4[4], cl[4], cr[4], k);
  }
  {  // Alfven pair: characteristics 1, 5
    Real t1[5], t5[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {
      const Real *q  = c + (long)(i - 2) * o.s;
      Real const svy = __ldg(q + o.v2), svz = __ldg(q + o.v3);
      Real const by = __ldg(q + o.b2) * e.sign * e.isd, bz = __ldg(q + o.b3) * e.sign * e.isd;
      t1[i] = RC(0.5) * (e.beta_y * (bz + svz) - e.beta_z * (by + svy));
      t5[i] = RC(0.5) * (e.beta_y * (bz - svz) - e.beta_z * (by - svy));
    }
    ppm_single(t1[0], t1[1], t1[2], t1[3], t1[4], cl[1], cr[1], k);
    ppm_single(t5[0], t5[1], t5[2], t5[3], t5[4], cl[5], cr[5], k);
  }
  {  // entropy: characteristic 3
    Real t3[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {
      const Real *q = c + (long)(i - 2) * o.s;
      t3[i]         = __ldg(q) - rdiv<true>(__ldg(q + o.p), e.da2, k);
    }
    ppm_single(t3[0], t3[1], t3[2], t3[3], t3[4], cl[3], cr[3], k);
  }
  L = e.unproject(cl);
  R = e.unproject(cr);
  mhd_floor(L, R);
}

template <class K>
__device__ __forceinline__ void mhd_plmc(const Real *__restrict__ c, MOff const &o, K const &k, MPrim &L, MPrim &R)
{
  MPrim const wi = load_mprim(c, o, 0), wm = load_mprim(c, o, -1), wp = load_mprim(c, o, 1);
  Eigen e;
  e.init(wi, k);
  Real const im[7] = {wm.d, wm.vx, wm.vy, wm.vz, wm.p, wm.by, wm.bz};
  Real const ii[7] = {wi.d, wi.vx, wi.vy, wi.vz, wi.p, wi.by, wi.bz};
  Real const ip[7] = {wp.d, wp.vx, wp.vy, wp.vz, wp.p, wp.by, wp.bz};
  Real dL[7], dR[7], dC[7], dG[7];
#pragma unroll
  for (int f = 0; f < 7; f++) {
    dL[f] = RC(1.0) * (ii[f] - im[f]);
    dR[f] = RC(1.0) * (ip[f] - ii[f]);
    dC[f] = RC(0.5) * (ip[f] - im[f]);
    dG[f] = vl_slope(dL[f], dR[f], k);
  }
  Real aL[7], aR[7], aC[7], aG[7], am[7];
  e.project(dL[0], dL[1], dL[2], dL[3], dL[4], dL[5], dL[6], aL, k);
  e.project(dR[0], dR[1], dR[2], dR[3], dR[4], dR[5], dR[6], aR, k);
  e.project(dC[0], dC[1], dC[2], dC[3], dC[4], dC[5], dC[6], aC, k);
  e.project(dG[0], dG[1], dG[2], dG[3], dG[4], dG[5], dG[6], aG, k);
#pragma unroll
  for (int f = 0; f < 7; f++) am[f] = mono_slope(aL[f], aR[f], aC[f], aG[f]);
  MPrim const dm = e.unproject(am);
  Real const dmv[7] = {dm.d, dm.vx, dm.vy, dm.vz, dm.p, dm.by, dm.bz};
  Real l[7], r[7];
#pragma unroll
  for (int f = 0; f < 7; f++) {
    l[f] = ii[f] + RC(1.0) * RC(0.5) * dmv[f];   // Calc_Interface_Linear: state + sign * 0.5 * slope
    r[f] = ii[f] + RC(-1.0) * RC(0.5) * dmv[f];
    clamp_pair(l[f], r[f], im[f], ii[f], ip[f]);
  }
  L = MPrim{l[0], l[1], l[2], l[3], l[4], RC(0.0), l[5], l[6]};
  R = MPrim{r[0], r[1], r[2], r[3], r[4], RC(0.0), r[5], r[6]};
  mhd_floor(L, R);
}

// Write_Data (reconstruction.h:918-968): interface primitives -> (d, m1, m2, m3, E, By, Bz); the energy uses the
// FACE-centred normal field of the face the state sits on
template <class K>
__device__ __forceinline__ M7 mhd_interface(MPrim const &q, Real bface, K const &k)
{
  M7 s;
  s.v[0] = q.d;
  s.v[1] = q.d * q.vx;
  s.v[2] = q.d * q.vy;
  s.v[3] = q.d * q.vz;
  s.v[4] = mhd_energy(q.p, q.d, q.vx, q.vy, q.vz, k, bface, q.by, q.bz);
  s.v[5] = q.by;
  s.v[6] = q.bz;
  return s;
}

// ---- HLLD ----------------------------------------------------------------------------------------------------------
struct HState {
  Real d, vx, vy, vz, E, by, bz, pg, pt;
};
struct HStar {
  Real d, vy, vz, E, by, bz;
};
template <class K>
__device__ __forceinline__ HState hlld_load(M7 const &q, Real bx, K const &k)
{
  Real const gamma = k.gamma;
  HState s;
  s.d  = max_s(q.v[0], kTiny);
  RDen const dd = rden(s.d, k);
  s.vx = rdiv(q.v[1], dd, k);
  s.vy = rdiv(q.v[2], dd, k);
  s.vz = rdiv(q.v[3], dd, k);
  s.E  = max_s(q.v[4], kTiny);
  s.by = q.v[5];
  s.bz = q.v[6];
  s.pg = mhd_pressure(s.E, s.d, s.vx, s.vy, s.vz, gamma, bx, s.by, s.bz);
  s.pt = total_pressure(s.pg, bx, s.by, s.bz);
  return s;
}
__device__ __forceinline__ M7 hlld_nonstar(HState const &s, Real bx)
{
  M7 f;
  f.v[0] = s.d * s.vx;
  f.v[1] = f.v[0] * s.vx + s.pt - bx * bx;
  f.v[2] = f.v[0] * s.vy - bx * s.by;
  f.v[3] = f.v[0] * s.vz - bx * s.bz;
  f.v[5] = s.by * s.vx - bx * s.vy;
  f.v[6] = s.bz * s.vx - bx * s.vz;
  f.v[4] = s.vx * (s.E + s.pt) - bx * (s.vx * bx + ((s.vy * s.by) + (s.vz * s.bz)));
  return f;
}
template <class K>
__device__ __forceinline__ HStar hlld_star(HState const &s, Real S, Real SM, Real bx, Real pts, K const &k)
{
  HStar o;
  RDen const dsm = rden(S - SM, k);
  o.d = rdiv<true>(s.d * (S - s.vx), dsm, k);
  if (fabs(s.d * (S - s.vx) * (S - SM) - (bx * bx)) < pts * kHlldSmall) {
    o.vy = s.vy;
    o.vz = s.vz;
    o.by = s.by;
    o.bz = s.bz;
  } else {
    RDen const dn = rden(s.d * (S - s.vx) * (S - SM) - (bx * bx), k);
    Real coef     = rdiv(bx * (SM - s.vx), dn, k);
    o.vy          = s.vy - s.by * coef;
    o.vz          = s.vz - s.bz * coef;
    Real t        = (S - s.vx);
    t             = t * t;
    coef          = rdiv(s.d * t - (bx * bx), dn, k);
    o.by          = s.by * coef;
    o.bz          = s.bz * coef;
  }
  o.E = rdiv<true>(s.E * (S - s.vx) - s.pt * s.vx + pts * SM +
                 bx * (dot3(s.vx, s.vy, s.vz, bx, s.by, s.bz) - dot3(SM, o.vy, o.vz, bx, o.by, o.bz)),
             dsm, k);
  return o;
}
__device__ __forceinline__ M7 hlld_star_flux(HStar const &st, HState const &s, M7 const &f, Real SM, Real S)
{
  M7 o;
  o.v[0] = f.v[0] + S * (st.d - s.d);
  o.v[1] = f.v[1] + S * (st.d * SM - s.d * s.vx);
  o.v[2] = f.v[2] + S * (st.d * st.vy - s.d * s.vy);
  o.v[3] = f.v[3] + S * (st.d * st.vz - s.d * s.vz);
  o.v[4] = f.v[4] + S * (st.E - s.E);
  o.v[5] = f.v[5] + S * (st.by - s.by);
  o.v[6] = f.v[6] + S * (st.bz - s.bz);
  return o;
}
__device__ __forceinline__ M7 hlld_dstar_flux(Real vyd, Real vzd, Real byd, Real bzd, Real Ed, HStar const &st,
                                              HState const &s, M7 const &f, Real SM, Real S, Real Sstar)
{
  M7 o;
  Real const sd = Sstar - S;
  o.v[0] = f.v[0] - S * s.d - sd * st.d + Sstar * st.d;
  o.v[1] = f.v[1] - S * (s.d * s.vx) - sd * (st.d * SM) + Sstar * (st.d * SM);
  o.v[2] = f.v[2] - S * (s.d * s.vy) - sd * (st.d * st.vy) + Sstar * (st.d * vyd);
  o.v[3] = f.v[3] - S * (s.d * s.vz) - sd * (st.d * st.vz) + Sstar * (st.d * vzd);
  o.v[4] = f.v[4] - S * s.E - sd * st.E + Sstar * Ed;
  o.v[5] = f.v[5] - S * s.by - sd * st.by + Sstar * byd;
  o.v[6] = f.v[6] - S * s.bz - sd * st.bz + Sstar * bzd;
  return o;
}

// Calculate_HLLD_Fluxes_CUDA for one interface.  Returns false when no branch of the reference fires (NaN input);
// the reference kernel then leaves its flux array untouched -- here the flux is zero.
template <class K>
__device__ __forceinline__ bool hlld_flux(M7 const &ql, M7 const &qr, Real bx, K const &k, M7 &f)
{
  HState const l = hlld_load(ql, bx, k), r = hlld_load(qr, bx, k);
  Real const cfl  = magnetosonic(l.d, l.pg, bx, l.by, l.bz, k, RC(1.0));
  Real const cfr  = magnetosonic(r.d, r.pg, bx, r.by, r.bz, k, RC(1.0));
  Real const cmax = fmax(cfl, cfr);
  Real const SL = fmin(l.vx, r.vx) - cmax, SR = fmax(l.vx, r.vx) + cmax;
  M7 const fl = hlld_nonstar(l, bx);
  if (SL > RC(0.0)) {
    f = fl;
    return true;
  }
  M7 const fr = hlld_nonstar(r, bx);
  if (SR < RC(0.0)) {
    f = fr;
    return true;
  }
  Real const rdiff = SR - r.vx, ldiff = SL - l.vx;
  // (rdiff * r.d - ldiff * l.d) and (r.d * (SR - r.vx) - l.d * (SL - l.vx)) are the same products: one denominator
  RDen const dsm = rden(rdiff * r.d - ldiff * l.d, k);
  Real const SM  = rdiv(rdiff * r.d * r.vx - ldiff * l.d * l.vx - r.pt + l.pt, dsm, k);
  Real const pts = rdiv<true>(r.d * l.pt * (SR - r.vx) - l.d * r.pt * (SL - l.vx) +
                            l.d * r.d * (SR - r.vx) * (SL - l.vx) * (r.vx - l.vx),
                        dsm, k);
  HStar const sl = hlld_star(l, SL, SM, bx, pts, k);
  Real const SLs = SM + RC(-1.0) * alfven_speed(bx, sl.d, k);
  if (SLs > RC(0.0) && SL <= RC(0.0)) {
    f = hlld_star_flux(sl, l, fl, SM, SL);
    return true;
  }
  HStar const sr = hlld_star(r, SR, SM, bx, pts, k);
  Real const SRs = SM + RC(1.0) * alfven_speed(bx, sr.d, k);
  if (SRs <= RC(0.0) && SR >= RC(0.0)) {
    f = hlld_star_flux(sr, r, fr, SM, SR);
    return true;
  }
  Real vyd, vzd, byd, bzd, EdL = RC(0.0), EdR = RC(0.0);
  if (RC(0.5) * (bx * bx) < kHlldSmall * pts) {
    if (SM >= RC(0.0)) {
      vyd = sl.vy; vzd = sl.vz; byd = sl.by; bzd = sl.bz; EdL = sl.E;
    } else {
      vyd = sr.vy; vzd = sr.vz; byd = sr.by; bzd = sr.bz; EdR = sr.E;
    }
  } else {
    Real const sdl = msq(sl.d, k), sdr = msq(sr.d, k);
    Real const inv = rdiv<true>(RC(1.0), rden(sdl + sdr, k), k);
    Real const sg  = copysign(RC(1.0), bx);
    vyd = inv * (sdl * sl.vy + sdr * sr.vy + sg * (sr.by - sl.by));
    vzd = inv * (sdl * sl.vz + sdr * sr.vz + sg * (sr.bz - sl.bz));
    byd = inv * (sdl * sr.by + sdr * sl.by + sg * (sdl * sdr) * (sr.vy - sl.vy));
    bzd = inv * (sdl * sr.bz + sdr * sl.bz + sg * (sdl * sdr) * (sr.vz - sl.vz));
    Real const vdb = dot3(SM, vyd, vzd, bx, byd, bzd);
    EdL = sl.E - sdl * sg * (dot3(SM, sl.vy, sl.vz, bx, sl.by, sl.bz) - vdb);
    EdR = sr.E + sdr * sg * (dot3(SM, sr.vy, sr.vz, bx, sr.by, sr.bz) - vdb);
  }
  if (SM > RC(0.0) && SLs <= RC(0.0)) {
    f = hlld_dstar_flux(vyd, vzd, byd, bzd, EdL, sl, l, fl, SM, SL, SLs);
    return true;
  }
  if (SRs > RC(0.0) && SM <= RC(0.0)) {
    f = hlld_dstar_flux(vyd, vzd, byd, bzd, EdR, sr, r, fr, SM, SR, SRs);
    return true;
  }
#pragma unroll
  for (int k = 0; k < 7; k++) f.v[k] = RC(0.0);
  return false;
}

// mhdInverseCrossingTime (hydro_cuda.cu:420-435)
__device__ __forceinline__ Real mhd_inverse_crossing_time(Real E, Real d, Real vx, Real vy, Real vz, Real bx, Real by,
                                                          Real bz, Real dx, Real dy, Real dz, Real gamma)
{
  KC const kb   = make_kc(gamma);  // once per cell and step: the per-quotient fallback branch is fine here
  Real const p  = mhd_pressure(E, d, vx, vy, vz, gamma, bx, by, bz);
  Real const cf = magnetosonic(d, p, bx, by, bz, kb, RC(1.0));
  Real m        = fmax((fabs(vx) + cf) / dx, (fabs(vy) + cf) / dy);
  m             = fmax(m, (fabs(vz) + cf) / dz);
  m             = fmax(m, RC(0.0));
  return m;
}

}  // namespace CHL_NS
