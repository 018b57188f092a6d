/*
 * vl_oracle.c -- CPU restatement of Cholla's Van Leer hydro hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the shipped product (cholla_b200/, the C-ABI
 * library) links, imports or calls this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may.  It exists to be the checker that
 * the hand-written sm_100a kernels are compared against.
 *
 * It restates, stage by stage and in the reference's floating-point operation order, the
 * algorithm of (all paths relative to the reference tree, commit 8478c9e):
 *   src/integrators/VL_{1,2,3}D_cuda.cu      six-stage predictor/corrector driver
 *   src/reconstruction/pcm_cuda.cu           piecewise-constant interfaces
 *   src/reconstruction/plmc_cuda.cu          PLM, characteristic limiting
 *   src/reconstruction/ppmc_cuda.cu:540-695  PPMC_VL
 *   src/reconstruction/reconstruction.h      shared helpers (hydro branches)
 *   src/riemann_solvers/hllc_cuda.cu, roe_cuda.cu, hll_cuda.cu;  src/reconstruction/plmp_cuda.cu, ppmp_cuda.cu (-DVL branches)
 *   src/hydro/hydro_cuda.cu                  conserved update, Calc_dt, temperature floor
 *   src/grid/cuda_boundaries.cu, src/grid/boundary_conditions.cpp:419-  ghost cells
 * Build with -O2 -ffp-contract=off (the reference device code is built -fmad=false,
 * Makefile:148), so every + and * rounds exactly as the reference's does.
 *
 * Parity pin: checked against the reference's own unit-test fiducials (the JSON fixtures under tests/golden,
 * extracted from plmc_cuda_tests.cu, ppmc_cuda_tests.cu, reconstruction_tests.cu,
 * hydro_cuda_tests.cu, hllc_cuda_tests.cu) by tests/test_oracle_golden.py, and against the
 * reference's own kernels compiled for sm_100 (oracle/_ref) by tests/test_gpu_reference.py.
 *
 * Scope: hydro (5 fields; density, momentum xyz, Energy), fp64.  Layout is the reference's:
 * SoA, field stride N = nx*ny*nz, cell id = x + nx*(y + ny*z), ghosts included.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_TINY 1.0e-20 /* global.h:25 */

enum { ORC_PCM = 0, ORC_PLMC = 1, ORC_PPMC = 2, ORC_PLMP = 3, ORC_PPMP = 4 };
enum { ORC_HLLC = 0, ORC_ROE = 1, ORC_HLL = 3, ORC_EXACT = 4 }; /* 2 is HLLD (mhd_oracle.c) */

typedef struct {
  double d, vx, vy, vz, p;
} prim_t;

/* global.h:144 -- note SIGN(0) = +1 */
static inline double sign_macro(double a) { return (a < 0.) ? -1. : 1.; }

/* math_utilities.h:78-98: a1*b1 + ((a2*b2) + (a3*b3)) */
static inline double square_mag(double a, double b, double c) { return a * a + ((b * b) + (c * c)); }

static inline void dir_perm(int dir, int *o1, int *o2, int *o3)
{ /* plmc_cuda.cu:41-58 */
  if (dir == 0) { *o1 = 1; *o2 = 2; *o3 = 3; }
  else if (dir == 1) { *o1 = 2; *o2 = 3; *o3 = 1; }
  else { *o1 = 3; *o2 = 1; *o3 = 2; }
}

/* reconstruction.h:139-212 (hydro, no DE) + hydro_utilities.h:34-45 */
static prim_t load_prim(const double *U, size_t id, size_t N, int o1, int o2, int o3, double gamma)
{
  prim_t w;
  w.d  = U[id];
  w.vx = U[(size_t)o1 * N + id] / w.d;
  w.vy = U[(size_t)o2 * N + id] / w.d;
  w.vz = U[(size_t)o3 * N + id] / w.d;
  double pr = U[4 * N + id] - 0.5 * w.d * square_mag(w.vx, w.vy, w.vz);
  w.p = fmax((gamma - 1.) * pr, ORC_TINY);
  return w;
}

/* reconstruction.h:918-968 + hydro_utilities.h:106-118 */
static void store_interface(prim_t q, double *Q, size_t id, size_t N, int o1, int o2, int o3, double gamma)
{
  Q[id]                  = q.d;
  Q[(size_t)o1 * N + id] = q.d * q.vx;
  Q[(size_t)o2 * N + id] = q.d * q.vy;
  Q[(size_t)o3 * N + id] = q.d * q.vz;
  Q[4 * N + id]          = (fmax(q.p, ORC_TINY) / (gamma - 1.)) + 0.5 * q.d * square_mag(q.vx, q.vy, q.vz);
}

/* reconstruction.h:95-119 */
static int outside_guard(int order, int nx, int ny, int nz, int x, int y, int z)
{
  int out = x < order - 1 || x >= nx - order;
  if (ny > 1) out = y < order - 1 || y >= ny - order || out;
  if (nz > 1) out = z < order - 1 || z >= nz - order || out;
  else out = z >= nz || out;
  return out;
}

/* ---- characteristic projections, hydro (reconstruction.h:423-431, 482-488) ---- */
typedef struct { double a0, a1, a2, a3, a4; } char_t;

static char_t to_char(prim_t c, prim_t s, double a, double a2)
{
  char_t o;
  o.a0 = -c.d * s.vx / (2.0 * a) + s.p / (2.0 * a2);
  o.a1 = s.d - s.p / (a2);
  o.a2 = s.vy;
  o.a3 = s.vz;
  o.a4 = c.d * s.vx / (2.0 * a) + s.p / (2.0 * a2);
  return o;
}

static prim_t from_char(prim_t c, char_t s, double a, double a2)
{
  prim_t o;
  o.d  = s.a0 + s.a1 + s.a4;
  o.vx = a / c.d * (s.a4 - s.a0);
  o.vy = s.a2;
  o.vz = s.a3;
  o.p  = a2 * (s.a0 + s.a4);
  return o;
}

/* ---- PLMC pieces (reconstruction.h:224-296, 512-559, 645-723) ---- */
static inline double vl_slope(double l, double r) { return (l * r > 0.0) ? 2.0 * l * r / (l + r) : 0.0; }

static inline double mono(double l, double r, double c, double g)
{
  if (l * r > 0.0) {
    double la = 2.0 * fmin(fabs(l), fabs(r));
    double lb = fmin(fabs(c), fabs(g));
    return copysign(fmin(la, lb), c);
  }
  return 0.0;
}

static inline void clamp_pair(double *l_iph, double *r_imh, double vm, double vi, double vp)
{
  *r_imh = fmax(fmin(vi, vm), *r_imh);
  *r_imh = fmin(fmax(vi, vm), *r_imh);
  *l_iph = fmax(fmin(vi, vp), *l_iph);
  *l_iph = fmin(fmax(vi, vp), *l_iph);
}

static prim_t prim_diff(prim_t l, prim_t r, double coef)
{
  prim_t s;
  s.d = coef * (r.d - l.d); s.vx = coef * (r.vx - l.vx); s.vy = coef * (r.vy - l.vy);
  s.vz = coef * (r.vz - l.vz); s.p = coef * (r.p - l.p);
  return s;
}

/* ---- test hooks: the building blocks above as the reference's helper-level unit tests call them
 * (reconstruction_tests.cu:177-680, hydro_utilities_tests.cpp:225-236, math_utilities_tests.cpp:45-76); read by
 * tests/test_oracle_helpers_golden.py only ---- */
static prim_t prim_of(const double *v) { prim_t w = {v[0], v[1], v[2], v[3], v[4]}; return w; }
static void prim_out(prim_t w, double *o) { o[0] = w.d; o[1] = w.vx; o[2] = w.vy; o[3] = w.vz; o[4] = w.p; }
void orc_test_load_data(const double *U, long id, long N, int o1, int o2, int o3, double gamma, double *out5)
{ prim_out(load_prim(U, (size_t)id, (size_t)N, o1, o2, o3, gamma), out5); }
void orc_test_write_data(const double *prim5, double *Q, long id, long N, int o1, int o2, int o3, double gamma)
{ store_interface(prim_of(prim5), Q, (size_t)id, (size_t)N, o1, o2, o3, gamma); }
int orc_test_thread_guard(int order, int nx, int ny, int nz, int x, int y, int z)
{ return outside_guard(order, nx, ny, nz, x, y, z); }
void orc_test_compute_slope(const double *l5, const double *r5, double coef, double *out5)
{ prim_out(prim_diff(prim_of(l5), prim_of(r5), coef), out5); }
void orc_test_van_leer_slope(const double *l5, const double *r5, double *out5)
{ for (int i = 0; i < 5; i++) out5[i] = vl_slope(l5[i], r5[i]); }
void orc_test_monotonize_characteristic(const double *prim5, const double *aL, const double *aR, const double *aC,
                                        const double *aG, double a, double a2, double *out5)
{
  char_t am = {mono(aL[0], aR[0], aC[0], aG[0]), mono(aL[1], aR[1], aC[1], aG[1]), mono(aL[2], aR[2], aC[2], aG[2]),
               mono(aL[3], aR[3], aC[3], aG[3]), mono(aL[4], aR[4], aC[4], aG[4])};
  prim_out(from_char(prim_of(prim5), am, a, a2), out5);
}
void orc_test_plm_limit_interfaces(double *l_iph5, double *r_imh5, const double *vm5, const double *vi5, const double *vp5)
{ for (int i = 0; i < 5; i++) clamp_pair(&l_iph5[i], &r_imh5[i], vm5[i], vi5[i], vp5[i]); }
double orc_test_square_magnitude(double a, double b, double c) { return square_mag(a, b, c); }

/* one cell of PLMC_cuda (plmc_cuda.cu:60-128, 287-299).  trace = 0: a -DVL build, where the #ifndef VL block is dead;
 * trace = 1: a -DSIMPLE build, which evolves the interface values over dt/2 along the characteristics
 * (plmc_cuda.cu:130-285) with dtodx = dt / dx of the sweep direction. */
static void plmc_cell_t(prim_t wm, prim_t wi, prim_t wp, double gamma, int trace, double dtodx, prim_t *L, prim_t *R)
{
  double a  = sqrt(gamma * wi.p / wi.d);
  double a2 = a * a;
  prim_t dL = prim_diff(wm, wi, 1.0), dR = prim_diff(wi, wp, 1.0), dC = prim_diff(wm, wp, 0.5);
  prim_t dG;
  dG.d = vl_slope(dL.d, dR.d); dG.vx = vl_slope(dL.vx, dR.vx); dG.vy = vl_slope(dL.vy, dR.vy);
  dG.vz = vl_slope(dL.vz, dR.vz); dG.p = vl_slope(dL.p, dR.p);
  char_t aL = to_char(wi, dL, a, a2), aR = to_char(wi, dR, a, a2), aC = to_char(wi, dC, a, a2),
         aG = to_char(wi, dG, a, a2);
  char_t am;
  am.a0 = mono(aL.a0, aR.a0, aC.a0, aG.a0);
  am.a1 = mono(aL.a1, aR.a1, aC.a1, aG.a1);
  am.a2 = mono(aL.a2, aR.a2, aC.a2, aG.a2);
  am.a3 = mono(aL.a3, aR.a3, aC.a3, aG.a3);
  am.a4 = mono(aL.a4, aR.a4, aC.a4, aG.a4);
  prim_t dm = from_char(wi, am, a, a2);
  prim_t l, r;
  l.d = wi.d + 1.0 * 0.5 * dm.d;    r.d = wi.d + -1.0 * 0.5 * dm.d;
  l.vx = wi.vx + 1.0 * 0.5 * dm.vx; r.vx = wi.vx + -1.0 * 0.5 * dm.vx;
  l.vy = wi.vy + 1.0 * 0.5 * dm.vy; r.vy = wi.vy + -1.0 * 0.5 * dm.vy;
  l.vz = wi.vz + 1.0 * 0.5 * dm.vz; r.vz = wi.vz + -1.0 * 0.5 * dm.vz;
  l.p = wi.p + 1.0 * 0.5 * dm.p;    r.p = wi.p + -1.0 * 0.5 * dm.p;
  clamp_pair(&l.d, &r.d, wm.d, wi.d, wp.d);
  clamp_pair(&l.vx, &r.vx, wm.vx, wi.vx, wp.vx);
  clamp_pair(&l.vy, &r.vy, wm.vy, wi.vy, wp.vy);
  clamp_pair(&l.vz, &r.vz, wm.vz, wi.vz, wp.vz);
  clamp_pair(&l.p, &r.p, wm.p, wi.p, wp.p);
  if (trace) {
    double lm = wi.vx - a, l0 = wi.vx, lp = wi.vx + a;
    /* integrate the linear profile over the domain of dependence of the fastest wave (plmc_cuda.cu:140-156) */
    double qx = -0.5 * fmin(lm, 0.0) * dtodx;
    r.d = r.d + qx * dm.d; r.vx = r.vx + qx * dm.vx; r.vy = r.vy + qx * dm.vy; r.vz = r.vz + qx * dm.vz;
    r.p = r.p + qx * dm.p;
    qx = 0.5 * fmax(lp, 0.0) * dtodx;
    l.d = l.d - qx * dm.d; l.vx = l.vx - qx * dm.vx; l.vy = l.vy - qx * dm.vy; l.vz = l.vz - qx * dm.vz;
    l.p = l.p - qx * dm.p;
    /* characteristic tracing, Stone eqs. 42-43 (plmc_cuda.cu:170-283).  The reference also adds the terms whose
     * wave-speed difference is identically zero (lambda_p - lambda_p, lambda_m - lambda_m): +0 for finite data */
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;
    if (lm >= 0) {
      double ld = lp - lm;
      s0 += ld * (-wi.d * dm.vx / (2 * a) + dm.p / (2 * a2));
      s1 += ld * (dm.vx / 2.0 - dm.p / (2 * a * wi.d));
      s4 += ld * (-wi.d * dm.vx * a / 2.0 + dm.p / 2.0);
    }
    if (l0 >= 0) {
      double ld = lp - l0;
      s0 += ld * (dm.d - dm.p / (a2));
      s2 += ld * dm.vy;
      s3 += ld * dm.vz;
    }
    if (lp >= 0) {
      double ld = lp - lp;
      s0 += ld * (wi.d * dm.vx / (2 * a) + dm.p / (2 * a2));
      s1 += ld * (dm.vx / 2.0 + dm.p / (2 * a * wi.d));
      s4 += ld * (wi.d * dm.vx * a / 2.0 + dm.p / 2.0);
    }
    l.d += 0.5 * dtodx * s0; l.vx += 0.5 * dtodx * s1; l.vy += 0.5 * dtodx * s2; l.vz += 0.5 * dtodx * s3;
    l.p += 0.5 * dtodx * s4;
    s0 = s1 = s2 = s3 = s4 = 0;
    if (lm <= 0) {
      double ld = lm - lm;
      s0 += ld * (-wi.d * dm.vx / (2 * a) + dm.p / (2 * a2));
      s1 += ld * (dm.vx / 2.0 - dm.p / (2 * a * wi.d));
      s4 += ld * (-wi.d * dm.vx * a / 2.0 + dm.p / 2.0);
    }
    if (l0 <= 0) {
      double ld = lm - l0;
      s0 += ld * (dm.d - dm.p / (a2));
      s2 += ld * dm.vy;
      s3 += ld * dm.vz;
    }
    if (lp <= 0) {
      double ld = lm - lp;
      s0 += ld * (wi.d * dm.vx / (2 * a) + dm.p / (2 * a2));
      s1 += ld * (dm.vx / 2.0 + dm.p / (2 * a * wi.d));
      s4 += ld * (wi.d * dm.vx * a / 2.0 + dm.p / 2.0);
    }
    r.d += 0.5 * dtodx * s0; r.vx += 0.5 * dtodx * s1; r.vy += 0.5 * dtodx * s2; r.vz += 0.5 * dtodx * s3;
    r.p += 0.5 * dtodx * s4;
  }
  r.d = fmax(r.d, ORC_TINY); l.d = fmax(l.d, ORC_TINY);
  r.p = fmax(r.p, ORC_TINY); l.p = fmax(l.p, ORC_TINY);
  *L = l; *R = r;
}
/* exposed for the tests: one PLMC cell with the SIMPLE build's tracing; w* = (d, vx, vy, vz, p) */
void orc_plmc_trace_cell(const double *wm, const double *wi, const double *wp, double gamma, double dtodx, double *L,
                         double *R)
{
  prim_t m = {wm[0], wm[1], wm[2], wm[3], wm[4]}, i = {wi[0], wi[1], wi[2], wi[3], wi[4]},
         p = {wp[0], wp[1], wp[2], wp[3], wp[4]}, l, r;
  plmc_cell_t(m, i, p, gamma, 1, dtodx, &l, &r);
  L[0] = l.d; L[1] = l.vx; L[2] = l.vy; L[3] = l.vz; L[4] = l.p;
  R[0] = r.d; R[1] = r.vx; R[2] = r.vy; R[3] = r.vz; R[4] = r.p;
}

/* ---- PPM_Single_Variable (reconstruction.h:794-901) ---- */
static double ppm_limit_face(double ci, double cim1, double face, double s2_i, double s2_im1)
{
  const double C2 = 1.25;
  double s2_c = 3.0 * (cim1 + ci - 2.0 * face);
  double lim = 0.0;
  if (sign_macro(s2_c) == sign_macro(s2_im1) && sign_macro(s2_c) == sign_macro(s2_i)) {
    lim = sign_macro(s2_c) * fmin(C2 * fabs(s2_im1), fmin(C2 * fabs(s2_i), fabs(s2_c)));
  }
  double dl = face - cim1, dr = ci - face;
  if (dl * dr < 0.0) return 0.5 * (cim1 + ci) - lim / 6.0;
  return face;
}

void orc_ppm_single_variable(double cm2, double cm1, double c0, double cp1, double cp2, double *L_iph, double *R_imh)
{
  const double C2 = 1.25;
  double sl = (c0 - cm1), sr = (cp1 - c0);
  double avg_m1 = 0.5 * sl + 0.5 * (cm1 - cm2);
  double avg_0  = 0.5 * sr + 0.5 * sl;
  double avg_p1 = 0.5 * (cp2 - cp1) + 0.5 * sr;
  double R = 0.5 * (cm1 + c0) + (avg_m1 - avg_0) / 6.0;
  double L = 0.5 * (c0 + cp1) + (avg_0 - avg_p1) / 6.0;
  double s2_m1 = cm2 + c0 - 2.0 * cm1;
  double s2_0  = cm1 + cp1 - 2.0 * c0;
  double s2_p1 = c0 + cp2 - 2.0 * cp1;
  R = ppm_limit_face(c0, cm1, R, s2_0, s2_m1);
  L = ppm_limit_face(cp1, c0, L, s2_p1, s2_0);
  double s2_f = 6.0 * (R + L - 2.0 * c0);
  double s2_lim = 0.0;
  if (sign_macro(s2_m1) == sign_macro(s2_0) && sign_macro(s2_m1) == sign_macro(s2_p1) &&
      sign_macro(s2_m1) == sign_macro(s2_f)) {
    s2_lim = sign_macro(s2_f) * fmin(fmin(C2 * fabs(s2_m1), C2 * fabs(s2_0)), fmin(C2 * fabs(s2_p1), fabs(s2_f)));
  }
  double cmax = fmax(fabs(cm2), fabs(cm1));
  cmax = fmax(cmax, fabs(c0));
  cmax = fmax(cmax, fabs(cp1));
  cmax = fmax(cmax, fabs(cp2));
  double rho = (fabs(s2_f) > (1.0e-12) * cmax) ? s2_lim / s2_f : 0.0;
  double fl = c0 - R, fr = L - c0;
  if ((fl * fr) <= 0.0 || ((cp1 - c0) * (c0 - cm1)) <= 0.0) {
    if (rho <= (1.0 - (1.0e-12))) {
      R = c0 - rho * fl;
      L = c0 + rho * fr;
    }
  } else {
    if (fabs(fl) >= 2.0 * fabs(fr)) R = c0 - 2.0 * fr;
    if (fabs(fr) >= 2.0 * fabs(fl)) L = c0 + 2.0 * fl;
  }
  *L_iph = L; *R_imh = R;
}

/* one cell of PPMC_VL (ppmc_cuda.cu:576-684, hydro) */
static void ppmc_cell(prim_t wm2, prim_t wm1, prim_t wi, prim_t wp1, prim_t wp2, double gamma, prim_t *L, prim_t *R)
{
  double a  = sqrt(gamma * wi.p / wi.d);
  double a2 = a * a;
  char_t c0 = to_char(wi, wi, a, a2), cm1 = to_char(wi, wm1, a, a2), cm2 = to_char(wi, wm2, a, a2),
         cp1 = to_char(wi, wp1, a, a2), cp2 = to_char(wi, wp2, a, a2);
  char_t cl, cr;
  orc_ppm_single_variable(cm2.a0, cm1.a0, c0.a0, cp1.a0, cp2.a0, &cl.a0, &cr.a0);
  orc_ppm_single_variable(cm2.a1, cm1.a1, c0.a1, cp1.a1, cp2.a1, &cl.a1, &cr.a1);
  orc_ppm_single_variable(cm2.a2, cm1.a2, c0.a2, cp1.a2, cp2.a2, &cl.a2, &cr.a2);
  orc_ppm_single_variable(cm2.a3, cm1.a3, c0.a3, cp1.a3, cp2.a3, &cl.a3, &cr.a3);
  orc_ppm_single_variable(cm2.a4, cm1.a4, c0.a4, cp1.a4, cp2.a4, &cl.a4, &cr.a4);
  prim_t l = from_char(wi, cl, a, a2), r = from_char(wi, cr, a, a2);
  r.d = fmax(r.d, ORC_TINY); l.d = fmax(l.d, ORC_TINY);
  r.p = fmax(r.p, ORC_TINY); l.p = fmax(l.p, ORC_TINY);
  *L = l; *R = r;
}

/* ---- PLMP under -DVL (plmp_cuda.cu:21-310): van Leer-limited slopes of the PRIMITIVE variables ---- */
/* Interface_Values_PLM (plmp_cuda.cu:312-365); sgn_CUDA(0) = +1 (global_cuda.h:41-48) */
static void plmp_interface(double qm, double qi, double qp, double *qL, double *qR)
{
  double dL = qi - qm, dR = qp - qi, dC = 0.5 * (qp - qm), dG;
  if (dL * dR > 0.0) dG = 2.0 * dL * dR / (dL + dR);
  else dG = 0.0;
  double a = fmin(fabs(dL), fabs(dR)), b = fmin(fabs(dC), fabs(dG));
  double m = ((dC < 0) ? -1 : 1) * fmin(2.0 * a, b);
  *qL = qi - 0.5 * m;
  *qR = qi + 0.5 * m;
}
/* the kernel's own primitive load: plain left-to-right kinetic-energy sum (plmp_cuda.cu:107-118) */
static prim_t plmp_load(const double *U, size_t id, size_t N, int o1, int o2, int o3, double gamma)
{
  prim_t w;
  w.d  = U[id];
  w.vx = U[(size_t)o1 * N + id] / w.d;
  w.vy = U[(size_t)o2 * N + id] / w.d;
  w.vz = U[(size_t)o3 * N + id] / w.d;
  w.p  = (U[4 * N + id] - 0.5 * w.d * (w.vx * w.vx + w.vy * w.vy + w.vz * w.vz)) * (gamma - 1.0);
  w.p  = fmax(w.p, ORC_TINY);
  return w;
}
static void plmp_store(prim_t q, double *Q, size_t id, size_t N, int o1, int o2, int o3, double gamma)
{
  Q[id]                  = q.d;
  Q[(size_t)o1 * N + id] = q.d * q.vx;
  Q[(size_t)o2 * N + id] = q.d * q.vy;
  Q[(size_t)o3 * N + id] = q.d * q.vz;
  Q[4 * N + id]          = q.p / (gamma - 1.0) + 0.5 * q.d * (q.vx * q.vx + q.vy * q.vy + q.vz * q.vz);
}
static void plmp_dir(const double *U, double *QL, double *QR, int nx, int ny, int nz, double gamma, int dir)
{
  size_t N = (size_t)nx * ny * nz;
  int o1, o2, o3;
  dir_perm(dir, &o1, &o2, &o3);
  long stride = dir == 0 ? 1 : (dir == 1 ? nx : (long)nx * ny);
  int n[3] = {nx, ny, nz};
#pragma omp parallel for collapse(2) schedule(static)
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++) {
        int c = dir == 0 ? x : (dir == 1 ? y : z);
        if (c < 1 || c >= n[dir] - 2) continue; /* xs = 1, xe = nx - 2 along the sweep, every transverse cell */
        size_t id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
        prim_t wi = plmp_load(U, id, N, o1, o2, o3, gamma);
        prim_t wm = plmp_load(U, id - stride, N, o1, o2, o3, gamma);
        prim_t wp = plmp_load(U, id + stride, N, o1, o2, o3, gamma);
        prim_t lo, hi; /* lo: the cell's low face (i-1/2), hi: its high face (i+1/2) */
        plmp_interface(wm.d, wi.d, wp.d, &lo.d, &hi.d);
        plmp_interface(wm.vx, wi.vx, wp.vx, &lo.vx, &hi.vx);
        plmp_interface(wm.vy, wi.vy, wp.vy, &lo.vy, &hi.vy);
        plmp_interface(wm.vz, wi.vz, wp.vz, &lo.vz, &hi.vz);
        plmp_interface(wm.p, wi.p, wp.p, &lo.p, &hi.p);
        lo.d = fmax(lo.d, ORC_TINY); hi.d = fmax(hi.d, ORC_TINY);
        lo.p = fmax(lo.p, ORC_TINY); hi.p = fmax(hi.p, ORC_TINY);
        plmp_store(lo, QR, id - stride, N, o1, o2, o3, gamma); /* bounds_R[i-1]: right state of face i-1/2 */
        plmp_store(hi, QL, id, N, o1, o2, o3, gamma);          /* bounds_L[i]:   left state of face i+1/2 */
      }
}

/* ---- PPMP under -DVL without STEEPENING / FLATTENING (ppmp_cuda.cu:28-760; no build type defines them) ---- */
/* Calculate_Slope (ppmp_cuda.cu:677-712): the same van Leer-limited slope as Interface_Values_PLM */
static double ppmp_slope(double qm, double qi, double qp)
{
  double dL = qi - qm, dR = qp - qi, dC = 0.5 * (qp - qm), dG;
  if (dL * dR > 0.0) dG = 2.0 * dL * dR / (dL + dR);
  else dG = 0.0;
  double a = fmin(fabs(dL), fabs(dR)), b = fmin(fabs(dC), fabs(dG));
  return ((dC < 0) ? -1 : 1) * fmin(2.0 * a, b);
}
/* Interface_Values_PPM (ppmp_cuda.cu:719-745) */
static void ppmp_interface(double qm, double qi, double qp, double sm, double si, double sp, double *qL, double *qR)
{
  *qL = 0.5 * (qi + qm) - (1.0 / 6.0) * (si - sm);
  *qR = 0.5 * (qp + qi) - (1.0 / 6.0) * (sp - si);
  if ((*qR - qi) * (qi - *qL) <= 0) *qL = *qR = qi;
  if (6.0 * (*qR - *qL) * (qi - 0.5 * (*qL + *qR)) > (*qR - *qL) * (*qR - *qL)) *qL = 3.0 * qi - 2.0 * (*qR);
  if (6.0 * (*qR - *qL) * (qi - 0.5 * (*qL + *qR)) < -(*qR - *qL) * (*qR - *qL)) *qR = 3.0 * qi - 2.0 * (*qL);
  *qL = fmax(fmin(qi, qm), *qL);
  *qL = fmin(fmax(qi, qm), *qL);
  *qR = fmax(fmin(qi, qp), *qR);
  *qR = fmin(fmax(qi, qp), *qR);
}
static void ppmp_var(double q2m, double qm, double qi, double qp, double q2p, double *lo, double *hi)
{
  double sm = ppmp_slope(q2m, qm, qi), si = ppmp_slope(qm, qi, qp), sp = ppmp_slope(qi, qp, q2p);
  ppmp_interface(qm, qi, qp, sm, si, sp, lo, hi);
}
static void ppmp_dir(const double *U, double *QL, double *QR, int nx, int ny, int nz, double gamma, int dir)
{
  size_t N = (size_t)nx * ny * nz;
  int o1, o2, o3;
  dir_perm(dir, &o1, &o2, &o3);
  long stride = dir == 0 ? 1 : (dir == 1 ? nx : (long)nx * ny);
  int n[3] = {nx, ny, nz};
#pragma omp parallel for collapse(2) schedule(static)
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++) {
        int c = dir == 0 ? x : (dir == 1 ? y : z);
        if (c < 2 || c >= n[dir] - 3) continue; /* xs = 2, xe = nx - 3 along the sweep, every transverse cell */
        size_t id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
        prim_t w[5];
        for (int k = 0; k < 5; k++) w[k] = plmp_load(U, id + (k - 2) * stride, N, o1, o2, o3, gamma); /* same load */
        prim_t lo, hi;
        ppmp_var(w[0].d, w[1].d, w[2].d, w[3].d, w[4].d, &lo.d, &hi.d);
        ppmp_var(w[0].vx, w[1].vx, w[2].vx, w[3].vx, w[4].vx, &lo.vx, &hi.vx);
        ppmp_var(w[0].vy, w[1].vy, w[2].vy, w[3].vy, w[4].vy, &lo.vy, &hi.vy);
        ppmp_var(w[0].vz, w[1].vz, w[2].vz, w[3].vz, w[4].vz, &lo.vz, &hi.vz);
        ppmp_var(w[0].p, w[1].p, w[2].p, w[3].p, w[4].p, &lo.p, &hi.p);
        lo.d = fmax(lo.d, ORC_TINY); hi.d = fmax(hi.d, ORC_TINY);
        lo.p = fmax(lo.p, ORC_TINY); hi.p = fmax(hi.p, ORC_TINY);
        plmp_store(lo, QR, id - stride, N, o1, o2, o3, gamma);
        plmp_store(hi, QL, id, N, o1, o2, o3, gamma);
      }
}

/* Full-grid reconstruction pass in one direction: what one launch of PLMC_cuda / PPMC_VL does.
 * Q_L[id] = left state at face id+1/2 (from cell id); Q_R[id - stride] = right state at
 * face id-1/2 (from cell id).  Cells outside the guard are left untouched. */
static void reconstruct_t(int recon, const double *U, double *QL, double *QR, int nx, int ny, int nz, double gamma,
                          int dir, int trace, double dtodx);
void orc_reconstruct(int recon, const double *U, double *QL, double *QR, int nx, int ny, int nz, double gamma, int dir)
{
  reconstruct_t(recon, U, QL, QR, nx, ny, nz, gamma, dir, 0, 0.0);
}
static void reconstruct_t(int recon, const double *U, double *QL, double *QR, int nx, int ny, int nz, double gamma,
                          int dir, int trace, double dtodx)
{
  size_t N = (size_t)nx * ny * nz;
  int o1, o2, o3;
  dir_perm(dir, &o1, &o2, &o3);
  long stride = dir == 0 ? 1 : (dir == 1 ? nx : (long)nx * ny);
  int order = recon == ORC_PPMC ? 3 : 2;
#pragma omp parallel for collapse(2) schedule(static)
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++) {
        if (outside_guard(order, nx, ny, nz, x, y, z)) continue;
        size_t id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
        prim_t wi = load_prim(U, id, N, o1, o2, o3, gamma);
        prim_t wm = load_prim(U, id - stride, N, o1, o2, o3, gamma);
        prim_t wp = load_prim(U, id + stride, N, o1, o2, o3, gamma);
        prim_t L, R;
        if (recon == ORC_PPMC) {
          prim_t wm2 = load_prim(U, id - 2 * stride, N, o1, o2, o3, gamma);
          prim_t wp2 = load_prim(U, id + 2 * stride, N, o1, o2, o3, gamma);
          ppmc_cell(wm2, wm, wi, wp, wp2, gamma, &L, &R);
        } else {
          plmc_cell_t(wm, wi, wp, gamma, trace, dtodx, &L, &R);
        }
        store_interface(L, QL, id, N, o1, o2, o3, gamma);
        store_interface(R, QR, id - stride, N, o1, o2, o3, gamma);
      }
}

/* PCM (pcm_cuda.cu:13-424): L[id] = U[id], R[id - stride] = U[id] */
static void pcm_dir(const double *U, double *QL, double *QR, int nx, int ny, int nz, int dir)
{
  size_t N = (size_t)nx * ny * nz;
  long stride = dir == 0 ? 1 : (dir == 1 ? nx : (long)nx * ny);
  for (int f = 0; f < 5; f++) {
#pragma omp parallel for schedule(static)
    for (long id = 0; id < (long)N; id++) {
      QL[f * N + id] = U[f * N + id];
      int x = id % nx, y = (id / nx) % ny, z = id / ((long)nx * ny);
      int c = dir == 0 ? x : (dir == 1 ? y : z);
      if (c > 0) QR[f * N + id - stride] = U[f * N + id];
    }
  }
}

/* ---- Riemann solvers: one interface ---- */
typedef struct { double d, mx, my, mz, E; } cons_t;

/* Passive fields of a -DSCALAR / -DDE build riding along one interface: np conserved values per side (the scalars,
 * then GasEnergy last when de != 0: grid_enum.h:33-78) and their fluxes.  NULL = a plain hydro build. */
typedef struct {
  int np, de;
  const double *ql, *qr;
  double *fq;
} pass_t;
#define ORC_DE_ETA_1 0.001 /* global.h:65-68 (non-cosmology) */
#define ORC_DE_ETA_2 0.035 /* global.h:71-73 */
/* hydro_utilities.h:136-151 Get_Pressure_From_DE */
static inline double de_pressure(double E, double U_total, double U_advected, double gamma)
{
  double U = (U_total / E > ORC_DE_ETA_1) ? U_total : U_advected;
  return fmax(U * (gamma - 1.0), ORC_TINY);
}
double orc_test_de_pressure(double E, double U_total, double U_advected, double gamma)
{ return de_pressure(E, U_total, U_advected, gamma); }
/* the solvers' own load of one side's pressure (hllc_cuda.cu:107-113, roe_cuda.cu:109-115, hll_cuda.cu:106-112) */
static inline double side_pressure(cons_t u, double vx, double vy, double vz, double gamma, const pass_t *ps, int right)
{
  if (ps && ps->de) {
    double E_kin = 0.5 * u.d * (vx * vx + vy * vy + vz * vz);
    double dge   = right ? ps->qr[ps->np - 1] : ps->ql[ps->np - 1];
    return fmax(de_pressure(u.E, u.E - E_kin, dge, gamma), ORC_TINY);
  }
  return fmax((u.E - 0.5 * u.d * (vx * vx + vy * vy + vz * vz)) * (gamma - 1.0), ORC_TINY);
}

/* hllc_cuda.cu:103-305 */
static cons_t hllc_flux_x(cons_t l, cons_t r, double gamma, const pass_t *pv)
{
  cons_t f;
  double g1 = gamma - 1.0;
  int const np = pv ? pv->np : 0;
  double vxl = l.mx / l.d, vyl = l.my / l.d, vzl = l.mz / l.d;
  double pl = side_pressure(l, vxl, vyl, vzl, gamma, pv, 0);
  double vxr = r.mx / r.d, vyr = r.my / r.d, vzr = r.mz / r.d;
  double pr = side_pressure(r, vxr, vyr, vzr, gamma, pv, 1);
  double Hl = (l.E + pl) / l.d, Hr = (r.E + pr) / r.d;
  double sl = sqrt(l.d), sr = sqrt(r.d);
  double vx = (sl * vxl + sr * vxr) / (sl + sr);
  double vy = (sl * vyl + sr * vyr) / (sl + sr);
  double vz = (sl * vzl + sr * vzr) / (sl + sr);
  double H  = (sl * Hl + sr * Hr) / (sl + sr);
  double vsq = (vx * vx + vy * vy + vz * vz);
  double asq = g1 * (H - 0.5 * vsq);
  double a = sqrt(asq);
  double lm = vx - a, lp = vx + a;
  double cfl = sqrt(gamma * pl / l.d), cfr = sqrt(gamma * pr / r.d);
  double Sl = fmin(lm, vxl - cfl), Sr = fmax(lp, vxr + cfr);
  /* etah = 0: S = sgn(S)*fmax(|S|,0), sgn(0)=+1 (global_cuda.h:41-48) */
  Sl = ((Sl < 0) ? -1 : 1) * fmax(fabs(Sl), 0.0);
  Sr = ((Sr < 0) ? -1 : 1) * fmax(fabs(Sr), 0.0);
  double fdl = l.mx, fmxl = l.mx * vxl + pl, fmyl = l.my * vxl, fmzl = l.mz * vxl, fEl = (l.E + pl) * vxl;
  double fdr = r.mx, fmxr = r.mx * vxr + pr, fmyr = r.my * vxr, fmzr = r.mz * vxr, fEr = (r.E + pr) * vxr;
  if (Sl > 0.0) {
    f.d = fdl; f.mx = fmxl; f.my = fmyl; f.mz = fmzl; f.E = fEl;
    for (int i = 0; i < np; i++) pv->fq[i] = pv->ql[i] * vxl; /* hllc_cuda.cu:183-190, 211-218 */
    return f;
  }
  if (Sr < 0.0) {
    f.d = fdr; f.mx = fmxr; f.my = fmyr; f.mz = fmzr; f.E = fEr;
    for (int i = 0; i < np; i++) pv->fq[i] = pv->qr[i] * vxr;
    return f;
  }
  double Sm = (r.d * vxr * (Sr - vxr) - l.d * vxl * (Sl - vxl) + pl - pr) / (r.d * (Sr - vxr) - l.d * (Sl - vxl));
  double ps = l.d * (vxl - Sl) * (vxl - Sm) + pl;
  double dls = l.d * (Sl - vxl) / (Sl - Sm);
  double mxls = (l.mx * (Sl - vxl) + ps - pl) / (Sl - Sm);
  double myls = dls * vyl, mzls = dls * vzl;
  double Els = (l.E * (Sl - vxl) - pl * vxl + ps * Sm) / (Sl - Sm);
  double drs = r.d * (Sr - vxr) / (Sr - Sm);
  double mxrs = (r.mx * (Sr - vxr) + ps - pr) / (Sr - Sm);
  double myrs = drs * vyr, mzrs = drs * vzr;
  double Ers = (r.E * (Sr - vxr) - pr * vxr + ps * Sm) / (Sr - Sm);
  f.d  = 0.5 * (fdl + fdr + (Sr - fabs(Sm)) * drs + (Sl + fabs(Sm)) * dls - Sl * l.d - Sr * r.d);
  f.mx = 0.5 * (fmxl + fmxr + (Sr - fabs(Sm)) * mxrs + (Sl + fabs(Sm)) * mxls - Sl * l.mx - Sr * r.mx);
  f.my = 0.5 * (fmyl + fmyr + (Sr - fabs(Sm)) * myrs + (Sl + fabs(Sm)) * myls - Sl * l.my - Sr * r.my);
  f.mz = 0.5 * (fmzl + fmzr + (Sr - fabs(Sm)) * mzrs + (Sl + fabs(Sm)) * mzls - Sl * l.mz - Sr * r.mz);
  f.E  = 0.5 * (fEl + fEr + (Sr - fabs(Sm)) * Ers + (Sl + fabs(Sm)) * Els - Sl * l.E - Sr * r.E);
  for (int i = 0; i < np; i++) { /* hllc_cuda.cu:119-131, 249-283 */
    double dql = pv->ql[i], dqr = pv->qr[i];
    double sql = dql / l.d, sqr = dqr / r.d;
    double qls = dls * sql, qrs = drs * sqr;
    pv->fq[i] = 0.5 * (dql * vxl + dqr * vxr + (Sr - fabs(Sm)) * qrs + (Sl + fabs(Sm)) * qls - Sl * dql - Sr * dqr);
  }
  return f;
}
static cons_t hllc_flux(cons_t l, cons_t r, double gamma) { return hllc_flux_x(l, r, gamma, NULL); }

/* hll_cuda.cu:62-262: Davis wave-speed bounds from the two sound speeds, single-state HLL flux */
static cons_t hll_flux_x(cons_t l, cons_t r, double gamma, const pass_t *ps)
{
  cons_t f;
  int const np = ps ? ps->np : 0;
  double vxl = l.mx / l.d, vyl = l.my / l.d, vzl = l.mz / l.d;
  double pl = side_pressure(l, vxl, vyl, vzl, gamma, ps, 0);
  double vxr = r.mx / r.d, vyr = r.my / r.d, vzr = r.mz / r.d;
  double pr = side_pressure(r, vxr, vyr, vzr, gamma, ps, 1);
  double cfl = sqrt(gamma * pl / l.d), cfr = sqrt(gamma * pr / r.d);
  double Sl = fmin(vxr - cfr, vxl - cfl), Sr = fmax(vxl + cfl, vxr + cfr);
  double fdl = l.mx, fmxl = l.mx * vxl + pl, fmyl = l.my * vxl, fmzl = l.mz * vxl, fEl = (l.E + pl) * vxl;
  double fdr = r.mx, fmxr = r.mx * vxr + pr, fmyr = r.my * vxr, fmzr = r.mz * vxr, fEr = (r.E + pr) * vxr;
  if (Sl > 0.0) {
    f.d = fdl; f.mx = fmxl; f.my = fmyl; f.mz = fmzl; f.E = fEl;
    for (int i = 0; i < np; i++) ps->fq[i] = ps->ql[i] * vxl; /* hll_cuda.cu:181-188 */
    return f;
  }
  if (Sr < 0.0) {
    f.d = fdr; f.mx = fmxr; f.my = fmyr; f.mz = fmzr; f.E = fEr;
    for (int i = 0; i < np; i++) ps->fq[i] = ps->qr[i] * vxr;
    return f;
  }
  for (int i = 0; i < np; i++) /* hll_cuda.cu:243-250 */
    ps->fq[i] = ((Sr * (ps->ql[i] * vxl)) - (Sl * (ps->qr[i] * vxr)) + Sl * Sr * (ps->qr[i] - ps->ql[i])) / (Sr - Sl);
  f.d  = ((Sr * fdl) - (Sl * fdr) + Sl * Sr * (r.d - l.d)) / (Sr - Sl);
  f.mx = ((Sr * fmxl) - (Sl * fmxr) + Sl * Sr * (r.mx - l.mx)) / (Sr - Sl);
  f.my = ((Sr * fmyl) - (Sl * fmyr) + Sl * Sr * (r.my - l.my)) / (Sr - Sl);
  f.mz = ((Sr * fmzl) - (Sl * fmzr) + Sl * Sr * (r.mz - l.mz)) / (Sr - Sl);
  f.E  = ((Sr * fEl) - (Sl * fEr) + Sl * Sr * (r.E - l.E)) / (Sr - Sl);
  return f;
}
static cons_t hll_flux(cons_t l, cons_t r, double gamma) { return hll_flux_x(l, r, gamma, NULL); }

/* roe_cuda.cu:105-398 (etah = 0) */
static cons_t roe_flux_x(cons_t l, cons_t r, double gamma, const pass_t *ps)
{
  cons_t f;
  double g1 = gamma - 1.0, etah = 0.0;
  int const np = ps ? ps->np : 0;
  double vxl = l.mx / l.d, vyl = l.my / l.d, vzl = l.mz / l.d;
  double pl = side_pressure(l, vxl, vyl, vzl, gamma, ps, 0);
  double vxr = r.mx / r.d, vyr = r.my / r.d, vzr = r.mz / r.d;
  double pr = side_pressure(r, vxr, vyr, vzr, gamma, ps, 1);
  double Hl = (l.E + pl) / l.d, Hr = (r.E + pr) / r.d;
  double sl = sqrt(l.d), sr = sqrt(r.d);
  double vx = (sl * vxl + sr * vxr) / (sl + sr);
  double vy = (sl * vyl + sr * vyr) / (sl + sr);
  double vz = (sl * vzl + sr * vzr) / (sl + sr);
  double H  = (sl * Hl + sr * Hr) / (sl + sr);
  double vsq = (vx * vx + vy * vy + vz * vz);
  double asq = g1 * fmax((H - 0.5 * vsq), ORC_TINY);
  double a = sqrt(asq);
  double lm = vx - a, l0 = vx, lp = vx + a;
  double fdl = l.mx, fmxl = l.mx * vxl + pl, fmyl = l.mx * vyl, fmzl = l.mx * vzl, fEl = (l.E + pl) * vxl;
  double fdr = r.mx, fmxr = r.mx * vxr + pr, fmyr = r.mx * vyr, fmzr = r.mx * vzr, fEr = (r.E + pr) * vxr;
  if (lm >= 0.0) {
    f.d = fdl; f.mx = fmxl; f.my = fmyl; f.mz = fmzl; f.E = fEl;
    for (int i = 0; i < np; i++) ps->fq[i] = l.mx * (ps->ql[i] / l.d); /* roe_cuda.cu:116-123, 174-181 */
    return f;
  }
  if (lp <= 0.0) {
    f.d = fdr; f.mx = fmxr; f.my = fmyr; f.mz = fmzr; f.E = fEr;
    for (int i = 0; i < np; i++) ps->fq[i] = r.mx * (ps->qr[i] / r.d);
    return f;
  }
  double dd = r.d - l.d, dmx = r.mx - l.mx, dmy = r.my - l.my, dmz = r.mz - l.mz, dE = r.E - l.E;
  double Na = 0.5 / asq, coeff;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;
  double a0 = dd * Na * (0.5 * g1 * vsq + vx * a) - dmx * Na * (g1 * vx + a) - dmy * Na * g1 * vy -
              dmz * Na * g1 * vz + dE * Na * g1;
  coeff = a0 * fmax(fabs(lm), etah);
  s0 += coeff; s1 += coeff * (vx - a); s2 += coeff * vy; s3 += coeff * vz; s4 += coeff * (H - vx * a);
  double a1 = -dd * vy + dmy;
  coeff = a1 * fmax(fabs(l0), etah);
  s2 += coeff; s4 += coeff * vy;
  double a2 = -dd * vz + dmz;
  coeff = a2 * fmax(fabs(l0), etah);
  s3 += coeff; s4 += coeff * vz;
  double a3 = dd * (1.0 - Na * g1 * vsq) + dmx * g1 * vx / asq + dmy * g1 * vy / asq + dmz * g1 * vz / asq -
              dE * g1 / asq;
  coeff = a3 * fmax(fabs(l0), etah);
  s0 += coeff; s1 += coeff * vx; s2 += coeff * vy; s3 += coeff * vz; s4 += coeff * 0.5 * vsq;
  double a4 = dd * Na * (0.5 * g1 * vsq - vx * a) - dmx * Na * (g1 * vx - a) - dmy * Na * g1 * vy -
              dmz * Na * g1 * vz + dE * Na * g1;
  coeff = a4 * fmax(fabs(lp), etah);
  s0 += coeff; s1 += coeff * (vx + a); s2 += coeff * vy; s3 += coeff * vz; s4 += coeff * (H + vx * a);
  int hlle = 0;
  double t0 = l.d + a0, t1 = l.mx + a0 * (vx - a), t2 = l.my + a0 * vy, t3 = l.mz + a0 * vz,
         t4 = l.E + a0 * (H - vx * a);
  if (l0 > lm) {
    if (t0 <= 0.0) hlle = 1;
    if (t4 - 0.5 * (t1 * t1 + t2 * t2 + t3 * t3) / t0 < 0.0) hlle = 2;
  }
  t0 += a3 + a4;
  t1 += a3 * vx;
  t2 += a1 + a3 * vy;
  t3 += a2 + a3 * vz;
  t4 += a1 * vy + a2 * vz + a3 * 0.5 * vsq;
  if (lp > l0) {
    if (t0 <= 0.0) hlle = 1;
    if (t4 - 0.5 * (t1 * t1 + t2 * t2 + t3 * t3) / t0 < 0.0) hlle = 2;
  }
  if (hlle != 0) {
    double cfl = sqrt(gamma * pl / l.d), cfr = sqrt(gamma * pr / r.d);
    double bm = fmin(fmin(lm, vxl - cfl), 0.0), bp = fmax(fmax(lp, vxr + cfr), 0.0);
    fdl = l.mx - bm * l.d;             fdr = r.mx - bp * r.d;
    fmxl = l.mx * (vxl - bm) + pl;     fmxr = r.mx * (vxr - bp) + pr;
    fmyl = l.my * (vxl - bm);          fmyr = r.my * (vxr - bp);
    fmzl = l.mz * (vxl - bm);          fmzr = r.mz * (vxr - bp);
    fEl = l.E * (vxl - bm) + pl * vxl; fEr = r.E * (vxr - bp) + pr * vxr;
    double tmp = 0.5 * (bp + bm) / (bp - bm);
    f.d  = 0.5 * (fdl + fdr) + (fdl - fdr) * tmp;
    f.mx = 0.5 * (fmxl + fmxr) + (fmxl - fmxr) * tmp;
    f.my = 0.5 * (fmyl + fmyr) + (fmyl - fmyr) * tmp;
    f.mz = 0.5 * (fmzl + fmzr) + (fmzl - fmzr) * tmp;
    f.E  = 0.5 * (fEl + fEr) + (fEl - fEr) * tmp;
    for (int i = 0; i < np; i++) { /* roe_cuda.cu:343-371 */
      double fl = ps->ql[i] * (vxl - bm), fr = ps->qr[i] * (vxr - bp);
      ps->fq[i] = 0.5 * (fl + fr) + (fl - fr) * tmp;
    }
    return f;
  }
  f.d  = 0.5 * (fdl + fdr - s0);
  f.mx = 0.5 * (fmxl + fmxr - s1);
  f.my = 0.5 * (fmyl + fmyr - s2);
  f.mz = 0.5 * (fmzl + fmzr - s3);
  f.E  = 0.5 * (fEl + fEr - s4);
  for (int i = 0; i < np; i++) /* upwinded on the sign of the mass flux, roe_cuda.cu:381-396 */
    ps->fq[i] = (f.d >= 0.0) ? f.d * (ps->ql[i] / l.d) : f.d * (ps->qr[i] / r.d);
  return f;
}
static cons_t roe_flux(cons_t l, cons_t r, double gamma) { return roe_flux_x(l, r, gamma, NULL); }

/* one launch of Calculate_{HLLC,Roe}_Fluxes_CUDA: every cell, F[id] = flux(Q_L[id], Q_R[id]) */
/* exact_cuda.cu:21-334 (Toro 1999, 4.9).  PARITY UNPINNED ON THE CPU: the reference evaluates its powers with the
 * device's powf (single precision, not correctly rounded) and no CPU libm reproduces those bits, so this restatement
 * agrees with the reference's kernels to float round-off (~1e-7) only.  The bit-level check of the CUDA path for this
 * solver is the reference's own -DEXACT build in tests/test_gpu_reference.py. */
static double exact_pow(double a, double b) { return (double)powf((float)a, (float)b); }
static void exact_prefun(double *f, double *fd, double p, double dk, double pk, double ck, double gamma)
{
  if (p <= pk) {
    *f  = (2.0 / (gamma - 1.0)) * ck * (exact_pow(p / pk, (gamma - 1.0) / (2.0 * gamma)) - 1.0);
    *fd = (1.0 / (dk * ck)) * exact_pow((p / pk), -((gamma + 1.0) / (2.0 * gamma)));
  } else {
    double qrt = sqrt(((2.0 / (gamma + 1.0)) / dk) / ((((gamma - 1.0) / (gamma + 1.0)) * pk) + p));
    *f  = (p - pk) * qrt;
    *fd = (1.0 - 0.5 * (p - pk) / ((((gamma - 1.0) / (gamma + 1.0)) * pk) + p)) * qrt;
  }
}
static cons_t exact_flux(cons_t l, cons_t r, double gamma)
{
  double dl = l.d, vxl = l.mx / dl, vyl = l.my / dl, vzl = l.mz / dl;
  double pl = fmax((l.E - 0.5 * dl * (vxl * vxl + vyl * vyl + vzl * vzl)) * (gamma - 1.0), ORC_TINY);
  double dr = r.d, vxr = r.mx / dr, vyr = r.my / dr, vzr = r.mz / dr;
  double pr = fmax((r.E - 0.5 * dr * (vxr * vxr + vyr * vyr + vzr * vzr)) * (gamma - 1.0), ORC_TINY);
  double cl = sqrt(gamma * pl / dl), cr = sqrt(gamma * pr / dr);
  const double TOL = 1.0e-6;
  /* guessp_CUDA */
  double ppv = 0.5 * (pl + pr) + 0.125 * (vxl - vxr) * (dl + dr) * (cl + cr);
  if (ppv < 0.0) ppv = 0.0;
  double gl = sqrt((2.0 / ((gamma + 1.0) * dl)) / (((gamma - 1.0) / (gamma + 1.0)) * pl + ppv));
  double gr = sqrt((2.0 / ((gamma + 1.0) * dr)) / (((gamma - 1.0) / (gamma + 1.0)) * pr + ppv));
  double pold = (gl * pl + gr * pr - (vxr - vxl)) / (gl + gr);
  if (pold < 0.0) pold = TOL;
  /* starpv_CUDA */
  double pm = pold, fl = 0, fld = 0, fr = 0, frd = 0;
  for (int i = 0; i <= 20; i++) {
    exact_prefun(&fl, &fld, pold, dl, pl, cl, gamma);
    exact_prefun(&fr, &frd, pold, dr, pr, cr, gamma);
    pm = pold - (fl + fr + vxr - vxl) / (fld + frd);
    double change = 2.0 * fabs((pm - pold) / (pm + pold));
    if (change <= TOL) break;
    if (pm < 0.0) pm = TOL;
    pold = pm;
  }
  double vm = 0.5 * (vxl + vxr + fr - fl);
  /* sample_CUDA */
  double ds, vs, ps, c;
  double g1 = (gamma - 1.0) / (gamma + 1.0);
  if (vm >= 0) {
    if (pm <= pl) {
      if (vxl - cl >= 0) { ds = dl; vs = vxl; ps = pl; }
      else if (vm - cl * exact_pow(pm / pl, (gamma - 1.0) / (2.0 * gamma)) < 0) {
        ds = dl * exact_pow(pm / pl, 1.0 / gamma); vs = vm; ps = pm;
      } else {
        c  = (2.0 / (gamma + 1.0)) * (cl + ((gamma - 1.0) / 2.0) * vxl);
        vs = c;
        ds = dl * exact_pow(c / cl, 2.0 / (gamma - 1.0));
        ps = pl * exact_pow(c / cl, 2.0 * gamma / (gamma - 1.0));
      }
    } else {
      double sl = vxl - cl * sqrt(((gamma + 1.0) / (2.0 * gamma)) * (pm / pl) + ((gamma - 1.0) / (2.0 * gamma)));
      if (sl >= 0) { ds = dl; vs = vxl; ps = pl; }
      else { ds = dl * (pm / pl + g1) / ((pm / pl) * g1 + 1.0); vs = vm; ps = pm; }
    }
  } else {
    if (pm > pr) {
      double sr = vxr + cr * sqrt(((gamma + 1.0) / (2.0 * gamma)) * (pm / pr) + ((gamma - 1.0) / (2.0 * gamma)));
      if (sr <= 0) { ds = dr; vs = vxr; ps = pr; }
      else { ds = dr * (pm / pr + g1) / ((pm / pr) * g1 + 1.0); vs = vm; ps = pm; }
    } else {
      if (vxr + cr <= 0) { ds = dr; vs = vxr; ps = pr; }
      else if (vm + cr * exact_pow(pm / pr, (gamma - 1.0) / (2.0 * gamma)) >= 0) {
        ds = dr * exact_pow(pm / pr, (1.0 / gamma)); vs = vm; ps = pm;
      } else {
        c  = (2.0 / (gamma + 1.0)) * (cr - ((gamma - 1.0) / 2.0) * vxr);
        vs = -c;
        ds = dr * exact_pow(c / cr, 2.0 / (gamma - 1.0));
        ps = pr * exact_pow(c / cr, 2.0 * gamma / (gamma - 1.0));
      }
    }
  }
  double vy = vs >= 0 ? vyl : vyr, vz = vs >= 0 ? vzl : vzr;
  cons_t f;
  f.d  = ds * vs;
  f.mx = ds * vs * vs + ps;
  f.my = ds * vs * vy;
  f.mz = ds * vs * vz;
  double Es = (ps / (gamma - 1.0)) + 0.5 * ds * (vs * vs + vy * vy + vz * vz);
  f.E = (Es + ps) * vs;
  return f;
}
static cons_t flux_of(int solver, cons_t l, cons_t r, double gamma)
{
  if (solver == ORC_ROE) return roe_flux(l, r, gamma);
  if (solver == ORC_HLL) return hll_flux(l, r, gamma);
  if (solver == ORC_EXACT) return exact_flux(l, r, gamma);
  return hllc_flux(l, r, gamma);
}

void orc_riemann(int solver, const double *QL, const double *QR, double *F, int nx, int ny, int nz, double gamma,
                 int dir)
{
  size_t N = (size_t)nx * ny * nz;
  int o1, o2, o3;
  dir_perm(dir, &o1, &o2, &o3);
#pragma omp parallel for schedule(static)
  for (long id = 0; id < (long)N; id++) {
    cons_t l = {QL[id], QL[o1 * N + id], QL[o2 * N + id], QL[o3 * N + id], QL[4 * N + id]};
    cons_t r = {QR[id], QR[o1 * N + id], QR[o2 * N + id], QR[o3 * N + id], QR[4 * N + id]};
    cons_t f = flux_of(solver, l, r, gamma);
    F[id] = f.d; F[o1 * N + id] = f.mx; F[o2 * N + id] = f.my; F[o3 * N + id] = f.mz; F[4 * N + id] = f.E;
  }
}

void orc_riemann_single(int solver, const double *l5, const double *r5, double gamma, double *f5)
{
  cons_t l = {l5[0], l5[1], l5[2], l5[3], l5[4]}, r = {r5[0], r5[1], r5[2], r5[3], r5[4]};
  cons_t f = flux_of(solver, l, r, gamma);
  f5[0] = f.d; f5[1] = f.mx; f5[2] = f.my; f5[3] = f.mz; f5[4] = f.E;
}

/* ---- static gravity (-DSTATIC_GRAV) ------------------------------------------------------------------------------
 * calc_g_{1,2,3}D (src/gravity/static_grav.h:16-210) evaluated for every cell of the local array, and the source terms
 * of Update_Conserved_Variables_{1,2,3}D (hydro_cuda.cu:59-68, 141-153, 301-317).  PARITY UNPINNED ON THE CPU for the
 * profiles that call log / pow / atan2 / cos / sin: the device library's results are not correctly rounded and no CPU
 * libm reproduces their bits (the Rayleigh-Taylor field, custom_grav = 2 in 2-D, is exact).  The bit-level check of the
 * CUDA path is the reference's own -DSTATIC_GRAV build (tests/test_gpu_gravity.py). */
#define ORC_GN 4.49451e-18     /* global.h:29 */
#define ORC_MSUN_CGS 1.98847e33 /* global.h:35 */
static inline double pow2(double x) { return x * x; }
static void disk_halo_3d(double x_pos, double y_pos, double z_pos, double M_vir, double M_d, double R_d, double z_d,
                         double R_vir, double c_vir, double *gx, double *gy, double *gz)
{ /* static_grav.h:142-199 */
  double r_disk = sqrt(x_pos * x_pos + y_pos * y_pos);
  double r_halo = sqrt(x_pos * x_pos + y_pos * y_pos + z_pos * z_pos);
  double M_h = M_vir - M_d, R_h = R_vir / c_vir;
  double phi_0_h = ORC_GN * M_h / (log(1.0 + c_vir) - c_vir / (1.0 + c_vir));
  double x = r_halo / R_h;
  double a_halo = -phi_0_h * (log(1 + x) - x / (1 + x)) / (r_halo * r_halo);
  double a_halo_r = a_halo * (r_disk / r_halo), a_halo_z = a_halo * (z_pos / r_halo);
  double a_disk_r = -ORC_GN * M_d * r_disk * pow(r_disk * r_disk + pow2(R_d + sqrt(z_pos * z_pos + z_d * z_d)), -1.5);
  double a_disk_z = -ORC_GN * M_d * z_pos * (R_d + sqrt(z_pos * z_pos + z_d * z_d)) /
                    (pow(r_disk * r_disk + pow2(R_d + sqrt(z_pos * z_pos + z_d * z_d)), 1.5) * sqrt(z_pos * z_pos + z_d * z_d));
  *gx = (x_pos / r_disk) * (a_disk_r + a_halo_r);
  *gy = (y_pos / r_disk) * (a_disk_r + a_halo_r);
  *gz = a_disk_z + a_halo_z;
}
void orc_static_gravity_field(double *g, int nx, int ny, int nz, int n_ghost, int custom_grav, int x_off, int y_off,
                              int z_off, double dx, double dy, double dz, double xbound, double ybound, double zbound)
{
  size_t N = (size_t)nx * ny * nz;
  int ndim = 1 + (ny > 1) + (nz > 1);
  for (int zid = 0; zid < nz; zid++)
    for (int yid = 0; yid < ny; yid++)
      for (int xid = 0; xid < nx; xid++) {
        size_t id = (size_t)xid + (size_t)nx * ((size_t)yid + (size_t)ny * zid);
        double gx = 0, gy = 0, gz = 0;
        double x_pos = (x_off + xid - n_ghost + 0.5) * dx + xbound;
        if (ndim == 1) {
          if (custom_grav == 1) { /* static_grav.h:22-55 */
            double r_disk = 13.9211647546, r_halo = sqrt(x_pos * x_pos + r_disk * r_disk);
            double M_vir = 1.0e12, M_d = 6.5e10, M_h = M_vir - M_d, R_vir = 261, c_vir = 20.0, R_h = R_vir / c_vir;
            double R_d = 3.5, z_d = 3.5 / 5.0;
            double phi_0_h = ORC_GN * M_h / (log(1.0 + c_vir) - c_vir / (1.0 + c_vir));
            double x = r_halo / R_h;
            double a_halo = -phi_0_h * (log(1 + x) - x / (1 + x)) / (r_halo * r_halo);
            double a_disk_z = -ORC_GN * M_d * x_pos * (R_d + sqrt(x_pos * x_pos + z_d * z_d)) /
                              (pow(r_disk * r_disk + pow2(R_d + sqrt(x_pos * x_pos + z_d * z_d)), 1.5) *
                               sqrt(x_pos * x_pos + z_d * z_d));
            gx = (x_pos / r_halo) * a_halo + a_disk_z;
          }
        } else if (ndim == 2) { /* static_grav.h:60-124 */
          double y_pos = (y_off + yid - n_ghost + 0.5) * dy + ybound;
          double r = sqrt(x_pos * x_pos + y_pos * y_pos), phi = atan2(y_pos, x_pos);
          if (custom_grav == 1) {
            if (r < 0.2) { gx = -cos(phi) * 25.0 * r; gy = -sin(phi) * 25.0 * r; }
            else if (r >= 0.2 && r < 0.4) {
              gx = -cos(phi) * (4.0 - 20.0 * r + 25.0 * r * r) / r;
              gy = -sin(phi) * (4.0 - 20.0 * r + 25.0 * r * r) / r;
            }
          } else if (custom_grav == 2) { gx = 0; gy = -1; }
          else if (custom_grav == 3) {
            double M = 1 * ORC_MSUN_CGS;
            gx = -cos(phi) * ORC_GN * M / (r * r);
            gy = -sin(phi) * ORC_GN * M / (r * r);
          } else if (custom_grav == 4) {
            double M_vir = 1.0e12, M_d = 6.5e10, M_h = M_vir - M_d, R_vir = 261, c_vir = 20, R_s = R_vir / c_vir, R_d = 3.5;
            double x = r / R_s;
            double a_d = ORC_GN * M_d * r * pow(r * r + R_d * R_d, -1.5);
            double a_h = ORC_GN * M_h * (log(1 + x) - x / (1 + x)) / ((log(1 + c_vir) - c_vir / (1 + c_vir)) * r * r);
            double a = a_d + a_h;
            gx = -cos(phi) * a; gy = -sin(phi) * a;
          }
        } else {
          double y_pos = (y_off + yid - n_ghost + 0.5) * dy + ybound, z_pos = (z_off + zid - n_ghost + 0.5) * dz + zbound;
          if (custom_grav == 1) disk_halo_3d(x_pos, y_pos, z_pos, 1.0e12, 6.5e10, 3.5, 3.5 / 5.0, 261., 20.0, &gx, &gy, &gz);
          else if (custom_grav == 2) disk_halo_3d(x_pos, y_pos, z_pos, 5.0e10, 1.0e10, 0.8, 0.15, 0.8 / 0.015, 10.0, &gx, &gy, &gz);
        }
        g[id] = gx;
        if (ndim > 1) g[N + id] = gy;
        if (ndim > 2) g[2 * N + id] = gz;
      }
}
/* the field the next orc_vl_step / orc_vl_step_x calls add the source terms of ([ndim][N]; NULL = a build without gravity) */
static const double *g_grav = NULL;
void orc_set_gravity(const double *g) { g_grav = g; }
/* -DGRAVITY (hydro_cuda.cu:317-379): potential with ghost cells, [N]; five_point = -DGRAVITY_5_POINTS_GRADIENT.  3-D only. */
static const double *g_pot = NULL;
static int g_pot5 = 0;
void orc_set_gravity_potential(const double *pot, int five_point) { g_pot = pot; g_pot5 = five_point; }
static void gravity_potential_cell(double *U, size_t N, size_t id, long sy, long sz, double dt, double dx, double dy,
                                   double dz, double d, const double *m_old)
{
  double d_inv = 1.0 / d, vx = m_old[0] * d_inv, vy = m_old[1] * d_inv, vz = m_old[2] * d_inv;
  double d_n = U[id], d_inv_n = 1.0 / d_n;
  double vx_n = U[N + id] * d_inv_n, vy_n = U[2 * N + id] * d_inv_n, vz_n = U[3 * N + id] * d_inv_n;
  const double *p = g_pot;
  double gx, gy, gz;
  if (g_pot5) { /* y and z divide by 12 dx as well, as the reference does */
    gx = -1 * (-p[id + 2] + 8 * p[id + 1] - 8 * p[id - 1] + p[id - 2]) / (12 * dx);
    gy = -1 * (-p[id + 2 * sy] + 8 * p[id + sy] - 8 * p[id - sy] + p[id - 2 * sy]) / (12 * dx);
    gz = -1 * (-p[id + 2 * sz] + 8 * p[id + sz] - 8 * p[id - sz] + p[id - 2 * sz]) / (12 * dx);
  } else {
    gx = -0.5 * (p[id + 1] - p[id - 1]) / dx;
    gy = -0.5 * (p[id + sy] - p[id - sy]) / dy;
    gz = -0.5 * (p[id + sz] - p[id - sz]) / dz;
  }
  U[N + id] += 0.5 * dt * gx * (d + d_n);
  U[2 * N + id] += 0.5 * dt * gy * (d + d_n);
  U[3 * N + id] += 0.5 * dt * gz * (d + d_n);
  U[4 * N + id] += 0.5 * dt * (gx * (d * vx + d_n * vx_n) + gy * (d * vy + d_n * vy_n) + gz * (d * vz + d_n * vz_n));
}
/* one cell: old = (rho, momenta)^n read before the flux update, U already updated (and density-floored) */
static void gravity_source_cell(double *U, size_t N, size_t id, int ndim, double dt, double d, const double *m_old)
{
  double d_inv = 1.0 / d, d_n = U[id], d_inv_n = 1.0 / d_n;
  double v[3], v_n[3];
  for (int c = 0; c < ndim; c++) { v[c] = m_old[c] * d_inv; v_n[c] = U[(size_t)(1 + c) * N + id] * d_inv_n; }
  double dE = 0.25 * dt * g_grav[id] * (d + d_n) * (v[0] + v_n[0]);
  for (int c = 1; c < ndim; c++) dE = dE + 0.25 * dt * g_grav[(size_t)c * N + id] * (d + d_n) * (v[c] + v_n[c]);
  for (int c = 0; c < ndim; c++) U[(size_t)(1 + c) * N + id] += 0.5 * dt * g_grav[(size_t)c * N + id] * (d + d_n);
  U[4 * N + id] += dE;
}

/* ---- one full VL step (VL_3D_cuda.cu:38-360; VL_2D_cuda.cu; VL_1D_cuda.cu) ---- */
void orc_vl_step(double *U, int nx, int ny, int nz, int n_ghost, double dx, double dy, double dz, double dt,
                 double gamma, int recon, int solver, double density_floor)
{
  size_t N = (size_t)nx * ny * nz;
  int ndim = 1 + (ny > 1) + (nz > 1);
  double dd[3] = {dx, dy, dz};
  long st[3] = {1, nx, (long)nx * ny};
  double *Uh = calloc(5 * N, sizeof(double));
  double *QL[3], *QR[3], *F[3];
  for (int d = 0; d < ndim; d++) {
    QL[d] = calloc(5 * N, sizeof(double));
    QR[d] = calloc(5 * N, sizeof(double));
    F[d]  = calloc(5 * N, sizeof(double));
  }
  /* stage 1-2: PCM + first-order fluxes */
  for (int d = 0; d < ndim; d++) {
    pcm_dir(U, QL[d], QR[d], nx, ny, nz, d);
    orc_riemann(solver, QL[d], QR[d], F[d], nx, ny, nz, gamma, d);
  }
  /* stage 3: half-step update on all but the outermost ring (VL_3D_cuda.cu:379-495) */
  {
    double h = 0.5 * dt;
    double dto[3] = {h / dx, h / dy, h / dz};
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < nz; z++)
      for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) {
          if (!(x > 0 && x < nx - 1)) continue;
          if (ndim > 1 && !(y > 0 && y < ny - 1)) continue;
          if (ndim > 2 && !(z > 0 && z < nz - 1)) continue;
          size_t id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
          for (int f = 0; f < 5; f++) {
            double v = U[f * N + id];
            for (int d = 0; d < ndim; d++) v = v + dto[d] * (F[d][f * N + id - st[d]] - F[d][f * N + id]);
            Uh[f * N + id] = v;
          }
          if (ndim == 3 && Uh[id] < density_floor) { /* DENSITY_FLOOR, 3-D half step only */
            double d0 = Uh[id];
            Uh[id] = density_floor;
            for (int f = 1; f < 5; f++) Uh[f * N + id] *= (density_floor / d0);
          }
        }
  }
  /* stage 4-5: second-order interfaces from the half-step state, then fluxes */
  for (int d = 0; d < ndim; d++) {
    /* VL_2D_cuda.cu:95-120 has no -DPCM branch: the 2-D corrector re-solves the step-1 interface arrays (copies of U^n) */
    if (recon == ORC_PCM) pcm_dir(ndim == 2 ? U : Uh, QL[d], QR[d], nx, ny, nz, d);
    else if (recon == ORC_PLMP) plmp_dir(Uh, QL[d], QR[d], nx, ny, nz, gamma, d);
    else if (recon == ORC_PPMP) ppmp_dir(Uh, QL[d], QR[d], nx, ny, nz, gamma, d);
    else orc_reconstruct(recon, Uh, QL[d], QR[d], nx, ny, nz, gamma, d);
    orc_riemann(solver, QL[d], QR[d], F[d], nx, ny, nz, gamma, d);
  }
  /* stage 6: full update of the real cells, in place (hydro_cuda.cu:20-300) */
  {
    double dto[3] = {dt / dx, dt / dy, dt / dz};
    int g = n_ghost;
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < nz; z++)
      for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) {
          if (!(x > g - 1 && x < nx - g)) continue;
          if (ndim > 1 && !(y > g - 1 && y < ny - g)) continue;
          if (ndim > 2 && !(z > g - 1 && z < nz - g)) continue;
          size_t id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
          double d_old = U[id], m_old[3] = {U[N + id], U[2 * N + id], U[3 * N + id]};
          for (int f = 0; f < 5; f++) {
            double s = dto[0] * (F[0][f * N + id - st[0]] - F[0][f * N + id]);
            for (int d = 1; d < ndim; d++) s = s + dto[d] * (F[d][f * N + id - st[d]] - F[d][f * N + id]);
            U[f * N + id] += s;
          }
          if (ndim == 3 && U[id] < density_floor && U[id] > 0) { /* hydro_cuda.cu:279-291 */
            double d0 = U[id];
            U[id] = density_floor;
            for (int f = 1; f < 5; f++) U[f * N + id] *= (density_floor / d0);
          }
          if (g_grav) gravity_source_cell(U, N, id, ndim, dt, d_old, m_old);
          if (g_pot && ndim == 3) gravity_potential_cell(U, N, id, st[1], st[2], dt, dx, dy, dz, d_old, m_old);
          /* the rho<=0 neighbour-averaging branch (hydro_cuda.cu:292-297) races in the
           * reference and is not restated */
        }
  }
  (void)dd;
  free(Uh);
  for (int d = 0; d < ndim; d++) { free(QL[d]); free(QR[d]); free(F[d]); }
}

/* ==== -DSCALAR / -DDE builds =================================================================================
 * ns passive scalars ride in fields 5 .. 5+ns-1 and, with de, the advected internal energy (GasEnergy) in the last
 * field nf-1 = 5+ns (grid_enum.h:33-78).  The scalars never feed back on the five hydro fields; DE replaces every
 * pressure on the path by Get_Pressure_From_DE (hydro_utilities.h:136-151), adds the p div v term to the GasEnergy
 * updates and ends the step with Select_Internal_Energy + Sync_Energies (hydro_cuda.cu:754-1099).
 * Restated for the Van Leer integrator with PCM / PLMC / PPMC and HLLC / Roe / HLL. */
#define ORC_MAXP 8
static prim_t load_prim_x(const double *U, size_t id, size_t N, int o1, int o2, int o3, double gamma, int ns, int de,
                          double *q)
{ /* reconstruction.h:139-212 */
  prim_t w;
  w.d  = U[id];
  w.vx = U[(size_t)o1 * N + id] / w.d;
  w.vy = U[(size_t)o2 * N + id] / w.d;
  w.vz = U[(size_t)o3 * N + id] / w.d;
  if (de) {
    double E = U[4 * N + id], ge = U[(size_t)(5 + ns) * N + id];
    double E_non_thermal = 0.5 * w.d * square_mag(w.vx, w.vy, w.vz); /* Calc_Kinetic_Energy_From_Velocity */
    w.p   = de_pressure(E, E - E_non_thermal, ge, gamma);
    q[ns] = ge / w.d;
  } else {
    double pr = U[4 * N + id] - 0.5 * w.d * square_mag(w.vx, w.vy, w.vz);
    w.p = fmax((gamma - 1.) * pr, ORC_TINY);
  }
  for (int i = 0; i < ns; i++) q[i] = U[(size_t)(5 + i) * N + id] / w.d;
  return w;
}
static void store_interface_x(prim_t w, const double *q, double *Q, size_t id, size_t N, int o1, int o2, int o3,
                              double gamma, int np)
{ /* reconstruction.h:918-968 */
  store_interface(w, Q, id, N, o1, o2, o3, gamma);
  for (int i = 0; i < np; i++) Q[(size_t)(5 + i) * N + id] = w.d * q[i];
}
/* a field without characteristics under PLMC: slopes, Monotonize, Calc_Interface_Linear, Plm_Limit_Interfaces
 * (reconstruction.h:224-296, 519-556, 645-723) */
static void plm_passive(double qm, double qi, double qp, double *l, double *r)
{
  double dL = 1.0 * (qi - qm), dR = 1.0 * (qp - qi), dC = 0.5 * (qp - qm), dG = vl_slope(dL, dR);
  double dm = mono(dL, dR, dC, dG);
  *l = qi + 1.0 * 0.5 * dm;
  *r = qi + -1.0 * 0.5 * dm;
  clamp_pair(l, r, qm, qi, qp);
}
static void reconstruct_x(int recon, const double *U, double *QL, double *QR, int nx, int ny, int nz, double gamma,
                          int dir, int ns, int de)
{
  size_t N = (size_t)nx * ny * nz;
  int o1, o2, o3, np = ns + de;
  dir_perm(dir, &o1, &o2, &o3);
  long stride = dir == 0 ? 1 : (dir == 1 ? nx : (long)nx * ny);
  int order = recon == ORC_PPMC ? 3 : 2;
#pragma omp parallel for collapse(2) schedule(static)
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++) {
        if (outside_guard(order, nx, ny, nz, x, y, z)) continue;
        size_t id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
        double qi[ORC_MAXP], qm[ORC_MAXP], qp[ORC_MAXP], qm2[ORC_MAXP], qp2[ORC_MAXP], ql[ORC_MAXP], qr[ORC_MAXP];
        prim_t wi = load_prim_x(U, id, N, o1, o2, o3, gamma, ns, de, qi);
        prim_t wm = load_prim_x(U, id - stride, N, o1, o2, o3, gamma, ns, de, qm);
        prim_t wp = load_prim_x(U, id + stride, N, o1, o2, o3, gamma, ns, de, qp);
        prim_t L, R;
        if (recon == ORC_PPMC) {
          prim_t wm2 = load_prim_x(U, id - 2 * stride, N, o1, o2, o3, gamma, ns, de, qm2);
          prim_t wp2 = load_prim_x(U, id + 2 * stride, N, o1, o2, o3, gamma, ns, de, qp2);
          ppmc_cell(wm2, wm, wi, wp, wp2, gamma, &L, &R);
          for (int i = 0; i < np; i++) /* ppmc_cuda.cu:660-671 */
            orc_ppm_single_variable(qm2[i], qm[i], qi[i], qp[i], qp2[i], &ql[i], &qr[i]);
        } else {
          plmc_cell_t(wm, wi, wp, gamma, 0, 0.0, &L, &R);
          for (int i = 0; i < np; i++) plm_passive(qm[i], qi[i], qp[i], &ql[i], &qr[i]);
        }
        store_interface_x(L, ql, QL, id, N, o1, o2, o3, gamma, np);
        store_interface_x(R, qr, QR, id - stride, N, o1, o2, o3, gamma, np);
      }
}
/* PLMP_cuda / PPMP_cuda of a -DDE / -DSCALAR build under -DVL (plmp_cuda.cu:108-262, ppmp_cuda.cu:155-283, 353-374):
 * their own load (plain kinetic-energy sum of the rotated velocities, Get_Pressure_From_DE) and the limiter of the
 * primitive variables applied to every passive field per unit mass */
static prim_t plmp_load_x(const double *U, size_t id, size_t N, int o1, int o2, int o3, double gamma, int ns, int de, double *q)
{
  prim_t w;
  w.d  = U[id];
  w.vx = U[(size_t)o1 * N + id] / w.d;
  w.vy = U[(size_t)o2 * N + id] / w.d;
  w.vz = U[(size_t)o3 * N + id] / w.d;
  if (de) {
    double E = U[4 * N + id], E_kin = 0.5 * w.d * (w.vx * w.vx + w.vy * w.vy + w.vz * w.vz);
    double dge = U[(size_t)(5 + ns) * N + id];
    w.p   = de_pressure(E, E - E_kin, dge, gamma);
    q[ns] = dge / w.d;
  } else {
    w.p = (U[4 * N + id] - 0.5 * w.d * (w.vx * w.vx + w.vy * w.vy + w.vz * w.vz)) * (gamma - 1.0);
  }
  w.p = fmax(w.p, ORC_TINY);
  for (int i = 0; i < ns; i++) q[i] = U[(size_t)(5 + i) * N + id] / w.d;
  return w;
}
static void pxmp_dir_x(int recon, const double *U, double *QL, double *QR, int nx, int ny, int nz, double gamma, int dir,
                       int ns, int de)
{
  size_t N = (size_t)nx * ny * nz;
  int o1, o2, o3, np = ns + de;
  dir_perm(dir, &o1, &o2, &o3);
  long stride = dir == 0 ? 1 : (dir == 1 ? nx : (long)nx * ny);
  int n[3] = {nx, ny, nz};
  int const ppm = recon == ORC_PPMP;
#pragma omp parallel for collapse(2) schedule(static)
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++) {
        int c = dir == 0 ? x : (dir == 1 ? y : z);
        if (ppm ? (c < 2 || c >= n[dir] - 3) : (c < 1 || c >= n[dir] - 2)) continue;
        size_t id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
        prim_t w[5];
        double q[5][ORC_MAXP], qlo[ORC_MAXP], qhi[ORC_MAXP];
        for (int k = ppm ? 0 : 1; k < (ppm ? 5 : 4); k++)
          w[k] = plmp_load_x(U, id + (k - 2) * stride, N, o1, o2, o3, gamma, ns, de, q[k]);
        prim_t lo, hi;
        if (ppm) {
          ppmp_var(w[0].d, w[1].d, w[2].d, w[3].d, w[4].d, &lo.d, &hi.d);
          ppmp_var(w[0].vx, w[1].vx, w[2].vx, w[3].vx, w[4].vx, &lo.vx, &hi.vx);
          ppmp_var(w[0].vy, w[1].vy, w[2].vy, w[3].vy, w[4].vy, &lo.vy, &hi.vy);
          ppmp_var(w[0].vz, w[1].vz, w[2].vz, w[3].vz, w[4].vz, &lo.vz, &hi.vz);
          ppmp_var(w[0].p, w[1].p, w[2].p, w[3].p, w[4].p, &lo.p, &hi.p);
          for (int i = 0; i < np; i++) ppmp_var(q[0][i], q[1][i], q[2][i], q[3][i], q[4][i], &qlo[i], &qhi[i]);
        } else {
          plmp_interface(w[1].d, w[2].d, w[3].d, &lo.d, &hi.d);
          plmp_interface(w[1].vx, w[2].vx, w[3].vx, &lo.vx, &hi.vx);
          plmp_interface(w[1].vy, w[2].vy, w[3].vy, &lo.vy, &hi.vy);
          plmp_interface(w[1].vz, w[2].vz, w[3].vz, &lo.vz, &hi.vz);
          plmp_interface(w[1].p, w[2].p, w[3].p, &lo.p, &hi.p);
          for (int i = 0; i < np; i++) plmp_interface(q[1][i], q[2][i], q[3][i], &qlo[i], &qhi[i]);
        }
        lo.d = fmax(lo.d, ORC_TINY); hi.d = fmax(hi.d, ORC_TINY);
        lo.p = fmax(lo.p, ORC_TINY); hi.p = fmax(hi.p, ORC_TINY);
        plmp_store(lo, QR, id - stride, N, o1, o2, o3, gamma);
        plmp_store(hi, QL, id, N, o1, o2, o3, gamma);
        for (int i = 0; i < np; i++) {
          QR[(size_t)(5 + i) * N + id - stride] = lo.d * qlo[i];
          QL[(size_t)(5 + i) * N + id]          = hi.d * qhi[i];
        }
      }
}
static void pcm_dir_x(const double *U, double *QL, double *QR, int nx, int ny, int nz, int dir, int nf)
{
  size_t N = (size_t)nx * ny * nz;
  long stride = dir == 0 ? 1 : (dir == 1 ? nx : (long)nx * ny);
  for (int f = 0; f < nf; f++) {
#pragma omp parallel for schedule(static)
    for (long id = 0; id < (long)N; id++) {
      QL[f * N + id] = U[f * N + id];
      int x = id % nx, y = (id / nx) % ny, z = id / ((long)nx * ny);
      int c = dir == 0 ? x : (dir == 1 ? y : z);
      if (c > 0) QR[f * N + id - stride] = U[f * N + id];
    }
  }
}
static void riemann_x(int solver, const double *QL, const double *QR, double *F, int nx, int ny, int nz, double gamma,
                      int dir, int ns, int de)
{
  size_t N = (size_t)nx * ny * nz;
  int o1, o2, o3, np = ns + de;
  dir_perm(dir, &o1, &o2, &o3);
#pragma omp parallel for schedule(static)
  for (long id = 0; id < (long)N; id++) {
    cons_t l = {QL[id], QL[o1 * N + id], QL[o2 * N + id], QL[o3 * N + id], QL[4 * N + id]};
    cons_t r = {QR[id], QR[o1 * N + id], QR[o2 * N + id], QR[o3 * N + id], QR[4 * N + id]};
    double ql[ORC_MAXP], qr[ORC_MAXP], fq[ORC_MAXP];
    for (int i = 0; i < np; i++) {
      ql[i] = QL[(size_t)(5 + i) * N + id];
      qr[i] = QR[(size_t)(5 + i) * N + id];
    }
    pass_t ps = {np, de, ql, qr, fq};
    cons_t f = solver == ORC_ROE ? roe_flux_x(l, r, gamma, &ps)
                                 : (solver == ORC_HLL ? hll_flux_x(l, r, gamma, &ps) : hllc_flux_x(l, r, gamma, &ps));
    F[id] = f.d; F[o1 * N + id] = f.mx; F[o2 * N + id] = f.my; F[o3 * N + id] = f.mz; F[4 * N + id] = f.E;
    for (int i = 0; i < np; i++) F[(size_t)(5 + i) * N + id] = fq[i];
  }
}

/* one Van Leer step of a build with ns scalars and (de) dual energy: VL_{1,2,3}D_cuda.cu, hydro_cuda.cu:20-300,
 * 662-1099.  Returns -1 for combinations that are not restated. */
int orc_vl_step_x(double *U, int nx, int ny, int nz, int n_ghost, double dx, double dy, double dz, double dt,
                  double gamma, int recon, int solver, double density_floor, int ns, int de)
{
  if (recon < ORC_PCM || recon > ORC_PPMP) return -1;
  if (solver != ORC_HLLC && solver != ORC_ROE && solver != ORC_HLL) return -1;
  if (ns + de > ORC_MAXP) return -1;
  size_t N = (size_t)nx * ny * nz;
  int ndim = 1 + (ny > 1) + (nz > 1), nf = 5 + ns + de, GE = nf - 1;
  long st[3] = {1, nx, (long)nx * ny};
  double *Uh = calloc(nf * N, sizeof(double));
  double *QL[3], *QR[3], *F[3];
  for (int d = 0; d < ndim; d++) {
    QL[d] = calloc(nf * N, sizeof(double));
    QR[d] = calloc(nf * N, sizeof(double));
    F[d]  = calloc(nf * N, sizeof(double));
  }
  for (int d = 0; d < ndim; d++) {
    pcm_dir_x(U, QL[d], QR[d], nx, ny, nz, d, nf);
    riemann_x(solver, QL[d], QR[d], F[d], nx, ny, nz, gamma, d, ns, de);
  }
  { /* half-step update (VL_3D_cuda.cu:379-495, VL_2D_cuda.cu:178-260, VL_1D_cuda.cu:160-215) */
    double h = 0.5 * dt;
    double dto[3] = {h / dx, h / dy, h / dz};
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < nz; z++)
      for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) {
          if (!(x > 0 && x < nx - 1)) continue;
          if (ndim > 1 && !(y > 0 && y < ny - 1)) continue;
          if (ndim > 2 && !(z > 0 && z < nz - 1)) continue;
          size_t id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
          for (int f = 0; f < nf; f++) {
            double v = U[f * N + id];
            for (int d = 0; d < ndim; d++) v = v + dto[d] * (F[d][f * N + id - st[d]] - F[d][f * N + id]);
            Uh[f * N + id] = v;
          }
          if (de) {
            double d = U[id], d_inv = 1.0 / d;
            double vx = U[N + id] * d_inv, vy = U[2 * N + id] * d_inv, vz = U[3 * N + id] * d_inv, P;
            if (ndim == 3) { /* PRESSURE_DE, VL_3D_cuda.cu:408-426 */
              double E = U[4 * N + id], ge = U[(size_t)GE * N + id];
              double E_kin = 0.5 * d * square_mag(vx, vy, vz);
              P = fmax(de_pressure(E, E - E_kin, ge, gamma), ORC_TINY);
            } else { /* the 1-D and 2-D half updates use the plain pressure (VL_2D_cuda.cu:213, VL_1D_cuda.cu:183) */
              P = (U[4 * N + id] - 0.5 * d * (vx * vx + vy * vy + vz * vz)) * (gamma - 1.0);
            }
            double div = dto[0] * (U[1 * N + id - 1] / U[id - 1] - U[1 * N + id + 1] / U[id + 1]);
            for (int dd = 1; dd < ndim; dd++)
              div = div + dto[dd] * (U[(size_t)(1 + dd) * N + id - st[dd]] / U[id - st[dd]] -
                                     U[(size_t)(1 + dd) * N + id + st[dd]] / U[id + st[dd]]);
            Uh[(size_t)GE * N + id] = Uh[(size_t)GE * N + id] + 0.5 * P * (div);
          }
          if (ndim == 3 && Uh[id] < density_floor) { /* VL_3D_cuda.cu:479-493: momenta, Energy, GasEnergy; not the scalars */
            double d0 = Uh[id];
            Uh[id] = density_floor;
            for (int f = 1; f < 5; f++) Uh[f * N + id] *= (density_floor / d0);
            if (de) Uh[(size_t)GE * N + id] *= (density_floor / d0);
          }
        }
  }
  for (int d = 0; d < ndim; d++) {
    if (recon == ORC_PCM) pcm_dir_x(ndim == 2 ? U : Uh, QL[d], QR[d], nx, ny, nz, d, nf); /* VL_2D_cuda.cu:95-120 */
    else if (recon == ORC_PLMP || recon == ORC_PPMP) pxmp_dir_x(recon, Uh, QL[d], QR[d], nx, ny, nz, gamma, d, ns, de);
    else reconstruct_x(recon, Uh, QL[d], QR[d], nx, ny, nz, gamma, d, ns, de);
    riemann_x(solver, QL[d], QR[d], F[d], nx, ny, nz, gamma, d, ns, de);
  }
  double dto[3] = {dt / dx, dt / dy, dt / dz};
  int g = n_ghost;
#define ORC_REAL_CELL(x, y, z)                                                                      \
  ((x) > g - 1 && (x) < nx - g && (ndim < 2 || ((y) > g - 1 && (y) < ny - g)) && (ndim < 3 || ((z) > g - 1 && (z) < nz - g)))
  if (de) { /* Partial_Update_Advected_Internal_Energy_{1,2,3}D (hydro_cuda.cu:662-833): before the update, from U^n */
    double *gep = malloc(N * sizeof(double));
    memcpy(gep, U + (size_t)GE * N, N * sizeof(double));
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < nz; z++)
      for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) {
          if (!ORC_REAL_CELL(x, y, z)) continue;
          size_t id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
          double d = U[id], d_inv = 1.0 / d;
          double vx = U[N + id] * d_inv, vy = U[2 * N + id] * d_inv, vz = U[3 * N + id] * d_inv;
          double E = U[4 * N + id], ge = U[(size_t)GE * N + id];
          double E_kin = ndim == 3 ? 0.5 * d * square_mag(vx, vy, vz) : 0.5 * d * (vx * vx + vy * vy + vz * vz);
          double P = fmax(de_pressure(E, E - E_kin, ge, gamma), ORC_TINY);
          double div = dto[0] * (U[1 * N + id - 1] / U[id - 1] - U[1 * N + id + 1] / U[id + 1]);
          for (int dd = 1; dd < ndim; dd++)
            div = div + dto[dd] * (U[(size_t)(1 + dd) * N + id - st[dd]] / U[id - st[dd]] -
                                   U[(size_t)(1 + dd) * N + id + st[dd]] / U[id + st[dd]]);
          gep[id] = ge + 0.5 * P * (div);
        }
    memcpy(U + (size_t)GE * N, gep, N * sizeof(double));
    free(gep);
  }
#pragma omp parallel for collapse(2) schedule(static)
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++) {
        if (!ORC_REAL_CELL(x, y, z)) continue;
        size_t id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
        double d_old = U[id], m_old[3] = {U[N + id], U[2 * N + id], U[3 * N + id]};
        for (int f = 0; f < nf; f++) {
          double s = dto[0] * (F[0][f * N + id - st[0]] - F[0][f * N + id]);
          for (int d = 1; d < ndim; d++) s = s + dto[d] * (F[d][f * N + id - st[d]] - F[d][f * N + id]);
          U[f * N + id] += s;
        }
        if (ndim == 3 && U[id] < density_floor && U[id] > 0) { /* hydro_cuda.cu:279-291 */
          double d0 = U[id];
          U[id] = density_floor;
          for (int f = 1; f < 5; f++) U[f * N + id] *= (density_floor / d0);
          if (de) U[(size_t)GE * N + id] *= (density_floor / d0);
        }
        if (g_grav) gravity_source_cell(U, N, id, ndim, dt, d_old, m_old);
          if (g_pot && ndim == 3) gravity_potential_cell(U, N, id, st[1], st[2], dt, dx, dy, dz, d_old, m_old);
      }
  if (de) { /* Select_Internal_Energy_{1,2,3}D then Sync_Energies_{1,2,3}D (hydro_cuda.cu:835-1099) */
    double *sel = malloc(N * sizeof(double));
    int hi[3] = {nx - g - 1, ny - g - 1, nz - g - 1};
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < nz; z++)
      for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) {
          if (!ORC_REAL_CELL(x, y, z)) continue;
          size_t id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
          int c[3] = {x, y, z};
          double d = U[id], d_inv = 1.0 / d;
          double vx = U[N + id] * d_inv, vy = U[2 * N + id] * d_inv, vz = U[3 * N + id] * d_inv;
          double E = U[4 * N + id], U_advected = U[(size_t)GE * N + id];
          double U_total = E - 0.5 * d * (vx * vx + vy * vy + vz * vz);
          double Emax = E; /* neighbours clamped to the real cells; order imo, (E), ipo, jmo, jpo, kmo, kpo */
          for (int dd = 0; dd < ndim; dd++) {
            int lo = c[dd] - 1 > g ? c[dd] - 1 : g, up = c[dd] + 1 < hi[dd] ? c[dd] + 1 : hi[dd];
            size_t il = id + (size_t)((long)(lo - c[dd]) * st[dd]), iu = id + (size_t)((long)(up - c[dd]) * st[dd]);
            if (dd == 0) Emax = fmax(U[4 * N + il], E);
            else Emax = fmax(Emax, U[4 * N + il]);
            Emax = fmax(Emax, U[4 * N + iu]);
          }
          double Usel = ((U_total / E > ORC_DE_ETA_1) || (U_total / Emax > ORC_DE_ETA_2)) ? U_total : U_advected;
          sel[id] = fmax(Usel, ORC_TINY);
        }
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < nz; z++)
      for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) {
          if (!ORC_REAL_CELL(x, y, z)) continue;
          size_t id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
          double d = U[id], d_inv = 1.0 / d;
          double vx = U[N + id] * d_inv, vy = U[2 * N + id] * d_inv, vz = U[3 * N + id] * d_inv;
          U[(size_t)GE * N + id] = sel[id];
          U[4 * N + id] = 0.5 * d * (vx * vx + vy * vy + vz * vz) + sel[id];
        }
    free(sel);
  }
#undef ORC_REAL_CELL
  free(Uh);
  for (int d = 0; d < ndim; d++) { free(QL[d]); free(QR[d]); free(F[d]); }
  return 0;
}

/* Temperature_Floor_Kernel of a -DDE build (hydro_cuda.cu:1117-1153): the advected internal energy is floored too */
void orc_temperature_floor(double *U, int nx, int ny, int nz, int n_ghost, double U_floor);
void orc_temperature_floor_x(double *U, int nx, int ny, int nz, int n_ghost, double U_floor, int nf, int de)
{
  orc_temperature_floor(U, nx, ny, nz, n_ghost, U_floor);
  if (!de) return;
  size_t N = (size_t)nx * ny * nz;
  int g = n_ghost;
  for (int z = g; z < nz - g; z++)
    for (int y = g; y < ny - g; y++)
      for (int x = g; x < nx - g; x++) {
        size_t id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
        double d = U[id];
        if (U[(size_t)(nf - 1) * N + id] / d < U_floor) U[(size_t)(nf - 1) * N + id] = d * U_floor;
      }
}

/* hydro_cuda.cu:404-418 */
/* One step of the SIMPLE integrator (simple_{1,2,3}D_cuda.cu: reconstruct on U, Riemann, update over dt -- no
 * predictor; builds/make.type.hydro selects it for cosmology builds).  PCM and PLMC (with the characteristic
 * tracing the -DSIMPLE build of plmc_cuda.cu compiles in); the CTU forms of the other reconstructions are not restated. */
int orc_simple_step(double *U, int nx, int ny, int nz, int n_ghost, double dx, double dy, double dz, double dt,
                    double gamma, int recon, int solver, double density_floor)
{
  if (recon != ORC_PCM && recon != ORC_PLMC) return -1;
  size_t N = (size_t)nx * ny * nz;
  int ndim = 1 + (ny > 1) + (nz > 1);
  long st[3] = {1, nx, (long)nx * ny};
  double dto[3] = {dt / dx, dt / dy, dt / dz};
  double *QL[3], *QR[3], *F[3];
  for (int d = 0; d < ndim; d++) {
    QL[d] = calloc(5 * N, sizeof(double));
    QR[d] = calloc(5 * N, sizeof(double));
    F[d]  = calloc(5 * N, sizeof(double));
    if (recon == ORC_PCM) pcm_dir(U, QL[d], QR[d], nx, ny, nz, d);
    else reconstruct_t(recon, U, QL[d], QR[d], nx, ny, nz, gamma, d, 1, dto[d]);
    orc_riemann(solver, QL[d], QR[d], F[d], nx, ny, nz, gamma, d);
  }
  int g = n_ghost;
#pragma omp parallel for collapse(2) schedule(static)
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++) {
        if (!(x > g - 1 && x < nx - g)) continue;
        if (ndim > 1 && !(y > g - 1 && y < ny - g)) continue;
        if (ndim > 2 && !(z > g - 1 && z < nz - g)) continue;
        size_t id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
        for (int f = 0; f < 5; f++) {
          double s = dto[0] * (F[0][f * N + id - st[0]] - F[0][f * N + id]);
          for (int d = 1; d < ndim; d++) s = s + dto[d] * (F[d][f * N + id - st[d]] - F[d][f * N + id]);
          U[f * N + id] += s;
        }
        if (ndim == 3 && U[id] < density_floor && U[id] > 0) { /* hydro_cuda.cu:279-291 */
          double d0 = U[id];
          U[id] = density_floor;
          for (int f = 1; f < 5; f++) U[f * N + id] *= (density_floor / d0);
        }
      }
  for (int d = 0; d < ndim; d++) { free(QL[d]); free(QR[d]); free(F[d]); }
  return 0;
}

double orc_hydro_inverse_crossing_time(double E, double d, double d_inv, double vx, double vy, double vz, double dx,
                                       double dy, double dz, double gamma)
{
  double P  = (E - 0.5 * d * (vx * vx + vy * vy + vz * vz)) * (gamma - 1.0);
  double cs = sqrt(d_inv * gamma * P);
  double m = fmax((fabs(vx) + cs) / dx, (fabs(vy) + cs) / dy);
  m = fmax(m, (fabs(vz) + cs) / dz);
  m = fmax(m, 0.0);
  return m;
}

/* Calc_dt_{1,2,3}D (hydro_cuda.cu:437-560): max inverse crossing time over real cells */
double orc_calc_dti(const double *U, int nx, int ny, int nz, int n_ghost, double dx, double dy, double dz, double gamma)
{
  size_t N = (size_t)nx * ny * nz;
  int ndim = 1 + (ny > 1) + (nz > 1);
  if (nx == 1 && ny == 1 && nz == 1) ndim = 3; /* the reference unit test drives Calc_dt_3D on one cell */
  int g = n_ghost;
  double best = -DBL_MAX;
#pragma omp parallel for collapse(2) reduction(max : best) schedule(static)
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++) {
        if (!(x > g - 1 && x < nx - g)) continue;
        if (ndim > 1 && !(y > g - 1 && y < ny - g)) continue;
        if (ndim > 2 && !(z > g - 1 && z < nz - g)) continue;
        size_t id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
        double d = U[id], d_inv = 1.0 / d;
        double vx = U[N + id] * d_inv, vy = U[2 * N + id] * d_inv, vz = U[3 * N + id] * d_inv;
        double E = U[4 * N + id], v;
        if (ndim == 3) {
          v = orc_hydro_inverse_crossing_time(E, d, d_inv, vx, vy, vz, dx, dy, dz, gamma);
        } else {
          double P = (E - 0.5 * d * (vx * vx + vy * vy + vz * vz)) * (gamma - 1.0);
          P = fmax(P, ORC_TINY);
          double cs = sqrt(d_inv * gamma * P);
          v = (fabs(vx) + cs) / dx;
          if (ndim == 2) v = fmax(v, (fabs(vy) + cs) / dy);
        }
        best = fmax(best, v);
      }
  return best;
}

/* Temperature_Floor_Kernel (hydro_cuda.cu:1117-1153): the guard needs zid > n_ghost-1, so it is a
 * no-op when nz == 1 */
void orc_temperature_floor(double *U, int nx, int ny, int nz, int n_ghost, double U_floor)
{
  size_t N = (size_t)nx * ny * nz;
  int g = n_ghost;
#pragma omp parallel for collapse(2) schedule(static)
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++) {
        if (!(x > g - 1 && x < nx - g && y > g - 1 && y < ny - g && z > g - 1 && z < nz - g)) continue;
        size_t id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
        double d = U[id], d_inv = 1.0 / d;
        double vx = U[N + id] * d_inv, vy = U[2 * N + id] * d_inv, vz = U[3 * N + id] * d_inv;
        double E = U[4 * N + id];
        double Ekin = 0.5 * d * (vx * vx + vy * vy + vz * vz);
        double Ue = (E - Ekin) / d;
        if (Ue < U_floor) U[4 * N + id] = Ekin + d * U_floor;
      }
}

/* ---- ghost cells (cuda_boundaries.cu:71-303, boundary_conditions.cpp:419-) ---- */
static int map_index(int ig, int n, int flag, int face, int g, double *a)
{ /* FindIndex */
  if (face == 0) {
    switch (flag) {
      case 2: *a = -1.0; return 2 * g - ig - 1;
      case 3: return g;
      case 4: return -1;
      case 5: return ig;
      default: return ig + n - 2 * g; /* 1 and anything else: periodic */
    }
  }
  switch (flag) {
    case 2: *a = -1.0; return 2 * (n - g) - ig - 1;
    case 3: return n - g - 1;
    case 4: return -1;
    case 5: return ig;
    default: return ig - n + 2 * g;
  }
}

static long map_cell(int ig, int jg, int kg, double *a, const int *flags, int nx, int ny, int nz, int g)
{ /* SetBoundaryMapping */
  long idx = 0;
  int n[3] = {nx, ny, nz}, c[3] = {ig, jg, kg};
  long st[3] = {1, nx, (long)nx * ny};
  for (int d = 0; d < 3; d++) {
    if (n[d] <= 1) continue;
    int r;
    if (c[d] < g) r = map_index(c[d], n[d], flags[2 * d], 0, g, &a[d]);
    else if (c[d] >= n[d] - g) r = map_index(c[d], n[d], flags[2 * d + 1], 1, g, &a[d]);
    else r = c[d];
    if (r < 0) return -1;
    idx += st[d] * r;
  }
  return idx;
}

void orc_boundary_extents(int dir, int nx, int ny, int nz, int g, int *imin, int *imax)
{ /* Set_Boundary_Extents: x faces span real j,k; y faces all i, real k; z faces all i,j */
  int lo[3] = {0, 0, 0}, hi[3] = {nx, ny, nz};
  if (ny > 1) { lo[1] = g; hi[1] = ny - g; }
  if (nz > 1) { lo[2] = g; hi[2] = nz - g; }
  int n[3] = {nx, ny, nz};
  int ax = dir / 2;
  for (int d = 0; d < ax; d++) { lo[d] = 0; hi[d] = n[d]; }
  if (dir % 2 == 0) { lo[ax] = 0; hi[ax] = g; }
  else { lo[ax] = n[ax] - g; hi[ax] = n[ax]; }
  for (int d = 0; d < 3; d++) { imin[d] = lo[d]; imax[d] = hi[d]; }
}

/* the transmissive "diode" is compiled out of MHD builds (cuda_boundaries.cu:121-147, #ifndef MHD) */
static int g_diode = 1;
void orc_set_mhd_build(int mhd) { g_diode = !mhd; }

/* one call of Grid3D::Set_Boundaries(dir, flags) -> SetGhostCells */
void orc_set_ghost_cells(double *U, int nx, int ny, int nz, int n_fields, int g, const int *flags, int dir)
{
  if (flags[dir] == 4 || flags[dir] == 5) return;
  size_t N = (size_t)nx * ny * nz;
  int lo[3], hi[3];
  orc_boundary_extents(dir, nx, ny, nz, g, lo, hi);
  for (int k = lo[2]; k < hi[2]; k++)
    for (int j = lo[1]; j < hi[1]; j++)
      for (int i = lo[0]; i < hi[0]; i++) {
        double a[3] = {1., 1., 1.};
        size_t gidx = (size_t)i + (size_t)nx * ((size_t)j + (size_t)ny * k);
        long idx = map_cell(i, j, k, a, flags, nx, ny, nz, g);
        if (idx < 0) continue;
        for (int f = 0; f < n_fields; f++) U[gidx + f * N] = U[idx + f * N];
        if (flags[0] == 2 || flags[1] == 2) U[gidx + N] *= a[0];
        if (flags[2] == 2 || flags[3] == 2) U[gidx + 2 * N] *= a[1];
        if (flags[4] == 2 || flags[5] == 2) U[gidx + 3 * N] *= a[2];
        if (flags[dir] == 3 && g_diode) {
          size_t m = gidx + (dir / 2 + 1) * N;
          if ((dir % 2 == 0) ? (U[m] > 0.0) : (U[m] < 0.0)) {
            U[gidx + 4 * N] -= 0.5 * (U[m] * U[m]) / U[gidx];
            U[m] = 0.0;
          }
        }
      }
}

/* Grid3D::Set_Boundary_Conditions, non-MPI branch: faces 0..5 in order */
void orc_set_boundaries(double *U, int nx, int ny, int nz, int n_fields, int g, const int *flags)
{
  int ndim = 1 + (ny > 1) + (nz > 1);
  for (int dir = 0; dir < 2 * ndim; dir++) orc_set_ghost_cells(U, nx, ny, nz, n_fields, g, flags, dir);
}

/* PackBuffers3D / UnpackBuffers3D (cuda_boundaries.cu:11-69): buffer[field][k][j][i] */
void orc_pack(double *buf, const double *U, int nx, int ny, int n_fields, long n_cells, long idxoffset, int isize,
              int jsize, int ksize)
{
  long nb = (long)isize * jsize * ksize;
  for (long id = 0; id < nb; id++) {
    long k = id / ((long)isize * jsize), j = (id - k * isize * jsize) / isize, i = id - k * isize * jsize - j * isize;
    long idx = i + (j + k * ny) * nx + idxoffset;
    for (int f = 0; f < n_fields; f++) buf[id + f * nb] = U[idx + f * n_cells];
  }
}

void orc_unpack(const double *buf, double *U, int nx, int ny, int n_fields, long n_cells, long idxoffset, int isize,
                int jsize, int ksize)
{
  long nb = (long)isize * jsize * ksize;
  for (long id = 0; id < nb; id++) {
    long k = id / ((long)isize * jsize), j = (id - k * isize * jsize) / isize, i = id - k * isize * jsize - j * isize;
    long idx = i + (j + k * ny) * nx + idxoffset;
    for (int f = 0; f < n_fields; f++) U[idx + f * n_cells] = buf[id + f * nb];
  }
}

/* Scalar_Floor_Kernel (hydro_cuda.cu:1284-1307): real cells of a 3-D grid, one field */
void orc_scalar_floor(double *U, int nx, int ny, int nz, int n_ghost, int field_num, double scalar_floor)
{
  size_t N = (size_t)nx * ny * nz;
  int g = n_ghost;
  for (int z = g; z < nz - g; z++)
    for (int y = g; y < ny - g; y++)
      for (int x = g; x < nx - g; x++) {
        size_t id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
        if (U[(size_t)field_num * N + id] < scalar_floor) U[(size_t)field_num * N + id] = scalar_floor;
      }
}

/* Cooling_Update / cooling_kernel with CIE_cool (src/cooling/cooling_cuda.cu:20-196, 293-317): the -DCOOLING_GPU source
 * step.  PARITY UNPINNED ON THE CPU where log10 / pow are involved (device library bits); the bit-level check of the CUDA
 * path is the reference's own -DCOOLING_GPU build (tests/test_gpu_cooling.py). */
#define ORC_MP 1.672622e-24
#define ORC_KB 1.380658e-16
#define ORC_TIME_UNIT 3.15569e10
#define ORC_LENGTH_UNIT 3.08567758e21
#define ORC_MASS_UNIT 1.98847e33
#define ORC_DENSITY_UNIT (ORC_MASS_UNIT / (ORC_LENGTH_UNIT * ORC_LENGTH_UNIT * ORC_LENGTH_UNIT))
#define ORC_VELOCITY_UNIT (ORC_LENGTH_UNIT / ORC_TIME_UNIT)
#define ORC_ENERGY_UNIT (ORC_DENSITY_UNIT * ORC_VELOCITY_UNIT * ORC_VELOCITY_UNIT)
#define ORC_PRESSURE_UNIT (ORC_DENSITY_UNIT * ORC_VELOCITY_UNIT * ORC_VELOCITY_UNIT)
#define ORC_SP_ENERGY_UNIT (ORC_VELOCITY_UNIT * ORC_VELOCITY_UNIT)
double orc_cie_cool(double n, double T)
{
  double lambda = 0.0;
  if (log10(T) < 4.0) lambda = 0.0;
  else if (log10(T) >= 4.0 && log10(T) < 5.9) lambda = pow(10.0, (-1.3 * (log10(T) - 5.25) * (log10(T) - 5.25) - 21.25));
  else if (log10(T) >= 5.9 && log10(T) < 7.4) lambda = pow(10.0, (0.7 * (log10(T) - 7.1) * (log10(T) - 7.1) - 22.8));
  else lambda = pow(10.0, (0.45 * log10(T) - 26.065));
  return n * n * lambda;
}
void orc_cooling_update(double *U, int nx, int ny, int nz, int n_ghost, int n_fields, int de, double dt_in, double gamma)
{
  size_t N = (size_t)nx * ny * nz;
  int g = n_ghost;
  int js = ny == 1 ? 0 : g, je = ny == 1 ? 1 : ny - g, ks = nz == 1 ? 0 : g, ke = nz == 1 ? 1 : nz - g;
  double mu = 0.6;
  for (int z = ks; z < ke; z++)
    for (int y = js; y < je; y++)
      for (int x = g; x < nx - g; x++) {
        size_t id = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z);
        double dt = dt_in, d = U[id], E = U[4 * N + id], ge = 0.0;
        if (E < 0.0 || E != E) continue;
        double vx = U[N + id] / d, vy = U[2 * N + id] / d, vz = U[3 * N + id] / d;
        double p = (E - 0.5 * d * (vx * vx + vy * vy + vz * vz)) * (gamma - 1.0);
        p = fmax(p, ORC_TINY);
        if (de) { ge = U[(size_t)(n_fields - 1) * N + id] / d; ge = fmax(ge, ORC_TINY); }
        double n = d * ORC_DENSITY_UNIT / (mu * ORC_MP);
        double T_init = p * ORC_PRESSURE_UNIT / (n * ORC_KB);
        if (de) T_init = d * ge * (gamma - 1.0) * ORC_PRESSURE_UNIT / (n * ORC_KB);
        double T = T_init, cool = orc_cie_cool(n, T);
        double del_T = cool * dt * ORC_TIME_UNIT * (gamma - 1.0) / (n * ORC_KB);
        while (del_T / T > 0.01) {
          double dt_sub = 0.01 * T * n * ORC_KB / (cool * ORC_TIME_UNIT * (gamma - 1.0));
          T -= cool * dt_sub * ORC_TIME_UNIT * (gamma - 1.0) / (n * ORC_KB);
          dt -= dt_sub;
          cool = orc_cie_cool(n, T);
          del_T = cool * dt * ORC_TIME_UNIT * (gamma - 1.0) / (n * ORC_KB);
        }
        T -= del_T;
        del_T = T_init - T;
        E -= n * ORC_KB * del_T / ((gamma - 1.0) * ORC_ENERGY_UNIT);
        U[4 * N + id] = E;
        if (de) {
          ge -= ORC_KB * del_T / (mu * ORC_MP * (gamma - 1.0) * ORC_SP_ENERGY_UNIT);
          U[(size_t)(n_fields - 1) * N + id] = d * ge;
        }
      }
}
