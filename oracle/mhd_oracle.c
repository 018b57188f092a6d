/*
 * mhd_oracle.c -- CPU restatement of Cholla's Van Leer MHD hot path (3-D, 8 fields, HLLD + constrained transport).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the shipped product (cholla_b200/, the C-ABI library) links, imports or
 * calls this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may.
 *
 * It restates, kernel by kernel, with the reference's guards, array shapes and floating-point operation order
 * (paths relative to the reference tree, commit 8478c9e):
 *   src/integrators/VL_3D_cuda.cu:38-360          the MHD branch of the six-stage driver (scratch zero-initialised
 *                                                 once and kept between steps, like the reference's globals)
 *   src/reconstruction/pcm_cuda.cu:256-424        PCM_Reconstruction_3D (MHD)
 *   src/reconstruction/plmc_cuda.cu:24-300, ppmc_cuda.cu:540-695, reconstruction.h (MHD branches:
 *                                                 Load_Data, Compute_Eigenvectors, Primitive_To_Characteristic,
 *                                                 Characteristic_To_Primitive, Write_Data)
 *   src/riemann_solvers/hlld_cuda.cu, hlld_cuda.h HLLD (Miyoshi & Kusano 2005)
 *   src/mhd/ct_electric_fields.{h,cu}             CT edge EMFs (Stone & Gardiner 2009 eqs. 22-24)
 *   src/mhd/magnetic_update.cu                    face-centred induction update
 *   src/hydro/hydro_cuda.cu:172-300, 420-435, 514-560   conserved update, mhdInverseCrossingTime, Calc_dt_3D
 *   src/utils/mhd_utilities.h, hydro_utilities.h, math_utilities.h
 * Build with -O2 -ffp-contract=off (the reference device code is built -fmad=false, Makefile:148).
 *
 * Two CUDA math functions the reference calls have no correctly-rounded libm twin: rsqrt() (reconstruction.h:387)
 * and rhypot() (reconstruction.h:342).  They are restated as 1/sqrt() and 1/hypot(); results can differ from the
 * device's by an ulp in the MHD reconstruction, so CPU-vs-GPU MHD parity is a tolerance (1e-12 relative L1), while
 * bit-for-bit MHD parity is checked on the GPU against the reference's own kernels (oracle/_ref, -DMHD build).
 *
 * Parity pin: the reference's own unit-test fiducials for HLLD (hlld_cuda_tests.cu), CT electric fields
 * (ct_electric_fields_tests.cu), the magnetic update (magnetic_update_tests.cu), MHD PLMC/PPMC
 * (plmc_cuda_tests.cu, ppmc_cuda_tests.cu), mhdInverseCrossingTime (hydro_cuda_tests.cu) and the mhd utilities
 * (mhd_utilities_tests.cu), extracted to tests/golden/mhd_*.json and checked by tests/test_oracle_mhd_golden.py.
 *
 * Layout is the reference's: SoA, field stride N = nx*ny*nz, id = x + nx*(y + ny*z); fields density, momentum
 * x/y/z, Energy, magnetic x/y/z (grid_enum.h:21-78); magnetic_x[id] lives on the +x face of cell id.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TINY 1.0e-20       /* global.h:25 */
#define HLLD_SMALL 1.0e-8  /* hlld_cuda.h:35 */
enum { F_D = 0, F_MX = 1, F_MY = 2, F_MZ = 3, F_E = 4, F_BX = 5, F_BY = 6, F_BZ = 7, NF = 8, NQ = 7 };
enum { MORC_PCM = 0, MORC_PLMC = 1, MORC_PPMC = 2 };

/* math_utilities.h:78-98 */
static inline double dot3(double a1, double a2, double a3, double b1, double b2, double b3)
{
  return a1 * b1 + ((a2 * b2) + (a3 * b3));
}
static inline double sqmag(double a, double b, double c) { return dot3(a, b, c, a, b, c); }
/* mhd_utilities.h:75-79 */
static inline double mag_energy(double bx, double by, double bz) { return 0.5 * sqmag(bx, by, bz); }
/* mhd_utilities.h:118-124 */
static inline double total_pressure(double p, double bx, double by, double bz)
{
  return fmax(p + mag_energy(bx, by, bz), TINY);
}
/* mhd_utilities.h:48-63 */
static double magnetosonic(double d, double p, double bx, double by, double bz, double gamma, double wave)
{
  double const bx2 = bx * bx;
  double const b2  = bx2 + ((by * by) + (bz * bz));
  double const t1  = gamma * p + b2;
  double t2        = (t1 * t1) - 4. * gamma * p * bx2;
  t2               = sqrt(t2);
  return sqrt((t1 + wave * t2) / (2.0 * fmax(d, TINY)));
}
double morc_fast_speed(double d, double p, double bx, double by, double bz, double gamma)
{
  return magnetosonic(d, p, bx, by, bz, gamma, 1.0);
}
double morc_slow_speed(double d, double p, double bx, double by, double bz, double gamma)
{
  return magnetosonic(d, p, bx, by, bz, gamma, -1.0);
}
/* mhd_utilities.h:176-180 */
double morc_alfven_speed(double bx, double d) { return fabs(bx) / sqrt(fmax(d, TINY)); }
/* hydro_utilities.h:34-45 (MHD) */
static inline double pressure_primitive(double E, double d, double vx, double vy, double vz, double gamma, double bx,
                                        double by, double bz)
{
  double p = E - 0.5 * d * sqmag(vx, vy, vz);
  p -= mag_energy(bx, by, bz);
  return fmax((gamma - 1.) * p, TINY);
}
/* hydro_utilities.h:106-118 (MHD) */
static inline double energy_primitive(double P, double d, double vx, double vy, double vz, double gamma, double bx,
                                      double by, double bz)
{
  double e = (fmax(P, TINY) / (gamma - 1.)) + 0.5 * d * sqmag(vx, vy, vz);
  e += mag_energy(bx, by, bz);
  return e;
}

/* hooks for the reference's known answers of these two (hydro_utilities_tests.cpp:61-77, 155-171, 191-206; with zero
 * fields they are the hydro build's formulas) */
double morc_test_pressure_primitive(double E, double d, double vx, double vy, double vz, double gamma, double bx, double by,
                                    double bz)
{ return pressure_primitive(E, d, vx, vy, vz, gamma, bx, by, bz); }
double morc_test_energy_primitive(double P, double d, double vx, double vy, double vz, double gamma, double bx, double by,
                                  double bz)
{ return energy_primitive(P, d, vx, vy, vz, gamma, bx, by, bz); }

/* mhd_utilities.h:205-236 */
static void cell_centred_b(const double *U, size_t id, int x, int y, int z, size_t N, int nx, int ny, double *bx,
                           double *by, double *bz)
{
  *bx = (x > 0) ? 0.5 * (U[F_BX * N + id] + U[F_BX * N + id - 1]) : U[F_BX * N + id];
  *by = (y > 0) ? 0.5 * (U[F_BY * N + id] + U[F_BY * N + id - (size_t)nx]) : U[F_BY * N + id];
  *bz = (z > 0) ? 0.5 * (U[F_BZ * N + id] + U[F_BZ * N + id - (size_t)nx * ny]) : U[F_BZ * N + id];
}
void morc_cell_centred_b(const double *U, int nx, int ny, int nz, int x, int y, int z, double *out3)
{
  size_t const N = (size_t)nx * ny * nz, id = x + (size_t)nx * (y + (size_t)ny * z);
  cell_centred_b(U, id, x, y, z, N, nx, ny, out3, out3 + 1, out3 + 2);
}

static inline void dir_perm(int dir, int *o1, int *o2, int *o3)
{ /* plmc_cuda.cu:41-58 */
  if (dir == 0) { *o1 = 1; *o2 = 2; *o3 = 3; }
  else if (dir == 1) { *o1 = 2; *o2 = 3; *o3 = 1; }
  else { *o1 = 3; *o2 = 1; *o3 = 2; }
}

/* ------------------------------------------------------------------------------------------------------------
 * reconstruction
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  double d, vx, vy, vz, p, bx, by, bz; /* sweep frame; bx,by,bz cell centred */
} mprim_t;
typedef struct {
  double a[7];
} mchar_t;
typedef struct {
  double cf, cs, cf2, cs2, alpha_f, alpha_s, beta_y, beta_z, n_fs, sign, q_f, q_s, a_f, a_s, qp_f, qp_s, ap_f, ap_s;
} eigen_t;

/* reconstruction.h:139-212 (MHD, no DE) */
static mprim_t load_prim(const double *U, int x, int y, int z, int nx, int ny, size_t N, int o1, int o2, int o3,
                         double gamma)
{
  size_t const id = x + (size_t)nx * (y + (size_t)ny * z);
  mprim_t w;
  w.d  = U[id];
  w.vx = U[o1 * N + id] / w.d;
  w.vy = U[o2 * N + id] / w.d;
  w.vz = U[o3 * N + id] / w.d;
  double cx, cy, cz;
  cell_centred_b(U, id, x, y, z, N, nx, ny, &cx, &cy, &cz);
  if (o1 == 1) { w.bx = cx; w.by = cy; w.bz = cz; }
  else if (o1 == 2) { w.bx = cy; w.by = cz; w.bz = cx; }
  else { w.bx = cz; w.by = cx; w.bz = cy; }
  w.p = pressure_primitive(U[F_E * N + id], w.d, w.vx, w.vy, w.vz, gamma, w.bx, w.by, w.bz);
  return w;
}

/* reconstruction.h:918-968 (MHD): the normal field at the face the state sits on comes from the grid */
static void write_interface(mprim_t q, double *Q, const double *U, size_t id, size_t N, int o1, int o2, int o3,
                            double gamma)
{
  Q[id]          = q.d;
  Q[o1 * N + id] = q.d * q.vx;
  Q[o2 * N + id] = q.d * q.vy;
  Q[o3 * N + id] = q.d * q.vz;
  Q[5 * N + id]  = q.by;
  Q[6 * N + id]  = q.bz;
  double const bface = U[(F_BX + (o1 - 1)) * N + id];
  Q[F_E * N + id]    = energy_primitive(q.p, q.d, q.vx, q.vy, q.vz, gamma, bface, q.by, q.bz);
}

/* reconstruction.h:310-361 */
static eigen_t eigenvectors(mprim_t w, double a, double a2, double gamma)
{
  eigen_t e;
  e.cf  = morc_fast_speed(w.d, w.p, w.bx, w.by, w.bz, gamma);
  e.cs  = morc_slow_speed(w.d, w.p, w.bx, w.by, w.bz, gamma);
  e.cf2 = e.cf * e.cf;
  e.cs2 = e.cs * e.cs;
  double const denom = e.cf2 - e.cs2, num2 = e.cf2 - a2;
  if (denom <= 0.0 || num2 <= 0.0) {
    e.alpha_f = 1.0;
    e.alpha_s = 0.0;
  } else {
    double const num1 = a2 - e.cs2;
    if (num1 <= 0.0) {
      e.alpha_f = 0.0;
      e.alpha_s = 1.0;
    } else {
      e.alpha_f = sqrt(num1 / denom);
      e.alpha_s = sqrt(num2 / denom);
    }
  }
  double const bd = 1.0 / hypot(w.by, w.bz); /* rhypot */
  e.beta_y        = isfinite(bd) ? w.by * bd : 1.0;
  e.beta_z        = isfinite(bd) ? w.bz * bd : 0.0;
  e.sign          = copysign(1.0, w.bx);
  e.n_fs          = 0.5 / a2;
  e.qp_f          = e.sign * e.n_fs * e.alpha_f * e.cf;
  e.qp_s          = e.sign * e.n_fs * e.alpha_s * e.cs;
  e.q_f           = e.sign * e.alpha_f * e.cf;
  e.q_s           = e.sign * e.alpha_s * e.cs;
  e.a_f           = e.alpha_f * a * sqrt(w.d);
  e.a_s           = e.alpha_s * a * sqrt(w.d);
  e.ap_f          = 0.5 * e.alpha_f / (a * sqrt(w.d));
  e.ap_s          = 0.5 * e.alpha_s / (a * sqrt(w.d));
  return e;
}

/* reconstruction.h:378-421 (MHD) */
static mchar_t to_char(mprim_t c, mprim_t s, eigen_t e, double a2)
{
  mchar_t o;
  double const isd = 1.0 / sqrt(c.d); /* rsqrt */
  o.a[0] = e.n_fs * e.alpha_f * (s.p / c.d - e.cf * s.vx) + e.qp_s * (e.beta_y * s.vy + e.beta_z * s.vz) +
           e.ap_s * (e.beta_y * s.by + e.beta_z * s.bz);
  o.a[1] = 0.5 * (e.beta_y * (s.bz * e.sign * isd + s.vz) - e.beta_z * (s.by * e.sign * isd + s.vy));
  o.a[2] = e.n_fs * e.alpha_s * (s.p / c.d - e.cs * s.vx) - e.qp_f * (e.beta_y * s.vy + e.beta_z * s.vz) -
           e.ap_f * (e.beta_y * s.by + e.beta_z * s.bz);
  o.a[3] = s.d - s.p / a2;
  o.a[4] = e.n_fs * e.alpha_s * (s.p / c.d + e.cs * s.vx) + e.qp_f * (e.beta_y * s.vy + e.beta_z * s.vz) -
           e.ap_f * (e.beta_y * s.by + e.beta_z * s.bz);
  o.a[5] = 0.5 * (e.beta_y * (s.bz * e.sign * isd - s.vz) - e.beta_z * (s.by * e.sign * isd - s.vy));
  o.a[6] = e.n_fs * e.alpha_f * (s.p / c.d + e.cf * s.vx) - e.qp_s * (e.beta_y * s.vy + e.beta_z * s.vz) +
           e.ap_s * (e.beta_y * s.by + e.beta_z * s.bz);
  return o;
}

/* reconstruction.h:450-481 (MHD) */
static mprim_t from_char(mprim_t c, mchar_t s, eigen_t e, double a2)
{
  mprim_t o;
  const double *a = s.a;
  o.d  = c.d * (e.alpha_f * (a[0] + a[6]) + e.alpha_s * (a[2] + a[4])) + a[3];
  o.vx = e.cf * e.alpha_f * (a[6] - a[0]) + e.cs * e.alpha_s * (a[4] - a[2]);
  o.vy = e.beta_y * (e.q_s * (a[0] - a[6]) + e.q_f * (a[4] - a[2])) + e.beta_z * (a[5] - a[1]);
  o.vz = e.beta_z * (e.q_s * (a[0] - a[6]) + e.q_f * (a[4] - a[2])) + e.beta_y * (a[1] - a[5]);
  o.p  = c.d * a2 * (e.alpha_f * (a[0] + a[6]) + e.alpha_s * (a[2] + a[4]));
  o.by = e.beta_y * (e.a_s * (a[0] + a[6]) - e.a_f * (a[2] + a[4])) - e.beta_z * e.sign * sqrt(c.d) * (a[5] + a[1]);
  o.bz = e.beta_z * (e.a_s * (a[0] + a[6]) - e.a_f * (a[2] + a[4])) + e.beta_y * e.sign * sqrt(c.d) * (a[5] + a[1]);
  o.bx = 0.0;
  return o;
}

static inline double vl_slope(double l, double r) { return (l * r > 0.0) ? 2.0 * l * r / (l + r) : 0.0; }
static inline double mono(double l, double r, double c, double g)
{
  if (l * r > 0.0) {
    double const la = 2.0 * fmin(fabs(l), fabs(r));
    double const lb = fmin(fabs(c), fabs(g));
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
static mprim_t pdiff(mprim_t l, mprim_t r, double coef)
{
  mprim_t s;
  s.d  = coef * (r.d - l.d);
  s.vx = coef * (r.vx - l.vx);
  s.vy = coef * (r.vy - l.vy);
  s.vz = coef * (r.vz - l.vz);
  s.p  = coef * (r.p - l.p);
  s.by = coef * (r.by - l.by);
  s.bz = coef * (r.bz - l.bz);
  s.bx = 0.0;
  return s;
}

/* plmc_cuda.cu:60-128, 287-300 (VL build: no characteristic tracing) */
void morc_plmc_cell(mprim_t wm, mprim_t wi, mprim_t wp, double gamma, mprim_t *L, mprim_t *R)
{
  double const a = sqrt(gamma * wi.p / wi.d), a2 = a * a;
  eigen_t const e  = eigenvectors(wi, a, a2, gamma);
  mprim_t const dL = pdiff(wm, wi, 1.0), dR = pdiff(wi, wp, 1.0), dC = pdiff(wm, wp, 0.5);
  mprim_t dG;
  dG.d  = vl_slope(dL.d, dR.d);
  dG.vx = vl_slope(dL.vx, dR.vx);
  dG.vy = vl_slope(dL.vy, dR.vy);
  dG.vz = vl_slope(dL.vz, dR.vz);
  dG.p  = vl_slope(dL.p, dR.p);
  dG.by = vl_slope(dL.by, dR.by);
  dG.bz = vl_slope(dL.bz, dR.bz);
  dG.bx = 0.0;
  mchar_t const aL = to_char(wi, dL, e, a2), aR = to_char(wi, dR, e, a2), aC = to_char(wi, dC, e, a2),
                aG = to_char(wi, dG, e, a2);
  mchar_t am;
  for (int k = 0; k < 7; k++) am.a[k] = mono(aL.a[k], aR.a[k], aC.a[k], aG.a[k]);
  mprim_t const dm = from_char(wi, am, e, a2);
  mprim_t l, r;
  l.d  = wi.d + 1.0 * 0.5 * dm.d;    r.d  = wi.d + -1.0 * 0.5 * dm.d;
  l.vx = wi.vx + 1.0 * 0.5 * dm.vx;  r.vx = wi.vx + -1.0 * 0.5 * dm.vx;
  l.vy = wi.vy + 1.0 * 0.5 * dm.vy;  r.vy = wi.vy + -1.0 * 0.5 * dm.vy;
  l.vz = wi.vz + 1.0 * 0.5 * dm.vz;  r.vz = wi.vz + -1.0 * 0.5 * dm.vz;
  l.p  = wi.p + 1.0 * 0.5 * dm.p;    r.p  = wi.p + -1.0 * 0.5 * dm.p;
  l.by = wi.by + 1.0 * 0.5 * dm.by;  r.by = wi.by + -1.0 * 0.5 * dm.by;
  l.bz = wi.bz + 1.0 * 0.5 * dm.bz;  r.bz = wi.bz + -1.0 * 0.5 * dm.bz;
  l.bx = r.bx = 0.0;
  clamp_pair(&l.d, &r.d, wm.d, wi.d, wp.d);
  clamp_pair(&l.vx, &r.vx, wm.vx, wi.vx, wp.vx);
  clamp_pair(&l.vy, &r.vy, wm.vy, wi.vy, wp.vy);
  clamp_pair(&l.vz, &r.vz, wm.vz, wi.vz, wp.vz);
  clamp_pair(&l.p, &r.p, wm.p, wi.p, wp.p);
  clamp_pair(&l.by, &r.by, wm.by, wi.by, wp.by);
  clamp_pair(&l.bz, &r.bz, wm.bz, wi.bz, wp.bz);
  /* plmc_cuda.cu:287-291: floors applied after the (dead, under VL) characteristic tracing block */
  r.d = fmax(r.d, TINY);
  l.d = fmax(l.d, TINY);
  r.p = fmax(r.p, TINY);
  l.p = fmax(l.p, TINY);
  *L = l;
  *R = r;
}

/* PPM_Single_Variable (reconstruction.h:794-901): the pinned restatement in vl_oracle.c, same library */
extern void orc_ppm_single_variable(double cm2, double cm1, double c0, double cp1, double cp2, double *L_iph,
                                    double *R_imh);
#define ppm_single orc_ppm_single_variable

/* ppmc_cuda.cu:576-684 */
void morc_ppmc_cell(mprim_t wm2, mprim_t wm1, mprim_t wi, mprim_t wp1, mprim_t wp2, double gamma, mprim_t *L,
                    mprim_t *R)
{
  double const a = sqrt(gamma * wi.p / wi.d), a2 = a * a;
  eigen_t const e  = eigenvectors(wi, a, a2, gamma);
  mchar_t const c0 = to_char(wi, wi, e, a2), cm1 = to_char(wi, wm1, e, a2), cm2 = to_char(wi, wm2, e, a2),
                cp1 = to_char(wi, wp1, e, a2), cp2 = to_char(wi, wp2, e, a2);
  mchar_t cl, cr;
  for (int k = 0; k < 7; k++) ppm_single(cm2.a[k], cm1.a[k], c0.a[k], cp1.a[k], cp2.a[k], &cl.a[k], &cr.a[k]);
  mprim_t l = from_char(wi, cl, e, a2), r = from_char(wi, cr, e, a2);
  r.d = fmax(r.d, TINY);
  l.d = fmax(l.d, TINY);
  r.p = fmax(r.p, TINY);
  l.p = fmax(l.p, TINY);
  *L = l;
  *R = r;
}

/* entry for the PLMC/PPMC fiducial tests: one launch of PLMC_cuda / PPMC_VL on an 8-field grid */
void morc_reconstruct(int recon, const double *U, double *QL, double *QR, int nx, int ny, int nz, double gamma, int dir)
{
  size_t const N = (size_t)nx * ny * nz;
  int o1, o2, o3;
  dir_perm(dir, &o1, &o2, &o3);
  int const order = recon == MORC_PPMC ? 3 : 2;
  int const sx = dir == 0, sy = dir == 1, sz = dir == 2;
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++) {
        /* Thread_Guard<order>, reconstruction.h:104-126 */
        int out = x < order - 1 || x >= nx - order;
        if (ny > 1) out = out || y < order - 1 || y >= ny - order;
        if (nz > 1) out = out || z < order - 1 || z >= nz - order;
        else out = out || z >= nz;
        if (out) continue;
        mprim_t const wi  = load_prim(U, x, y, z, nx, ny, N, o1, o2, o3, gamma);
        mprim_t const wm1 = load_prim(U, x - sx, y - sy, z - sz, nx, ny, N, o1, o2, o3, gamma);
        mprim_t const wp1 = load_prim(U, x + sx, y + sy, z + sz, nx, ny, N, o1, o2, o3, gamma);
        mprim_t L, R;
        if (recon == MORC_PPMC) {
          mprim_t const wm2 = load_prim(U, x - 2 * sx, y - 2 * sy, z - 2 * sz, nx, ny, N, o1, o2, o3, gamma);
          mprim_t const wp2 = load_prim(U, x + 2 * sx, y + 2 * sy, z + 2 * sz, nx, ny, N, o1, o2, o3, gamma);
          morc_ppmc_cell(wm2, wm1, wi, wp1, wp2, gamma, &L, &R);
        } else {
          morc_plmc_cell(wm1, wi, wp1, gamma, &L, &R);
        }
        size_t const id = x + (size_t)nx * (y + (size_t)ny * z);
        write_interface(L, QL, U, id, N, o1, o2, o3, gamma);
        size_t const idm = (x - sx) + (size_t)nx * ((y - sy) + (size_t)ny * (z - sz));
        write_interface(R, QR, U, idm, N, o1, o2, o3, gamma);
      }
}

/* pcm_cuda.cu:256-424 (MHD): the three directions at once */
static void pcm_3d(const double *U, double *QL[3], double *QR[3], int nx, int ny, int nz)
{
  size_t const N = (size_t)nx * ny * nz;
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++) {
        size_t const id = x + (size_t)nx * (y + (size_t)ny * z);
        double b[3];
        cell_centred_b(U, id, x, y, z, N, nx, ny, &b[0], &b[1], &b[2]);
        int const c[3]      = {x, y, z};
        size_t const str[3] = {1, (size_t)nx, (size_t)nx * ny};
        for (int d = 0; d < 3; d++) {
          for (int f = 0; f < 5; f++) QL[d][f * N + id] = U[f * N + id];
          QL[d][5 * N + id] = b[(d + 1) % 3];
          QL[d][6 * N + id] = b[(d + 2) % 3];
          if (c[d] > 0) {
            size_t const im = id - str[d];
            for (int f = 0; f < 5; f++) QR[d][f * N + im] = U[f * N + id];
            QR[d][5 * N + im] = b[(d + 1) % 3];
            QR[d][6 * N + im] = b[(d + 2) % 3];
          }
        }
      }
}

/* ------------------------------------------------------------------------------------------------------------
 * HLLD (hlld_cuda.cu)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  double d, vx, vy, vz, E, by, bz, pg, pt;
} hstate_t;
typedef struct {
  double d, vy, vz, E, by, bz;
} hstar_t;
typedef struct {
  double d, mx, my, mz, E, by, bz;
} hflux_t;

/* loadState, hlld_cuda.cu:166-207; q = (d, m1, m2, m3, E, By, Bz) in the sweep frame */
static hstate_t hlld_load(const double q[7], double bx, double gamma)
{
  hstate_t s;
  s.d  = fmax(q[0], TINY);
  s.vx = q[1] / s.d;
  s.vy = q[2] / s.d;
  s.vz = q[3] / s.d;
  s.E  = fmax(q[4], TINY);
  s.by = q[5];
  s.bz = q[6];
  s.pg = pressure_primitive(s.E, s.d, s.vx, s.vy, s.vz, gamma, bx, s.by, s.bz);
  s.pt = total_pressure(s.pg, bx, s.by, s.bz);
  return s;
}
/* nonStarFluxes, hlld_cuda.cu:264-285 */
static hflux_t hlld_nonstar(hstate_t s, double bx)
{
  hflux_t f;
  f.d  = s.d * s.vx;
  f.mx = f.d * s.vx + s.pt - bx * bx;
  f.my = f.d * s.vy - bx * s.by;
  f.mz = f.d * s.vz - bx * s.bz;
  f.by = s.by * s.vx - bx * s.vy;
  f.bz = s.bz * s.vx - bx * s.vz;
  f.E  = s.vx * (s.E + s.pt) - bx * (s.vx * bx + ((s.vy * s.by) + (s.vz * s.bz)));
  return f;
}
/* computeStarState, hlld_cuda.cu:335-382 */
static hstar_t hlld_star(hstate_t s, double S, double SM, double bx, double pts)
{
  hstar_t o;
  o.d = s.d * (S - s.vx) / (S - SM);
  if (fabs(s.d * (S - s.vx) * (S - SM) - (bx * bx)) < pts * HLLD_SMALL) {
    o.vy = s.vy;
    o.vz = s.vz;
    o.by = s.by;
    o.bz = s.bz;
  } else {
    double const denom = s.d * (S - s.vx) * (S - SM) - (bx * bx);
    double coef        = bx * (SM - s.vx) / denom;
    o.vy               = s.vy - s.by * coef;
    o.vz               = s.vz - s.bz * coef;
    double t           = (S - s.vx);
    t                  = t * t;
    coef               = (s.d * t - (bx * bx)) / denom;
    o.by               = s.by * coef;
    o.bz               = s.bz * coef;
  }
  o.E = (s.E * (S - s.vx) - s.pt * s.vx + pts * SM +
         bx * (dot3(s.vx, s.vy, s.vz, bx, s.by, s.bz) - dot3(SM, o.vy, o.vz, bx, o.by, o.bz))) /
        (S - SM);
  return o;
}
/* starFluxes, hlld_cuda.cu:387-407 */
static hflux_t hlld_star_flux(hstar_t st, hstate_t s, hflux_t f, double SM, double S)
{
  hflux_t o;
  o.d  = f.d + S * (st.d - s.d);
  o.mx = f.mx + S * (st.d * SM - s.d * s.vx);
  o.my = f.my + S * (st.d * st.vy - s.d * s.vy);
  o.mz = f.mz + S * (st.d * st.vz - s.d * s.vz);
  o.E  = f.E + S * (st.E - s.E);
  o.by = f.by + S * (st.by - s.by);
  o.bz = f.bz + S * (st.bz - s.bz);
  return o;
}
/* computeDoubleStarFluxes, hlld_cuda.cu:479-507 */
static hflux_t hlld_dstar_flux(double vyd, double vzd, double byd, double bzd, double Ed, hstar_t st, hstate_t s,
                               hflux_t f, double SM, double S, double Sstar)
{
  hflux_t o;
  double const sd = Sstar - S;
  o.d  = f.d - S * s.d - sd * st.d + Sstar * st.d;
  o.mx = f.mx - S * (s.d * s.vx) - sd * (st.d * SM) + Sstar * (st.d * SM);
  o.my = f.my - S * (s.d * s.vy) - sd * (st.d * st.vy) + Sstar * (st.d * vyd);
  o.mz = f.mz - S * (s.d * s.vz) - sd * (st.d * st.vz) + Sstar * (st.d * vzd);
  o.E  = f.E - S * s.E - sd * st.E + Sstar * Ed;
  o.by = f.by - S * s.by - sd * st.by + Sstar * byd;
  o.bz = f.bz - S * s.bz - sd * st.bz + Sstar * bzd;
  return o;
}

/* Calculate_HLLD_Fluxes_CUDA for one interface, hlld_cuda.cu:31-161.  Returns 0 when no branch fires (NaN input):
 * the reference kernel then leaves the flux array untouched. */
int morc_hlld_single(const double ql[7], const double qr[7], double bx, double gamma, double out[7])
{
  hstate_t const l = hlld_load(ql, bx, gamma), r = hlld_load(qr, bx, gamma);
  double const cfl = morc_fast_speed(l.d, l.pg, bx, l.by, l.bz, gamma);
  double const cfr = morc_fast_speed(r.d, r.pg, bx, r.by, r.bz, gamma);
  double const cmax = fmax(cfl, cfr);
  double const SL = fmin(l.vx, r.vx) - cmax, SR = fmax(l.vx, r.vx) + cmax;
  hflux_t fl = hlld_nonstar(l, bx), f;
  int done   = 0;
  if (SL > 0.0) {
    f    = fl;
    done = 1;
  }
  hflux_t fr;
  if (!done) {
    fr = hlld_nonstar(r, bx);
    if (SR < 0.0) {
      f    = fr;
      done = 1;
    }
  }
  if (!done) {
    double const rdiff = SR - r.vx, ldiff = SL - l.vx;
    double const SM = (rdiff * r.d * r.vx - ldiff * l.d * l.vx - r.pt + l.pt) / (rdiff * r.d - ldiff * l.d);
    double const pts = (r.d * l.pt * (SR - r.vx) - l.d * r.pt * (SL - l.vx) +
                        l.d * r.d * (SR - r.vx) * (SL - l.vx) * (r.vx - l.vx)) /
                       (r.d * (SR - r.vx) - l.d * (SL - l.vx));
    hstar_t const sl  = hlld_star(l, SL, SM, bx, pts);
    double const SLs  = SM + -1 * morc_alfven_speed(bx, sl.d);
    if (SLs > 0.0 && SL <= 0.0) {
      f    = hlld_star_flux(sl, l, fl, SM, SL);
      done = 1;
    }
    if (!done) {
      hstar_t const sr = hlld_star(r, SR, SM, bx, pts);
      double const SRs = SM + 1 * morc_alfven_speed(bx, sr.d);
      if (SRs <= 0.0 && SR >= 0.0) {
        f    = hlld_star_flux(sr, r, fr, SM, SR);
        done = 1;
      }
      if (!done) {
        /* computeDoubleStarState, hlld_cuda.cu:412-474 */
        double vyd, vzd, byd, bzd, EdL = 0.0, EdR = 0.0;
        if (0.5 * (bx * bx) < HLLD_SMALL * pts) {
          if (SM >= 0.0) {
            vyd = sl.vy; vzd = sl.vz; byd = sl.by; bzd = sl.bz; EdL = sl.E;
          } else {
            vyd = sr.vy; vzd = sr.vz; byd = sr.by; bzd = sr.bz; EdR = sr.E;
          }
        } else {
          double const sdl = sqrt(sl.d), sdr = sqrt(sr.d);
          double const inv = 1.0 / (sdl + sdr);
          double const sg  = copysign(1.0, bx);
          vyd = inv * (sdl * sl.vy + sdr * sr.vy + sg * (sr.by - sl.by));
          vzd = inv * (sdl * sl.vz + sdr * sr.vz + sg * (sr.bz - sl.bz));
          byd = inv * (sdl * sr.by + sdr * sl.by + sg * (sdl * sdr) * (sr.vy - sl.vy));
          bzd = inv * (sdl * sr.bz + sdr * sl.bz + sg * (sdl * sdr) * (sr.vz - sl.vz));
          double const vdb = dot3(SM, vyd, vzd, bx, byd, bzd);
          EdL = sl.E - sdl * sg * (dot3(SM, sl.vy, sl.vz, bx, sl.by, sl.bz) - vdb);
          EdR = sr.E + sdr * sg * (dot3(SM, sr.vy, sr.vz, bx, sr.by, sr.bz) - vdb);
        }
        if (SM > 0.0 && SLs <= 0.0) {
          f    = hlld_dstar_flux(vyd, vzd, byd, bzd, EdL, sl, l, fl, SM, SL, SLs);
          done = 1;
        } else if (SRs > 0.0 && SM <= 0.0) {
          f    = hlld_dstar_flux(vyd, vzd, byd, bzd, EdR, sr, r, fr, SM, SR, SRs);
          done = 1;
        }
      }
    }
  }
  if (!done) return 0;
  out[0] = f.d; out[1] = f.mx; out[2] = f.my; out[3] = f.mz; out[4] = f.E; out[5] = f.by; out[6] = f.bz;
  return 1;
}

/* one launch of Calculate_HLLD_Fluxes_CUDA: every cell, no guard */
static void hlld_dir(const double *QL, const double *QR, const double *Bface, double *F, size_t N, double gamma, int dir)
{
  int o1, o2, o3;
  dir_perm(dir, &o1, &o2, &o3);
#pragma omp parallel for schedule(static)
  for (long id = 0; id < (long)N; id++) {
    double ql[7], qr[7], f[7];
    ql[0] = QL[id]; ql[1] = QL[o1 * N + id]; ql[2] = QL[o2 * N + id]; ql[3] = QL[o3 * N + id]; ql[4] = QL[4 * N + id];
    ql[5] = QL[5 * N + id]; ql[6] = QL[6 * N + id];
    qr[0] = QR[id]; qr[1] = QR[o1 * N + id]; qr[2] = QR[o2 * N + id]; qr[3] = QR[o3 * N + id]; qr[4] = QR[4 * N + id];
    qr[5] = QR[5 * N + id]; qr[6] = QR[6 * N + id];
    if (!morc_hlld_single(ql, qr, Bface[id], gamma, f)) continue;
    /* returnFluxes, hlld_cuda.cu:290-312 */
    F[id] = f[0]; F[o1 * N + id] = f[1]; F[o2 * N + id] = f[2]; F[o3 * N + id] = f[3]; F[4 * N + id] = f[4];
    F[5 * N + id] = f[5]; F[6 * N + id] = f[6];
  }
}

/* ------------------------------------------------------------------------------------------------------------
 * constrained transport (ct_electric_fields.{h,cu}) and the magnetic update (magnetic_update.cu)
 * ---------------------------------------------------------------------------------------------------------- */
static double ct_slope(const double *flux, const double *U, double sign, int ctdir, int cq1, int cq2, int fq1, int fq2,
                       int x, int y, int z, int nx, int ny, size_t N)
{
  int const m1 = (ctdir + 1) % 3, m2 = (ctdir + 2) % 3;
  int const xc = x - (cq1 == 0) - (cq2 == 0), yc = y - (cq1 == 1) - (cq2 == 1), zc = z - (cq1 == 2) - (cq2 == 2);
  long const ic = xc + (long)nx * (yc + (long)ny * zc);
  long const ifl = (x - (fq1 == 0) - (fq2 == 0)) + (long)nx * ((y - (fq1 == 1) - (fq2 == 1)) + (long)ny * (z - (fq1 == 2) - (fq2 == 2)));
  long const ib2 = (xc - (m1 == 0)) + (long)nx * ((yc - (m1 == 1)) + (long)ny * (zc - (m1 == 2)));
  long const ib3 = (xc - (m2 == 0)) + (long)nx * ((yc - (m2 == 1)) + (long)ny * (zc - (m2 == 2)));
  double const d  = U[ic];
  double const M2 = U[ic + (m1 + F_MX) * N], M3 = U[ic + (m2 + F_MX) * N];
  double const B2 = 0.5 * (U[ic + (m1 + F_BX) * N] + U[ib2 + (m1 + F_BX) * N]);
  double const B3 = 0.5 * (U[ic + (m2 + F_BX) * N] + U[ib3 + (m2 + F_BX) * N]);
  double const ec = (M3 * B2 - M2 * B3) / d;
  double const ef = sign * flux[ifl + ((sign == 1) + 5) * N];
  return ef - ec;
}
#define IDX(a, b, c) ((long)(a) + (long)nx * ((long)(b) + (long)ny * (long)(c)))
#define UPWIND(sv, A, B) ((sv) > 0.0 ? (A) : ((sv) < 0.0 ? (B) : 0.5 * ((A) + (B))))
void morc_ct_electric_fields(const double *Fx, const double *Fy, const double *Fz, const double *U, double *E, int nx,
                             int ny, int nz)
{
  size_t const N = (size_t)nx * ny * nz;
#pragma omp parallel for schedule(static)
  for (int z = 1; z < nz; z++)
    for (int y = 1; y < ny; y++)
      for (int x = 1; x < nx; x++) {
        long const id = IDX(x, y, z);
        double su, syp, syn, szp, szn, sxp, sxn, a, b;
#define SL(fl, sg, cd, c1, c2, f1, f2) ct_slope(fl, U, sg, cd, c1, c2, f1, f2, x, y, z, nx, ny, N)
        /* X electric field (ct_electric_fields.cu:65-129) */
        su = Fz[IDX(x, y, z - 1)];
        a = SL(Fy, -1, 0, 2, -1, 1, 2); b = SL(Fy, -1, 0, -1, -1, 1, -1);
        syp = UPWIND(su, a, b);
        su = Fz[IDX(x, y - 1, z - 1)];
        a = SL(Fy, -1, 0, 1, 2, 1, 2); b = SL(Fy, -1, 0, 1, -1, 1, -1);
        syn = UPWIND(su, a, b);
        su = Fy[IDX(x, y - 1, z)];
        a = SL(Fz, 1, 0, 1, -1, 1, 2); b = SL(Fz, 1, 0, -1, -1, 2, -1);
        szp = UPWIND(su, a, b);
        su = Fy[IDX(x, y - 1, z - 1)];
        a = SL(Fz, 1, 0, 1, 2, 1, 2); b = SL(Fz, 1, 0, 2, -1, -1, 2);
        szn = UPWIND(su, a, b);
        {
          double const fyp = +Fz[IDX(x, y, z - 1) + 6 * N], fyn = +Fz[IDX(x, y - 1, z - 1) + 6 * N];
          double const fzp = -Fy[IDX(x, y - 1, z) + 5 * N], fzn = -Fy[IDX(x, y - 1, z - 1) + 5 * N];
          E[id + 0 * N] = 0.25 * (+fyp + fyn + fzp + fzn + syp + syn + szp + szn);
        }
        /* Y electric field (:131-194) */
        su = Fz[IDX(x, y, z - 1)];
        a = SL(Fx, 1, 1, 2, -1, 0, 2); b = SL(Fx, 1, 1, -1, -1, 0, -1);
        sxp = UPWIND(su, a, b);
        su = Fz[IDX(x - 1, y, z - 1)];
        a = SL(Fx, 1, 1, 0, 2, 0, 2); b = SL(Fx, 1, 1, 0, -1, 0, -1);
        sxn = UPWIND(su, a, b);
        su = Fx[IDX(x - 1, y, z)];
        a = SL(Fz, -1, 1, 0, -1, 0, 2); b = SL(Fz, -1, 1, -1, -1, 2, -1);
        szp = UPWIND(su, a, b);
        su = Fx[IDX(x - 1, y, z - 1)];
        a = SL(Fz, -1, 1, 0, 2, 0, 2); b = SL(Fz, -1, 1, 2, -1, 2, -1);
        szn = UPWIND(su, a, b);
        {
          double const fxp = -Fz[IDX(x, y, z - 1) + 5 * N], fxn = -Fz[IDX(x - 1, y, z - 1) + 5 * N];
          double const fzp = +Fx[IDX(x - 1, y, z) + 6 * N], fzn = +Fx[IDX(x - 1, y, z - 1) + 6 * N];
          E[id + 1 * N] = 0.25 * (+fxp + fxn + fzp + fzn + sxp + sxn + szp + szn);
        }
        /* Z electric field (:196-260) */
        su = Fx[IDX(x - 1, y, z)];
        a = SL(Fy, 1, 2, 0, -1, 0, 1); b = SL(Fy, 1, 2, -1, -1, 1, -1);
        syp = UPWIND(su, a, b);
        su = Fx[IDX(x - 1, y - 1, z)];
        a = SL(Fy, 1, 2, 0, 1, 0, 1); b = SL(Fy, 1, 2, 1, -1, 1, -1);
        syn = UPWIND(su, a, b);
        su = Fy[IDX(x, y - 1, z)];
        a = SL(Fx, -1, 2, 1, -1, 0, 1); b = SL(Fx, -1, 2, -1, -1, 0, -1);
        sxp = UPWIND(su, a, b);
        su = Fy[IDX(x - 1, y - 1, z)];
        a = SL(Fx, -1, 2, 0, 1, 0, 1); b = SL(Fx, -1, 2, 0, -1, 0, -1);
        sxn = UPWIND(su, a, b);
        {
          double const fxp = +Fy[IDX(x, y - 1, z) + 6 * N], fxn = +Fy[IDX(x - 1, y - 1, z) + 6 * N];
          double const fyp = -Fx[IDX(x - 1, y, z) + 5 * N], fyn = -Fx[IDX(x - 1, y - 1, z) + 5 * N];
          E[id + 2 * N] = 0.25 * (+fxp + fxn + fyp + fyn + sxp + sxn + syp + syn);
        }
#undef SL
      }
}

/* Update_Magnetic_Field_3D, magnetic_update.cu:21-80 */
void morc_update_magnetic(const double *src, double *dst, const double *E, int nx, int ny, int nz, double dt, double dx,
                          double dy, double dz)
{
  size_t const N = (size_t)nx * ny * nz;
  double const dtodx = dt / dx, dtody = dt / dy, dtodz = dt / dz;
  for (int z = 1; z < nz - 1; z++)
    for (int y = 1; y < ny - 1; y++)
      for (int x = 1; x < nx - 1; x++) {
        long const id = IDX(x, y, z);
        double const ex1 = E[IDX(x, y + 1, z)], ex2 = E[IDX(x, y, z + 1)], ex3 = E[IDX(x, y + 1, z + 1)];
        double const ey1 = E[IDX(x + 1, y, z) + N], ey2 = E[IDX(x, y, z + 1) + N], ey3 = E[IDX(x + 1, y, z + 1) + N];
        double const ez1 = E[IDX(x + 1, y, z) + 2 * N], ez2 = E[IDX(x, y + 1, z) + 2 * N],
                     ez3 = E[IDX(x + 1, y + 1, z) + 2 * N];
        dst[id + F_BX * N] = src[id + F_BX * N] + dtodz * (ey3 - ey1) + dtody * (ez1 - ez3);
        dst[id + F_BY * N] = src[id + F_BY * N] + dtodx * (ez3 - ez2) + dtodz * (ex1 - ex3);
        dst[id + F_BZ * N] = src[id + F_BZ * N] + dtody * (ex3 - ex2) + dtodx * (ey2 - ey3);
      }
}

/* ------------------------------------------------------------------------------------------------------------
 * the step
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  int nx, ny, nz;
  double *half, *QL[3], *QR[3], *F[3], *E;
} morc_ctx;

morc_ctx *morc_create(int nx, int ny, int nz)
{
  size_t const N = (size_t)nx * ny * nz;
  morc_ctx *c    = (morc_ctx *)calloc(1, sizeof(morc_ctx));
  c->nx = nx; c->ny = ny; c->nz = nz;
  c->half = (double *)calloc(NF * N, sizeof(double));
  for (int d = 0; d < 3; d++) {
    c->QL[d] = (double *)calloc(NQ * N, sizeof(double));
    c->QR[d] = (double *)calloc(NQ * N, sizeof(double));
    c->F[d]  = (double *)calloc(NQ * N, sizeof(double));
  }
  c->E = (double *)calloc(3 * N, sizeof(double));
  return c;
}
void morc_destroy(morc_ctx *c)
{
  if (!c) return;
  free(c->half);
  for (int d = 0; d < 3; d++) { free(c->QL[d]); free(c->QR[d]); free(c->F[d]); }
  free(c->E);
  free(c);
}
const double *morc_flux(morc_ctx *c, int dir) { return c->F[dir]; }
const double *morc_efield(morc_ctx *c) { return c->E; }
const double *morc_half(morc_ctx *c) { return c->half; }

/* VL_Algorithm_3D_CUDA, MHD branch */
void morc_vl_step(morc_ctx *c, double *U, int n_ghost, double dx, double dy, double dz, double dt, double gamma,
                  int recon, double density_floor)
{
  int const nx = c->nx, ny = c->ny, nz = c->nz;
  size_t const N = (size_t)nx * ny * nz;
  size_t const str[3] = {1, (size_t)nx, (size_t)nx * ny};
  /* steps 1, 2, 2.5 */
  pcm_3d(U, c->QL, c->QR, nx, ny, nz);
  for (int d = 0; d < 3; d++) hlld_dir(c->QL[d], c->QR[d], U + (F_BX + d) * N, c->F[d], N, gamma, d);
  morc_ct_electric_fields(c->F[0], c->F[1], c->F[2], U, c->E, nx, ny, nz);
  /* step 3: Update_Conserved_Variables_3D_half (VL_3D_cuda.cu:379-495), hydro fields only */
  {
    double const h = 0.5 * dt, dtodx = h / dx, dtody = h / dy, dtodz = h / dz;
    for (int z = 1; z < nz - 1; z++)
      for (int y = 1; y < ny - 1; y++)
        for (int x = 1; x < nx - 1; x++) {
          size_t const id = IDX(x, y, z);
          for (int f = 0; f < 5; f++)
            c->half[f * N + id] = U[f * N + id] + dtodx * (c->F[0][f * N + id - str[0]] - c->F[0][f * N + id]) +
                                  dtody * (c->F[1][f * N + id - str[1]] - c->F[1][f * N + id]) +
                                  dtodz * (c->F[2][f * N + id - str[2]] - c->F[2][f * N + id]);
          if (c->half[id] < density_floor) {
            double const d0 = c->half[id];
            c->half[id]     = density_floor;
            for (int f = 1; f < 5; f++) c->half[f * N + id] *= (density_floor / d0);
          }
        }
    morc_update_magnetic(U, c->half, c->E, nx, ny, nz, h, dx, dy, dz);
  }
  /* step 4 */
  if (recon == MORC_PCM) pcm_3d(c->half, c->QL, c->QR, nx, ny, nz);
  else
    for (int d = 0; d < 3; d++) morc_reconstruct(recon, c->half, c->QL[d], c->QR[d], nx, ny, nz, gamma, d);
  /* steps 5, 5.5 */
  for (int d = 0; d < 3; d++) hlld_dir(c->QL[d], c->QR[d], c->half + (F_BX + d) * N, c->F[d], N, gamma, d);
  morc_ct_electric_fields(c->F[0], c->F[1], c->F[2], c->half, c->E, nx, ny, nz);
  /* step 6: Update_Conserved_Variables_3D (hydro_cuda.cu:172-300) on real cells, then the magnetic update */
  {
    double const dtodx = dt / dx, dtody = dt / dy, dtodz = dt / dz;
    int const g = n_ghost;
    for (int z = g; z < nz - g; z++)
      for (int y = g; y < ny - g; y++)
        for (int x = g; x < nx - g; x++) {
          size_t const id = IDX(x, y, z);
          for (int f = 0; f < 5; f++)
            U[f * N + id] += dtodx * (c->F[0][f * N + id - str[0]] - c->F[0][f * N + id]) +
                             dtody * (c->F[1][f * N + id - str[1]] - c->F[1][f * N + id]) +
                             dtodz * (c->F[2][f * N + id - str[2]] - c->F[2][f * N + id]);
          if (U[id] < density_floor && U[id] > 0) {
            double const d0 = U[id];
            U[id]           = density_floor;
            for (int f = 1; f < 5; f++) U[f * N + id] *= (density_floor / d0);
          }
        }
    morc_update_magnetic(U, U, c->E, nx, ny, nz, dt, dx, dy, dz);
  }
}

/* mhdInverseCrossingTime, hydro_cuda.cu:420-435 */
double morc_mhd_inverse_crossing_time(double E, double d, double d_inv, double vx, double vy, double vz, double bx,
                                      double by, double bz, double dx, double dy, double dz, double gamma)
{
  (void)d_inv;
  double const p  = pressure_primitive(E, d, vx, vy, vz, gamma, bx, by, bz);
  double const cf = morc_fast_speed(d, p, bx, by, bz, gamma);
  double m        = fmax((fabs(vx) + cf) / dx, (fabs(vy) + cf) / dy);
  m               = fmax(m, (fabs(vz) + cf) / dz);
  m               = fmax(m, 0.0);
  return m;
}
/* Calc_dt_3D (MHD), hydro_cuda.cu:514-560 */
double morc_calc_dti(const double *U, int nx, int ny, int nz, int g, double dx, double dy, double dz, double gamma)
{
  size_t const N = (size_t)nx * ny * nz;
  double m       = -DBL_MAX;
  for (int z = g; z < nz - g; z++)
    for (int y = g; y < ny - g; y++)
      for (int x = g; x < nx - g; x++) {
        size_t const id = IDX(x, y, z);
        double const d = U[id], di = 1.0 / d;
        double const vx = U[N + id] * di, vy = U[2 * N + id] * di, vz = U[3 * N + id] * di;
        double bx, by, bz;
        cell_centred_b(U, id, x, y, z, N, nx, ny, &bx, &by, &bz);
        m = fmax(m, morc_mhd_inverse_crossing_time(U[4 * N + id], d, di, vx, vy, vz, bx, by, bz, dx, dy, dz, gamma));
      }
  return m;
}

/* max |div B| over real cells (magnetic_divergence.cu): a parity-independent property check */
double morc_max_divergence(const double *U, int nx, int ny, int nz, int g, double dx, double dy, double dz)
{
  size_t const N = (size_t)nx * ny * nz;
  double m       = 0.0;
  for (int z = g; z < nz - g; z++)
    for (int y = g; y < ny - g; y++)
      for (int x = g; x < nx - g; x++) {
        size_t const id = IDX(x, y, z);
        double const div = (U[F_BX * N + id] - U[F_BX * N + id - 1]) / dx +
                           (U[F_BY * N + id] - U[F_BY * N + id - nx]) / dy +
                           (U[F_BZ * N + id] - U[F_BZ * N + id - (size_t)nx * ny]) / dz;
        m = fmax(m, fabs(div));
      }
  return m;
}

/* mhd::calculateMagneticDivergence as the reference runs it (magnetic_divergence.cu:33-76): every cell whose three low
 * faces exist -- x, y, z > 1, ghost cells included -- not only the real cells; what chl_mhd_max_divergence returns and what
 * the reference's unit test pins (magnetic_divergence_tests.cu:27-70) */
double morc_check_magnetic_divergence(const double *U, int nx, int ny, int nz, double dx, double dy, double dz)
{
  size_t const N = (size_t)nx * ny * nz;
  double m       = 0.0;
  for (int z = 2; z < nz; z++)
    for (int y = 2; y < ny; y++)
      for (int x = 2; x < nx; x++) {
        size_t const id = IDX(x, y, z);
        double const div = ((U[F_BX * N + id] - U[F_BX * N + id - 1]) / dx) +
                           ((U[F_BY * N + id] - U[F_BY * N + id - nx]) / dy) +
                           ((U[F_BZ * N + id] - U[F_BZ * N + id - (size_t)nx * ny]) / dz);
        m = fmax(m, fabs(div));
      }
  return m;
}

/* ---- entry points for the reference's small-unit fiducials (reconstruction_tests.cu:51-175) ---- */
static mprim_t prim_from8(const double p[8])
{
  mprim_t w;
  w.d = p[0]; w.vx = p[1]; w.vy = p[2]; w.vz = p[3]; w.p = p[4]; w.bx = p[5]; w.by = p[6]; w.bz = p[7];
  return w;
}
/* helper-level hooks (reconstruction_tests.cu:205-249, 336-392, 574-611, MHD branches); tests/test_oracle_helpers_golden.py */
static void prim_to8(mprim_t o, double out[8])
{
  out[0] = o.d; out[1] = o.vx; out[2] = o.vy; out[3] = o.vz; out[4] = o.p; out[5] = o.bx; out[6] = o.by; out[7] = o.bz;
}
void morc_test_load_data(const double *U, int x, int y, int z, int nx, int ny, long N, int o1, int o2, int o3, double gamma,
                         double out[8])
{
  prim_to8(load_prim(U, x, y, z, nx, ny, (size_t)N, o1, o2, o3, gamma), out);
}
void morc_test_write_data(const double prim[8], double *Q, const double *U, long id, long N, int o1, int o2, int o3,
                          double gamma)
{
  write_interface(prim_from8(prim), Q, U, (size_t)id, (size_t)N, o1, o2, o3, gamma);
}
void morc_test_monotonize_characteristic(const double prim[8], const double aL[7], const double aR[7], const double aC[7],
                                         const double aG[7], const double eig[18], double a2, double out[8])
{
  eigen_t e;
  memcpy(&e, eig, 18 * sizeof(double));
  mchar_t am;
  for (int i = 0; i < 7; i++) am.a[i] = mono(aL[i], aR[i], aC[i], aG[i]);
  mprim_t o = from_char(prim_from8(prim), am, e, a2);
  o.bx = 0.0;
  prim_to8(o, out);
}
void morc_eigenvectors(const double prim[8], double gamma, double out[18])
{
  mprim_t const w = prim_from8(prim);
  double const a = sqrt(gamma * w.p / w.d);
  eigen_t const e = eigenvectors(w, a, a * a, gamma);
  memcpy(out, &e, 18 * sizeof(double));
}
void morc_prim_to_char(const double prim[8], const double slope[8], const double eig[18], double gamma, double out[7])
{
  mprim_t const w = prim_from8(prim);
  double const a = sqrt(gamma * w.p / w.d);
  eigen_t e;
  memcpy(&e, eig, 18 * sizeof(double));
  mchar_t const c = to_char(w, prim_from8(slope), e, a * a);
  memcpy(out, c.a, 7 * sizeof(double));
}
void morc_char_to_prim(const double prim[8], const double chr[7], const double eig[18], double gamma, double out[8])
{
  mprim_t const w = prim_from8(prim);
  double const a = sqrt(gamma * w.p / w.d);
  eigen_t e;
  memcpy(&e, eig, 18 * sizeof(double));
  mchar_t c;
  memcpy(c.a, chr, 7 * sizeof(double));
  mprim_t const o = from_char(w, c, e, a * a);
  out[0] = o.d; out[1] = o.vx; out[2] = o.vy; out[3] = o.vz; out[4] = o.p; out[5] = 0.0; out[6] = o.by; out[7] = o.bz;
}
