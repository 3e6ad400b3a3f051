/*
 * wo_thermo.c -- oracle (TEST INFRASTRUCTURE): power tables, IAPWS-97 and
 * IFC-67 water/steam thermodynamics restated from the reference:
 *   src/powertable.F90, src/IAPWS.F90, src/IFC67.F90, src/thermodynamics.F90,
 *   src/utils.F90:651-709 (newton1d).
 * Operation order follows the reference (power tables are multiplication
 * chains, sums run in array order, Horner forms as written).
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* power table: src/powertable.F90                                     */
/* ------------------------------------------------------------------ */

void wo_powertable_init(wo_powertable *t) { memset(t, 0, sizeof(*t)); }

#define PT_PROD(t, k, i) ((t)->product[2 * ((i) - (t)->lower) + (k)])
#define PT_REQ(t, i) ((t)->required[(i) - (t)->lower])

/* powertable.F90:80-91 */
static int pt_product_configured(const wo_powertable *t, int i) {
  return (abs(i) <= 1) || (PT_PROD(t, 0, i) != 0 && PT_PROD(t, 1, i) != 0);
}

/* powertable.F90:95-131 */
static void pt_configure_product(wo_powertable *t, int i) {
  if (abs(i) > 1 && !pt_product_configured(t, i)) {
    int s = (i >= 0) ? 1 : -1;
    int i2 = (int)lround(i / 2.0); /* Fortran nint: half away from zero */
    for (int c = i2; (s > 0) ? (c >= s) : (c <= s); c -= s) {
      int j = i - c;
      if (pt_product_configured(t, c) && pt_product_configured(t, j)) {
        PT_PROD(t, 0, i) = c;
        PT_PROD(t, 1, i) = j;
        break;
      }
    }
    if (!pt_product_configured(t, i)) {
      if (i % 2 == 0) {
        pt_configure_product(t, i2);
        PT_PROD(t, 0, i) = i2;
        PT_PROD(t, 1, i) = i2;
        PT_REQ(t, i2) = 2;
      } else {
        int j = i - s;
        pt_configure_product(t, j);
        PT_PROD(t, 0, i) = s;
        PT_PROD(t, 1, i) = j;
        PT_REQ(t, j) = 2;
      }
    }
  }
}

/* powertable.F90:135-245 */
void wo_powertable_configure(wo_powertable *t, const int *powers, int n) {
  int minp = 0, maxp = 1;
  for (int k = 0; k < n; k++) {
    if (powers[k] < minp) minp = powers[k];
    if (powers[k] > maxp) maxp = powers[k];
  }
  int enlarge = (t->power != NULL);
  int old_lower = 0, old_upper = 0;
  int *old_required = NULL;
  if (enlarge) {
    old_lower = t->lower;
    old_upper = t->upper;
    old_required = t->required;
    if (minp < t->lower) t->lower = minp;
    if (maxp > t->upper) t->upper = maxp;
    free(t->power);
    free(t->product);
    t->required = NULL;
  } else {
    t->lower = minp;
    t->upper = maxp;
  }
  int sz = t->upper - t->lower + 1;
  t->power = (double *)calloc(sz, sizeof(double));
  t->product = (int *)calloc(2 * sz, sizeof(int));
  t->required = (int *)calloc(sz, sizeof(int));
  t->power[0 - t->lower] = 1.0;
  if (enlarge) {
    for (int i = old_lower; i <= old_upper; i++)
      if (old_required[i - old_lower] == 1) PT_REQ(t, i) = 1;
    free(old_required);
  }
  for (int k = 0; k < n; k++)
    if (abs(powers[k]) > 1) PT_REQ(t, powers[k]) = 1;
  for (int s = 1; s >= -1; s -= 2) {
    int u = (s > 0) ? t->upper : t->lower;
    for (int i = s * 2; (s > 0) ? (i <= u) : (i >= u); i += s)
      if (PT_REQ(t, i) > 0) pt_configure_product(t, i);
  }
  /* set_powerlist: powertable.F90:214-245 */
  free(t->list);
  int cnt = 0;
  for (int i = t->lower; i <= t->upper; i++)
    if (PT_REQ(t, i) > 0) cnt++;
  t->nlist = cnt;
  t->list = (int *)malloc(3 * (cnt > 0 ? cnt : 1) * sizeof(int));
  int m = 0;
  for (int s = 1; s >= -1; s -= 2) {
    int u = (s > 0) ? t->upper : t->lower;
    for (int p = s * 2; (s > 0) ? (p <= u) : (p >= u); p += s)
      if (PT_REQ(t, p) > 0) {
        t->list[3 * m + 0] = PT_PROD(t, 0, p);
        t->list[3 * m + 1] = PT_PROD(t, 1, p);
        t->list[3 * m + 2] = p;
        m++;
      }
  }
}

/* powertable.F90:261-278 */
void wo_powertable_compute(wo_powertable *t, double val) {
  double *pw = t->power - t->lower;
  pw[1] = val;
  if (t->lower < 0) pw[-1] = 1.0 / val;
  for (int m = 0; m < t->nlist; m++)
    pw[t->list[3 * m + 2]] = pw[t->list[3 * m + 0]] * pw[t->list[3 * m + 1]];
}

double wo_powertable_get(const wo_powertable *t, int i) { return t->power[i - t->lower]; }

void wo_powertable_destroy(wo_powertable *t) {
  free(t->power);
  free(t->product);
  free(t->required);
  free(t->list);
  memset(t, 0, sizeof(*t));
}

void wo_powertable_eval(const int *powers, int n, double val, const int *query, int nq, double *out) {
  wo_powertable t;
  wo_powertable_init(&t);
  wo_powertable_configure(&t, powers, n);
  wo_powertable_compute(&t, val);
  for (int k = 0; k < nq; k++) out[k] = wo_powertable_get(&t, query[k]);
  wo_powertable_destroy(&t);
}

/* ------------------------------------------------------------------ */
/* thermodynamics object                                               */
/* ------------------------------------------------------------------ */

struct wo_thermo {
  int id;
  double tcriticalk, tcritical, pcritical, dcritical;
  double r1_max_temperature;
  /* IAPWS power tables: IAPWS.F90:468-473, 566-575, 661-666, 738-741 */
  wo_powertable r1_pi, r1_pj;
  wo_powertable r2_pj0, r2_pi, r2_pj;
  wo_powertable r3_pi, r3_pj;
  wo_powertable v_pi, v_pj, v_pk;
};

/* ---- IAPWS-97 coefficient tables (IAPWS.F90:48-233) ---- */
static const double sat_n[10] = {
    0.11670521452767e4, -0.72421316703206e6, -0.17073846940092e2, 0.12020824702470e5,
    -0.32325550322333e7, 0.14915108613530e2, -0.48232657361591e4, 0.40511340542057e6,
    -0.23855557567849,  0.65017534844798e3};

static const double visc_h0[4] = {1.67752, 2.20462, 0.6366564, -0.241605};
static const double visc_h1[21] = {
    5.20094e-1, 8.50895e-2,  -1.08374,    -2.89555e-1, 2.22531e-1,  9.99115e-1, 1.88797,
    1.26613,    1.20573e-1,  -2.81378e-1, -9.06851e-1, -7.72479e-1, -4.89837e-1, -2.57040e-1,
    1.61913e-1, 2.57399e-1,  -3.25372e-2, 6.98452e-2,  8.72102e-3,  -4.35673e-3, -5.93264e-4};
static const int visc_I[21] = {0, 1, 2, 3, 0, 1, 2, 3, 5, 0, 1, 2, 3, 4, 0, 1, 0, 3, 4, 3, 5};
static const int visc_J[21] = {0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 3, 3, 4, 4, 5, 6, 6};
static const int visc_K[4] = {0, 1, 2, 3};

static const double r1_n[34] = {
    0.14632971213167,     -0.84548187169114,    -0.37563603672040e1,  0.33855169168385e1,
    -0.95791963387872,    0.15772038513228,     -0.16616417199501e-1, 0.81214629983568e-3,
    0.28319080123804e-3,  -0.60706301565874e-3, -0.18990068218419e-1, -0.32529748770505e-1,
    -0.21841717175414e-1, -0.52838357969930e-4, -0.47184321073267e-3, -0.30001780793026e-3,
    0.47661393906987e-4,  -0.44141845330846e-5, -0.72694996297594e-15, -0.31679644845054e-4,
    -0.28270797985312e-5, -0.85205128120103e-9, -0.22425281908000e-5, -0.65171222895601e-6,
    -0.14341729937924e-12, -0.40516996860117e-6, -0.12734301741641e-8, -0.17424871230634e-9,
    -0.68762131295531e-18, 0.14478307828521e-19, 0.26335781662795e-22, -0.11947622640071e-22,
    0.18228094581404e-23, -0.93537087292458e-25};
static const int r1_I[34] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 2, 2, 2,
                             2, 2, 3, 3, 3, 4, 4, 4, 5, 8, 8, 21, 23, 29, 30, 31, 32};
static const int r1_J[34] = {-2, -1, 0,  1,  2,  3,  4,   5,  -9,  -7,  -1,  0,   1,   3,   -3,  0,  1,
                             3,  17, -4, 0,  6,  -5, -2,  10, -8,  -11, -6,  -29, -31, -38, -39, -40, -41};

static const double r2_n0[9] = {-0.96927686500217e1, 0.10086655968018e2,  -0.56087911283020e-2,
                                0.71452738081455e-1, -0.40710498223928,   0.14240819171444e1,
                                -0.43839511319450e1, -0.28408632460772,   0.21268463753307e-1};
static const int r2_J0[9] = {0, 1, -5, -4, -3, -2, -1, 2, 3};
static const double r2_n[43] = {
    -0.17731742473213e-2,  -0.17834862292358e-1,  -0.45996013696365e-1,  -0.57581259083432e-1,
    -0.50325278727930e-1,  -0.33032641670203e-4,  -0.18948987516315e-3,  -0.39392777243355e-2,
    -0.43797295650573e-1,  -0.26674547914087e-4,  0.20481737692309e-7,   0.43870667284435e-6,
    -0.32277677238570e-4,  -0.15033924542148e-2,  -0.40668253562649e-1,  -0.78847309559367e-9,
    0.12790717852285e-7,   0.48225372718507e-6,   0.22922076337661e-5,   -0.16714766451061e-10,
    -0.21171472321355e-2,  -0.23895741934104e2,   -0.59059564324270e-17, -0.12621808899101e-5,
    -0.38946842435739e-1,  0.11256211360459e-10,  -0.82311340897998e1,   0.19809712802088e-7,
    0.10406965210174e-18,  -0.10234747095929e-12, -0.10018179379511e-8,  -0.80882908646985e-10,
    0.10693031879409,      -0.33662250574171,     0.89185845355421e-24,  0.30629316876232e-12,
    -0.42002467698208e-5,  -0.59056029685639e-25, 0.37826947613457e-5,   -0.12768608934681e-14,
    0.73087610595061e-28,  0.55414715350778e-16,  -0.94369707241210e-6};
static const int r2_I[43] = {1, 1, 1, 1, 1,  2,  2,  2,  2,  2,  3,  3,  3,  3,  3,  4,  4,  4,  5,  6,  6, 6,
                             7, 7, 7, 8, 8,  9,  10, 10, 10, 16, 16, 18, 20, 20, 20, 21, 22, 23, 24, 24, 24};
static const int r2_J[43] = {0,  1,  2,  3,  6,  1,  2,  4,  7,  36, 0,  1,  3,  6,  35, 1,  2,  3,  7,  3, 16, 35,
                             0,  11, 25, 8,  36, 13, 4,  10, 14, 29, 50, 57, 20, 35, 48, 21, 53, 39, 26, 40, 58};

static const double r3_n[40] = {
    0.10658070028513e1,  -0.15732845290239e2,  0.20944396974307e2,  -0.76867707878716e1,
    0.26185947787954e1,  -0.28080781148620e1,  0.12053369696517e1,  -0.84566812812502e-2,
    -0.12654315477714e1, -0.11524407806681e1,  0.88521043984318,    -0.64207765181607,
    0.38493460186671,    -0.85214708824206,    0.48972281541877e1,  -0.30502617256965e1,
    0.39420536879154e-1, 0.12558408424308,     -0.27999329698710,   0.13899799569460e1,
    -0.20189915023570e1, -0.82147637173963e-2, -0.47596035734923,   0.43984074473500e-1,
    -0.44476435428739,   0.90572070719733,     0.70522450087967,    0.10770512626332,
    -0.32913623258954,   -0.50871062041158,    -0.22175400873096e-1, 0.94260751665092e-1,
    0.16436278447961,    -0.13503372241348e-1, -0.14834345352472e-1, 0.57922953628084e-3,
    0.32308904703711e-2, 0.80964802996215e-4,  -0.16557679795037e-3, -0.44923899061815e-4};
static const int r3_I[40] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3,
                             3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 6, 6, 6, 7, 8, 9, 9, 10, 10, 11};
static const int r3_J[40] = {0, 0, 1,  2, 7, 10, 12, 23, 2, 6, 15, 17, 0, 2, 6,  7, 22, 26, 0, 2,
                             4, 16, 26, 0, 2, 4,  26, 1,  3, 26, 0,  2,  26, 2, 26, 2, 26, 0,  1, 26};

static const double b23_n[5] = {0.34805185628969e3, -0.11671859879975e1, 0.10192970039326e-2,
                                0.57254459862746e3, 0.13918839778870e2};

static void shifted(const int *a, int n, int *out) {
  for (int i = 0; i < n; i++) out[i] = a[i] - 1;
}

wo_thermo *wo_thermo_create(int id, int extrapolate) {
  wo_thermo *th = (wo_thermo *)calloc(1, sizeof(wo_thermo));
  th->id = id;
  th->r1_max_temperature = extrapolate ? 360.0 : 350.0; /* IAPWS.F90:456-483, IFC67.F90:235-249 */
  if (id == WO_THERMO_IAPWS) {
    /* IAPWS.F90:273-276 */
    th->tcriticalk = 647.096;
    th->tcritical = th->tcriticalk - WO_TC_K;
    th->pcritical = 22.064e6;
    th->dcritical = 322.0;
    int tmp[64];
    /* region 1: IAPWS.F90:468-473 */
    wo_powertable_configure(&th->r1_pi, r1_I, 34);
    shifted(r1_I, 34, tmp);
    wo_powertable_configure(&th->r1_pi, tmp, 34);
    wo_powertable_configure(&th->r1_pj, r1_J, 34);
    shifted(r1_J, 34, tmp);
    wo_powertable_configure(&th->r1_pj, tmp, 34);
    /* region 2: IAPWS.F90:566-575 */
    wo_powertable_configure(&th->r2_pj0, r2_J0, 9);
    shifted(r2_J0, 9, tmp);
    wo_powertable_configure(&th->r2_pj0, tmp, 9);
    wo_powertable_configure(&th->r2_pi, r2_I, 43);
    shifted(r2_I, 43, tmp);
    wo_powertable_configure(&th->r2_pi, tmp, 43);
    int m1 = -1;
    wo_powertable_configure(&th->r2_pi, &m1, 1);
    wo_powertable_configure(&th->r2_pj, r2_J, 43);
    shifted(r2_J, 43, tmp);
    wo_powertable_configure(&th->r2_pj, tmp, 43);
    /* region 3: IAPWS.F90:661-666 */
    wo_powertable_configure(&th->r3_pi, r3_I, 40);
    shifted(r3_I, 40, tmp);
    wo_powertable_configure(&th->r3_pi, tmp, 40);
    wo_powertable_configure(&th->r3_pj, r3_J, 40);
    shifted(r3_J, 40, tmp);
    wo_powertable_configure(&th->r3_pj, tmp, 40);
    /* viscosity: IAPWS.F90:738-741 */
    wo_powertable_configure(&th->v_pi, visc_I, 21);
    wo_powertable_configure(&th->v_pj, visc_J, 21);
    wo_powertable_configure(&th->v_pk, visc_K, 4);
  } else {
    /* IFC67.F90:157-160 */
    th->tcriticalk = 647.3;
    th->tcritical = th->tcriticalk - WO_TC_K;
    th->pcritical = 22.12e6;
    th->dcritical = 322.0;
  }
  return th;
}

void wo_thermo_destroy(wo_thermo *th) {
  if (!th) return;
  if (th->id == WO_THERMO_IAPWS) {
    wo_powertable_destroy(&th->r1_pi);
    wo_powertable_destroy(&th->r1_pj);
    wo_powertable_destroy(&th->r2_pj0);
    wo_powertable_destroy(&th->r2_pi);
    wo_powertable_destroy(&th->r2_pj);
    wo_powertable_destroy(&th->r3_pi);
    wo_powertable_destroy(&th->r3_pj);
    wo_powertable_destroy(&th->v_pi);
    wo_powertable_destroy(&th->v_pj);
    wo_powertable_destroy(&th->v_pk);
  }
  free(th);
}

int wo_thermo_id(const wo_thermo *th) { return th->id; }
double wo_thermo_tcritical(const wo_thermo *th) { return th->tcritical; }
double wo_thermo_pcritical(const wo_thermo *th) { return th->pcritical; }

/* ------------------------------------------------------------------ */
/* IAPWS-97                                                            */
/* ------------------------------------------------------------------ */

/* IAPWS.F90:503-542 */
static int iapws_region1(wo_thermo *th, const double param[2], double props[2]) {
  double p = param[0], t = param[1];
  if (t <= th->r1_max_temperature && p <= 100.e6) {
    const double pstar = 16.53e6, tstar = 1386.0;
    double tk = t + WO_TC_K;
    double rt = WO_RCONST * tk;
    double pi = p / pstar;
    double tau = tstar / tk;
    wo_powertable_compute(&th->r1_pi, 7.1 - pi);
    wo_powertable_compute(&th->r1_pj, tau - 1.222);
    const double *PI = th->r1_pi.power - th->r1_pi.lower;
    const double *PJ = th->r1_pj.power - th->r1_pj.lower;
    double s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < 34; i++) {
      double nI = r1_n[i] * r1_I[i], nJ = r1_n[i] * r1_J[i];
      s1 += nI * PI[r1_I[i] - 1] * PJ[r1_J[i]];
      s2 += nJ * PI[r1_I[i]] * PJ[r1_J[i] - 1];
    }
    double gampi = -s1, gamt = s2;
    props[0] = pstar / (rt * gampi);
    props[1] = rt * (tau * gamt - pi * gampi);
    return 0;
  }
  return 1;
}

/* IAPWS.F90:596-639 */
static int iapws_region2(wo_thermo *th, const double param[2], double props[2]) {
  double p = param[0], t = param[1];
  if (t <= 800.0 && p <= 100.e6) {
    const double pstar = 1.0e6, tstar = 540.0;
    double tk = t + WO_TC_K;
    double rt = WO_RCONST * tk;
    double pi = p / pstar;
    double tau = tstar / tk;
    wo_powertable_compute(&th->r2_pj0, tau);
    wo_powertable_compute(&th->r2_pi, pi);
    wo_powertable_compute(&th->r2_pj, tau - 0.5);
    const double *PJ0 = th->r2_pj0.power - th->r2_pj0.lower;
    const double *PI = th->r2_pi.power - th->r2_pi.lower;
    const double *PJ = th->r2_pj.power - th->r2_pj.lower;
    double gamt0 = 0.0, gampir = 0.0, gamtr = 0.0;
    for (int i = 0; i < 9; i++) gamt0 += (r2_n0[i] * r2_J0[i]) * PJ0[r2_J0[i] - 1];
    for (int i = 0; i < 43; i++) gampir += (r2_n[i] * r2_I[i]) * PI[r2_I[i] - 1] * PJ[r2_J[i]];
    for (int i = 0; i < 43; i++) gamtr += (r2_n[i] * r2_J[i]) * PI[r2_I[i]] * PJ[r2_J[i] - 1];
    double gampi = PI[-1] + gampir;
    props[0] = pstar / (rt * gampi);
    props[1] = rt * (tau * (gamt0 + gamtr) - pi * gampi);
    return 0;
  }
  return 1;
}

/* IAPWS.F90:689-727 ; param = (density, temperature) -> (pressure, internal energy) */
static int iapws_region3(wo_thermo *th, const double param[2], double props[2]) {
  double d = param[0], t = param[1];
  double tk = t + WO_TC_K;
  double rt = WO_RCONST * tk;
  double tau = th->tcriticalk / tk;
  double delta = d / th->dcritical;
  wo_powertable_compute(&th->r3_pi, delta);
  wo_powertable_compute(&th->r3_pj, tau);
  const double *PI = th->r3_pi.power - th->r3_pi.lower;
  const double *PJ = th->r3_pj.power - th->r3_pj.lower;
  double s1 = 0.0, s2 = 0.0;
  for (int i = 0; i < 40; i++) s1 += (r3_n[i] * r3_I[i]) * PI[r3_I[i] - 1] * PJ[r3_J[i]];
  for (int i = 0; i < 40; i++) s2 += (r3_n[i] * r3_J[i]) * PI[r3_I[i]] * PJ[r3_J[i] - 1];
  double phidelta = r3_n[0] * PI[-1] + s1;
  double phitau = s2;
  props[0] = d * rt * delta * phidelta;
  props[1] = rt * tau * phitau;
  return (props[0] > 100.0e6) ? 1 : 0;
}

/* IAPWS.F90:412-443 */
static double iapws_viscosity(wo_thermo *th, double temperature, double density) {
  double tk = temperature + WO_TC_K;
  double tau = tk / th->tcriticalk;
  double del = density / th->dcritical;
  wo_powertable_compute(&th->v_pk, 1.0 / tau);
  const double *PK = th->v_pk.power - th->v_pk.lower;
  wo_powertable_compute(&th->v_pi, PK[1] - 1.0);
  wo_powertable_compute(&th->v_pj, del - 1.0);
  const double *PI = th->v_pi.power - th->v_pi.lower;
  const double *PJ = th->v_pj.power - th->v_pj.lower;
  double s0 = 0.0;
  for (int k = 0; k < 4; k++) s0 += visc_h0[k] * PK[k];
  double mu0 = 100.0 * sqrt(tau) / s0;
  double s1 = 0.0;
  for (int i = 0; i < 21; i++) s1 += PI[visc_I[i]] * visc_h1[i] * PJ[visc_J[i]];
  double mu1 = exp(del * s1);
  return 1.0e-6 * mu0 * mu1;
}

/* IAPWS.F90:762-789 */
static int iapws_sat_pressure(const wo_thermo *th, double t, double *p) {
  if (t >= 0.0 && t <= th->tcritical) {
    const double *n = sat_n - 1; /* 1-based */
    double tk = t + WO_TC_K;
    double theta = tk + n[9] / (tk - n[10]);
    double theta2 = theta * theta;
    double a = theta2 + n[1] * theta + n[2];
    double b = n[3] * theta2 + n[4] * theta + n[5];
    double c = n[6] * theta2 + n[7] * theta + n[8];
    double x = 2.0 * c / (-b + sqrt(b * b - 4.0 * a * c));
    x = x * x;
    *p = 1.0e6 * x * x;
    return 0;
  }
  return 1;
}

/* IAPWS.F90:793-818 */
static int iapws_sat_temperature(const wo_thermo *th, double p, double *t) {
  if (p >= 611.213 && p <= th->pcritical) {
    const double *n = sat_n - 1;
    double beta2 = sqrt(p / 1.0e6);
    double beta = sqrt(beta2);
    double e = beta2 + n[3] * beta + n[6];
    double f = n[1] * beta2 + n[4] * beta + n[7];
    double g = n[2] * beta2 + n[5] * beta + n[8];
    double d = 2.0 * g / (-f - sqrt(f * f - 4.0 * e * g));
    double x = n[10] + d;
    *t = 0.5 * (n[10] + d - sqrt(x * x - 4.0 * (n[9] + n[10] * d))) - WO_TC_K;
    return 0;
  }
  return 1;
}

/* IAPWS.F90:824-851 */
double wo_boundary23_pressure(double t) {
  double tk = t + WO_TC_K;
  return 1.0e6 * (b23_n[0] + tk * (b23_n[1] + tk * b23_n[2]));
}
double wo_boundary23_temperature(double p) {
  return b23_n[3] + sqrt((p / 1.0e6 - b23_n[4]) / b23_n[2]) - WO_TC_K;
}

/* IAPWS.F90:317-365 */
static int iapws_phase_composition(const wo_thermo *th, int region, double pressure, double temperature) {
  int phases = 0;
  if (region == 4) {
    phases = 3;
  } else if (temperature <= th->tcritical) {
    if (region == 1) phases = 1;
    else if (region == 2) phases = 2;
    else if (region == 3) {
      double ps;
      if (iapws_sat_pressure(th, temperature, &ps) == 0) phases = (pressure >= ps) ? 1 : 2;
      else phases = -1;
    }
  } else {
    phases = (pressure <= th->pcritical) ? 2 : 4;
  }
  return phases;
}

/* ------------------------------------------------------------------ */
/* IFC-67                                                              */
/* ------------------------------------------------------------------ */

/* IFC67.F90:606-633 */
static int ifc67_sat_pressure(const wo_thermo *th, double t, double *p) {
  const double A1 = -7.691234564, A2 = -2.608023696e1, A3 = -1.681706546e2, A4 = 6.423285504e1,
               A5 = -1.189646225e2, A6 = 4.167117320, A7 = 2.097506760e1, A8 = 1.0e9, A9 = 6.0;
  if (t >= 1.0 && t <= th->tcritical) {
    double TC = (t + WO_TC_K) / th->tcriticalk;
    double X1 = 1.0 - TC;
    double X2 = X1 * X1;
    double SC = A5 * X1 + A4;
    SC = SC * X1 + A3;
    SC = SC * X1 + A2;
    SC = SC * X1 + A1;
    SC = SC * X1;
    double PC = exp(SC / (TC * (1.0 + A6 * X1 + A7 * X2)) - X1 / (A8 * X2 + A9));
    *p = PC * th->pcritical;
    return 0;
  }
  return 1;
}

/* IFC67.F90:637-676 with newton1d_general (utils.F90:651-709) */
static int ifc67_sat_temperature(const wo_thermo *th, double p, double *tout) {
  const int maxit = 200;
  const double ftol = 1.e-10, xtol = 1.e-10, inc = 1.e-8;
  if (p >= 0.0061e5 && p <= th->pcritical) {
    double x = fmax(4606.0 / (24.02 - log(p)) - WO_TC_K, 5.0);
    double ftolp = ftol * p;
    double delx = inc * x;
    int found = 0, err = 0;
    for (int i = 1; i <= maxit; i++) {
      double ps;
      err = ifc67_sat_pressure(th, x, &ps);
      double fx = p - ps;
      if (err == 0) {
        if (fabs(fx) <= ftolp) {
          found = 1;
          break;
        } else {
          err = ifc67_sat_pressure(th, x + delx, &ps);
          double fxd = p - ps;
          if (err == 0) {
            double df = (fxd - fx) / delx;
            double dx = -fx / df;
            x = x + dx;
            if (fabs(dx) <= xtol) {
              found = 1;
              break;
            }
          } else
            break;
        }
      } else
        break;
    }
    if (err == 0 && !found) err = 1;
    *tout = x;
    return err;
  }
  return 1;
}

/* IFC67.F90:265-374 */
static int ifc67_region1(const wo_thermo *th, const double param[2], double props[2]) {
  const double A1 = 6.824687741e3, A2 = -5.422063673e2, A3 = -2.096666205e4, A4 = 3.941286787e4,
               A5 = -13.466555478e4, A6 = 29.707143084e4, A7 = -4.375647096e5, A8 = 42.954208335e4,
               A9 = -27.067012452e4, A10 = 9.926972482e4, A11 = -16.138168904e3, A12 = 7.982692717,
               A13 = -2.616571843e-2, A14 = 1.522411790e-3, A15 = 2.284279054e-2, A16 = 2.421647003e2,
               A17 = 1.269716088e-10, A18 = 2.074838328e-7, A19 = 2.174020350e-8, A20 = 1.105710498e-9,
               A21 = 1.293441934e1, A22 = 1.308119072e-5, A23 = 6.047626338e-14;
  const double SA1 = 8.438375405e-1, SA2 = 5.362162162e-4, SA3 = 1.72, SA4 = 7.342278489e-2,
               SA5 = 4.975858870e-2, SA6 = 6.537154300e-1, SA7 = 1.150e-6, SA8 = 1.51080e-5,
               SA9 = 1.41880e-1, SA10 = 7.002753165, SA11 = 2.995284926e-4, SA12 = 2.040e-1;
  (void)A3;
  double p = param[0], t = param[1];
  if (t <= th->r1_max_temperature && p <= 100.e6) {
    double TKR = (t + WO_TC_K) / th->tcriticalk;
    double TKR2 = TKR * TKR;
    double TKR3 = TKR * TKR2;
    double TKR4 = TKR2 * TKR2;
    double TKR5 = TKR2 * TKR3;
    double TKR6 = TKR4 * TKR2;
    double TKR7 = TKR4 * TKR3;
    double TKR8 = TKR4 * TKR4;
    double TKR9 = TKR4 * TKR5;
    double TKR10 = TKR4 * TKR6;
    double TKR11 = TKR * TKR10;
    double TKR18 = TKR8 * TKR10;
    double TKR19 = TKR8 * TKR11;
    double TKR20 = TKR10 * TKR10;
    (void)TKR9;
    double PNMR = p / th->pcritical;
    double PNMR2 = PNMR * PNMR;
    double PNMR3 = PNMR * PNMR2;
    double PNMR4 = PNMR * PNMR3;
    double Y = 1.0 - SA1 * TKR2 - SA2 / TKR6;
    double ZP = SA3 * Y * Y - 2.0 * SA4 * TKR + 2.0 * SA5 * PNMR;
    if (ZP >= 0.0) {
      double Z = Y + sqrt(ZP);
      double CZ = pow(Z, 5.0 / 17.0);
      double PAR1 = A12 * SA5 / CZ;
      double CC1 = SA6 - TKR;
      double CC2 = CC1 * CC1;
      double CC4 = CC2 * CC2;
      double CC8 = CC4 * CC4;
      double CC10 = CC2 * CC8;
      double AA1 = SA7 + TKR19;
      double PAR2 = A13 + A14 * TKR + A15 * TKR2 + A16 * CC10 + A17 / AA1;
      double PAR3 = (A18 + 2.0 * A19 * PNMR + 3.0 * A20 * PNMR2) / (SA8 + TKR11);
      double DD1 = SA10 + PNMR;
      double DD2 = DD1 * DD1;
      double DD4 = DD2 * DD2;
      double PAR4 = A21 * TKR18 * (SA9 + TKR2) * (-3.0 / DD4 + SA11);
      double PAR5 = 3.0 * A22 * (SA12 - TKR) * PNMR2 + 4.0 * A23 / TKR20 * PNMR3;
      double VMKR = PAR1 + PAR2 - PAR3 - PAR4 + PAR5;
      double V = VMKR * 3.17e-3;
      double D = 1.0 / V;
      double YD = -2.0 * SA1 * TKR + 6.0 * SA2 / TKR7;
      double SNUM = A10 + A11 * TKR;
      SNUM = SNUM * TKR + A9;
      SNUM = SNUM * TKR + A8;
      SNUM = SNUM * TKR + A7;
      SNUM = SNUM * TKR + A6;
      SNUM = SNUM * TKR + A5;
      SNUM = SNUM * TKR + A4;
      SNUM = SNUM * TKR2 - A2;
      double PRT1 = A12 * (Z * (17.0 * (Z / 29.0 - Y / 12.0) + 5.0 * TKR * YD / 12.0) + SA4 * TKR -
                           (SA3 - 1.0) * TKR * Y * YD) / CZ;
      double PRT2 = PNMR * (A13 - A15 * TKR2 + A16 * (9.0 * TKR + SA6) * CC8 * CC1 +
                            A17 * (19.0 * TKR19 + AA1) / (AA1 * AA1));
      double BB1 = SA8 + TKR11;
      double BB2 = BB1 * BB1;
      double PRT3 = (11.0 * TKR11 + BB1) / BB2 * (A18 * PNMR + A19 * PNMR2 + A20 * PNMR3);
      double EE1 = SA10 + PNMR;
      double EE3 = EE1 * EE1 * EE1;
      double PRT4 = A21 * TKR18 * (17.0 * SA9 + 19.0 * TKR2) * (1.0 / EE3 + SA11 * PNMR);
      double PRT5 = A22 * SA12 * PNMR3 + 21.0 * A23 / TKR20 * PNMR4;
      double ENTR = A1 * TKR - SNUM + PRT1 + PRT2 - PRT3 + PRT4 + PRT5;
      double H = ENTR * 70120.4;
      double U = H - p * V;
      props[0] = D;
      props[1] = U;
      return 0;
    }
    return 1;
  }
  return 1;
}

/* IFC67.F90:378-396 */
static double ifc67_region1_viscosity(const wo_thermo *th, double temperature, double pressure) {
  double ex = 247.8 / (temperature + 133.15);
  double phi = 1.0467 * (temperature - 31.85);
  double ps = 0.0;
  ifc67_sat_pressure(th, temperature, &ps);
  double am = 1.0 + phi * (pressure - ps) * 1.0e-11;
  return 1.0e-7 * am * 241.4 * pow(10.0, ex);
}

/* IFC67.F90:425-576 */
static int ifc67_region2(const wo_thermo *th, const double param[2], double props[2]) {
  const double B0 = 16.83599274, B01 = 28.56067796, B02 = -54.38923329, B03 = 0.4330662834,
               B04 = -0.6547711697, B05 = 8.565182058e-2, B11 = 6.670375918e-2, B12 = 1.388983801,
               B21 = 8.390104328e-2, B22 = 2.614670893e-2, B23 = -3.373439453e-2, B31 = 4.520918904e-1,
               B32 = 1.069036614e-1, B41 = -5.975336707e-1, B42 = -8.847535804e-2, B51 = 5.958051609e-1,
               B52 = -5.159303373e-1, B53 = 2.075021122e-1, B61 = 1.190610271e-1, B62 = -9.867174132e-2,
               B71 = 1.683998803e-1, B72 = -5.809438001e-2, B81 = 6.552390126e-3, B82 = 5.710218649e-4,
               B90 = 1.936587558e2, B91 = -1.388522425e3, B92 = 4.126607219e3, B93 = -6.508211677e3,
               B94 = 5.745984054e3, B95 = -2.693088365e3, B96 = 5.235718623e2;
  const double SB = 7.633333333e-1, SB61 = 4.006073948e-1, SB71 = 8.636081627e-2,
               SB81 = -8.532322921e-1, SB82 = 3.460208861e-1;
  (void)B02;
  double P = param[0], T = param[1];
  if (T <= 800.0 && P <= 100.e6) {
    double THETA = (T + WO_TC_K) / th->tcriticalk;
    double BETA = P / th->pcritical;
    double RI1 = 4.260321148;
    double X = exp(SB * (1.0 - THETA));
    double X2 = X * X;
    double X3 = X2 * X;
    double X4 = X3 * X;
    double X5 = X4 * X;
    double X6 = X5 * X;
    double X8 = X6 * X2;
    double X10 = X6 * X4;
    double X11 = X10 * X;
    double X14 = X10 * X4;
    double X18 = X14 * X4;
    double X19 = X18 * X;
    double X24 = X18 * X6;
    double X27 = X24 * X3;
    double THETA2 = THETA * THETA;
    double THETA3 = THETA2 * THETA;
    double THETA4 = THETA3 * THETA;
    double BETA2 = BETA * BETA;
    double BETA3 = BETA2 * BETA;
    double BETA4 = BETA3 * BETA;
    double BETA5 = BETA4 * BETA;
    double BETA6 = BETA5 * BETA;
    double BETA7 = BETA6 * BETA;
    double BETAL = 15.74373327 - 34.17061978 * THETA + 19.31380707 * THETA2;
    double DBETAL = -34.17061978 + 38.62761414 * THETA;
    double R = BETA / BETAL;
    double R2 = R * R;
    double R4 = R2 * R2;
    double R6 = R4 * R2;
    double R10 = R6 * R4;
    double CHI2 = RI1 * THETA / BETA;
    double SC = (B11 * X10 + B12) * X3;
    CHI2 = CHI2 - SC;
    SC = B21 * X18 + B22 * X2 + B23 * X;
    CHI2 = CHI2 - 2.0 * BETA * SC;
    SC = (B31 * X8 + B32) * X10;
    CHI2 = CHI2 - 3.0 * BETA2 * SC;
    SC = (B41 * X11 + B42) * X14;
    CHI2 = CHI2 - 4.0 * BETA3 * SC;
    SC = (B51 * X8 + B52 * X4 + B53) * X24;
    CHI2 = CHI2 - 5.0 * BETA4 * SC;
    double SD1 = 1.0 / BETA4 + SB61 * X14;
    double SD2 = 1.0 / BETA5 + SB71 * X19;
    double SD3 = 1.0 / BETA6 + (SB81 * X27 + SB82) * X27;
    double SD12 = SD1 * SD1;
    double SD22 = SD2 * SD2;
    double SD32 = SD3 * SD3;
    double SN = (B61 * X + B62) * X11;
    CHI2 = CHI2 - SN / SD12 * 4.0 / BETA5;
    SN = (B71 * X6 + B72) * X18;
    CHI2 = CHI2 - SN / SD22 * 5.0 / BETA6;
    SN = (B81 * X10 + B82) * X14;
    CHI2 = CHI2 - SN / SD32 * 6.0 / BETA7;
    SC = B96;
    SC = SC * X + B95;
    SC = SC * X + B94;
    SC = SC * X + B93;
    SC = SC * X + B92;
    SC = SC * X + B91;
    SC = SC * X + B90;
    CHI2 = CHI2 + 11.0 * R10 * SC;
    double V = CHI2 * 0.00317;
    double D = 1.0 / V;
    double OS1 = SB * THETA;
    double EPS2 = B0 * THETA - (-B01 + B03 * THETA2 + 2.0 * B04 * THETA3 + 3.0 * B05 * THETA4);
    SC = (B11 * (1.0 + 13.0 * OS1) * X10 + B12 * (1.0 + 3.0 * OS1)) * X3;
    EPS2 = EPS2 - BETA * SC;
    SC = B21 * (1.0 + 18.0 * OS1) * X18 + B22 * (1.0 + 2.0 * OS1) * X2 + B23 * (1.0 + OS1) * X;
    EPS2 = EPS2 - BETA2 * SC;
    SC = (B31 * (1.0 + 18.0 * OS1) * X8 + B32 * (1.0 + 10.0 * OS1)) * X10;
    EPS2 = EPS2 - BETA3 * SC;
    SC = (B41 * (1.0 + 25.0 * OS1) * X11 + B42 * (1.0 + 14.0 * OS1)) * X14;
    EPS2 = EPS2 - BETA4 * SC;
    SC = (B51 * (1.0 + 32.0 * OS1) * X8 + B52 * (1.0 + 28.0 * OS1) * X4 + B53 * (1.0 + 24.0 * OS1)) * X24;
    EPS2 = EPS2 - BETA5 * SC;
    double SN6 = 14.0 * SB61 * X14;
    double SN7 = 19.0 * SB71 * X19;
    double SN8 = (54.0 * SB81 * X27 + 27.0 * SB82) * X27;
    double OS5 = 1.0 + 11.0 * OS1 - OS1 * SN6 / SD1;
    SC = (B61 * X * (OS1 + OS5) + B62 * OS5) * (X11 / SD1);
    EPS2 = EPS2 - SC;
    double OS6 = 1.0 + 24.0 * OS1 - OS1 * SN7 / SD2;
    SC = (B71 * X6 * OS6 + B72 * (OS6 - 6.0 * OS1)) * (X18 / SD2);
    EPS2 = EPS2 - SC;
    double OS7 = 1.0 + 24.0 * OS1 - OS1 * SN8 / SD3;
    SC = (B81 * X10 * OS7 + B82 * (OS7 - 10.0 * OS1)) * (X14 / SD3);
    EPS2 = EPS2 - SC;
    double OS2 = 1.0 + THETA * 10.0 * DBETAL / BETAL;
    SC = (OS2 + 6.0 * OS1) * B96;
    SC = SC * X + (OS2 + 5.0 * OS1) * B95;
    SC = SC * X + (OS2 + 4.0 * OS1) * B94;
    SC = SC * X + (OS2 + 3.0 * OS1) * B93;
    SC = SC * X + (OS2 + 2.0 * OS1) * B92;
    SC = SC * X + (OS2 + OS1) * B91;
    SC = SC * X + OS2 * B90;
    EPS2 = EPS2 + BETA * R10 * SC;
    double H = EPS2 * 70120.4;
    double U = H - P * V;
    props[0] = D;
    props[1] = U;
    return 0;
  }
  return 1;
}

/* IFC67.F90:580-600 */
static double ifc67_region2_viscosity(double temperature, double density) {
  double v1 = 0.407 * temperature + 80.4;
  if (temperature <= 350.0)
    return 1.0e-7 * (v1 - density * (1858.0 - 5.9 * temperature) * 1.0e-3);
  return 1.0e-7 * (v1 + density * (0.353 + density * (676.5e-6 + density * 102.1e-9)));
}

/* IFC67.F90:200-222 */
static int ifc67_phase_composition(int region) {
  switch (region) {
    case 1: return 1;
    case 2: return 2;
    case 4: return 3;
    default: return 0;
  }
}

/* ------------------------------------------------------------------ */
/* dispatch                                                            */
/* ------------------------------------------------------------------ */

int wo_region_properties(wo_thermo *th, int region, const double param[2], double props[2]) {
  if (th->id == WO_THERMO_IAPWS) {
    if (region == 1) return iapws_region1(th, param, props);
    if (region == 2) return iapws_region2(th, param, props);
    if (region == 3) return iapws_region3(th, param, props);
  } else {
    if (region == 1) return ifc67_region1(th, param, props);
    if (region == 2) return ifc67_region2(th, param, props);
  }
  props[0] = props[1] = 0.0;
  return 1;
}

double wo_region_viscosity(wo_thermo *th, int region, double temperature, double pressure, double density) {
  if (th->id == WO_THERMO_IAPWS) return iapws_viscosity(th, temperature, density);
  if (region == 1) return ifc67_region1_viscosity(th, temperature, pressure);
  return ifc67_region2_viscosity(temperature, density);
}

int wo_saturation_pressure(const wo_thermo *th, double t, double *p) {
  return th->id == WO_THERMO_IAPWS ? iapws_sat_pressure(th, t, p) : ifc67_sat_pressure(th, t, p);
}
int wo_saturation_temperature(const wo_thermo *th, double p, double *t) {
  return th->id == WO_THERMO_IAPWS ? iapws_sat_temperature(th, p, t) : ifc67_sat_temperature(th, p, t);
}
int wo_phase_composition(const wo_thermo *th, int region, double pressure, double temperature) {
  return th->id == WO_THERMO_IAPWS ? iapws_phase_composition(th, region, pressure, temperature)
                                   : ifc67_phase_composition(region);
}
