/*
 * wo_curves.c -- oracle (TEST INFRASTRUCTURE): linear interpolation tables,
 * Brent root finder, relative permeability and capillary pressure curves.
 * Restated from src/interpolation.F90:202-581, src/root_finder.F90:127-248,
 * src/relative_permeability.F90:197-558, src/capillary_pressure.F90:159-358.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- interpolation table (linear) ---- */

/* interpolation.F90:309-353: coordinates sorted at init */
void wo_table_init(wo_table *t, const double *x, const double *v, int n, int dim) {
  t->n = n;
  t->dim = dim;
  t->index = 1;
  t->x = (double *)malloc(n * sizeof(double));
  t->val = (double *)malloc((size_t)n * dim * sizeof(double));
  int *perm = (int *)malloc(n * sizeof(int));
  for (int i = 0; i < n; i++) perm[i] = i;
  for (int i = 1; i < n; i++) { /* stable insertion sort */
    int p = perm[i], j = i - 1;
    while (j >= 0 && x[perm[j]] > x[p]) {
      perm[j + 1] = perm[j];
      j--;
    }
    perm[j + 1] = p;
  }
  for (int i = 0; i < n; i++) {
    t->x[i] = x[perm[i]];
    for (int d = 0; d < dim; d++) t->val[d + dim * i] = v[d + dim * perm[i]];
  }
  free(perm);
}

void wo_table_destroy(wo_table *t) {
  free(t->x);
  free(t->val);
  memset(t, 0, sizeof(*t));
}

/* interpolation.F90:202-306 (1-based index semantics kept: index in [0, n]) */
void wo_table_find(wo_table *t, double x) {
  const double *val = t->x - 1; /* 1-based */
  int size = t->n;
  if (x <= val[1]) {
    t->index = 0;
  } else if (x >= val[size]) {
    t->index = size;
  } else {
    int i1, i2;
    /* bracket */
    if (t->index >= 1 && t->index <= size) {
      i1 = t->index;
      int inc = 1, found = 0;
      if (x >= val[i1]) {
        i2 = i1;
        while (!found) {
          i2 = i1 + inc;
          if (i2 > size) {
            i2 = size + 1;
            found = 1;
          } else if (x >= val[i2]) {
            i1 = i2;
            inc = inc + inc;
          } else
            found = 1;
        }
      } else {
        i2 = i1;
        while (!found) {
          i1 = i2 - inc;
          if (i1 < 1) {
            i1 = 0;
            found = 1;
          } else if (x < val[i1]) {
            i2 = i1;
            inc = inc + inc;
          } else
            found = 1;
        }
      }
    } else {
      i1 = 0;
      i2 = size + 1;
    }
    /* bisect */
    while (i2 - i1 > 1) {
      int im = (i1 + i2) / 2;
      if (x >= val[im]) i1 = im;
      else i2 = im;
    }
    t->index = i1;
  }
}

/* interpolation.F90:388-403, 494-510 */
void wo_table_interpolate_at_index(const wo_table *t, double x, double *y) {
  int dim = t->dim;
  if (t->index <= 0) {
    for (int d = 0; d < dim; d++) y[d] = t->val[d];
  } else if (t->index >= t->n) {
    for (int d = 0; d < dim; d++) y[d] = t->val[d + dim * (t->n - 1)];
  } else {
    int i = t->index - 1; /* 0-based lower point */
    double xi = (x - t->x[i]) / (t->x[i + 1] - t->x[i]);
    for (int d = 0; d < dim; d++) y[d] = (1.0 - xi) * t->val[d + dim * i] + xi * t->val[d + dim * (i + 1)];
  }
}

void wo_table_interpolate(wo_table *t, double x, double *y) {
  wo_table_find(t, x);
  wo_table_interpolate_at_index(t, x, y);
}

/* interpolation.F90:407-438, 563-581; component is 1-based */
int wo_table_find_component_at_index(const wo_table *t, double yi, int component, double *x) {
  const double tol = 1.e-8;
  if (t->index <= 0 || t->index >= t->n) return 1;
  int i = t->index - 1;
  double v1 = t->val[(component - 1) + t->dim * i];
  double v2 = t->val[(component - 1) + t->dim * (i + 1)];
  double vmax = fmax(fabs(v1), fabs(v2));
  if (fabs(v2 - v1) >= tol * vmax) {
    double vs1 = v1 / vmax, vs2 = v2 / vmax;
    double ys = yi / vmax;
    double xi = (ys - vs1) / (vs2 - vs1);
    *x = (1.0 - xi) * t->x[i] + xi * t->x[i + 1];
    return 0;
  }
  return 1;
}

/* ---- Brent root finder: root_finder.F90:63-248 ---- */

void wo_root_finder_init(wo_root_finder *r) {
  r->interval[0] = 0.0;
  r->interval[1] = 1.0;
  r->root_tolerance = 1.e-8;
  r->function_tolerance = 1.e-8;
  r->max_iterations = 100;
  r->iterations = 0;
  r->root = 0.0;
  r->err = 0;
}

void wo_root_finder_find(wo_root_finder *self, wo_root_fn f, void *ctx) {
  double a, b, c, d = 0.0, e = 0.0, fa, fb, fc, dx, p, pc, q, r, s;
  int iter, found = 0;
  const double small = 1.e-16;
  self->iterations = 0;
  self->err = 0;
  self->root = 0.0;
  a = self->interval[0];
  b = self->interval[1];
  fa = f(a, ctx);
  fb = f(b, ctx);
  if (fa * fb > 0.0) {
    self->err = 1; /* ROOT_FINDER_INTERVAL_NOT_BRACKETED */
    return;
  }
  c = b;
  fc = fb;
  for (iter = 1; iter <= self->max_iterations; iter++) {
    if (fb * fc > 0.0) {
      c = a;
      fc = fa;
      d = b - a;
      e = d;
    }
    if (fabs(fc) < fabs(fb)) {
      a = b;
      b = c;
      c = a;
      fa = fb;
      fb = fc;
      fc = fa;
    }
    dx = 0.5 * (c - b);
    if (fabs(dx) <= self->root_tolerance || fabs(fb) <= self->function_tolerance) {
      found = 1;
      break;
    }
    if (fabs(e) >= self->root_tolerance && fabs(fa) > fabs(fb)) {
      s = fb / fa;
      if (fabs(a - c) <= small) {
        p = 2.0 * dx * s;
        q = 1.0 - s;
      } else {
        q = fa / fc;
        r = fb / fc;
        p = s * (2.0 * dx * q * (q - r) - (b - a) * (r - 1.0));
        q = (q - 1.0) * (r - 1.0) * (s - 1.0);
      }
      if (p > 0.0) q = -q;
      else p = -p;
      pc = fmin(3.0 * dx * q - fabs(self->root_tolerance * q), fabs(e * q));
      if (2.0 * p < pc) {
        e = d;
        d = p / q;
      } else {
        d = dx;
        e = d;
      }
    } else {
      d = dx;
      e = d;
    }
    a = b;
    fa = fb;
    if (fabs(d) > self->root_tolerance) b = b + d;
    else b = b + copysign(self->root_tolerance, dx);
    fb = f(b, ctx);
  }
  self->root = b;
  self->iterations = iter;
  if (!found) self->err = 2; /* ROOT_FINDER_ITERATIONS_EXCEEDED */
}

/* ---- relative permeability: relative_permeability.F90 ---- */

static double table2_interp(const double *xs, const double *ys, int n, double x) {
  wo_table t;
  double y;
  wo_table_init(&t, xs, ys, n, 1);
  wo_table_interpolate(&t, x, &y);
  wo_table_destroy(&t);
  return y;
}

void wo_relperm_values(const wo_relperm *rp, double sl, double out[2]) {
  switch (rp->type) {
    case WO_RP_FULLY_MOBILE: /* :197 */
      out[0] = 1.0;
      out[1] = 1.0;
      break;
    case WO_RP_LINEAR: { /* :213-259 */
      double lx[2] = {rp->p[0], rp->p[1]}, ly[2] = {0.0, 1.0};
      double vx[2] = {rp->p[2], rp->p[3]}, vy[2] = {0.0, 1.0};
      out[0] = table2_interp(lx, ly, 2, sl);
      out[1] = table2_interp(vx, vy, 2, 1.0 - sl);
      break;
    }
    case WO_RP_PICKENS: /* :297-308 */
      out[0] = pow(sl, rp->p[0]);
      out[1] = 1.0;
      break;
    case WO_RP_COREY: { /* :349-370 */
      double slr = rp->p[0], ssr = rp->p[1];
      double sv = 1.0 - sl;
      if (sv < ssr) {
        out[0] = 1.0;
        out[1] = 0.0;
      } else if (sv > 1.0 - slr) {
        out[0] = 0.0;
        out[1] = 1.0;
      } else {
        double sstar = (sl - slr) / (1.0 - slr - ssr);
        double sstar2 = sstar * sstar;
        out[0] = sstar2 * sstar2;
        out[1] = (1.0 - 2.0 * sstar + sstar2) * (1.0 - sstar2);
      }
      break;
    }
    case WO_RP_GRANT: { /* :399-420 */
      double slr = rp->p[0], ssr = rp->p[1];
      double sv = 1.0 - sl;
      if (sv < ssr) {
        out[0] = 1.0;
        out[1] = 0.0;
      } else if (sv > 1.0 - slr) {
        out[0] = 0.0;
        out[1] = 1.0;
      } else {
        double sstar = (sl - slr) / (1.0 - slr - ssr);
        double sstar2 = sstar * sstar;
        out[0] = sstar2 * sstar2;
        out[1] = 1.0 - out[0];
      }
      break;
    }
    case WO_RP_VAN_GENUCHTEN: { /* :461-491 */
      double lambda = rp->p[0], slr = rp->p[1], sls = rp->p[2], ssr = rp->p[4];
      int sum_unity = rp->p[3] != 0.0;
      double sstar = (sl - slr) / (sls - slr);
      if (sstar < 0.0) out[0] = 0.0;
      else if (sstar < 1.0) {
        double b = 1.0 - pow(1.0 - pow(sstar, 1.0 / lambda), lambda);
        out[0] = sqrt(sstar) * (b * b);
      } else
        out[0] = 1.0;
      if (sum_unity) out[1] = 1.0 - out[0];
      else {
        double s_hat = (sl - slr) / (1.0 - slr - ssr);
        double s_hat2 = s_hat * s_hat;
        out[1] = (1.0 - 2.0 * s_hat + s_hat2) * (1.0 - s_hat2);
        out[1] = fmin(1.0, out[1]);
      }
      break;
    }
    case WO_RP_TABLE: /* :547-558 (linear interpolation only) */
      out[0] = table2_interp(rp->lx, rp->ly, rp->nl, sl);
      out[1] = table2_interp(rp->vx, rp->vy, rp->nv, 1.0 - sl);
      break;
    default:
      out[0] = out[1] = 0.0;
  }
}

/* ---- capillary pressure: capillary_pressure.F90 ---- */

double wo_cappress_value(const wo_cappress *cp, double sl, double t) {
  (void)t;
  switch (cp->type) {
    case WO_CP_ZERO: /* :159 */
      return 0.0;
    case WO_CP_LINEAR: { /* :176-218 */
      double pressure = fabs(cp->p[2]);
      double xs[2] = {cp->p[0], cp->p[1]}, ys[2] = {-pressure, 0.0};
      return table2_interp(xs, ys, 2, sl);
    }
    case WO_CP_VAN_GENUCHTEN: { /* :273-305 */
      const double eps = 1.e-3;
      double P0 = fabs(cp->p[0]), lambda = cp->p[1], slr = cp->p[2], sls = cp->p[3];
      double Pmax = fabs(cp->p[4]);
      int apply_Pmax = cp->p[5] != 0.0;
      double c;
      if (sl < 1.0) {
        double sstar = (sl - slr) / (sls - slr);
        if (sstar < 0.0) c = -Pmax;
        else if (sstar < 1.0) c = -P0 * pow(pow(sstar, -1.0 / lambda) - 1.0, 1.0 - lambda);
        else c = 0.0;
        c = fmin(0.0, c);
        if (apply_Pmax) c = fmax(-Pmax, c);
        if (sl > 1.0 - eps) c = c * (1.0 - sl) / eps;
      } else
        c = 0.0;
      return c;
    }
    case WO_CP_TABLE: /* :349-358 */
      return table2_interp(cp->x, cp->y, cp->n, sl);
    default:
      return 0.0;
  }
}
