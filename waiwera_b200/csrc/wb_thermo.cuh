// wb_thermo.cuh -- stateless water/steam thermodynamics for device code.
//
// Computes what the reference's thermodynamics objects compute
// (src/IAPWS.F90, src/IFC67.F90, src/thermodynamics.F90) but with no mutable
// scratch (the reference's powertable / interpolation caches make its objects
// non re-entrant): everything lives in registers, so a
// thread can evaluate any cell, any region, any number of times.
//
// All quantities SI; temperatures in deg C as in the reference.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define WB_HD __host__ __device__ __forceinline__
#define WB_HDN __host__ __device__ __noinline__
#else
#define WB_HD static inline
#define WB_HDN
#endif

#include "wb_iapws_gen.cuh"

#define WB_RCONST 0.461526e3 /* src/thermodynamics.F90:36 */
#define WB_TC_K 273.15       /* src/thermodynamics.F90:38 */

#define WB_THERMO_IAPWS 0
#define WB_THERMO_IFC67 1

struct WbThermo {
  int id;
  double tcriticalk, tcritical, pcritical, dcritical;
  double r1_max_temperature;
};

// src/IAPWS.F90:273-276, 456-483 ; src/IFC67.F90:156-160, 235-249
inline WbThermo wb_thermo_make(int id, int extrapolate) {
  WbThermo th;
  th.id = id;
  th.r1_max_temperature = extrapolate ? 360.0 : 350.0;
  if (id == WB_THERMO_IAPWS) {
    th.tcriticalk = 647.096;
    th.pcritical = 22.064e6;
  } else {
    th.tcriticalk = 647.3;
    th.pcritical = 22.12e6;
  }
  th.tcritical = th.tcriticalk - WB_TC_K;
  th.dcritical = 322.0;
  return th;
}

// ---------------------------------------------------------------- IAPWS-97

// src/IAPWS.F90:503-542
WB_HD int wb_iapws_region1(const WbThermo &th, double p, double t, double &rho, double &u) {
  if (t <= th.r1_max_temperature && p <= 100.e6) {
    const double pstar = 16.53e6, tstar = 1386.0;
    const double tk = t + WB_TC_K;
    const double rt = WB_RCONST * tk;
    const double pi = p / pstar;
    const double tau = tstar / tk;
    double s1, s2;
    wb_iapws_r1_sums(7.1 - pi, tau - 1.222, s1, s2);
    const double gampi = -s1, gamt = s2;
    rho = pstar / (rt * gampi);
    u = rt * (tau * gamt - pi * gampi);
    return 0;
  }
  return 1;
}

// src/IAPWS.F90:596-639
WB_HD int wb_iapws_region2(const WbThermo &th, double p, double t, double &rho, double &u) {
  if (t <= 800.0 && p <= 100.e6) {
    const double pstar = 1.0e6, tstar = 540.0;
    const double tk = t + WB_TC_K;
    const double rt = WB_RCONST * tk;
    const double pi = p / pstar;
    const double tau = tstar / tk;
    double gamt0, gampir, gamtr, pim1;
    wb_iapws_r2_sums(pi, tau, tau - 0.5, gamt0, gampir, gamtr, pim1);
    const double gampi = pim1 + gampir;
    rho = pstar / (rt * gampi);
    u = rt * (tau * (gamt0 + gamtr) - pi * gampi);
    return 0;
  }
  return 1;
}

// src/IAPWS.F90:689-727 : (density, temperature) -> (pressure, internal energy)
WB_HD int wb_iapws_region3(const WbThermo &th, double d, double t, double &p, double &u) {
  const double tk = t + WB_TC_K;
  const double rt = WB_RCONST * tk;
  const double tau = th.tcriticalk / tk;
  const double delta = d / th.dcritical;
  double s1, s2, dm1;
  wb_iapws_r3_sums(delta, tau, s1, s2, dm1);
  const double phidelta = 0.10658070028513e1 * dm1 + s1;
  p = d * rt * delta * phidelta;
  u = rt * tau * s2;
  return (p > 100.0e6) ? 1 : 0;
}

// src/IAPWS.F90:412-443
WB_HD double wb_iapws_viscosity(const WbThermo &th, double temperature, double density) {
  const double tk = temperature + WB_TC_K;
  const double tau = tk / th.tcriticalk;
  const double del = density / th.dcritical;
  double s0, s1;
  wb_iapws_visc_sums(1.0 / tau, del - 1.0, s0, s1);
  const double mu0 = 100.0 * sqrt(tau) / s0;
  const double mu1 = exp(del * s1);
  return 1.0e-6 * mu0 * mu1;
}

#define WB_SAT_N1 0.11670521452767e4
#define WB_SAT_N2 -0.72421316703206e6
#define WB_SAT_N3 -0.17073846940092e2
#define WB_SAT_N4 0.12020824702470e5
#define WB_SAT_N5 -0.32325550322333e7
#define WB_SAT_N6 0.14915108613530e2
#define WB_SAT_N7 -0.48232657361591e4
#define WB_SAT_N8 0.40511340542057e6
#define WB_SAT_N9 -0.23855557567849
#define WB_SAT_N10 0.65017534844798e3

// src/IAPWS.F90:762-789
WB_HD int wb_iapws_sat_pressure(const WbThermo &th, double t, double &p) {
  if (t >= 0.0 && t <= th.tcritical) {
    const double tk = t + WB_TC_K;
    const double theta = tk + WB_SAT_N9 / (tk - WB_SAT_N10);
    const double theta2 = theta * theta;
    const double a = theta2 + WB_SAT_N1 * theta + WB_SAT_N2;
    const double b = WB_SAT_N3 * theta2 + WB_SAT_N4 * theta + WB_SAT_N5;
    const double c = WB_SAT_N6 * theta2 + WB_SAT_N7 * theta + WB_SAT_N8;
    double x = 2.0 * c / (-b + sqrt(b * b - 4.0 * a * c));
    x = x * x;
    p = 1.0e6 * x * x;
    return 0;
  }
  return 1;
}

// src/IAPWS.F90:793-818
WB_HD int wb_iapws_sat_temperature(const WbThermo &th, double p, double &t) {
  if (p >= 611.213 && p <= th.pcritical) {
    const double beta2 = sqrt(p / 1.0e6);
    const double beta = sqrt(beta2);
    const double e = beta2 + WB_SAT_N3 * beta + WB_SAT_N6;
    const double f = WB_SAT_N1 * beta2 + WB_SAT_N4 * beta + WB_SAT_N7;
    const double g = WB_SAT_N2 * beta2 + WB_SAT_N5 * beta + WB_SAT_N8;
    const double d = 2.0 * g / (-f - sqrt(f * f - 4.0 * e * g));
    const double x = WB_SAT_N10 + d;
    t = 0.5 * (WB_SAT_N10 + d - sqrt(x * x - 4.0 * (WB_SAT_N9 + WB_SAT_N10 * d))) - WB_TC_K;
    return 0;
  }
  return 1;
}

// src/IAPWS.F90:317-365
WB_HD int wb_iapws_phase_composition(const WbThermo &th, int region, double pressure, double temperature) {
  int phases = 0;
  if (region == 4) {
    phases = 3;
  } else if (temperature <= th.tcritical) {
    if (region == 1) phases = 1;
    else if (region == 2) phases = 2;
    else if (region == 3) {
      double ps;
      if (wb_iapws_sat_pressure(th, temperature, ps) == 0) phases = (pressure >= ps) ? 1 : 2;
      else phases = -1;
    }
  } else {
    phases = (pressure <= th.pcritical) ? 2 : 4;
  }
  return phases;
}

// ---------------------------------------------------------------- IFC-67

// src/IFC67.F90:606-633
WB_HD int wb_ifc67_sat_pressure(const WbThermo &th, double t, double &p) {
  const double A1 = -7.691234564, A2 = -2.608023696e1, A3 = -1.681706546e2, A4 = 6.423285504e1,
               A5 = -1.189646225e2, A6 = 4.167117320, A7 = 2.097506760e1, A8 = 1.0e9, A9 = 6.0;
  if (t >= 1.0 && t <= th.tcritical) {
    const double TC = (t + WB_TC_K) / th.tcriticalk;
    const double X1 = 1.0 - TC;
    const double X2 = X1 * X1;
    double SC = A5 * X1 + A4;
    SC = SC * X1 + A3;
    SC = SC * X1 + A2;
    SC = SC * X1 + A1;
    SC = SC * X1;
    const double PC = exp(SC / (TC * (1.0 + A6 * X1 + A7 * X2)) - X1 / (A8 * X2 + A9));
    p = PC * th.pcritical;
    return 0;
  }
  return 1;
}

// src/IFC67.F90:637-676 : Newton iteration with a forward-difference slope
// (src/utils.F90:651-709), data-dependent trip count.
WB_HD int wb_ifc67_sat_temperature(const WbThermo &th, double p, double &tout) {
  const int maxit = 200;
  const double ftol = 1.e-10, xtol = 1.e-10, inc = 1.e-8;
  if (p >= 0.0061e5 && p <= th.pcritical) {
    double x = fmax(4606.0 / (24.02 - log(p)) - WB_TC_K, 5.0);
    const double ftolp = ftol * p;
    const double delx = inc * x;
    int found = 0, err = 0;
    for (int i = 1; i <= maxit; i++) {
      double ps;
      err = wb_ifc67_sat_pressure(th, x, ps);
      if (err) break;
      const double fx = p - ps;
      if (fabs(fx) <= ftolp) {
        found = 1;
        break;
      }
      err = wb_ifc67_sat_pressure(th, x + delx, ps);
      if (err) break;
      const double fxd = p - ps;
      const double df = (fxd - fx) / delx;
      const double dx = -fx / df;
      x = x + dx;
      if (fabs(dx) <= xtol) {
        found = 1;
        break;
      }
    }
    if (err == 0 && !found) err = 1;
    tout = x;
    return err;
  }
  return 1;
}

// src/IFC67.F90:265-374
WB_HD int wb_ifc67_region1(const WbThermo &th, double p, double t, double &rho, double &uout) {
  const double A1 = 6.824687741e3, A2 = -5.422063673e2, A4 = 3.941286787e4, A5 = -13.466555478e4,
               A6 = 29.707143084e4, A7 = -4.375647096e5, A8 = 42.954208335e4, A9 = -27.067012452e4,
               A10 = 9.926972482e4, A11 = -16.138168904e3, A12 = 7.982692717, A13 = -2.616571843e-2,
               A14 = 1.522411790e-3, A15 = 2.284279054e-2, A16 = 2.421647003e2, A17 = 1.269716088e-10,
               A18 = 2.074838328e-7, A19 = 2.174020350e-8, A20 = 1.105710498e-9, A21 = 1.293441934e1,
               A22 = 1.308119072e-5, A23 = 6.047626338e-14;
  const double SA1 = 8.438375405e-1, SA2 = 5.362162162e-4, SA3 = 1.72, SA4 = 7.342278489e-2,
               SA5 = 4.975858870e-2, SA6 = 6.537154300e-1, SA7 = 1.150e-6, SA8 = 1.51080e-5,
               SA9 = 1.41880e-1, SA10 = 7.002753165, SA11 = 2.995284926e-4, SA12 = 2.040e-1;
  if (t <= th.r1_max_temperature && p <= 100.e6) {
    const double TKR = (t + WB_TC_K) / th.tcriticalk;
    const double TKR2 = TKR * TKR;
    const double TKR3 = TKR * TKR2;
    const double TKR4 = TKR2 * TKR2;
    const double TKR6 = TKR4 * TKR2;
    const double TKR7 = TKR4 * TKR3;
    const double TKR8 = TKR4 * TKR4;
    const double TKR10 = TKR4 * TKR6;
    const double TKR11 = TKR * TKR10;
    const double TKR18 = TKR8 * TKR10;
    const double TKR19 = TKR8 * TKR11;
    const double TKR20 = TKR10 * TKR10;
    const double PNMR = p / th.pcritical;
    const double PNMR2 = PNMR * PNMR;
    const double PNMR3 = PNMR * PNMR2;
    const double PNMR4 = PNMR * PNMR3;
    const double Y = 1.0 - SA1 * TKR2 - SA2 / TKR6;
    const double ZP = SA3 * Y * Y - 2.0 * SA4 * TKR + 2.0 * SA5 * PNMR;
    if (ZP >= 0.0) {
      const double Z = Y + sqrt(ZP);
      const double CZ = pow(Z, 5.0 / 17.0);
      const double PAR1 = A12 * SA5 / CZ;
      const double CC1 = SA6 - TKR;
      const double CC2 = CC1 * CC1;
      const double CC4 = CC2 * CC2;
      const double CC8 = CC4 * CC4;
      const double CC10 = CC2 * CC8;
      const double AA1 = SA7 + TKR19;
      const double PAR2 = A13 + A14 * TKR + A15 * TKR2 + A16 * CC10 + A17 / AA1;
      const double PAR3 = (A18 + 2.0 * A19 * PNMR + 3.0 * A20 * PNMR2) / (SA8 + TKR11);
      const double DD1 = SA10 + PNMR;
      const double DD2 = DD1 * DD1;
      const double DD4 = DD2 * DD2;
      const double PAR4 = A21 * TKR18 * (SA9 + TKR2) * (-3.0 / DD4 + SA11);
      const double PAR5 = 3.0 * A22 * (SA12 - TKR) * PNMR2 + 4.0 * A23 / TKR20 * PNMR3;
      const double VMKR = PAR1 + PAR2 - PAR3 - PAR4 + PAR5;
      const double V = VMKR * 3.17e-3;
      const double D = 1.0 / V;
      const double YD = -2.0 * SA1 * TKR + 6.0 * SA2 / TKR7;
      double SNUM = A10 + A11 * TKR;
      SNUM = SNUM * TKR + A9;
      SNUM = SNUM * TKR + A8;
      SNUM = SNUM * TKR + A7;
      SNUM = SNUM * TKR + A6;
      SNUM = SNUM * TKR + A5;
      SNUM = SNUM * TKR + A4;
      SNUM = SNUM * TKR2 - A2;
      const double PRT1 = A12 * (Z * (17.0 * (Z / 29.0 - Y / 12.0) + 5.0 * TKR * YD / 12.0) + SA4 * TKR -
                                 (SA3 - 1.0) * TKR * Y * YD) / CZ;
      const double PRT2 = PNMR * (A13 - A15 * TKR2 + A16 * (9.0 * TKR + SA6) * CC8 * CC1 +
                                  A17 * (19.0 * TKR19 + AA1) / (AA1 * AA1));
      const double BB1 = SA8 + TKR11;
      const double BB2 = BB1 * BB1;
      const double PRT3 = (11.0 * TKR11 + BB1) / BB2 * (A18 * PNMR + A19 * PNMR2 + A20 * PNMR3);
      const double EE1 = SA10 + PNMR;
      const double EE3 = EE1 * EE1 * EE1;
      const double PRT4 = A21 * TKR18 * (17.0 * SA9 + 19.0 * TKR2) * (1.0 / EE3 + SA11 * PNMR);
      const double PRT5 = A22 * SA12 * PNMR3 + 21.0 * A23 / TKR20 * PNMR4;
      const double ENTR = A1 * TKR - SNUM + PRT1 + PRT2 - PRT3 + PRT4 + PRT5;
      const double H = ENTR * 70120.4;
      rho = D;
      uout = H - p * V;
      return 0;
    }
    return 1;
  }
  return 1;
}

// src/IFC67.F90:378-396
WB_HD double wb_ifc67_region1_viscosity(const WbThermo &th, double temperature, double pressure) {
  const double ex = 247.8 / (temperature + 133.15);
  const double phi = 1.0467 * (temperature - 31.85);
  double ps = 0.0;
  wb_ifc67_sat_pressure(th, temperature, ps);
  const double am = 1.0 + phi * (pressure - ps) * 1.0e-11;
  return 1.0e-7 * am * 241.4 * pow(10.0, ex);
}

// src/IFC67.F90:425-576
WB_HD int wb_ifc67_region2(const WbThermo &th, double P, double T, double &rho, double &uout) {
  const double B0 = 16.83599274, B01 = 28.56067796, B03 = 0.4330662834, B04 = -0.6547711697,
               B05 = 8.565182058e-2, B11 = 6.670375918e-2, B12 = 1.388983801, B21 = 8.390104328e-2,
               B22 = 2.614670893e-2, B23 = -3.373439453e-2, B31 = 4.520918904e-1, B32 = 1.069036614e-1,
               B41 = -5.975336707e-1, B42 = -8.847535804e-2, B51 = 5.958051609e-1, B52 = -5.159303373e-1,
               B53 = 2.075021122e-1, B61 = 1.190610271e-1, B62 = -9.867174132e-2, B71 = 1.683998803e-1,
               B72 = -5.809438001e-2, B81 = 6.552390126e-3, B82 = 5.710218649e-4, B90 = 1.936587558e2,
               B91 = -1.388522425e3, B92 = 4.126607219e3, B93 = -6.508211677e3, B94 = 5.745984054e3,
               B95 = -2.693088365e3, B96 = 5.235718623e2;
  const double SB = 7.633333333e-1, SB61 = 4.006073948e-1, SB71 = 8.636081627e-2, SB81 = -8.532322921e-1,
               SB82 = 3.460208861e-1;
  if (T <= 800.0 && P <= 100.e6) {
    const double THETA = (T + WB_TC_K) / th.tcriticalk;
    const double BETA = P / th.pcritical;
    const double RI1 = 4.260321148;
    const double X = exp(SB * (1.0 - THETA));
    const double X2 = X * X;
    const double X3 = X2 * X;
    const double X4 = X3 * X;
    const double X5 = X4 * X;
    const double X6 = X5 * X;
    const double X8 = X6 * X2;
    const double X10 = X6 * X4;
    const double X11 = X10 * X;
    const double X14 = X10 * X4;
    const double X18 = X14 * X4;
    const double X19 = X18 * X;
    const double X24 = X18 * X6;
    const double X27 = X24 * X3;
    const double THETA2 = THETA * THETA;
    const double THETA3 = THETA2 * THETA;
    const double THETA4 = THETA3 * THETA;
    const double BETA2 = BETA * BETA;
    const double BETA3 = BETA2 * BETA;
    const double BETA4 = BETA3 * BETA;
    const double BETA5 = BETA4 * BETA;
    const double BETA6 = BETA5 * BETA;
    const double BETA7 = BETA6 * BETA;
    const double BETAL = 15.74373327 - 34.17061978 * THETA + 19.31380707 * THETA2;
    const double DBETAL = -34.17061978 + 38.62761414 * THETA;
    const double R = BETA / BETAL;
    const double R2 = R * R;
    const double R4 = R2 * R2;
    const double R6 = R4 * R2;
    const double R10 = R6 * R4;
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
    const double SD1 = 1.0 / BETA4 + SB61 * X14;
    const double SD2 = 1.0 / BETA5 + SB71 * X19;
    const double SD3 = 1.0 / BETA6 + (SB81 * X27 + SB82) * X27;
    const double SD12 = SD1 * SD1;
    const double SD22 = SD2 * SD2;
    const double SD32 = SD3 * SD3;
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
    const double V = CHI2 * 0.00317;
    const double D = 1.0 / V;
    const double OS1 = SB * THETA;
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
    const double SN6 = 14.0 * SB61 * X14;
    const double SN7 = 19.0 * SB71 * X19;
    const double SN8 = (54.0 * SB81 * X27 + 27.0 * SB82) * X27;
    const double OS5 = 1.0 + 11.0 * OS1 - OS1 * SN6 / SD1;
    SC = (B61 * X * (OS1 + OS5) + B62 * OS5) * (X11 / SD1);
    EPS2 = EPS2 - SC;
    const double OS6 = 1.0 + 24.0 * OS1 - OS1 * SN7 / SD2;
    SC = (B71 * X6 * OS6 + B72 * (OS6 - 6.0 * OS1)) * (X18 / SD2);
    EPS2 = EPS2 - SC;
    const double OS7 = 1.0 + 24.0 * OS1 - OS1 * SN8 / SD3;
    SC = (B81 * X10 * OS7 + B82 * (OS7 - 10.0 * OS1)) * (X14 / SD3);
    EPS2 = EPS2 - SC;
    const double OS2 = 1.0 + THETA * 10.0 * DBETAL / BETAL;
    SC = (OS2 + 6.0 * OS1) * B96;
    SC = SC * X + (OS2 + 5.0 * OS1) * B95;
    SC = SC * X + (OS2 + 4.0 * OS1) * B94;
    SC = SC * X + (OS2 + 3.0 * OS1) * B93;
    SC = SC * X + (OS2 + 2.0 * OS1) * B92;
    SC = SC * X + (OS2 + OS1) * B91;
    SC = SC * X + OS2 * B90;
    EPS2 = EPS2 + BETA * R10 * SC;
    const double H = EPS2 * 70120.4;
    rho = D;
    uout = H - P * V;
    return 0;
  }
  return 1;
}

// src/IFC67.F90:580-600
WB_HD double wb_ifc67_region2_viscosity(double temperature, double density) {
  const double v1 = 0.407 * temperature + 80.4;
  if (temperature <= 350.0) return 1.0e-7 * (v1 - density * (1858.0 - 5.9 * temperature) * 1.0e-3);
  return 1.0e-7 * (v1 + density * (0.353 + density * (676.5e-6 + density * 102.1e-9)));
}

// src/IFC67.F90:200-222
WB_HD int wb_ifc67_phase_composition(int region) {
  return region == 1 ? 1 : region == 2 ? 2 : region == 4 ? 3 : 0;
}

// ---------------------------------------------------------------- dispatch
// (the reference dispatches through class(region_type) pointers,
// src/thermodynamics.F90:46-104; here a branch on the formulation id that is
// uniform across the whole grid)

// region 1 or 2 properties at (p, t): density and internal energy
WB_HD int wb_region_properties(const WbThermo &th, int region, double p, double t, double &rho, double &u) {
  if (th.id == WB_THERMO_IAPWS) {
    if (region == 1) return wb_iapws_region1(th, p, t, rho, u);
    if (region == 2) return wb_iapws_region2(th, p, t, rho, u);
    if (region == 3) return wb_iapws_region3(th, p, t, rho, u);
  } else {
    if (region == 1) return wb_ifc67_region1(th, p, t, rho, u);
    if (region == 2) return wb_ifc67_region2(th, p, t, rho, u);
  }
  rho = 0.0;
  u = 0.0;
  return 1;
}

WB_HD double wb_region_viscosity(const WbThermo &th, int region, double temperature, double pressure,
                                 double density) {
  if (th.id == WB_THERMO_IAPWS) return wb_iapws_viscosity(th, temperature, density);
  if (region == 1) return wb_ifc67_region1_viscosity(th, temperature, pressure);
  return wb_ifc67_region2_viscosity(temperature, density);
}

WB_HD int wb_saturation_pressure(const WbThermo &th, double t, double &p) {
  return th.id == WB_THERMO_IAPWS ? wb_iapws_sat_pressure(th, t, p) : wb_ifc67_sat_pressure(th, t, p);
}
WB_HD int wb_saturation_temperature(const WbThermo &th, double p, double &t) {
  return th.id == WB_THERMO_IAPWS ? wb_iapws_sat_temperature(th, p, t) : wb_ifc67_sat_temperature(th, p, t);
}
WB_HD int wb_phase_composition(const WbThermo &th, int region, double pressure, double temperature) {
  return th.id == WB_THERMO_IAPWS ? wb_iapws_phase_composition(th, region, pressure, temperature)
                                  : wb_ifc67_phase_composition(region);
}
