/* oracle/shell.h -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Element-buffer mirror of one 4-node shell group (ITY=3): G_BUFEL_ / L_BUFEL_ fields the path
 * touches (common_source/modules/mat_elem/elbufdef_mod.F90:739-1013, 1184-1300), component-major
 * (k*nel+i) exactly as the reference stores them. */
#ifndef ORACLE_SHELL_H
#define ORACLE_SHELL_H
#include "oracle.h"
#include "../include/or_quadrature.h"

struct OrcShellGroup {
  int nel = 0, nft = 0;
  int law = 0;                     /* MLW: 2 or 36 */
  orgpu_prop_shell prop{};
  orgpu_law2 m2{};
  orgpu_law36 m36{};
  /* GBUF */
  std::vector<double> FOR, MOM, EINT, THK, OFF, STRA, EPSD, HOURG, SMSTR, THKE;
  /* LBUF per integration point */
  struct Lbuf { std::vector<double> sig, pla, epsd, temp, sigb, off, plap; std::vector<int> vartmp;
                std::vector<double> dfmax, foff; };    /* /FAIL/JOHNSON: FBUF%FLOC%DAMMX, FBUF%FLOC%OFF of the point */
  orgpu_fail fail{};               /* failure model of the group's material (irupt = 0: none) */
  std::vector<Lbuf> ip;
  int nvartmp = 0;
  int nhourg = 0;                  /* 12 QEPH, 5 BT */
};

/* shared by shell_qeph.cpp and shell_bt.cpp */
struct ShellMatIn {                /* what CMAIN3/MULAWC take from the element routine */
  double exx, eyy, exy, exz, eyz, kxx, kyy, kxy;
  double area, thk0, off, nu, g, a11, a12, gs, shf, rho, ssp, epsd_pg, dt1c;
};
struct ShellMatOut { double ssp, viscmx, sigy, zcfac1, zcfac2, ssp_eq, vol0; };

void orc_cmain3(const Oracle& o, OrcShellGroup& g, int i, bool flag_zcfac, ShellMatIn& in, ShellMatOut& out);
void orc_vinter(const std::vector<double>& TF, int iad /*0-based pair index of first point*/, int npts,
                int& ipos, double x, double& dydx, double& y);
#endif
