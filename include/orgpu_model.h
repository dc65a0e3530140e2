/* orgpu_model.h -- plain-C model description shared by the C-ABI library (include/orgpu.h)
 * and the CPU oracle (oracle/).  INTERFACE ONLY: no algorithm lives here.
 *
 * These structs carry what the reference Engine keeps in PM(NPROPM,mat), GEO(NPROPG,pid),
 * MAT_PARAM(mat)%UPARAM/IPARAM, IPARG(NPARG,ng) and a few /COMMON/ scalars, restricted to
 * the slots the hot path reads:
 *   - LAW2 solid  : engine/source/materials/mat/mat002/m2law.F:135-169
 *   - LAW2 shell  : engine/source/materials/mat/mat002/sigeps02c.F:91-123
 *   - LAW36 shell : engine/source/materials/mat/mat036/sigeps36c.F:197-260
 *   - LAW36 solid : engine/source/materials/mat/mat036/sigeps36.F:170-207
 *   - solid prop  : engine/source/materials/mat_share/mqviscb.F:201-207, solid/solide/shvis3.F:164-168
 *   - shell prop  : engine/source/elements/sh3n/coquedk/cncoef3.F, shell/coque/ccoef3.F:124-168
 *   - groups      : engine/source/elements/forintc.F:254-300 (IPARG slots)
 * All reals are fp64 (my_real = DOUBLE PRECISION, engine/share/r8/my_real.inc).
 */
#ifndef ORGPU_MODEL_H
#define ORGPU_MODEL_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORGPU_MAXFUNC36 10   /* max yield curves (NRATE) carried for LAW36 */

/* element family (IPARG(5)=ITY and IPARG(23)=JHBE) */
enum { ORGPU_FAM_BRICK = 1,      /* ITY=1, SFORC3            */
       ORGPU_FAM_SHELL_BT = 3,   /* ITY=3, JHBE<11,  CFORC3   */
       ORGPU_FAM_SHELL_QEPH = 24,/* ITY=3, JHBE=21..29, CZFORC3 */
       ORGPU_FAM_SH3N = 7        /* ITY=7, Ish3n 1/2, C3FORC3   */ };

/* /MAT/LAW2 (PLAS_JOHNS / PLAS_ZERIL).  uparam(1:11), iparam(1:4), therm. */
typedef struct orgpu_law2 {
  double rho0;      /* mat_param%rho  = PM(1)  (reference density) */
  double young;     /* mat_param%young = PM(20) */
  double nu;        /* PM(21) */
  double shear;     /* mat_param%shear = PM(22) */
  double bulk;      /* mat_param%bulk  = PM(32) */
  double ca, cb, cn;        /* uparam(1:3)  A, B, n            */
  double epmx;              /* uparam(4)    eps_p max          */
  double sigmx;             /* uparam(5)    sigma max          */
  double cc;                /* uparam(6)    C (strain rate)    */
  double epdr;              /* uparam(7)    eps_dot_0          */
  double fisokin;           /* uparam(8)    0=isotropic 1=kinematic */
  double asrate;            /* uparam(9)    2*pi*Fcut          */
  double z3;                /* uparam(10)   m (JC) / C3 (ZA)   */
  double z4;                /* uparam(11)   C4 (ZA)            */
  double tref, tmelt, rhocp;/* mat_param%therm                 */
  double tini;              /* initial element temperature     */
  double pshift;            /* PM(88) pressure shift           */
  double a11, a12, ssp;     /* shells: PM(24), PM(25), PM(27)  */
  double gsr, a11sr, a12sr, nusr; /* QEPH: PM(12), PM(13), PM(14), PM(190) = sqrt(G, A11, A12, nu)
                                     (starter/source/materials/mat/hm_read_mat.F90:1638-1641) */
  int iform, icc, vp, israte; /* iparam(1:4) */
  int has_temp;             /* ELBUF L_TEMP>0 (adiabatic heating tracked) */
} orgpu_law2;

/* /MAT/LAW36 (PLAS_TAB).  Values are the Starter-computed UPARAM entries SIGEPS36C reads
 * (starter/source/materials/mat/mat036/hm_read_mat36.F:268-320, engine sigeps36c.F:166-200);
 * built path: VP=0, FISOKIN=0, IFAIL 0 / 1 / 2, no E(epsp) / pressure scaling. */
typedef struct orgpu_law36 {
  double rho0, young, nu, shear, bulk;
  double a11, a12, ssp;     /* PM(24), PM(25), PM(27) as CNCOEF3B reads them          */
  double gsr, a11sr, a12sr, nusr; /* PM(12), PM(13), PM(14), PM(190): sqrt(G, A11, A12, nu)
                                     (starter/source/materials/mat/hm_read_mat.F90:1638-1641) */
  double a1u, a2u;          /* UPARAM(3) = E/(1-nu^2), UPARAM(4) = nu*UPARAM(3)        */
  double g3;                /* UPARAM(2*nrate+11) = 3G                                 */
  double g2;                /* UPARAM(2*nrate+10) = 2G                  (solids, sigeps36.F:181) */
  double ssp3d;             /* UPARAM(2*nrate+13) = sqrt((K+4G/3)/rho0) (solids, sigeps36.F:184) */
  double soundsp;           /* UPARAM(2*nrate+18) shell sound speed                    */
  double nu_mnu, t_pnu, u_mnu; /* UPARAM(2*nrate+19..21): nu/(1-nu), 3/(1+nu), 1/(1-nu) */
  double epsmax;            /* UPARAM(2*nrate+7)  (INFINITY when unset)                */
  double epsr1, epsr2, epsf;/* UPARAM(2*nrate+8), (2*nrate+9), (2*nrate+15): tensile failure strains (IFAIL = 2;
                               INFINITY, 2*INFINITY, 3*INFINITY when unset, hm_read_mat36.F:246-248) */
  double fisokin;           /* UPARAM(2*nrate+14) must be 0                            */
  double asrate;            /* PM(9) = 2*pi*Fcut                                       */
  double rate[ORGPU_MAXFUNC36];   /* UPARAM(6+j)        strain rates                   */
  double yfac[ORGPU_MAXFUNC36];   /* UPARAM(6+nrate+j)  curve scale factors            */
  int    ifunc[ORGPU_MAXFUNC36];  /* 0-based curve ids into the function table         */
  int    nrate;             /* UPARAM(1)                                               */
  int    israte;            /* IPM(3)                                                  */
  int    vp, ifail, yldcheck, ismooth; /* UPARAM(2*nrate+26..29); vp = 1 (curves on the plastic strain rate): shells only,
                                           needs nrate > 1 as the Starter enforces (hm_read_mat36.F:199) */
} orgpu_law36;

/* Function table TF/NPF for LAW36 curves: curve c occupies points
 * [npf[c], npf[c+1]) of (x,y) pairs: x = tf[2*p], y = tf[2*p+1]. */

/* /PROP/SOLID (IGTYP 14) slots read on the path */
typedef struct orgpu_prop_solid {
  double qa, qb;        /* GEO(14), GEO(15) bulk viscosity            */
  double cns1, cns2;    /* GEO(16), GEO(17) Navier-Stokes viscosity   */
  double hcoef;         /* GEO(13) hourglass coefficient h            */
  double dtmin;         /* GEO(172)                                    */
  int    jhbe;          /* Isolid -> IPARG(23): 0,1,2                  */
  int    ismstr;        /* IPARG(9): 1,2,4                             */
  int    ipla;          /* IPARG(29) = IPLAST (default 1, starter sgrtails.F:510): radial-return variant of SIGEPS36 */
  int    istrain;       /* IPARG(44): MULAW accumulates LBUF%STRA (mulaw.F90:886-892)   */
  int    jcvt;          /* IPARG(37) = Iframe - 1: 0 global frame (Jaumann rate, SROTA3), 1 Belytschko co-rotational frame (SRCOOR3 / SRROTA3) */
  int    pad;
} orgpu_prop_solid;

/* /PROP/SHELL (IGTYP 1) slots read on the path (starter hm_read_prop01.F:156-262) */
typedef struct orgpu_prop_shell {
  double thick;         /* GEO(1)  THKE                                           */
  double h1, h2, h3;    /* BT: GEO(13:15) hm, hf, hr ; QEPH: h1 = GEO(13) = Dn    */
  double srh1, srh2, srh3; /* GEO(18:20)                                          */
  double shf;           /* GEO(38) shear factor (5/6; 0 when NPT=1)               */
  double shfsr;         /* GEO(100) = sqrt(GEO(38)) (starter hm_read_properties.F:796) */
  double cvis;          /* QEPH: GEO(17) = hourglass factor FAC1 (default 1)      */
  double dm;            /* GROUP_PARAM%VISC_DM membrane damping                   */
  int    npt;           /* IPARG(6)                                               */
  int    ismstr;        /* IPARG(9)                                               */
  int    ithk;          /* IPARG(28)                                              */
  int    ipla;          /* IPARG(29)                                              */
  int    ihbe;          /* Ishell: 1..4 BT (JHBE<11), 24 QEPH                     */
  int    istrain;       /* IPARG(44): accumulate GBUF%STRA                        */
} orgpu_prop_shell;

/* Engine-wide scalars (COMMON blocks / engine deck) the path reads */
/* /FAIL/JOHNSON on shells: FAIL_JOHNSON_C (engine/source/materials/fail/johnson_cook/fail_johnson_c.F:111-130, called from
 * mulawc.F90:2118-2127 after the law, per integration point) and the element deletion rule of FAIL_SETOFF_C for one layer
 * (fail/fail_setoff_c.F:123-186).  IXFEM = 0, local (no /NONLOCAL), D5 = 0 (no temperature term). */
typedef struct orgpu_fail {
  int irupt;            /* 0: none, 1: /FAIL/JOHNSON (mat_param%fail(ifl)%irupt)                              */
  int pad;
  double d1, d2, d3, d4, d5;   /* UPARAM(1:5): eps_f = (D1 + D2 exp(D3 sigma*)) (1 + D4 ln(max(1, epsp / EPSP0)))   */
  double epsp0;         /* UPARAM(6)  reference strain rate                                                       */
  double epsf_min;      /* UPARAM(12) lower bound of the failure strain                                           */
  double pthk;          /* fail%pthk: > 0 broken fraction of the thickness, < 0 ratio of broken points, 0: the property's */
  double pthickg;       /* GEO(42,PID): P_thickfail of the property                                               */
} orgpu_fail;

typedef struct orgpu_control {
  double dtfac_brick;   /* DTFAC1(1)  /DT/BRICK scale                 */
  double dtfac_shell;   /* DTFAC1(3)  /DT/SHELL scale                 */
  double dtmx;          /* DTMX (EP20 when unset)                     */
  double dt_init;       /* DT2 carried in from the restart (becomes DT1 of cycle 0) */
  double dt2old_init;   /* DT2OLD carried in from the restart         */
  double tt_init;       /* TT                                          */
  int    iroddl;        /* rotational dofs present (shells)           */
  int    nodadt;        /* NODADT: 1 = /DT/NODA nodal time step (bricks and QEPH shells; no /DT/NODA/CST) */
  double dtfac_node;    /* DTFAC1(11) /DT/NODA scale                  */
  double dtfac_sh3n;    /* DTFAC1(7)  /DT/SH_3N scale                 */
} orgpu_control;

#ifdef __cplusplus
}
#endif
#endif
