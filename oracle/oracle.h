/* oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (C++17, scalar, libm, compiled with -ffp-contract=off like the reference's
 * -ffp-contract=off / -no-fma builds) of the OpenRadioss Engine routines on the explicit
 * element cycle hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may link or call anything under oracle/.  The product path
 * (openradioss_b200/csrc) never includes this directory.
 *
 * PARITY PINNING: the reference Engine is Fortran and cannot be compiled in this image (no
 * Fortran compiler), and its own QA suite holds no routine-level vectors for this path
 * (SURVEY.md 8c).  PINNED by executing reference code: Belytschko-Tsay + LAW2 membrane /
 * transverse-shear response against the reference's own CUDA shell path built unmodified into
 * oracle/_ref (tests/test_ref_gpu_pin.py, golden vectors in tests/golden/refgpu_bt_law2_*.npz replayed
 * by tests/test_golden_refgpu.py: forces, moments, time step to 1e-12).  PARITY UNPINNED for the rest
 * (bending, QEPH, LAW36, bricks, assembly, update): there the restatement stands on line-by-line
 * correspondence to the cited Fortran and analytic patch tests (tests/test_oracle_*.py) -- see DESIGN.md.
 *
 * The oracle keeps the reference's data model: Fortran-ordered nodal arrays X(3,NUMNOD),
 * element groups of NEL<=MVSIZ elements with ELBUF-style component-major state
 * (sig[(k)*nel+i]), IXS(11,*) / IXC(7,*) connectivity, FSKY(8,LSKY) + ADSKY + IADS/IADC.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <vector>
#include <cmath>
#include <cstdint>
#include <algorithm>
#include "../include/orgpu_model.h"
#include "../include/or_constants.h"

#define MVSIZ 129   /* engine/share/spe_inc/mvsiz_p.inc:54-55 */

struct OrcSolidGroup {          /* one element group, ITY=1 (forint.F -> SFORC3) */
  int nel = 0, nft = 0;         /* nft: offset of first element in IXS/IADS (0-based) */
  int law = 2;                  /* MLW: 2 (M2LAW) or 36 (MULAW -> SIGEPS36) */
  orgpu_law2 mat;               /* MLW=2 (for MLW=36 only rho0 = PM(1) is mirrored here) */
  orgpu_law36 m36;              /* MLW=36 */
  orgpu_prop_solid prop;
  /* ELBUF (gbuf == lbuf for 1-IP solids), component-major (k*nel+i) */
  std::vector<double> sig;      /* 6*nel */
  std::vector<double> eint, rho, qvis, pla, epsd, vol, off, temp, dmg;
  std::vector<double> smstr;    /* SAV(nel,21) */
  std::vector<double> stra, wpla; /* MLW=36: LBUF%STRA(6*nel) when ISTRAIN>0, LBUF%WPLA */
  std::vector<double> sigb;     /* MLW=2, FISOKIN>0: LBUF%SIGB(6*nel), the back stress of the kinematic hardening */
  orgpu_fail fail{};            /* /FAIL/JOHNSON of the group's material (MLW=2; irupt = 0: none) */
  std::vector<double> dfmax;    /* FBUF%FLOC%DAMMX */
  std::vector<int> vartmp;      /* MLW=36: VARTMP(nel,2+NRATE) table cursors */
};

struct OrcShellGroup;           /* defined in shell files */

struct Oracle {
  int numnod = 0;
  orgpu_control ctl{};
  /* nodal arrays, Fortran (3,NUMNOD) column-major -> [3*n+c]  (nodal_arrays.F90:125-176) */
  std::vector<double> X, V, VR, D, DR, A, AR, MS, IN, STIFN, STIFR;
  std::vector<double> FEXT, MEXT;     /* constant external nodal loads (3,N) */
  std::vector<double> LA, LAR;        /* this cycle's nodal load sums (3,N): FEXT * f + the load records in record order */
  int iparit = 1;                     /* IPARIT: 1 /PARITH/ON -- FORCE leaves its records in FSKY rows BEHIND the element rows of a node
                                         (force.F90:714-1034, Starter order domdec2.F:2363-2388), ASSPAR4 adds them last;
                                         0 /PARITH/OFF -- FORCE adds to A before the element loop (force.F90:182-312) */
  std::vector<int>    ICODT, ICODR;   /* BCS codes (bit 4=x,2=y,1=z fixed), bcs10.F */
  std::vector<int> ITAB;              /* user node ids (NELTST of a nodal time step) */
  int LF_FUNC=-1; double LF_FCX=1.0;  /* time function of the concentrated loads (force.F90:195-196, 235) */
  std::vector<int> CL_IB; std::vector<double> CL_FAC;   /* ... or the load records themselves: (node, direction, function), (FCY, FCX) */
  std::vector<int> IBFV;              /* imposed velocities (3,n): node, direction, curve (fixvel.F) */
  std::vector<double> VEL;            /* (4,n): FAC, STARTT, STOPT, FACX */
  std::vector<double> FV_DW;          /* VEL(4,N) of the Engine: DT2*DW of the last cycle, booked into WFEXT at the next one */
  std::vector<int> IGRV, IBGRV;       /* gravity loads (3,n): node count, direction, curve; node lists (gravit.F) */
  std::vector<double> AGRV;           /* (2,n): FCY, FCX */
  /* connectivity */
  std::vector<int> IXS;   /* (11,NUMELS) 1-based nodes in 2..9, user id in 11 */
  std::vector<int> IXC;   /* (7,NUMELC)  1-based nodes in 2..5, user id in 7  */
  std::vector<int> IXTG;  /* (6,NUMELTG) mat, 1-based nodes in 2..4, pid, user id in 6 */
  int numels = 0, numelc = 0, numeltg = 0;
  /* /PARITH/ON tables (parith_on_mod.F90:39-74) */
  std::vector<int> ADSKY;  /* numnod+1, 1-based slot addresses */
  std::vector<int> IADS;   /* (8,NUMELS) 1-based slot of each brick corner */
  std::vector<int> IADC;   /* (4,NUMELC) */
  std::vector<int> IADTG;  /* (3,NUMELTG) */
  std::vector<double> FSKY;/* (8,LSKY) */
  int lsky = 0;
  std::vector<OrcSolidGroup> sgroups;
  std::vector<OrcShellGroup*> cgroups;
  std::vector<OrcShellGroup*> tgroups;   /* 3-node shell groups (ITY=7), processed after the 4-node shells */
  /* LAW36 function table */
  std::vector<double> TF; std::vector<int> NPF;
  /* time-step bookkeeping (resol.F:2721, 6124-6128, 6352, 6494-6497) */
  double TT = 0, DT1 = 0, DT2 = 0, DT12 = 0, DT2OLD = 0;
  double DT2T = 0; int NELTST = 0, ITYPTST = 0;
  long NCYCLE = 0;
  int nthreads = 1;
  /* balances on print cycles (bilan.cpp): IPRI, parts, PARTSAV(1:6,part), the global line of ECRIT */
  int ipri = 0, npart = 1;
  std::vector<int> IPARTC, IPARTS, IPARTTG;      /* 0-based part of each element */
  std::vector<double> GVOLC, GVOLTG;             /* GBUF%VOL of the shells */
  std::vector<double> PARTSAV;                   /* (6,npart), zeroed at the start of a print cycle */
  double ENCIN = 0, ENROT = 0, ENINT = 0, WFEXT = 0, XMOMT = 0, YMOMT = 0, ZMOMT = 0, XMASS = 0;
};

/* solid.cpp */
void orc_sforc3(Oracle& o, OrcSolidGroup& g, double& dt2t, int& neltst, int& ityptst);
/* shell_qeph.cpp / shell_bt.cpp */
void orc_czforc3(Oracle& o, OrcShellGroup& g, double& dt2t, int& neltst, int& ityptst);
void orc_cforc3(Oracle& o, OrcShellGroup& g, double& dt2t, int& neltst, int& ityptst);
/* shell_c3.cpp */
void orc_c3forc3(Oracle& o, OrcShellGroup& g, double& dt2t, int& neltst, int& ityptst);
/* bilan.cpp */
void orc_bilan_shell(Oracle& o,int elem,int nn,const int* nodes,double eint1,double eint2,double rho,double off);
void orc_bilan_solid(Oracle& o,int elem,const double vx[8],const double vy[8],const double vz[8],
                     double eint,double vol,double rho,double vnew,double off);
void orc_ecrit(Oracle& o);
/* assembly.cpp */
void orc_asspar4(Oracle& o);
void orc_accele(Oracle& o);
void orc_gravit(Oracle& o);
void orc_bcs(Oracle& o);
void orc_fixvel(Oracle& o);
void orc_dtnoda(Oracle& o);
void orc_velocity(Oracle& o);
void orc_depla(Oracle& o);
void orc_cycle(Oracle& o);

#endif
