/* oracle/assembly.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Restates: ASSPAR4 (engine/source/assembly/asspar4.F:164-181), ACCELE (accele.F:65-131),
 * BCS10-style fixed dof masks (constraints/general/bcs/bcs10.F), VELOCITY (velocity.F:57-89),
 * DEPLA (displacement.F:91-113) and the RESOL time-step bookkeeping
 * (engine/source/engine/resol.F:2721-2722, 4165-4171, 6124-6128, 6352, 6494-6497, 8599-8608).
 */
#include "oracle.h"
#include <omp.h>

struct OrcShellGroup;
void orc_shell_dispatch(Oracle& o, OrcShellGroup& g, double& dt2t, int& neltst, int& ityptst);

/* ASSPAR4: left fold of slots ADSKY(N)..ADSKY(N+1)-1 into the EXISTING A/AR/STIFN/STIFR */
void orc_asspar4(Oracle& o)
{
  const int n=o.numnod;
  #pragma omp parallel for schedule(static) num_threads(o.nthreads)
  for(int N=0;N<n;N++){
    int nct=o.ADSKY[N]-1; int nc=o.ADSKY[N+1]-o.ADSKY[N];
    for(int k=nct;k<nct+nc;k++){
      const double* f=&o.FSKY[8*(size_t)k];
      o.A[3*N]  =o.A[3*N]  +f[0];
      o.A[3*N+1]=o.A[3*N+1]+f[1];
      o.A[3*N+2]=o.A[3*N+2]+f[2];
      o.AR[3*N]  =o.AR[3*N]  +f[3];
      o.AR[3*N+1]=o.AR[3*N+1]+f[4];
      o.AR[3*N+2]=o.AR[3*N+2]+f[5];
      o.STIFN[N]=o.STIFN[N]+f[6];
      o.STIFR[N]=o.STIFR[N]+f[7];
    }
    /* /PARITH/ON: the load records are the LAST rows of a node's skyline (pseudo-elements appended after all elements,
     * starter/source/spmd/domdec2.F:2363-2388; FORCE fills them, force.F90:714-1034) */
    if(o.iparit!=0 && !o.LA.empty()){
      for(int c=0;c<3;c++){ o.A[3*N+c]=o.A[3*N+c]+o.LA[3*N+c]; o.AR[3*N+c]=o.AR[3*N+c]+o.LAR[3*N+c]; }
    }
  }
}

/* ACCELE accele.F:65-131 (N2D=0, NMULT=0) */
void orc_accele(Oracle& o)
{
  const int n=o.numnod;
  #pragma omp parallel for schedule(static) num_threads(o.nthreads)
  for(int N=0;N<n;N++){
    if(o.MS[N]>K_ZERO){
      double rtmp=K_ONE/o.MS[N];
      o.A[3*N]*=rtmp; o.A[3*N+1]*=rtmp; o.A[3*N+2]*=rtmp;
    } else { o.A[3*N]=K_ZERO; o.A[3*N+1]=K_ZERO; o.A[3*N+2]=K_ZERO; }
    if(o.ctl.iroddl!=0){
      if(o.IN[N]>K_ZERO){
        double rtmp=K_ONE/o.IN[N];
        o.AR[3*N]*=rtmp; o.AR[3*N+1]*=rtmp; o.AR[3*N+2]*=rtmp;
      } else { o.AR[3*N]=K_ZERO; o.AR[3*N+1]=K_ZERO; o.AR[3*N+2]=K_ZERO; }
    }
  }
}

static double orc_finter(const Oracle& o,int f,double XX);
/* GRAVIT  engine/source/loads/general/grav/gravit.F:84-160 (ISK<=1 global frame, N2D=0, no sensor: TS=TT, ISMOOTH=0):
 * A0/GAMA :103-119, A(N2,N1)=A(N2,N1)+AA :149-153.  Called after ACCELE and before BCS10 (resol.F:6921, 7123, 7322), so AA is an
 * acceleration.  The external work WFEXT (:152) is energy bookkeeping outside the path. */
void orc_gravit(Oracle& o)
{
  const int ngrav=(int)(o.IGRV.size()/3);
  int IAD=0;
  for(int NL=0;NL<ngrav;NL++){
    const double FCY=o.AGRV[2*NL], FCX=o.AGRV[2*NL+1];
    const int NN=o.IGRV[3*NL], N2=o.IGRV[3*NL+1], IFUNC=o.IGRV[3*NL+2];
    const double TS=o.TT;
    const double GAMA = IFUNC>=0 ? FCY*orc_finter(o,IFUNC,TS*FCX) : FCY;
    const double AA=GAMA;
    for(int J=IAD;J<IAD+NN;J++){ const int N1=std::abs(o.IBGRV[J]); o.A[3*(N1-1)+(N2-1)]=o.A[3*(N1-1)+(N2-1)]+AA; }
    IAD+=NN;
  }
}

/* BCS10 (global-frame codes only): zero the acceleration of fixed dofs.
 * code bits as in ICODT/ICODR: 4 -> x, 2 -> y, 1 -> z   (bcs10.F, skew 0 branch) */
void orc_bcs(Oracle& o)
{
  const int n=o.numnod;
  if(o.ICODT.empty()) return;
  #pragma omp parallel for schedule(static) num_threads(o.nthreads)
  for(int N=0;N<n;N++){
    int c=o.ICODT[N];
    if(c&4) o.A[3*N]=K_ZERO; if(c&2) o.A[3*N+1]=K_ZERO; if(c&1) o.A[3*N+2]=K_ZERO;
    if(o.ctl.iroddl!=0){
      int r=o.ICODR[N];
      if(r&4) o.AR[3*N]=K_ZERO; if(r&2) o.AR[3*N+1]=K_ZERO; if(r&1) o.AR[3*N+2]=K_ZERO;
    }
  }
}

/* FINTER (engine/source/tools/curve/finter.F:165-356): fewer than 20 segments -> the classical walk (:210-229); otherwise the
 * two end segments first (:236-289), a dichotomy down to fewer than 20 segments (:294-330), then the walk over what is left
 * (:343-356).  TF holds (x,y) pairs; curve f = points NPF[f] .. NPF[f+1]-1 (0-based) */
static double orc_finter_seg(const double* TF,int i0,int I,double DX1,double DX2)
{
  double DIV0=TF[2*(i0+I)]-TF[2*(i0+I-1)];
  double DIV=std::max(std::fabs(DIV0),K_EM16);
  DIV=std::copysign(DIV,DIV0);
  double DERI=(TF[2*(i0+I)+1]-TF[2*(i0+I-1)+1])/DIV;
  if(DX1<=DX2) return TF[2*(i0+I-1)+1]+DX1*DERI;
  return TF[2*(i0+I)+1]-DX2*DERI;
}
static double orc_finter(const Oracle& o,int f,double XX)
{
  const int i0=o.NPF[f], n=o.NPF[f+1]-o.NPF[f];
  const double* TF=o.TF.data();
  if(n==1) return TF[2*i0+1];
  const int POINT_NBR=n-1, MIN_GAP=20;
  double DX2=TF[2*i0]-XX;
  if(POINT_NBR<MIN_GAP){
    for(int I=1;I<n;I++){
      double DX1=-DX2;
      DX2=TF[2*(i0+I)]-XX;
      if(DX2>=K_ZERO || I==n-1) return orc_finter_seg(TF,i0,I,DX1,DX2);
    }
    return K_ZERO;
  }
  { double DX1=-DX2; DX2=TF[2*(i0+1)]-XX;                               /* first shot (a): the first segment */
    if(DX2>=K_ZERO) return orc_finter_seg(TF,i0,1,DX1,DX2); }
  { DX2=TF[2*(i0+n-1)]-XX; double DX1=-DX2;                              /* first shot (b): beyond the last point */
    if(DX2<=K_ZERO){ if(DX1==K_ZERO && DX2==K_ZERO) return TF[2*(i0+n-1)+1]; return orc_finter_seg(TF,i0,n-1,DX1,DX2); } }
  int FIRST=1, LAST=POINT_NBR, COUNTER=0; bool BOOL=true;
  while(BOOL){
    const int MIDDLE=(LAST-FIRST)/2+FIRST;
    const double DX2_FIRST=TF[2*(i0+FIRST)]-XX, DX2_LAST=TF[2*(i0+LAST)]-XX, DX2_MIDDLE=TF[2*(i0+MIDDLE)]-XX;
    const double PRODUCT_FM=DX2_FIRST*DX2_MIDDLE, PRODUCT_ML=DX2_MIDDLE*DX2_LAST;
    if(PRODUCT_FM<0) LAST=MIDDLE; else if(PRODUCT_ML<0) FIRST=MIDDLE; else BOOL=false;
    if(LAST-FIRST<MIN_GAP) BOOL=false;
    COUNTER=COUNTER+1;
    if(COUNTER>POINT_NBR){ COUNTER=-1; BOOL=false; }
  }
  if(COUNTER==-1){ FIRST=1; LAST=POINT_NBR; }
  DX2=TF[2*(i0+FIRST-1)]-XX;
  for(int J=FIRST;J<=LAST;J++){
    double DX1=-DX2;
    DX2=TF[2*(i0+J)]-XX;
    if(DX2>=K_ZERO || J==LAST) return orc_finter_seg(TF,i0,J,DX1,DX2);
  }
  return K_ZERO;
}
double orc_load_scale(const Oracle& o){ return o.LF_FUNC>=0 ? orc_finter(o,o.LF_FUNC,o.TT*o.LF_FCX) : K_ONE; }

/* FIXVEL (engine/source/constraints/general/impvel/fixvel.F), restricted to imposed velocities
 * (IBFV(7,N)=1) in the global frame without sensor: TSC=(TT+HALF*DT2)*FACX (:146), YC from VINTERDP
 * (vinterdp.F:35-70, cursor from 0), YC=YC*FAC (:344), A(J,I)=(YC-V(J,I))/DT12 (:375-377).  Called after
 * BCS (resol.F:7322) and before VELOCITY (resol.F:8947), with TT = time at the start of the cycle. */
void orc_fixvel(Oracle& o)
{
  const int nfx=(int)(o.IBFV.size()/3);
  for(int N=0;N<nfx;N++){
    const double FAC=o.VEL[4*N], STARTT=o.VEL[4*N+1], STOPT=o.VEL[4*N+2], FACX=o.VEL[4*N+3];
    if(o.TT<STARTT) continue;
    if(o.TT>STOPT) continue;
    const int I=o.IBFV[3*N]-1, J=o.IBFV[3*N+1]-1, L=o.IBFV[3*N+2];
    if(o.FV_DW.size()!=(size_t)nfx) o.FV_DW.assign(nfx,K_ZERO);
    o.WFEXT=o.WFEXT+o.FV_DW[N]; o.FV_DW[N]=K_ZERO;      /* fixvel.F:342-344 (RW_SMS = 1): the DT2 half booked last cycle */
    const double TSC=(o.TT+K_HALF*o.DT2)*FACX;
    const int IAD=o.NPF[L], NP=o.NPF[L+1]-o.NPF[L];
    const double* TF=o.TF.data();
    int IPOS=0;
    for(int JJ=1;JJ<=NP-2;JJ++){ if(TSC>TF[2*(IAD+IPOS+1)]) IPOS++; else break; }
    const double TF1J1=TF[2*(IAD+IPOS)], TF2J1=TF[2*(IAD+IPOS)+1], TF1J2=TF[2*(IAD+IPOS+1)], TF2J2=TF[2*(IAD+IPOS+1)+1];
    const double DYDX=(TF2J2-TF2J1)/(TF1J2-TF1J1);
    double YC=TF2J1+DYDX*(TSC-TF1J1);
    YC=YC*FAC;
    YC=(YC-o.V[3*I+J])/o.DT12;
    const double AOLD=o.A[3*I+J];
    o.A[3*I+J]=YC;
    /* work of the imposed velocity (fixvel.F:391-394, 834-837): DW = 1/4 MS (A DT12 + 2 V)(A - AOLD); WFEXT gets DT1*DW now and DT2*DW at the next cycle */
    const double DW=K_FOURTH*o.MS[I]*(o.A[3*I+J]*o.DT12+K_TWO*o.V[3*I+J])*(o.A[3*I+J]-AOLD);
    o.WFEXT=o.WFEXT+o.DT1*DW;
    o.FV_DW[N]=o.FV_DW[N]+o.DT2*DW;                     /* VEL(4,N) = VEL(4,N) + DT2*DW (:837) */
  }
}

/* VELOCITY velocity.F:57-89 */
void orc_velocity(Oracle& o)
{
  const int n=o.numnod; const double DT12=o.DT12;
  #pragma omp parallel for schedule(static) num_threads(o.nthreads)
  for(int N=0;N<n;N++){
    for(int c=0;c<3;c++){ o.V[3*N+c]=o.V[3*N+c]+DT12*o.A[3*N+c]; o.A[3*N+c]=K_ZERO; }
    if(o.ctl.iroddl!=0) for(int c=0;c<3;c++){ o.VR[3*N+c]=o.VR[3*N+c]+DT12*o.AR[3*N+c]; o.AR[3*N+c]=K_ZERO; }
  }
}

/* DEPLA displacement.F:91-103 (IRESP=0; DR not advanced: ISECUT=IISROT=IMPOSE_DR=IDROT=0) */
void orc_depla(Oracle& o)
{
  const int n=o.numnod; const double DT2=o.DT2;
  #pragma omp parallel for schedule(static) num_threads(o.nthreads)
  for(int N=0;N<n;N++){
    for(int c=0;c<3;c++){
      double VDT=DT2*o.V[3*N+c];
      o.D[3*N+c]=o.D[3*N+c]+VDT;
      o.X[3*N+c]=o.X[3*N+c]+VDT;
    }
  }
}

/* element force phase: this cycle's nodal loads (into A / AR at once with /PARITH/OFF, kept for ASSPAR4 with /PARITH/ON), then all groups write FSKY */
void orc_forces(Oracle& o)
{
  const int n=o.numnod;
  /* resol.F: A/AR hold external nodal loads when the element loop starts (FORCE, resol.F:2929);
   * STIFN/STIFR restart from zero each cycle */
  const double fs=orc_load_scale(o);   /* force.F90:235, 301-312: AA = FCY*FINTER(IFUN,TT*FCX) */
  const bool loaded = !o.FEXT.empty() || !o.MEXT.empty() || !o.CL_IB.empty();
  if(loaded){ o.LA.assign((size_t)3*n,K_ZERO); o.LAR.assign((size_t)3*n,K_ZERO); } else { o.LA.clear(); o.LAR.clear(); }
  if(loaded) for(int i=0;i<3*n;i++){ o.LA[i]= o.FEXT.empty()? K_ZERO : o.FEXT[i]*fs; o.LAR[i]= o.MEXT.empty()? K_ZERO : o.MEXT[i]*fs; }
  /* FORCE record by record (force.F90:188-312, IFUN = 1, no sensor, global frame): AA = FCY*FINTER(N3,TS*FCX); the records
   * of one node and direction are summed in record order */
  for(size_t NL=0; NL<o.CL_IB.size()/3; NL++){
    const int N1=o.CL_IB[3*NL]-1, N2=o.CL_IB[3*NL+1], N3=o.CL_IB[3*NL+2];
    const double FCY=o.CL_FAC[2*NL], FCX=o.CL_FAC[2*NL+1];
    const double AA = N3>=0 ? FCY*orc_finter(o,N3,o.TT*FCX) : FCY;
    if(N2<=3) o.LA[3*N1+N2-1]=o.LA[3*N1+N2-1]+AA;
    else      o.LAR[3*N1+N2-4]=o.LAR[3*N1+N2-4]+AA;
  }
  /* /PARITH/OFF: A += AA before the element loop (force.F90:182-312); /PARITH/ON: the records wait in their FSKY rows, which
   * ASSPAR4 reaches after the element rows (orc_asspar4) */
  for(int i=0;i<3*n;i++){ o.A[i]= (loaded && o.iparit==0)? o.LA[i] : K_ZERO; o.AR[i]= (loaded && o.iparit==0)? o.LAR[i] : K_ZERO; }
  /* with /DT/NODA the nodal stiffnesses restart from EM20 (dtnoda.F:336-338, rotational part alike) */
  const double st0 = o.ctl.nodadt!=0 ? K_EM20 : K_ZERO;
  for(int i=0;i<n;i++){ o.STIFN[i]=st0; o.STIFR[i]=st0; }
  double DT2T=o.DT2; int NELTST=0, ITYPTST=0;   /* thread mins are merged with strict "<" (resol.F:4165-4171) */
  /* shells first (FORINTC resol.F:4138), then solids (FORINT resol.F:4225) */
  const int ncg=(int)o.cgroups.size(), nsg=(int)o.sgroups.size(), ntg=(int)o.tgroups.size();
  if(o.nthreads<=1){
    for(int g=0;g<ncg;g++) orc_shell_dispatch(o,*o.cgroups[g],DT2T,NELTST,ITYPTST);
    for(int g=0;g<ntg;g++) orc_c3forc3(o,*o.tgroups[g],DT2T,NELTST,ITYPTST);      /* ITY=7 groups follow ITY=3 in the group list */
    for(int g=0;g<nsg;g++) orc_sforc3(o,o.sgroups[g],DT2T,NELTST,ITYPTST);
  } else {
    /* OpenMP over groups as forintc.F:238 (!$OMP DO SCHEDULE(DYNAMIC,1)); thread-private DT2TT */
    #pragma omp parallel num_threads(o.nthreads)
    {
      double dt2tt=o.DT2; int nelt=0, ityp=0;
      #pragma omp for schedule(dynamic,1) nowait
      for(int g=0;g<ncg;g++) orc_shell_dispatch(o,*o.cgroups[g],dt2tt,nelt,ityp);
      #pragma omp for schedule(dynamic,1) nowait
      for(int g=0;g<ntg;g++) orc_c3forc3(o,*o.tgroups[g],dt2tt,nelt,ityp);
      #pragma omp for schedule(dynamic,1)
      for(int g=0;g<nsg;g++) orc_sforc3(o,o.sgroups[g],dt2tt,nelt,ityp);
      #pragma omp critical
      { if(dt2tt<DT2T){ DT2T=dt2tt; NELTST=nelt; ITYPTST=ityp; } }
    }
  }
  o.DT2T=DT2T; o.NELTST=NELTST; o.ITYPTST=ITYPTST;
}

/* DTNODA (engine/source/time_step/dtnoda.F:221-260, 324-334 translations; :445-462 + fold rotations), NODADT>0,
 * IDTMIN(11)=0, Lagrangian: DTN = DTFAC1(11)*SQRT(TWO*MS/STIFN) over nodes with MS>0 (STIFN>0), strict "<" in node
 * order, NELTST=ITAB(N), ITYPTST=11; then the same with IN / STIFR when IRODDL/=0.  Called after ASSPAR4 (resol.F:6066). */
void orc_dtnoda(Oracle& o)
{
  const int n=o.numnod; const double fac=o.ctl.dtfac_node;
  for(int N=0;N<n;N++){
    if(o.STIFN[N]<=K_ZERO) continue;
    if(!(o.MS[N]>K_ZERO)) continue;
    const double DTN=fac*std::sqrt(K_TWO*o.MS[N]/o.STIFN[N]);
    if(DTN<o.DT2T){ o.DT2T=DTN; o.NELTST=o.ITAB.empty()? N+1 : o.ITAB[N]; o.ITYPTST=11; }
  }
  if(o.ctl.iroddl!=0){
    for(int N=0;N<n;N++){
      if(o.STIFR[N]<=K_ZERO) continue;
      if(!(o.IN[N]>K_ZERO)) continue;
      const double DTN=fac*std::sqrt(K_TWO*o.IN[N]/o.STIFR[N]);
      if(DTN<o.DT2T){ o.DT2T=DTN; o.NELTST=o.ITAB.empty()? N+1 : o.ITAB[N]; o.ITYPTST=11; }
    }
  }
}

/* one pass of RESOL restricted to the hot path */
void orc_cycle(Oracle& o)
{
  o.DT1=o.DT2;                    /* resol.F:2721 */
  o.DT2=K_EP06;                   /* resol.F:2722 */
  if(o.ipri) std::fill(o.PARTSAV.begin(),o.PARTSAV.end(),K_ZERO);
  orc_forces(o);
  orc_asspar4(o);
  if(o.ctl.nodadt!=0) orc_dtnoda(o);
  if(o.DT2T<o.DT2) o.DT2=o.DT2T;  /* resol.F:6124-6128 */
  {                               /* resol.F:6352: DT2=MIN(DT2,1.1*DT2OLD,DTMX) -- 1.1 is a REAL*4 literal */
    double c11=(double)1.1f;
    o.DT2=std::min(o.DT2,std::min(c11*o.DT2OLD,o.ctl.dtmx));
  }
  o.DT2OLD=o.DT2;                 /* resol.F:6494 */
  o.DT12=K_HALF*(o.DT1+o.DT2);    /* resol.F:6496 */
  orc_accele(o);
  orc_gravit(o);                  /* resol.F:7123 */
  orc_bcs(o);
  orc_fixvel(o);
  if(o.ipri) orc_ecrit(o);        /* SORTIE_MAIN -> ECRIT, resol.F:8523 */
  orc_velocity(o);
  orc_depla(o);
  o.TT=o.TT+o.DT2; o.NCYCLE++;    /* resol.F:8599-8608 */
}
