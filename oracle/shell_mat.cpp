/* oracle/shell_mat.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Through-thickness material loop of 4-node shells, restated per element from:
 *   CMAIN3    engine/source/materials/mat_share/cmain3.F:204-352      (NPT>0 Radioss-law branch)
 *   LAYINI    engine/source/elements/shell/coque/layini.F:246-254     (IGTYP 1: THKLY=WF, POSLY=Z0)
 *   MULAWC    engine/source/materials/mat_share/mulawc.F90:542-604 (pre), 718-1114 (IP frame),
 *             2630-2662 (stress store, FOR/MOM), 2818-2845 (ZCFAC, SSP_EQ), 2934-3091 (tail)
 *   SIGEPS36C engine/source/materials/mat/mat036/sigeps36c.F:171-661  (VP=0; IPLAS 0/1/2)
 *   VINTER    engine/source/tools/curve/vinter.F:100-130
 *   SIGEPS02C engine/source/materials/mat/mat002/sigeps02c.F:91-230
 *   M2CPLR    engine/source/materials/mat/mat002/m2cplr.F:108-507     (FISOKIN=0)
 * Built path: IGTYP=1 (/PROP/SHELL), no failure model, no thermal coupling (JTHE=0), no
 * non-local, NPG=1.  Expressions keep the Fortran evaluation order (left to right).
 */
#include "shell.h"
/* through-thickness tables (coqini.F) with a test-only override of one row: the reference's own CUDA kernels use the
 * mid-point rule (shell_strain_material_kernel.cu:696-701); loading it here lets them pin the bending path */
static double g_quad[3][121]; static bool g_quad_init=false;
const double* orc_quad_tab(int which){
  if(!g_quad_init){ for(int i=0;i<121;i++){ g_quad[0][i]=OR_Z0[i]; g_quad[1][i]=OR_WF[i]; g_quad[2][i]=OR_WM[i]; } g_quad_init=true; }
  return g_quad[which];
}
extern "C" void orc_set_quadrature(void*,int npt,const double* z0,const double* wf,const double* wm){
  orc_quad_tab(0);
  if(!z0){ g_quad_init=false; return; }                       /* NULL: back to the coqini.F tables */
  for(int i=0;i<npt;i++){ g_quad[0][(npt-1)*11+i]=z0[i]; g_quad[1][(npt-1)*11+i]=wf[i]; g_quad[2][(npt-1)*11+i]=wm[i]; }
}


/* VINTER: monotone forward walk of the persistent cursor, then linear interpolation.
 * Curve points are (x,y) pairs TF[2*p], TF[2*p+1], p in [iad, iad+npts). */
void orc_vinter(const std::vector<double>& TF, int iad, int npts, int& ipos, double x, double& dydx, double& y)
{
  /* ILEN = NPF(f+1)/2 - IAD - IPOS = npts-1-ipos ; at most ILEN-1 advances (vinter.F:104-113) */
  const int ilen=npts-1-ipos;
  for(int j=1;j<=ilen-1;j++){
    int j1=ipos+iad+1;
    if(x>TF[2*(size_t)j1]) ipos=ipos+1; else break;
  }
  const int j1=ipos+iad, j2=j1+1;
  dydx=(TF[2*(size_t)j2+1]-TF[2*(size_t)j1+1])/(TF[2*(size_t)j2]-TF[2*(size_t)j1]);
  y=TF[2*(size_t)j1+1]+dydx*(x-TF[2*(size_t)j1]);
}

namespace {

struct IpIO {             /* one integration point, one element */
  double depsxx,depsyy,depsxy,depsyz,depszx, epspxx,epspyy,epspxy;
  double sigoxx,sigoyy,sigoxy,sigoyz,sigozx;
  double signxx,signyy,signxy,signyz,signzx;
  double thklyl;
  double epsxx,epsyy,epsxy;   /* total strains at the point (mulawc.F90:856-862), IFAIL = 2 */
};

/* ---- SIGEPS36C, VP=0 branch, one element ------------------------------------------------ */
void sigeps36c(const Oracle& o, const orgpu_law36& m, int ipla, double asrate, const ShellMatIn& in, IpIO& s,
               double& pla, double& epsd, int* vartmp, double& off, double& thk, double& ssp, double& viscmax,
               double& etse, double& yld_out, double* sigb /*SIGBXX, SIGBYY, SIGBXY of the point*/,
               double& plap /*UVAR(2): filtered plastic strain rate of the point (VP = 1)*/, double dt1)
{
  const int NITER=3;
  /* VP = 1 (sigeps36c.F:665-923, 976-982): the curves are interpolated on the PLASTIC strain rate UVAR(2) instead of the total
   * one, the return is always the three Newton steps of Iplas = 1, LBUF%EPSD is left alone; the Starter keeps VP = 0 for a
   * single curve (hm_read_mat36.F:199) */
  const bool vp1=(m.vp==1);
  if(vp1) ipla=1;
  const int nrate=m.nrate;
  const double E=m.young, A1=m.a1u, A2=m.a2u, G=m.shear, G3=m.g3;
  const double NU_MNU=m.nu_mnu, T_PNU=m.t_pnu, U_MNU=m.u_mnu, FISOKIN=m.fisokin;
  const double GS=in.gs;
  viscmax=K_ZERO; etse=K_ONE; ssp=m.soundsp;
  const double PFAC=K_ONE, FACYLDI=K_ONE;
  /* damage factor on the largest in-plane principal strain (sigeps36c.F:256-264) */
  double FAIL=K_ONE, EPST=K_ZERO;
  if(m.ifail==2){
    EPST=K_HALF*(s.epsxx+s.epsyy+std::sqrt((s.epsxx-s.epsyy)*(s.epsxx-s.epsyy)+s.epsxy*s.epsxy));
    FAIL=std::max(K_EM20,std::min(K_ONE,(m.epsr2-EPST)/(m.epsr2-m.epsr1)));
  }
  /* elastic predictor (sigeps36c.F:272-284) from the stress shifted by the back stress (zero unless FISOKIN > 0) */
  s.sigoxx=s.sigoxx-sigb[0]; s.sigoyy=s.sigoyy-sigb[1]; s.sigoxy=s.sigoxy-sigb[2];
  double DPLA_I=K_ZERO;                                   /* plastic strain increment of the point: drives the back stress */
  s.signxx=s.sigoxx+A1*s.depsxx+A2*s.depsyy;
  s.signyy=s.sigoyy+A2*s.depsxx+A1*s.depsyy;
  s.signxy=s.sigoxy+G*s.depsxy;
  s.signyz=s.sigoyz+GS*s.depsyz;
  s.signzx=s.sigozx+GS*s.depszx;
  /* strain rate (:288-296) */
  if(vp1){ /* :665-: no total strain rate */ }
  else if(m.israte==0){
    epsd=K_HALF*( std::fabs(s.epspxx+s.epspyy)
         + std::sqrt( (s.epspxx-s.epspyy)*(s.epspxx-s.epspyy) + s.epspxy*s.epspxy ) );
  } else {
    epsd=asrate*in.epsd_pg+(K_ONE-asrate)*epsd;
  }
  /* yield (:317-462) */
  double YLD,H;
  if(nrate==1){
    int ipos=vartmp[2];
    const int f=m.ifunc[0];
    double dydx1,y1;
    orc_vinter(o.TF,o.NPF[f],o.NPF[f+1]-o.NPF[f],ipos,pla,dydx1,y1);
    const double YFAC1=m.yfac[0]*FACYLDI;
    vartmp[2]=ipos;
    const double FACT=FAIL*PFAC*YFAC1;
    H=dydx1*FACT;
    if(FISOKIN==K_ZERO) YLD=y1*FACT;                                   /* :329-337 */
    else if(FISOKIN==K_ONE){ const double YLD0=o.TF[2*(size_t)o.NPF[f]+1]; YLD=YLD0*FACT; }
    else { const double YLD0=o.TF[2*(size_t)o.NPF[f]+1]; YLD=((K_ONE-FISOKIN)*y1+FISOKIN*YLD0)*FACT; }
  } else {
    const double rate_x = vp1 ? plap : epsd;                            /* :705 PLAP = UVAR(2) */
    int JJ=1;
    for(int J=2;J<=nrate-1;J++) if(rate_x>=m.rate[J-1]) JJ=J;
    double RFAC,YFAC1,YFAC2;
    if(m.ismooth==2){
      double EPSP1=std::max(m.rate[JJ-1],K_EM20), EPSP2=m.rate[JJ];
      RFAC=std::log(std::max(rate_x,K_EM20)/EPSP1)/std::log(EPSP2/EPSP1);
    } else {
      double EPSP1=m.rate[JJ-1], EPSP2=m.rate[JJ];
      RFAC=(rate_x-EPSP1)/(EPSP2-EPSP1);
    }
    YFAC1=m.yfac[JJ-1]*FACYLDI; YFAC2=m.yfac[JJ]*FACYLDI;
    const int J1=JJ,J2=JJ+1;
    const int f1=m.ifunc[J1-1], f2=m.ifunc[J2-1];
    int ipos1=vartmp[1+J1], ipos2=vartmp[1+J2];
    double dydx1,y1,dydx2,y2;
    orc_vinter(o.TF,o.NPF[f1],o.NPF[f1+1]-o.NPF[f1],ipos1,pla,dydx1,y1);
    orc_vinter(o.TF,o.NPF[f2],o.NPF[f2+1]-o.NPF[f2],ipos2,pla,dydx2,y2);
    const double FAC=RFAC;
    if(FISOKIN==K_ZERO){                                               /* :383-404 */
      y1=y1*YFAC1; y2=y2*YFAC2;
      YLD=FAIL*(y1+FAC*(y2-y1));
      YLD=std::max(YLD,K_EM20);
      dydx1=dydx1*YFAC1; dydx2=dydx2*YFAC2;
      H=FAIL*(dydx1+FAC*(dydx2-dydx1));
      YLD=YLD*std::max(K_ZERO,PFAC);
      H=H*std::max(K_ZERO,PFAC);
    } else if(FISOKIN==K_ONE){                                         /* :405-429: the yield stress stays the curves' first point */
      dydx1=dydx1*YFAC1; dydx2=dydx2*YFAC2;
      H=FAIL*(dydx1+FAC*(dydx2-dydx1));
      y1=o.TF[2*(size_t)o.NPF[f1]+1]; y2=o.TF[2*(size_t)o.NPF[f2]+1];
      y1=y1*YFAC1; y2=y2*YFAC2;
      YLD=FAIL*(y1+FAC*(y2-y1));
      YLD=YLD*std::max(K_ZERO,PFAC);
      H=H*std::max(K_ZERO,PFAC);
    } else {                                                           /* :430-460 mixed hardening */
      y1=y1*YFAC1; y2=y2*YFAC2;
      YLD=FAIL*(y1+FAC*(y2-y1));
      YLD=std::max(YLD,K_EM20);
      dydx1=dydx1*YFAC1; dydx2=dydx2*YFAC2;
      H=FAIL*(dydx1+FAC*(dydx2-dydx1));
      y1=o.TF[2*(size_t)o.NPF[f1]+1]; y2=o.TF[2*(size_t)o.NPF[f2]+1];
      y1=y1*YFAC1; y2=y2*YFAC2;
      YLD=(K_ONE-FISOKIN)*YLD+FISOKIN*(FAIL*(y1+FAC*(y2-y1)));
      YLD=YLD*std::max(K_ZERO,PFAC);
      H=H*std::max(K_ZERO,PFAC);
    }
    vartmp[1+J1]=ipos1; vartmp[1+J2]=ipos2;
  }
  if(m.yldcheck==1) YLD=std::max(YLD,K_EM20);
  /* projection (:472-661) */
  if(ipla==0){
    const double NU3=K_ONE-NU_MNU;
    double SVM2=s.signxx*s.signxx+s.signyy*s.signyy-s.signxx*s.signyy+K_THREE*s.signxy*s.signxy;
    if(SVM2>YLD*YLD){
      double SVM=std::sqrt(SVM2);
      double R=YLD/SVM;
      s.signxx=s.signxx*R; s.signyy=s.signyy*R; s.signxy=s.signxy*R;
      double DPLA=off*SVM*(K_ONE-R)/(G3+H);
      DPLA_I=DPLA;
      pla=pla+DPLA;
      double DEZZ;
      if(YLD!=0) DEZZ=DPLA*K_HALF*(s.signxx+s.signyy)/YLD; else DEZZ=K_ZERO;
      DEZZ=-(s.depsxx+s.depsyy)*NU_MNU-NU3*DEZZ;
      thk=thk+DEZZ*s.thklyl*off;
      etse=H/(H+E);
    }
  } else if(ipla==1){
    H=std::max(K_ZERO,H);
    double S1=s.signxx+s.signyy, S2=s.signxx-s.signyy, S3=s.signxy;
    const double AA=K_FOURTH*S1*S1;
    const double BB=K_THREE_OVER_4*S2*S2+K_THREE*S3*S3;
    const double SVM2=AA+BB;
    { double DEZZ=-(s.depsxx+s.depsyy)*NU_MNU; thk=thk+DEZZ*s.thklyl*off; }
    if(SVM2>YLD*YLD && off==K_ONE){
      double SVM=std::sqrt(SVM2);
      double DPLA_J=(SVM-YLD)/(G3+H);
      etse=H/(H+E);
      const double HI=H*(K_ONE-FISOKIN);
      const double HK=K_TWO_THIRD*H*FISOKIN;
      const double NU3=K_ONE-NU_MNU;
      double DR=K_ZERO,PP=K_ONE,QQ=K_ONE;
      for(int N=1;N<=NITER;N++){
        DPLA_I=DPLA_J;
        double YLD_I=YLD+HI*DPLA_I;
        DR=K_HALF*E*DPLA_I/YLD_I;
        double AAA=K_THREE*HK/E;
        double NU11=U_MNU+AAA, NU21=T_PNU+AAA;
        PP=K_ONE/(K_ONE+DR*NU11);
        QQ=K_ONE/(K_ONE+DR*NU21);
        double P2=PP*PP, Q2=QQ*QQ;
        double F=AA*P2+BB*Q2-YLD_I*YLD_I;
        double DF=-(AA*NU11*P2*PP+BB*NU21*Q2*QQ)*(E-K_TWO*DR*HI)/YLD_I-K_TWO*HI*YLD_I;
        DF=std::copysign(std::max(std::fabs(DF),K_EM20),DF);
        if(DPLA_I>K_ZERO) DPLA_J=std::max(K_ZERO,DPLA_I-F/DF); else DPLA_J=K_ZERO;
      }
      pla=pla+DPLA_I;
      S1=(s.signxx+s.signyy)*PP;
      S2=(s.signxx-s.signyy)*QQ;
      s.signxx=K_HALF*(S1+S2);
      s.signyy=K_HALF*(S1-S2);
      s.signxy=s.signxy*QQ;
      { double DEZZ=-NU3*DR*S1/E; thk=thk+DEZZ*s.thklyl*off; }
      YLD=YLD+HI*DPLA_I;
    }
  } else {            /* IPLAS == 2 */
    H=std::max(K_ZERO,H);
    double SVM2=s.signxx*s.signxx+s.signyy*s.signyy-s.signxx*s.signyy+K_THREE*s.signxy*s.signxy;
    { double DEZZ=-(s.depsxx+s.depsyy)*NU_MNU; thk=thk+DEZZ*s.thklyl*off; }
    const double YLD2=YLD*YLD;
    if(SVM2>YLD2 && off==K_ONE){
      const double NU3=K_ONE-NU_MNU;
      double A=(SVM2-YLD2)/(K_FIVE*SVM2+K_THREE*(-s.signxx*s.signyy+s.signxy*s.signxy));
      double S1=(K_ONE-K_TWO*A)*s.signxx+A*s.signyy;
      double S2=A*s.signxx+(K_ONE-K_TWO*A)*s.signyy;
      double S3=(K_ONE-K_THREE*A)*s.signxy;
      s.signxx=S1; s.signyy=S2; s.signxy=S3;
      double SVM=std::sqrt(SVM2);
      double DPLA=off*(SVM-YLD)/(G3+H);
      DPLA_I=DPLA;
      double HK=H*(K_ONE-FISOKIN);
      YLD=YLD+HK*DPLA;
      SVM=std::sqrt(s.signxx*s.signxx+s.signyy*s.signyy-s.signxx*s.signyy+K_THREE*s.signxy*s.signxy);
      double R=std::min(K_ONE,YLD/std::max(K_EM20,SVM));
      s.signxx=s.signxx*R; s.signyy=s.signyy*R; s.signxy=s.signxy*R;
      pla=pla+DPLA;
      double DEZZ=DPLA*K_HALF*(s.signxx+s.signyy)/YLD;
      DEZZ=-NU3*DEZZ;
      thk=thk+DEZZ*s.thklyl*off;
      etse=H/(H+E);
    }
  }
  /* plastic strain rate filter (sigeps36c.F:976-982) */
  if(vp1){ const double DTINV=K_ONE/std::max(dt1,K_EM20); plap=asrate*DPLA_I*DTINV+(K_ONE-asrate)*plap; }
  /* kinematic part of the hardening (sigeps36c.F:986-1002): the back stress grows along the new stress, which gets it back */
  if(FISOKIN>K_ZERO){
    const double HKIN=FISOKIN*H;
    const double ALPHA=HKIN*DPLA_I/YLD;
    const double SIGPXX=ALPHA*s.signxx, SIGPYY=ALPHA*s.signyy, SIGPXY=ALPHA*s.signxy;
    sigb[0]=sigb[0]+SIGPXX; sigb[1]=sigb[1]+SIGPYY; sigb[2]=sigb[2]+SIGPXY;
    s.signxx=s.signxx+sigb[0]; s.signyy=s.signyy+sigb[1]; s.signxy=s.signxy+sigb[2];
  }
  /* IFAIL = 1: failure on the maximum plastic strain (sigeps36c.F:928-938, no non-local): the element starts its
   * deletion; MULAWC completes it in the same cycle (mulawc.F90:2937-2941) */
  if(m.ifail==1){ if(off==K_ONE && pla>m.epsmax) off=K_FOUR_OVER_5; }
  else if(m.ifail==2){ if(off==K_ONE && (pla>m.epsmax || EPST>m.epsf)) off=K_FOUR_OVER_5; }   /* :940-950 */
  yld_out=YLD;
}

/* ---- SIGEPS02C + M2CPLR, one element (FISOKIN=0) ------------------------------------------ */
void sigeps02c(const orgpu_law2& m, int ipla, int npttot, double dt1, double asrate, const ShellMatIn& in, IpIO& s,
               double& pla, double& epsd, double& temp, bool has_temp, double& off, double off_old, int& ioff_duct,
               double& epchk, double& thk, double& etse, double& sigy, double* sigb /*SIGBAKXX, SIGBAKYY, SIGBAKXY*/)
{
  const double FISOKIN=m.fisokin;
  const int NMAX=3;
  const double SMALL=K_EM7;
  const int iform=m.iform, icc=m.icc, vp=m.vp, israte=m.israte;
  const double young=m.young, g=m.shear, nu=m.nu;
  const double a11=young/(K_ONE-nu*nu);
  const double a12=a11*nu;
  double ca=m.ca, cb=m.cb; const double cn=m.cn, epmx=m.epmx; double ymax=m.sigmx; const double cc=m.cc;
  double epdr=m.epdr; epdr=std::max(epdr*dt1,K_EM20);
  const double tref=m.tref, tmelt=m.tmelt, rhocp=m.rhocp;
  double z3,z4,m_exp,tstar=K_ZERO;
  double tempel= has_temp? temp : K_ZERO;
  if(iform==1){ z3=m.z3; z4=m.z4; m_exp=K_ONE;
    if(has_temp) tstar=std::max(K_ZERO,(tempel-tref)/std::max(tmelt-tref,K_EM20));   /* mulawc.F90:1104-1110 */
  } else { z3=K_ZERO; z4=K_ZERO; m_exp=m.z3; tstar=std::max(K_ZERO,(tempel-tref)/(tmelt-tref)); }
  double EZZ=K_ZERO, epsdot=K_ZERO;
  if(vp==1){ epsdot=epsd*dt1; }
  else if(vp==2){ epsd=asrate*in.epsd_pg+(K_ONE-asrate)*epsd; epsdot=epsd*dt1; }
  else if(vp==3){
    double DAV=(s.epspxx+s.epspyy)*K_THIRD;
    double DEVE1=s.epspxx-DAV, DEVE2=s.epspyy-DAV, DEVE3=-DAV, DEVE4=K_HALF*s.epspxy;
    epsdot=K_HALF*(DEVE1*DEVE1+DEVE2*DEVE2+DEVE3*DEVE3)+DEVE4*DEVE4;
    epsdot=std::sqrt(K_THREE*epsdot)/K_THREE_HALF;
    if(israte>0) epsdot=asrate*epsdot+(K_ONE-asrate)*epsd;
    epsd=epsdot; epsdot=epsdot*dt1;
  }
  /* ---- M2CPLR */
  double CA=ca, CB=cb, YMAX=ymax, H=K_ZERO, DPLA=K_ZERO, YLD;
  etse=K_ONE;
  s.signxx=s.sigoxx; s.signyy=s.sigoyy; s.signxy=s.sigoxy; s.signyz=s.sigoyz; s.signzx=s.sigozx;
  if(FISOKIN>K_ZERO){ s.signxx=s.signxx-sigb[0]; s.signyy=s.signyy-sigb[1]; s.signxy=s.signxy-sigb[2]; }   /* m2cplr.F:115-121 */
  s.signxx=s.signxx+a11*s.depsxx+a12*s.depsyy;
  s.signyy=s.signyy+a12*s.depsxx+a11*s.depsyy;
  s.signxy=s.signxy+g*s.depsxy;
  s.signyz=s.signyz+in.gs*s.depsyz;
  s.signzx=s.signzx+in.gs*s.depszx;
  double EPSP=epsdot, LOGEP=K_ZERO, Q;
  if(cc!=K_ZERO){
    if(iform==0){
      if(israte==0&&vp==2) EPSP=std::max(std::max(std::fabs(s.depsxx),std::fabs(s.depsyy)),K_HALF*std::fabs(s.depsxy));
      EPSP=std::max(EPSP,epdr);
      LOGEP=std::log(EPSP/epdr);
      if(tstar==K_ZERO) Q=(K_ONE+cc*LOGEP);
      else Q=(K_ONE+cc*LOGEP)*(K_ONE-std::exp(m_exp*std::log(tstar)));
      Q=std::max(Q,K_EM20);
      CA=CA*Q; CB=CB*Q;
      if(icc==1) YMAX=YMAX*Q;
    } else if(iform==1){
      if(israte==0&&vp==2) EPSP=std::max(std::max(std::fabs(s.depsxx),std::fabs(s.depsyy)),K_HALF*std::fabs(s.depsxy));
      EPSP=std::max(EPSP,K_EM20);
      LOGEP=std::log(EPSP/epdr);
      Q=LOGEP;
      Q=cc*std::exp((-z3+z4*Q)*tempel);
      if(icc==1) YMAX=YMAX+Q;
      CA=CA+Q;
    }
  } else if(iform==0){
    if(tstar!=K_ZERO){
      Q=K_ONE-std::exp(m_exp*std::log(tstar));
      Q=std::max(Q,K_EM20);
      CA=CA*Q; CB=CB*Q;
    }
  }
  if(pla==K_ZERO) YLD=CA;
  else { double BETA=CB*(K_ONE-m.fisokin); YLD=CA+BETA*std::exp(cn*std::log(pla)); }
  YLD=std::min(YLD,YMAX);
  const double offp=off_old;                 /* M2CPLR receives OFF_OLD as OFF (sigeps02c.F:161) */
  if(ipla==0){
    double SVM=std::sqrt(s.signxx*s.signxx+s.signyy*s.signyy-s.signxx*s.signyy+K_THREE*s.signxy*s.signxy);
    double R=std::min(K_ONE,YLD/(SVM+K_EM15));
    if(R<K_ONE){
      s.signxx=s.signxx*R; s.signyy=s.signyy*R; s.signxy=s.signxy*R;
      DPLA=offp*std::max(K_ZERO,(SVM-YLD)/young);
      double S1=K_HALF*(s.signxx+s.signyy);
      EZZ=DPLA*S1/YLD;
      pla=pla+DPLA;
      epchk=std::max(pla,epchk);
      if(YLD>=YMAX) H=K_ZERO; else H=cn*CB*std::exp((cn-K_ONE)*std::log(pla+SMALL));
      etse=H/(H+young);
    }
  } else if(ipla==1){
    double S1=s.signxx+s.signyy, S2=s.signxx-s.signyy, S3=s.signxy;
    const double A=K_FOURTH*S1*S1;
    const double B=K_THREE_OVER_4*S2*S2+K_THREE*S3*S3;
    const double SVM=std::sqrt(A+B);
    if(SVM>YLD && offp==K_ONE){
      const double NU1=K_ONE/(K_ONE-nu), NU2=K_ONE/(K_ONE+nu);
      if(YLD>=YMAX) H=K_ZERO; else H=cn*CB*std::exp((cn-K_ONE)*std::log(pla+SMALL));
      double DPLA_J=(SVM-YLD)/(K_THREE*g+H);
      etse=H/(H+young);
      double DPLA_I=K_ZERO,DR=K_ZERO,P=K_ONE,Qq=K_ONE;
      if(FISOKIN==K_ZERO){                                 /* m2cplr.F:289-318 */
        const double ANU1=A*NU1, BNU2=K_THREE*B*NU2, H2=K_TWO*H;
        for(int N=1;N<=NMAX;N++){
          DPLA_I=DPLA_J;
          double PLA_I=pla+DPLA_I;
          DPLA=DPLA_J;
          double YLD_I;
          if(PLA_I==K_ZERO) YLD_I=std::min(YMAX,CA);
          else YLD_I=std::min(YMAX,CA+CB*std::exp(cn*std::log(PLA_I)));
          DR=K_HALF*young*DPLA_I/YLD_I;
          P=K_ONE/(K_ONE+DR*NU1);
          Qq=K_ONE/(K_ONE+K_THREE*DR*NU2);
          double P2=P*P, Q2=Qq*Qq;
          double F=A*P2+B*Q2-YLD_I*YLD_I;
          double DF=-(ANU1*P2*P+BNU2*Q2*Qq)*(young-DR*H2)/YLD_I-H2*YLD_I;
          if(DPLA_I>K_ZERO) DPLA_J=std::max(K_ZERO,DPLA_I-F/DF); else DPLA_J=K_ZERO;
        }
      } else {                                             /* :319-363 kinematic / mixed hardening */
        double BETA=H*FISOKIN;
        const double HI=H-BETA, HK=K_TWO_THIRD*BETA;
        const double AAA=K_THREE*HK/young;
        const double NU11=NU1+AAA, NU21=K_THREE*NU2+AAA;
        const double ANU1=A*NU11, BNU2=B*NU21, H2=K_TWO*HI;
        for(int N=1;N<=NMAX;N++){
          DPLA_I=DPLA_J;
          double PLA_I=pla+DPLA_I;
          DPLA=DPLA_J;
          BETA=K_ONE-FISOKIN;
          double YLD_I;
          if(PLA_I==K_ZERO) YLD_I=std::min(YMAX,CA);
          else YLD_I=std::min(YMAX,CA+BETA*CB*std::exp(cn*std::log(PLA_I)));
          DR=K_HALF*young*DPLA_I/YLD_I;
          P=K_ONE/(K_ONE+DR*NU11);
          Qq=K_ONE/(K_ONE+DR*NU21);
          double P2=P*P, Q2=Qq*Qq;
          double F=A*P2+B*Q2-YLD_I*YLD_I;
          double DF=-(ANU1*P2*P+BNU2*Q2*Qq)*(young-DR*H2)/YLD_I-H2*YLD_I;
          if(DPLA_I>K_ZERO) DPLA_J=std::max(K_ZERO,DPLA_I-F/DF); else DPLA_J=K_ZERO;
        }
      }
      pla=pla+DPLA_I;
      epchk=std::max(pla,epchk);
      S1=(s.signxx+s.signyy)*P;
      S2=(s.signxx-s.signyy)*Qq;
      s.signxx=K_HALF*(S1+S2);
      s.signyy=K_HALF*(S1-S2);
      s.signxy=s.signxy*Qq;
      EZZ=DR*S1/young;
      if(FISOKIN>K_ZERO){                                  /* :474-487: the yield stress at the new plastic strain, isotropic part only */
        const double BETA=K_ONE-FISOKIN;
        if(pla==K_ZERO) YLD=CA; else YLD=std::min(YMAX,CA+BETA*CB*std::exp(cn*std::log(pla)));
      }
    }
  } else {
    double SVM2=s.signxx*s.signxx+s.signyy*s.signyy-s.signxx*s.signyy+K_THREE*s.signxy*s.signxy;
    double SVM=std::sqrt(SVM2);
    const double YLD2=YLD*YLD;
    if(SVM2>YLD2 && offp==K_ONE){
      if(YLD>=YMAX) H=K_ZERO; else H=cn*CB*std::exp((cn-K_ONE)*std::log(pla+SMALL));
      etse=H/(H+young);
      double AA=(SVM2-YLD2)/(K_FIVE*SVM2+K_THREE*(-s.signxx*s.signyy+s.signxy*s.signxy));
      double S1=(K_ONE-K_TWO*AA)*s.signxx+AA*s.signyy;
      double S2=AA*s.signxx+(K_ONE-K_TWO*AA)*s.signyy;
      double S3=(K_ONE-K_THREE*AA)*s.signxy;
      s.signxx=S1; s.signyy=S2; s.signxy=S3;
      DPLA=offp*(SVM-YLD)/(K_THREE*g+H);
      pla=pla+DPLA;
      YLD=YLD+H*DPLA;
      SVM=std::sqrt(s.signxx*s.signxx+s.signyy*s.signyy-s.signxx*s.signyy+K_THREE*s.signxy*s.signxy);
      double R=std::min(K_ONE,YLD/std::max(K_EM20,SVM));
      s.signxx=s.signxx*R; s.signyy=s.signyy*R; s.signxy=s.signxy*R;
      EZZ=DPLA*K_HALF*(s.signxx+s.signyy)/YLD;
    }
  }
  /* kinematic part (m2cplr.F:488-499): the back stress grows along the new shifted stress, which then gets it back */
  if(FISOKIN>K_ZERO){
    const double HKIN=FISOKIN*H;
    const double ALPHA=HKIN*DPLA/YLD;
    sigb[0]=sigb[0]+ALPHA*s.signxx; sigb[1]=sigb[1]+ALPHA*s.signyy; sigb[2]=sigb[2]+ALPHA*s.signxy;
    s.signxx=s.signxx+sigb[0]; s.signyy=s.signyy+sigb[1]; s.signxy=s.signxy+sigb[2];
  }
  /* ---- back in SIGEPS02C (:172-230) */
  if(vp==1){ epsdot=DPLA/std::max(K_EM20,dt1); epsd=asrate*epsdot+(K_ONE-asrate)*epsd; }
  sigy=sigy+YLD/npttot;
  if(off==off_old && off>K_ZERO){
    if(off==K_ONE && epchk>=epmx){ off=K_FOUR_OVER_5; ioff_duct=1; }
    else if(off<K_ONE) off=off*K_FOUR_OVER_5;
  }
  EZZ=-(s.depsxx+s.depsyy)*nu-(K_ONE-K_TWO*nu)*EZZ;
  EZZ=EZZ/(K_ONE-nu);
  thk=thk+EZZ*s.thklyl*off;
  if(rhocp>K_ZERO && has_temp) temp=tempel+sigy*DPLA/rhocp;
}

} // namespace

/* ---- CMAIN3 -> LAYINI -> MULAWC for one element ------------------------------------------- */
void orc_cmain3(const Oracle& o, OrcShellGroup& g, int i, bool flag_zcfac, ShellMatIn& in, ShellMatOut& out)
{
  const int nel=g.nel, npt=g.prop.npt;
  const double dt1=o.DT1;
  const double DM=g.prop.dm;
  double* FOR=g.FOR.data(); double* MOM=g.MOM.data();
#define F_(k) FOR[(size_t)(k-1)*nel+i]
#define M_(k) MOM[(size_t)(k-1)*nel+i]
  /* mulawc.F90:542-546 */
  double degmb=F_(1)*in.exx+F_(2)*in.eyy+F_(3)*in.exy+F_(4)*in.eyz+F_(5)*in.exz;
  double degfx=M_(1)*in.kxx+M_(2)*in.kyy+M_(3)*in.kxy;
  const double vol0=in.area*in.thk0;
  double thkn=g.THK[i];
  for(int k=1;k<=5;k++) F_(k)=K_ZERO;
  for(int k=1;k<=3;k++) M_(k)=K_ZERO;
  double sigy=out.sigy;
  if(g.law==2 || !flag_zcfac) sigy=K_ZERO;
  double zcfac1=K_ZERO, zcfac2= flag_zcfac? K_ONE : K_ZERO;
  double etse=K_ONE;
  double off=in.off; const double off_old=off;
  int ioff_duct=0;
  double epchk=K_ZERO, viscmx=K_ZERO, ssp=in.ssp, ssp_eq=K_ZERO;
  const double dtinv=dt1/std::max(dt1*dt1,K_EM20);
  /* strain-rate filter of the layer material (mulawc.F90:695-700) */
  const int israte= (g.law==36)? g.m36.israte : g.m2.israte;
  const double pm9= (g.law==36)? g.m36.asrate : g.m2.asrate;
  double asrate; if(israte>0) asrate=std::min(K_ONE,pm9*dt1); else asrate=K_ONE;
  for(int ipt=1;ipt<=npt;ipt++){
    OrcShellGroup::Lbuf& lb=g.ip[ipt-1];
    const double thkly=orc_quad_tab(1)[(npt-1)*11+(ipt-1)];          /* WF, layini.F:250 */
    const double posly=orc_quad_tab(0)[(npt-1)*11+(ipt-1)]+K_ZERO;   /* Z0, layini.F:251 (ZSHIFT=0) */
    const double wmc=orc_quad_tab(2)[(npt-1)*11+(ipt-1)];            /* WM, mulawc.F90:771-773 */
    IpIO s;
    const double pla0=lb.pla[i];                                   /* mulawc.F90: PLA0 kept for the failure models */
    s.thklyl=thkly*in.thk0;
    const double zt=posly*in.thk0;
    s.depsxx=in.exx+zt*in.kxx;
    s.depsyy=in.eyy+zt*in.kyy;
    s.depsxy=in.exy+zt*in.kxy;
    s.depsyz=in.eyz; s.depszx=in.exz;
    s.epspxx=s.depsxx*dtinv; s.epspyy=s.depsyy*dtinv; s.epspxy=s.depsxy*dtinv;
    { const double* GS_=g.STRA.data(); s.epsxx=GS_[i]+zt*GS_[5*nel+i]; s.epsyy=GS_[nel+i]+zt*GS_[6*nel+i]; s.epsxy=GS_[2*nel+i]+zt*GS_[7*nel+i]; }
    s.sigoxx=lb.sig[i]; s.sigoyy=lb.sig[nel+i]; s.sigoxy=lb.sig[2*nel+i]; s.sigoyz=lb.sig[3*nel+i]; s.sigozx=lb.sig[4*nel+i];
    if(g.law==36){
      double sb[3]={lb.sigb[i],lb.sigb[nel+i],lb.sigb[2*nel+i]};
      sigeps36c(o,g.m36,g.prop.ipla,asrate,in,s,lb.pla[i],lb.epsd[i],&lb.vartmp[(size_t)g.nvartmp*i],off,thkn,ssp,viscmx,etse,sigy,sb,lb.plap[i],dt1);
      lb.sigb[i]=sb[0]; lb.sigb[nel+i]=sb[1]; lb.sigb[2*nel+i]=sb[2];
    } else {
      double sb[3]={lb.sigb[i],lb.sigb[nel+i],lb.sigb[2*nel+i]};
      sigeps02c(g.m2,g.prop.ipla,npt,dt1,asrate,in,s,lb.pla[i],lb.epsd[i],lb.temp[i],g.m2.has_temp!=0,off,off_old,ioff_duct,
                epchk,thkn,etse,sigy,sb);
      lb.sigb[i]=sb[0]; lb.sigb[nel+i]=sb[1]; lb.sigb[2*nel+i]=sb[2];
    }
    viscmx=std::max(DM,viscmx);
    double sigoff=K_ONE;                                           /* mulawc.F90:2037 */
    if(g.fail.irupt==1){
      /* mulawc.F90:2064-2069: DPLA = LBUF%PLA - PLA0, EPSD = LBUF%EPSD; :2118-2127 FAIL_JOHNSON_C (fail_johnson_c.F:111-130) */
      const double DPLA=lb.pla[i]-pla0, EPSP=lb.epsd[i];
      const orgpu_fail& f=g.fail;
      if(off==K_ONE && lb.foff[i]==K_ONE && DPLA>K_ZERO){
        const double P=K_THIRD*(s.signxx+s.signyy);
        const double SVM=std::sqrt(s.signxx*s.signxx+s.signyy*s.signyy-s.signxx*s.signyy+K_THREE*s.signxy*s.signxy);
        double EPSF=f.d3*P/std::max(K_EM20,SVM);
        EPSF=(f.d1+f.d2*std::exp(EPSF));
        if(f.d4!=K_ZERO) EPSF=EPSF*(K_ONE+f.d4*std::log(std::max(K_ONE,EPSP/f.epsp0)));
        EPSF=std::max(EPSF,f.epsf_min);
        if(EPSF>K_ZERO) lb.dfmax[i]=lb.dfmax[i]+DPLA/EPSF;
        if(lb.dfmax[i]>=K_ONE) lb.foff[i]=K_ZERO;
      }
      lb.dfmax[i]=std::min(K_ONE,lb.dfmax[i]);                     /* :143-145 */
      if(lb.foff[i]==K_ZERO){ lb.off[i]=K_ZERO; sigoff=K_ZERO; }   /* mulawc.F90:2608-2616 */
    }
    lb.sig[i]=s.signxx*sigoff; lb.sig[nel+i]=s.signyy*sigoff; lb.sig[2*nel+i]=s.signxy*sigoff;   /* :2633-2637 */
    lb.sig[3*nel+i]=s.signyz*sigoff; lb.sig[4*nel+i]=s.signzx*sigoff;
    F_(1)=F_(1)+thkly*s.signxx; F_(2)=F_(2)+thkly*s.signyy; F_(3)=F_(3)+thkly*s.signxy;
    F_(4)=F_(4)+thkly*s.signyz; F_(5)=F_(5)+thkly*s.signzx;
    M_(1)=M_(1)+wmc*s.signxx; M_(2)=M_(2)+wmc*s.signyy; M_(3)=M_(3)+wmc*s.signxy;
    if(g.law!=2){
      if(flag_zcfac){ zcfac1=zcfac1+etse*thkly; zcfac2=std::min(etse,zcfac2); }
    } else {
      if(flag_zcfac){ zcfac1=zcfac1+etse/npt; zcfac2=std::min(etse,zcfac2); }
    }
    ssp_eq=ssp_eq+ssp*thkly;
  }
  /* FAIL_SETOFF_C, NLAY = 1 (mulawc.F90:2912-2920 -> fail_setoff_c.F:123-186) */
  if(g.fail.irupt==1){
    double PTHKF=g.fail.pthk; const double P_THICKG=g.fail.pthickg;
    if(PTHKF>K_ZERO){ PTHKF=std::min(PTHKF,std::fabs(P_THICKG)); PTHKF=std::max(std::min(PTHKF,K_ONE-K_EM06),K_EM06); }
    else if(PTHKF<K_ZERO){ PTHKF=std::max(PTHKF,-std::fabs(P_THICKG)); PTHKF=std::min(std::max(PTHKF,-K_ONE+K_EM06),-K_EM06); }
    else PTHKF=P_THICKG;
    double THFACT=K_ZERO, NPFAIL=K_ZERO;
    for(int ipt=1;ipt<=npt;ipt++){
      if(off==K_ONE){
        if(g.ip[ipt-1].foff[i]<K_ONE){
          THFACT=THFACT+orc_quad_tab(1)[(npt-1)*11+(ipt-1)];
          NPFAIL=NPFAIL+K_ONE/npt;
        }
      }
    }
    if(off==K_ONE){
      if(((THFACT>=PTHKF)&&(PTHKF>K_ZERO)) || ((NPFAIL>=std::fabs(PTHKF))&&(PTHKF<K_ZERO))) off=K_FOUR_OVER_5;
    }
  }
  /* tail (mulawc.F90:2934-3091) */
  if((off==K_FOUR_OVER_5 && ioff_duct==0) || (off>K_ZERO && off_old<K_EM01)) off=K_ZERO;
  g.THK[i]=std::max(thkn,K_EM30);
  const double fact=K_ONEP414*DM;
  const double visc=fact*ssp*std::sqrt(in.area)*dtinv*in.rho;
  F_(1)=F_(1)+visc*(in.exx+K_HALF*in.eyy);
  F_(2)=F_(2)+visc*(in.eyy+K_HALF*in.exx);
  F_(3)=F_(3)+visc*in.exy*K_THIRD;
  for(int k=1;k<=5;k++) F_(k)=F_(k)*off*K_ONE;
  for(int k=1;k<=3;k++) M_(k)=M_(k)*off*K_ONE;
  degmb=degmb+F_(1)*in.exx+F_(2)*in.eyy+F_(3)*in.exy+F_(4)*in.eyz+F_(5)*in.exz;
  degfx=degfx+M_(1)*in.kxx+M_(2)*in.kyy+M_(3)*in.kxy;
  const double vol2=K_HALF*vol0;
  g.EINT[i]=g.EINT[i]+degmb*vol2;
  g.EINT[nel+i]=g.EINT[nel+i]+degfx*in.thk0*vol2;
#undef F_
#undef M_
  in.off=off;
  out.ssp=ssp; out.viscmx=viscmx; out.sigy=sigy; out.zcfac1=zcfac1; out.zcfac2=zcfac2; out.ssp_eq=ssp_eq; out.vol0=vol0;
}
