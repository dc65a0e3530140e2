/* oracle/solid.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of the 8-node brick internal-force path, Lagrangian (JLAG=1), JCVT=0, ISORTH=0,
 * ISROT=0, NPT=1, no ALE/Euler/thermal-FE/nonlocal branches:
 *   SFORC3   engine/source/elements/solid/solide/sforc3.F:131 (call order :476-1764)
 * Each block below names the routine and lines it follows.  Statement order and expression
 * shapes are kept so that, compiled with -ffp-contract=off, the arithmetic is the reference's.
 */
#include "oracle.h"
#include "shell.h"   /* orc_vinter */

namespace {

struct Vec { double v[MVSIZ]; double& operator[](int i){return v[i];} double operator[](int i) const {return v[i];} };

/* SLEN  solid/solide/slen.F:68-88 : face "area" metric E*G-F*F, running max into AREAM */
inline void slen(const Vec& X1,const Vec& X2,const Vec& X3,const Vec& X4,
                 const Vec& Y1,const Vec& Y2,const Vec& Y3,const Vec& Y4,
                 const Vec& Z1,const Vec& Z2,const Vec& Z3,const Vec& Z4,
                 int j, double (*AREA)[6], Vec& AREAM, int nel)
{
  for (int i=0;i<nel;i++){
    double X13=X3[i]-X1[i], X24=X4[i]-X2[i];
    double Y13=Y3[i]-Y1[i], Y24=Y4[i]-Y2[i];
    double Z13=Z3[i]-Z1[i], Z24=Z4[i]-Z2[i];
    double FS1=X13-X24, FT1=X13+X24;
    double FS2=Y13-Y24, FT2=Y13+Y24;
    double FS3=Z13-Z24, FT3=Z13+Z24;
    double E=FS1*FS1+FS2*FS2+FS3*FS3;
    double F=FS1*FT1+FS2*FT2+FS3*FT3;
    double G=FT1*FT1+FT2*FT2+FT3*FT3;
    AREA[i][j]=E*G-F*F;
    AREAM[i]=std::max(AREA[i][j],AREAM[i]);
  }
}

/* jacobian block of SDERI3 (sderi3.F:105-149 and :218-260): fills JAC1..9, the three cofactors, VOLDP */
struct JacOut { double J1,J2,J3,J4,J5,J6,J7,J8,J9,c5968,c6749,c4857,voldp; };
inline JacOut sderi_jac(const double* x,const double* y,const double* z){ /* x[0..7] = X1..X8 */
  JacOut r;
  double X17=x[6]-x[0], X28=x[7]-x[1], X35=x[4]-x[2], X46=x[5]-x[3];
  double Y17=y[6]-y[0], Y28=y[7]-y[1], Y35=y[4]-y[2], Y46=y[5]-y[3];
  double Z17=z[6]-z[0], Z28=z[7]-z[1], Z35=z[4]-z[2], Z46=z[5]-z[3];
  r.J1=X17+X28-X35-X46;
  r.J2=Y17+Y28-Y35-Y46;
  r.J3=Z17+Z28-Z35-Z46;
  double X_17_46=X17+X46, X_28_35=X28+X35;
  double Y_17_46=Y17+Y46, Y_28_35=Y28+Y35;
  double Z_17_46=Z17+Z46, Z_28_35=Z28+Z35;
  r.J4=X_17_46+X_28_35; r.J5=Y_17_46+Y_28_35; r.J6=Z_17_46+Z_28_35;
  r.J7=X_17_46-X_28_35; r.J8=Y_17_46-Y_28_35; r.J9=Z_17_46-Z_28_35;
  r.c5968=r.J5*r.J9-r.J6*r.J8;
  r.c6749=r.J6*r.J7-r.J4*r.J9;
  r.c4857=r.J4*r.J8-r.J5*r.J7;
  r.voldp=K_ONE_OVER_64*(r.J1*r.c5968+r.J2*r.c6749+r.J3*r.c4857);
  return r;
}

} // namespace

/* MSTRAIN_RATE  materials/mat_share/mstrain_rate.F:60-91 */
static void mstrain_rate(int nel,int israte,double asrate,double* epsd,int idev,
                         const double* e1,const double* e2,const double* e3,
                         const double* e4,const double* e5,const double* e6)
{
  double epsdot[MVSIZ];
  if (israte>=0){
    if (idev==0){
      for(int i=0;i<nel;i++){
        double E1=e1[i],E2=e2[i],E3=e3[i],E4=K_HALF*e4[i],E5=K_HALF*e5[i],E6=K_HALF*e6[i];
        double epsp=E1*E1+E2*E2+E3*E3+K_TWO*(E4*E4+E5*E5+E6*E6);
        epsdot[i]=std::sqrt(epsp);
      }
    } else {
      for(int i=0;i<nel;i++){
        double dav=(e1[i]+e2[i]+e3[i])*K_THIRD;
        double E1=e1[i]-dav,E2=e2[i]-dav,E3=e3[i]-dav,E4=K_HALF*e4[i],E5=K_HALF*e5[i],E6=K_HALF*e6[i];
        double epsp=K_HALF*(E1*E1+E2*E2+E3*E3)+E4*E4+E5*E5+E6*E6;
        epsdot[i]=std::sqrt(K_THREE*epsp)/K_THREE_HALF;
      }
    }
  }
  if (israte==0){ for(int i=0;i<nel;i++) epsd[i]=epsdot[i]; }
  else if (israte>0){ for(int i=0;i<nel;i++) epsd[i]=asrate*epsdot[i]+(K_ONE-asrate)*epsd[i]; }
}

/* MQVISCB  materials/mat_share/mqviscb.F:44 -- Lagrangian 3-D branch (IMPL=0, N2D=0, JTHE=0,
 * IDTMINS/=2, ALE_OR_EULER=0, NODADT=0, IDTMIN(1)=0), NPG=1 => FACPG=1.
 * Outputs QVIS(=QNEW), SSP_EQ, STI, lowers DT2T/NELTST/ITYPTST. */
static void mqviscb(const Oracle& o,const OrcSolidGroup& g,int nel,const int* ngl,
                    const double* off,const double* rho,const double* ssp,double* sti,
                    double& dt2t,int& neltst,int& ityptst,const double* offg,
                    const double* vol /*VNEW*/,const double* vd2,const double* deltax,const double* vis,
                    const double* d1,const double* d2,const double* d3,
                    double* qvis,double* ssp_eq,const double* vol0,const double* rhoref,double facq0)
{
  const int ismstr=g.prop.ismstr;
  double dd[MVSIZ],al[MVSIZ],dtx[MVSIZ],ad[MVSIZ],qx[MVSIZ],cx[MVSIZ],rho0[MVSIZ],nrho[MVSIZ];
  /* :141-153 IMPL==ZERO branch */
  for(int i=0;i<nel;i++){ dd[i]=-d1[i]-d2[i]-d3[i]; ad[i]=K_ZERO; al[i]=K_ZERO; cx[i]=ssp[i]+std::sqrt(vd2[i]); }
  const double visi=K_ONE, facq=facq0;
  /* :178-191 */
  const double facpg=1.0;
  for(int i=0;i<nel;i++){
    if(off[i]==K_ONE){
      if(vol[i]>K_ZERO) al[i]=std::pow(vol[i],1.0/3.0);   /* VOL**(REAL_ONE/REAL_THREE), :184 */
      else al[i]=K_ZERO;
      ad[i]=std::max(K_ZERO,dd[i]);
    }
  }
  /* :195-220 */
  for(int i=0;i<nel;i++){ rho0[i]=g.mat.rho0; nrho[i]=std::sqrt(rhoref[i]*rho0[i]); }
  const double qa=facq*g.prop.qa, qb=facq*g.prop.qb;
  const double cns1_0=facpg*g.prop.cns1, cns2_0=facpg*g.prop.cns2;
  const double qaa_0=qa*qa;
  for(int i=0;i<nel;i++){
    double cns1=cns1_0*al[i]*nrho[i]*ssp[i]*off[i];
    double cns2=cns2_0*al[i]*nrho[i]*ssp[i]*off[i];
    double qaa=qaa_0*ad[i];
    qx[i]=qb*ssp[i]+al[i]*qaa
         + visi*K_TWO*vis[i]/std::max(K_EM20,rho[i]*deltax[i])
         + (cns1+visi*cns2)/std::max(K_EM20,rhoref[i]*deltax[i]);
    qvis[i]=rho[i]*ad[i]*al[i]*(qaa*al[i]+qb*ssp[i]);
  }
  /* :259-262 */
  for(int i=0;i<nel;i++){
    ssp_eq[i]=std::max(K_EM20,qx[i]+std::sqrt(qx[i]*qx[i]+cx[i]*cx[i]));
    dtx[i]=deltax[i]/ssp_eq[i];
  }
  /* :308-416 : KDTSMSTR==1.AND.ISMSTR==1 is the only first-branch case reachable with IDTMIN(1)=0 */
  const bool kdtsm = (ismstr==1);   /* KDTSMSTR defaults to 1 */
  for(int i=0;i<nel;i++){
    sti[i]=K_ZERO;
    if(off[i]==K_ZERO||offg[i]<K_ZERO) continue;
    double tidt=K_ONE/dtx[i], trho, tvol;
    if(kdtsm && offg[i]>K_ONE){ trho=rho0[i]*tidt; tvol=vol0[i]*tidt; }
    else                      { trho=rho[i]*tidt;  tvol=vol[i]*tidt; }
    sti[i]=trho*tvol;
  }
  for(int i=0;i<nel;i++) dtx[i]=o.ctl.dtfac_brick*dtx[i];
  /* NODADT==0 : :351-355 / :411-415 ; with /DT/NODA the element does not lower DT2T (:351, :411, :621) */
  if(o.ctl.nodadt!=0) return;
  for(int i=0;i<nel;i++)
    if(vol[i]>K_ZERO && (off[i]!=K_ZERO && offg[i]>=K_ZERO)) dt2t=std::min(dtx[i],dt2t);
  /* :621-631 argmin bookkeeping (note: ">" so the LAST element at the minimum wins) */
  for(int i=0;i<nel;i++){
    if(dtx[i]>dt2t || off[i]<=K_ZERO || offg[i]<=K_ZERO) continue;
    if(vol[i]<=K_ZERO) continue;
    dt2t=dtx[i]; neltst=ngl[i]; ityptst=1;
  }
}

/* M2LAW  materials/mat/mat002/m2law.F:38 -- explicit (IMPL_S=0), IEOS=0, JSPH=0, JTHE=0 */
static void m2law(const Oracle& o,OrcSolidGroup& g,int nel,const int* ngl,
                  const double* off,double* sig /*6*nel comp-major*/,double* eint,const double* rho,
                  double* qold,double* epxe,double* epsd,const double* vol,double* stifn,
                  double& dt2t,int& neltst,int& ityptst,const double* offg,
                  const double* amu,const double* vol_avg,double* ssp,const double* dvol,
                  const double* vnew,const double* vd2,const double* deltax,const double* vis,
                  const double* d1,const double* d2,const double* d3,const double* d4,const double* d5,const double* d6,
                  double* qnew,double* ssp_eq,
                  const double* sold1,const double* sold2,const double* sold3,
                  const double* sold4,const double* sold5,const double* sold6,
                  const double* tstar,double* tempel,double* dmg,const double* rhoref,double* sigbak_of_call /*6*nel comp-major, or null*/,
                  double* dpla_out,double* epsp_out /* what MMAIN hands the failure models, or null */)
{
  const orgpu_law2& m=g.mat;
  const double facq0=K_ONE;
  const int iform=m.iform, icc=m.icc, vp=m.vp, israte=m.israte;
  const double rho0=m.rho0, bulk=m.bulk, g0=m.shear;
  const double ca0=m.ca, cb=m.cb, cn=m.cn, epmx=m.epmx, sigm0=m.sigmx, cc=m.cc, epdr=m.epdr;
  const double fisokin=m.fisokin;
  double asrate=m.asrate; const double z3=m.z3;
  asrate=std::min(K_ONE,asrate*o.DT1);                       /* :155 */
  const double rhocp=m.rhocp;
  double rhocpi = (rhocp>K_ZERO)? K_ONE/rhocp : K_ZERO;      /* :160-164 */
  double z4=0; if(iform==1) z4=m.z4;
  double G[MVSIZ],CA[MVSIZ],SIGMX[MVSIZ],AJ2[MVSIZ],DAV[MVSIZ],EPD[MVSIZ],AK[MVSIZ],QH[MVSIZ],SIGY[MVSIZ],DPLA[MVSIZ];
  auto S=[&](int i,int k)->double&{ return sig[k*nel+i]; };
  const double DT1=o.DT1;
  for(int i=0;i<nel;i++){ G[i]=g0*off[i]; CA[i]=ca0; SIGMX[i]=sigm0; }   /* :171-176 */
  double* SIGBAK = nullptr;                                                /* LBUF%SIGB of the elements of this call */
  auto B=[&](int i,int k)->double&{ return SIGBAK[k*nel+i]; };
  std::vector<double> SIGE;
  if(fisokin>K_ZERO){                                                      /* :181-190 kinematic hardening: stress shifted by the back stress */
    SIGBAK = sigbak_of_call;
    for(int i=0;i<nel;i++) for(int k=0;k<6;k++) S(i,k)=S(i,k)-B(i,k);
  }
  for(int i=0;i<nel;i++){                                                  /* :192-210 */
    double P=-K_THIRD*(S(i,0)+S(i,1)+S(i,2));
    DAV[i]=-K_THIRD*(d1[i]+d2[i]+d3[i]);
    double G1=DT1*G[i];
    double G2=K_TWO*G1;
    ssp[i]=std::sqrt((K_ONEP333*G[i]+bulk)/rho0);
    S(i,0)=S(i,0)+P+G2*(d1[i]+DAV[i]);
    S(i,1)=S(i,1)+P+G2*(d2[i]+DAV[i]);
    S(i,2)=S(i,2)+P+G2*(d3[i]+DAV[i]);
    S(i,3)=S(i,3)+G1*d4[i];
    S(i,4)=S(i,4)+G1*d5[i];
    S(i,5)=S(i,5)+G1*d6[i];
    AJ2[i]=K_HALF*(S(i,0)*S(i,0)+S(i,1)*S(i,1)+S(i,2)*S(i,2))
          + S(i,3)*S(i,3)+S(i,4)*S(i,4)+S(i,5)*S(i,5);
    AJ2[i]=std::sqrt(K_THREE*AJ2[i]);
  }
  if(fisokin>K_ZERO){ SIGE.resize((size_t)6*nel); for(int i=0;i<nel;i++) for(int k=0;k<6;k++) SIGE[k*nel+i]=S(i,k); }   /* :213-222 */
  const int idev=vp-2;                                                     /* :226-228 */
  mstrain_rate(nel,israte,asrate,epsd,idev,d1,d2,d3,d4,d5,d6);
  if(epsp_out) for(int i=0;i<nel;i++) epsp_out[i]=epsd[i];                 /* :229 EPSP = EPSD */
  for(int i=0;i<nel;i++) EPD[i]=K_ONE;                                     /* :231 */
  if(cc!=K_ZERO){                                                          /* :233-260 */
    if(vp==1){ for(int i=0;i<nel;i++){ EPD[i]=std::max(epsd[i],epdr); EPD[i]=std::log(EPD[i]/epdr);} }
    else     { for(int i=0;i<nel;i++){ EPD[i]=std::max(epsd[i],K_EM15); EPD[i]=std::log(EPD[i]/epdr);} }
    if(iform==0){
      for(int i=0;i<nel;i++){
        double MT=std::max(K_EM15,z3);
        EPD[i]=std::max(K_ZERO,EPD[i]);
        EPD[i]=(K_ONE+cc*EPD[i])*(K_ONE-std::pow(tstar[i],MT));
        if(icc==1) SIGMX[i]=sigm0*EPD[i];
      }
    } else if(iform==1){
      for(int i=0;i<nel;i++){
        EPD[i]=cc*std::exp((-z3+z4*EPD[i])*tempel[i]);
        if(icc==1) SIGMX[i]=sigm0+EPD[i];
        CA[i]=ca0+EPD[i];
        EPD[i]=K_ONE;
      }
    }
  } else if(iform==0){                                                     /* :261-265 */
    double MT=std::max(K_EM15,z3);
    for(int i=0;i<nel;i++){ EPD[i]=K_ONE-std::pow(tstar[i],MT); if(icc==1) SIGMX[i]=sigm0*EPD[i]; }
  }
  if(fisokin==K_ZERO){
  /* isotropic hardening :269-299 */
  if(cn==K_ONE){
    for(int i=0;i<nel;i++){ AK[i]=CA[i]+cb*epxe[i]; QH[i]=cb*EPD[i]; }
  } else {
    for(int i=0;i<nel;i++){
      if(epxe[i]>K_ZERO){
        AK[i]=CA[i]+cb*std::pow(epxe[i],cn);
        if(cn>K_ONE) QH[i]=(cb*cn*std::pow(epxe[i],(cn-K_ONE)))*EPD[i];
        else         QH[i]=(cb*cn/std::pow(epxe[i],(K_ONE-cn)))*EPD[i];
      } else { AK[i]=CA[i]; QH[i]=K_ZERO; }
    }
  }
  for(int i=0;i<nel;i++){
    AK[i]=AK[i]*EPD[i];
    if(SIGMX[i]<AK[i]){ AK[i]=SIGMX[i]; QH[i]=K_ZERO; }
    SIGY[i]=AK[i];
    if(epxe[i]>epmx){ AK[i]=K_ZERO; QH[i]=K_ZERO; }
  }
  } else {
  /* kinematic / mixed hardening :300-337: the isotropic share BETA of the hardening stays in the yield stress */
  for(int i=0;i<nel;i++){
    const double BETA=K_ONE-fisokin;
    if(cn==K_ONE){
      SIGY[i]=CA[i]+cb*epxe[i];
      AK[i]=CA[i]+BETA*cb*epxe[i];
      QH[i]=cb*EPD[i];
    } else {
      if(epxe[i]>K_ZERO){
        SIGY[i]=CA[i]+cb*std::pow(epxe[i],cn);
        AK[i]=CA[i]+BETA*cb*std::pow(epxe[i],cn);
        if(cn>K_ONE) QH[i]=(cb*cn*std::pow(epxe[i],(cn-K_ONE)))*EPD[i];
        else         QH[i]=(cb*cn/std::pow(epxe[i],(K_ONE-cn)))*EPD[i];
      } else { AK[i]=CA[i]; SIGY[i]=CA[i]; QH[i]=K_ZERO; }
    }
    AK[i]=AK[i]*EPD[i];
    SIGY[i]=SIGY[i]*EPD[i];
    if(SIGMX[i]<AK[i]){ AK[i]=SIGMX[i]; QH[i]=K_ZERO; }
    SIGY[i]=std::min(SIGY[i],SIGMX[i]);
    if(epxe[i]>epmx){ AK[i]=K_ZERO; QH[i]=K_ZERO; }
  }
  }
  /* radial return :339-354 */
  for(int i=0;i<nel;i++){
    double SCALE=std::min(K_ONE,AK[i]/std::max(AJ2[i],K_EM15));
    DPLA[i]=(K_ONE-SCALE)*AJ2[i]/std::max(K_THREE*G[i]+QH[i],K_EM15);
    AK[i]=AK[i]+(K_ONE-fisokin)*DPLA[i]*QH[i];
    SCALE=std::min(K_ONE,AK[i]/std::max(AJ2[i],K_EM15));
    S(i,0)=SCALE*S(i,0); S(i,1)=SCALE*S(i,1); S(i,2)=SCALE*S(i,2);
    S(i,3)=SCALE*S(i,3); S(i,4)=SCALE*S(i,4); S(i,5)=SCALE*S(i,5);
    epxe[i]=epxe[i]+DPLA[i];
  }
  if(fisokin>K_ZERO){                                                      /* :364-390 back stress along the plastic corrector */
    for(int i=0;i<nel;i++){
      double DS[6]; for(int k=0;k<6;k++) DS[k]=SIGE[k*nel+i]-S(i,k);
      const double HKIN=K_TWO_THIRD*fisokin*QH[i];
      const double ALPHA=HKIN/std::max(K_TWO*G[i]+HKIN,K_EM15);
      for(int k=0;k<6;k++) B(i,k)=B(i,k)+ALPHA*DS[k];
      for(int k=0;k<6;k++) S(i,k)=S(i,k)+B(i,k);
    }
  }
  /* :393-407 MQVISCB */
  double bid[MVSIZ]; for(int i=0;i<MVSIZ;i++) bid[i]=K_ZERO; (void)bid;
  mqviscb(o,g,nel,ngl,off,rho,ssp,stifn,dt2t,neltst,ityptst,offg,vnew,vd2,deltax,vis,d1,d2,d3,
          qnew,ssp_eq,vol,rhoref,facq0);
  const double DTA=K_HALF*DT1;                                             /* :421 */
  for(int i=0;i<nel;i++) if(epxe[i]>epmx && dmg[i]==K_ZERO) dmg[i]=K_ONE;  /* :425-431 */
  /* IEOS==0 :433-453 */
  double pnew;
  for(int i=0;i<nel;i++){
    pnew=bulk*amu[i];
    S(i,0)=(S(i,0)-pnew)*off[i];
    S(i,1)=(S(i,1)-pnew)*off[i];
    S(i,2)=(S(i,2)-pnew)*off[i];
    S(i,3)=S(i,3)*off[i]; S(i,4)=S(i,4)*off[i]; S(i,5)=S(i,5)*off[i];
  }
  for(int i=0;i<nel;i++){
    double E1=d1[i]*(sold1[i]+S(i,0));
    double E2=d2[i]*(sold2[i]+S(i,1));
    double E3=d3[i]*(sold3[i]+S(i,2));
    double E4=d4[i]*(sold4[i]+S(i,3));
    double E5=d5[i]*(sold5[i]+S(i,4));
    double E6=d6[i]*(sold6[i]+S(i,5));
    double EINC=vol_avg[i]*(E1+E2+E3+E4+E5+E6)*DTA-K_HALF*dvol[i]*(qold[i]+qnew[i]);
    eint[i]=(eint[i]+EINC*off[i])/std::max(K_EM15,vol[i]);
  }
  for(int i=0;i<nel;i++){ qold[i]=qnew[i]; }                               /* :477-481 (DEFP,SIGY outputs unused) */
  if(dpla_out) for(int i=0;i<nel;i++) dpla_out[i]=DPLA[i];
  if(vp==1){                                                               /* :539-544 */
    for(int i=0;i<nel;i++){
      double plap=DPLA[i]/std::max(K_EM20,DT1);
      epsd[i]=asrate*plap+(K_ONE-asrate)*epsd[i];
    }
  }
  /* :552-561 adiabatic heating (JTHE=0, RHOCP>0) */
  if(rhocp>K_ZERO){
    for(int i=0;i<nel;i++){ SIGY[i]=std::max(SIGY[i],AK[i]); tempel[i]=tempel[i]+SIGY[i]*DPLA[i]*rhocpi; }
  }
}

/* SIGEPS36  materials/mat/mat036/sigeps36.F:35 -- the solid (3-D) LAW36, explicit (IMPL_S=0), built
 * envelope as for shells: VP=0, FISOKIN=0, IFAIL=0, OPTE=0, CE1=0, PFUN=0, IEOS=0.
 * uparam slots: :170-196 ; predictor :266-290 ; yield from the tables :398-525 ; projection :530-603 ;
 * pressure :1453-1460 ; OFF relaxation :1507-1510. */
static void sigeps36(const Oracle& o,const orgpu_law36& m,int nel,int ipla,
                     const double* de1,const double* de2,const double* de3,const double* de4,const double* de5,const double* de6,
                     const double* so1,const double* so2,const double* so3,const double* so4,const double* so5,const double* so6,
                     double* s1,double* s2,double* s3,double* s4,double* s5,double* s6,
                     double* soundsp,double* viscmax,double* off,const double* epsp,double* yld,double* pla,double* dpla1,
                     const double* amu,int* vartmp /*(nel,nvartmp) column-major: VARTMP(I,k) -> vartmp[(k-1)*nel+i]*/,
                     const double* es /*total strains (6,nel) comp-major, IFAIL = 2*/)
{
  const int nrate=m.nrate;
  const double G=m.shear, G2=m.g2, G3=m.g3, BULK=m.bulk, SSP=m.ssp3d, FISOKIN=m.fisokin;
  double P0[MVSIZ],H[MVSIZ],R[MVSIZ];
  for(int i=0;i<nel;i++) soundsp[i]=SSP;                                   /* :199-207 */
  for(int i=0;i<nel;i++){                                                  /* :266-281 */
    double DAV=(de1[i]+de2[i]+de3[i])*K_THIRD;
    P0[i]=-(so1[i]+so2[i]+so3[i])*K_THIRD;
    s1[i]=so1[i]+P0[i]+G2*(de1[i]-DAV);
    s2[i]=so2[i]+P0[i]+G2*(de2[i]-DAV);
    s3[i]=so3[i]+P0[i]+G2*(de3[i]-DAV);
  }
  for(int i=0;i<nel;i++){ s4[i]=so4[i]+G*de4[i]; s5[i]=so5[i]+G*de5[i]; s6[i]=so6[i]+G*de6[i]; }   /* :283-285 */
  for(int i=0;i<nel;i++){ viscmax[i]=K_ZERO; dpla1[i]=K_ZERO; }            /* :292-295 (PFAC=1) */
  double FAIL[MVSIZ],EPSTT[MVSIZ];
  for(int i=0;i<nel;i++){ FAIL[i]=K_ONE; EPSTT[i]=K_ZERO; }
  if(m.ifail>1){                                                           /* :299-360 largest principal strain, 4 Newton steps */
    const double EPSR1F=std::min(m.epsr1,m.epsf);
    for(int i=0;i<nel;i++){
      const double EXX=es[i],EYY=es[nel+i],EZZ=es[2*nel+i],EXY=es[3*nel+i],EYZ=es[4*nel+i],EZX=es[5*nel+i];
      double DAV=(EXX+EYY+EZZ)*K_THIRD;
      double E1=EXX-DAV,E2=EYY-DAV,E3=EZZ-DAV,E4=K_HALF*EXY,E5=K_HALF*EYZ,E6=K_HALF*EZX;
      double E42=E4*E4,E52=E5*E5,E62=E6*E6;
      double C=-K_HALF*(E1*E1+E2*E2+E3*E3)-E42-E52-E62;
      double EPST=std::sqrt(-C*K_THIRD);
      double EPSR1DAV=EPSR1F-DAV;
      if(EPST+EPST<EPSR1DAV) continue;
      double D=-E1*E2*E3+E1*E52+E2*E62+E3*E42-K_TWO*E4*E5*E6;
      double EPST2=EPST*EPST;
      double Y=(EPST2+C)*EPST+D;
      if(std::fabs(Y)>K_EM8){
        bool out=false;
        EPST=K_ONEP75*EPST;
        for(int it=0;it<4;it++){
          EPST2=EPST*EPST; Y=(EPST2+C)*EPST+D;
          double YP=K_THREE*EPST2+C;
          EPST=EPST-Y/YP;
          if(it<3 && EPST<EPSR1DAV){ out=true; break; }
        }
        if(out) continue;
      }
      EPST=EPST+DAV;
      EPSTT[i]=EPST;
      FAIL[i]=std::max(K_EM20,std::min(K_ONE,(m.epsr2-EPST)/(m.epsr2-m.epsr1)));
    }
  }
  auto VT=[&](int i,int k)->int&{ return vartmp[(size_t)(k-1)*nel+i]; };
  if(nrate==1){                                                            /* :398-414 */
    const int f=m.ifunc[0];
    for(int i=0;i<nel;i++){
      double dydx,y1; int ipos=VT(i,3);
      orc_vinter(o.TF,o.NPF[f],o.NPF[f+1]-o.NPF[f],ipos,pla[i],dydx,y1);
      VT(i,3)=ipos;
      double YFAC=m.yfac[0]*K_ONE;                  /* FACYLDI = 1 (no L_FAC_YLD) */
      double FACT=FAIL[i]*K_ONE*YFAC;
      H[i]=dydx*FACT;
      yld[i]=y1*FACT;                               /* FISOKIN == 0 */
    }
  } else {                                                                 /* :420-482 */
    for(int i=0;i<nel;i++){
      int JJ=1;
      for(int J=2;J<=nrate-1;J++) if(epsp[i]>=m.rate[J-1]) JJ=J;
      double RFAC;
      if(m.ismooth==2){
        double EPSP1=std::max(m.rate[JJ-1],K_EM20), EPSP2=m.rate[JJ];
        RFAC=std::log(std::max(epsp[i],K_EM20)/EPSP1)/std::log(EPSP2/EPSP1);
      } else {
        double EPSP1=m.rate[JJ-1], EPSP2=m.rate[JJ];
        RFAC=(epsp[i]-EPSP1)/(EPSP2-EPSP1);
      }
      const int J1=JJ,J2=JJ+1;
      int ipos1=VT(i,J1+2), ipos2=VT(i,J2+2);
      double YFAC1=m.yfac[J1-1]*K_ONE, YFAC2=m.yfac[J2-1]*K_ONE;
      const int f1=m.ifunc[J1-1], f2=m.ifunc[J2-1];
      double dydx1,y1,dydx2,y2;
      orc_vinter(o.TF,o.NPF[f1],o.NPF[f1+1]-o.NPF[f1],ipos1,pla[i],dydx1,y1);
      orc_vinter(o.TF,o.NPF[f2],o.NPF[f2+1]-o.NPF[f2],ipos2,pla[i],dydx2,y2);
      VT(i,J1+2)=ipos1; VT(i,J2+2)=ipos2;
      y1=y1*YFAC1; y2=y2*YFAC2;
      double FAC=RFAC, CC=FAIL[i]*K_ONE;
      yld[i]=(y1+FAC*(y2-y1))*CC;
      dydx1=dydx1*YFAC1; dydx2=dydx2*YFAC2;
      H[i]=(dydx1+FAC*(dydx2-dydx1))*CC;
    }
  }
  if(m.yldcheck==1) for(int i=0;i<nel;i++) yld[i]=std::max(yld[i],K_EM20);   /* :519-523 */
  /* projection :527-603 */
  for(int i=0;i<nel;i++){
    double VM=K_THREE*(K_HALF*(s1[i]*s1[i]+s2[i]*s2[i]+s3[i]*s3[i])+s4[i]*s4[i]+s5[i]*s5[i]+s6[i]*s6[i]);
    if(!(VM>yld[i]*yld[i])) continue;
    VM=std::sqrt(VM);
    R[i]=yld[i]/std::max(VM,K_EM20);
    if(ipla==0){
      s1[i]*=R[i]; s2[i]*=R[i]; s3[i]*=R[i]; s4[i]*=R[i]; s5[i]*=R[i]; s6[i]*=R[i];
      pla[i]=pla[i]+(K_ONE-R[i])*VM/std::max(G3+H[i],K_EM20);
      dpla1[i]=(K_ONE-R[i])*VM/std::max(G3+H[i],K_EM20);
    } else if(ipla==2){
      s1[i]*=R[i]; s2[i]*=R[i]; s3[i]*=R[i]; s4[i]*=R[i]; s5[i]*=R[i]; s6[i]*=R[i];
      pla[i]=pla[i]+(K_ONE-R[i])*VM/std::max(G3,K_EM20);
      dpla1[i]=(K_ONE-R[i])*VM/std::max(G3,K_EM20);
    } else {
      double DPLA=(K_ONE-R[i])*VM/std::max(G3+H[i],K_EM20);
      yld[i]=std::max(yld[i]+(K_ONE-FISOKIN)*DPLA*H[i],K_ZERO);
      R[i]=std::min(K_ONE,yld[i]/std::max(VM,K_EM20));
      s1[i]*=R[i]; s2[i]*=R[i]; s3[i]*=R[i]; s4[i]*=R[i]; s5[i]*=R[i]; s6[i]*=R[i];
      pla[i]=pla[i]+DPLA;
      dpla1[i]=DPLA;
    }
  }
  for(int i=0;i<nel;i++){ double P=BULK*amu[i]; s1[i]-=P; s2[i]-=P; s3[i]-=P; }   /* :1453-1460 (IEOS=0) */
  for(int i=0;i<nel;i++){                                                    /* :1507-1510 */
    if(off[i]<K_EM01) off[i]=K_ZERO;
    if(off[i]<K_ONE) off[i]=off[i]*K_FOUR_OVER_5;
  }
  if(m.ifail==1){                                                            /* :1546-1555 (IFAIL=1, no non-local) */
    for(int i=0;i<nel;i++) if(pla[i]>m.epsmax && off[i]==K_ONE) off[i]=K_FOUR_OVER_5;
  }
  else if(m.ifail==2){                                                       /* :1524-1533 */
    for(int i=0;i<nel;i++) if((pla[i]>m.epsmax || EPSTT[i]>m.epsf) && off[i]==K_ONE) off[i]=K_FOUR_OVER_5;
  }
}

/* MULAW  materials/mat_share/mulaw.F90 -- "user type" law driver for solids, MTN=36, isotropic global frame
 * (JCVT=0, ISORTH=0), no /VISC, no /FAIL, no non-local, ISVIS=0, JSPH=0.
 * :668-700 (old values, EP=D*OFF) ; :846-884 (DE, SO, strain rotation) ; :886-905 (ISTRAIN) ; :1049-1052
 * (rate filter) ; :1133-1166 (MSTRAIN_RATE IDEV=1, SIGEPS36) ; :2187-2219 (plastic work) ; :2876-2890
 * (SIG=S*OFF) ; :2895-2915 (SSP, MQVISCB) ; :3000-3016 (internal energy, QOLD). */
static void mulaw36(const Oracle& o,OrcSolidGroup& g,int nel,const int* ngl,
                    double* off,double* sig,double* eint,const double* rho,double* qold,double* defp,double* epsd,
                    const double* vol,double* stifn,double& dt2t,int& neltst,int& ityptst,const double* offg,
                    const double* amu,const double* vol_avg,double* ssp,const double* dvol,
                    const double* voln,const double* vd2,const double* deltax,double* vis,
                    const double* d1,const double* d2,const double* d3,const double* d4,const double* d5,const double* d6,
                    double* q,double* ssp_eq,
                    const double* sold1,const double* sold2,const double* sold3,
                    const double* sold4,const double* sold5,const double* sold6,
                    const double* wxx,const double* wyy,const double* wzz,const double* rhoref)
{
  const orgpu_law36& m=g.m36;
  const double DT1=o.DT1, facq0=K_ONE;
  double defp0[MVSIZ],ep1[MVSIZ],ep2[MVSIZ],ep3[MVSIZ],ep4[MVSIZ],ep5[MVSIZ],ep6[MVSIZ];
  double de1[MVSIZ],de2[MVSIZ],de3[MVSIZ],de4[MVSIZ],de5[MVSIZ],de6[MVSIZ];
  double so1[MVSIZ],so2[MVSIZ],so3[MVSIZ],so4[MVSIZ],so5[MVSIZ],so6[MVSIZ];
  double s1[MVSIZ],s2[MVSIZ],s3[MVSIZ],s4[MVSIZ],s5[MVSIZ],s6[MVSIZ];
  double dpla[MVSIZ],sigy[MVSIZ],viscmax[MVSIZ];
  auto S=[&](int i,int k)->double&{ return sig[k*nel+i]; };
  for(int i=0;i<nel;i++) defp0[i]=defp[i];
  for(int i=0;i<nel;i++){
    vis[i]=K_ZERO;
    ep1[i]=d1[i]*off[i]; ep2[i]=d2[i]*off[i]; ep3[i]=d3[i]*off[i];
    ep4[i]=d4[i]*off[i]; ep5[i]=d5[i]*off[i]; ep6[i]=d6[i]*off[i];
  }
  for(int i=0;i<nel;i++){
    de1[i]=ep1[i]*DT1; de2[i]=ep2[i]*DT1; de3[i]=ep3[i]*DT1; de4[i]=ep4[i]*DT1; de5[i]=ep5[i]*DT1; de6[i]=ep6[i]*DT1;
    so1[i]=S(i,0); so2[i]=S(i,1); so3[i]=S(i,2); so4[i]=S(i,3); so5[i]=S(i,4); so6[i]=S(i,5);
  }
  if(g.prop.istrain>0){                                   /* L_STRA>0: LBUF%STRA rotated (:860-884) then incremented (:886-892) */
    auto E=[&](int i,int k)->double&{ return g.stra[(size_t)(k-1)*nel+i]; };
    for(int i=0;i<nel;i++){
      double wxxf=wxx[i]*off[i], wyyf=wyy[i]*off[i], wzzf=wzz[i]*off[i];
      double q1=E(i,4)*wzzf, q2=E(i,6)*wyyf, q3=E(i,5)*wxxf;
      double ss1=E(i,1)-q1+q2, ss2=E(i,2)+q1-q3, ss3=E(i,3)-q2+q3;
      double ss4=E(i,4)+2.*wzzf*(E(i,1)-E(i,2))+wyyf*E(i,5)-wxxf*E(i,6);
      double ss5=E(i,5)+2.*wxxf*(E(i,2)-E(i,3))+wzzf*E(i,6)-wyyf*E(i,4);
      double ss6=E(i,6)+2.*wyyf*(E(i,3)-E(i,1))+wxxf*E(i,4)-wzzf*E(i,5);
      E(i,1)=ss1; E(i,2)=ss2; E(i,3)=ss3; E(i,4)=ss4; E(i,5)=ss5; E(i,6)=ss6;
    }
    for(int i=0;i<nel;i++){
      E(i,1)=E(i,1)+de1[i]; E(i,2)=E(i,2)+de2[i]; E(i,3)=E(i,3)+de3[i];
      E(i,4)=E(i,4)+de4[i]; E(i,5)=E(i,5)+de5[i]; E(i,6)=E(i,6)+de6[i];
    }
  }
  const int israte=m.israte; double asrate=K_ZERO;
  if(israte>0) asrate=std::min(K_ONE,m.asrate*DT1);
  mstrain_rate(nel,israte,asrate,epsd,1,ep1,ep2,ep3,ep4,ep5,ep6);
  sigeps36(o,m,nel,g.prop.ipla,de1,de2,de3,de4,de5,de6,so1,so2,so3,so4,so5,so6,s1,s2,s3,s4,s5,s6,
           ssp,viscmax,off,epsd,sigy,defp,dpla,amu,g.vartmp.data(),g.stra.data());
  /* plastic work (L_PLA>0, no L_SEQ) */
  for(int i=0;i<nel;i++){
    dpla[i]=defp[i]-defp0[i];
    double vm0=std::sqrt(K_HALF*((so1[i]-so2[i])*(so1[i]-so2[i])+(so2[i]-so3[i])*(so2[i]-so3[i])+(so3[i]-so1[i])*(so3[i]-so1[i]))
                         +K_THREE*(so4[i]*so4[i]+so5[i]*so5[i]+so6[i]*so6[i]));
    double vm =std::sqrt(K_HALF*((s1[i]-s2[i])*(s1[i]-s2[i])+(s2[i]-s3[i])*(s2[i]-s3[i])+(s3[i]-s1[i])*(s3[i]-s1[i]))
                         +K_THREE*(s4[i]*s4[i]+s5[i]*s5[i]+s6[i]*s6[i]));
    g.wpla[i]=g.wpla[i]+K_HALF*(vm0+vm)*dpla[i]*voln[i];
  }
  for(int i=0;i<nel;i++){
    S(i,0)=s1[i]*off[i]; S(i,1)=s2[i]*off[i]; S(i,2)=s3[i]*off[i];
    S(i,3)=s4[i]*off[i]; S(i,4)=s5[i]*off[i]; S(i,5)=s6[i]*off[i];
  }
  for(int i=0;i<nel;i++) if(ssp[i]==K_ZERO) ssp[i]=std::sqrt(m.bulk/m.rho0);
  mqviscb(o,g,nel,ngl,off,rho,ssp,stifn,dt2t,neltst,ityptst,offg,voln,vd2,deltax,vis,d1,d2,d3,
          q,ssp_eq,vol,rhoref,facq0);
  for(int i=0;i<nel;i++){
    double p2=-(sold1[i]+S(i,0)+sold2[i]+S(i,1)+sold3[i]+S(i,2))*K_THIRD;
    double e1=d1[i]*(sold1[i]+S(i,0)+p2+K_TWO*K_ZERO);
    double e2=d2[i]*(sold2[i]+S(i,1)+p2+K_TWO*K_ZERO);
    double e3=d3[i]*(sold3[i]+S(i,2)+p2+K_TWO*K_ZERO);
    double e4=d4[i]*(sold4[i]+S(i,3)+K_TWO*K_ZERO);
    double e5=d5[i]*(sold5[i]+S(i,4)+K_TWO*K_ZERO);
    double e6=d6[i]*(sold6[i]+S(i,5)+K_TWO*K_ZERO);
    double einc=off[i]*(vol_avg[i]*DT1*(e1+e2+e3+e4+e5+e6+K_ZERO)-dvol[i]*(q[i]+qold[i]+p2))*K_HALF;
    eint[i]=eint[i]+einc;
  }
  for(int i=0;i<nel;i++) qold[i]=q[i];
}

/* SFORC3 */
void orc_sforc3(Oracle& o, OrcSolidGroup& g, double& dt2t, int& neltst, int& ityptst)
{
  const int nel=g.nel, nft=g.nft;
  const int ismstr=g.prop.ismstr, jhbe=g.prop.jhbe, jcvt=g.prop.jcvt;
  const double DT1=o.DT1;
  static thread_local Vec X1,X2,X3,X4,X5,X6,X7,X8,Y1,Y2,Y3,Y4,Y5,Y6,Y7,Y8,Z1,Z2,Z3,Z4,Z5,Z6,Z7,Z8;
  static thread_local Vec XD[8],YD[8],ZD[8],VX[8],VY[8],VZ[8];
  int NC[8][MVSIZ], NGL[MVSIZ];
  double OFF[MVSIZ],RHOO[MVSIZ],VIS[MVSIZ],VD2[MVSIZ];
  double* OFFG=g.off.data();
  double* SAV=g.smstr.data();           /* SAV(I,k) -> SAV[(k-1)*nel+i] */
  auto sav=[&](int i,int k)->double&{ return SAV[(k-1)*nel+i]; };
  const int* ixs=&o.IXS[(size_t)11*nft];
  /* ---- SCOOR3  scoor3.F:107-386 (IRESP=0, ISORTH=0, ISROT=0, JLAG/=0) */
  for(int i=0;i<nel;i++){
    VIS[i]=K_ZERO; NGL[i]=ixs[11*i+10];
    for(int k=0;k<8;k++) NC[k][i]=ixs[11*i+1+k];
    RHOO[i]=g.rho[i];
  }
  double OFF_L=K_ZERO;
  for(int i=0;i<nel;i++) for(int k=0;k<8;k++){
    int n=NC[k][i]-1;
    XD[k][i]=o.X[3*n]; YD[k][i]=o.X[3*n+1]; ZD[k][i]=o.X[3*n+2];
  }
  /* ---- SRCOOR3 (srcoor3.F:143-723, JCVT /= 0: Belytschko's co-rotational frame; ISORTH=0, IRESP=0): the frame from the
   * iso-parametric axes of the CURRENT coordinates (SREPISO3 srepiso3.F:80-111, SORTHO3 sortho3.F:75-150), then coordinates --
   * or the saved small-strain reference, which already lives in that frame -- and velocities in it */
  static thread_local Vec R11,R12,R13,R21,R22,R23,R31,R32,R33;
  if(jcvt!=0){
    for(int i=0;i<nel;i++){
      const double X17=XD[6][i]-XD[0][i], X28=XD[7][i]-XD[1][i], X35=XD[4][i]-XD[2][i], X46=XD[5][i]-XD[3][i];
      const double Y17=YD[6][i]-YD[0][i], Y28=YD[7][i]-YD[1][i], Y35=YD[4][i]-YD[2][i], Y46=YD[5][i]-YD[3][i];
      const double Z17=ZD[6][i]-ZD[0][i], Z28=ZD[7][i]-ZD[1][i], Z35=ZD[4][i]-ZD[2][i], Z46=ZD[5][i]-ZD[3][i];
      const double A17=X17+X46, A28=X28+X35, B17=Y17+Y46, B28=Y28+Y35, C17=Z17+Z46, C28=Z28+Z35;
      const double RX=X17+X28-X35-X46, RY=Y17+Y28-Y35-Y46, RZ=Z17+Z28-Z35-Z46;
      const double SX=A17+A28, SY=B17+B28, SZ=C17+C28;
      const double TX=A17-A28, TY=B17-B28, TZ=C17-C28;
      double aa=std::sqrt(RX*RX+RY*RY+RZ*RZ); if(aa!=K_ZERO) aa=K_ONE/aa;
      double Ux=RX*aa, Uy=RY*aa, Uz=RZ*aa;
      aa=std::sqrt(SX*SX+SY*SY+SZ*SZ); if(aa!=K_ZERO) aa=K_ONE/aa;
      double Vx=SX*aa, Vy=SY*aa, Vz=SZ*aa;
      aa=std::sqrt(TX*TX+TY*TY+TZ*TZ); if(aa!=K_ZERO) aa=K_ONE/aa;
      double Wx=TX*aa, Wy=TY*aa, Wz=TZ*aa;
      for(int N=0;N<3;N++){                                      /* NITER = 3 */
        const double e1x=Vy*Wz-Vz*Wy+Ux, e1y=Vz*Wx-Vx*Wz+Uy, e1z=Vx*Wy-Vy*Wx+Uz;
        const double e2x=Wy*Uz-Wz*Uy+Vx, e2y=Wz*Ux-Wx*Uz+Vy, e2z=Wx*Uy-Wy*Ux+Vz;
        const double e3x=Uy*Vz-Uz*Vy+Wx, e3y=Uz*Vx-Ux*Vz+Wy, e3z=Ux*Vy-Uy*Vx+Wz;
        double bb=std::sqrt(e1x*e1x+e1y*e1y+e1z*e1z); if(bb!=K_ZERO) bb=K_ONE/bb;
        Ux=e1x*bb; Uy=e1y*bb; Uz=e1z*bb;
        bb=std::sqrt(e2x*e2x+e2y*e2y+e2z*e2z); if(bb!=K_ZERO) bb=K_ONE/bb;
        Vx=e2x*bb; Vy=e2y*bb; Vz=e2z*bb;
        bb=std::sqrt(e3x*e3x+e3y*e3y+e3z*e3z); if(bb!=K_ZERO) bb=K_ONE/bb;
        Wx=e3x*bb; Wy=e3y*bb; Wz=e3z*bb;
      }
      const double e1x=Ux, e1y=Uy, e1z=Uz;
      double e3x=e1y*Vz-e1z*Vy, e3y=e1z*Vx-e1x*Vz, e3z=e1x*Vy-e1y*Vx;
      aa=std::sqrt(e3x*e3x+e3y*e3y+e3z*e3z); if(aa!=K_ZERO) aa=K_ONE/aa;
      e3x=e3x*aa; e3y=e3y*aa; e3z=e3z*aa;
      const double e2x=e3y*e1z-e3z*e1y, e2y=e3z*e1x-e3x*e1z, e2z=e3x*e1y-e3y*e1x;
      /* SORTHO3(..., E1X=R11, E1Y=R12(!), ...) as SRCOOR3 passes them: R11,R12,R13 receive e1x,e2x,e3x; R21.. e1y,e2y,e3y; R31.. e1z,e2z,e3z */
      R11[i]=e1x; R12[i]=e2x; R13[i]=e3x; R21[i]=e1y; R22[i]=e2y; R23[i]=e3y; R31[i]=e1z; R32[i]=e2z; R33[i]=e3z;
    }
  }
  if(jcvt!=0 && ismstr<=4){                           /* srcoor3.F:328-420 */
    for(int i=0;i<nel;i++){
      if(std::fabs(OFFG[i])>K_ONE){
        for(int k=0;k<7;k++){ XD[k][i]=sav(i,3*k+1); YD[k][i]=sav(i,3*k+2); ZD[k][i]=sav(i,3*k+3); }
        XD[7][i]=K_ZERO; YD[7][i]=K_ZERO; ZD[7][i]=K_ZERO;
        OFF[i]=std::fabs(OFFG[i])-K_ONE;
        OFF_L=std::min(OFF_L,OFFG[i]);
      } else {
        for(int k=0;k<8;k++){
          const double XDL=R11[i]*XD[k][i]+R21[i]*YD[k][i]+R31[i]*ZD[k][i];
          const double YDL=R12[i]*XD[k][i]+R22[i]*YD[k][i]+R32[i]*ZD[k][i];
          const double ZDL=R13[i]*XD[k][i]+R23[i]*YD[k][i]+R33[i]*ZD[k][i];
          XD[k][i]=XDL; YD[k][i]=YDL; ZD[k][i]=ZDL;
        }
        OFF[i]=std::fabs(OFFG[i]);
        OFF_L=std::min(OFF_L,OFFG[i]);
      }
    }
  } else
  if(ismstr<=4){                                      /* :218-253 (JLAG>0) */
    for(int i=0;i<nel;i++){
      if(std::fabs(OFFG[i])>K_ONE){
        for(int k=0;k<7;k++){ XD[k][i]=sav(i,3*k+1); YD[k][i]=sav(i,3*k+2); ZD[k][i]=sav(i,3*k+3); }
        XD[7][i]=K_ZERO; YD[7][i]=K_ZERO; ZD[7][i]=K_ZERO;
        OFF[i]=std::fabs(OFFG[i])-K_ONE;
        OFF_L=std::min(OFF_L,OFFG[i]);
      } else {
        OFF[i]=std::fabs(OFFG[i]);
        OFF_L=std::min(OFF_L,OFFG[i]);
      }
    }
  } else {
    for(int i=0;i<nel;i++){ OFF[i]=std::fabs(OFFG[i]); OFF_L=std::min(OFF_L,OFFG[i]); }
  }
  Vec* Xs[8]={&X1,&X2,&X3,&X4,&X5,&X6,&X7,&X8};
  Vec* Ys[8]={&Y1,&Y2,&Y3,&Y4,&Y5,&Y6,&Y7,&Y8};
  Vec* Zs[8]={&Z1,&Z2,&Z3,&Z4,&Z5,&Z6,&Z7,&Z8};
  for(int i=0;i<nel;i++) for(int k=0;k<8;k++){ (*Xs[k])[i]=XD[k][i]; (*Ys[k])[i]=YD[k][i]; (*Zs[k])[i]=ZD[k][i]; }
  for(int i=0;i<nel;i++) for(int k=0;k<8;k++){
    int n=NC[k][i]-1;
    VX[k][i]=o.V[3*n]; VY[k][i]=o.V[3*n+1]; VZ[k][i]=o.V[3*n+2];
  }
  static thread_local Vec VGX[8],VGY[8],VGZ[8];          /* velocities in the global frame: what SRBILAN books (srcoor3.F:231-244) */
  if(jcvt!=0 && o.ipri) for(int i=0;i<nel;i++) for(int k=0;k<8;k++){ VGX[k][i]=VX[k][i]; VGY[k][i]=VY[k][i]; VGZ[k][i]=VZ[k][i]; }
  if(jcvt!=0){
    /* srcoor3.F:672-681: CALL SRROTA3(R11,R12,R13,R21,R22,R23,R31,R32,R33, VX.., VY.., VZ..) -- SRROTA3's dummies are
     * (R11,R21,R31,R12,R22,R32,R13,R23,R33), so its X = R11*VX + R21*VY + R31*VZ reads R11*VX + R12*VY + R13*VZ here... with
     * the actual arguments in that order the velocity gets  V_loc = t(R) V:  x' = R11 vx + R21 vy + R31 vz */
    for(int i=0;i<nel;i++) for(int k=0;k<8;k++){
      const double X=R11[i]*VX[k][i]+R21[i]*VY[k][i]+R31[i]*VZ[k][i];
      const double Y=R12[i]*VX[k][i]+R22[i]*VY[k][i]+R32[i]*VZ[k][i];
      const double Z=R13[i]*VX[k][i]+R23[i]*VY[k][i]+R33[i]*VZ[k][i];
      VX[k][i]=X; VY[k][i]=Y; VZ[k][i]=Z;
    }
  }
  if(OFF_L<K_ZERO){
    for(int i=0;i<nel;i++) if(OFFG[i]<K_ZERO) for(int k=0;k<8;k++){ VX[k][i]=K_ZERO; VY[k][i]=K_ZERO; VZ[k][i]=K_ZERO; }
  }
  for(int i=0;i<nel;i++) VD2[i]=K_ZERO;               /* :358-386 */

  /* ---- SDERI3  sderi3.F:105-336 */
  double VOLN[MVSIZ],VOLDP[MVSIZ];
  double JAC1[MVSIZ],JAC2[MVSIZ],JAC3[MVSIZ],JAC4[MVSIZ],JAC5[MVSIZ],JAC6[MVSIZ],JAC7[MVSIZ],JAC8[MVSIZ],JAC9[MVSIZ];
  double C5968[MVSIZ],C6749[MVSIZ],C4857[MVSIZ];
  double PX1[MVSIZ],PX2[MVSIZ],PX3[MVSIZ],PX4[MVSIZ],PY1[MVSIZ],PY2[MVSIZ],PY3[MVSIZ],PY4[MVSIZ],PZ1[MVSIZ],PZ2[MVSIZ],PZ3[MVSIZ],PZ4[MVSIZ];
  double PX1H1[MVSIZ],PX2H1[MVSIZ],PX3H1[MVSIZ],PX4H1[MVSIZ],PX1H2[MVSIZ],PX2H2[MVSIZ],PX3H2[MVSIZ],PX4H2[MVSIZ],PX1H3[MVSIZ],PX2H3[MVSIZ],PX3H3[MVSIZ],PX4H3[MVSIZ];
  auto do_jac=[&](int i){
    double x[8],y[8],z[8]; for(int k=0;k<8;k++){x[k]=XD[k][i];y[k]=YD[k][i];z[k]=ZD[k][i];}
    JacOut r=sderi_jac(x,y,z);
    JAC1[i]=r.J1;JAC2[i]=r.J2;JAC3[i]=r.J3;JAC4[i]=r.J4;JAC5[i]=r.J5;JAC6[i]=r.J6;JAC7[i]=r.J7;JAC8[i]=r.J8;JAC9[i]=r.J9;
    C5968[i]=r.c5968;C6749[i]=r.c6749;C4857[i]=r.c4857; VOLDP[i]=r.voldp; VOLN[i]=r.voldp;
  };
  for(int i=0;i<nel;i++) do_jac(i);
  /* SCHKJABT3  solid/solide4/schkjabt3.F:75-138 (JLAG/=0, INCONV=1, INEG_V default: switch to small strain) */
  {
    int nnega=0, index[MVSIZ]; int icor=0;
    for(int i=0;i<nel;i++){
      if(OFF[i]==K_ZERO) VOLN[i]=K_ONE;
      else if(OFFG[i]>K_ONE) {}
      else if(VOLN[i]<=K_ZERO) icor=1;
    }
    if(icor>0){
      for(int i=0;i<nel;i++)
        if(VOLN[i]<=K_ZERO && OFFG[i]<=K_ONE && OFFG[i]!=K_ZERO){ index[nnega++]=i; }
    }
    if(nnega>0){                                     /* sderi3.F:157-263 */
      for(int j=0;j<nnega;j++){
        int i=index[j];
        for(int k=0;k<7;k++){ XD[k][i]=sav(i,3*k+1); YD[k][i]=sav(i,3*k+2); ZD[k][i]=sav(i,3*k+3); }
        XD[7][i]=K_ZERO; YD[7][i]=K_ZERO; ZD[7][i]=K_ZERO;
      }
      for(int j=0;j<nnega;j++){ int i=index[j]; do_jac(i); OFFG[i]=K_TWO; }
    }
  }
  for(int i=0;i<nel;i++){                            /* :266-301 */
    double DETT=K_ONE_OVER_64/VOLN[i];
    double JACI1=DETT*C5968[i], JACI4=DETT*C6749[i], JACI7=DETT*C4857[i];
    double JACI2=DETT*(JAC3[i]*JAC8[i]-JAC2[i]*JAC9[i]);
    double JACI5=DETT*(JAC1[i]*JAC9[i]-JAC3[i]*JAC7[i]);
    double JACI8=DETT*(JAC2[i]*JAC7[i]-JAC1[i]*JAC8[i]);
    double JACI3=DETT*(JAC2[i]*JAC6[i]-JAC3[i]*JAC5[i]);
    double JACI6=DETT*(JAC3[i]*JAC4[i]-JAC1[i]*JAC6[i]);
    double JACI9=DETT*(JAC1[i]*JAC5[i]-JAC2[i]*JAC4[i]);
    double JAC12=JACI1+JACI2, JAC45=JACI4+JACI5, JAC78=JACI7+JACI8;
    PX1[i]=-JAC12-JACI3; PY1[i]=-JAC45-JACI6; PZ1[i]=-JAC78-JACI9;
    PX2[i]=-JAC12+JACI3; PY2[i]=-JAC45+JACI6; PZ2[i]=-JAC78+JACI9;
    JAC12=JACI1-JACI2; JAC45=JACI4-JACI5; JAC78=JACI7-JACI8;
    PX3[i]=JAC12+JACI3; PY3[i]=JAC45+JACI6; PZ3[i]=JAC78+JACI9;
    PX4[i]=JAC12-JACI3; PY4[i]=JAC45-JACI6; PZ4[i]=JAC78-JACI9;
  }
  if(jhbe!=0){                                       /* :303-336 */
    for(int i=0;i<nel;i++){
      double HX=(XD[0][i]-XD[1][i]+XD[2][i]-XD[3][i]+XD[4][i]-XD[5][i]+XD[6][i]-XD[7][i]);
      double HY=(YD[0][i]-YD[1][i]+YD[2][i]-YD[3][i]+YD[4][i]-YD[5][i]+YD[6][i]-YD[7][i]);
      double HZ=(ZD[0][i]-ZD[1][i]+ZD[2][i]-ZD[3][i]+ZD[4][i]-ZD[5][i]+ZD[6][i]-ZD[7][i]);
      PX1H1[i]=PX1[i]*HX+PY1[i]*HY+PZ1[i]*HZ; PX2H1[i]=PX2[i]*HX+PY2[i]*HY+PZ2[i]*HZ;
      PX3H1[i]=PX3[i]*HX+PY3[i]*HY+PZ3[i]*HZ; PX4H1[i]=PX4[i]*HX+PY4[i]*HY+PZ4[i]*HZ;
    }
    for(int i=0;i<nel;i++){
      double HX=(XD[0][i]+XD[1][i]-XD[2][i]-XD[3][i]-XD[4][i]-XD[5][i]+XD[6][i]+XD[7][i]);
      double HY=(YD[0][i]+YD[1][i]-YD[2][i]-YD[3][i]-YD[4][i]-YD[5][i]+YD[6][i]+YD[7][i]);
      double HZ=(ZD[0][i]+ZD[1][i]-ZD[2][i]-ZD[3][i]-ZD[4][i]-ZD[5][i]+ZD[6][i]+ZD[7][i]);
      PX1H2[i]=PX1[i]*HX+PY1[i]*HY+PZ1[i]*HZ; PX2H2[i]=PX2[i]*HX+PY2[i]*HY+PZ2[i]*HZ;
      PX3H2[i]=PX3[i]*HX+PY3[i]*HY+PZ3[i]*HZ; PX4H2[i]=PX4[i]*HX+PY4[i]*HY+PZ4[i]*HZ;
    }
    for(int i=0;i<nel;i++){
      double HX=(XD[0][i]-XD[1][i]-XD[2][i]+XD[3][i]-XD[4][i]+XD[5][i]+XD[6][i]-XD[7][i]);
      double HY=(YD[0][i]-YD[1][i]-YD[2][i]+YD[3][i]-YD[4][i]+YD[5][i]+YD[6][i]-YD[7][i]);
      double HZ=(ZD[0][i]-ZD[1][i]-ZD[2][i]+ZD[3][i]-ZD[4][i]+ZD[5][i]+ZD[6][i]-ZD[7][i]);
      PX1H3[i]=PX1[i]*HX+PY1[i]*HY+PZ1[i]*HZ; PX2H3[i]=PX2[i]*HX+PY2[i]*HY+PZ2[i]*HZ;
      PX3H3[i]=PX3[i]*HX+PY3[i]*HY+PZ3[i]*HZ; PX4H3[i]=PX4[i]*HX+PY4[i]*HY+PZ4[i]*HZ;
    }
  }
  /* ---- SDLEN3  sdlen3.F:89-168 (Lagrangian branch; IDTS6=0; MTN/=5,41) -- uses X1..Z8 copies */
  double DELTAX[MVSIZ];
  {
    static thread_local double AREA[MVSIZ][6]; Vec AREAM; double XIOFF[MVSIZ];
    for(int i=0;i<nel;i++){ XIOFF[i]=K_ONE; AREAM[i]=K_EM20; }
    slen(X1,X2,X3,X4,Y1,Y2,Y3,Y4,Z1,Z2,Z3,Z4,0,AREA,AREAM,nel);
    slen(X5,X6,X7,X8,Y5,Y6,Y7,Y8,Z5,Z6,Z7,Z8,1,AREA,AREAM,nel);
    slen(X1,X2,X6,X5,Y1,Y2,Y6,Y5,Z1,Z2,Z6,Z5,2,AREA,AREAM,nel);
    slen(X2,X3,X7,X6,Y2,Y3,Y7,Y6,Z2,Z3,Z7,Z6,3,AREA,AREAM,nel);
    slen(X3,X4,X8,X7,Y3,Y4,Y8,Y7,Z3,Z4,Z8,Z7,4,AREA,AREAM,nel);
    slen(X4,X1,X5,X8,Y4,Y1,Y5,Y8,Z4,Z1,Z5,Z8,5,AREA,AREAM,nel);
    for(int i=0;i<nel;i++) DELTAX[i]=K_FOUR*VOLN[i]*XIOFF[i]/std::sqrt(AREAM[i]);
  }
  /* ---- SDEFO3  sdefo3.F:115-271 (INTEG8*JEUL=0, ISCAU=0, JCVT=0, ISROT=0) */
  double DXX[MVSIZ],DYY[MVSIZ],DZZ[MVSIZ],DXY[MVSIZ],DXZ[MVSIZ],DYX[MVSIZ],DYZ[MVSIZ],DZX[MVSIZ],DZY[MVSIZ];
  double D4[MVSIZ],D5[MVSIZ],D6[MVSIZ],WXX[MVSIZ],WYY[MVSIZ],WZZ[MVSIZ];
  for(int i=0;i<nel;i++){
    double VX17=VX[0][i]-VX[6][i], VX28=VX[1][i]-VX[7][i], VX35=VX[2][i]-VX[4][i], VX46=VX[3][i]-VX[5][i];
    double VY17=VY[0][i]-VY[6][i], VY28=VY[1][i]-VY[7][i], VY35=VY[2][i]-VY[4][i], VY46=VY[3][i]-VY[5][i];
    double VZ17=VZ[0][i]-VZ[6][i], VZ28=VZ[1][i]-VZ[7][i], VZ35=VZ[2][i]-VZ[4][i], VZ46=VZ[3][i]-VZ[5][i];
    DXX[i]=PX1[i]*VX17+PX2[i]*VX28+PX3[i]*VX35+PX4[i]*VX46;
    DYY[i]=PY1[i]*VY17+PY2[i]*VY28+PY3[i]*VY35+PY4[i]*VY46;
    DZZ[i]=PZ1[i]*VZ17+PZ2[i]*VZ28+PZ3[i]*VZ35+PZ4[i]*VZ46;
    DXY[i]=PY1[i]*VX17+PY2[i]*VX28+PY3[i]*VX35+PY4[i]*VX46;
    DXZ[i]=PZ1[i]*VX17+PZ2[i]*VX28+PZ3[i]*VX35+PZ4[i]*VX46;
    DYX[i]=PX1[i]*VY17+PX2[i]*VY28+PX3[i]*VY35+PX4[i]*VY46;
    DYZ[i]=PZ1[i]*VY17+PZ2[i]*VY28+PZ3[i]*VY35+PZ4[i]*VY46;
    DZX[i]=PX1[i]*VZ17+PX2[i]*VZ28+PX3[i]*VZ35+PX4[i]*VZ46;
    DZY[i]=PY1[i]*VZ17+PY2[i]*VZ28+PY3[i]*VZ35+PY4[i]*VZ46;
  }
  const double DT1D2=K_HALF*DT1;
  if(jcvt!=0){                                       /* sdefo3.F:158-220 (ISMSTR /= 11, IMPL_S = 0): no spin in the co-rotating frame, second-order strain rate */
    for(int i=0;i<nel;i++){
      WXX[i]=K_ZERO; WYY[i]=K_ZERO; WZZ[i]=K_ZERO;
      double EXX=DXX[i],EYY=DYY[i],EZZ=DZZ[i],EXY=DXY[i],EYX=DYX[i],EXZ=DXZ[i],EZX=DZX[i],EYZ=DYZ[i],EZY=DZY[i];
      DXX[i]=DXX[i]-DT1D2*(EXX*EXX+EYX*EYX+EZX*EZX);
      DYY[i]=DYY[i]-DT1D2*(EYY*EYY+EZY*EZY+EXY*EXY);
      DZZ[i]=DZZ[i]-DT1D2*(EZZ*EZZ+EXZ*EXZ+EYZ*EYZ);
      double AAA=DT1D2*(EXX*EXY+EYX*EYY+EZX*EZY);
      DXY[i]=DXY[i]-AAA; DYX[i]=DYX[i]-AAA; D4[i]=DXY[i]+DYX[i];
      AAA=DT1D2*(EYY*EYZ+EZY*EZZ+EXY*EXZ);
      DYZ[i]=DYZ[i]-AAA; DZY[i]=DZY[i]-AAA; D5[i]=DYZ[i]+DZY[i];
      AAA=DT1D2*(EZZ*EZX+EXZ*EXX+EYZ*EYX);
      DXZ[i]=DXZ[i]-AAA; DZX[i]=DZX[i]-AAA; D6[i]=DXZ[i]+DZX[i];
    }
  } else
  if(jhbe>=2){                                       /* :222-256 */
    for(int i=0;i<nel;i++){
      double EXX=DXX[i],EYY=DYY[i],EZZ=DZZ[i],EXY=DXY[i],EYX=DYX[i],EXZ=DXZ[i],EZX=DZX[i],EYZ=DYZ[i],EZY=DZY[i];
      DXX[i]=DXX[i]-DT1D2*(EXX*EXX+EYX*EYX+EZX*EZX);
      DYY[i]=DYY[i]-DT1D2*(EYY*EYY+EZY*EZY+EXY*EXY);
      DZZ[i]=DZZ[i]-DT1D2*(EZZ*EZZ+EXZ*EXZ+EYZ*EYZ);
      double AAA=DT1D2*(EXX*EXY+EYX*EYY+EZX*EZY);
      DXY[i]=DXY[i]-AAA; DYX[i]=DYX[i]-AAA; D4[i]=DXY[i]+DYX[i];
      AAA=DT1D2*(EYY*EYZ+EZY*EZZ+EXY*EXZ);
      DYZ[i]=DYZ[i]-AAA; DZY[i]=DZY[i]-AAA; D5[i]=DYZ[i]+DZY[i];
      AAA=DT1D2*(EZZ*EZX+EXZ*EXX+EYZ*EYX);
      DXZ[i]=DXZ[i]-AAA; DZX[i]=DZX[i]-AAA; D6[i]=DXZ[i]+DZX[i];
      double PXX2=PX1[i]*PX1[i]+PX2[i]*PX2[i]+PX3[i]*PX3[i]+PX4[i]*PX4[i];
      double PYY2=PY1[i]*PY1[i]+PY2[i]*PY2[i]+PY3[i]*PY3[i]+PY4[i]*PY4[i];
      double PZZ2=PZ1[i]*PZ1[i]+PZ2[i]*PZ2[i]+PZ3[i]*PZ3[i]+PZ4[i]*PZ4[i];
      WZZ[i]=DT1*(PYY2*DYX[i]-PXX2*DXY[i])/(PXX2+PYY2);
      WXX[i]=DT1*(PZZ2*DZY[i]-PYY2*DYZ[i])/(PYY2+PZZ2);
      WYY[i]=DT1*(PXX2*DXZ[i]-PZZ2*DZX[i])/(PZZ2+PXX2);
    }
  } else {                                           /* :258-269 */
    for(int i=0;i<nel;i++){
      D4[i]=DXY[i]+DYX[i]; D5[i]=DYZ[i]+DZY[i]; D6[i]=DXZ[i]+DZX[i];
      WZZ[i]=DT1D2*(DYX[i]-DXY[i]);
      WYY[i]=DT1D2*(DXZ[i]-DZX[i]);
      WXX[i]=DT1D2*(DZY[i]-DYZ[i]);
    }
  }
  /* sforc3.F:790 */
  double DIVDE[MVSIZ]; for(int i=0;i<nel;i++) DIVDE[i]=DT1*(DXX[i]+DYY[i]+DZZ[i]);
  /* ---- SRHO3  srho3.F:110-239 (JLAG, IMPL_S=0, MTN/=115, IRESP=0) */
  double DVOL[MVSIZ]; double* VOLO=g.vol.data(); double* RHON=g.rho.data(); double* EINT=g.eint.data();
  {
    double RHON_OLD[MVSIZ]; for(int i=0;i<nel;i++) RHON_OLD[i]=RHON[i];
    const double RHO0=g.mat.rho0;
    if(o.TT==K_ZERO && ismstr==1){ for(int i=0;i<nel;i++) if(OFFG[i]>K_ONE) VOLO[i]=VOLN[i]; }
    for(int i=0;i<nel;i++){
      if(OFFG[i]==K_ZERO && VOLN[i]==K_ONE) VOLN[i]=VOLO[i];
      DVOL[i]=VOLN[i]-(RHO0/RHON[i])*VOLO[i];
      RHON[i]=RHO0*(VOLO[i]/VOLN[i]);
      EINT[i]=EINT[i]*VOLO[i];
    }
    if(ismstr<=4){
      for(int i=0;i<nel;i++) if(OFFG[i]>K_ONE){
        double DVDP=DIVDE[i]; double RHOREF=RHON[i];
        RHON[i]=RHON_OLD[i]-RHOREF*DVDP;
        RHON[i]=std::max(RHON[i],K_EM30);
        DVOL[i]=VOLN[i]*DVDP;
      }
    }
  }
  /* ---- SROTA3  srota3.F:72-94 */
  double S1[MVSIZ],S2[MVSIZ],S3[MVSIZ],S4[MVSIZ],S5[MVSIZ],S6[MVSIZ];
  double* SIG=g.sig.data(); auto SG=[&](int i,int k)->double&{ return SIG[k*nel+i]; };
  for(int i=0;i<nel;i++){ S1[i]=SG(i,0);S2[i]=SG(i,1);S3[i]=SG(i,2);S4[i]=SG(i,3);S5[i]=SG(i,4);S6[i]=SG(i,5); }
  if(jcvt==0)                                        /* JCVT /= 0: SRMALLA3 (srmall3.F) only copies the old stress, it lives in the co-rotating frame */
  for(int i=0;i<nel;i++){
    double Q1=K_TWO*S4[i]*WZZ[i], Q2=K_TWO*S6[i]*WYY[i], Q3=K_TWO*S5[i]*WXX[i];
    SG(i,0)=S1[i]-Q1+Q2;
    SG(i,1)=S2[i]+Q1-Q3;
    SG(i,2)=S3[i]-Q2+Q3;
    SG(i,3)=S4[i]+WZZ[i]*(S1[i]-S2[i])+WYY[i]*S5[i]-WXX[i]*S6[i];
    SG(i,4)=S5[i]+WXX[i]*(S2[i]-S3[i])+WZZ[i]*S6[i]-WYY[i]*S4[i];
    SG(i,5)=S6[i]+WYY[i]*(S3[i]-S1[i])+WXX[i]*S4[i]-WZZ[i]*S5[i];
  }
  /* ---- SMALLA3  smalla3.F:118-173 (ISMSTR<=4, JLAG>0; not called for JCVT /= 0, sforc3.F:807-822) */
  if(ismstr<=4 && jcvt==0){
    for(int i=0;i<nel;i++) if(OFFG[i]>K_ONE){
      for(int k=0;k<7;k++){
        double X=sav(i,3*k+1),Y=sav(i,3*k+2),Z=sav(i,3*k+3);
        sav(i,3*k+1)=X-Y*WZZ[i]+Z*WYY[i];
        sav(i,3*k+2)=Y-Z*WXX[i]+X*WZZ[i];
        sav(i,3*k+3)=Z-X*WYY[i]+Y*WXX[i];
      }
    }
  }
  /* ---- S8SAV3  s8sav3.F:66-90 (ISMSTR<=3 or ISMSTR==4 with JLAG) */
  if(ismstr<=4){
    for(int i=0;i<nel;i++) if(std::fabs(OFFG[i])<=K_ONE){
      for(int k=0;k<7;k++){
        sav(i,3*k+1)=XD[k][i]-XD[7][i];
        sav(i,3*k+2)=YD[k][i]-YD[7][i];
        sav(i,3*k+3)=ZD[k][i]-ZD[7][i];
      }
    }
  }
  /* ---- MMAIN  materials/mat_share/mmain.F90:640-925 (pre-law section, LAW2 call) */
  double STI[MVSIZ],QVIS[MVSIZ],CXX[MVSIZ],SSP_EQ[MVSIZ];
  {
    double QOLD[MVSIZ],RHO0[MVSIZ],VOL_AVG[MVSIZ],AMU[MVSIZ],RHOREF[MVSIZ],TSTAR[MVSIZ];
    for(int i=0;i<nel;i++) QOLD[i]=g.qvis[i];                 /* :597 */
    for(int i=0;i<nel;i++){                                   /* :668-746 */
      RHO0[i]=g.mat.rho0;
      VOL_AVG[i]=VOLN[i]-K_HALF*DVOL[i];
      AMU[i]=g.rho[i]/RHO0[i]-K_ONE;                          /* mtn==2 branch :693 */
    }
    if(ismstr==1){ for(int i=0;i<nel;i++) RHOREF[i]=RHO0[i]; }
    else if(ismstr==2){
      for(int i=0;i<nel;i++){
        if(std::fabs(OFFG[i])<=K_ONE) RHOREF[i]=g.rho[i];
        else RHOREF[i]=RHO0[i]*g.vol[i]/std::max(K_EM20,VOLN[i]);
      }
    } else { for(int i=0;i<nel;i++) RHOREF[i]=g.rho[i]; }
    /* tstar :793-800 */
    if(g.mat.has_temp){
      for(int i=0;i<nel;i++) TSTAR[i]=std::max(K_ZERO,(g.temp[i]-g.mat.tref)/std::max((g.mat.tmelt-g.mat.tref),K_EM20));
    } else { for(int i=0;i<nel;i++) TSTAR[i]=K_ZERO; }
    double vecnul[MVSIZ]; for(int i=0;i<nel;i++) vecnul[i]=K_ZERO;
    double DPLA_F[MVSIZ],EPSP_F[MVSIZ]; for(int i=0;i<nel;i++){ DPLA_F[i]=K_ZERO; EPSP_F[i]=K_ZERO; }
    double* el_temp = g.mat.has_temp ? g.temp.data() : vecnul;
    if(g.law==36){
      /* mmain.F90:1899-1960 MULAW, then :1996-2004 energy -> energy density */
      mulaw36(o,g,nel,NGL,OFF,SIG,EINT,RHON,g.qvis.data(),g.pla.data(),g.epsd.data(),g.vol.data(),STI,
              dt2t,neltst,ityptst,OFFG,AMU,VOL_AVG,CXX,DVOL,VOLN,VD2,DELTAX,VIS,
              DXX,DYY,DZZ,D4,D5,D6,QVIS,SSP_EQ,S1,S2,S3,S4,S5,S6,WXX,WYY,WZZ,RHOREF);
      for(int i=0;i<nel;i++){
        if(g.vol[i]>K_ZERO) EINT[i]=EINT[i]/std::max(g.vol[i],K_EM20);
        else EINT[i]=K_ZERO;
      }
    } else
    m2law(o,g,nel,NGL,OFF,SIG,EINT,RHON,g.qvis.data(),g.pla.data(),g.epsd.data(),g.vol.data(),STI,
          dt2t,neltst,ityptst,OFFG,AMU,VOL_AVG,CXX,DVOL,VOLN,VD2,DELTAX,VIS,
          DXX,DYY,DZZ,D4,D5,D6,QVIS,SSP_EQ,S1,S2,S3,S4,S5,S6,TSTAR,el_temp,g.dmg.data(),RHOREF,
          g.mat.fisokin>K_ZERO ? g.sigb.data() : nullptr, DPLA_F, EPSP_F);
    /* m2law stored QNEW into QVIS and then QOLD(=lbuf%qvis) = QNEW */
    /* mmain.F90 tail: l_temp>0 entropy heating of the artificial viscosity */
    if(g.mat.has_temp){
      double cv=g.mat.rhocp/g.mat.rho0;                       /* eos%cv==0 -> cp = rhocp/rho0 */
      if(cv>K_ZERO){
        for(int i=0;i<nel;i++) if(OFF[i]==K_ONE){
          double mcv=g.rho[i]*VOLN[i]*cv;
          double qheat=-K_HALF*(QOLD[i]+g.qvis[i])*DVOL[i];
          double dtemp=qheat/mcv;
          g.temp[i]=g.temp[i]+dtemp;
          g.temp[i]=std::max(K_ZERO,g.temp[i]);
        }
      }
    }
    /* failure models of MMAIN (mmain.F90:2250-2262 NFAIL>0, MTN<28; :2288-2300; :2410-2416 FAIL_JOHNSON, fail_johnson.F:95-141 with
     * Ifail_so = 1; the stress is left alone, so the energy correction of :2788-2802 is a multiply and a divide by the volume) */
    if(g.law==2 && g.fail.irupt==1){
      const orgpu_fail& f=g.fail;
      for(int i=0;i<nel;i++){
        if(OFF[i]<(double)0.1f) OFF[i]=K_ZERO;                  /* REAL*4 literals 0.1, 0.8 */
        if(OFF[i]<K_ONE) OFF[i]=OFF[i]*(double)0.8f;
      }
      for(int i=0;i<nel;i++){
        if(OFF[i]==K_ONE){
          if(DPLA_F[i]!=K_ZERO){
            const double* S_=SIG;
            const double sxx_=S_[i], syy_=S_[nel+i], szz_=S_[2*nel+i], sxy_=S_[3*nel+i], syz_=S_[4*nel+i], szx_=S_[5*nel+i];
            const double P=K_THIRD*(sxx_+syy_+szz_);
            const double SXX=sxx_-P, SYY=syy_-P, SZZ=szz_-P;
            double SVM=K_HALF*(SXX*SXX+SYY*SYY+SZZ*SZZ)+sxy_*sxy_+szx_*szx_+syz_*syz_;
            SVM=std::sqrt(K_THREE*SVM);
            double EPSF=f.d3*P/std::max(K_EM20,SVM);
            EPSF=f.d1+f.d2*std::exp(EPSF);
            if(f.d4!=K_ZERO) EPSF=EPSF*(K_ONE+f.d4*std::log(std::max(K_ONE,EPSP_F[i]/f.epsp0)));
            EPSF=std::max(EPSF,f.epsf_min);
            if(EPSF>K_ZERO) g.dfmax[i]=g.dfmax[i]+DPLA_F[i]/EPSF;
            g.dfmax[i]=std::min(K_ONE,g.dfmax[i]);
          }
          if(g.dfmax[i]>=K_ONE && OFF[i]==K_ONE) OFF[i]=K_FOUR_OVER_5;
        }
      }
      for(int i=0;i<nel;i++){ const double e=EINT[i]*g.vol[i]; EINT[i]=e/std::max(g.vol[i],K_EM20); }
    }
  }
  /* ---- SMALLB3  smallb3.F:65-81 */
  if(ismstr==1||ismstr==3){ for(int i=0;i<nel;i++) if(OFFG[i]>K_ZERO) OFFG[i]=K_TWO; }
  for(int i=0;i<nel;i++){
    if(OFF[i]<K_ONE){
      if(OFF[i]==K_ZERO) OFFG[i]=K_ZERO;
      else if(OFFG[i]>K_ONE) OFFG[i]=K_ONE+OFF[i];
      else OFFG[i]=OFF[i];
    }
  }
  if(o.ipri){                                          /* SBILAN sforc3.F:1436-1458 */
    for(int i=0;i<nel;i++){
      double vx[8],vy[8],vz[8]; for(int k=0;k<8;k++){ vx[k]=VX[k][i]; vy[k]=VY[k][i]; vz[k]=VZ[k][i]; }
      if(jcvt!=0) for(int k=0;k<8;k++){ vx[k]=VGX[k][i]; vy[k]=VGY[k][i]; vz[k]=VGZ[k][i]; }   /* SRBILAN (sforc3.F:1459-1470): the same sums, from the global velocities */
      orc_bilan_solid(o,nft+i,vx,vy,vz,g.eint[i],g.vol[i],g.rho[i],VOLN[i],OFFG[i]);
    }
  }
  /* ---- SHVIS3  shvis3.F:164-412 (INVSTR>=35, FLUID=0) */
  double F1[8][MVSIZ],F2[8][MVSIZ],F3[8][MVSIZ];   /* F1[k]=F1(k+1): x-force on node k+1 */
  {
    double CAQ[MVSIZ],FCL[MVSIZ],FCQ[MVSIZ];
    for(int i=0;i<nel;i++) CAQ[i]=K_FOURTH*OFF[i]*g.prop.hcoef;
    const double* RHO=g.rho.data();
    if(ismstr==1){
      for(int i=0;i<nel;i++){ FCL[i]=CAQ[i]*g.mat.rho0*std::pow(VOLN[i],K_TWO_THIRD); FCQ[i]=FCL[i]*CAQ[i]*K_HUNDRED; FCL[i]=FCL[i]*CXX[i]; }
    } else if(ismstr==2){
      for(int i=0;i<nel;i++){
        if(OFFG[i]>K_ONE){
          double AA=g.mat.rho0*g.vol[i]/std::max(K_EM20,VOLN[i]);
          FCL[i]=CAQ[i]*AA*std::pow(VOLN[i],K_TWO_THIRD);
        } else FCL[i]=CAQ[i]*RHO[i]*std::pow(VOLN[i],K_TWO_THIRD);
        FCQ[i]=FCL[i]*CAQ[i]*K_HUNDRED; FCL[i]=FCL[i]*CXX[i];
      }
    } else {
      for(int i=0;i<nel;i++){ FCL[i]=CAQ[i]*RHO[i]*std::pow(VOLN[i],K_TWO_THIRD); FCQ[i]=FCL[i]*CAQ[i]*K_HUNDRED; FCL[i]=FCL[i]*CXX[i]; }
    }
    double HGX[4][MVSIZ],HGY[4][MVSIZ],HGZ[4][MVSIZ];
    double G_[8][3][MVSIZ];
    if(jhbe==0){
      for(int i=0;i<nel;i++){
        auto hg=[&](Vec* V,double (*H)[MVSIZ]){
          double V3478=V[2][i]-V[3][i]-V[6][i]+V[7][i];
          double V2358=V[1][i]-V[2][i]-V[4][i]+V[7][i];
          double V1467=V[0][i]-V[3][i]-V[5][i]+V[6][i];
          double V1256=V[0][i]-V[1][i]-V[4][i]+V[5][i];
          H[0][i]=V1467-V2358; H[1][i]=V1467+V2358; H[2][i]=V1256-V3478; H[3][i]=V1256+V3478;
        };
        hg(VX,HGX); hg(VY,HGY); hg(VZ,HGZ);
      }
    } else {
      for(int i=0;i<nel;i++){
        G_[0][0][i]= K_ONE-PX1H1[i]; G_[1][0][i]=-K_ONE-PX2H1[i]; G_[2][0][i]= K_ONE-PX3H1[i]; G_[3][0][i]=-K_ONE-PX4H1[i];
        G_[4][0][i]= K_ONE+PX3H1[i]; G_[5][0][i]=-K_ONE+PX4H1[i]; G_[6][0][i]= K_ONE+PX1H1[i]; G_[7][0][i]=-K_ONE+PX2H1[i];
        G_[0][1][i]= K_ONE-PX1H2[i]; G_[1][1][i]= K_ONE-PX2H2[i]; G_[2][1][i]=-K_ONE-PX3H2[i]; G_[3][1][i]=-K_ONE-PX4H2[i];
        G_[4][1][i]=-K_ONE+PX3H2[i]; G_[5][1][i]=-K_ONE+PX4H2[i]; G_[6][1][i]= K_ONE+PX1H2[i]; G_[7][1][i]= K_ONE+PX2H2[i];
        G_[0][2][i]= K_ONE-PX1H3[i]; G_[1][2][i]=-K_ONE-PX2H3[i]; G_[2][2][i]=-K_ONE-PX3H3[i]; G_[3][2][i]= K_ONE-PX4H3[i];
        G_[4][2][i]=-K_ONE+PX3H3[i]; G_[5][2][i]= K_ONE+PX4H3[i]; G_[6][2][i]= K_ONE+PX1H3[i]; G_[7][2][i]=-K_ONE+PX2H3[i];
        for(int m=0;m<3;m++){
          auto dot=[&](Vec* V){ return G_[0][m][i]*V[0][i]+G_[1][m][i]*V[1][i]+G_[2][m][i]*V[2][i]+G_[3][m][i]*V[3][i]
                                      +G_[4][m][i]*V[4][i]+G_[5][m][i]*V[5][i]+G_[6][m][i]*V[6][i]+G_[7][m][i]*V[7][i]; };
          HGX[m][i]=dot(VX); HGY[m][i]=dot(VY); HGZ[m][i]=dot(VZ);
        }
        HGX[3][i]=VX[0][i]-VX[1][i]+VX[2][i]-VX[3][i]-VX[4][i]+VX[5][i]-VX[6][i]+VX[7][i];
        HGY[3][i]=VY[0][i]-VY[1][i]+VY[2][i]-VY[3][i]-VY[4][i]+VY[5][i]-VY[6][i]+VY[7][i];
        HGZ[3][i]=VZ[0][i]-VZ[1][i]+VZ[2][i]-VZ[3][i]-VZ[4][i]+VZ[5][i]-VZ[6][i]+VZ[7][i];
      }
    }
    for(int i=0;i<nel;i++){
      double HX[4],HY[4],HZ[4];
      for(int m=0;m<4;m++){
        HX[m]=HGX[m][i]*(FCL[i]+std::fabs(HGX[m][i])*FCQ[i]);
        HY[m]=HGY[m][i]*(FCL[i]+std::fabs(HGY[m][i])*FCQ[i]);
        HZ[m]=HGZ[m][i]*(FCL[i]+std::fabs(HGZ[m][i])*FCQ[i]);
      }
      auto fill=[&](double (*F)[MVSIZ],const double* H){
        if(jhbe==0){
          F[0][i]=-H[0]-H[1]-H[2]-H[3];
          F[1][i]= H[0]-H[1]+H[2]+H[3];
          F[2][i]=-H[0]+H[1]+H[2]-H[3];
          F[3][i]= H[0]+H[1]-H[2]+H[3];
          F[4][i]=-H[0]+H[1]+H[2]+H[3];
          F[5][i]= H[0]+H[1]-H[2]-H[3];
          F[6][i]=-H[0]-H[1]-H[2]+H[3];
          F[7][i]= H[0]-H[1]+H[2]-H[3];
        } else {
          F[0][i]=-G_[0][0][i]*H[0]-G_[0][1][i]*H[1]-G_[0][2][i]*H[2]-H[3];
          F[1][i]=-G_[1][0][i]*H[0]-G_[1][1][i]*H[1]-G_[1][2][i]*H[2]+H[3];
          F[2][i]=-G_[2][0][i]*H[0]-G_[2][1][i]*H[1]-G_[2][2][i]*H[2]-H[3];
          F[3][i]=-G_[3][0][i]*H[0]-G_[3][1][i]*H[1]-G_[3][2][i]*H[2]+H[3];
          F[4][i]=-G_[4][0][i]*H[0]-G_[4][1][i]*H[1]-G_[4][2][i]*H[2]+H[3];
          F[5][i]=-G_[5][0][i]*H[0]-G_[5][1][i]*H[1]-G_[5][2][i]*H[2]-H[3];
          F[6][i]=-G_[6][0][i]*H[0]-G_[6][1][i]*H[1]-G_[6][2][i]*H[2]+H[3];
          F[7][i]=-G_[7][0][i]*H[0]-G_[7][1][i]*H[1]-G_[7][2][i]*H[2]-H[3];
        }
      };
      fill(F1,HX); fill(F2,HY); fill(F3,HZ);
    }
  }
  /* ---- SFINT3  sfint3.F:257-323 (Lagrangian: SVIS=0) */
  for(int i=0;i<nel;i++){
    double QVIS_LOC=QVIS[i], VOL_LOC=VOLN[i];
    double s1=(SG(i,0)+K_ZERO-QVIS_LOC)*VOL_LOC;
    double s2=(SG(i,1)+K_ZERO-QVIS_LOC)*VOL_LOC;
    double s3=(SG(i,2)+K_ZERO-QVIS_LOC)*VOL_LOC;
    double s4=(SG(i,3)+K_ZERO)*VOL_LOC;
    double s5=(SG(i,4)+K_ZERO)*VOL_LOC;
    double s6=(SG(i,5)+K_ZERO)*VOL_LOC;
    const double* PX[4]={PX1,PX2,PX3,PX4}; const double* PY[4]={PY1,PY2,PY3,PY4}; const double* PZ[4]={PZ1,PZ2,PZ3,PZ4};
    const int a[4]={0,1,2,3}, b[4]={6,7,4,5};   /* pairs 1-7, 2-8, 3-5, 4-6 */
    for(int k=0;k<4;k++){
      double FINT=s1*PX[k][i]+s4*PY[k][i]+s6*PZ[k][i];
      F1[a[k]][i]=F1[a[k]][i]-FINT; F1[b[k]][i]=F1[b[k]][i]+FINT;
      FINT=s2*PY[k][i]+s4*PX[k][i]+s5*PZ[k][i];
      F2[a[k]][i]=F2[a[k]][i]-FINT; F2[b[k]][i]=F2[b[k]][i]+FINT;
      FINT=s3*PZ[k][i]+s6*PX[k][i]+s5*PY[k][i];
      F3[a[k]][i]=F3[a[k]][i]-FINT; F3[b[k]][i]=F3[b[k]][i]+FINT;
    }
  }
  if(jcvt!=0){                                        /* sforc3.F:1634-1645 SRROTA3(R11,R21,R31,R12,...): F_global = R F_local */
    for(int i=0;i<nel;i++) for(int k=0;k<8;k++){
      const double X=R11[i]*F1[k][i]+R12[i]*F2[k][i]+R13[i]*F3[k][i];
      const double Y=R21[i]*F1[k][i]+R22[i]*F2[k][i]+R23[i]*F3[k][i];
      const double Z=R31[i]*F1[k][i]+R32[i]*F2[k][i]+R33[i]*F3[k][i];
      F1[k][i]=X; F2[k][i]=Y; F3[k][i]=Z;
    }
  }
  /* ---- SCUMU3P  scumu3p.F:104-309 (IPARTSPH=0, JTHE>=0, IVECTOR=0) */
  {
    double off_l=K_ZERO; for(int i=0;i<nel;i++) off_l=std::min(off_l,OFFG[i]);
    if(off_l<K_ZERO) for(int i=0;i<nel;i++) if(OFFG[i]<K_ZERO) for(int k=0;k<8;k++){F1[k][i]=K_ZERO;F2[k][i]=K_ZERO;F3[k][i]=K_ZERO;}
    for(int i=0;i<nel;i++) STI[i]=K_FOURTH*STI[i];
    for(int i=0;i<nel;i++){
      int ii=i+nft;
      for(int k=0;k<8;k++){
        int K=o.IADS[8*(size_t)ii+k]-1;
        double* f=&o.FSKY[8*(size_t)K];
        f[0]=F1[k][i]; f[1]=F2[k][i]; f[2]=F3[k][i]; f[6]=STI[i];
      }
    }
  }
}
