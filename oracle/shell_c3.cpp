/* oracle/shell_c3.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * 3-node C0 shell (ITY=7, Ish3n 1/2: IFRAM_OLD=1), restated per element from C3FORC3
 * (engine/source/elements/sh3n/coque3n/c3forc3.F:296-720; IGTYP=1, no drilling dof, no XFEM / thermal /
 * non-local / drape) and the routines it calls:
 *   C3COOR3  coque3n/c3coor3.F      gather, OFF, deleted-element velocity reset
 *   C3EVEC3  coque3n/c3evec3.F:89-153   frame: e1 along 1-2, e3 normal, AREA = |x31 x x32| / 2
 *   C3DERI3  coque3n/c3deri3.F      local coordinates, small-strain reference SMSTR(3), PX1/PY1/PY2, ALDT
 *   C3COEF3  coque3n/c3coef3.F      THK0/VOL0, material constants, SHF
 *   C3DEFO3  coque3n/c3defo3.F      membrane / shear strain rates with the rigid-rotation correction
 *   C3CURV3  coque3n/c3curv3.F      curvature rates, shear from the nodal rotations
 *   C3STRA3  coque3n/c3stra3.F      increments, GBUF%STRA
 *   epsd_pg  c3forc3.F:538-556 ; CMAIN3 -> oracle/shell_mat.cpp (ISH3N in the place of IHBE)
 *   C3DT3    coque3n/c3dt3.F        element dt (DTFAC1(7), ITYPTST=7), STI / STIR (NODADT 0 and 1)
 *   C3FINT3  coque3n/c3fint3.F ; C3FCUM3 c3fcum3.F ; C3MCUM3 c3mcum3.F : internal forces local -> global
 *   C3UPDT3P coque3n/c3updt3.F      corner rows into FSKY(8,IADTG)
 */
#include "shell.h"

void orc_c3forc3(Oracle& o, OrcShellGroup& g, double& DT2T, int& NELTST, int& ITYPTST)
{
  const int nel=g.nel;
  const int ISMSTR=g.prop.ismstr, ITHK=g.prop.ithk, NPT=g.prop.npt, ISH3N=g.prop.ihbe;
  const double DT1=o.DT1;
  for(int i=0;i<nel;i++){
    const int* ix=&o.IXTG[(size_t)6*(g.nft+i)];
    const int nn[3]={ix[1]-1,ix[2]-1,ix[3]-1};
    const int NGL=ix[5];
    const double* X=o.X.data();
    double OFFG=g.OFF[i];
    /* ---- C3COOR3 */
    double XG[3],YG[3],ZG[3],VL[3][3],VRL[3][3];
    for(int k=0;k<3;k++){
      XG[k]=X[3*nn[k]]; YG[k]=X[3*nn[k]+1]; ZG[k]=X[3*nn[k]+2];
      for(int c=0;c<3;c++){ VL[k][c]=o.V[3*nn[k]+c]; VRL[k][c]=o.VR[3*nn[k]+c]; }
    }
    double THK0=g.THKE[i];
    const double DT1C=DT1;
    double OFF=std::min(K_ONE,std::fabs(OFFG));
    if(OFFG<K_ZERO) for(int k=0;k<3;k++) for(int c=0;c<3;c++){ VL[k][c]=K_ZERO; VRL[k][c]=K_ZERO; }
    /* ---- C3EVEC3 (IFRAM_OLD=1) */
    double E1X,E1Y,E1Z,E2X,E2Y,E2Z,E3X,E3Y,E3Z,AREA;
    const double X21G=XG[1]-XG[0],Y21G=YG[1]-YG[0],Z21G=ZG[1]-ZG[0];
    const double X31G=XG[2]-XG[0],Y31G=YG[2]-YG[0],Z31G=ZG[2]-ZG[0];
    {
      const double X32=XG[2]-XG[1],Y32=YG[2]-YG[1],Z32=ZG[2]-ZG[1];
      E1X=X21G; E1Y=Y21G; E1Z=Z21G;
      double SUM=std::sqrt(E1X*E1X+E1Y*E1Y+E1Z*E1Z);
      E1X=E1X/SUM; E1Y=E1Y/SUM; E1Z=E1Z/SUM;
      E3X=Y31G*Z32-Z31G*Y32; E3Y=Z31G*X32-X31G*Z32; E3Z=X31G*Y32-Y31G*X32;
      SUM=std::sqrt(E3X*E3X+E3Y*E3Y+E3Z*E3Z);
      E3X=E3X/SUM; E3Y=E3Y/SUM; E3Z=E3Z/SUM;
      AREA=K_HALF*SUM;
      E2X=E3Y*E1Z-E3Z*E1Y; E2Y=E3Z*E1X-E3X*E1Z; E2Z=E3X*E1Y-E3Y*E1X;
      SUM=std::sqrt(E2X*E2X+E2Y*E2Y+E2Z*E2Z);
      E2X=E2X/SUM; E2Y=E2Y/SUM; E2Z=E2Z/SUM;
    }
    /* ---- C3DERI3 (ISMSTR /= 3) */
    double X2,Y2,X3,Y3,PX1,PY1,PY2,ALDT; const double ALPE=K_ONE;
    {
      X2=E1X*X21G+E1Y*Y21G+E1Z*Z21G; Y2=E2X*X21G+E2Y*Y21G+E2Z*Z21G;
      X3=E1X*X31G+E1Y*Y31G+E1Z*Z31G; Y3=E2X*X31G+E2Y*Y31G+E2Z*Z31G;
      (void)Y2;
      double* SM=g.SMSTR.data();
      if(ISMSTR==1||ISMSTR==2){
        if(OFFG==K_TWO){ X2=SM[i]; X3=SM[nel+i]; Y3=SM[2*nel+i]; AREA=K_HALF*X2*Y3; }
        else { SM[i]=X2; SM[nel+i]=X3; SM[2*nel+i]=Y3; }
      }
      if(ISMSTR==1){ if(OFFG==K_ONE) OFFG=K_TWO; }
      Y3=std::copysign(std::max(K_EM15,std::fabs(Y3)),Y3);
      PX1=-K_HALF*Y3; PY1=K_HALF*(X3-X2); PY2=-K_HALF*X3;
      double AL1=X2*X2, AL2=(X3-X2)*(X3-X2)+Y3*Y3, AL3=X3*X3+Y3*Y3;
      double ALMAX=std::max(std::max(AL1,AL2),AL3);
      ALDT=K_TWO*AREA/std::sqrt(ALMAX);
    }
    /* ---- C3COEF3 */
    if(ITHK>0&&ISMSTR!=3) THK0=g.THK[i];
    const double THK02=THK0*THK0;
    double VOL0=THK0*AREA;
    double RHO,YM,NU,G,A11,A12,SSP;
    if(g.law==36){ const orgpu_law36& m=g.m36; RHO=m.rho0; YM=m.young; NU=m.nu; G=m.shear; A11=m.a11; A12=m.a12; SSP=m.ssp; }
    else         { const orgpu_law2& m=g.m2;   RHO=m.rho0; YM=m.young; NU=m.nu; G=m.shear; A11=m.a11; A12=m.a12; SSP=m.ssp; }
    (void)YM;
    double SHF;
    if(NPT==1) SHF=K_ZERO;
    else { double FAC1=K_TWO*(K_ONE+NU)*THK02; const int ISH=0; double FSH=g.prop.shf;
           SHF=FSH*(K_ONE-ISH+ISH*FAC1/(FSH*AREA+FAC1)); }
    const double GS=G*SHF;
    /* ---- C3DEFO3 */
    double EXX,EYY,EXY,EYZ,EZX;
    {
      double VX1=E1X*VL[0][0]+E1Y*VL[0][1]+E1Z*VL[0][2];
      double VX2=E1X*VL[1][0]+E1Y*VL[1][1]+E1Z*VL[1][2];
      double VX3=E1X*VL[2][0]+E1Y*VL[2][1]+E1Z*VL[2][2];
      double VY3=E2X*VL[2][0]+E2Y*VL[2][1]+E2Z*VL[2][2];
      double VY2=E2X*VL[1][0]+E2Y*VL[1][1]+E2Z*VL[1][2];
      double VY1=E2X*VL[0][0]+E2Y*VL[0][1]+E2Z*VL[0][2];
      double VZ1=E3X*VL[0][0]+E3Y*VL[0][1]+E3Z*VL[0][2];
      double VZ2=E3X*VL[1][0]+E3Y*VL[1][1]+E3Z*VL[1][2];
      double VZ3=E3X*VL[2][0]+E3Y*VL[2][1]+E3Z*VL[2][2];
      const double DT1V4=K_FOURTH*DT1;
      double DT1V4B=DT1V4; if(ISH3N<2) DT1V4B=K_ZERO;
      double VZ12=VZ1-VZ2, VZ13=VZ1-VZ3, VZ23=VZ2-VZ3;
      double TMP1=DT1V4*VZ12/(PY1+PY2);
      double TMP2=(PY1*VZ1+PY2*VZ2)/(PY1+PY2);
      TMP2=DT1V4*(TMP2-VZ3)/PX1;
      double VY12=VY1-VY2;
      double TMP11=DT1V4B*VY12/(PY1+PY2);
      double TMP22=(PY1*VX1+PY2*VX2)/(PY1+PY2);
      TMP22=DT1V4B*(TMP22-VX3)/PX1;
      double VX10=VX1,VX20=VX2,VX30=VX3;
      VX1=VX1-VZ1*TMP1-VY1*TMP11;
      VX2=VX2-VZ2*TMP1-VY2*TMP11;
      VX3=VX3-VZ3*TMP1-VY3*TMP11;
      VY1=VY1-VZ1*TMP2-VX10*TMP22;
      VY2=VY2-VZ2*TMP2-VX20*TMP22;
      VY3=VY3-VZ3*TMP2-VX30*TMP22;
      double VX12=VX1-VX2; VY12=VY1-VY2;
      double VX13=VX1-VX3, VY13=VY1-VY3, VX23=VX2-VX3, VY23=VY2-VY3;
      EXX=PX1*VX12;
      EYY=PY1*VY13+PY2*VY23;
      EXY=PY1*VX13+PY2*VX23+PX1*VY12;
      EYZ=PY1*VZ13+PY2*VZ23;
      EZX=PX1*VZ12;
    }
    /* ---- C3CURV3 */
    double KXX,KYY,KXY;
    {
      double RX1=E1X*VRL[0][0]+E1Y*VRL[0][1]+E1Z*VRL[0][2];
      double RY1=E2X*VRL[0][0]+E2Y*VRL[0][1]+E2Z*VRL[0][2];
      double RY2=E2X*VRL[1][0]+E2Y*VRL[1][1]+E2Z*VRL[1][2];
      double RX2=E1X*VRL[1][0]+E1Y*VRL[1][1]+E1Z*VRL[1][2];
      double RX3=E1X*VRL[2][0]+E1Y*VRL[2][1]+E1Z*VRL[2][2];
      double RY3=E2X*VRL[2][0]+E2Y*VRL[2][1]+E2Z*VRL[2][2];
      double RX12T=RX1-RX2, RX13T=RX1-RX3, RX23T=RX2-RX3;
      KYY=-PY1*RX13T-PY2*RX23T;
      KXY=PX1*RX12T;
      double RY12T=RY1-RY2, RY13T=RY1-RY3, RY23T=RY2-RY3;
      KXX=PX1*RY12T;
      KXY=PY1*RY13T+PY2*RY23T-KXY;
      double RYAVT=PX1*(PX1*(-RX1+RX2)
                       +(K_TWO*PY1+K_THREE*PY2)*RY1
                       +(K_THREE*PY1+K_TWO*PY2)*RY2
                       +(PY1+PY2)*RY3);
      double RXAVT=-PX1*(+(K_TWO*PY1+PY2)*RX1
                         +(PY1+K_TWO*PY2)*RX2
                         +K_THREE*(PY1+PY2)*RX3)
                   +PY1*(PY1+K_TWO*PY2)*RY1
                   -PY2*(K_TWO*PY1+PY2)*RY2
                   +(PY2*PY2-PY1*PY1)*RY3;
      EZX=EZX+RYAVT*K_THIRD;
      EYZ=EYZ+RXAVT*K_THIRD;
    }
    /* ---- C3STRA3 (ISMSTR /= 10, 11) */
    ShellMatIn mi;
    {
      double FAC1=DT1/AREA;
      mi.exx=EXX*FAC1; mi.eyy=EYY*FAC1; mi.exy=EXY*FAC1; mi.eyz=EYZ*FAC1; mi.exz=EZX*FAC1;
      mi.kxx=KXX*FAC1; mi.kyy=KYY*FAC1; mi.kxy=KXY*FAC1;
      if(g.prop.istrain!=0){
        double* S=g.STRA.data();
        S[i]=S[i]+mi.exx; S[nel+i]=S[nel+i]+mi.eyy; S[2*nel+i]=S[2*nel+i]+mi.exy;
        S[3*nel+i]=S[3*nel+i]+mi.eyz; S[4*nel+i]=S[4*nel+i]+mi.exz;
        S[5*nel+i]=S[5*nel+i]+mi.kxx; S[6*nel+i]=S[6*nel+i]+mi.kyy; S[7*nel+i]=S[7*nel+i]+mi.kxy;
      }
    }
    {                                                     /* c3forc3.F:546-556 */
      const double dtinv=DT1/std::max(DT1*DT1,K_EM20);
      double thk=g.THK[i];
      double eps_k2=(mi.kxx*mi.kxx+mi.kyy*mi.kyy+mi.kxx*mi.kyy+K_FOURTH*(mi.kxy*mi.kxy))*K_ONE_OVER_9*(thk*thk);
      double eps_m2=K_FOUR_OVER_3*(mi.exx*mi.exx+mi.eyy*mi.eyy+mi.exx*mi.eyy+K_FOURTH*(mi.exy*mi.exy));
      mi.epsd_pg=std::sqrt(eps_k2+eps_m2)*dtinv;
      g.EPSD[i]=K_ONE*mi.epsd_pg+(K_ONE-K_ONE)*g.EPSD[i];
    }
    /* ---- CMAIN3 */
    mi.area=AREA; mi.thk0=THK0; mi.off=OFF; mi.nu=NU; mi.g=G; mi.a11=A11; mi.a12=A12; mi.gs=GS; mi.shf=SHF;
    mi.rho=RHO; mi.ssp=SSP; mi.dt1c=DT1C;
    ShellMatOut mo; mo.sigy=K_EP30;
    orc_cmain3(o,g,i,false,mi,mo);
    OFF=mi.off; SSP=mo.ssp; VOL0=mo.vol0; (void)VOL0;
    if(o.ipri) orc_bilan_shell(o,g.nft+i,3,nn,g.EINT[i],g.EINT[nel+i],RHO,OFF);   /* C3BILAN c3forc3.F:616 */
    /* ---- C3DT3 (IGTYP=1, ZOFFSET=0, IDTMIN(7)=0) */
    double STI,STIR;
    {
      double VISCMX=mo.viscmx;
      VISCMX=std::sqrt(K_ONE+VISCMX*VISCMX)-VISCMX;
      ALDT=ALDT*VISCMX/std::sqrt(ALPE);
      const double F_OSET=K_ONE+K_HALF*std::fabs(K_ZERO)/THK0;
      if(o.ctl.nodadt!=0){
        if(OFF==K_ZERO){ STI=K_ZERO; STIR=K_ZERO; }
        else {
          double ATHK=AREA*THK0;
          STI=ATHK*F_OSET*A11/(ALDT*ALDT);
          STIR=STI*(THK0*THK0*K_ONE_OVER_12+K_HALF*SHF*AREA*G/A11)+STI*K_ZERO*K_ZERO;
        }
      } else {
        const double F_DTE=K_ONE/std::sqrt(F_OSET);
        double DT=o.ctl.dtfac_sh3n*F_DTE*ALDT/SSP;
        if(OFFG>K_ZERO&&OFF!=K_ZERO&&DT<DT2T){ DT2T=DT; NELTST=NGL; ITYPTST=7; }
        STI=AREA*THK0*F_OSET*A11/(ALDT*ALDT);
        STI=K_ZEP81*K_ZEP81*STI*OFF;
        STIR=K_ZERO;
      }
    }
    /* ---- C3FINT3 (NFOR = FOR, NMOM = MOM: C3SROTO3 with IFRAM_OLD=1) */
    double FX[3],FY[3],FZ[3],MX[3],MY[3];
    {
      const double* FO=g.FOR.data(); const double* MO=g.MOM.data();
      double F1=FO[i]*THK0, F3=FO[2*nel+i]*THK0;
      FX[0]= F1*PX1+F3*PY1;
      FX[1]=-F1*PX1+F3*PY2;
      FX[2]=-FX[0]-FX[1];
      double F2=FO[nel+i]*THK0;
      FY[0]=F2*PY1+F3*PX1;
      FY[1]=F2*PY2-F3*PX1;
      FY[2]=-FY[0]-FY[1];
      double F4=FO[3*nel+i]*THK0, F5=FO[4*nel+i]*THK0;
      FZ[0]= F5*PX1+F4*PY1;
      FZ[1]=-F5*PX1+F4*PY2;
      FZ[2]=-FZ[0]-FZ[1];
      double TH2=THK0*THK0;
      double M2=MO[nel+i]*TH2, M3=MO[2*nel+i]*TH2;
      MX[0]=-M2*PY1-M3*PX1;
      MX[1]=-M2*PY2+M3*PX1;
      MX[2]=-MX[0]-MX[1];
      double M1=MO[i]*TH2;
      MY[0]= M1*PX1+M3*PY1;
      MY[1]=-M1*PX1+M3*PY2;
      MY[2]=-MY[0]-MY[1];
      double M4=F4*K_THIRD, M5=F5*K_THIRD;
      M5=M5*PX1;
      MY[0]=MY[0]+M5*(K_TWO*PY1+K_THREE*PY2)+M4*PY1*(PY1+K_TWO*PY2);
      MY[1]=MY[1]+M5*(K_THREE*PY1+K_TWO*PY2)-M4*PY2*(K_TWO*PY1+PY2);
      MY[2]=MY[2]+M5*(PY1+PY2)+M4*(PY2*PY2-PY1*PY1);
      M5=M5*PX1;
      M4=M4*PX1;
      MX[0]=MX[0]-M5-M4*(K_TWO*PY1+PY2);
      MX[1]=MX[1]+M5-M4*(PY1+K_TWO*PY2);
      MX[2]=MX[2]   -M4*K_THREE*(PY1+PY2);
    }
    /* ---- C3FCUM3 / C3MCUM3 (ISIGI /= 5) */
    double F[3][3],M[3][3];
    for(int J=0;J<3;J++){
      F[0][J]=E1X*FX[J]+E2X*FY[J]+E3X*FZ[J];
      F[1][J]=E1Y*FX[J]+E2Y*FY[J]+E3Y*FZ[J];
      F[2][J]=E1Z*FX[J]+E2Z*FY[J]+E3Z*FZ[J];
      M[0][J]=E1X*MX[J]+E2X*MY[J];
      M[1][J]=E1Y*MX[J]+E2Y*MY[J];
      M[2][J]=E1Z*MX[J]+E2Z*MY[J];
    }
    /* ---- C3UPDT3P */
    if(OFF<K_ONE) OFFG=OFF;
    if(OFFG<K_ZERO){ for(int J=0;J<3;J++) for(int I=0;I<3;I++){ F[I][J]=K_ZERO; M[I][J]=K_ZERO; } STI=K_ZERO; STIR=K_ZERO; }
    for(int J=0;J<3;J++){
      const int K=o.IADTG[(size_t)3*(g.nft+i)+J]-1;
      double* f=&o.FSKY[(size_t)8*K];
      f[0]=-F[0][J]; f[1]=-F[1][J]; f[2]=-F[2][J];
      f[3]=-M[0][J]; f[4]=-M[1][J]; f[5]=-M[2][J];
      f[6]=STI; f[7]=STIR;
    }
    g.OFF[i]=OFFG;
  }
}
