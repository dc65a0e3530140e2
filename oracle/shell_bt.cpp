/* oracle/shell_bt.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Belytschko-Tsay 4-node shell (Ishell 1, 3, 4; NPT > 1), restated per element from CFORC3
 * (engine/source/elements/shell/coque/cforc3.F:403-751, ISHFRAM=0, no XFEM / thermal / non-local)
 * and the routines it calls:
 *   CCOOR3  coque/ccoor3.F:60-140     gather, OFF, deleted-element velocity reset
 *   CNVEC3  coque/cnvec3.F:75-141     convected orthonormal frame (ISHFRAM=0)
 *   CDERI3  coque/cderi3.F:85-190     local coordinates, small-strain reference, PX/PY, AREA, VHX/VHY
 *   CCOEF3  coque/ccoef3.F:70-190     THK0/VOL0, material constants, hourglass coefficients, SHF
 *   CDLEN3  coque/cdlen3.F:60-125     characteristic length
 *   CDEFO3  coque/cdefo3.F:65-175     membrane / shear strain rates (IHBE<=1 | 2,3 | 4 branches)
 *   CCURV3  coque/ccurv3.F:60-100     curvature rates
 *   CSTRA3  coque/cstra3.F:85-215     increments, GBUF%STRA
 *   epsd_pg cforc3.F:533-552 ; CMAIN3 -> oracle/shell_mat.cpp
 *   CHVIS3  coque/chvis3.F:120-420    visco-elastic hourglass forces (NODADT=0: STI untouched)
 *   CDT3    coque/cdt3.F:111-232      element dt, arg-min, STI = 0.81*1/2*VOL0*YM/ALDT^2
 *   CFINT3  coque/cfint3.F:147-236    internal forces local -> global
 *   CUPDT3P coque/cupdt3.F:1017-1175  corner rows into FSKY(8,IADC)
 */
#include "shell.h"

void orc_cforc3(Oracle& o, OrcShellGroup& g, double& DT2T, int& NELTST, int& ITYPTST)
{
  const int nel=g.nel;
  const int ISMSTR=g.prop.ismstr, ITHK=g.prop.ithk, NPT=g.prop.npt, IHBE=g.prop.ihbe;
  const double DT1=o.DT1;
  const double HELAS=K_HALF, HVISC=K_HALF, HVLIN=K_ZERO;       /* radioss2.F:641-643 */
  for(int i=0;i<nel;i++){
    const int* ix=&o.IXC[(size_t)7*(g.nft+i)];
    const int nn[4]={ix[1]-1,ix[2]-1,ix[3]-1,ix[4]-1};
    const int NGL=ix[6];
    const double* X=o.X.data();
    double OFFG=g.OFF[i];
    /* ---- CCOOR3 */
    double XG[4],YG[4],ZG[4],VL[4][3],VRL[4][3];
    for(int k=0;k<4;k++){
      XG[k]=X[3*nn[k]]; YG[k]=X[3*nn[k]+1]; ZG[k]=X[3*nn[k]+2];
      for(int c=0;c<3;c++){ VL[k][c]=o.V[3*nn[k]+c]; VRL[k][c]=o.VR[3*nn[k]+c]; }
    }
    double THK0=g.THKE[i];
    const double DT1C=DT1;
    double OFF=std::min(K_ONE,std::fabs(OFFG));
    if(OFFG<K_ZERO) for(int k=0;k<4;k++) for(int c=0;c<3;c++){ VL[k][c]=K_ZERO; VRL[k][c]=K_ZERO; }
    /* ---- CNVEC3 (ISHFRAM=0) */
    double E1X,E1Y,E1Z,E2X,E2Y,E2Z,E3X,E3Y,E3Z;
    {
      double X21=XG[1]-XG[0],X32=XG[2]-XG[1],X34=XG[2]-XG[3],X41=XG[3]-XG[0];
      double Y21=YG[1]-YG[0],Y32=YG[2]-YG[1],Y34=YG[2]-YG[3],Y41=YG[3]-YG[0];
      double Z21=ZG[1]-ZG[0],Z32=ZG[2]-ZG[1],Z34=ZG[2]-ZG[3],Z41=ZG[3]-ZG[0];
      E1X=(X21+X34); E1Y=(Y21+Y34); E1Z=(Z21+Z34);
      E2X=(X32+X41); E2Y=(Y32+Y41); E2Z=(Z32+Z41);
      E3X=E1Y*E2Z-E1Z*E2Y; E3Y=E1Z*E2X-E1X*E2Z; E3Z=E1X*E2Y-E1Y*E2X;
      double SUMA=E3X*E3X+E3Y*E3Y+E3Z*E3Z;
      SUMA=K_ONE/std::max(std::sqrt(SUMA),K_EM20);
      E3X=E3X*SUMA; E3Y=E3Y*SUMA; E3Z=E3Z*SUMA;
      double S1=E1X*E1X+E1Y*E1Y+E1Z*E1Z, S2=E2X*E2X+E2Y*E2Y+E2Z*E2Z;
      SUMA=std::sqrt(S1/S2);
      E1X=E1X+(E2Y*E3Z-E2Z*E3Y)*SUMA;
      E1Y=E1Y+(E2Z*E3X-E2X*E3Z)*SUMA;
      E1Z=E1Z+(E2X*E3Y-E2Y*E3X)*SUMA;
      SUMA=E1X*E1X+E1Y*E1Y+E1Z*E1Z;
      SUMA=K_ONE/std::max(std::sqrt(SUMA),K_EM20);
      E1X=E1X*SUMA; E1Y=E1Y*SUMA; E1Z=E1Z*SUMA;
      E2X=E3Y*E1Z-E3Z*E1Y; E2Y=E3Z*E1X-E3X*E1Z; E2Z=E3X*E1Y-E3Y*E1X;
    }
    /* ---- CDERI3 */
    double STI=K_ZERO,STIR=K_ZERO;
    double X2,Y2,X3,Y3,X4,Y4,Z2;
    {
      double X21=XG[1]-XG[0],Y21=YG[1]-YG[0],Z21=ZG[1]-ZG[0];
      double X31=XG[2]-XG[0],Y31=YG[2]-YG[0],Z31=ZG[2]-ZG[0];
      double X41=XG[3]-XG[0],Y41=YG[3]-YG[0],Z41=ZG[3]-ZG[0];
      X2=E1X*X21+E1Y*Y21+E1Z*Z21; Y2=E2X*X21+E2Y*Y21+E2Z*Z21;
      Y3=E2X*X31+E2Y*Y31+E2Z*Z31; X3=E1X*X31+E1Y*Y31+E1Z*Z31;
      X4=E1X*X41+E1Y*Y41+E1Z*Z41; Y4=E2X*X41+E2Y*Y41+E2Z*Z41;
      Z2=E3X*X21+E3Y*Y21+E3Z*Z21;
    }
    double* SM=g.SMSTR.data();
    if(ISMSTR==1||ISMSTR==2){
      if(std::fabs(OFFG)==K_TWO){
        X2=SM[i]; Y2=SM[nel+i]; X3=SM[2*nel+i]; Y3=SM[3*nel+i]; X4=SM[4*nel+i]; Y4=SM[5*nel+i]; Z2=K_ZERO;
      } else {
        SM[i]=X2; SM[nel+i]=Y2; SM[2*nel+i]=X3; SM[3*nel+i]=Y3; SM[4*nel+i]=X4; SM[5*nel+i]=Y4;
      }
      if(ISMSTR==1){ if(OFFG==K_ONE) OFFG=K_TWO; }
    }
    const double PX1=K_HALF*(Y2-Y4), PY1=K_HALF*(X4-X2), PX2=K_HALF*Y3, PY2=-K_HALF*X3;
    const double AREA=std::max(K_TWO*(PY2*PX1-PY1*PX2),K_EM20);
    const double VHX=(-X2+X3-X4)/AREA, VHY=(-Y2+Y3-Y4)/AREA;
    /* ---- CCOEF3 */
    const double ALPE=K_ONE;
    double VOL0,VOL00,THK02;
    if(ITHK>0&&ISMSTR!=3){ VOL00=THK0*AREA; THK0=g.THK[i]; VOL0=THK0*AREA; THK02=THK0*THK0; }
    else { VOL00=THK0*AREA; VOL0=VOL00; THK02=THK0*THK0; }
    double RHO,YM,NU,G,A11,A12,SSP;
    if(g.law==36){ const orgpu_law36& m=g.m36; RHO=m.rho0; YM=m.young; NU=m.nu; G=m.shear; A11=m.a11; A12=m.a12; SSP=m.ssp; }
    else         { const orgpu_law2& m=g.m2;   RHO=m.rho0; YM=m.young; NU=m.nu; G=m.shear; A11=m.a11; A12=m.a12; SSP=m.ssp; }
    const double H1=g.prop.h1,H2=g.prop.h2,H3=g.prop.h3,SRH1=g.prop.srh1,SRH2=g.prop.srh2,SRH3=g.prop.srh3;
    double SHF;
    if(NPT==1) SHF=K_ZERO;
    else { double FAC1TMP=2.*(1.+NU)*THK02; const int ISH=0; double FSH=g.prop.shf;
           SHF=FSH*(1.-ISH+ISH*FAC1TMP/(FSH*AREA+FAC1TMP)); }
    const double GS=G*SHF;
    /* ---- CDLEN3 */
    double ALDT;
    {
      double AL1=X2*X2+Y2*Y2;
      double AL2=(X3-X2)*(X3-X2)+(Y3-Y2)*(Y3-Y2);
      double AL6=X3*X3+Y3*Y3;
      double AL3=(X4-X3)*(X4-X3)+(Y4-Y3)*(Y4-Y3);
      double AL4=X4*X4+Y4*Y4;
      double AL5=(X4-X2)*(X4-X2)+(Y4-Y2)*(Y4-Y2);
      double ALMIN=std::min(std::min(AL1,AL2),AL4);
      double ALQUAD=std::min(std::min(AL3,AL5),AL6);
      if(AL3!=K_ZERO) ALMIN=std::min(ALMIN,ALQUAD);
      double DTDYN=AREA*AREA/std::max(std::max(AL5,AL6),K_EM20);
      ALDT=std::max(DTDYN,ALMIN);
      double DTHOUR=K_HALF*(ALMIN+ALDT)/std::max(H1,H2);
      if(IHBE!=0){ if(DTHOUR<ALDT) ALDT=DTHOUR; } else ALDT=std::min(ALDT,DTHOUR);
      ALDT=std::sqrt(ALDT);
    }
    /* ---- CDEFO3 */
    double VX[4],VY[4],VZ[4],EXX,EYY,EXY,EXZ,EYZ;
    {
      for(int k=0;k<4;k++) VX[k]=E1X*VL[k][0]+E1Y*VL[k][1]+E1Z*VL[k][2];
      VY[3]=E2X*VL[3][0]+E2Y*VL[3][1]+E2Z*VL[3][2];
      VY[2]=E2X*VL[2][0]+E2Y*VL[2][1]+E2Z*VL[2][2];
      VY[1]=E2X*VL[1][0]+E2Y*VL[1][1]+E2Z*VL[1][2];
      VY[0]=E2X*VL[0][0]+E2Y*VL[0][1]+E2Z*VL[0][2];
      for(int k=0;k<4;k++) VZ[k]=E3X*VL[k][0]+E3Y*VL[k][1]+E3Z*VL[k][2];
      double VZ13=VZ[0]-VZ[2], VZ24=VZ[1]-VZ[3];
      EYZ=PY1*VZ13+PY2*VZ24;
      EXZ=PX1*VZ13+PX2*VZ24;
      if(IHBE<=1){
        Z2=K_ZERO;
        double DT1V4=K_FOURTH*DT1C;
        double TMP2A=PY2+PY1;
        double TMP3A=std::copysign(std::max(std::fabs(TMP2A),K_EM20),TMP2A);
        double TMP1A=DT1V4*(VZ13-VZ24)*(VZ13-VZ24)/TMP3A;
        double VX13=VX[0]-VX[2], VX24=VX[1]-VX[3];
        VX13=VX13-TMP1A; VX24=VX24+TMP1A;
        EXX=PX1*VX13+PX2*VX24;
        EXY=PY1*VX13+PY2*VX24;
        double TMP1B=PX2-PX1;
        double TMP3B=std::copysign(std::max(std::fabs(TMP1B),K_EM20),TMP1B);
        double TMP2B=DT1V4*(VZ13+VZ24)*(VZ13+VZ24)/TMP3B;
        double VY13=VY[0]-VY[2], VY24=VY[1]-VY[3];
        VY13=VY13+TMP2B; VY24=VY24+TMP2B;
        EXY=EXY+PX1*VY13+PX2*VY24;
        EYY=PY1*VY13+PY2*VY24;
      } else if(IHBE==2||IHBE==3){
        double DT1V4=K_HALF*DT1C;
        double GZX=EXZ/AREA, EXZZ2=GZX*Z2, EXZ2=GZX*GZX*DT1V4;
        VX[2]=VX[2]-EXZ2*X3-VX[0];
        VX[1]=VX[1]+EXZZ2-EXZ2*X2-VX[0];
        VX[3]=VX[3]+EXZZ2-EXZ2*X4-VX[0];
        VX[0]=K_ZERO;
        double GZY=EYZ/AREA, EYZZ2=GZY*Z2, EYZ2=GZY*GZY*DT1V4;
        VY[2]=VY[2]-EYZ2*Y3-VY[0];
        VY[1]=VY[1]+EYZZ2-EYZ2*Y2-VY[0];
        VY[3]=VY[3]+EYZZ2-EYZ2*Y4-VY[0];
        VY[0]=K_ZERO;
        double ZZZ=(EXZ2+EYZ2)*Z2;
        VZ[2]=VZ[2]-GZY*Y3-GZX*X3-VZ[0];
        VZ[1]=VZ[1]-GZY*Y2-GZX*X2-ZZZ-VZ[0];
        VZ[3]=VZ[3]-GZY*Y4-GZX*X4-ZZZ-VZ[0];
        VZ[0]=K_ZERO;
        double VX13=-VX[2], VX24=VX[1]-VX[3];
        EXX=PX1*VX13+PX2*VX24;
        EXY=PY1*VX13+PY2*VX24;
        double VY13=-VY[2], VY24=VY[1]-VY[3];
        EXY=EXY+PX1*VY13+PX2*VY24;
        EYY=PY1*VY13+PY2*VY24;
      } else {          /* IHBE == 4 */
        double DT1V4=K_HALF*DT1C;
        double ZZ2=K_HALF*Z2;
        double GZX=EXZ/AREA, EXZZ2=GZX*ZZ2, EXZ2=GZX*GZX*DT1V4, EXZ2PY2=EXZ2*PY2, EXZ2PY1=EXZ2*PY1;
        VX[0]=VX[0]-EXZZ2-EXZ2PY2; VX[2]=VX[2]-EXZZ2+EXZ2PY2;
        VX[1]=VX[1]+EXZZ2+EXZ2PY1; VX[3]=VX[3]+EXZZ2-EXZ2PY1;
        double GZY=EYZ/AREA, EYZZ2=GZY*ZZ2, EYZ2=GZY*GZY*DT1V4, EYZ2PX2=EYZ2*PX2, EYZ2PX1=EYZ2*PX1;
        VY[0]=VY[0]-EYZZ2+EYZ2PX2; VY[2]=VY[2]-EYZZ2-EYZ2PX2;
        VY[1]=VY[1]+EYZZ2-EYZ2PX1; VY[3]=VY[3]+EYZZ2+EYZ2PX1;
        double VX13=VX[0]-VX[2], VX24=VX[1]-VX[3];
        EXX=PX1*VX13+PX2*VX24;
        EXY=PY1*VX13+PY2*VX24;
        double VY13=VY[0]-VY[2], VY24=VY[1]-VY[3];
        EXY=EXY+PX1*VY13+PX2*VY24;
        EYY=PY1*VY13+PY2*VY24;
      }
    }
    /* ---- CCURV3 */
    double RX[4],RY[4],KXX,KYY,KXY;
    {
      for(int k=0;k<4;k++) RX[k]=E1X*VRL[k][0]+E1Y*VRL[k][1]+E1Z*VRL[k][2];
      for(int k=0;k<4;k++) RY[k]=E2X*VRL[k][0]+E2Y*VRL[k][1]+E2Z*VRL[k][2];
      double RX13TA=RX[0]-RX[2], RXAVTA=RX[0]+RX[1]+RX[2]+RX[3], RX24TA=RX[1]-RX[3];
      KYY=-PY1*RX13TA-PY2*RX24TA;
      KXY=PX1*RX13TA+PX2*RX24TA;
      double RY13TA=RY[0]-RY[2], RYAVTA=RY[0]+RY[1]+RY[2]+RY[3], RY24TA=RY[1]-RY[3];
      KXX=PX1*RY13TA+PX2*RY24TA;
      KXY=PY1*RY13TA+PY2*RY24TA-KXY;
      EXZ=EXZ+RYAVTA*(.25*AREA);
      EYZ=EYZ-RXAVTA*(.25*AREA);
    }
    /* ---- CSTRA3 */
    ShellMatIn mi;
    {
      double FAC1=DT1C/AREA;
      mi.exx=EXX*FAC1; mi.eyy=EYY*FAC1; mi.exy=EXY*FAC1; mi.eyz=EYZ*FAC1; mi.exz=EXZ*FAC1;
      mi.kxx=KXX*FAC1; mi.kyy=KYY*FAC1; mi.kxy=KXY*FAC1;
      if(g.prop.istrain!=0){
        double* S=g.STRA.data();
        S[i]=S[i]+mi.exx; S[nel+i]=S[nel+i]+mi.eyy; S[2*nel+i]=S[2*nel+i]+mi.exy;
        S[3*nel+i]=S[3*nel+i]+mi.eyz; S[4*nel+i]=S[4*nel+i]+mi.exz;
        S[5*nel+i]=S[5*nel+i]+mi.kxx; S[6*nel+i]=S[6*nel+i]+mi.kyy; S[7*nel+i]=S[7*nel+i]+mi.kxy;
      }
    }
    {
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
    OFF=mi.off; SSP=mo.ssp; VOL0=mo.vol0;
    double VISCMX=mo.viscmx;
    VISCMX=std::sqrt(K_ONE+VISCMX*VISCMX)-VISCMX;
    /* ---- CHVIS3 */
    double H11,H12,H13,H21,H22,H23,H31,H32,H33,B11,B12,B13,B14,B21,B22,B23,B24;
    {
      double* HOUR=g.HOURG.data();
#define HR(k) HOUR[(size_t)(k-1)*nel+i]
      const double SR2D2=std::sqrt(K_TWO)*K_HALF;
      double GAMA1,GAMA2,GAMA3,GAMA4;
      if(ISMSTR!=1&&ISMSTR!=11&&IHBE>=1){
        double PX1V=PX1*VHX, PX2V=PX2*VHX, PY1V=PY1*VHY, PY2V=PY2*VHY;
        GAMA1=OFF*(K_ONE-PX1V-PY1V); GAMA3=OFF*(K_ONE+PX1V+PY1V);
        GAMA2=OFF*(-K_ONE-PX2V-PY2V); GAMA4=OFF*(-K_ONE+PX2V+PY2V);
      } else { GAMA1=OFF; GAMA3=OFF; GAMA2=-OFF; GAMA4=-OFF; }
      double SHFPR3=SHF/(K_THREE*(K_ONE+NU));
      double HVISH1=HVISC*H1, HVISH2=HVISC*H2;
      double R0=K_FOURTH*RHO; double R1=R0*K_HUNDRED; R0=R0*HVLIN;
      double A1=R1*HVISH1;
      double A2=R0*SR2D2*SRH1;
      double SRSHFPR3=std::sqrt(SHFPR3);
      double A3=R1*HVISH2*SRSHFPR3;
      double A4=R0*SR2D2*SRH2*SRSHFPR3;
      double HH3=HELAS*H3;
      double A5=HH3*R1*K_ZEP072169;
      HH3=SR2D2*SRH3;
      double A6=HH3*R0*K_ZEP072169;
      R0=K_FOURTH*YM*HELAS;
      double A7=H1*R0, A8=H2*R0*SHFPR3;
      double T2A=THK02*AREA, TSA=std::sqrt(T2A);
      double H1Q=A1*TSA, H1L=A2*SSP*TSA, H2Q=A3*THK02, H2L=A4*SSP*THK02, H3Q=A5*T2A, H3L=A6*SSP*T2A;
      double TD=THK0*DT1C;
      double HH1=A7*TD;
      double B1=PX1*PX1+PY1*PY1, B2=PX2*PX2+PY2*PY2;
      double HH2=A8*THK02*TD/(B1+B2);
      if(ix[3]==ix[4]){ H1Q=H1L=H2Q=H2L=H3Q=H3L=HH1=HH2=K_ZERO; }
      double HG1,HG2;
      const bool plain=(ISMSTR==1||ISMSTR==11||IHBE<1);
      if(plain){ HG1=(VX[0]-VX[1]+VX[2]-VX[3])*OFF; HG2=(VY[0]-VY[1]+VY[2]-VY[3])*OFF; }
      else { HG1=VX[0]*GAMA1+VX[1]*GAMA2+VX[2]*GAMA3+VX[3]*GAMA4; HG2=VY[0]*GAMA1+VY[1]*GAMA2+VY[2]*GAMA3+VY[3]*GAMA4; }
      HR(1)=HR(1)+HG1*HH1;
      HR(2)=HR(2)+HG2*HH1;
      double HOUR1A=HR(1)+HG1*(H1L+H1Q*std::fabs(HG1));
      H11=HOUR1A*GAMA1; H12=HOUR1A*GAMA2; H13=HOUR1A*GAMA3;
      double HOUR2A=HR(2)+HG2*(H1L+H1Q*std::fabs(HG2));
      H21=HOUR2A*GAMA1; H22=HOUR2A*GAMA2; H23=HOUR2A*GAMA3;
      if(plain) HG1=(VZ[0]-VZ[1]+VZ[2]-VZ[3])*OFF;
      else HG1=VZ[0]*GAMA1+VZ[1]*GAMA2+VZ[2]*GAMA3+VZ[3]*GAMA4;
      HR(3)=HR(3)+HG1*HH2;
      double HOUR3A=HR(3)+HG1*(H2L+H2Q*std::fabs(HG1));
      H31=HOUR3A*GAMA1; H32=HOUR3A*GAMA2; H33=HOUR3A*GAMA3;
      HG1=RX[0]-RX[1]+RX[2]-RX[3];
      HG2=RY[0]-RY[1]+RY[2]-RY[3];
      HR(4)=HG1*(H3L+H3Q*std::fabs(HG1));
      HR(5)=HG2*(H3L+H3Q*std::fabs(HG2));
      B11=HR(4)*OFF; B12=-HR(4)*OFF; B13=HR(4)*OFF; B14=-HR(4)*OFF;
      B21=HR(5)*OFF; B22=-HR(5)*OFF; B23=HR(5)*OFF; B24=-HR(5)*OFF;
#undef HR
      /* STIFFNESS - DT (chvis3.F:196-207, 242-253; NODADT/=0, IGTYP=1, NADMESH=0); STI enters as ZERO (cderi3.F:91) */
      if(o.ctl.nodadt!=0){
        double SCALE=std::max(std::max(GAMA1*GAMA1,GAMA2*GAMA2),std::max(GAMA3*GAMA3,GAMA4*GAMA4))
                    *DT1C*std::max(std::max(HH1+H1L,HH2+H2L),H3L)/std::max(DT1C*DT1C,K_EM20);
        STI=K_ZERO+SCALE;
        if(OFF==K_ZERO){ STI=K_ZERO; STIR=K_ZERO; }
        else {
          double VV=VISCMX*VISCMX*ALPE;
          STI=STI+std::max(B1,B2)*THK0*A11/(AREA*VV);
          STIR=STI*(THK02*K_ONE_OVER_12+AREA*K_ONE_OVER_9);
        }
      }
    }
    if(o.ipri) orc_bilan_shell(o,g.nft+i,4,nn,g.EINT[i],g.EINT[nel+i],RHO,OFF);   /* CBILAN cforc3.F:648 */
    /* ---- CDT3 (NODADT=0, IDTMIN(3)=2 with DTMIN1(3)=0: no deletion); not called with /DT/NODA (cforc3.F:668) */
    if(o.ctl.nodadt==0){
      ALDT=ALDT*VISCMX/std::sqrt(ALPE);
      double DT=o.ctl.dtfac_shell*ALDT/SSP;
      if(OFFG>K_ZERO&&OFF!=K_ZERO&&DT<DT2T){ DT2T=DT; NELTST=NGL; ITYPTST=3; }
      double DIVM=std::max(ALDT*ALDT,K_EM20);
      STI=K_HALF*VOL0*YM/DIVM;
      STI=K_ZEP81*STI*OFF;
      STIR=K_ZERO;
    }
    (void)VOL00;
    /* ---- CFINT3 */
    double F[3][4],M[3][4];
    {
      const double* FO=g.FOR.data(); const double* MO=g.MOM.data();
      double F1A=FO[i]*THK0, F2A=FO[nel+i]*THK0, F3A=FO[2*nel+i]*THK0, F4A=FO[3*nel+i]*THK0, F5A=FO[4*nel+i]*THK0;
      double M4=F4A*AREA, M5=F5A*AREA;
      double F12=F1A*PX2+F3A*PY2, F22=F2A*PY2+F3A*PX2, F32=F5A*PX2+F4A*PY2;
      double F11=F1A*PX1+F3A*PY1, F21=F2A*PY1+F3A*PX1, F31=F5A*PX1+F4A*PY1;
      double G11=F11+H11, G13=H13-F11, G21=F21+H21, G23=H23-F21, G31=F31+H31, G33=H33-F31;
      double G12=F12+H12, G22=F22+H22, G32=F32+H32;
      F[0][0]=E1X*G11+E2X*G21+E3X*G31; F[0][1]=E1X*G12+E2X*G22+E3X*G32; F[0][2]=E1X*G13+E2X*G23+E3X*G33;
      F[1][0]=E1Y*G11+E2Y*G21+E3Y*G31; F[1][1]=E1Y*G12+E2Y*G22+E3Y*G32; F[1][2]=E1Y*G13+E2Y*G23+E3Y*G33;
      F[2][0]=E1Z*G11+E2Z*G21+E3Z*G31; F[2][1]=E1Z*G12+E2Z*G22+E3Z*G32; F[2][2]=E1Z*G13+E2Z*G23+E3Z*G33;
      F[0][3]=-F[0][0]-F[0][1]-F[0][2];
      F[1][3]=-F[1][0]-F[1][1]-F[1][2];
      F[2][3]=-F[2][0]-F[2][1]-F[2][2];
      if(IHBE>=2&&std::abs(NPT)!=1){ M4=M4+(H21+H23)*Z2; M5=M5+(H11+H13)*Z2; }
      double M1A=MO[i]*THK02, M2A=MO[nel+i]*THK02, M3A=MO[2*nel+i]*THK02;
      M4=M4*K_FOURTH; M5=M5*K_FOURTH;
      double M11=-M2A*PY1-M3A*PX1, M21=M1A*PX1+M3A*PY1, M12=-M2A*PY2-M3A*PX2, M22=M1A*PX2+M3A*PY2;
      double Q11=M11-M4+B11, Q13=-M11-M4+B13, Q12=M12-M4+B12, Q14=-M12-M4+B14;
      double Q21=M21+M5+B21, Q23=-M21+M5+B23, Q22=M22+M5+B22, Q24=-M22+M5+B24;
      const double Q1[4]={Q11,Q12,Q13,Q14}, Q2[4]={Q21,Q22,Q23,Q24};
      for(int J=0;J<4;J++){ M[0][J]=E1X*Q1[J]+E2X*Q2[J]; M[1][J]=E1Y*Q1[J]+E2Y*Q2[J]; M[2][J]=E1Z*Q1[J]+E2Z*Q2[J]; }
    }
    /* ---- CUPDT3P */
    if(OFF<K_ONE) OFFG=OFF;
    if(OFFG<K_ZERO){ for(int J=0;J<4;J++) for(int I=0;I<3;I++){ F[I][J]=K_ZERO; M[I][J]=K_ZERO; } STI=K_ZERO; STIR=K_ZERO; }
    for(int J=0;J<4;J++){
      const int K=o.IADC[(size_t)4*(g.nft+i)+J]-1;
      double* f=&o.FSKY[(size_t)8*K];
      f[0]=-F[0][J]; f[1]=-F[1][J]; f[2]=-F[2][J];
      f[3]=-M[0][J]; f[4]=-M[1][J]; f[5]=-M[2][J];
      f[6]=STI; f[7]=STIR;
    }
    g.OFF[i]=OFFG;
  }
}
