/* oracle/shell_qeph.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * QEPH 4-node shell (Ishell=24), restated per element from CZFORC3
 * (engine/source/elements/shell/coquez/czforc3.F:380-823, ISROT=0, IORTH=0, no XFEM / thermal /
 * non-local branches) and the routines it calls:
 *   CZCORC1   coquez/czcorc.F:143-611       gather, local frame, small strain, char. length,
 *                                           2nd-order rigid-rotation correction
 *   CLSKEW3   sh3n/coquedk/cdkcoor3.F:336-397 (IREP=0)
 *   CZCORP5   coquez/czcorp5.F:83-346       warped-element projection (explicit: full projection)
 *   CNCOEF3B  sh3n/coquedk/cncoef3.F:77-284 (IGTYP=1 branch)
 *   CZDEF     coquez/czdef.F:115-200        strain rates + hourglass rates
 *   CZSTRA3   coquez/czstra3.F:77-119       strain increments, GBUF%STRA
 *   epsd_pg   czforc3.F:582-591
 *   CMAIN3    -> oracle/shell_mat.cpp
 *   CNDT3     sh3n/coquedk/cndt3.F:85-318   (NODADT=0, IDTMINS=0, IDTMIN(3)=0)
 *   CZFINTCE  coquez/czfintce.F:63-104
 *   CZFINTN1  coquez/czfintn.F:86-453       physical hourglass stabilisation (MTN/=58)
 *   CZPROJ1 / CZPROJN coquez/czproj.F:1206-1467 (IFINI=0)
 *   CUPDTN3P  coque/cupdtn3.F:545-699       corner rows into FSKY(8,IADC)
 */
#include "shell.h"

void orc_czforc3(Oracle& o, OrcShellGroup& g, double& DT2T, int& NELTST, int& ITYPTST)
{
  const int nel=g.nel;
  const int ISMSTR=g.prop.ismstr, ITHK=g.prop.ithk, NPT=g.prop.npt;
  const double DT1=o.DT1;
  const double TOL=K_EM8;                     /* IRESP=0 */
  for(int i=0;i<nel;i++){
    const int* ix=&o.IXC[(size_t)7*(g.nft+i)];
    const int n1=ix[1]-1,n2=ix[2]-1,n3=ix[3]-1,n4=ix[4]-1;
    const int NGL=ix[6];
    const double* X=o.X.data(); const double* V=o.V.data(); const double* VR=o.VR.data();
    double OFFG=g.OFF[i];
    double SIGY=K_EP30, ALPE=K_ONE;
    const double FAC1=g.prop.cvis;            /* GEO(17,PID) */
    /* ---- CZCORC1 */
    double RX=X[3*n2]+X[3*n3]-X[3*n1]-X[3*n4];
    double SX=X[3*n3]+X[3*n4]-X[3*n1]-X[3*n2];
    double RY=X[3*n2+1]+X[3*n3+1]-X[3*n1+1]-X[3*n4+1];
    double SY=X[3*n3+1]+X[3*n4+1]-X[3*n1+1]-X[3*n2+1];
    double RZ=X[3*n2+2]+X[3*n3+2]-X[3*n1+2]-X[3*n4+2];
    double SSZ=X[3*n3+2]+X[3*n4+2]-X[3*n1+2]-X[3*n2+2];
    /* CLSKEW3 (IREP=0) */
    double E1X,E1Y,E1Z,E2X,E2Y,E2Z,E3X,E3Y,E3Z,DETA1;
    {
      E3X=RY*SSZ-RZ*SY; E3Y=RZ*SX-RX*SSZ; E3Z=RX*SY-RY*SX;
      double DET=std::sqrt(E3X*E3X+E3Y*E3Y+E3Z*E3Z);
      if(DET<K_EM20 && OFFG!=K_ZERO) OFFG=K_ZERO;
      double OFF_LOC=K_ZERO; if(std::fabs(OFFG)!=K_ZERO) OFF_LOC=K_ONE;
      DET=std::max(K_EM20,DET);
      double CC=OFF_LOC/DET; CC=std::max(CC,K_EM20);
      E3X=E3X*CC; E3Y=E3Y*CC; E3Z=E3Z*CC;
      double C1C1=RX*RX+RY*RY+RZ*RZ, C2C2=SX*SX+SY*SY+SSZ*SSZ;
      double C2_1=K_ZERO,C1_1=K_ZERO;
      if(C1C1!=K_ZERO){ C2_1=std::sqrt(C2C2/std::max(K_EM20,C1C1)); C1_1=K_ONE; }
      else if(C2C2!=K_ZERO){ C2_1=K_ONE; C1_1=std::sqrt(C1C1/std::max(K_EM20,C2C2)); }
      E1X=RX*C2_1+(SY*E3Z-SSZ*E3Y)*C1_1;
      E1Y=RY*C2_1+(SSZ*E3X-SX*E3Z)*C1_1;
      E1Z=RZ*C2_1+(SX*E3Y-SY*E3X)*C1_1;
      double C1=std::sqrt(E1X*E1X+E1Y*E1Y+E1Z*E1Z);
      if(C1!=K_ZERO) C1=K_ONE/std::max(K_EM20,C1);
      E1X=E1X*C1; E1Y=E1Y*C1; E1Z=E1Z*C1;
      E2X=E3Y*E1Z-E3Z*E1Y; E2Y=E3Z*E1X-E3X*E1Z; E2Z=E3X*E1Y-E3Y*E1X;
      DETA1=DET;
    }
    const double R11=E1X,R12=E2X,R13=E3X,R21=E1Y,R22=E2Y,R23=E3Y,R31=E1Z,R32=E2Z,R33=E3Z;
    double AREA=K_FOURTH*DETA1;
    double AREA_I;
    { double OFF_LOC=K_ZERO; if(std::fabs(OFFG)!=K_ZERO) OFF_LOC=K_ONE; AREA_I=OFF_LOC/AREA; AREA_I=std::max(AREA_I,K_EM20); }
    /* VQ(a,b) = R_ab */
    const double VQ[3][3]={{R11,R12,R13},{R21,R22,R23},{R31,R32,R33}};
    double XL2,YL2,XL3,YL3,XL4,YL4,Z1;
    {
      double L0[3];
      for(int c=0;c<3;c++) L0[c]=K_FOURTH*(X[3*n3+c]+X[3*n4+c]+X[3*n1+c]+X[3*n2+c]);
      double XX=X[3*n2]-X[3*n1], YY=X[3*n2+1]-X[3*n1+1], ZZ=X[3*n2+2]-X[3*n1+2];
      XL2=R11*XX+R21*YY+R31*ZZ; YL2=R12*XX+R22*YY+R32*ZZ;
      XX=X[3*n1]-L0[0]; YY=X[3*n1+1]-L0[1]; ZZ=X[3*n1+2]-L0[2];
      Z1=R13*XX+R23*YY+R33*ZZ;
      XX=X[3*n3]-X[3*n1]; YY=X[3*n3+1]-X[3*n1+1]; ZZ=X[3*n3+2]-X[3*n1+2];
      XL3=R11*XX+R21*YY+R31*ZZ; YL3=R12*XX+R22*YY+R32*ZZ;
      XX=X[3*n4]-X[3*n1]; YY=X[3*n4+1]-X[3*n1+1]; ZZ=X[3*n4+2]-X[3*n1+2];
      XL4=R11*XX+R21*YY+R31*ZZ; YL4=R12*XX+R22*YY+R32*ZZ;
    }
    /* small strain */
    double* SM=g.SMSTR.data();
    if(ISMSTR==1||ISMSTR==2){
      if(std::fabs(OFFG)==K_TWO){
        XL2=SM[i]; YL2=SM[nel+i]; XL3=SM[2*nel+i]; YL3=SM[3*nel+i]; XL4=SM[4*nel+i]; YL4=SM[5*nel+i];
        Z1=K_ZERO;
        AREA=K_HALF*((XL2-XL4)*YL3-XL3*(YL2-YL4));
        AREA_I=K_ONE/std::max(K_EM20,AREA);
      } else {
        SM[i]=XL2; SM[nel+i]=YL2; SM[2*nel+i]=XL3; SM[3*nel+i]=YL3; SM[4*nel+i]=XL4; SM[5*nel+i]=YL4;
      }
    }
    if(ISMSTR==1){ if(OFFG==K_ONE) OFFG=K_TWO; }
    /* local corner coordinates */
    double COREL[2][4];
    double X13,X24,Y13,Y24,MX13,MX23,MX34,MY13,MY23,MY34,L13,L24,HS;
    {
      double LX=K_FOURTH*(XL2+XL3+XL4), LY=K_FOURTH*(YL2+YL3+YL4);
      COREL[0][0]=-LX; COREL[0][1]=XL2-LX; COREL[0][2]=XL3-LX; COREL[0][3]=XL4-LX;
      COREL[1][0]=-LY; COREL[1][1]=YL2-LY; COREL[1][2]=YL3-LY; COREL[1][3]=YL4-LY;
      X13=(COREL[0][0]-COREL[0][2])*K_HALF; X24=(COREL[0][1]-COREL[0][3])*K_HALF;
      Y13=(COREL[1][0]-COREL[1][2])*K_HALF; Y24=(COREL[1][1]-COREL[1][3])*K_HALF;
      MX13=(COREL[0][0]+COREL[0][2])*K_HALF; MX23=(COREL[0][1]+COREL[0][2])*K_HALF; MX34=(COREL[0][2]+COREL[0][3])*K_HALF;
      MY13=(COREL[1][0]+COREL[1][2])*K_HALF; MY23=(COREL[1][1]+COREL[1][2])*K_HALF; MY34=(COREL[1][2]+COREL[1][3])*K_HALF;
      L13=X13*X13+Y13*Y13; L24=X24*X24+Y24*Y24;
      double C1=COREL[0][1]*COREL[1][3]-COREL[1][1]*COREL[0][3];
      double C2=COREL[0][0]*COREL[1][2]-COREL[1][0]*COREL[0][2];
      HS=std::max(std::fabs(C1),std::fabs(C2))*AREA_I;
    }
    /* characteristic length */
    double LL,LM,FACN1,FACN2;
    {
      const double FACDT=K_FIVE_OVER_4;
      double rx=XL2+XL3-XL4, ry=YL2+YL3-YL4, sx=-XL2+XL3+XL4, sy=-YL2+YL3+YL4;
      double C1=std::sqrt(rx*rx+ry*ry), C2=std::sqrt(sx*sx+sy*sy);
      double S1=K_FOURTH*(std::max(C1,C2)/std::min(C1,C2)-K_ONE);
      double fac1=std::min(K_HALF,S1)+K_ONE;
      double fac2=K_FOUR*AREA/(C1*C2);
      fac2=(double)3.413f*std::max(K_ZERO,fac2-(double)0.7071f);
      fac2=(double)0.78f+(double)0.22f*fac2*fac2*fac2;
      double FACI=K_TWO*fac1*fac2;
      LL=std::max(L13,L24);
      LM=K_HALF*(L13+L24);
      FACN1=std::sqrt(L24/LL); FACN2=std::sqrt(L13/LL);
      S1=std::sqrt(FACI*(FACDT+HS)*LL);
      S1=std::max(S1,K_EM10);
      LL=AREA/S1;
    }
    /* nodal rotation rates in the local frame */
    double RL[2][4];
    { const int nn[4]={n1,n2,n3,n4};
      for(int k=0;k<4;k++){ int K=nn[k];
        RL[0][k]=VQ[0][0]*VR[3*K]+VQ[1][0]*VR[3*K+1]+VQ[2][0]*VR[3*K+2];
        RL[1][k]=VQ[0][1]*VR[3*K]+VQ[1][1]*VR[3*K+1]+VQ[2][1]*VR[3*K+2]; } }
    double V13[3],V24[3],VHI[3];
    {
      double VG13[3],VG24[3],VGHI[3];
      for(int c=0;c<3;c++){
        VG13[c]=V[3*n1+c]-V[3*n3+c];
        VG24[c]=V[3*n2+c]-V[3*n4+c];
        VGHI[c]=V[3*n1+c]-V[3*n2+c]+V[3*n3+c]-V[3*n4+c];
      }
      for(int c=0;c<3;c++){
        V13[c]=(VQ[0][c]*VG13[0]+VQ[1][c]*VG13[1]+VQ[2][c]*VG13[2]);
        V24[c]=(VQ[0][c]*VG24[0]+VQ[1][c]*VG24[1]+VQ[2][c]*VG24[2]);
        VHI[c]=(VQ[0][c]*VGHI[0]+VQ[1][c]*VGHI[1]+VQ[2][c]*VGHI[2]);
      }
    }
    /* 2nd-order rigid rotation correction (IMPL_S=0) */
    {
      const double DT05=K_HALF*DT1, DT025=K_FOURTH*DT1;
      double EXZ=Y24*V13[2]-Y13*V24[2];
      double EYZ=-X24*V13[2]+X13*V24[2];
      double DDRY=DT05*EXZ*AREA_I, DDRX=DT05*EYZ*AREA_I;
      double V13X=V13[0],V24X=V24[0],VHIX=VHI[0];
      double DDRZ1,DDRZ2;
      if(std::fabs(X13-X24)<K_EM10) DDRZ1=K_ZERO; else DDRZ1=DT025*(V13[1]-V24[1])/(X13-X24);
      V13[0]=V13[0]-DDRY*V13[2]-DDRZ1*V13[1];
      V24[0]=V24[0]-DDRY*V24[2]-DDRZ1*V24[1];
      VHI[0]=VHI[0]-DDRY*VHI[2]-DDRZ1*VHI[1];
      if(std::fabs(Y13+Y24)<K_EM10) DDRZ2=K_ZERO; else DDRZ2=DT025*(V13X+V24X)/(Y13+Y24);
      V13[1]=V13[1]-DDRX*V13[2]-DDRZ2*V13X;
      V24[1]=V24[1]-DDRX*V24[2]-DDRZ2*V24X;
      VHI[1]=VHI[1]-DDRX*VHI[2]-DDRZ2*VHIX;
    }
    /* ---- CZCORP5 */
    bool PLAT;
    double VQN[3][4]={{0}}, DI[6]={0}, DB[3][4]={{0}};
    {
      double Z2=Z1*Z1;
      if(Z2<LM*TOL || NPT==1){ Z1=K_ZERO; PLAT=true; }
      else {
        PLAT=false;
        const double A_4=AREA*K_FOURTH;
        double SZ1=MX13*Y24-MY13*X24;
        double SZ2=A_4+SZ1;
        double SZ=Z2*L24;
        double SL=K_ONE/std::sqrt(SZ+SZ2*SZ2);
        VQN[0][0]=-Z1*Y24; VQN[1][0]=Z1*X24; VQN[2][0]=SZ2*SL;
        VQN[0][2]=-VQN[0][0]; VQN[1][2]=-VQN[1][0];
        VQN[0][0]=VQN[0][0]*SL; VQN[1][0]=VQN[1][0]*SL;
        SZ2=A_4-SZ1;
        SL=K_ONE/std::sqrt(SZ+SZ2*SZ2);
        VQN[0][2]=VQN[0][2]*SL; VQN[1][2]=VQN[1][2]*SL; VQN[2][2]=SZ2*SL;
        SZ1=MX13*Y13-MY13*X13;
        SZ2=A_4+SZ1;
        SZ=Z2*L13;
        SL=K_ONE/std::sqrt(SZ+SZ2*SZ2);
        VQN[0][1]=-Z1*Y13; VQN[1][1]=Z1*X13; VQN[2][1]=SZ2*SL;
        VQN[0][3]=-VQN[0][1]; VQN[1][3]=-VQN[1][1];
        VQN[0][1]=VQN[0][1]*SL; VQN[1][1]=VQN[1][1]*SL;
        SZ2=A_4-SZ1;
        SL=K_ONE/std::sqrt(SZ+SZ2*SZ2);
        VQN[0][3]=VQN[0][3]*SL; VQN[1][3]=VQN[1][3]*SL; VQN[2][3]=SZ2*SL;
        double RR[3][4];
        { const int nn[4]={n1,n2,n3,n4};
          for(int k=0;k<4;k++){ int K=nn[k]; RR[0][k]=RL[0][k]; RR[1][k]=RL[1][k];
            RR[2][k]=VQ[0][2]*VR[3*K]+VQ[1][2]*VR[3*K+1]+VQ[2][2]*VR[3*K+2]; } }
        /* full projection */
        double AR[3],AD[4];
        AR[0]=-Z1*VHI[1]+Y13*V13[2]+Y24*V24[2]+MY13*VHI[2]+RR[0][0]+RR[0][1]+RR[0][2]+RR[0][3];
        AR[1]= Z1*VHI[0]-X13*V13[2]-X24*V24[2]-MX13*VHI[2]+RR[1][0]+RR[1][1]+RR[1][2]+RR[1][3];
        AR[2]= X13*V13[1]+X24*V24[1]+MX13*VHI[1]-Y13*V13[0]-Y24*V24[0]-MY13*VHI[0]+RR[2][0]+RR[2][1]+RR[2][2]+RR[2][3];
        for(int k=0;k<4;k++) AD[k]=VQN[0][k]*RR[0][k]+VQN[1][k]*RR[1][k]+VQN[2][k]*RR[2][k];
        double XX=COREL[0][0]*COREL[0][0]+COREL[0][1]*COREL[0][1]+COREL[0][2]*COREL[0][2]+COREL[0][3]*COREL[0][3];
        double YY=COREL[1][0]*COREL[1][0]+COREL[1][1]*COREL[1][1]+COREL[1][2]*COREL[1][2]+COREL[1][3]*COREL[1][3];
        double XY=COREL[0][0]*COREL[1][0]+COREL[0][1]*COREL[1][1]+COREL[0][2]*COREL[1][2]+COREL[0][3]*COREL[1][3];
        double XZ=(COREL[0][0]-COREL[0][1]+COREL[0][2]-COREL[0][3])*Z1;
        double YZ=(COREL[1][0]-COREL[1][1]+COREL[1][2]-COREL[1][3])*Z1;
        double ZZ=K_FOUR*Z2;
        double BTB[6];
        BTB[0]=VQN[0][0]*VQN[0][0]+VQN[0][1]*VQN[0][1]+VQN[0][2]*VQN[0][2]+VQN[0][3]*VQN[0][3];
        BTB[1]=VQN[1][0]*VQN[1][0]+VQN[1][1]*VQN[1][1]+VQN[1][2]*VQN[1][2]+VQN[1][3]*VQN[1][3];
        BTB[2]=VQN[2][0]*VQN[2][0]+VQN[2][1]*VQN[2][1]+VQN[2][2]*VQN[2][2]+VQN[2][3]*VQN[2][3];
        BTB[3]=VQN[0][0]*VQN[1][0]+VQN[0][1]*VQN[1][1]+VQN[0][2]*VQN[1][2]+VQN[0][3]*VQN[1][3];
        BTB[4]=VQN[0][0]*VQN[2][0]+VQN[0][1]*VQN[2][1]+VQN[0][2]*VQN[2][2]+VQN[0][3]*VQN[2][3];
        BTB[5]=VQN[1][0]*VQN[2][0]+VQN[1][1]*VQN[2][1]+VQN[1][2]*VQN[2][2]+VQN[1][3]*VQN[2][3];
        double D[6];
        D[0]=YY+ZZ+K_FOUR-BTB[0]; D[1]=XX+ZZ+K_FOUR-BTB[1]; D[2]=XX+YY+K_FOUR-BTB[2];
        D[3]=-XY-BTB[3]; D[4]=-XZ-BTB[4]; D[5]=-YZ-BTB[5];
        double ABC=D[0]*D[1]*D[2];
        double XXYZ2=D[0]*D[5]*D[5], YYXZ2=D[1]*D[4]*D[4], ZZXY2=D[2]*D[3]*D[3];
        double DETA=std::fabs(ABC+K_TWO*D[3]*D[4]*D[5]-XXYZ2-YYXZ2-ZZXY2);
        DETA=K_ONE/std::max(DETA,K_EM20);
        DI[0]=(ABC-XXYZ2)*DETA/std::max(D[0],K_EM20);
        DI[1]=(ABC-YYXZ2)*DETA/std::max(D[1],K_EM20);
        DI[2]=(ABC-ZZXY2)*DETA/std::max(D[2],K_EM20);
        DI[3]=(D[4]*D[5]-D[3]*D[2])*DETA;
        DI[4]=(D[5]*D[3]-D[4]*D[1])*DETA;
        DI[5]=(D[3]*D[4]-D[5]*D[0])*DETA;
        for(int J=0;J<4;J++){
          DB[0][J]=DI[0]*VQN[0][J]+DI[3]*VQN[1][J]+DI[4]*VQN[2][J];
          DB[1][J]=DI[3]*VQN[0][J]+DI[1]*VQN[1][J]+DI[5]*VQN[2][J];
          DB[2][J]=DI[4]*VQN[0][J]+DI[5]*VQN[1][J]+DI[2]*VQN[2][J];
        }
        double DBAD[3],ALR[3],ALD[4];
        for(int c=0;c<3;c++) DBAD[c]=DB[c][0]*AD[0]+DB[c][1]*AD[1]+DB[c][2]*AD[2]+DB[c][3]*AD[3];
        ALR[0]=DI[0]*AR[0]+DI[3]*AR[1]+DI[4]*AR[2]-DBAD[0];
        ALR[1]=DI[3]*AR[0]+DI[1]*AR[1]+DI[5]*AR[2]-DBAD[1];
        ALR[2]=DI[4]*AR[0]+DI[5]*AR[1]+DI[2]*AR[2]-DBAD[2];
        for(int k=0;k<4;k++)
          ALD[k]=AD[k]+VQN[0][k]*DBAD[0]+VQN[1][k]*DBAD[1]+VQN[2][k]*DBAD[2]-DB[0][k]*AR[0]-DB[1][k]*AR[1]-DB[2][k]*AR[2];
        double C1=K_TWO*ALR[2];
        V13[0]=V13[0]+C1*Y13; V24[0]=V24[0]+C1*Y24;
        VHI[0]=VHI[0]+K_FOUR*(ALR[2]*MY13-Z1*ALR[1]);
        V13[1]=V13[1]-C1*X13; V24[1]=V24[1]-C1*X24;
        VHI[1]=VHI[1]-K_FOUR*(ALR[2]*MX13-Z1*ALR[0]);
        V13[2]=V13[2]-K_TWO*(Y13*ALR[0]-X13*ALR[1]);
        V24[2]=V24[2]-K_TWO*(Y24*ALR[0]-X24*ALR[1]);
        VHI[2]=VHI[2]+K_FOUR*(MX13*ALR[1]-MY13*ALR[0]);
        for(int k=0;k<4;k++){
          RL[0][k]=RR[0][k]-ALR[0]-VQN[0][k]*ALD[k];
          RL[1][k]=RR[1][k]-ALR[1]-VQN[1][k]*ALD[k];
        }
      }
    }
    for(int c=0;c<3;c++){ V13[c]=V13[c]*AREA_I; V24[c]=V24[c]*AREA_I; VHI[c]=VHI[c]*K_FOURTH; }
    /* ---- CNCOEF3B (IGTYP=1, material constants from PM) */
    double THK0;
    if(ITHK>0) THK0=std::max(K_EM20,g.THK[i]); else THK0=g.THKE[i];
    const double THK02=THK0*THK0;
    double VOL0=THK0*AREA;
    const double DT1C=DT1;
    double RHO,YM,NU,G,A11,A12,SSP,GSR,A11SR,A12SR,NUSR;
    if(g.law==36){ const orgpu_law36& m=g.m36; RHO=m.rho0; YM=m.young; NU=m.nu; G=m.shear; A11=m.a11; A12=m.a12; SSP=m.ssp;
                   GSR=m.gsr; A11SR=m.a11sr; A12SR=m.a12sr; NUSR=m.nusr; }
    else         { const orgpu_law2& m=g.m2;   RHO=m.rho0; YM=m.young; NU=m.nu; G=m.shear; A11=m.a11; A12=m.a12; SSP=m.ssp;
                   GSR=m.gsr; A11SR=m.a11sr; A12SR=m.a12sr; NUSR=m.nusr; }
    double SHF,SHFSR;
    if(NPT==1){ SHF=K_ZERO; SHFSR=K_ZERO; } else { SHF=g.prop.shf; SHFSR=g.prop.shfsr; }
    const double GS=G*SHF;
    if(g.law>=24){ A12=NU*A11; A12SR=NUSR*A11SR; }
    double DN=g.prop.h1; if(DN==K_ZERO) DN=K_ZEP01+K_FIVEEM3;
    const double AMU=DN;
    const double ZOFFSET=K_ZERO*THK0;
    (void)YM;
    /* ---- CZDEF */
    double VDEF[8],VHG[6],OFF;
    {
      double R13v[2],R24v[2],RSOM[2],RHI[2];
      for(int c=0;c<2;c++){
        R13v[c]=(RL[c][0]-RL[c][2])*AREA_I;
        R24v[c]=(RL[c][1]-RL[c][3])*AREA_I;
        RSOM[c]=(RL[c][3]+RL[c][2]+RL[c][0]+RL[c][1])*AREA_I;
        RHI[c]=(RL[c][0]-RL[c][1]+RL[c][2]-RL[c][3])*K_FOURTH;
      }
      VDEF[0]=Y24*V13[0]-Y13*V24[0];
      VDEF[1]=-X24*V13[1]+X13*V24[1];
      double BXV2=Y24*V13[1]-Y13*V24[1];
      double BYV1=-X24*V13[0]+X13*V24[0];
      VDEF[2]=BXV2+BYV1;
      VDEF[5]=Y24*R13v[1]-Y13*R24v[1];
      VDEF[6]=X24*R13v[0]-X13*R24v[0];
      double BXR1=Y13*R24v[0]-Y24*R13v[0];
      double BYR2=-X24*R13v[1]+X13*R24v[1];
      VDEF[7]=BXR1+BYR2;
      double BCXY=AREA*K_FOURTH;
      double BCX=V13[2]-MY13*R13v[0]+MX13*R13v[1];
      double BCY=V24[2]+MY13*R24v[0]-MX13*R24v[1];
      VDEF[3]=Y24*BCX-Y13*BCY+BCXY*RSOM[1];
      VDEF[4]=X13*BCY-X24*BCX-BCXY*RSOM[0];
      VHG[0]=VHI[0]-MX13*VDEF[0]-MY13*BYV1;
      VHG[1]=VHI[1]-MX13*BXV2-MY13*VDEF[1];
      VHG[2]=RHI[1]-MX13*VDEF[5]-MY13*BYR2;
      VHG[3]=-RHI[0]-MX13*BXR1-MY13*VDEF[6];
      VHG[4]=(VHI[2]*4.-(MY13*RSOM[0]-MY23*(R13v[0]+R24v[0])+MX23*(R13v[1]+R24v[1])-MX13*RSOM[1])*AREA)*K_FOUR;
      VHG[5]=(VHI[2]*4.-(MY13*RSOM[0]-MY34*(R13v[0]-R24v[0])+MX34*(R13v[1]-R24v[1])-MX13*RSOM[1])*AREA)*K_FOUR;
      VHG[0]=VHG[0]+(Y24*V13[2]-Y13*V24[2])*Z1;
      VHG[1]=VHG[1]+(-X24*V13[2]+X13*V24[2])*Z1;
      double DETA=Z1*K_FOUR*AREA_I;
      VDEF[5]=VDEF[5]+(X13*V13[0]-X24*V24[0])*DETA;
      VDEF[6]=VDEF[6]+(Y13*V13[1]-Y24*V24[1])*DETA;
      VDEF[7]=VDEF[7]+(X13*V13[1]-X24*V24[1]+Y13*V13[0]-Y24*V24[0])*DETA;
      OFF=std::min(K_ONE,std::fabs(OFFG));
      if(OFFG<K_ZERO){ for(int k=0;k<8;k++) VDEF[k]=K_ZERO; for(int k=0;k<6;k++) VHG[k]=K_ZERO; }
    }
    /* ---- CZSTRA3 */
    ShellMatIn mi;
    mi.exx=VDEF[0]*DT1C; mi.eyy=VDEF[1]*DT1C; mi.exy=VDEF[2]*DT1C;
    mi.eyz=VDEF[4]*DT1C; mi.exz=VDEF[3]*DT1C;
    mi.kxx=VDEF[5]*DT1C; mi.kyy=VDEF[6]*DT1C; mi.kxy=VDEF[7]*DT1C;
    if(g.prop.istrain!=0){
      double* S=g.STRA.data();
      S[i]=S[i]+mi.exx; S[nel+i]=S[nel+i]+mi.eyy; S[2*nel+i]=S[2*nel+i]+mi.exy;
      S[3*nel+i]=S[3*nel+i]+mi.eyz; S[4*nel+i]=S[4*nel+i]+mi.exz;
      S[5*nel+i]=S[5*nel+i]+mi.kxx; S[6*nel+i]=S[6*nel+i]+mi.kyy; S[7*nel+i]=S[7*nel+i]+mi.kxy;
    }
    /* global element strain rate (czforc3.F:582-591) */
    {
      const double dtinv=DT1/std::max(DT1*DT1,K_EM20);
      const double asrate=K_ONE;
      double thk=g.THK[i];
      double eps_k2=(mi.kxx*mi.kxx+mi.kyy*mi.kyy+mi.kxx*mi.kyy+K_FOURTH*(mi.kxy*mi.kxy))*K_ONE_OVER_9*(thk*thk);
      double eps_m2=K_FOUR_OVER_3*(mi.exx*mi.exx+mi.eyy*mi.eyy+mi.exx*mi.eyy+K_FOURTH*(mi.exy*mi.exy));
      mi.epsd_pg=std::sqrt(eps_k2+eps_m2)*dtinv;
      g.EPSD[i]=asrate*mi.epsd_pg+(K_ONE-asrate)*g.EPSD[i];
    }
    /* ---- CMAIN3 */
    mi.area=AREA; mi.thk0=THK0; mi.off=OFF; mi.nu=NU; mi.g=G; mi.a11=A11; mi.a12=A12; mi.gs=GS; mi.shf=SHF;
    mi.rho=RHO; mi.ssp=SSP; mi.dt1c=DT1C;
    ShellMatOut mo; mo.sigy=SIGY;
    orc_cmain3(o,g,i,true,mi,mo);
    OFF=mi.off; SSP=mo.ssp; SIGY=mo.sigy; VOL0=mo.vol0;
    double VISCMX=mo.viscmx;
    const double ZCFAC[2]={mo.zcfac1,mo.zcfac2};
    if(o.ipri){ const int nd[4]={n1,n2,n3,n4}; orc_bilan_shell(o,g.nft+i,4,nd,g.EINT[i],g.EINT[nel+i],RHO,OFF); }   /* CBILAN czforc3.F:639 */
    /* ---- CNDT3 */
    double STI,STIR;
    {
      VISCMX=std::max(VISCMX,AMU);
      VISCMX=std::sqrt(K_ONE+VISCMX*VISCMX)-VISCMX;
      double ALDT=LL*VISCMX/std::sqrt(ALPE);
      double F_OSET=K_ONE+K_HALF*std::fabs(ZOFFSET)/THK0;
      double F_DTE=K_ONE/std::sqrt(F_OSET);
      double DT=o.ctl.dtfac_shell*F_DTE*ALDT/SSP;
      if(o.ctl.nodadt!=0){
        /* cndt3.F:96-101 F_OSET (IGTYP=1), :194-221 nodal stiffnesses; :231 and :294 return before the element dt */
        if(OFF==K_ZERO){ STI=K_ZERO; STIR=K_ZERO; }
        else {
          STI=K_HALF*F_OSET*VOL0*A11/(ALDT*ALDT);
          STIR=STI*(THK0*THK0+AREA)*K_ONE_OVER_12+STI*ZOFFSET*ZOFFSET;
        }
      } else {
      if(OFFG>K_ZERO && OFF!=K_ZERO && DT<DT2T){ DT2T=DT; NELTST=NGL; ITYPTST=3; }
      double DIVM=std::max(ALDT*ALDT,K_EM20);
      STI=K_HALF*F_OSET*VOL0*A11*OFF/DIVM;
      STIR=K_ZERO;
      }
    }
    /* ---- CZFINTCE */
    double VF[3][4]={{0}},VM[2][4]={{0}};
    const double* VS=g.FOR.data(); const double* MS_=g.MOM.data();
#define VSTRE(k) VS[(size_t)(k-1)*nel+i]
#define MSTRE(k) MS_[(size_t)(k-1)*nel+i]
    {
      double X13S8=X13*MSTRE(3), X24S8=X24*MSTRE(3), Y13S8=Y13*MSTRE(3), Y24S8=Y24*MSTRE(3);
      double S1=(MY34*MX23-MY23*MX34)*THK0;
      double S42S=S1*VSTRE(5), S52S=S1*VSTRE(4);
      VF[0][0]=THK0*(Y24*VSTRE(1)-X24*VSTRE(3));
      VF[1][0]=THK0*(-X24*VSTRE(2)+Y24*VSTRE(3));
      VF[2][0]=THK0*(-X24*VSTRE(4)+Y24*VSTRE(5));
      VM[0][0]=THK02*(X24*MSTRE(2)-Y24S8)-MY13*VF[2][0];
      VM[1][0]=THK02*(Y24*MSTRE(1)-X24S8)+MX13*VF[2][0];
      VM[0][2]=-S52S; VM[1][2]=S42S;
      VF[0][1]=THK0*(-Y13*VSTRE(1)+X13*VSTRE(3));
      VF[1][1]=THK0*(X13*VSTRE(2)-Y13*VSTRE(3));
      VF[2][1]=THK0*(X13*VSTRE(4)-Y13*VSTRE(5));
      VM[0][1]=THK02*(-X13*MSTRE(2)+Y13S8)+MY13*VF[2][1];
      VM[1][1]=THK02*(-Y13*MSTRE(1)+X13S8)-MX13*VF[2][1];
      VM[0][3]=VM[0][2]; VM[1][3]=VM[1][2];
      double C2=THK02*Z1*4.*AREA_I;
      VF[0][0]=VF[0][0]+C2*(X13*MSTRE(1)+Y13S8);
      VF[1][0]=VF[1][0]+C2*(Y13*MSTRE(2)+X13S8);
      VF[0][1]=VF[0][1]-C2*(X24*MSTRE(1)+Y24S8);
      VF[1][1]=VF[1][1]-C2*(Y24*MSTRE(2)+X24S8);
    }
    /* ---- CZFINTN1 (MTN /= 58) */
    {
      double* VG=g.HOURG.data();
#define VGLAS(k) VG[(size_t)(k-1)*nel+i]
      const double C7=K_FOUR_OVER_3, COEFH=K_ZEP999, COEF=K_ZEP85, STIER=K_FIVEP333, TOLh=K_EM18;
      double FBEND_V=K_THREEP464, FBEND, COEF1;
      if(NPT==0) COEF1=K_SIXTEEN; else COEF1=K_TWENTY5;
      if(NPT==1){ FBEND=K_ZERO; FBEND_V=K_ZERO; } else FBEND=K_ONE_OVER_12;
      const double UNDOUZSR=std::sqrt(K_ONE_OVER_12);
      double DHG[6]; for(int k=0;k<6;k++) DHG[k]=VHG[k]*DT1;
      double DGLAS[13];
      double C3=K_FOUR*AREA_I;
      double HXX=C3*MY34, HYY=C3*MX34, HXX_K=C3*MY23, HYY_K=C3*MX23;
      double CXX=HXX*DHG[0], CYY=HYY*DHG[1], CXX_K=HXX_K*DHG[0], CYY_K=HYY_K*DHG[1];
      double BXX=HXX*DHG[2], BYY=HYY*DHG[3], BXX_K=HXX_K*DHG[2], BYY_K=HYY_K*DHG[3];
      double C1M=A11*FAC1, C2M=A12*FAC1;
      const double C6=THK02*FBEND;
      double SS1=MY34*VGLAS(1)+MY23*VGLAS(7);
      double SS2=MX23*VGLAS(8)+MX34*VGLAS(2);
      double SF1=MY34*VGLAS(3)+MY23*VGLAS(9);
      double SF2=-MX23*VGLAS(10)-MX34*VGLAS(4);
      double SC5=MY34*VGLAS(5)+MX34*VGLAS(6);
      double SC6=MY23*VGLAS(11)+MX23*VGLAS(12);
      double C5=K_HALF*OFF*THK0*C7;
      const double ESX=SS1*DHG[0]+SS2*DHG[1];
      double ETMP1=C5*(ESX+K_FOURTH*(SC5*DHG[4]+SC6*DHG[5]));
      const double EMX=(SF1*DHG[2]-SF2*DHG[3])*C6;
      double ETMP2=C5*EMX;
      DGLAS[1]=C1M*CXX-C2M*CYY;   DGLAS[2]=C1M*CYY-C2M*CXX;
      DGLAS[3]=C1M*BXX-C2M*BYY;   DGLAS[4]=C1M*BYY-C2M*BXX;
      DGLAS[7]=C1M*CXX_K-C2M*CYY_K; DGLAS[8]=C1M*CYY_K-C2M*CXX_K;
      DGLAS[9]=C1M*BXX_K-C2M*BYY_K; DGLAS[10]=C1M*BYY_K-C2M*BXX_K;
      double C2=FAC1*G*SHF*K_ONE_OVER_64;
      DGLAS[5]=C2*HXX*DHG[4]; DGLAS[6]=C2*HYY*DHG[4];
      DGLAS[11]=C2*HXX_K*DHG[5]; DGLAS[12]=C2*HYY_K*DHG[5];
      for(int k=1;k<=12;k++) VGLAS(k)=VGLAS(k)+DGLAS[k];
      g.EINT[i]=g.EINT[i]+ETMP1; g.EINT[nel+i]=g.EINT[nel+i]+ETMP2;
      if(SIGY<K_ZEP9EP30){
        double UFAC=std::fabs(std::min(ZCFAC[0],ZCFAC[1])-K_ONE);
        double SIGY2=SIGY*SIGY, SVM=K_ZERO, SXY0=K_ZERO;
        if(UFAC<TOLh){
          SXY0=VSTRE(1)*VSTRE(1)+VSTRE(2)*VSTRE(2)-VSTRE(1)*VSTRE(2)+K_THREE*VSTRE(3)*VSTRE(3);
          double MXY0=MSTRE(1)*MSTRE(1)+MSTRE(2)*MSTRE(2)-MSTRE(1)*MSTRE(2)+K_THREE*MSTRE(3)*MSTRE(3);
          double CNN=COEF, CMM=COEF*THK0*K_ONE_OVER_16;
          double CNNX=CNN*VGLAS(1), CNNY=CNN*VGLAS(2), CNNX_K=CNN*VGLAS(7), CNNY_K=CNN*VGLAS(8);
          double CMMX=CMM*VGLAS(3), CMMY=CMM*VGLAS(4), CMMX_K=CMM*VGLAS(9), CMMY_K=CMM*VGLAS(10);
          SXY0=SXY0+CNNX*CNNX+CNNY*CNNY-CNNX*CNNY;
          MXY0=MXY0+CMMX*CMMX+CMMY*CMMY-CMMX*CMMY;
          SXY0=SXY0+CNNX_K*CNNX_K+CNNY_K*CNNY_K-CNNX_K*CNNY_K;
          MXY0=MXY0+CMMX_K*CMMX_K+CMMY_K*CMMY_K-CMMX_K*CMMY_K;
          SXY0=SXY0+std::fabs(CNNX*(K_TWO*CNNX_K-CNNY_K)+CNNY*(K_TWO*CNNY_K-CNNX_K));
          MXY0=MXY0+std::fabs(CMMX*(K_TWO*CMMX_K-CMMY_K)+CMMY*(K_TWO*CMMY_K-CMMX_K));
          SVM=SXY0+COEF1*MXY0;
        }
        if(UFAC>=TOLh || SVM>SIGY2){
          double EH1=std::min(SXY0/std::max(SIGY2,TOLh),K_ONE);
          EH1=std::max(COEFH*EH1,(K_ONE-ZCFAC[0]));
          double EH2=std::max(COEFH,(K_ONE-ZCFAC[1]));
          if(ESX<K_ZERO) EH1=K_ZERO;
          if(EMX<K_ZERO) EH2=K_ZERO;
          VGLAS(1)=VGLAS(1)-EH1*DGLAS[1]; VGLAS(2)=VGLAS(2)-EH1*DGLAS[2];
          VGLAS(7)=VGLAS(7)-EH1*DGLAS[7]; VGLAS(8)=VGLAS(8)-EH1*DGLAS[8];
          VGLAS(3)=VGLAS(3)-EH2*DGLAS[3]; VGLAS(4)=VGLAS(4)-EH2*DGLAS[4];
          VGLAS(9)=VGLAS(9)-EH2*DGLAS[9]; VGLAS(10)=VGLAS(10)-EH2*DGLAS[10];
        }
      }
      const double C8=C7*OFF;
      SS1=(MY34*VGLAS(1)+MY23*VGLAS(7))*C8;
      SS2=(MX23*VGLAS(8)+MX34*VGLAS(2))*C8;
      SF1=(MY34*VGLAS(3)+MY23*VGLAS(9))*C8;
      SF2=-(MX23*VGLAS(10)+MX34*VGLAS(4))*C8;
      const double HSURA=THK0*AREA_I;
      C2=C8*THK0;
      SC5=(MY34*VGLAS(5)+MX34*VGLAS(6))*C2;
      SC6=(MY23*VGLAS(11)+MX23*VGLAS(12))*C2;
      double SS3=SC5+SC6;
      const double HVL=AMU*std::sqrt(RHO*AREA*FAC1)*OFF;
      double SSV0=MY23*MY23, SSV1=MY34*MY34, SSV2=MX23*MX23, SSV3=MX34*MX34;
      double HXX_V=STIER*(SSV1+SSV0);
      double HXY_V=-STIER*(MY34*MX34+MY23*MX23);
      double HYY_V=STIER*(SSV2+SSV3);
      C2=HVL*GSR*SHFSR*UNDOUZSR;
      double CXZ_V=(SSV1+SSV3)*C2, CYZ_V=(SSV2+SSV0)*C2;
      double AUX=AREA_I*HVL;
      C1M=A11SR*AUX; C2M=A12SR*AUX;
      double CXX_V=C1M*HXX_V, CYY_V=C1M*HYY_V, CXY_V=C2M*HXY_V;
      double SS1_V=CXX_V*VHG[0]+CXY_V*VHG[1];
      double SS2_V=CYY_V*VHG[1]+CXY_V*VHG[0];
      double SF1_V=(CXX_V*VHG[2]+CXY_V*VHG[3])*FBEND_V;
      double SF2_V=(-CYY_V*VHG[3]-CXY_V*VHG[2])*FBEND_V;
      double SC5_V=CXZ_V*VHG[4]*HSURA;
      double SC6_V=CYZ_V*VHG[5]*HSURA;
      double SS3_V=SC5_V+SC6_V;
      SS1=SS1+SS1_V; SS2=SS2+SS2_V; SS3=SS3+SS3_V; SC5=SC5+SC5_V; SC6=SC6+SC6_V; SF1=SF1+SF1_V; SF2=SF2+SF2_V;
      double Y13S=MY13*SS3, X13S=MX13*SS3, Y34S6=MY34*SC6, Y23S5=MY23*SC5, X23S5=MX23*SC5, X34S6=MX34*SC6;
      C2=K_FOURTH*THK0;
      double B13=(MY13*X24-MX13*Y24)*HSURA;
      VF[0][0]=VF[0][0]+B13*SS1; VF[0][2]=C2*SS1;
      VF[1][0]=VF[1][0]+B13*SS2; VF[1][2]=C2*SS2;
      VF[2][2]=SS3;
      double B24=(MX13*Y13-MY13*X13)*HSURA;
      VF[0][1]=VF[0][1]+B24*SS1; VF[0][3]=-VF[0][2];
      VF[1][1]=VF[1][1]+B24*SS2; VF[1][3]=-VF[1][2];
      VF[2][3]=-VF[2][2];
      C3=C6*B13; double C4=C6*C2;
      VM[0][0]=VM[0][0]+C3*SF2+Y23S5+Y34S6;
      VM[0][2]=VM[0][2]+C4*SF2-Y13S;
      VM[1][0]=VM[1][0]+C3*SF1-X23S5-X34S6;
      VM[1][2]=VM[1][2]+C4*SF1+X13S;
      C3=C6*B24;
      VM[0][1]=VM[0][1]+C3*SF2+Y23S5-Y34S6;
      VM[0][3]=VM[0][3]-C4*SF2-Y13S;
      VM[1][1]=VM[1][1]+C3*SF1-X23S5+X34S6;
      VM[1][3]=VM[1][3]-C4*SF1+X13S;
      C2=Z1*HSURA;
      VF[2][0]=VF[2][0]+C2*(SS1*Y24-SS2*X24);
      VF[2][1]=VF[2][1]+C2*(-SS1*Y13+SS2*X13);
      double ESY=((SS1-SS1_V)*DHG[0]+(SS2-SS2_V)*DHG[1])*THK0+K_FOURTH*((SC5-SC5_V)*DHG[4]+(SC6-SC6_V)*DHG[5]);
      ETMP1=K_HALF*ESY;
      double EMY=(SF1-SF1_V)*DHG[2]-(SF2-SF2_V)*DHG[3];
      ETMP2=K_HALF*C6*EMY*THK0;
      g.EINT[i]=g.EINT[i]+ETMP1; g.EINT[nel+i]=g.EINT[nel+i]+ETMP2;
#undef VGLAS
    }
#undef VSTRE
#undef MSTRE
    /* ---- CZPROJ1 / CZPROJN (IFINI=0) */
    double Fg[3][4],Mg[3][4];
    {
      double FL[3][4],ML[2][4],MM[3][4];
      for(int c=0;c<3;c++){
        FL[c][0]=VF[c][0]+VF[c][2]; FL[c][1]=VF[c][1]+VF[c][3];
        FL[c][2]=-VF[c][0]+VF[c][2]; FL[c][3]=-VF[c][1]+VF[c][3];
      }
      for(int c=0;c<2;c++){
        ML[c][0]=VM[c][0]+VM[c][2]; ML[c][1]=VM[c][1]+VM[c][3];
        ML[c][2]=-VM[c][0]+VM[c][2]; ML[c][3]=-VM[c][1]+VM[c][3];
      }
      if(PLAT){
        for(int J=0;J<4;J++) for(int I=0;I<3;I++){
          Fg[I][J]=VQ[I][0]*FL[0][J]+VQ[I][1]*FL[1][J]+VQ[I][2]*FL[2][J];
          Mg[I][J]=VQ[I][0]*ML[0][J]+VQ[I][1]*ML[1][J];
        }
      } else {
        double AR[3],AD[4],DBAD[3],ALR[3],ALD[4];
        AR[0]=-Z1*(FL[1][0]-FL[1][1]+FL[1][2]-FL[1][3])
              +COREL[1][0]*FL[2][0]+ML[0][0]+COREL[1][1]*FL[2][1]+ML[0][1]
              +COREL[1][2]*FL[2][2]+ML[0][2]+COREL[1][3]*FL[2][3]+ML[0][3];
        AR[1]= Z1*(FL[0][0]-FL[0][1]+FL[0][2]-FL[0][3])
              -COREL[0][0]*FL[2][0]+ML[1][0]-COREL[0][1]*FL[2][1]+ML[1][1]
              -COREL[0][2]*FL[2][2]+ML[1][2]-COREL[0][3]*FL[2][3]+ML[1][3];
        AR[2]=-COREL[1][0]*FL[0][0]+COREL[0][0]*FL[1][0]-COREL[1][1]*FL[0][1]+COREL[0][1]*FL[1][1]
              -COREL[1][2]*FL[0][2]+COREL[0][2]*FL[1][2]-COREL[1][3]*FL[0][3]+COREL[0][3]*FL[1][3];
        for(int k=0;k<4;k++) AD[k]=VQN[0][k]*ML[0][k]+VQN[1][k]*ML[1][k];
        for(int c=0;c<3;c++) DBAD[c]=DB[c][0]*AD[0]+DB[c][1]*AD[1]+DB[c][2]*AD[2]+DB[c][3]*AD[3];
        ALR[0]=DI[0]*AR[0]+DI[3]*AR[1]+DI[4]*AR[2]-DBAD[0];
        ALR[1]=DI[3]*AR[0]+DI[1]*AR[1]+DI[5]*AR[2]-DBAD[1];
        ALR[2]=DI[4]*AR[0]+DI[5]*AR[1]+DI[2]*AR[2]-DBAD[2];
        for(int k=0;k<4;k++)
          ALD[k]=AD[k]+VQN[0][k]*DBAD[0]+VQN[1][k]*DBAD[1]+VQN[2][k]*DBAD[2]-DB[0][k]*AR[0]-DB[1][k]*AR[1]-DB[2][k]*AR[2];
        double C1=Z1*ALR[1];
        FL[0][0]=FL[0][0]-C1+COREL[1][0]*ALR[2]; FL[0][1]=FL[0][1]+C1+COREL[1][1]*ALR[2];
        FL[0][2]=FL[0][2]-C1+COREL[1][2]*ALR[2]; FL[0][3]=FL[0][3]+C1+COREL[1][3]*ALR[2];
        C1=Z1*ALR[0];
        FL[1][0]=FL[1][0]+C1-COREL[0][0]*ALR[2]; FL[1][1]=FL[1][1]-C1-COREL[0][1]*ALR[2];
        FL[1][2]=FL[1][2]+C1-COREL[0][2]*ALR[2]; FL[1][3]=FL[1][3]-C1-COREL[0][3]*ALR[2];
        for(int J=0;J<4;J++){
          FL[2][J]=FL[2][J]-COREL[1][J]*ALR[0]+COREL[0][J]*ALR[1];
          MM[0][J]=ML[0][J]-ALR[0]-VQN[0][J]*ALD[J];
          MM[1][J]=ML[1][J]-ALR[1]-VQN[1][J]*ALD[J];
          MM[2][J]=-ALR[2]-VQN[2][J]*ALD[J];
        }
        for(int J=0;J<4;J++) for(int I=0;I<3;I++){
          Fg[I][J]=VQ[I][0]*FL[0][J]+VQ[I][1]*FL[1][J]+VQ[I][2]*FL[2][J];
          Mg[I][J]=VQ[I][0]*MM[0][J]+VQ[I][1]*MM[1][J]+VQ[I][2]*MM[2][J];
        }
      }
    }
    /* ---- CUPDTN3P */
    if(OFF<K_ONE) OFFG=OFF;
    if(OFFG<K_ZERO){ for(int J=0;J<4;J++) for(int I=0;I<3;I++){ Fg[I][J]=K_ZERO; Mg[I][J]=K_ZERO; } STI=K_ZERO; STIR=K_ZERO; }
    const double FACN[2]={FACN1,FACN2};
    for(int J=0;J<4;J++){
      const int K=o.IADC[(size_t)4*(g.nft+i)+J]-1;
      double* f=&o.FSKY[(size_t)8*K];
      f[0]=-Fg[0][J]; f[1]=-Fg[1][J]; f[2]=-Fg[2][J];
      f[3]=-Mg[0][J]; f[4]=-Mg[1][J]; f[5]=-Mg[2][J];
      f[6]=STI*FACN[J&1]; f[7]=STIR*FACN[J&1];
    }
    g.OFF[i]=OFFG;
  }
}
