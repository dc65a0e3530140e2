/* oracle/bilan.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Energy / momentum / mass balances as the Engine books them on print cycles (IPRI = 1):
 *   CBILAN  engine/source/elements/shell/coque/cbilan.F:183-275      4-node shells -> PARTSAV(1:6,part)
 *   C3BILAN engine/source/elements/sh3n/coque3n/c3bilan.F:150-167, 282-300   3-node shells
 *   SBILAN  engine/source/elements/solid/solide/sbilan.F:110-157      bricks (mean of the squared nodal velocities)
 *   ECRIT   engine/source/output/ecrit.F:178-240, 322-352             global line of the listing:
 *           ENCIN = sum 1/2 MS |V + DT1/2 A|^2 with A after every kinematic condition, V = V(n-1/2);
 *           ENROT alike with IN, VR, AR; ENINT = sum PARTSAV(1,:); momenta and mass from the nodes
 *   FIXVEL  engine/source/constraints/general/impvel/fixvel.F:391-394, 834  work of the imposed velocities -> WFEXT
 * The element routines call the hooks where the reference calls its xBILAN (czforc3.F:639, cforc3.F:648,
 * c3forc3.F:616, sforc3.F:1436): after the material law, before the hourglass / time-step parts. */
#include "oracle.h"
#include <cstring>

static inline void psav_add(Oracle& o,int part,const double c[6])
{
  double* p=&o.PARTSAV[(size_t)6*part];
  #pragma omp critical(orc_partsav)
  { for(int k=0;k<6;k++) p[k]=p[k]+c[k]; }
}

/* shells: nn = 4 (CBILAN) or 3 (C3BILAN); nodes 0-based */
void orc_bilan_shell(Oracle& o,int elem,int nn,const int* nodes,double eint1,double eint2,double rho,double off)
{
  const int part = nn==4 ? (o.IPARTC.empty()?0:o.IPARTC[elem]) : (o.IPARTTG.empty()?0:o.IPARTTG[elem]);
  const double gvol = nn==4 ? o.GVOLC[elem] : o.GVOLTG[elem];
  const double* V=o.V.data();
  double vxa=K_ZERO,vya=K_ZERO,vza=K_ZERO;
  for(int k=0;k<nn;k++){ vxa=vxa+V[3*nodes[k]]; vya=vya+V[3*nodes[k]+1]; vza=vza+V[3*nodes[k]+2]; }
  double va2=K_ZERO;
  for(int c=0;c<3;c++) for(int k=0;k<nn;k++){ const double v=V[3*nodes[k]+c]; va2=va2+v*v; }
  const double xmas=rho*gvol;
  const double ei=eint1+eint2;
  const double ek= nn==4 ? xmas*va2*K_ONE_OVER_8 : xmas*va2*K_ONE_OVER_6;
  const double xmas25= nn==4 ? xmas*K_FOURTH : xmas*K_THIRD;
  const double c[6]={ei,ek,xmas25*vxa,xmas25*vya,xmas25*vza, off!=K_ZERO? xmas : K_ZERO};
  psav_add(o,part,c);
}

/* bricks: v = the nodal velocities the force routine gathered (zeroed for a dying element) */
void orc_bilan_solid(Oracle& o,int elem,const double vx[8],const double vy[8],const double vz[8],
                     double eint,double vol,double rho,double vnew,double off)
{
  const int part = o.IPARTS.empty()?0:o.IPARTS[elem];
  double vxa=vx[0]+vx[1]+vx[2]+vx[3]+vx[4]+vx[5]+vx[6]+vx[7];
  double vya=vy[0]+vy[1]+vy[2]+vy[3]+vy[4]+vy[5]+vy[6]+vy[7];
  double vza=vz[0]+vz[1]+vz[2]+vz[3]+vz[4]+vz[5]+vz[6]+vz[7];
  double va2=vx[0]*vx[0]+vx[1]*vx[1]+vx[2]*vx[2]+vx[3]*vx[3]+vx[4]*vx[4]+vx[5]*vx[5]+vx[6]*vx[6]+vx[7]*vx[7]
            +vy[0]*vy[0]+vy[1]*vy[1]+vy[2]*vy[2]+vy[3]*vy[3]+vy[4]*vy[4]+vy[5]*vy[5]+vy[6]*vy[6]+vy[7]*vy[7]
            +vz[0]*vz[0]+vz[1]*vz[1]+vz[2]*vz[2]+vz[3]*vz[3]+vz[4]*vz[4]+vz[5]*vz[5]+vz[6]*vz[6]+vz[7]*vz[7];
  vxa=vxa*K_ONE_OVER_8; vya=vya*K_ONE_OVER_8; vza=vza*K_ONE_OVER_8; va2=va2*K_ONE_OVER_8;
  const double xmas=rho*vnew;                   /* FILL = 1 */
  const double c[6]={eint*vol, xmas*va2*K_HALF, xmas*vxa, xmas*vya, xmas*vza, off>=K_ONE? xmas : K_ZERO};
  psav_add(o,part,c);
}

/* ECRIT, called between the kinematic conditions and VELOCITY (sortie_main at resol.F:8523) */
void orc_ecrit(Oracle& o)
{
  const int n=o.numnod; const double DT05=K_HALF*o.DT1;
  double encin=K_ZERO,enrot=K_ZERO,xm=K_ZERO,ym=K_ZERO,zm=K_ZERO,mass=K_ZERO;
  for(int I=0;I<n;I++){
    const double MAS=o.MS[I];
    const double VX=o.V[3*I]+DT05*o.A[3*I], VY=o.V[3*I+1]+DT05*o.A[3*I+1], VZ=o.V[3*I+2]+DT05*o.A[3*I+2];
    encin=encin+(VX*VX+VY*VY+VZ*VZ)*K_HALF*MAS;
    xm=xm+VX*MAS; ym=ym+VY*MAS; zm=zm+VZ*MAS; mass=mass+MAS;
  }
  if(o.ctl.iroddl!=0){
    for(int I=0;I<n;I++){
      const double VX=o.VR[3*I]+DT05*o.AR[3*I], VY=o.VR[3*I+1]+DT05*o.AR[3*I+1], VZ=o.VR[3*I+2]+DT05*o.AR[3*I+2];
      enrot=enrot+(VX*VX+VY*VY+VZ*VZ)*K_HALF*o.IN[I];
    }
  }
  double enint=K_ZERO;
  for(int m=0;m<o.npart;m++) enint=enint+o.PARTSAV[(size_t)6*m];
  o.ENCIN=encin; o.ENROT=enrot; o.ENINT=enint; o.XMOMT=xm; o.YMOMT=ym; o.ZMOMT=zm; o.XMASS=mass;
}

extern "C" {
/* part (0-based) of every element and GBUF%VOL of the shells (initial area x thickness, Starter output) */
void orc_set_parts(void* h,int npart,const int* ipartc,const int* iparts,const int* iparttg,const double* gvolc,const double* gvoltg)
{
  Oracle* o=(Oracle*)h; o->npart=npart>0?npart:1;
  if(ipartc) o->IPARTC.assign(ipartc,ipartc+o->numelc);
  if(iparts) o->IPARTS.assign(iparts,iparts+o->numels);
  if(iparttg) o->IPARTTG.assign(iparttg,iparttg+o->numeltg);
  if(gvolc) o->GVOLC.assign(gvolc,gvolc+o->numelc);
  if(gvoltg) o->GVOLTG.assign(gvoltg,gvoltg+o->numeltg);
  o->PARTSAV.assign((size_t)6*o->npart,0.0);
}
void orc_set_print(void* h,int ipri){ Oracle* o=(Oracle*)h; o->ipri=ipri; if(o->PARTSAV.empty()) o->PARTSAV.assign((size_t)6*o->npart,0.0); }
/* out: ENCIN, ENROT, ENINT, WFEXT, XMOMT, YMOMT, ZMOMT, XMASS of the last cycle; partsav (6,npart) */
void orc_get_balance(void* h,double* out,double* partsav){
  Oracle* o=(Oracle*)h;
  out[0]=o->ENCIN; out[1]=o->ENROT; out[2]=o->ENINT; out[3]=o->WFEXT; out[4]=o->XMOMT; out[5]=o->YMOMT; out[6]=o->ZMOMT; out[7]=o->XMASS;
  if(partsav) memcpy(partsav,o->PARTSAV.data(),sizeof(double)*6*(size_t)o->npart);
}
}
