/* oracle/shell.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).  Shell group bookkeeping: what
 * FORINTC (engine/source/elements/forintc.F:238-483) decodes from IPARG to pick CFORC3 (JHBE<11)
 * or CZFORC3 (JHBE 21..29), and the ELBUF allocation the Starter does for /PROP/SHELL groups. */
#include "shell.h"
#include <cstdlib>
#include <cstdio>

void orc_shell_group_free(OrcShellGroup* g){ delete g; }

void orc_shell_dispatch(Oracle& o, OrcShellGroup& g, double& dt2t, int& neltst, int& ityptst)
{
  if(g.prop.ihbe>=21 && g.prop.ihbe<=29) orc_czforc3(o,g,dt2t,neltst,ityptst);   /* forintc.F:383-385 */
  else orc_cforc3(o,g,dt2t,neltst,ityptst);                                       /* forintc.F:452    */
}

OrcShellGroup* orc_shell_group_new(int nel,int nft,int law,const void* mat,const orgpu_prop_shell* prop)
{
  OrcShellGroup* g=new OrcShellGroup();
  g->nel=nel; g->nft=nft; g->law=law; g->prop=*prop;
  if(law==36){ g->m36=*(const orgpu_law36*)mat; g->nvartmp=2+g->m36.nrate; }
  else       { g->m2=*(const orgpu_law2*)mat; g->nvartmp=0; }
  const bool qeph=(prop->ihbe>=21&&prop->ihbe<=29);
  g->nhourg= qeph? 12 : 5;
  g->FOR.assign(5*nel,0); g->MOM.assign(3*nel,0); g->EINT.assign(2*nel,0);
  g->THK.assign(nel,prop->thick); g->THKE.assign(nel,prop->thick); g->OFF.assign(nel,1.0);
  g->STRA.assign(8*nel,0); g->EPSD.assign(nel,0); g->HOURG.assign((size_t)g->nhourg*nel,0); g->SMSTR.assign(6*nel,0);
  g->ip.resize(prop->npt);
  for(auto& lb:g->ip){
    lb.sig.assign(5*nel,0); lb.pla.assign(nel,0); lb.epsd.assign(nel,0);
    lb.temp.assign(nel, law==2? g->m2.tini : 0.0); lb.sigb.assign(3*nel,0); lb.off.assign(nel,1.0);
    lb.dfmax.assign(nel,0.0); lb.foff.assign(nel,1.0); lb.plap.assign(nel,0.0);
    lb.vartmp.assign((size_t)(g->nvartmp>0?g->nvartmp:1)*nel,0);
  }
  return g;
}

/* fields: 0 for(5) 1 mom(3) 2 eint(2) 3 thk 4 off 5 stra(8) 6 epsd 7 hourg(12) 8 smstr(6)
 *         9 sig(5*npt) 10 pla(npt) 11 epsd_ip(npt) 12 temp(npt) ; out[k*numelc+e] */
void orc_shell_group_state(const OrcShellGroup& g,int field,size_t ne,double* out)
{
  auto cp=[&](const std::vector<double>& v,int nc,int k0){ for(int k=0;k<nc;k++) for(int i=0;i<g.nel;i++) out[(size_t)(k0+k)*ne+g.nft+i]=v[(size_t)k*g.nel+i]; };
  switch(field){
    case 0: cp(g.FOR,5,0); break; case 1: cp(g.MOM,3,0); break; case 2: cp(g.EINT,2,0); break;
    case 3: cp(g.THK,1,0); break; case 4: cp(g.OFF,1,0); break; case 5: cp(g.STRA,8,0); break;
    case 6: cp(g.EPSD,1,0); break; case 7: cp(g.HOURG,g.nhourg,0); break; case 8: cp(g.SMSTR,6,0); break;
    case 9:  for(size_t p=0;p<g.ip.size();p++) cp(g.ip[p].sig,5,5*(int)p); break;
    case 10: for(size_t p=0;p<g.ip.size();p++) cp(g.ip[p].pla,1,(int)p); break;
    case 11: for(size_t p=0;p<g.ip.size();p++) cp(g.ip[p].epsd,1,(int)p); break;
    case 12: for(size_t p=0;p<g.ip.size();p++) cp(g.ip[p].temp,1,(int)p); break;
    case 13: for(size_t p=0;p<g.ip.size();p++) cp(g.ip[p].sigb,3,3*(int)p); break;   /* LBUF%SIGB (kinematic hardening) */
    case 14: for(size_t p=0;p<g.ip.size();p++) cp(g.ip[p].dfmax,1,(int)p); break;     /* /FAIL/JOHNSON damage */
    case 15: for(size_t p=0;p<g.ip.size();p++) cp(g.ip[p].foff,1,(int)p); break;      /* ... and point flag */
    case 16: for(size_t p=0;p<g.ip.size();p++) cp(g.ip[p].plap,1,(int)p); break;      /* LAW36 VP = 1: UVAR(2), filtered plastic strain rate */
  }
}

/* same fields the other way (restart / initial state) */
void orc_shell_group_state_up(OrcShellGroup& g,int field,size_t ne,const double* in)
{
  auto cp=[&](std::vector<double>& v,int nc,int k0){ for(int k=0;k<nc;k++) for(int i=0;i<g.nel;i++) v[(size_t)k*g.nel+i]=in[(size_t)(k0+k)*ne+g.nft+i]; };
  switch(field){
    case 0: cp(g.FOR,5,0); break; case 1: cp(g.MOM,3,0); break; case 2: cp(g.EINT,2,0); break;
    case 3: cp(g.THK,1,0); break; case 4: cp(g.OFF,1,0); break; case 5: cp(g.STRA,8,0); break;
    case 6: cp(g.EPSD,1,0); break; case 7: cp(g.HOURG,g.nhourg,0); break; case 8: cp(g.SMSTR,6,0); break;
    case 9:  for(size_t p=0;p<g.ip.size();p++) cp(g.ip[p].sig,5,5*(int)p); break;
    case 10: for(size_t p=0;p<g.ip.size();p++) cp(g.ip[p].pla,1,(int)p); break;
    case 11: for(size_t p=0;p<g.ip.size();p++) cp(g.ip[p].epsd,1,(int)p); break;
    case 12: for(size_t p=0;p<g.ip.size();p++) cp(g.ip[p].temp,1,(int)p); break;
    case 13: for(size_t p=0;p<g.ip.size();p++) cp(g.ip[p].sigb,3,3*(int)p); break;
    case 14: for(size_t p=0;p<g.ip.size();p++) cp(g.ip[p].dfmax,1,(int)p); break;
    case 15: for(size_t p=0;p<g.ip.size();p++) cp(g.ip[p].foff,1,(int)p); break;
    case 16: for(size_t p=0;p<g.ip.size();p++) cp(g.ip[p].plap,1,(int)p); break;
  }
}
