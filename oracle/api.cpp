/* oracle/api.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).  C entry points used through
 * ctypes by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.  The call
 * surface mirrors include/orgpu.h one-to-one (orc_* vs orgpu_*) so the parity tests drive
 * both sides with the same arrays. */
#include "shell.h"
#include <cstring>
#include <cstdio>
#include <omp.h>

void orc_shell_group_free(OrcShellGroup*);
OrcShellGroup* orc_shell_group_new(int nel,int nft,int law,const void* mat,const orgpu_prop_shell* prop);
void orc_shell_group_state(const OrcShellGroup& g,int field,size_t ne,double* out);
void orc_forces(Oracle& o);
void orc_shell_group_state_up(OrcShellGroup& g,int field,size_t ne,const double* in);

extern "C" {

void* orc_create(int numnod, const orgpu_control* ctl)
{
  Oracle* o=new Oracle();
  o->numnod=numnod; o->ctl=*ctl;
  size_t n=numnod;
  o->X.assign(3*n,0); o->V.assign(3*n,0); o->VR.assign(3*n,0); o->D.assign(3*n,0); o->DR.assign(3*n,0);
  o->A.assign(3*n,0); o->AR.assign(3*n,0); o->MS.assign(n,0); o->IN.assign(n,0);
  o->STIFN.assign(n,0); o->STIFR.assign(n,0);
  o->TT=ctl->tt_init; o->DT2=ctl->dt_init; o->DT2OLD=ctl->dt2old_init; o->DT1=0; o->DT12=0;
  return o;
}
void orc_destroy(void* h){
  Oracle* o=(Oracle*)h;
  for(auto* g:o->cgroups) orc_shell_group_free(g);
  for(auto* g:o->tgroups) orc_shell_group_free(g);
  delete o;
}
void orc_set_threads(void* h,int nt){ ((Oracle*)h)->nthreads = nt>0? nt : omp_get_max_threads(); }

/* nodal arrays, Fortran (3,NUMNOD); any pointer may be NULL (left unchanged) */
void orc_upload_nodes(void* h,const double* X,const double* V,const double* VR,const double* D,
                      const double* MS,const double* IN)
{
  Oracle* o=(Oracle*)h; size_t n=o->numnod;
  if(X)  memcpy(o->X.data(),X,24*n);
  if(V)  memcpy(o->V.data(),V,24*n);
  if(VR) memcpy(o->VR.data(),VR,24*n);
  if(D)  memcpy(o->D.data(),D,24*n);
  if(MS) memcpy(o->MS.data(),MS,8*n);
  if(IN) memcpy(o->IN.data(),IN,8*n);
}
void orc_set_loads(void* h,const double* FEXT,const double* MEXT){
  Oracle* o=(Oracle*)h; size_t n=o->numnod;
  if(FEXT) o->FEXT.assign(FEXT,FEXT+3*n); else o->FEXT.clear();
  if(MEXT) o->MEXT.assign(MEXT,MEXT+3*n); else o->MEXT.clear();
}
void orc_set_bcs(void* h,const int* icodt,const int* icodr){
  Oracle* o=(Oracle*)h; size_t n=o->numnod;
  o->ICODT.assign(icodt,icodt+n);
  if(icodr) o->ICODR.assign(icodr,icodr+n); else o->ICODR.assign(n,0);
}
void orc_set_cloads(void* h,int nload,const int* ib,const double* fac){
  Oracle* o=(Oracle*)h; o->CL_IB.assign(ib,ib+(size_t)3*nload); o->CL_FAC.assign(fac,fac+(size_t)2*nload);
}
void orc_set_parith(void* h,int iparit){ ((Oracle*)h)->iparit=iparit; }   /* IPARIT: where FORCE's records enter the nodal sum (oracle.h) */
void orc_set_load_function(void* h,int ifunc,double fcx){ Oracle* o=(Oracle*)h; o->LF_FUNC=ifunc; o->LF_FCX=fcx; }
void orc_set_fixvel(void* h,int nfxvel,const int* ibfv /*(3,n)*/,const double* vel /*(4,n)*/){
  Oracle* o=(Oracle*)h; o->IBFV.assign(ibfv,ibfv+(size_t)3*nfxvel); o->VEL.assign(vel,vel+(size_t)4*nfxvel);
}
void orc_set_gravity(void* h,int ngrav,const int* igrv /*(3,n)*/,const double* agrv /*(2,n)*/,const int* ib,int lib){
  Oracle* o=(Oracle*)h; o->IGRV.assign(igrv,igrv+(size_t)3*ngrav); o->AGRV.assign(agrv,agrv+(size_t)2*ngrav); o->IBGRV.assign(ib,ib+lib);
}
/* connectivity + /PARITH/ON tables (all 1-based, Fortran layout) */
void orc_set_solids(void* h,int numels,const int* ixs /*(11,numels)*/,const int* iads /*(8,numels)*/){
  Oracle* o=(Oracle*)h; o->numels=numels;
  o->IXS.assign(ixs,ixs+(size_t)11*numels); o->IADS.assign(iads,iads+(size_t)8*numels);
}
void orc_set_shells(void* h,int numelc,const int* ixc /*(7,numelc)*/,const int* iadc /*(4,numelc)*/){
  Oracle* o=(Oracle*)h; o->numelc=numelc;
  o->IXC.assign(ixc,ixc+(size_t)7*numelc); o->IADC.assign(iadc,iadc+(size_t)4*numelc);
}
void orc_set_sh3n(void* h,int numeltg,const int* ixtg /*(6,numeltg)*/,const int* iadtg /*(3,numeltg)*/){
  Oracle* o=(Oracle*)h; o->numeltg=numeltg;
  o->IXTG.assign(ixtg,ixtg+(size_t)6*numeltg); o->IADTG.assign(iadtg,iadtg+(size_t)3*numeltg);
}
void orc_set_pon(void* h,const int* adsky /*numnod+1*/,int lsky){
  Oracle* o=(Oracle*)h; o->ADSKY.assign(adsky,adsky+o->numnod+1); o->lsky=lsky;
  o->FSKY.assign((size_t)8*lsky,0.0);
}
void orc_set_functions(void* h,int nfunc,const int* npf /*nfunc+1*/,const double* tf){
  Oracle* o=(Oracle*)h; o->NPF.assign(npf,npf+nfunc+1); o->TF.assign(tf,tf+(size_t)2*npf[nfunc]);
}

/* one solid group: elements [nft, nft+nel) of IXS; vol0 = initial volumes (Starter output) */
int orc_add_solid_group(void* h,int nel,int nft,const orgpu_law2* mat,const orgpu_prop_solid* prop,
                        const double* vol0)
{
  Oracle* o=(Oracle*)h;
  if(nel>MVSIZ-1) return -1;
  OrcSolidGroup g; g.nel=nel; g.nft=nft; g.mat=*mat; g.prop=*prop;
  if(mat->fisokin>0.0) g.sigb.assign(6*nel,0);             /* LBUF%SIGB (m2law.F:181-190, 364-390) */
  g.sig.assign(6*nel,0); g.eint.assign(nel,0); g.rho.assign(nel,mat->rho0); g.qvis.assign(nel,0);
  g.pla.assign(nel,0); g.epsd.assign(nel,0); g.vol.assign(vol0,vol0+nel); g.off.assign(nel,1.0);
  g.temp.assign(nel,mat->tini); g.dmg.assign(nel,0); g.smstr.assign(21*nel,0);
  o->sgroups.push_back(std::move(g));
  return (int)o->sgroups.size()-1;
}

/* same, any built law: law = 2 (orgpu_law2) or 36 (orgpu_law36: MMAIN -> MULAW -> SIGEPS36) */
int orc_add_solid_group_law(void* h,int nel,int nft,int law,const void* mat,const orgpu_prop_solid* prop,
                            const double* vol0)
{
  if(law==2) return orc_add_solid_group(h,nel,nft,(const orgpu_law2*)mat,prop,vol0);
  if(law!=36) return -3;
  Oracle* o=(Oracle*)h;
  if(nel>MVSIZ-1) return -1;
  const orgpu_law36* m=(const orgpu_law36*)mat;
  if(m->fisokin!=0.0 || m->vp!=0 || m->ifail<0 || m->ifail>2) return -2;
  if(m->ifail==2 && prop->istrain==0) return -4;
  OrcSolidGroup g; g.nel=nel; g.nft=nft; g.law=36; g.m36=*m; g.prop=*prop;
  g.mat=orgpu_law2{}; g.mat.rho0=m->rho0;
  g.sig.assign(6*nel,0); g.eint.assign(nel,0); g.rho.assign(nel,m->rho0); g.qvis.assign(nel,0);
  g.pla.assign(nel,0); g.epsd.assign(nel,0); g.vol.assign(vol0,vol0+nel); g.off.assign(nel,1.0);
  g.temp.assign(nel,0); g.dmg.assign(nel,0); g.smstr.assign(21*nel,0);
  g.stra.assign(6*nel,0); g.wpla.assign(nel,0); g.vartmp.assign((size_t)(2+m->nrate)*nel,0);
  o->sgroups.push_back(std::move(g));
  return (int)o->sgroups.size()-1;
}

/* one shell group: elements [nft, nft+nel) of IXC; law = 2 (orgpu_law2) or 36 (orgpu_law36) */
int orc_add_shell_group(void* h,int nel,int nft,int law,const void* mat,const orgpu_prop_shell* prop)
{
  Oracle* o=(Oracle*)h;
  if(nel>MVSIZ-1) return -1;
  if(law!=2 && law!=36) return -2;
  if(prop->npt<1 || prop->npt>10) return -3;
  if(law==36){ const orgpu_law36* m=(const orgpu_law36*)mat; if(m->vp<0 || m->vp>1 || (m->vp==1 && m->nrate<2)) return -5; }   /* hm_read_mat36.F:199 */
  o->cgroups.push_back(orc_shell_group_new(nel,nft,law,mat,prop));
  return (int)o->cgroups.size()-1;
}

/* one 3-node shell group (ITY=7): elements [nft, nft+nel) of IXTG; prop->ihbe carries Ish3n (1, 2) */
int orc_add_sh3n_group(void* h,int nel,int nft,int law,const void* mat,const orgpu_prop_shell* prop)
{
  Oracle* o=(Oracle*)h;
  if(nel>MVSIZ-1) return -1;
  if(law!=2 && law!=36) return -2;
  if(prop->npt<1 || prop->npt>10) return -3;
  if(prop->ihbe!=1 && prop->ihbe!=2) return -4;
  if(law==36){ const orgpu_law36* m=(const orgpu_law36*)mat; if(m->vp<0 || m->vp>1 || (m->vp==1 && m->nrate<2)) return -5; }
  OrcShellGroup* g=orc_shell_group_new(nel,nft,law,mat,prop);
  g->nhourg=0; g->HOURG.clear(); g->SMSTR.assign(3*nel,0);
  o->tgroups.push_back(g);
  return (int)o->tgroups.size()-1;
}

/* /FAIL/JOHNSON for one LAW2 solid group (mirror of orgpu_set_solid_group_fail) */
int orc_set_solid_group_fail(void* h,int group,const orgpu_fail* f)
{
  Oracle* o=(Oracle*)h;
  if(group<0 || group>=(int)o->sgroups.size()) return -1;
  if(f->irupt!=0 && (f->irupt!=1 || f->d5!=0.0 || o->sgroups[group].law!=2)) return -2;
  o->sgroups[group].fail=*f; o->sgroups[group].dfmax.assign(o->sgroups[group].nel,0.0);
  return 0;
}

/* /FAIL/JOHNSON for one shell group (mirror of orgpu_set_shell_group_fail) */
int orc_set_shell_group_fail(void* h,int sh3n,int group,const orgpu_fail* f)
{
  Oracle* o=(Oracle*)h;
  auto& gs = sh3n? o->tgroups : o->cgroups;
  if(group<0 || group>=(int)gs.size()) return -1;
  if(f->irupt!=0 && (f->irupt!=1 || f->d5!=0.0)) return -2;
  gs[group]->fail=*f;
  return 0;
}

void orc_finalize(void*){}

/* phases (same split as the device library) */
void orc_forces_phase(void* h,double dt1){
  Oracle* o=(Oracle*)h; o->DT1=dt1; o->DT2=K_EP06;
  if(o->ipri) std::fill(o->PARTSAV.begin(),o->PARTSAV.end(),K_ZERO);
  orc_forces(*o);
}
void orc_assemble(void* h){ Oracle* o=(Oracle*)h; orc_asspar4(*o); if(o->ctl.nodadt!=0) orc_dtnoda(*o); }
void orc_set_itab(void* h,const int* itab){ Oracle* o=(Oracle*)h; o->ITAB.assign(itab,itab+o->numnod); }
void orc_advance(void* h,double dt12,double dt2){
  Oracle* o=(Oracle*)h; o->DT12=dt12; o->DT2=dt2;
  orc_accele(*o); orc_gravit(*o); orc_bcs(*o); orc_fixvel(*o);
  if(o->ipri) orc_ecrit(*o);
  orc_velocity(*o); orc_depla(*o); o->TT+=dt2; o->NCYCLE++;
}
void orc_run_cycles(void* h,int ncycles){ Oracle* o=(Oracle*)h; for(int c=0;c<ncycles;c++) orc_cycle(*o); }

/* read-back */
void orc_get_time(void* h,double* out /*tt,dt1,dt2,dt12,dt2t*/,int* iout /*neltst,ityptst,ncycle*/){
  Oracle* o=(Oracle*)h;
  out[0]=o->TT; out[1]=o->DT1; out[2]=o->DT2; out[3]=o->DT12; out[4]=o->DT2T;
  iout[0]=o->NELTST; iout[1]=o->ITYPTST; iout[2]=(int)o->NCYCLE;
}
void orc_download_nodes(void* h,double* X,double* V,double* VR,double* D,double* A,double* AR,
                        double* STIFN,double* STIFR){
  Oracle* o=(Oracle*)h; size_t n=o->numnod;
  if(X) memcpy(X,o->X.data(),24*n); if(V) memcpy(V,o->V.data(),24*n); if(VR) memcpy(VR,o->VR.data(),24*n);
  if(D) memcpy(D,o->D.data(),24*n); if(A) memcpy(A,o->A.data(),24*n); if(AR) memcpy(AR,o->AR.data(),24*n);
  if(STIFN) memcpy(STIFN,o->STIFN.data(),8*n); if(STIFR) memcpy(STIFR,o->STIFR.data(),8*n);
}
void orc_download_fsky(void* h,double* fsky){ Oracle* o=(Oracle*)h; memcpy(fsky,o->FSKY.data(),64*(size_t)o->lsky); }

/* solid state of all groups concatenated in element order, component-major over NUMELS:
 * sig[k*numels+e] ; fields: 0 sig(6) 1 eint 2 rho 3 qvis 4 pla 5 epsd 6 vol 7 off 8 temp 9 smstr(21) 10 stra(6) 11 wpla 12 sigb(6) 13 dfmax */
void orc_download_solid_state(void* h,int field,double* out){
  Oracle* o=(Oracle*)h; size_t ne=o->numels;
  for(auto& g:o->sgroups){
    auto cp=[&](const std::vector<double>& v,int nc){ for(int k=0;k<nc;k++) for(int i=0;i<g.nel;i++) out[k*ne+g.nft+i]=v[(size_t)k*g.nel+i]; };
    switch(field){
      case 0: cp(g.sig,6); break; case 1: cp(g.eint,1); break; case 2: cp(g.rho,1); break;
      case 3: cp(g.qvis,1); break; case 4: cp(g.pla,1); break; case 5: cp(g.epsd,1); break;
      case 6: cp(g.vol,1); break; case 7: cp(g.off,1); break; case 8: cp(g.temp,1); break;
      case 9: cp(g.smstr,21); break;
      case 10: if(!g.stra.empty()) cp(g.stra,6); break; case 11: if(!g.wpla.empty()) cp(g.wpla,1); break;
      case 12: if(!g.sigb.empty()) cp(g.sigb,6); break;
      case 13: if(!g.dfmax.empty()) cp(g.dfmax,1); break;
    }
  }
}

/* the reverse (restart / initial state hand-over, /INIBRI, /INISHE): same fields, same layout */
void orc_upload_solid_state(void* h,int field,const double* in){
  Oracle* o=(Oracle*)h; size_t ne=o->numels;
  for(auto& g:o->sgroups){
    auto cp=[&](std::vector<double>& v,int nc){ if(v.empty()) return; for(int k=0;k<nc;k++) for(int i=0;i<g.nel;i++) v[(size_t)k*g.nel+i]=in[k*ne+g.nft+i]; };
    switch(field){
      case 0: cp(g.sig,6); break; case 1: cp(g.eint,1); break; case 2: cp(g.rho,1); break;
      case 3: cp(g.qvis,1); break; case 4: cp(g.pla,1); break; case 5: cp(g.epsd,1); break;
      case 6: cp(g.vol,1); break; case 7: cp(g.off,1); break; case 8: cp(g.temp,1); break;
      case 9: cp(g.smstr,21); break; case 10: cp(g.stra,6); break; case 11: cp(g.wpla,1); break; case 12: cp(g.sigb,6); break; case 13: cp(g.dfmax,1); break;
    }
  }
}
void orc_upload_shell_state(void* h,int field,const double* in){
  Oracle* o=(Oracle*)h;
  for(auto* g:o->cgroups) orc_shell_group_state_up(*g,field,(size_t)o->numelc,in);
}
void orc_upload_sh3n_state(void* h,int field,const double* in){
  Oracle* o=(Oracle*)h;
  for(auto* g:o->tgroups){
    if(field==7) continue;
    if(field==8){ for(int k=0;k<3;k++) for(int i=0;i<g->nel;i++) g->SMSTR[(size_t)k*g->nel+i]=in[(size_t)k*o->numeltg+g->nft+i]; continue; }
    orc_shell_group_state_up(*g,field,(size_t)o->numeltg,in);
  }
}
void orc_set_time(void* h,double tt,double dt2,double dt2old,long long ncycle){
  Oracle* o=(Oracle*)h; o->TT=tt; o->DT2=dt2; o->DT2OLD=dt2old; o->NCYCLE=(long)ncycle;
}

/* corner rows of n FSKY slots (0-based) out of / into the skyline: what SPMD_EXCH2_A_PON packs
 * (spmd_exch2_a_pon.F:545-557) and unpacks (:1190-1201); rows are 8 doubles */
void orc_pack_rows(void* h,int n,const int* slots,double* buf){
  Oracle* o=(Oracle*)h;
  for(int j=0;j<n;j++) memcpy(buf+8*(size_t)j,&o->FSKY[8*(size_t)slots[j]],64);
}
void orc_unpack_rows(void* h,int n,const int* slots,const double* buf){
  Oracle* o=(Oracle*)h;
  for(int j=0;j<n;j++) memcpy(&o->FSKY[8*(size_t)slots[j]],buf+8*(size_t)j,64);
}

/* /PARITH/OFF exchange of nodal partial sums, SPMD_EXCH_A (engine/source/mpi/forces/spmd_exch_a.F): pack :153-166 (IRODDL/=0:
 * A(1:3), AR(1:3), STIFN, STIFR of the frontier nodes FR_ELEM shared with one neighbour), add :517-528 (neighbours in rank
 * order, nodes in list order).  nodes are 0-based; buf is (8,n). */
void orc_pack_nodes(void* h,int n,const int* nodes,double* buf){
  Oracle* o=(Oracle*)h;
  for(int j=0;j<n;j++){ const int N=nodes[j]; double* b=buf+8*(size_t)j;
    b[0]=o->A[3*N]; b[1]=o->A[3*N+1]; b[2]=o->A[3*N+2]; b[3]=o->AR[3*N]; b[4]=o->AR[3*N+1]; b[5]=o->AR[3*N+2]; b[6]=o->STIFN[N]; b[7]=o->STIFR[N]; }
}
void orc_add_nodes(void* h,int n,const int* nodes,const double* buf){
  Oracle* o=(Oracle*)h;
  for(int j=0;j<n;j++){ const int N=nodes[j]; const double* b=buf+8*(size_t)j;
    o->A[3*N]=o->A[3*N]+b[0]; o->A[3*N+1]=o->A[3*N+1]+b[1]; o->A[3*N+2]=o->A[3*N+2]+b[2];
    o->AR[3*N]=o->AR[3*N]+b[3]; o->AR[3*N+1]=o->AR[3*N+1]+b[4]; o->AR[3*N+2]=o->AR[3*N+2]+b[5];
    o->STIFN[N]=o->STIFN[N]+b[6]; o->STIFR[N]=o->STIFR[N]+b[7]; }
}

void orc_download_shell_state(void* h,int field,double* out){
  Oracle* o=(Oracle*)h;
  for(auto* g:o->cgroups) orc_shell_group_state(*g,field,(size_t)o->numelc,out);
}

/* 3-node shells: same fields as the 4-node shells (7 hourg: none, 8 smstr(3)); out[k*numeltg+e] */
void orc_download_sh3n_state(void* h,int field,double* out){
  Oracle* o=(Oracle*)h;
  for(auto* g:o->tgroups){
    if(field==7) continue;
    if(field==8){ for(int k=0;k<3;k++) for(int i=0;i<g->nel;i++) out[(size_t)k*o->numeltg+g->nft+i]=g->SMSTR[(size_t)k*g->nel+i]; continue; }
    orc_shell_group_state(*g,field,(size_t)o->numeltg,out);
  }
}

} // extern "C"
