// Design experiment behind the list traversal kernel (csrc/traverse.cu, traverse_f32_list_kernel): how uniform are the
// per-body accept / open decisions inside a warp's group of G Morton-consecutive bodies, and how sharp is a
// bounding-box test of the whole group?  CPU only, no library code:
//     gcc -O2 -o classify_nodes tools/classify_nodes.c -lm && ./classify_nodes 1000000 64 [plummer]
// Builds the capped pyramid (cap 10) over a seeded uniform disk (or projected Plummer sphere), walks it with the
// reference's per-body semantics (project.cu:608-670) for sampled groups and prints, per group: nodes touched, nodes
// with a full body mask that every body accepts / opens / that are mixed, the same for partial masks, what a
// conservative box test decides, how many of the accepted nodes are "far" (dmin^2 >= diag^2 / 16 and / 64), and how
// little sub-group boxes (halves, quarters, eighths of the group) would help.
// experiment: how uniform are the per-lane decisions inside a warp group?
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#define F 9
static uint64_t s=88172645463325252ull; static double rnd(){ s^=s<<13; s^=s>>7; s^=s<<17; return (s>>11)*(1.0/9007199254740992.0);}
typedef struct {double m,cx,cy; uint32_t cnt;} Cell;
static Cell* lev[F+1];
static int cmpk(const void*a,const void*b){ uint64_t x=*(const uint64_t*)a,y=*(const uint64_t*)b; return x<y?-1:x>y;}
int main(int argc,char**argv){
  int N=argc>1?atoi(argv[1]):1000000; int G=argc>2?atoi(argv[2]):64; int plummer = argc>3?atoi(argv[3]):0;
  double *px=malloc(8*N),*py=malloc(8*N),*pm=malloc(8*N);
  for(int i=0;i<N;i++){
    if(!plummer){ double r=0.1*sqrt(rnd()),ph=2*M_PI*rnd(); px[i]=r*cos(ph);py[i]=r*sin(ph);}
    else { double r; do{ double u=rnd(); r=0.02/sqrt(pow(u,-2.0/3.0)-1);}while(r>0.1); double cz=2*rnd()-1,ph=2*M_PI*rnd(),sn=sqrt(1-cz*cz); px[i]=r*sn*cos(ph);py[i]=r*sn*sin(ph);}
    pm[i]=pow(10,-1+rnd()*(log10(0.5)+1)); }
  double xmin=1e9,xmax=-1e9,ymin=1e9,ymax=-1e9; for(int i=0;i<N;i++){ if(px[i]<xmin)xmin=px[i]; if(px[i]>xmax)xmax=px[i]; if(py[i]<ymin)ymin=py[i]; if(py[i]>ymax)ymax=py[i];}
  double pad=0.1*fmax(xmax-xmin,ymax-ymin); xmin-=pad;xmax+=pad;ymin-=pad;ymax+=pad;
  double W=xmax-xmin,H=ymax-ymin;
  uint64_t* ks=malloc(8*N);
  for(int i=0;i<N;i++){ uint32_t ix=(uint32_t)((px[i]-xmin)/W*512),iy=(uint32_t)((py[i]-ymin)/H*512); uint32_t k=0; for(int b=0;b<9;b++){k|=((ix>>b)&1)<<(2*b); k|=((iy>>b)&1)<<(2*b+1);} ks[i]=((uint64_t)k<<32)|i; }
  qsort(ks,N,8,cmpk);
  for(int l=0;l<=F;l++) lev[l]=calloc((size_t)1<<(2*l),sizeof(Cell));
  for(int j=0;j<N;j++){ uint32_t k=ks[j]>>32,i=(uint32_t)ks[j]; Cell*c=&lev[F][k]; c->m+=pm[i]; c->cx+=pm[i]*px[i]; c->cy+=pm[i]*py[i]; c->cnt++; }
  for(size_t k=0;k<(1u<<(2*F));k++){ Cell*c=&lev[F][k]; if(c->m>0){c->cx/=c->m;c->cy/=c->m;} }
  for(int l=F-1;l>=0;l--) for(size_t k=0;k<((size_t)1<<(2*l));k++){ Cell*c=&lev[l][k]; for(int q=0;q<4;q++){Cell*d=&lev[l+1][4*k+q]; c->m+=d->m;c->cx+=d->m*d->cx;c->cy+=d->m*d->cy;c->cnt+=d->cnt;} if(c->m>0){c->cx/=c->m;c->cy/=c->m;} }
  double size[F+1]; for(int l=0;l<=F;l++) size[l]=fmax(W,H)/(1<<l);
  // walk
  long pAbb=0,pObb=0,pMbb=0; long subM[4]={0,0,0,0}; long nearA=0, farA=0, nearB=0, farB=0; long U=0,V=0,Afull=0,Ofull=0,Mfull=0,Apart=0,Opart=0,Mpart=0,Z=0, Abb=0,Obb=0,Mbb=0, inter=0, iters=0, Aleaf=0;
  long nodesA_lvl[F+1]={0}, nodesM_lvl[F+1]={0}, nodesO_lvl[F+1]={0};
  int nw=0; int stride = (N/G)/2000; if(stride<1)stride=1;
  typedef struct{int l; uint32_t k; uint64_t mask;} SE; SE st[256];
  for(int g=0; g*G<N; g+=stride){ nw++;
    int j0=g*G, nb = (j0+G<=N)?G:N-j0; uint64_t live = nb==64?~0ull:((1ull<<nb)-1);
    double bx0=1e9,bx1=-1e9,by0=1e9,by1=-1e9; double bxs[64],bys[64]; uint32_t bk[64]; double sb[3][8][4];
    for(int b=0;b<nb;b++){ uint32_t i=(uint32_t)ks[j0+b]; bxs[b]=px[i];bys[b]=py[i]; bk[b]=ks[j0+b]>>32; if(bxs[b]<bx0)bx0=bxs[b]; if(bxs[b]>bx1)bx1=bxs[b]; if(bys[b]<by0)by0=bys[b]; if(bys[b]>by1)by1=bys[b]; }
    for(int qi=0;qi<3;qi++){ int Q=2<<qi; int per=64/Q; for(int q=0;q<Q;q++){ double a0=1e9,a1=-1e9,c0=1e9,c1=-1e9; for(int b=q*per;b<(q+1)*per&&b<nb;b++){ if(bxs[b]<a0)a0=bxs[b]; if(bxs[b]>a1)a1=bxs[b]; if(bys[b]<c0)c0=bys[b]; if(bys[b]>c1)c1=bys[b]; } sb[qi][q][0]=a0;sb[qi][q][1]=a1;sb[qi][q][2]=c0;sb[qi][q][3]=c1; } }
    int top=0; 
    // root handled as node at level 0: evaluate
    // generic evaluate function inline
    #define EVAL(L,K,MASK) do{ Cell*c=&lev[L][K]; U++; int nact=__builtin_popcountll(MASK); V+=nact; \
      if(!(c->m>1e-15)){Z++; break;} int leaf = (c->cnt<=1|| (L)==F); uint64_t om=0; int nacc=0; \
      for(int b=0;b<nb;b++) if((MASK>>b)&1){ double dx=c->cx-bxs[b],dy=c->cy-bys[b]; double d=sqrt(dx*dx+dy*dy)+1e-15; int acc = leaf || size[L]/d<0.5; if(acc){ nacc++; int self = leaf && c->cnt==1 && bk[b]>>(2*(F-(L)))==(K) ; if(!self) inter++; } else om|=1ull<<b; } \
      int full = (MASK==live); \
      if(om==0){ if(full){Afull++; nodesA_lvl[L]++; if(leaf)Aleaf++;} else Apart++; } else if(nacc==0){ if(full){Ofull++;nodesO_lvl[L]++;} else Opart++; } else { if(full){Mfull++;nodesM_lvl[L]++;} else Mpart++; } \
      /* bbox test */ { double ddx = c->cx<bx0?bx0-c->cx:(c->cx>bx1?c->cx-bx1:0), ddy=c->cy<by0?by0-c->cy:(c->cy>by1?c->cy-by1:0); double dmin=sqrt(ddx*ddx+ddy*ddy); double fx=fmax(fabs(c->cx-bx0),fabs(c->cx-bx1)),fy=fmax(fabs(c->cy-by0),fabs(c->cy-by1)); double dmax=sqrt(fx*fx+fy*fy); \
         if(leaf || size[L]/(dmin+1e-15)<0.5*0.9999) {if(full)Abb++; else pAbb++;} else if(size[L]/(dmax+1e-15)>=0.5*1.0001) {if(full)Obb++; else pObb++;} else {if(full)Mbb++; else pMbb++;} } \
      if(!leaf){ for(int qi=0;qi<3;qi++){ int Q=2<<qi; int per=64/Q; int anyM=0; for(int q=0;q<Q;q++){ uint64_t qm = (per==64?~0ull:(((1ull<<per)-1)<<(q*per))) & MASK; if(!qm) continue; double *B=sb[qi][q]; double ddx = c->cx<B[0]?B[0]-c->cx:(c->cx>B[1]?c->cx-B[1]:0), ddy=c->cy<B[2]?B[2]-c->cy:(c->cy>B[3]?c->cy-B[3]:0); double dmn=sqrt(ddx*ddx+ddy*ddy); double fx=fmax(fabs(c->cx-B[0]),fabs(c->cx-B[1])),fy=fmax(fabs(c->cy-B[2]),fabs(c->cy-B[3])); double dmx=sqrt(fx*fx+fy*fy); if(!(size[L]/(dmn+1e-15)<0.5*0.9999) && !(size[L]/(dmx+1e-15)>=0.5*1.0001)) anyM=1; } subM[qi]+=anyM; } } \
      { double ddx = c->cx<bx0?bx0-c->cx:(c->cx>bx1?c->cx-bx1:0), ddy=c->cy<by0?by0-c->cy:(c->cy>by1?c->cy-by1:0); double dmn2=ddx*ddx+ddy*ddy; double gd2=(bx1-bx0)*(bx1-bx0)+(by1-by0)*(by1-by0); if(om==0){ if(dmn2>=gd2/16) farA++; else nearA++; if(dmn2>=gd2/64) farB++; else nearB++; } } \
      if(om){ st[top].l=L; st[top].k=K; st[top].mask=om; top++; } }while(0)
    EVAL(0,0,live);
    while(top>0){ SE e=st[--top]; iters++; for(int q=0;q<4;q++){ int L=e.l+1; uint32_t K=4*e.k+q; uint64_t MK=e.mask; EVAL(L,K,MK); } }
  }
  printf("N=%d G=%d warps sampled %d\n",N,G,nw);
  printf("per warp: iterations %.1f  node evals (union) %.1f  active body-evals/body %.1f  interactions/body %.1f lane-eff %.3f\n",(double)iters/nw,(double)U/nw,(double)V/nw/G,(double)inter/nw/G,(double)V/((double)U*G));
  printf("per warp nodes: zero %.1f | full-mask: allaccept %.1f (leaf %.1f) allopen %.1f mixed %.1f | partial-mask: acc %.1f open %.1f mixed %.1f\n",(double)Z/nw,(double)Afull/nw,(double)Aleaf/nw,(double)Ofull/nw,(double)Mfull/nw,(double)Apart/nw,(double)Opart/nw,(double)Mpart/nw);
  printf("bbox test on full-mask nodes: A %.1f O %.1f M %.1f\n",(double)Abb/nw,(double)Obb/nw,(double)Mbb/nw);
  printf("M nodes with sub-group boxes: halves %.1f quarters %.1f eighths %.1f ; ideal-accept nodes far %.1f near %.1f (/64: far %.1f near %.1f)\n",(double)subM[0]/nw,(double)subM[1]/nw,(double)subM[2]/nw,(double)farA/nw,(double)nearA/nw,(double)farB/nw,(double)nearB/nw);
  printf("bbox test on partial-mask nodes: A %.1f O %.1f M %.1f\n",(double)pAbb/nw,(double)pObb/nw,(double)pMbb/nw);
  if(0)for(int l=0;l<=F;l++) printf(" L%d: A %.1f O %.1f M %.1f\n",l,(double)nodesA_lvl[l]/nw,(double)nodesO_lvl[l]/nw,(double)nodesM_lvl[l]/nw);
  return 0; }
