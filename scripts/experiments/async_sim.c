// experiment (CPU): work done by an ASYNCHRONOUS parallel push (K workers, pushes land `lat` steps after
// the pop) under different deferral policies.  Models csrc/push_async.cuh well enough to compare schedules.
//   policy 0: eager (pop anything > eps)
//   policy 1: global threshold levels theta_k = theta0 / rho^k with a deferred list (Policy A in the notes)
// gcc -O2 -o build/async_sim scripts/experiments/async_sim.c -lm
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define A 0.15
static int V; static int64_t M, W; static int directed; static int32_t *pairs;
static int32_t *in_ptr, *in_col, *deg; static double *p, *r;
static void build(int64_t pos){
  memset(in_ptr,0,sizeof(int32_t)*(V+1)); memset(deg,0,sizeof(int32_t)*V);
  for(int64_t i=pos-W;i<pos;++i){int a=pairs[2*i],b=pairs[2*i+1]; in_ptr[b+1]++; deg[a]++; if(!directed){in_ptr[a+1]++; deg[b]++;}}
  for(int u=0;u<V;++u) in_ptr[u+1]+=in_ptr[u];
  int32_t *f=malloc(sizeof(int32_t)*V); memcpy(f,in_ptr,sizeof(int32_t)*V);
  for(int64_t i=pos-W;i<pos;++i){int a=pairs[2*i],b=pairs[2*i+1]; in_col[f[b]++]=a; if(!directed) in_col[f[a]++]=b;}
  free(f);
}
static int64_t T,F,STEPS,LEVELS;
typedef struct { int v; double add; } Ev;
#define RING 256
static Ev *evb[RING]; static int evn[RING], evc[RING];
static unsigned rng=12345; static inline unsigned rnd(){rng=rng*1664525u+1013904223u; return rng>>8;}
static void ev_push(int t,int v,double add){int b=t%RING; if(evn[b]==evc[b]){evc[b]=evc[b]?evc[b]*2:1024; evb[b]=realloc(evb[b],sizeof(Ev)*evc[b]);} evb[b][evn[b]].v=v; evb[b][evn[b]].add=add; evn[b]++;}
static void push_phase(int phase, double eps, int K, int lat, double rho){
  int cap=8*V+1024; int32_t *q=malloc(sizeof(int32_t)*cap); int64_t qh=0,qt=0;
  int32_t *defer=malloc(sizeof(int32_t)*cap); int64_t nd=0;
  double sgn = phase? -1.0: 1.0; long inflight=0;
  double mx=0; for(int u=0;u<V;++u){double x=sgn*r[u]; if(x>mx) mx=x;}
  if(mx<=eps){free(q);free(defer);return;}
  double theta = rho>1 ? mx/rho : eps; if(theta<eps) theta=eps;
  for(int u=0;u<V;++u){double x=sgn*r[u]; if(x>theta) q[(qt++)%cap]=u; else if(x>eps) defer[nd++]=u;}
  int step=0;
  while(1){
    int b=step%RING;
    for(int i=0;i<evn[b];++i){ int v=evb[b][i].v; double add=evb[b][i].add; double old=r[v]; r[v]=old+add; T++; inflight--;
        double xo=sgn*old, xn=sgn*(old+add);
        if(!(xo>theta)&&xn>theta) q[(qt++)%cap]=v;
        else if(!(xo>eps)&&xn>eps) defer[nd++]=v; }
    evn[b]=0;
    int popped=0;
    while(popped<K && qh<qt){int u=q[(qh++)%cap]; double ru=r[u]; if(ru==0) continue; r[u]=0; p[u]+=A*ru; F++; popped++;
      for(int j=in_ptr[u];j<in_ptr[u+1];++j){int v=in_col[j]; ev_push(step+lat+(int)(rnd()%(unsigned)lat), v, (1-A)*ru/(deg[v]+1)); inflight++;} }
    step++;
    if(inflight==0 && qh==qt){
      if(theta<=eps) break;
      theta/=rho; if(theta<eps) theta=eps; LEVELS++;
      int64_t keep=0; for(int64_t i=0;i<nd;++i){int u=defer[i]; double x=sgn*r[u]; if(x>theta) q[(qt++)%cap]=u; else if(x>eps) defer[keep++]=u;} nd=keep;
    }
  }
  STEPS+=step;
  free(q);free(defer);
}
int main(int argc,char**argv){
  const char*fn=argv[1]; directed=atoi(argv[2]); int src=atoi(argv[3]); int K=atoi(argv[4]); int lat=atoi(argv[5]); double rho=atof(argv[6]); int nb=atoi(argv[7]);
  double eps=1e-9; FILE*f=fopen(fn,"rb"); fseek(f,0,SEEK_END); long sz=ftell(f); rewind(f); if(fread(&V,4,1,f)!=1) return 1; M=(sz-4)/8; pairs=malloc(8*M); if(fread(pairs,8,M,f)!=(size_t)M) return 1; fclose(f);
  W=(int64_t)(M*0.1); int64_t B=(int64_t)(0.01*W); int64_t Ew=directed?W:2*W;
  in_ptr=malloc(sizeof(int32_t)*(V+1)); in_col=malloc(sizeof(int32_t)*Ew); deg=malloc(sizeof(int32_t)*V); p=calloc(V,8); r=calloc(V,8);
  int32_t*deg0=malloc(sizeof(int32_t)*V);
  build(W); r[src]=1; push_phase(0,eps,K,lat,rho);
  printf("init: steps %lld levels %lld pops %lld traversed %lld\n",(long long)STEPS,(long long)LEVELS,(long long)F,(long long)T);
  int64_t pos=W;
  for(int k=0;k<nb;++k){ T=F=STEPS=LEVELS=0; memcpy(deg0,deg,sizeof(int32_t)*V);
    int64_t left=pos-W; pos+=B; int32_t*pd=deg0;
    int64_t nent=directed?2*B:4*B; int32_t*u1=malloc(4*nent),*v1=malloc(4*nent); char*ins=malloc(nent); int64_t c=0;
    for(int64_t i=0;i<B;++i){u1[c]=pairs[2*(left+i)];v1[c]=pairs[2*(left+i)+1];ins[c++]=0;}
    for(int64_t i=0;i<B;++i){u1[c]=pairs[2*(pos-B+i)];v1[c]=pairs[2*(pos-B+i)+1];ins[c++]=1;}
    if(!directed){int64_t len=c; for(int64_t i=0;i<len;++i){u1[c]=v1[i];v1[c]=u1[i];ins[c++]=ins[i];}}
    build(pos);
    for(int64_t i=0;i<c;++i){int u=u1[i],v=v1[i]; double add=(1-A)*p[v]-p[u]-A*r[u]+A*(u==src); if(ins[i]){pd[u]++; r[u]+=add/(pd[u]+1)/A;} else {pd[u]--; r[u]-=add/(pd[u]+1)/A;}}
    free(u1);free(v1);free(ins);
    push_phase(0,eps,K,lat,rho); push_phase(1,eps,K,lat,rho);
    double mx=0; for(int u=0;u<V;++u) if(fabs(r[u])>mx) mx=fabs(r[u]);
    printf("batch %d: steps %lld levels %lld pops %lld traversed %lld  max|r|/eps %.3f\n",k,(long long)STEPS,(long long)LEVELS,(long long)F,(long long)T,mx/eps);
  }
  return 0;
}
