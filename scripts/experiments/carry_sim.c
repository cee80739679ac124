// experiment (CPU): level-synchronous (Jacobi) push where iteration k only pops frontier items with
// |r| > theta_k = max(eps, theta0 * gamma^k); the others are carried to the next frontier untouched.
// gcc -O2 -o build/carry_sim scripts/experiments/carry_sim.c -lm
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
static int64_t T,F,C,IT;
static void push_phase(int phase, double eps, double gamma, double theta_scale){
  int32_t *ft=malloc(sizeof(int32_t)*V), *ft2=malloc(sizeof(int32_t)*V); double *fr=malloc(sizeof(double)*V);
  double sgn=phase?-1:1; int n=0; double mx=0;
  for(int u=0;u<V;++u){double x=sgn*r[u]; if(x>eps){ft[n++]=u; if(x>mx) mx=x;}}
  double theta = gamma<1 ? mx*theta_scale : eps; if(theta<eps) theta=eps;
  while(n){
    IT++;
    int n2=0, npop=0;
    for(int i=0;i<n;++i){int u=ft[i]; double x=sgn*r[u]; if(x>theta){fr[i]=r[u]; p[u]+=A*r[u]; r[u]=0; npop++;} else {fr[i]=0; ft2[n2++]=u; C++;}}
    F+=npop;
    for(int i=0;i<n;++i){ double ru=fr[i]; if(ru==0) continue; int u=ft[i];
      for(int j=in_ptr[u];j<in_ptr[u+1];++j){int v=in_col[j]; double add=(1-A)*ru/(deg[v]+1); double old=r[v]; r[v]=old+add; T++;
        double xo=sgn*old, xn=sgn*(old+add); if(!(xo>eps)&&xn>eps) ft2[n2++]=v; }
    }
    int32_t*t=ft;ft=ft2;ft2=t;n=n2;
    theta*=gamma; if(theta<eps) theta=eps;
  }
  free(ft);free(ft2);free(fr);
}
int main(int argc,char**argv){
  const char*fn=argv[1]; directed=atoi(argv[2]); int src=atoi(argv[3]); double gamma=atof(argv[4]); double ts=atof(argv[5]); int nb=atoi(argv[6]);
  double eps=1e-9; FILE*f=fopen(fn,"rb"); fseek(f,0,SEEK_END); long sz=ftell(f); rewind(f); if(fread(&V,4,1,f)!=1) return 1; M=(sz-4)/8; pairs=malloc(8*M); if(fread(pairs,8,M,f)!=(size_t)M) return 1; fclose(f);
  W=(int64_t)(M*0.1); int64_t B=(int64_t)(0.01*W); int64_t Ew=directed?W:2*W;
  in_ptr=malloc(sizeof(int32_t)*(V+1)); in_col=malloc(sizeof(int32_t)*Ew); deg=malloc(sizeof(int32_t)*V); p=calloc(V,8); r=calloc(V,8);
  int32_t*deg0=malloc(sizeof(int32_t)*V);
  build(W); r[src]=1; push_phase(0,eps,gamma,ts);
  printf("init: iters %lld pops %lld carried %lld traversed %lld\n",(long long)IT,(long long)F,(long long)C,(long long)T);
  int64_t pos=W;
  for(int k=0;k<nb;++k){ T=F=C=IT=0; memcpy(deg0,deg,sizeof(int32_t)*V);
    int64_t left=pos-W; pos+=B; int32_t*pd=deg0;
    int64_t nent=directed?2*B:4*B; int32_t*u1=malloc(4*nent),*v1=malloc(4*nent); char*ins=malloc(nent); int64_t c=0;
    for(int64_t i=0;i<B;++i){u1[c]=pairs[2*(left+i)];v1[c]=pairs[2*(left+i)+1];ins[c++]=0;}
    for(int64_t i=0;i<B;++i){u1[c]=pairs[2*(pos-B+i)];v1[c]=pairs[2*(pos-B+i)+1];ins[c++]=1;}
    if(!directed){int64_t len=c; for(int64_t i=0;i<len;++i){u1[c]=v1[i];v1[c]=u1[i];ins[c++]=ins[i];}}
    build(pos);
    for(int64_t i=0;i<c;++i){int u=u1[i],v=v1[i]; double add=(1-A)*p[v]-p[u]-A*r[u]+A*(u==src); if(ins[i]){pd[u]++; r[u]+=add/(pd[u]+1)/A;} else {pd[u]--; r[u]-=add/(pd[u]+1)/A;}}
    free(u1);free(v1);free(ins);
    push_phase(0,eps,gamma,ts); push_phase(1,eps,gamma,ts);
    double mx=0; for(int u=0;u<V;++u) if(fabs(r[u])>mx) mx=fabs(r[u]);
    printf("batch %d: iters %lld pops %lld carried %lld traversed %lld  max|r|/eps %.3f\n",k,(long long)IT,(long long)F,(long long)C,(long long)T,mx/eps);
  }
  return 0;
}
