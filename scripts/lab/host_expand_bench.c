/* lab: cost of one mnvh_expand (65 536 envs, ~26 000 sonar returns) per thread count:
   gcc -O3 -pthread host_expand_bench.c ../../distributional_rl_navigation_b200/csrc_host/mnv_host.c -o /tmp/hb && /tmp/hb THREADS [FIRST_CPU] */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>
#include "../../include/mnv_host.h"
static double now(){struct timespec t; clock_gettime(CLOCK_MONOTONIC,&t); return t.tv_sec+1e-9*t.tv_nsec;}
int main(int argc,char**argv){
  int nt=atoi(argv[1]); int first=argc>2?atoi(argv[2]):-1; int64_t E=65536; int D=26; int64_t G=E/32;
  float* obs=calloc(E*D,4); float* head=malloc(E*16); uint8_t* skip=calloc(E,1);
  uint32_t* mask[4]; uint32_t* dir[4]; float* vals[4];
  for(int k=0;k<4;k++){ mask[k]=calloc(E,4); dir[k]=calloc(G,4); vals[k]=malloc(E*11*8); uint32_t pos=0;
    for(int64_t g=0;g<G;g++){ dir[k][g]=pos; for(int64_t e=32*g;e<32*g+32;e++){ uint32_t m=0; for(int b=0;b<11;b++) if(rand()%1000<36) m|=1u<<b; mask[k][e]=m; pos+=__builtin_popcount(m);} }
    for(uint32_t i=0;i<2*pos;i++) vals[k][i]=1.0f+i; }
  for(int64_t i=0;i<E*4;i++) head[i]=i;
  for(int64_t e=0;e<E;e++) skip[e]=(rand()%1000<2);
  mnvh_pool* p=mnvh_create(nt,E,D,first);
  for(int it=0;it<20;it++) mnvh_expand(p,obs,head,skip,mask[it%4],dir[it%4],vals[it%4]);
  double t0=now(); int n=300;
  for(int it=0;it<n;it++) mnvh_expand(p,obs,head,skip,mask[it%4],dir[it%4],vals[it%4]);
  printf("%d threads (first cpu %d): %.1f us\n",nt,first,(now()-t0)/n*1e6);
  mnvh_destroy(p); return 0; }
