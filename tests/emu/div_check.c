// TEST INFRASTRUCTURE - the reciprocal + two-correction division of csrc/epilogue_staged.cuh (struct DivBy) against the
// hardware IEEE division, bit for bit: every sqrt(abar_t) of the scaled-linear DDIM schedule and 2000 random divisors
// (including all-ones significands) x N random dividends each (argv[1], default 400000 -> 1.2e9 quotients).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline float mk(float a, float b, float nb, float y){ float q=a*y; float r=fmaf(nb,q,a); q=fmaf(r,y,q); r=fmaf(nb,q,a); q=fmaf(r,y,q); return q; }
static uint64_t s=88172645463325252ull; static inline uint64_t rnd(){ s^=s<<13; s^=s>>7; s^=s<<17; return s; }
int main(int argc, char** argv){
  const int N = argc > 1 ? atoi(argv[1]) : 400000;
  // divisors: sqrt(alpha_bar_t) of the scaled-linear schedule (fp32 math like torch) + random divisors
  static float bs[3000]; int nb_=0;
  { double b0=sqrt(0.00085), b1=sqrt(0.012); float ac=1.f; for(int t=0;t<1000;t++){ float beta=(float)(b0+(b1-b0)*t/999.0); beta=beta*beta; ac*= (1.f-beta); bs[nb_++]=sqrtf(ac);} }
  for(int i=0;i<2000;i++){ uint32_t u=(uint32_t)rnd(); u=(u&0x007fffff)|((uint32_t)(100+rnd()%56)<<23); float f; memcpy(&f,&u,4); bs[nb_++]=f; }
  bs[nb_-1]=1.0f; { uint32_t u=0x3f7fffff; memcpy(&bs[nb_-2],&u,4);} { uint32_t u=0x3effffff; memcpy(&bs[nb_-3],&u,4);}
  long long bad=0, tot=0;
  for(int i=0;i<nb_;i++){ float b=bs[i], y=1.0f/b, nb=-b;
    for(int j=0;j<N;j++){ uint32_t u=(uint32_t)rnd(); uint32_t e=(j&7)==0? (uint32_t)(rnd()%200+27) : (uint32_t)(rnd()%16+120); u=(u&0x807fffff)|(e<<23); float a; memcpy(&a,&u,4);
      float q=mk(a,b,nb,y), ref=a/b; tot++; if(memcmp(&q,&ref,4)){ if(bad<5) printf("MISMATCH a=%a b=%a got %a want %a\n",a,b,q,ref); bad++; } } }
  printf("checked %lld divisions, %lld mismatches\n",tot,bad); return bad!=0; }
