// biquad_state_merge.c - can a channel be split in TIME?  A second trajectory of the fixed-point biquad cascade is started from
// a zeroed (y1, y2, residual) state somewhere inside the stream (input history correct) and we look for the first sample at
// which its state equals the true one.  It never does: the 14-bit residual is an exact carry (a marginally stable mode), so
// a speculative warm-up cannot reproduce the reference bit for bit.  Build: gcc -O2 -o merge biquad_state_merge.c -lm
// Run:   ./merge 236552419 473104839 236552419 175469220 -47937074 1049016272 -1483533003 1049016272 1483533003 -1024290721 KIND AMP TRIALS
//        (the sketch's low-pass and notch, a1/a2 as stored i.e. negated; KIND 0 noise, 1 two tones, 2 tone+noise, 3 silence, 4 DC)
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
static inline int32_t smulw(int32_t a, int32_t b16){ return (int32_t)(((int64_t)a * (int16_t)b16) >> 16); }
static inline int32_t ssat16(int32_t v){ return v > 32767 ? 32767 : v < -32768 ? -32768 : v; }
typedef struct { int32_t x1,x2,y1,y2; uint32_t sum; } St;
static inline int32_t step(const int32_t *c, St *s, int32_t x0){
  s->sum += smulw(c[0],x0)+smulw(c[1],s->x1)+smulw(c[2],s->x2)+smulw(c[3],s->y1)+smulw(c[4],s->y2);
  int32_t y0 = ssat16((int32_t)s->sum >> 14); s->sum &= 0x3FFF; s->x2=s->x1; s->x1=x0; s->y2=s->y1; s->y1=y0; return y0; }
int main(int argc, char **argv){
  int32_t c1[5], c2[5]; for (int i=0;i<5;i++) c1[i]=atoi(argv[1+i]); for (int i=0;i<5;i++) c2[i]=atoi(argv[6+i]);
  int kind = atoi(argv[11]); double amp = atof(argv[12]); int trials = atoi(argv[13]);
  const int N = 1<<17; int16_t *x = malloc(N*2); St *t1 = malloc(N*sizeof(St)), *t2 = malloc(N*sizeof(St));
  long hist1[64]={0}, hist2[64]={0}; int never1=0, never2=0; long max1=0,max2=0; int lated=0; long latediff=0, latecnt=0;
  for (int tr=0; tr<trials; tr++){
    srand(1000+tr);
    double f1 = 200+ (rand()%6000), f2 = 300 + (rand()%3000), ph = rand()%1000;
    for (int i=0;i<N;i++){ double v;
      if (kind==0) v = amp*( (rand()/(double)RAND_MAX)*2-1 );
      else if (kind==1) v = amp*0.5*(sin(2*M_PI*f1*i/44117.0+ph)+sin(2*M_PI*f2*i/44117.0));
      else if (kind==2) v = amp*0.6*(sin(2*M_PI*f1*i/44117.0+ph)) + 30*((rand()/(double)RAND_MAX)*2-1);
      else if (kind==3) v = 0; else v = amp; // silence / DC
      x[i] = (int16_t)ssat16((int32_t)lrint(v)); }
    St a={0},b={0};
    for (int i=0;i<N;i++){ int32_t y = step(c1,&a,x[i]); t1[i]=a; step(c2,&b,y); t2[i]=b; }
    // speculative start at t0: zero y/sum, correct x history for stage 1; stage 2 x history = speculative stage-1 outputs
    for (int s=0;s<4;s++){ int t0 = 20000 + 20000*s + (rand()%128);
      St p=t1[t0-1]; p.y1=p.y2=0; p.sum=0; St q={0};
      long m1=-1, m2=-1;
      for (int i=t0;i<N;i++){ int32_t y=step(c1,&p,x[i]); step(c2,&q,y);
        if (i >= N-1000) { int dd = abs(q.y1 - t2[i].y1); if (dd > lated) lated = dd; latediff += (q.y1 != t2[i].y1); latecnt++; }
        if (m1<0 && p.y1==t1[i].y1 && p.y2==t1[i].y2 && p.sum==t1[i].sum) m1=i-t0;
        if (m1>=0 && m2<0 && q.x1==t2[i].x1 && q.x2==t2[i].x2 && q.y1==t2[i].y1 && q.y2==t2[i].y2 && q.sum==t2[i].sum){ m2=i-t0; break; } }
      if (m1<0) never1++; else { if (m1>max1) max1=m1; int bb=0; while ((1L<<bb) <= m1) bb++; hist1[bb]++; }
      if (m2<0) never2++; else { if (m2>max2) max2=m2; int bb=0; while ((1L<<bb) <= m2) bb++; hist2[bb]++; } } }
  printf("kind %d amp %.0f: stage1 never %d max %ld | cascade never %d max %ld | last 1000 samples of unmerged runs: max |dy| %d, %.1f%% of outputs differ\n  cascade log2 hist:", kind, amp, never1, max1, never2, max2, lated, latecnt ? 100.0*latediff/latecnt : 0.0);
  for (int b=0;b<20;b++) printf(" %ld", hist2[b]); printf("\n"); return 0; }
