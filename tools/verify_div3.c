#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
int main(){
    const float y = 1.0f/3.0f;   /* == 0.3333333432674407958984375f */
    uint64_t bad=0, first=0; int have=0;
    uint32_t lo_bad=0xffffffff, hi_bad=0;
    #pragma omp parallel for reduction(+:bad)
    for (uint64_t i=0;i<0x7f800000ull;i++){
        uint32_t u=(uint32_t)i; float x; memcpy(&x,&u,4);
        float q=x*y; float r=fmaf(-3.0f,q,x); float q2=fmaf(r,y,q);
        float want=x/3.0f;
        if (memcmp(&q2,&want,4)) { bad++;
            #pragma omp critical
            { if(u<lo_bad)lo_bad=u; if(u>hi_bad)hi_bad=u; } }
    }
    printf("bad=%llu lo=%08x hi=%08x\n",(unsigned long long)bad,lo_bad,hi_bad);
    return 0;
}
