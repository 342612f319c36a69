/* oracle/check_div_const.c — TEST INFRASTRUCTURE (never linked into the product).
 * Justifies the 3-instruction constant division in csrc/ofsv_common.cuh (norm_flow): run once with
 *   gcc -O2 -fopenmp -ffp-contract=off -mfma -o /tmp/chk oracle/check_div_const.c -lm && /tmp/chk     (about 2 minutes on 8 cores)
 * Result recorded in DESIGN.md: 0 mismatches for every float 2^-60 <= |x| < 2^20 and every size below. */
// exhaustive check: for constants c = (S-1)/2, is  q0=RN(x*rc); r=fma(-q0,c,x); q1=fma(r,rc,q0)  ==  x/c  for all floats |x| <= 2^20 ?
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
int main(int argc, char** argv) {
  int sizes[] = {2,3,4,5,6,7,8,9,10,12,13,16,17,20,24,26,28,31,32,33,35,36,40,48,52,56,64,96,104,112,128,160,208,224,256,320,384,448,480,512,640,832,1024};
  int ns = sizeof(sizes)/sizeof(int);
  long long bad_total = 0;
  #pragma omp parallel for schedule(dynamic) reduction(+:bad_total)
  for (int si = 0; si < ns; ++si) {
    const float c = (float)((sizes[si] - 1.0) / 2.0);
    const float rc = (float)(1.0 / ((sizes[si] - 1.0) / 2.0));
    long long bad = 0;
    for (uint32_t b = 0x21800000u; b < 0x49800000u; ++b) {       // positive floats up to 2^20 (sign symmetric)
      float x; memcpy(&x, &b, 4);
      const float q0 = x * rc;
      const float r = fmaf(-q0, c, x);
      const float q1 = fmaf(r, rc, q0);
      const float t = x / c;
      if (q1 != t) { if (bad < 3) printf("S=%d x=%a got %a want %a\n", sizes[si], x, q1, t); ++bad; }
    }
    printf("S=%d c=%g: %lld mismatches\n", sizes[si], c, bad);
    bad_total += bad;
  }
  printf("TOTAL %lld\n", bad_total);
  return 0;
}
