// Host check behind logf_ref (csrc/ulcb_encode.cu): the restated glibc logf against the host libm over every positive finite float.
// gcc -O2 -fopenmp -ffp-contract=off -mfma scripts/check_logf.c -o /tmp/check_logf -lm && /tmp/check_logf   (5 s on 8 cores)
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static const double T[16][2] = {
 {0x1.661ec79f8f3bep+0,-0x1.57bf7808caadep-2},{0x1.571ed4aaf883dp+0,-0x1.2bef0a7c06ddbp-2},{0x1.49539f0f010bp+0,-0x1.01eae7f513a67p-2},
 {0x1.3c995b0b80385p+0,-0x1.b31d8a68224e9p-3},{0x1.30d190c8864a5p+0,-0x1.6574f0ac07758p-3},{0x1.25e227b0b8eap+0,-0x1.1aa2bc79c81p-3},
 {0x1.1bb4a4a1a343fp+0,-0x1.a4e76ce8c0e5ep-4},{0x1.12358f08ae5bap+0,-0x1.1973c5a611cccp-4},{0x1.0953f419900a7p+0,-0x1.252f438e10c1ep-5},
 {0x1p+0,0x0p+0},{0x1.e608cfd9a47acp-1,0x1.aa5aa5df25984p-5},{0x1.ca4b31f026aap-1,0x1.c5e53aa362eb4p-4},
 {0x1.b2036576afce6p-1,0x1.526e57720db08p-3},{0x1.9c2d163a1aa2dp-1,0x1.bc2860d22477p-3},{0x1.886e6037841edp-1,0x1.1058bc8a07ee1p-2},
 {0x1.767dcf5534862p-1,0x1.4043057b6ee09p-2}};
static const double Ln2 = 0x1.62e42fefa39efp-1, A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
static inline float my_logf(float x, int usefma) {
  uint32_t ix; memcpy(&ix,&x,4);
  if (ix == 0x3f800000) return 0;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix*2 == 0) return -INFINITY; if (ix == 0x7f800000) return x; if ((ix & 0x80000000u) || ix*2 >= 0xff000000u) return NAN;
    float xs = x * 0x1p23f; memcpy(&ix,&xs,4); ix -= 23u << 23;
  }
  uint32_t tmp = ix - 0x3f330000u;
  int i = (tmp >> 19) % 16;
  int k = (int32_t)tmp >> 23;
  uint32_t iz = ix - (tmp & 0xff800000u);
  float zf; memcpy(&zf,&iz,4);
  double z = zf, invc = T[i][0], logc = T[i][1], r, y0, r2, y;
  if (usefma) { r = fma(z, invc, -1.0); y0 = fma((double)k, Ln2, logc); r2 = r*r; y = fma(A1, r, A2); y = fma(A0, r2, y); y = fma(y, r2, y0 + r); }
  else { r = z*invc - 1; y0 = logc + (double)k*Ln2; r2 = r*r; y = A1*r + A2; y = A0*r2 + y; y = y*r2 + (y0 + r); }
  return (float)y;
}
int main() {
  long bad_f = 0, bad_n = 0, bad_dl = 0; 
  #pragma omp parallel for reduction(+:bad_f,bad_n,bad_dl) schedule(static)
  for (long u = 1; u < 0x7f800000L; u++) {
    uint32_t ix = (uint32_t)u; float x; memcpy(&x,&ix,4);
    float ref = logf(x);
    float a = my_logf(x,1), b = my_logf(x,0), c = (float)log((double)x);
    uint32_t ur, ua, ub, uc; memcpy(&ur,&ref,4); memcpy(&ua,&a,4); memcpy(&ub,&b,4); memcpy(&uc,&c,4);
    bad_f += ur != ua; bad_n += ur != ub; bad_dl += ur != uc;
  }
  printf("mismatches vs host logf over all positive finite floats: fma-variant %ld, no-fma %ld, (float)log(double) %ld\n", bad_f, bad_n, bad_dl);
}
