// Exhaustive check that wb_div1000() (python-world_b200/csrc/wb_platform.h: reciprocal multiply + two FMAs) equals the
// IEEE division a / 1000.0 for every frame time a = j * period the kernels form (j < 5e7, periods 0.5 .. 10 ms):
//   gcc -O2 -o /tmp/check_div1000 tools/check_div1000.c -lm && /tmp/check_div1000    ->  "mismatches 0"
#include <math.h>
#include <stdio.h>
int main() {
  const double r = 1.0 / 1000.0;
  long bad = 0, n = 0;
  double periods[] = {1.0, 5.0, 2.5, 10.0, 2.0, 7.0, 0.5, 1.25, 3.3};
  for (int p = 0; p < 9; ++p)
    for (long j = 0; j < 50000000; ++j) {
      const double a = (double)j * periods[p];
      const double q0 = a * r;
      const double e = fma(-q0, 1000.0, a);
      const double q1 = fma(e, r, q0);
      if (q1 != a / 1000.0) { if (bad < 5) printf("bad j=%ld p=%g q1=%.17g true=%.17g\n", j, periods[p], q1, a / 1000.0); ++bad; }
      ++n;
    }
  printf("checked %ld, mismatches %ld\n", n, bad);
  return 0;
}
