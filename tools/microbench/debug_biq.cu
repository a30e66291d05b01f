// debug: print the criteria of biquadratic_film_coeffs for one hBN-like point
#include <cstdio>
#include "../../hyperbolic_optics_b200/csrc/ho_berreman.cuh"
using namespace ho;
__global__ void k() {
  Sym6<double> e;
  e.xx = mk<double>(7.38135757, 0.0063377); e.yy = e.xx; e.zz = mk<double>(4.95457993, 0.04740561);
  e.xy = mk<double>(0, 0); e.xz = e.xy; e.yz = e.xy;
  ZRow<double> z;
  Delta<double, true> D = berreman_nm<double>(1.7, e, mk<double>(1, 0), z);
  cx<double> c3, c2, c1, c0;
  charpoly(D, c3, c2, c1, c0);
  printf("c3 %g %g c2 %g %g c1 %g %g c0 %g %g\n", c3.re, c3.im, c2.re, c2.im, c1.re, c1.im, c0.re, c0.im);
  cx<double> y0, y1;
  quad_roots(c2, c0, y0, y1);
  printf("y0 %g %g y1 %g %g\n", y0.re, y0.im, y1.re, y1.im);
  cx<double> c = mk<double>(0, -0.0628);
  cx<double> z0 = c * csqrt(y0), z1 = c * csqrt(y1);
  printf("z0 %g %g z1 %g %g thin %d\n", z0.re, z0.im, z1.re, z1.im, (int)thin_film(ho_abs(z0.re), ho_abs(z1.re)));
  cx<double> a0, a1, a2, a3;
  bool ok = biquadratic_film_coeffs(c3, c2, c1, c0, c, a0, a1, a2, a3);
  printf("ok %d a0 %g %g a1 %g %g a2 %g %g a3 %g %g\n", (int)ok, a0.re, a0.im, a1.re, a1.im, a2.re, a2.im, a3.re, a3.im);
}
int main() { k<<<1, 1>>>(); cudaDeviceSynchronize(); return 0; }
