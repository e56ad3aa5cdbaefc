// mathd.cuh -- sin / cos evaluated to ~2^-68 and rounded once, for the panorama cameras.
//
// Camera::GenerateEnvRay / GenerateStereoEnvRay (camera.cc:242-329) call the host's libm.  CUDA's double sin / cos are
// allowed 1-2 ulp, glibc's are correctly rounded in ~99.9 % of their results, so 44 % of the device's panorama rays
// differed from the reference's in the last bit of a component.  sincos_rn() computes both functions in double-double
// arithmetic (error < 2^-66 relative, i.e. the correctly rounded double except when the exact value lies within
// ~2^-13 ulp of a rounding boundary), which is what a good host libm returns: rays now agree bit for bit wherever
// glibc itself rounds correctly.  Not a libm: arguments beyond 2^19 * pi/2 (the panorama angles are in [0, 2 pi]) and
// non-finite ones go to the CUDA functions.
//
// Method (the classic one, constants are Taylor coefficients and the three 33-bit pieces of pi/2 of Cody-Waite
// reduction): k = rint(x * 2/pi); r = x - k*pi/2 as a double-double (k * piece is exact for |k| < 2^20);
// sin(r) = r + r z (s1 + z (s2 + Q(z))), cos(r) = 1 - z/2 + z^2 (c2 + z (c3 + Q'(z))), z = r^2, with the two leading
// coefficients and every product / sum around them in double-double and the tails Q, Q' (|Q| < 1.3e-4) in double;
// the quadrant k mod 4 selects and signs the results.  Compiles as host code too (tests/test_mathd.py holds it
// against mpmath and against the host's libm on CPU).
#ifndef MB200_MATHD_CUH
#define MB200_MATHD_CUH

#include <cmath>

#ifdef __CUDACC__
#define MB200_HD __host__ __device__ __forceinline__
#else
#define MB200_HD inline
#endif

namespace mb200 {
namespace mathd {

struct dd {
  double hi, lo;
};

MB200_HD dd two_sum(double a, double b) {
  const double s = a + b, bb = s - a;
  return dd{s, (a - (s - bb)) + (b - bb)};
}
MB200_HD dd quick_two_sum(double a, double b) { // |a| >= |b|
  const double s = a + b;
  return dd{s, b - (s - a)};
}
MB200_HD dd two_prod(double a, double b) {
  const double p = a * b;
  return dd{p, fma(a, b, -p)};
}
MB200_HD dd add(dd a, dd b) {
  dd s = two_sum(a.hi, b.hi);
  const dd t = two_sum(a.lo, b.lo);
  s.lo += t.hi;
  s = quick_two_sum(s.hi, s.lo);
  s.lo += t.lo;
  return quick_two_sum(s.hi, s.lo);
}
MB200_HD dd add_d(dd a, double b) {
  dd s = two_sum(a.hi, b);
  s.lo += a.lo;
  return quick_two_sum(s.hi, s.lo);
}
MB200_HD dd mul(dd a, dd b) {
  dd p = two_prod(a.hi, b.hi);
  p.lo += a.hi * b.lo + a.lo * b.hi;
  return quick_two_sum(p.hi, p.lo);
}
MB200_HD dd mul_d(dd a, double b) {
  dd p = two_prod(a.hi, b);
  p.lo += a.lo * b;
  return quick_two_sum(p.hi, p.lo);
}

// sin and cos of a reduced argument |r| <= pi/4 (+ a few ulp), as double-doubles
MB200_HD void sincos_reduced(dd r, dd &s, dd &c) {
  dd z = two_prod(r.hi, r.hi);
  z.lo += 2.0 * r.hi * r.lo;
  z = quick_two_sum(z.hi, z.lo);
  const double x = z.hi;
  // tails in double: sin  z (s3 + z (s4 + ... s10)),  cos  z (c4 + z (c5 + ... c10))
  const double qs = x * (-0.0001984126984126984 + x * (2.7557319223985893e-06 + x * (-2.505210838544172e-08 +
                    x * (1.6059043836821613e-10 + x * (-7.647163731819816e-13 + x * (2.8114572543455206e-15 +
                    x * (-8.22063524662433e-18 + x * 1.9572941063391263e-20)))))));
  const double qc = x * (2.48015873015873e-05 + x * (-2.755731922398589e-07 + x * (2.08767569878681e-09 +
                    x * (-1.1470745597729725e-11 + x * (4.779477332387385e-14 + x * (-1.5619206968586225e-16 +
                    x * 4.110317623312165e-19))))));
  // sin: r + r z (s1 + z (s2 + qs))
  dd w = add_d(dd{0.008333333333333333, 1.1564823173178714e-19}, qs);
  w = add(dd{-0.16666666666666666, -9.25185853854297e-18}, mul(z, w));
  s = add(r, mul(mul(r, z), w));
  // cos: 1 - z/2 + z^2 (c2 + z (c3 + qc))
  dd v = add_d(dd{-0.001388888888888889, 5.300543954373577e-20}, qc);
  v = add(dd{0.041666666666666664, 2.3129646346357427e-18}, mul(z, v));
  const dd half_z = dd{-0.5 * z.hi, -0.5 * z.lo};
  c = add(add_d(half_z, 1.0), mul(mul(z, z), v));
}

// sin(x) and cos(x), each rounded once from a ~2^-68-accurate value.
MB200_HD void sincos_rn(double x, double &sn, double &cs) {
  const double ax = fabs(x);
  if (!(ax < 823549.6)) { // 2^19 * pi/2; also NaN / inf: the library functions
    sn = sin(x), cs = cos(x);
    return;
  }
  if (x == 0.0) { // keeps the sign of zero
    sn = x, cs = 1.0;
    return;
  }
  const double k = rint(x * 0.6366197723675814);
  // pi/2 = p1 + p2 + p3 + p3t, p1..p3 33 bits each (k * p exact)
  const double p1 = 1.57079632673412561417e+00, p2 = 6.07710050630396597660e-11, p3 = 2.02226624871116645580e-21,
               p3t = 8.47842766036889956997e-32;
  dd r = two_sum(x, -k * p1);
  r = add(r, two_sum(-k * p2, -k * p3));
  r = add_d(r, -k * p3t);
  dd s, c;
  sincos_reduced(r, s, c);
  const int q = (int)((long long)k & 3LL);
  const double sv = s.hi, cv = c.hi; // normalised double-doubles: hi is the sum rounded to nearest
  switch (q) {
    case 0: sn = sv, cs = cv; break;
    case 1: sn = cv, cs = -sv; break;
    case 2: sn = -sv, cs = -cv; break;
    default: sn = -cv, cs = sv; break;
  }
}

} // namespace mathd
} // namespace mb200

#endif
