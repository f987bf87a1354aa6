// Deterministic double-precision sincos for the IK kernels: the same sequence of IEEE multiplies, adds and explicit FMAs on the
// GPU and in the CPU build of the same sources (tests/emu), so that - compiled without implicit FMA contraction - the solver
// gives bit-identical results on both (CUDA's and glibc's sincos are each within 1 ulp of the truth but not of each other,
// and the reference's rank-deficient trust-region steps amplify one ulp to decimetres: SURVEY.md 8c').
// Algorithm: Cody-Waite argument reduction by pi/2 in two pieces with a compensated subtraction, then the
// fdlibm kernel polynomials on [-pi/4, pi/4] (Sun Microsystems' freely distributable libm: k_sin.c / k_cos.c coefficients).
// Error < 1 ulp for |x| < 1e5 (checked against glibc in tests/test_oracle_golden.py::test_det_sincos).
#pragma once

namespace mvmc {

__host__ __device__ __forceinline__ void det_sincos(double x, double* sn, double* cs) {
    const double invpio2 = 6.36619772367581382433e-01;
    const double pio2_1 = 1.57079632673412561417e+00, pio2_1t = 6.07710050650619224932e-11;
    const double big = 6755399441055744.0;   // 1.5 * 2^52: (t + big) - big rounds t to the nearest integer
    // n = nearest integer to x * 2/pi; y0 + y1 = x - n pi/2: the leading 33 bits of pi/2 come off exactly (|n| < 2^20), the
    // next piece with a compensated subtraction (about 85 bits of pi/2 in all)
    const double fn = (x * invpio2 + big) - big;
    const double r0 = x - fn * pio2_1;
    const double w1 = fn * pio2_1t;
    const double ra = r0 - w1;
    const double wa = w1 - (r0 - ra);
    const double y0 = ra - wa;
    const double y1 = (ra - y0) - wa;
    const int n = (int)fn;
    // kernels on |y| <= pi/4
    const double z = y0 * y0;
    // sin
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double v = z * y0;
    const double rs = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    const double ks = y0 - ((z * (0.5 * y1 - v * rs) - y1) - v * S1);
    // cos
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double rc = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    const double hz = 0.5 * z;
    const double wc = 1.0 - hz;
    const double kc = wc + (((1.0 - wc) - hz) + (z * rc - y0 * y1));
    switch (n & 3) {
        case 0: *sn = ks; *cs = kc; break;
        case 1: *sn = kc; *cs = -ks; break;
        case 2: *sn = -ks; *cs = -kc; break;
        default: *sn = -kc; *cs = ks; break;
    }
}

}  // namespace mvmc
