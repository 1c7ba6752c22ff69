// Stand-in for boost::math::legendre_p (three-term recurrence).  TEST INFRASTRUCTURE ONLY.
#ifndef SRI_BOOST_SHIM_LEGENDRE
#define SRI_BOOST_SHIM_LEGENDRE
namespace boost { namespace math {
inline double legendre_p(int l, double x) {
    if (l < 0) l = -l - 1;
    double p0 = 1.0, p1 = x;
    if (l == 0) return p0;
    for (int k = 1; k < l; ++k) { const double p2 = ((2 * k + 1) * x * p1 - k * p0) / (k + 1); p0 = p1; p1 = p2; }
    return p1;
}
}}  // namespace boost::math
#endif
