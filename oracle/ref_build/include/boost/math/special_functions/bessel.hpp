// TEST INFRASTRUCTURE ONLY.  BH (Boost.Math) is not in the image and is unpinned by the reference
// (DESCRIPTION:26).  boost::math::cyl_bessel_k -> std::cyl_bessel_k, the substitution the reference
// itself shipped in v0.1.5 (NEWS.md:11-15).
#ifndef GPV_REF_STUB_BOOST_BESSEL_HPP
#define GPV_REF_STUB_BOOST_BESSEL_HPP
#include <cmath>
namespace boost { namespace math {
inline double cyl_bessel_k(double v, double x) { return std::cyl_bessel_k(v, x); }
}}
#endif
