// TEST INFRASTRUCTURE ONLY.  boost::math::tgamma -> std::tgamma (see bessel.hpp).
#ifndef GPV_REF_STUB_BOOST_GAMMA_HPP
#define GPV_REF_STUB_BOOST_GAMMA_HPP
#include <cmath>
namespace boost { namespace math {
inline double tgamma(double x) { return std::tgamma(x); }
}}
#endif
