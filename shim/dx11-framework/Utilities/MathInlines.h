// shim for stoyannk/dx11-framework Utilities/MathInlines.h (external, unpinned, not in the
// reference repo).  Semantics ASSUMED: plain min / max / clamp.
#pragma once
#include <type_traits>
namespace StMath
{
template <typename A, typename B>
inline typename std::common_type<A, B>::type max_value(A a, B b)
{
	typedef typename std::common_type<A, B>::type T;
	return T(a) > T(b) ? T(a) : T(b);
}
template <typename A, typename B>
inline typename std::common_type<A, B>::type min_value(A a, B b)
{
	typedef typename std::common_type<A, B>::type T;
	return T(a) < T(b) ? T(a) : T(b);
}
template <typename T>
inline T clamp_value(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }
}
