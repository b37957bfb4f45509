// Stand-in for glare-core utils/Vector.h (glare-core is NOT vendored in the reference tree:
// /root/reference/CMakeLists.txt:107 locates it through $GLARE_CORE_TRUNK_DIR).  Only what
// audio/AudioResampler.{h,cpp} uses: js::Vector<float,16> with resizeNoCopy/resize/data/size/[],
// the int64 typedef and myClamp.  TEST INFRASTRUCTURE: used to compile the reference file unmodified.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstdlib>
#include <cassert>
#include <cmath>
#include <vector>
#include <algorithm>

typedef long long int64;            // same underlying types as Resonance's base/integral_types.h:44,59
typedef unsigned long long uint64;
typedef uint32_t uint32;
typedef uint8_t uint8;

template <class T> inline T myClamp(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }
template <class T> inline T myMin(T a, T b) { return a < b ? a : b; }
template <class T> inline T myMax(T a, T b) { return a > b ? a : b; }

namespace js {
template <class T, size_t alignment>
class Vector {
public:
	Vector() {}
	explicit Vector(size_t n) : v(n) {}
	void resize(size_t n) { v.resize(n); }
	void resizeNoCopy(size_t n) { v.resize(n); }
	size_t size() const { return v.size(); }
	T* data() { return v.data(); }
	const T* data() const { return v.data(); }
	T& operator[](size_t i) { return v[i]; }
	const T& operator[](size_t i) const { return v[i]; }
	T* begin() { return v.data(); }
	T* end() { return v.data() + v.size(); }
private:
	std::vector<T> v;
};
}
