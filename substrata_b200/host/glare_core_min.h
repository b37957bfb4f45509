// glare_core_min.h -- minimal stand-ins for the glare-core types that appear on the surface of audio/AudioEngine.h.
//
// glare-core (utils/, maths/, physics/) is a separate repository that the reference tree does not vendor
// (CMakeLists.txt:107, env GLARE_CORE_TRUNK_DIR), so the semantics below are the ones the audio code relies on
// (SURVEY.md 8b lists the call sites).  Inside Substrata these come from glare-core itself and this header is not
// used: define SBM_USE_GLARE_CORE before including AudioEngine.h and include the real headers instead.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

typedef uint32_t uint32;
typedef int64_t int64;
typedef uint8_t uint8;

namespace glare {
class Exception : public std::runtime_error {
public:
	explicit Exception(const std::string& s) : std::runtime_error(s) {}
	std::string what_str() const { return what(); }
};
}  // namespace glare

// maths/Vec4f.h: 4 x f32, 16-byte aligned; points carry w = 1, vectors w = 0 (AudioEngine.cpp:71,131-134).
struct alignas(16) Vec4f {
	float x[4];
	Vec4f() : x{0, 0, 0, 0} {}
	Vec4f(float a, float b, float c, float d) : x{a, b, c, d} {}
	float& operator[](size_t i) { return x[i]; }
	float operator[](size_t i) const { return x[i]; }
	bool isFinite() const { return std::isfinite(x[0]) && std::isfinite(x[1]) && std::isfinite(x[2]) && std::isfinite(x[3]); }
	float length() const { return sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]); }
};
inline Vec4f operator-(const Vec4f& a, const Vec4f& b) { return Vec4f(a[0] - b[0], a[1] - b[1], a[2] - b[2], a[3] - b[3]); }
inline Vec4f operator*(const Vec4f& a, float f) { return Vec4f(a[0] * f, a[1] * f, a[2] * f, a[3] * f); }
inline float dot(const Vec4f& a, const Vec4f& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]; }
// maths/Vec4f.h projectOntoDir: projection of `a` onto the (not necessarily unit) direction `dir`.
inline Vec4f projectOntoDir(const Vec4f& a, const Vec4f& dir) {
	const float d2 = dot(dir, dir);
	return d2 > 0 ? dir * (dot(a, dir) / d2) : Vec4f();
}

// maths/Quat.h: v = (x, y, z, w), passed as SetHeadRotation(v[0], v[1], v[2], v[3]) (AudioEngine.cpp:988).
struct Quatf {
	Vec4f v;
	Quatf() : v(0, 0, 0, 1) {}
	Quatf(float x, float y, float z, float w) : v(x, y, z, w) {}
};

namespace Maths {
inline float sign(float x) { return x < 0 ? -1.f : (x > 0 ? 1.f : 0.f); }
}
template <class T> inline T myMin(T a, T b) { return a < b ? a : b; }
template <class T> inline T myMax(T a, T b) { return a > b ? a : b; }
template <class T> inline T myClamp(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }

// utils/ThreadSafeRefCounted.h + utils/Reference.h: intrusive reference counting.
class ThreadSafeRefCounted {
public:
	ThreadSafeRefCounted() : refcount(0) {}
	virtual ~ThreadSafeRefCounted() {}
	void incRefCount() const { refcount.fetch_add(1, std::memory_order_relaxed); }
	int64_t decRefCount() const { return refcount.fetch_sub(1, std::memory_order_acq_rel) - 1; }
	int64_t getRefCount() const { return refcount.load(); }
private:
	mutable std::atomic<int64_t> refcount;
};

template <class T> class Reference {
public:
	Reference() : ob(nullptr) {}
	Reference(T* p) : ob(p) { if (ob) ob->incRefCount(); }
	Reference(const Reference& o) : ob(o.ob) { if (ob) ob->incRefCount(); }
	template <class U> Reference(const Reference<U>& o) : ob(o.ptr()) { if (ob) ob->incRefCount(); }
	~Reference() { release(); }
	Reference& operator=(const Reference& o) { if (o.ob) o.ob->incRefCount(); release(); ob = o.ob; return *this; }
	T* operator->() const { return ob; }
	T& operator*() const { return *ob; }
	T* ptr() const { return ob; }
	T* getPointer() const { return ob; }
	bool isNull() const { return ob == nullptr; }
	bool nonNull() const { return ob != nullptr; }
	explicit operator bool() const { return ob != nullptr; }
	bool operator<(const Reference& o) const { return ob < o.ob; }
	bool operator==(const Reference& o) const { return ob == o.ob; }
	bool operator!=(const Reference& o) const { return ob != o.ob; }
private:
	void release() { if (ob && ob->decRefCount() == 0) delete ob; ob = nullptr; }
	T* ob;
};
template <class T> using VRef = Reference<T>;   // utils/VRef.h: a Reference that is never null

namespace js {
// utils/Vector.h: contiguous vector with the given alignment; the audio code uses resize / resizeNoCopy / data / size / [] / iteration.
template <class T, size_t alignment> class Vector {
public:
	Vector() {}
	explicit Vector(size_t n) : v(n) {}
	void resize(size_t n) { v.resize(n); }
	void resizeNoCopy(size_t n) { v.resize(n); }
	size_t size() const { return v.size(); }
	bool empty() const { return v.empty(); }
	T* data() { return v.data(); }
	const T* data() const { return v.data(); }
	T& operator[](size_t i) { return v[i]; }
	const T& operator[](size_t i) const { return v[i]; }
	void push_back(const T& t) { v.push_back(t); }
	typename std::vector<T>::iterator begin() { return v.begin(); }
	typename std::vector<T>::iterator end() { return v.end(); }
	typename std::vector<T>::const_iterator begin() const { return v.begin(); }
	typename std::vector<T>::const_iterator end() const { return v.end(); }
private:
	std::vector<T> v;
};

// physics/jscol_aabbox.h, as used by setCurentRoomDimensions (AudioEngine.cpp:753-760).
struct AABBox {
	Vec4f min_, max_;
	float axisLength(int i) const { return max_[i] - min_[i]; }
	Vec4f centroid() const { return Vec4f((min_[0] + max_[0]) * 0.5f, (min_[1] + max_[1]) * 0.5f, (min_[2] + max_[2]) * 0.5f, 1.f); }
};
}  // namespace js

// utils/CircularBuffer.h: FIFO of T; read from the front, enqueue to the back (AudioEngine.cpp:212,464,470,543).
template <class T> class CircularBuffer {
public:
	CircularBuffer() : head(0), count(0) {}
	size_t size() const { return count; }
	bool empty() const { return count == 0; }
	void clear() { head = 0; count = 0; }
	void pushBackNItems(const T* src, size_t n) {
		if (count + n > buf.size()) grow(count + n);
		for (size_t i = 0; i < n; ++i) buf[(head + count + i) % buf.size()] = src[i];
		count += n;
	}
	void push_back(const T& t) { pushBackNItems(&t, 1); }
	void popFrontNItems(T* dst, size_t n) {
		for (size_t i = 0; i < n; ++i) dst[i] = buf[(head + i) % buf.size()];
		popFrontNItems(n);
	}
	void popFrontNItems(size_t n) { head = buf.empty() ? 0 : (head + n) % buf.size(); count -= n; }
private:
	void grow(size_t need) {
		std::vector<T> nb(need * 2 + 16);
		for (size_t i = 0; i < count; ++i) nb[i] = buf[(head + i) % buf.size()];
		buf.swap(nb);
		head = 0;
	}
	std::vector<T> buf;
	size_t head, count;
};

// utils/Mutex.h, utils/Lock.h.  glare-core's Mutex is recursive (CRITICAL_SECTION / a recursive pthread mutex):
// addSourceFromStreamingSoundFile holds AudioEngine::mutex while it calls addSource, which takes it again
// (AudioEngine.cpp:1085,1137,892).
class Mutex {
public:
	void lock() { m.lock(); }
	void unlock() { m.unlock(); }
private:
	std::recursive_mutex m;
};
class Lock {
public:
	explicit Lock(Mutex& m_) : m(m_) { m.lock(); }
	~Lock() { m.unlock(); }
	Lock(const Lock&) = delete;
	Lock& operator=(const Lock&) = delete;
private:
	Mutex& m;
};
#ifndef GUARDED_BY
#define GUARDED_BY(x)
#endif
#ifndef GLARE_ALIGNED_16_NEW_DELETE
#define GLARE_ALIGNED_16_NEW_DELETE
#endif

// utils/AtomicInt.h
namespace glare {
class AtomicInt {
public:
	AtomicInt(int64_t v = 0) : val(v) {}
	operator int64_t() const { return val.load(); }
	AtomicInt& operator=(int64_t v) { val.store(v); return *this; }
	int64_t increment() { return val.fetch_add(1); }
	int64_t decrement() { return val.fetch_sub(1); }
private:
	std::atomic<int64_t> val;
};
}  // namespace glare

// utils/MessageableThread.h + utils/ThreadManager.h, as AudioEngine uses them (AudioEngine.cpp:340-345,557-560,711-726,816):
// threads are objects with doRun() / kill(); the manager starts them and, on killThreadsBlocking(), asks each to stop and joins.
class MessageableThread : public ThreadSafeRefCounted {
public:
	virtual ~MessageableThread() {}
	virtual void doRun() = 0;
	virtual void kill() = 0;
};
class ThreadManager {
public:
	~ThreadManager() { killThreadsBlocking(); }
	void addThread(const Reference<MessageableThread>& t) {
		objs.push_back(t);
		MessageableThread* raw = t.ptr();
		threads.emplace_back([raw]() { raw->doRun(); });
	}
	void killThreadsBlocking() {
		for (auto& o : objs) o->kill();
		for (auto& th : threads) if (th.joinable()) th.join();
		threads.clear(); objs.clear();
	}
	size_t getNumThreads() const { return objs.size(); }
private:
	std::vector<Reference<MessageableThread>> objs;
	std::vector<std::thread> threads;
};

// utils/MemMappedFile.h (SharedMemMappedFile) and audio/MP3AudioStreamer.h (MP3AudioStreamerDataSource): opaque here -- the
// mp3 codec is outside the accelerated path; they only travel through AddSourceFromStreamingSoundFileParams.
class SharedMemMappedFile : public ThreadSafeRefCounted {};
class MP3AudioStreamerDataSource : public ThreadSafeRefCounted {};
