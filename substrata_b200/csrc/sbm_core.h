// sbm_core.h -- data layout in HBM and the exact-arithmetic scalar helpers shared by the host and the kernels.
//
// Everything here is `__host__ __device__` so that (a) the per-source parameter kernel (k1_params.cu) and
// (b) the host-side emulation used by the CPU tests (tests/native/emu.cc -- test infrastructure, not shipped)
// execute literally the same statements.  Compile host users with -ffp-contract=off and the device user with
// -fmad=false: the reference is an SSE4.1 build without FMA (cmake/shared_cxx_settings.cmake:271) and several
// results feed integer truncations or thresholds, so operation order and rounding must match.
//
// "RA/" = resonance-audio/resonance_audio/ in the reference tree; "AE.cpp" = audio/AudioEngine.cpp;
// "AR.cpp" = audio/AudioResampler.cpp.
#pragma once

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define SBM_HD __host__ __device__ __forceinline__
#else
#define SBM_HD inline
#endif

namespace sbm {

// ------------------------------------------------------------------------------------------------
// Constants (RA/base/constants_and_types.h:80,84,156,168).
// ------------------------------------------------------------------------------------------------
constexpr float kNeg60dB = 0.001f;
constexpr float kEpsilonFloat = 1e-6f;
constexpr float kNearFieldThreshold = 1.0f;
constexpr float kUnitRampLength = 2048.0f;
constexpr float kMaxDistance = 500.0f;             // RA/base/source_parameters.h:62 (Substrata never changes it)
constexpr float kOcclusionSmoothing = 0.75f;       // RA/graph/occlusion_node.cc:33
constexpr int kNumSh = 16;                         // 3rd-order ambisonics (RA/config/source_config.cc:54-63)
constexpr int kShRaw = 15;
constexpr int kTableAz = 91, kTableEl = 91;        // RA/ambisonics/ambisonic_lookup_table.cc:31-35
constexpr int kBusRows = 19;                       // 16 ambisonic + stereo L/R + the mono reflections send (room effects)
constexpr int kRowRefl = 18;
constexpr int kMaxFrames = 1024;
constexpr double kFx = 4294967296.0;               // 2^32: scale of the 32.32 fixed-point read-head coordinates

// Source kinds (== vraudio::RenderingMode for objects) and feeder types.
constexpr uint32_t kKindPan = 0, kKindHrtf = 3, kKindStereo = 100;
constexpr uint32_t kFeedBuffer = 0, kFeedRing = 1, kFeedMix = 2;

// SrcStatic.flags
constexpr uint32_t kFlagAlive = 1u << 0;            // slot holds a live, active source
constexpr uint32_t kFlagLooping = 1u << 1;
constexpr uint32_t kFlagRemoveOnFinish = 1u << 2;
constexpr uint32_t kFlagFeedShift = 4;              // 2 bits
constexpr uint32_t kFlagKindShift = 8;              // 8 bits

// WorkRec.flags
constexpr uint32_t kRecSkip = 1u << 0;              // contributes nothing this buffer
constexpr uint32_t kRecResample = 1u << 1;          // linear SRC from `rate` to the engine rate (AR.cpp)
constexpr uint32_t kRecMix = 1u << 2;               // pitch/doppler mix sources (AE.cpp:435-458)
constexpr uint32_t kRecLpf = 1u << 3;               // occlusion low-pass active
constexpr uint32_t kRecStereo = 1u << 4;            // non-spatial: L = R = y
constexpr uint32_t kRecHrtf = 1u << 5;              // ambisonic encode (16 channels) instead of pan (2)
constexpr uint32_t kRecWrap = 1u << 6;              // window index wraps modulo len (looping buffer / ring)
constexpr uint32_t kRecMix1 = 1u << 7;              // exactly one mix source, expressed as a window + fixed-point read head
constexpr uint32_t kRecRampCareful = 1u << 8;       // ambisonic: a channel ramp ends inside the buffer on a value the -60 dB rule then zeroes (konst != end value)

// ------------------------------------------------------------------------------------------------
// HBM layout.  One slot per live source; SoA arrays indexed by slot, written by the host (params) or
// by the kernels (state).  Per buffer K1 condenses a slot into one 128-byte WorkRec that K2 consumes.
// ------------------------------------------------------------------------------------------------
struct __attribute__((aligned(16))) F4 { float x, y, z, w; };   // position xyz + volume
struct __attribute__((aligned(16))) U4 { uint32_t x, y, z, w; };

struct SrcStatic {              // 32 B, host-written on attach
	const float* pcm;           // device pointer to the source's PCM (shared buffer, or ring storage)
	uint32_t len;               // samples in pcm (loop length / ring capacity)
	int32_t rate;               // AudioSource::sampling_rate
	uint32_t flags;             // kFlag*
	uint32_t mix_first;         // first entry in the MixZ array
	uint32_t nz;                // number of mix sources
	uint32_t id;                // public source id (for the finished list)
};

struct MixZ {                   // 32 B: glare::MixSource (AudioEngine.h:61-69), host-written part
	const float* pcm;
	uint32_t len;
	float mix_factor;
	double delta;               // source_delta
	double pad;
};
// sound_file_i lives in two device arrays indexed like MixZ: mix_pos (current) and mix_pos0 (value before this
// buffer's advance, which K2 evaluates the window from).

struct Resamp {                 // 24 B: glare::AudioResampler state (AudioResampler.h:37-41)
	int64_t prev_dest_dst_coords;
	int64_t prev_samples_0_src_coords;
	float prev_samples[2];
};

struct RampDesc {               // gain over one buffer: i < n ? ramp(start, inc, i) : konst
	float konst;                // effective multiplier of the constant part (0 / 1 / gain by the -60 dB rules)
	float start;
	float inc;
	uint32_t n;                 // ramp samples inside this buffer (0 = no ramp)
};

struct __attribute__((aligned(128))) WorkRec {   // 128 B, K1 -> K2, one per entry of the render lists (list order)
	const float* pcm;           //  0  window source (shared buffer, ring storage, or the single mix source's buffer)
	uint32_t len;               //  8  samples in pcm (loop length / ring capacity)
	uint32_t start;             // 12  index in pcm of window sample 0
	uint32_t n_src;             // 16  window samples (= source samples consumed this buffer for buffer / ring feeds)
	uint32_t avail;             // 20  window samples that exist; beyond: zeros (end of buffer / ring underflow)
	uint32_t flags;             // 24  kRec*
	uint32_t slot;              // 28
	unsigned long long x0_fx;   // 32  32.32 fixed point: coordinate of output sample 0 in {prev0, prev1, window...} (resample,
	                            //     AR.cpp:101-103) or in the window (single mix source, AE.cpp:443-452)
	unsigned long long step_fx; // 40  32.32 fixed point: coordinate step per output sample (src_rate / dst_rate, or source_delta)
	float prev0, prev1;         // 48  resampler history
	float lpf_coef, lpf_prev;   // 56
	RampDesc direct;            // 64  GainNode (kDirect / kInput)
	RampDesc ch[2];             // 80  pan L/R (pan mode only; ambisonic gains live in ShGains / ShRamps)
	uint32_t ch_ramp_mask;      // 112 bit c set: bus channel c ramps this buffer
	float mix_factor;           // 116 kRecMix1: MixSource::mix_factor
	uint32_t mix_first;         // 120 kRecMix: first entry in the MixZ array
	uint32_t nz;                // 124 kRecMix: number of mix sources
};

// What the mix kernel needs to render one source on its lean paths, in final form (128 B, one per entry of the render lists,
// written by K1 -- k1_params.cu, or the parameter warps of the fused kernel -- and streamed into K2's record ring by TMA).
// The full WorkRec is only read back (from global memory) by the rarely taken careful paths.
struct __attribute__((aligned(16))) MixRec {
	uint32_t mode;              //   0  kMd*
	uint32_t nramp;             //   4  frames from the start of the buffer in which some gain of this source still ramps
	uint32_t w_off;             //   8  float offset, inside the source's stage slot, of element 0 of the array the read head indexes:
	                            //      window sample 0, or two before it for the resampler's {prev0, prev1, window...}
	float gd;                   //  12  direct gain * mix factor (constant part)
	unsigned long long x0, step;   //  16  32.32 read head
	float g0, g1;               //  32  stereo kernel: direct gain * mix factor * bus gain (constant parts)
	uint32_t n1b;               //  40  index (same coordinates as w_off) from which `wadj` applies (loop wrap), 0xffffffff = never
	uint32_t wadj;              //  44
	uint32_t slot;              //  48
	float mfx;                  //  52  MixSource::mix_factor where the lean paths must apply it to the samples themselves (ramping direct gain), else 1
	float prev0, prev1;         //  56  AudioResampler::prev_samples
	float d_konst, d_start, d_inc;   //  64  GainNode ramp
	uint32_t dn_nsrc;           //  76  low 16 bits: ramp frames of the direct gain; high 16: window samples (n_src)
	float c0_konst, c0_start, c0_inc;   //  80  stereo bus gain L
	uint32_t cn;                //  92  low 16 bits: ramp frames of bus gain L; high 16: of bus gain R
	float c1_konst, c1_start, c1_inc;   //  96  stereo bus gain R
	uint32_t pad;               // 108
	const float* pcm;           // 112  fetch plan: TMA segment 1 = pcm[a0 .. a0 + seg1), segment 2 = pcm[0 .. seg2)
	uint32_t a0;                // 120
	uint32_t segs;              // 124  low 16 bits: seg1 floats; high 16: seg2 floats; 0 = window not staged (careful path reads global memory)
};
static_assert(sizeof(MixRec) == 128, "MixRec is 128 bytes");
constexpr uint32_t kMdSkip = 1, kMdGeneric = 2, kMdInterp = 4, kMdAdjust = 8, kMdResample = 16, kMdLpf = 32, kMdStereo = 64,
                   kMdRamp = 128, kMdShRamp = 256;

// Room effects (RA/graph/graph_manager.cc:144-151): the ramped gain of a sound object's send into the mono reflections mixer
// (GainMixerNode of AttenuationType::kReflections), 16 B per source, written by K1 and read by K2 only while room effects are on.
struct ReflSend { float konst, start, inc; uint32_t n; };

struct ShGains { float konst[kNumSh]; };                       // 64 B per HRTF source: effective constant gains
// Channel ramps of an ambisonic source (read only when ch_ramp_mask != 0; all 48 entries are written then): the gain of bus
// channel c at frame i is  konst[c] + (i < nf[c] ? stm[c] + i * inc[c] : 0)  with stm = ramp start - konst, nf = ramp frames in
// this buffer as a float (0 for a channel that does not ramp).  Where the ramp ends inside the buffer on konst itself (every case
// except the -60 dB rule zeroing the end value, kRecRampCareful) that is also  konst + stm + inc * min(i, nf)  -- branch-free.
struct ShRamps { float stm[kNumSh]; float inc[kNumSh]; float nf[kNumSh]; };

// Per-engine scalars that change every buffer (passed by value to K1).
struct Globals {
	float head_pos[3];
	float master;
	float head_q[4];            // x y z w
	uint32_t frames;
	int32_t rate;
	uint32_t parity;            // buffer index & 1
	uint32_t n_slots;
	uint32_t n_pan, n_sh;       // render-list lengths
	uint32_t room;              // room effects enabled: sound objects also feed the reflections send
	float refl_gain;            // ReflectionProperties::gain
	uint32_t list_identity;     // render-list entry k is slot k + list_base (sources created in one run, none removed): K1 skips the list load
	uint32_t list_base;
	uint32_t mix_identity;      // ... and its first mix source is MixZ k + mix_base (every listed source has mix sources, allocated in one run)
	uint32_t mix_base;
	uint32_t epoch;             // low 32 bits of buffer index + 1 (never 0): the value K1 stamps its finished tiles with
	uint32_t overlap;           // K1 runs as a persistent grid beside the mix kernel and publishes SlotArrays::ready tile by tile
};

constexpr uint32_t kK1Tile = 128;       // render-list entries per K1 tile (one thread each)

// Device-side flags / counters.
struct DevFlags {
	int pan_used[2];            // GainMixer of the stereo panner received >= 1 input (by buffer parity)
	int sh_used[2];             // GainMixer of the ambisonic encoder
	int st_used[2];             // a non-spatial source reached the stereo mixer
	int sh_flush_frames;        // AmbisonicBinauralDecoderNode::num_frames_processed_on_empty_input_
	int refl_used[2];           // the reflections send mixer received >= 1 input (by buffer parity)
	int foa_valid;              // the ReflectionsNode produced a first-order buffer this cycle (room kernel -> decode / finish)
	int sh1_flush_frames;       // the order-1 decoder's num_frames_processed_on_empty_input_
	int n_finished;             // entries in the finished list
	unsigned long long algo_bytes;  // accounting: sum of 4*n_src over rendered sources
	unsigned int n_rendered;
	int peer_timeout;               // set by the gather kernel when a peer rank's bus did not arrive in time (reported and cleared by K4b)
	int pipe_timeout;               // set by the mix kernel when K1's records of a tile did not appear in time (overlapped K1 / K2)
	// K1 spreads its per-warp accounting atomics over kStatParts padded slots (same-address atomics serialise in L2;
	// a block-wide reduction would hold every warp of a block until its slowest one is done); K4b folds them into the tail.
	unsigned long long part_bytes[32][4];      // [part][0] used, rest = 32-byte padding
	unsigned int part_rendered[32][8];
};
constexpr uint32_t kStatParts = 32;

// All per-slot arrays (device pointers), passed by value to K1.
struct SlotArrays {
	const SrcStatic* stat;      // host-written
	const F4* posvol;           // x y z volume
	const float* occl;          // occlusion intensity target
	const unsigned long long* ring_write;   // samples pushed so far (ring sources)
	uint32_t* live;             // device copy of "alive" (cleared by the feeder when a source stops)
	unsigned long long* read_i; // cur_read_i / ring read total
	long long* eof;
	Resamp* rs;
	float* dgain;               // GainNode current gain
	float* occ_cur;
	float* lpf_prev;
	float* cgain;               // GainMixer processor gains, quad-major: channels 4q .. 4q+3 of slot s = 16 bytes at float index 4 * (q * cstride + s)
	uint8_t* bits;              // bit0 direct-gain initialised, bit1 mixer processors exist, bit2 active
	const MixZ* mixz;
	double* mix_pos;
	double* mix_pos0;
	WorkRec* rec;               // [n_pan + n_sh] in render-list order (pan list first)
	MixRec* mrec;               // indexed like rec (CUDA build; the host emulation mixes straight from rec)
	ShGains* shg;               // indexed like rec
	ShRamps* shr;               // indexed like rec
	ReflSend* rfl;              // indexed like rec (room effects)
	float* rgain;               // [slot] the reflections mixer's processor gain of the source
	uint32_t* ready;            // [render-list index / kK1Tile] == Globals::epoch once K1 has written that tile's records
	const uint32_t* list_mix;   // [render-list index] SrcStatic::mix_first of the entry's slot (0xffffffff: no mix sources): lets K1 fetch the
	                            // first MixZ and read head together with the slot's state instead of one dependent round trip later
	int* finished;              // list of public ids stopped by the feeder
	int cstride;                // slots per quad plane of cgain (= the engine's slot capacity)
};

constexpr uint8_t kBitDirectInit = 1, kBitMixExists = 2, kBitMixActive = 4, kBitReflExists = 8, kBitReflActive = 16;

// ------------------------------------------------------------------------------------------------
// Bit helpers.
// ------------------------------------------------------------------------------------------------
SBM_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
	return __float_as_uint(f);
#else
	uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
SBM_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
	return __uint_as_float(u);
#else
	float f; memcpy(&f, &u, 4); return f;
#endif
}
// Single IEEE operations that must never be contracted into an FMA, whatever the translation unit's flags.
#if defined(__CUDA_ARCH__)
SBM_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
SBM_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
SBM_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
SBM_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
SBM_HD float fsqrt(float a) { return __fsqrt_rn(a); }
#else
SBM_HD float fadd(float a, float b) { return a + b; }
SBM_HD float fsub(float a, float b) { return a - b; }
SBM_HD float fmul(float a, float b) { return a * b; }
SBM_HD float fdiv(float a, float b) { return a / b; }
SBM_HD float fsqrt(float a) { return sqrtf(a); }
#endif

SBM_HD bool near_zero(float g) { return fabsf(g) < kNeg60dB; }                 // RA/dsp/gain.cc:94-96
SBM_HD bool near_unity(float g) { return fabsf(fsub(1.0f, g)) < kNeg60dB; }    // RA/dsp/gain.cc:98-100

// ------------------------------------------------------------------------------------------------
// ramp_advance: the value of `g` after n iterations of the float recurrence  g = g + inc
// (LinearGainRamp, RA/dsp/gain.cc:38-47), bit for bit, in O(#binades crossed) instead of O(n).
//
// While g and the exact sum g+inc lie in one binade [2^e, 2^(e+1)) the rounded step is a constant integer
// number q of ulps: with M the 24-bit significand of |g| and a = |inc|/ulp(g) = k + r (k integer, 0 <= r < 1),
// round-to-nearest-even gives q = k (r < 1/2), k+1 (r > 1/2) and, on an exact tie, the value that makes M even
// (after one tie step M is even and stays even, so q = k + (k & 1)).  A run of J such steps is M -> M +- J*q in
// integer arithmetic; steps that leave the binade (or start from 0 / denormals) are taken with one real float add,
// which is always exact by construction.
// ------------------------------------------------------------------------------------------------
SBM_HD float ramp_advance(float g, float inc, uint32_t n) {
	const uint32_t ib = f2u(inc);
	const int ei = (int)((ib >> 23) & 0xff);
	if (ei == 0 || ei == 255) {                          // inc is 0, denormal, inf or nan: plain loop semantics
		if ((ib & 0x7fffffffu) == 0) return g;
		for (; n > 0; --n) g = fadd(g, inc);
		return g;
	}
	const uint32_t mi = (ib & 0x7fffffu) | 0x800000u;    // significand of |inc|
	while (n > 0) {
		const uint32_t gb = f2u(g);
		const int eg = (int)((gb >> 23) & 0xff);
		const int s = eg - ei;
		uint32_t J = 0;
		const uint32_t M = (gb & 0x7fffffu) | 0x800000u;
		const bool grow = ((gb ^ ib) >> 31) == 0;        // same sign: |g| increases
		// Shrinking from the very bottom of a binade lands in the finer-grained binade below: real step.
		const bool bottom = !grow && M == 0x800000u;
		if (eg >= 1 && eg < 254 && s >= 1 && !bottom) {
			if (s > 24) return g;                        // |inc| < ulp/2: the sum rounds back to g forever
			const uint32_t k = mi >> s;
			const uint32_t rem = mi & ((1u << s) - 1u);
			const uint32_t half = 1u << (s - 1);
			const bool tie = rem == half;
			uint32_t q;
			bool ok = true;
			if (!tie) q = k + (rem > half ? 1u : 0u);
			else if ((M & 1u) == 0) q = k + (k & 1u);
			else { q = 0; ok = false; }                  // odd M on a tie: take one real step first
			if (ok) {
				if (q == 0) return g;                    // stuck
				if (grow) {
					// all J steps need  M_{j-1} + k + 1 <= 2^24
					const int64_t room = (int64_t)0x1000000 - (int64_t)k - 1 - (int64_t)M;
					if (room >= 0) J = (uint32_t)(room / q) + 1u;
				} else {
					// all J steps need  M_{j-1} - k - 1 >= 2^23
					const int64_t room = (int64_t)M - (int64_t)k - 1 - (int64_t)0x800000;
					if (room >= 0) J = (uint32_t)(room / q) + 1u;
				}
				if (J > n) J = n;
				if (J > 0) {
					const uint32_t Mj = grow ? M + J * q : M - J * q;
					uint32_t bits = gb & 0x80000000u;
					if (Mj == 0x1000000u) bits |= (uint32_t)(eg + 1) << 23;
					else bits |= ((uint32_t)eg << 23) | (Mj & 0x7fffffu);
					g = u2f(bits);
					n -= J;
					continue;
				}
			}
		}
		const float g2 = fadd(g, inc);                    // one real step (binade crossing, near zero, tie with odd M ...)
		--n;
		if (f2u(g2) == gb) return g;                      // fixed point
		g = g2;
	}
	return g;
}

// ------------------------------------------------------------------------------------------------
// atan2f / atanf: operation-for-operation restatement of the fdlibm-lineage float implementation that the
// reference links against (glibc 2.39 libm: sysdeps/ieee754/flt-32/e_atan2f.c, s_atanf.c; pinned version =
// the image's glibc "2.39-0ubuntu8.5").  The integer degree of these angles selects a lookup-table cell
// (RA/ambisonics/ambisonic_lookup_table.cc:107-117), so the result has to be the same float, not a close one.
// tests/test_host_math.py checks bit equality against the live libm on 10^7+ inputs.
// ------------------------------------------------------------------------------------------------
SBM_HD float exact_atanf(float x) {
	const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
	const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
	const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f, 9.0908870101e-02f,
	                      -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f, 4.9768779427e-02f, -3.6531571299e-02f,
	                      1.6285819933e-02f};
	const int32_t hx = (int32_t)f2u(x);
	const int32_t ix = hx & 0x7fffffff;
	int id;
	if (ix >= 0x4c800000) {
		if (ix > 0x7f800000) return fadd(x, x);
		return hx > 0 ? fadd(atanhi[3], atanlo[3]) : fsub(-atanhi[3], atanlo[3]);
	}
	if (ix < 0x3ee00000) {
		if (ix < 0x31000000) return x;                 // |x| < 2^-29 (1e30 + x > 1 always holds for finite x)
		id = -1;
	} else {
		x = fabsf(x);
		if (ix < 0x3f980000) {
			if (ix < 0x3f300000) { id = 0; x = fdiv(fsub(fmul(2.0f, x), 1.0f), fadd(2.0f, x)); }
			else { id = 1; x = fdiv(fsub(x, 1.0f), fadd(x, 1.0f)); }
		} else {
			if (ix < 0x401c0000) { id = 2; x = fdiv(fsub(x, 1.5f), fadd(1.0f, fmul(1.5f, x))); }
			else { id = 3; x = fdiv(-1.0f, x); }
		}
	}
	const float z = fmul(x, x);
	const float w = fmul(z, z);
	const float s1 = fmul(z, fadd(aT[0], fmul(w, fadd(aT[2], fmul(w, fadd(aT[4], fmul(w, fadd(aT[6], fmul(w, fadd(aT[8], fmul(w, aT[10])))))))))));
	const float s2 = fmul(w, fadd(aT[1], fmul(w, fadd(aT[3], fmul(w, fadd(aT[5], fmul(w, fadd(aT[7], fmul(w, aT[9])))))))));
	if (id < 0) return fsub(x, fmul(x, fadd(s1, s2)));
	const float r = fsub(atanhi[id], fsub(fsub(fmul(x, fadd(s1, s2)), atanlo[id]), x));
	return hx < 0 ? -r : r;
}

SBM_HD float exact_atan2f(float y, float x) {
	const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
	            pi_lo = -8.7422776573e-08f;
	const int32_t hx = (int32_t)f2u(x), hy = (int32_t)f2u(y);
	const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
	if (ix > 0x7f800000 || iy > 0x7f800000) return fadd(x, y);
	if (hx == 0x3f800000) return exact_atanf(y);
	const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
	if (iy == 0) {
		if (m < 2) return y;
		return m == 2 ? fadd(pi, tiny) : fsub(-pi, tiny);
	}
	if (ix == 0) return hy < 0 ? fsub(-pi_o_2, tiny) : fadd(pi_o_2, tiny);
	if (ix == 0x7f800000) {
		if (iy == 0x7f800000) {
			switch (m) {
				case 0: return fadd(pi_o_4, tiny);
				case 1: return fsub(-pi_o_4, tiny);
				case 2: return fadd(fmul(3.0f, pi_o_4), tiny);
				default: return fsub(fmul(-3.0f, pi_o_4), tiny);
			}
		}
		switch (m) {
			case 0: return 0.0f;
			case 1: return -0.0f;
			case 2: return fadd(pi, tiny);
			default: return fsub(-pi, tiny);
		}
	}
	if (iy == 0x7f800000) return hy < 0 ? fsub(-pi_o_2, tiny) : fadd(pi_o_2, tiny);
	const int k = (iy - ix) >> 23;
	float z;
	if (k > 60) z = fadd(pi_o_2, fmul(0.5f, pi_lo));
	else if (hx < 0 && k < -60) z = 0.0f;
	else z = exact_atanf(fabsf(fdiv(y, x)));
	switch (m) {
		case 0: return z;
		case 1: return -z;
		case 2: return fsub(pi, fsub(z, pi_lo));
		default: return fsub(fsub(z, pi_lo), pi);
	}
}

// ------------------------------------------------------------------------------------------------
// Scalar DSP parameter helpers.
// ------------------------------------------------------------------------------------------------
// ComputeLogarithmicDistanceAttenuation with Substrata's defaults min 0 / max 500 (RA/dsp/distance_attenuation.cc:26-52).
SBM_HD float log_distance_attenuation(const float* listener, const float* source) {
	const float dx = fsub(listener[0], source[0]), dy = fsub(listener[1], source[1]), dz = fsub(listener[2], source[2]);
	const float distance = fsqrt(fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz)));
	if (distance > kMaxDistance) return 0.0f;
	const float min_allowed = kNearFieldThreshold;      // max(min_distance = 0, kNearFieldThreshold)
	if (distance > min_allowed) {
		// interval = 500 - 1, att_max = 1 / (1 + interval), 1 - att_max: constants of the reference's expression, folded
		// at compile time with IEEE float arithmetic (identical values)
		constexpr float interval = kMaxDistance - kNearFieldThreshold;
		constexpr float att_max = 1.0f / (1.0f + interval);
		constexpr float one_minus_att_max = 1.0f - att_max;
		const float relative = fsub(distance, min_allowed);
		const float att = fdiv(1.0f, fadd(relative, 1.0f));
		return fdiv(fsub(att, att_max), one_minus_att_max);
	}
	return 1.0f;
}

// GetRelativeDirection (RA/base/misc_math.h:143-150) with Eigen's quaternion-vector product order
// (third_party/eigen/Eigen/src/Geometry/Quaternion.h:531-541), then SphericalAngle::FromWorldPosition
// (RA/base/spherical_angle.cc:42-49).
// GetRelativeDirection (RA/base/misc_math.h:143-150) with Eigen's quaternion-vector product order
// (third_party/eigen/Eigen/src/Geometry/Quaternion.h:531-541): r = conj(q_head) * (source - listener).
SBM_HD void source_relative(const float* head_pos, const float* q_xyzw, const float* src_pos, float* r) {
	const float v0 = fsub(src_pos[0], head_pos[0]), v1 = fsub(src_pos[1], head_pos[1]), v2 = fsub(src_pos[2], head_pos[2]);
	const float a0 = -q_xyzw[0], a1 = -q_xyzw[1], a2 = -q_xyzw[2], w = q_xyzw[3];
	float u0 = fsub(fmul(a1, v2), fmul(a2, v1));
	float u1 = fsub(fmul(a2, v0), fmul(a0, v2));
	float u2 = fsub(fmul(a0, v1), fmul(a1, v0));
	u0 = fadd(u0, u0); u1 = fadd(u1, u1); u2 = fadd(u2, u2);
	const float c0 = fsub(fmul(a1, u2), fmul(a2, u1));
	const float c1 = fsub(fmul(a2, u0), fmul(a0, u2));
	const float c2 = fsub(fmul(a0, u1), fmul(a1, u0));
	r[0] = fadd(fadd(v0, fmul(w, u0)), c0);
	r[1] = fadd(fadd(v1, fmul(w, u1)), c1);
	r[2] = fadd(fadd(v2, fmul(w, u2)), c2);
}

// ... then SphericalAngle::FromWorldPosition (RA/base/spherical_angle.cc:42-49).  The integer degree of these angles
// selects a lookup-table cell (ambisonic sources), so they must be the reference's floats bit for bit.
SBM_HD void source_direction(const float* head_pos, const float* q_xyzw, const float* src_pos, float* az, float* el) {
	float r[3];
	source_relative(head_pos, q_xyzw, src_pos, r);
	*az = exact_atan2f(-r[0], -r[2]);
	*el = exact_atan2f(r[1], fsqrt(fadd(fmul(r[0], r[0]), fmul(r[2], r[2]))));
}

// CalculateOcclusionFilterCoefficient with directivity 1 (alphas stay 0 in Substrata) + MonoPoleFilter::SetCoefficient
// clamp (RA/dsp/occlusion_calculator.cc:43-50, RA/dsp/mono_pole_filter.cc:30-32).
SBM_HD float occlusion_coefficient(float occlusion) {
	if (occlusion == 0.0f) return 0.0f;      // the expression below gives exactly 0 for 0 (1 - 1/1): skip its division
	float b = fadd(occlusion, 1.0f);
	b = fmul(b, b);
	b = fmul(b, b);
	const float factor = fdiv(1.0f, b);
	const float c = fmaxf(0.0f, fsub(1.0f, fmul(1.0f, factor)));
	return fmaxf(fminf(c, 1.0f), 0.0f);
}

// One GainProcessor::ApplyGain over a buffer of F frames, reduced to a descriptor (RA/dsp/gain_processor.cc:34-85).
// `cur` is the processor's gain on entry, updated to its gain on exit.  accumulate selects the GainMixer rules
// (near-unity not special-cased).  Returns the descriptor K2 evaluates.
SBM_HD RampDesc plan_gain(float* cur, bool initialised, float target, uint32_t F, bool accumulate) {
	RampDesc d;
	float c = initialised ? *cur : target;                       // first call: Reset(target), no ramp
	const uint32_t ramp = (uint32_t)(fmul(fabsf(fsub(target, c)), kUnitRampLength));
	d.start = c; d.inc = 0.0f; d.n = 0;
	if (ramp > 0) {
		d.n = ramp < F ? ramp : F;
#if defined(__CUDA_ARCH__)
		// On the device: increment and end value in closed form.  Both are continuous in their inputs (the end value feeds
		// the next buffer's ramp length only through |target - c| * 2048, which is 0 for a finished ramp either way), and
		// they differ from the reference's n running additions by < n ulp -- while K1's duration is the length of its
		// dependent instruction chain, of which a bit-exact division and up to 48 chained additions (or the closed-form
		// binade walk) per ramping gain, 17 gains per ambisonic source, were the largest part.
		d.inc = __fdividef(fsub(target, c), (float)ramp);
		c = fmaf((float)d.n, d.inc, c);
#else
		d.inc = fdiv(fsub(target, c), (float)ramp);
		// the value after d.n running additions, bit for bit: literally for short ramps (the usual few frames while the
		// listener moves), through the closed form for long ones
		if (d.n <= 48u) { for (uint32_t j = 0; j < d.n; ++j) c = fadd(c, d.inc); }
		else c = ramp_advance(c, d.inc, d.n);
#endif
	} else {
		c = target;
	}
	*cur = c;
	if (near_zero(c)) d.konst = 0.0f;
	else if (near_unity(c) && !accumulate) d.konst = 1.0f;
	else d.konst = c;
	return d;
}

}  // namespace sbm
