// ssb_kernels.cuh — sm_100a kernels of the spectral path-tracing hot path.
//
// Replaces Renderer::_render_pixel / _render_sample and everything they reach
// (reference src/renderer.cpp:103-308).  fp32 throughout (+ the f64
// islands the reference has: camera ray, zero-barycentric fallback, accumulation), compiled with
// -fmad=false so that no multiply-add is contracted: operation order is the reference's.
//
// Design (B200-first, not a translation of the reference's recursive std::function): a sorted WAVEFRONT.
// Per path depth, four launches on one stream (queue lengths stay on the device, no host round trip):
//   1. ssb_intersect_kernel  — the closest-hit query of every live path (depth 0: also creates the path: camera
//      ray, per-sample PCG32 seed, hero wavelength).  Small (64 registers, 32 warps/SM): the linear scan over the
//      quad list (ssb_isect.cuh) is a conservative plane/rectangle filter run converged by all lanes — two filter
//      entries per packed-fp32 instruction (FFMA2) — plus the exact watertight test on the surviving candidates,
//      nearest candidate first.  Writes the hit record, ends paths that miss, and counts hits per quad
//      (shared-memory histogram, one global atomic per quad per CTA).
//   2. ssb_bin_scan_kernel + 3. ssb_bin_scatter_kernel — counting sort of the hit paths by hit quad.
//   4. ssb_shade_kernel — everything the reference's lambda L does after the hit (emission, albedo / sRGB texture
//      upsampling, light sample, shadow query, BSDF sample, fold record), visiting paths grouped by hit quad so
//      that a warp shares material, light-sampling geometry class and branch behaviour; surviving paths are
//      stream-compacted (warp ballot + one atomic per warp) into dense records for the next depth.
// then ssb_fold_kernel (one thread per sample) folds the per-depth records backwards (the reference folds radiance
// on the way back UP its recursion, renderer.cpp:216,248 — kept exact) and converts to XYZ, and
// ssb_accumulate_kernel (one thread per pixel) adds sample*0.001f to the double XYZA accumulator in sample order
// (renderer.cpp:292-295): deterministic, no float/double atomics.
// History (profiles/): a register-resident megakernel reached 154-222 Msamples/s with 9.8 of 32 lanes active
// and instruction-cache thrash on 131 KB of SASS; the wavefront forms went from 500 to 840+.
// Path state lives in HBM as 32-byte (one sector) records: recA {origin, ignore | direction, lambda0} feeds the
// intersect stage, recH {hit point, quad | barycentrics} goes intersect -> shade, recR {PCG32 | sample id} feeds
// the shade stage.  The scene, materials, spectra, observer/basis tables, filter records and the sRGB LUT are one
// contiguous "blob" that each CTA pulls into shared memory with a single TMA bulk copy
// (cp.async.bulk.shared::cluster.global + mbarrier).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/ssb200.h"
#include "ssb_blob.hpp"
#include "ssb_isect.cuh"
#include "ssb_math.cuh"

namespace ssbk {

// (blob layout: ssb_blob.hpp; shared-memory view, ray/scene intersection: ssb_isect.cuh)

struct KParams {
	const unsigned char* blob;
	// Path state in HBM, one 32-byte sector per record so that every access moves whole sectors:
	//   recA[2] (ping-pong): ray        {origin.xyz, ignore (int bits)} {direction.xyz, lambda_0}   read by the intersect stage
	//   recR[2] (ping-pong): the rest   {PCG32 state, inc}              {sample id, lambda_0, -, -}  read by the shade stage
	//   recH               : closest hit {hit point.xyz, quad|tri<<31}  {barycentrics, -}            intersect -> shade
	// recA/recR are written densely (compacted) by the stage that creates the next ray; recH is indexed like the queue.
	float4* recA[2];
	float4* recR[2];
	float4* recH;
	// per-depth records for the backward fold, indexed [depth][sample id]: ONE 64-byte record (two whole sectors)
	//   {local radiance} {f_s} {n.l, pdf, -, -} {-}
	// per vertex.  The sample ids of a warp are scattered (paths are sorted by hit quad), so anything smaller than a sector
	// is a partial write that the L2 has to complete with a DRAM read before it can be written back: the three separate
	// 16/16/8-byte arrays of round 1 cost ~96 B written + ~96 B read per vertex for 40 B of payload (ncu: the shade launches
	// read 211 B per path for 68 B of records).  Whole records also mean the fold stage fetches no neighbour's bytes.
	float4* stk;
	// per-sample end of path, one 32-byte record: {value returned by the deepest L() call} {lambda_0, #records | hit<<16, -, -}
	float4* leaf;
	float* ff;          // dot(camera ray, camera dir), only when FLAT_FIELD_CORRECTION is off (renderer.cpp:265)
	uint32_t* counts;   // queue length per depth; counts[0] = samples in the pass
	// closest-hit records of the current depth (indexed like the input queue) and the sort-by-quad machinery
	uint32_t* hit_q;    // quad | tri << 31, or 0xffffffff for a miss (dense copy of recH[].w for the counting sort)
	uint32_t* order;    // queue positions of the paths that hit something, grouped by hit quad
	uint32_t* bin_count;   // [max_depth][SSB_MAX_QUADS] paths per hit quad
	uint32_t* bin_cursor;  // [max_depth][SSB_MAX_QUADS] scatter cursors (start at the bin offset)
	uint32_t* nhits;       // [max_depth] total hits (= length of `order`)
	float4* samples;    // [nsamp][npix_rect] per-sample (X,Y,Z,hit): fold stage -> in-order accumulation (aliases recA[0])
	double* accum;
	unsigned long long total_work;  // npix_rect * nsamp
	uint32_t width, height, x0, y0, rect_w, rect_h, sample_begin, nsamp;
	uint32_t indirect_only, upsampling, max_depth, els, flat_field;
	uint32_t render_mode;  // SSB_RENDER_SPECTRAL / SSB_RENDER_RGB
	uint32_t n_wavelengths;  // SAMPLE_WAVELENGTHS (2..4): channels >= n of every Hero are exactly 0
	uint32_t depth;     // depth processed by this launch
	uint32_t has_mirror;  // some material is a MaterialMirror: the shade stage then reads the incoming direction from recA
	uint32_t band_h, band_n, band_i;  // ssb_options.band_*: rows j with (j / band_h) % band_n == band_i (band_n <= 1: all rows of the rectangle)
	uint32_t scan_list;  // ssb_debug_intersect only: SSB_SCAN_LIST (the render kernels take the scan mode as a template parameter)
	float eps, lambda_min, lambda_step;
	unsigned long long seed;
	double pv_inv[16];
	float cam_pos[3], cam_dir[3];
	const float* jh_scale;
	const float* jh_data;
	uint32_t jh_res;
	uint32_t jh_prebaked;  // every DevTexture::coef is valid and holds the texels' JH coefficients
	const int32_t* meng_grid;
	const float* meng_points;
	uint32_t meng_grid_w, meng_grid_h, meng_npoints, meng_nsamples;
	float meng_xy_to_uv[6];
	float meng_sample_min, meng_sample_max;
};

struct Hero { float v[4]; };

#define SSB_PI_F 3.14159265358979323846f

// ------------------------------------------------------------------ small helpers (GLM scalar semantics)
__device__ __forceinline__ float glm_min(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float glm_max(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float glm_clamp(float x, float lo, float hi) { return glm_min(glm_max(x, lo), hi); }
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
	return (ax * bx + ay * by) + az * bz;
}

// a / b, exactly — with the one case that is both frequent and trivially known kept away from the IEEE division's slow
// path: a zero numerator over a positive (finite or infinite) denominator is that same zero.  (Channels of a hero sample
// that fall outside a measured spectrum — the Cornell box's albedos end at 700 nm — are exactly 0; ncu: the slow-path
// subroutine was entered from these quotients ~0.77 M times per launch, at 5-24 lanes.)  Lanes taking the shortcut divide
// 1 by 1.  The operands pass through an empty asm so that the optimiser cannot fold the selects back into "a / b, then
// select" (it did: r5a capture, FCHK on the raw numerator).
__device__ __forceinline__ float div_or_zero(float a, float b) {
	const bool zero = (a == 0.0f) && (b > 0.0f);
	float an = zero ? 1.0f : a, bn = zero ? 1.0f : b;
	asm("" : "+f"(an), "+f"(bn));
	const float q = an / bn;
	return zero ? a : q;
}
// the same for a divisor known at compile time (only the numerator needs hiding: the division keeps its constant reciprocal)
template <typename F>
__device__ __forceinline__ float div_const_or_zero(float a, F divide) {
	const bool zero = a == 0.0f;
	float an = zero ? 1.0f : a;
	asm("" : "+f"(an));
	const float q = divide(an);
	return zero ? a : q;
}

// ------------------------------------------------------------------ RNG (util/random.hpp:16-78)
struct Rng { unsigned long long state, inc; };
__device__ __forceinline__ uint32_t rng_next(Rng& r) {
	unsigned long long s = r.state;
	uint32_t xorshifted = (uint32_t)(((s >> 18u) ^ s) >> 27u);
	uint32_t rot = (uint32_t)(s >> 59u);
	uint32_t result = __funnelshift_r(xorshifted, xorshifted, rot);  // rotr32
	r.state = s * 6364136223846793005ull + r.inc;
	return result;
}
__device__ __forceinline__ float rand_1f(Rng& r) {  // libstdc++ generate_canonical<float,24>
	float ret = (float)rng_next(r) * 2.3283064365386963e-10f;  // /2^32, exact scaling
	if (ret >= 1.0f) ret = __uint_as_float(0x3f7fffffu);       // nextafter(1,0)
	return ret;
}
__device__ __forceinline__ double rand_1d(Rng& r) {  // generate_canonical<double,53>: (u0 + u1*2^32)/2^64
	double sum = (double)rng_next(r);
	sum += (double)rng_next(r) * 4294967296.0;
	double ret = sum * 5.421010862427522e-20;  // /2^64, exact scaling
	if (ret >= 1.0) ret = __longlong_as_double(0x3fefffffffffffffll);
	return ret;
}
__device__ __forceinline__ uint32_t rand_choice(Rng& r, uint32_t n) {  // Lemire, libstdc++ _S_nd
	unsigned long long product = (unsigned long long)rng_next(r) * (unsigned long long)n;
	uint32_t low = (uint32_t)product;
	if (low < n) {
		uint32_t threshold = (0u - n) % n;
		while (low < threshold) {
			product = (unsigned long long)rng_next(r) * (unsigned long long)n;
			low = (uint32_t)product;
		}
	}
	return (uint32_t)(product >> 32);
}
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

// Single, non-inlined copies of the exact libm restatements: the trace kernel calls them from ~20 sites and its
// instruction footprint must stay inside the instruction cache (ncu: `stalled_no_instruction` dominated v1).
#ifndef SSB_INLINE_MATH
// Measured (profiles/r5v_ab.txt, r5w_ab.txt; every call of a non-inlined function moves its arguments into fixed registers and
// anything passed by reference or returned as a struct through LOCAL MEMORY): inlining spec_hero4 +1.7 %, sincosf_x +0.3 %;
// acosf_x / acosf2_x / sinf_x inlined at their 3-4 call sites each cost 0.3-1 % (code size) and stay functions.
#define SSB_INLINE_MATH 24  // bit k set: inline 1 acosf_x, 2 acosf2_x, 4 sinf_x/cosf_x, 8 sincosf_x, 16 spec_hero4
#endif
#if !(SSB_INLINE_MATH & 1)
#define SSB_ATTR_ACOS __device__ __noinline__
#else
#define SSB_ATTR_ACOS __device__ __forceinline__
#endif
#if !(SSB_INLINE_MATH & 2)
#define SSB_ATTR_ACOS2 __device__ __noinline__
#else
#define SSB_ATTR_ACOS2 __device__ __forceinline__
#endif
#if !(SSB_INLINE_MATH & 4)
#define SSB_ATTR_SIN __device__ __noinline__
#else
#define SSB_ATTR_SIN __device__ __forceinline__
#endif
#if !(SSB_INLINE_MATH & 8)
#define SSB_ATTR_SINCOS __device__ __noinline__
#else
#define SSB_ATTR_SINCOS __device__ __forceinline__
#endif
#if !(SSB_INLINE_MATH & 16)
#define SSB_ATTR_SPEC __device__ __noinline__
#else
#define SSB_ATTR_SPEC __device__ __forceinline__
#endif
SSB_ATTR_ACOS float acosf_x(float x) { return ssbm::acosf_exact(x); }
// two arc cosines per call, evaluated with packed-fp32 instructions (ssbm::acosf_exact2: bit-identical per lane, ~45 % of
// the instructions of two scalar calls)
#ifndef SSB_ACOS_PAIRS
#define SSB_ACOS_PAIRS 1
#endif
SSB_ATTR_ACOS2 float2 acosf2_x(float x0, float x1) {
#if SSB_ACOS_PAIRS
	return ssbm::acosf_exact2(make_float2(x0, x1));
#else
	return make_float2(ssbm::acosf_exact(x0), ssbm::acosf_exact(x1));
#endif
}
SSB_ATTR_SIN float sinf_x(float x) { return ssbm::sinf_exact(x); }
SSB_ATTR_SIN float cosf_x(float x) { return ssbm::cosf_exact(x); }
// sin and cos of the same argument share glibc's argument reduction (ssbm::sincosf_exact): each result is the scalar
// function's, bit for bit, for ~40 % fewer instructions than two calls (measured: +0.8 % frame rate, profiles/r1p_tune.txt).
// (Returned by value: reference parameters of a non-inlined function would go through local memory.)
SSB_ATTR_SINCOS float2 sincosf_x(float x) {  // (sin, cos)
	float s, c;
	ssbm::sincosf_exact(x, &s, &c);
	return make_float2(s, c);
}

// _Spectrum::_sample_linear / _sample_nearest (spectrum.cpp:29-60)
__device__ __forceinline__ float spec_sample(const float* pool, const DevSpectrum& s, float lambda) {
	uint32_t n = s.n_filter & 0x7fffffffu;
	float i = (lambda - s.low) * s.recip;
	if (s.n_filter >> 31) {
		int ii = (int)roundf(i);
		return ((uint32_t)ii < n) ? pool[s.offset + ii] : 0.0f;
	}
	float i0f = floorf(i);
	float frac = i - i0f;
	int i0 = (int)i0f;
	int i1 = i0 + 1;
	float val0 = ((uint32_t)i0 < n) ? pool[s.offset + i0] : 0.0f;
	float val1 = ((uint32_t)i1 < n) ? pool[s.offset + i1] : 0.0f;
	return val0 * (1.0f - frac) + val1 * frac;
}
// _Spectrum::operator[] (spectrum.cpp:61-67)
SSB_ATTR_SPEC float4 spec_hero4(DevSpectrum s, float lambda_0, float step) {
	const float* pool = SceneView().pool();
	float4 h;
	h.x = spec_sample(pool, s, lambda_0 + 0.0f * step);
	h.y = spec_sample(pool, s, lambda_0 + 1.0f * step);
	h.z = spec_sample(pool, s, lambda_0 + 2.0f * step);
	h.w = spec_sample(pool, s, lambda_0 + 3.0f * step);
	return h;
}
// Three spectra tabulated on the same grid — the observer's xbar/ybar/zbar are columns of one CSV file (color.cpp:77-99) —
// share the index arithmetic of _sample_linear / _sample_nearest: the same operations on the same operands, evaluated once
// instead of three times.  Every returned value is spec_sample's, bit for bit.  Used by the fold stage (-0.13 ms per frame);
// the same sharing for the basis' r/g/b in the shade stage measured slower (+0.06 ms: profiles/r4b_ab_shared_grid.txt).
#ifndef SSB_SHARED_GRID
#define SSB_SHARED_GRID 1
#endif
__device__ __forceinline__ bool spec_same_grid(const DevSpectrum& a, const DevSpectrum& b, const DevSpectrum& c) {
	return a.n_filter == b.n_filter && a.n_filter == c.n_filter && a.low == b.low && a.low == c.low && a.recip == b.recip && a.recip == c.recip;
}
__device__ __forceinline__ void spec_sample3(const float* pool, const DevSpectrum& s0, const DevSpectrum& s1, const DevSpectrum& s2,
                                             float lambda, float& v0, float& v1, float& v2) {
	const uint32_t n = s0.n_filter & 0x7fffffffu;
	const float i = (lambda - s0.low) * s0.recip;
	if (s0.n_filter >> 31) {
		const int ii = (int)roundf(i);
		const bool in = (uint32_t)ii < n;
		v0 = in ? pool[s0.offset + ii] : 0.0f; v1 = in ? pool[s1.offset + ii] : 0.0f; v2 = in ? pool[s2.offset + ii] : 0.0f;
		return;
	}
	const float i0f = floorf(i);
	const float frac = i - i0f;
	const int i0 = (int)i0f, i1 = i0 + 1;
	const bool in0 = (uint32_t)i0 < n, in1 = (uint32_t)i1 < n;
	const float w0 = 1.0f - frac;
	v0 = (in0 ? pool[s0.offset + i0] : 0.0f) * w0 + (in1 ? pool[s0.offset + i1] : 0.0f) * frac;
	v1 = (in0 ? pool[s1.offset + i0] : 0.0f) * w0 + (in1 ? pool[s1.offset + i1] : 0.0f) * frac;
	v2 = (in0 ? pool[s2.offset + i0] : 0.0f) * w0 + (in1 ? pool[s2.offset + i1] : 0.0f) * frac;
}
// HeroSample = glm::vec<SAMPLE_WAVELENGTHS,float>: with fewer than 4 wavelengths the unused channels are held at exactly
// 0 (which every later per-channel operation and the pairwise dot product preserve).  spec_sample is bounds-checked,
// so sampling the unused wavelengths is harmless; they are cleared where a Hero enters the path state.  The test is
// uniform and false in the default configuration.
__device__ __forceinline__ void hero_clear_unused(const KParams& P, Hero& h) {
	if (P.n_wavelengths != 4u) {
		if (P.n_wavelengths < 3u) h.v[2] = 0.0f;
		h.v[3] = 0.0f;
	}
}
__device__ __forceinline__ Hero spec_hero(const KParams& P, const DevSpectrum& s, float lambda_0, float step) {
	float4 v = spec_hero4(s, lambda_0, step);
	Hero h;
	h.v[0] = v.x; h.v[1] = v.y; h.v[2] = v.z; h.v[3] = v.w;
	return h;
}

// ------------------------------------------------------------------ upsampling (util/color.cpp:166-232)
__device__ __forceinline__ int jh_find_interval(const float* values, int size_, float x) {  // rgb2spec.c:56-75
	int left = 0, last_interval = size_ - 2, size = last_interval;
	while (size > 0) {
		int half = size >> 1, middle = left + half + 1;
		if (values[middle] < x) { left = middle; size -= half + 1; }
		else size = half;
	}
	return left < last_interval ? left : last_interval;
}
// rgb2spec_fetch (rgb2spec.c:77-118), no FMA (parity build): l-RGB -> the three polynomial coefficients
__device__ __forceinline__ void jh_fetch(const float* __restrict__ jh_scale, const float* __restrict__ d, int res,
                                         float r, float g, float b, float coeff[3]) {
	// i = index of the largest component, later ones winning ties (rgb2spec.c:84-86), without a dynamically indexed local array
	int i = 0;
	float z = r;
	if (g >= z) { i = 1; z = g; }
	if (b >= z) { i = 2; z = b; }
	const float scale = (float)(res - 1) / z;
	const float x = (i == 0 ? g : (i == 1 ? b : r)) * scale, y = (i == 0 ? b : (i == 1 ? r : g)) * scale;
	uint32_t xu = (x != x) ? 0u : (uint32_t)x, yu = (y != y) ? 0u : (uint32_t)y;  // see oracle note on NaN
	uint32_t xi = min(xu, (uint32_t)(res - 2)), yi = min(yu, (uint32_t)(res - 2));
	uint32_t zi = (uint32_t)jh_find_interval(jh_scale, res, z);
	uint32_t offset = (((i * res + zi) * res + yi) * res + xi) * 3, dx = 3, dy = 3 * res, dz = 3 * res * res;
	float x1 = x - (float)xi, x0 = 1.f - x1, y1 = y - (float)yi, y0 = 1.f - y1;
	float z1 = (z - jh_scale[zi]) / (jh_scale[zi + 1] - jh_scale[zi]), z0 = 1.f - z1;
	for (int j = 0; j < 3; ++j) {
		coeff[j] = ((__ldg(d + offset) * x0 + __ldg(d + offset + dx) * x1) * y0 +
		            (__ldg(d + offset + dy) * x0 + __ldg(d + offset + dy + dx) * x1) * y1) * z0 +
		           ((__ldg(d + offset + dz) * x0 + __ldg(d + offset + dz + dx) * x1) * y0 +
		            (__ldg(d + offset + dz + dy) * x0 + __ldg(d + offset + dz + dy + dx) * x1) * y1) * z1;
		offset++;
	}
}
// rgb2spec_eval_precise (rgb2spec.c:129-133) at the hero wavelengths, no FMA (parity build)
__device__ __forceinline__ Hero jh_eval(const KParams& P, float c0, float c1, float c2, float lambda_0) {
	Hero h;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		float lambda = lambda_0 + (float)k * P.lambda_step;
		float xx = (c0 * lambda + c1) * lambda + c2;
		float yy = 1.f / sqrtf(xx * xx + 1.f);
		h.v[k] = (uint32_t)k < P.n_wavelengths ? (.5f * xx) * yy + .5f : 0.0f;
	}
	return h;
}
// Color::lrgb_to_specrefl, JH build (color.cpp:202-232): both steps per lookup, as the reference does
#ifndef SSB_INLINE_UPS
#define SSB_INLINE_UPS 1  // 1: the Jakob-Hanika / Meng upsampling functions inlined into their single call site (+1 % on those configurations)
#endif
#if SSB_INLINE_UPS
#define SSB_ATTR_UPS __device__ __forceinline__
#else
#define SSB_ATTR_UPS __device__ __noinline__
#endif
SSB_ATTR_UPS Hero jh_upsample(const KParams& P, float r, float g, float b, float lambda_0) {
	float coeff[3];
	jh_fetch(P.jh_scale, P.jh_data, (int)P.jh_res, r, g, b, coeff);
	return jh_eval(P, coeff[0], coeff[1], coeff[2], lambda_0);
}
// the second step alone, on coefficients fetched once per texel by ssb_bake_jh_kernel (ssb_options.prebaked_textures)
SSB_ATTR_UPS Hero jh_eval_baked(const KParams& P, float4 cf, float lambda_0) {
	return jh_eval(P, cf.x, cf.y, cf.z, lambda_0);
}
// spectrum_xyz_to_p (meng-et-al.-2015/spectrum_grid.h:13-137) for the hero wavelengths of one colour.  The reference calls
// it once per wavelength (color.cpp:196-199) and every call repeats the part that does not depend on the wavelength: the
// chromaticity, the grid cell, and — for a border cell — the walk around the triangle fan that ends with the barycentric
// weights (u, v, w) of the triangle holding the point.  Here that part runs ONCE per colour and only the spectral
// interpolation of the cell's corner points runs per wavelength: the same operations on the same operands, so every
// returned value is the per-call one bit for bit (as with the shared grid arithmetic of spec_sample3).
SSB_ATTR_UPS Hero meng_upsample(const KParams& P, float r, float g, float b, float lambda_0) {
	// color.cpp:189-200: xyz_rel = transpose(M) * 100 * lrgb
	const float M[9] = { 0.41231515f, 0.3576f, 0.1805f, 0.2126f, 0.7152f, 0.0722f, 0.01932727f, 0.1192f, 0.95063333f };
	float xyz[3];
	for (int rr = 0; rr < 3; ++rr)
		xyz[rr] = ((M[rr * 3 + 0] * 100.0f) * r + (M[rr * 3 + 1] * 100.0f) * g) + (M[rr * 3 + 2] * 100.0f) * b;
	Hero h;
	h.v[0] = h.v[1] = h.v[2] = h.v[3] = 0.0f;
	float xyY[3], uv[2];
	const float norm = (float)(1.0 / (double)((xyz[0] + xyz[1]) + xyz[2]));
	if (!(norm < 3.402823466e+38f)) return h;
	xyY[0] = xyz[0] * norm; xyY[1] = xyz[1] * norm; xyY[2] = xyz[1];
	uv[0] = P.meng_xy_to_uv[0] * xyY[0] + P.meng_xy_to_uv[1] * xyY[1] + P.meng_xy_to_uv[2];
	uv[1] = P.meng_xy_to_uv[3] * xyY[0] + P.meng_xy_to_uv[4] * xyY[1] + P.meng_xy_to_uv[5];
	if (uv[0] < 0.0f || uv[0] >= (float)P.meng_grid_w || uv[1] < 0.0f || uv[1] >= (float)P.meng_grid_h) return h;
	const int uvi[2] = { (int)uv[0], (int)uv[1] };
	const int cell_idx = uvi[0] + (int)P.meng_grid_w * uvi[1];
	const int32_t* __restrict__ cell = P.meng_grid + 8 * cell_idx;
	const int4 c0 = __ldg(reinterpret_cast<const int4*>(cell)), c1 = __ldg(reinterpret_cast<const int4*>(cell) + 1);
	const int inside = c0.x, num = c0.y;
	const int idx0 = c0.z, idx1 = c0.w;
	// (corner indices beyond the first two are needed at run-time positions: re-read from the cell — an L1 hit — rather than
	// from a dynamically indexed local array, which would live in local memory)
	auto idx_at = [&](int k) { return k == 0 ? idx0 : (k == 1 ? idx1 : __ldg(cell + 2 + k)); };
	(void)c1;
	const uint32_t stride = 5 + P.meng_nsamples;
	const float* __restrict__ pts = P.meng_points;
	const int ns = (int)P.meng_nsamples;
	// ---- the wavelength-independent part: which corner points are blended, with which weights
	//   inside: the bilinear weights of the four corners (layout 2 3 / 0 1);  border: p[0] * w + p[ia] * v + p[ib] * u
	int ia = -1, ib = -1;
	float wu = 0.0f, wv = 0.0f, ww = 0.0f, fu = 0.0f, fv = 0.0f;
	if (inside) {
		fu = uv[0] - (float)uvi[0]; fv = uv[1] - (float)uvi[1];
	} else {
#define SSB_MENG_UV(k, c) __ldg(pts + (size_t)stride * idx_at(k) + 3 + (c))
		const float p0u = SSB_MENG_UV(0, 0), p0v = SSB_MENG_UV(0, 1);
		const float p1u = SSB_MENG_UV(1, 0), p1v = SSB_MENG_UV(1, 1);
		const float ex = uv[0] - p0u, ey = uv[1] - p0v;
		float e0x = p1u - p0u, e0y = p1v - p0v;
		float uu = e0x * ey - ex * e0y;
		for (int i = 0; i < num - 1; i++) {
			float e1x, e1y;
			if (i == num - 2) { e1x = p1u - p0u; e1y = p1v - p0v; }
			else { e1x = SSB_MENG_UV(i + 2, 0) - p0u; e1y = SSB_MENG_UV(i + 2, 1) - p0v; }
			const float vv = ex * e1y - e1x * ey;
			const float area = e0x * e1y - e1x * e0y;
			const float u = uu / area, v = vv / area;
			const float w = 1.0f - u - v;
			if (u < 0.0f || v < 0.0f || w < 0.0f) { uu = -vv; e0x = e1x; e0y = e1y; continue; }
			ia = i + 1; ib = (i == num - 2) ? 1 : (i + 2);
			wu = u; wv = v; ww = w;
			break;
		}
#undef SSB_MENG_UV
	}
	// ---- per wavelength: the corner points' spectra at lambda (linear in the table's sample grid), blended
	for (int k = 0; k < 4; ++k) {  // (wavelengths past the last channel may lie outside the tables: not evaluated)
		if ((uint32_t)k >= P.n_wavelengths) break;
		const float lambda = lambda_0 + (float)k * P.lambda_step;
		const float sb = (lambda - P.meng_sample_min) / (P.meng_sample_max - P.meng_sample_min) * (float)(ns - 1);
		const int sb0 = (int)sb;
		const int sb1 = sb + 1 < (float)ns ? (int)(sb + 1) : ns - 1;
		const float sbf = sb - (float)sb0;
		auto corner = [&](int i) {
			const float* spectrum = pts + (size_t)stride * idx_at(i) + 5;
			return __ldg(spectrum + sb0) * (1.0f - sbf) + __ldg(spectrum + sb1) * sbf;
		};
		float interpolated_p = 0.0f;
		if (inside) {
			const float p0 = corner(0), p1 = corner(1), p2 = corner(2), p3 = corner(3);
			interpolated_p = p0 * (1.0f - fu) * (1.0f - fv) + p2 * (1.0f - fu) * fv + p3 * fu * fv + p1 * fu * (1.0f - fv);
		} else if (ia >= 0) {
			const float p0 = corner(0), pa = corner(ia), pb = corner(ib);
			interpolated_p = p0 * ww + pa * wv + pb * wu;
		}
		h.v[k] = interpolated_p / norm;
	}
	return h;
}

// Template parameter UPS of the shading code: one of SSB_UPSAMPLE_* for spectral transport, or SSB_UPS_RGB for the
// reference's RENDER_MODE_RGB build (three l-RGB channels carried in v[0..2], v[3] = 0: every per-channel operation of
// the reference's vec3 arithmetic is the same expression, and dot(f,f) = (f0^2+f1^2)+(f2^2+0) is unchanged by the zero).
#define SSB_UPS_RGB 0

// MaterialBase::evaluate_emission (material.hpp:96-104)
template <int UPS>
__device__ __forceinline__ Hero material_emission(const KParams& P, const SceneView& S, const DevMaterial& m, float lambda_0) {
	if (UPS == SSB_UPS_RGB) {
		Hero h; h.v[0] = m.emission_rgb[0]; h.v[1] = m.emission_rgb[1]; h.v[2] = m.emission_rgb[2]; h.v[3] = 0.0f;
		return h;
	}
	Hero h = spec_hero(P, m.emission, lambda_0, P.lambda_step);
	hero_clear_unused(P, h);
	return h;
}

// material albedo at (st, lambda_0): constant spectrum or sRGB texture + upsampling
// (material.cpp:45-97,120-143; color.cpp:166-232)
template <int UPS>
__device__ __forceinline__ Hero material_albedo(const KParams& P, const SceneView& S, const DevMaterial& m,
                                                float st_x, float st_y, float lambda_0) {
	if (m.albedo_mode == SSB_ALBEDO_CONSTANT) {
		if (UPS == SSB_UPS_RGB) {
			Hero h; h.v[0] = m.albedo_rgb[0]; h.v[1] = m.albedo_rgb[1]; h.v[2] = m.albedo_rgb[2]; h.v[3] = 0.0f;
			return h;
		}
		Hero h = spec_hero(P, m.albedo, lambda_0, P.lambda_step);
		hero_clear_unused(P, h);
		return h;
	}
	const DevTexture tex = S.textures()[m.texture];
	float index_x = st_x * (float)tex.width;
	float index_y = (float)tex.height - st_y * (float)tex.height;
	int i = (int)floorf(index_x), j = (int)floorf(index_y);
	i = max(i, 0); i = min(i, (int)tex.width - 1);
	j = max(j, 0); j = min(j, (int)tex.height - 1);
	const size_t texel = (size_t)j * tex.width + (size_t)i;
	if (UPS == SSB_UPSAMPLE_JH && P.jh_prebaked)  // uniform: the texel's coefficients were fetched once, at bake time
		return jh_eval_baked(P, __ldg(tex.coef + texel), lambda_0);
	uchar4 px = __ldg(tex.rgba + texel);
	float r = S.hdr()->srgb_lut[px.x], g = S.hdr()->srgb_lut[px.y], b = S.hdr()->srgb_lut[px.z];
	if (UPS == SSB_UPS_RGB) {  // material.cpp:64-66: the texel's l-RGB is the reflectance
		Hero h; h.v[0] = r; h.v[1] = g; h.v[2] = b; h.v[3] = 0.0f;
		return h;
	} else if (UPS == SSB_UPSAMPLE_OURS) {
		Hero br = spec_hero(P, S.hdr()->basis_r, lambda_0, P.lambda_step);
		Hero bg = spec_hero(P, S.hdr()->basis_g, lambda_0, P.lambda_step);
		Hero bb = spec_hero(P, S.hdr()->basis_b, lambda_0, P.lambda_step);
		Hero h;
#pragma unroll
		for (int k = 0; k < 4; ++k) h.v[k] = (r * br.v[k] + g * bg.v[k]) + b * bb.v[k];
		hero_clear_unused(P, h);
		return h;
	} else if (UPS == SSB_UPSAMPLE_JH) {
		return jh_upsample(P, r, g, b, lambda_0);
	} else {
		return meng_upsample(P, r, g, b, lambda_0);
	}
}

// ------------------------------------------------------------------ light sampling
// Math::SphericalTriangle (util/spherical-tri.cpp:18-124) + Math::rand_toward_sphericaltri
// (util/random.cpp:101-154, Arvo 1995) + PrimTri::get_rand_toward (geometry.cpp:103-116), fused.
__device__ __forceinline__ float underestimate_pi() { return __uint_as_float(0x40490FDAu); }

__device__ __forceinline__ void func_bar(float xx, float xy, float xz, float yx, float yy, float yz,
                                         float& ox, float& oy, float& oz) {  // random.cpp:139-144
	float d = dot3(xx, xy, xz, yx, yy, yz);
	float dx = xx - d * yx, dy = xy - d * yy, dz = xz - d * yz;
	float lensq = dot3(dx, dy, dz, dx, dy, dz);
	if (lensq == 0.0f) { ox = oy = oz = 0.0f; return; }
	float inv = 1.0f / sqrtf(lensq);
	ox = dx * inv; oy = dy * inv; oz = dz * inv;
}

__device__ __noinline__ float cos_double_cold(float x) { return (float)cos((double)x); }  // degenerate triangles only

// Returns (direction, pdf) BY VALUE: reference parameters of a non-inlined function live in local memory.
#ifndef SSB_INLINE_SPHTRI
#define SSB_INLINE_SPHTRI 1  // (one call site; +0.3 %)
#endif
#if SSB_INLINE_SPHTRI
__device__ __forceinline__
#else
__device__ __noinline__
#endif
float4 sample_spherical_triangle(int light_quad, int light_tri, float px, float py, float pz, float r0, float r1) {
	float wx, wy, wz, pdf;
	const ssb_tri& t = SceneView().quads()[light_quad].tri[light_tri];
	// unit vectors toward the vertices: glm::normalize(v - from) = v * (1/sqrt(dot))
	float Ax = t.v[0].pos[0] - px, Ay = t.v[0].pos[1] - py, Az = t.v[0].pos[2] - pz;
	float Bx = t.v[1].pos[0] - px, By = t.v[1].pos[1] - py, Bz = t.v[1].pos[2] - pz;
	float Cx = t.v[2].pos[0] - px, Cy = t.v[2].pos[1] - py, Cz = t.v[2].pos[2] - pz;
	float ia = 1.0f / sqrtf(dot3(Ax, Ay, Az, Ax, Ay, Az)); Ax *= ia; Ay *= ia; Az *= ia;
	float ib = 1.0f / sqrtf(dot3(Bx, By, Bz, Bx, By, Bz)); Bx *= ib; By *= ib; Bz *= ib;
	float ic = 1.0f / sqrtf(dot3(Cx, Cy, Cz, Cx, Cy, Cz)); Cx *= ic; Cy *= ic; Cz *= ic;

	float cos_a = glm_clamp(dot3(Bx, By, Bz, Cx, Cy, Cz), -1.0f, 1.0f);
	float cos_b = glm_clamp(dot3(Ax, Ay, Az, Cx, Cy, Cz), -1.0f, 1.0f);
	float cos_c = glm_clamp(dot3(Ax, Ay, Az, Bx, By, Bz), -1.0f, 1.0f);
	const float2 acos_ab = acosf2_x(cos_a, cos_b);
	float a = glm_clamp(acos_ab.x, 0.0f, underestimate_pi());
	float b = glm_clamp(acos_ab.y, 0.0f, underestimate_pi());
	float c = glm_clamp(acosf_x(cos_c), 0.0f, underestimate_pi());
	float sin_a = sinf_x(a), sin_b = sinf_x(b), sin_c = sinf_x(c);
	float numer0 = cos_a - cos_b * cos_c;
	float numer1 = cos_b - cos_c * cos_a;
	float numer2 = cos_c - cos_a * cos_b;
	float denom0 = sin_b * sin_c, denom1 = sin_c * sin_a, denom2 = sin_a * sin_b;
	float alpha, cos_alpha, surface_area;
	const float nan = __int_as_float(0x7fc00000);
	if (denom0 > 0 && denom1 > 0 && denom2 > 0) {
		cos_alpha = glm_clamp(numer0 / denom0, -1.0f, 1.0f);
		float cos_beta = glm_clamp(numer1 / denom1, -1.0f, 1.0f);
		float cos_gamma = glm_clamp(numer2 / denom2, -1.0f, 1.0f);
		const float2 acos_albe = acosf2_x(cos_alpha, cos_beta);
		alpha = glm_clamp(acos_albe.x, 0.0f, underestimate_pi());
		float beta = glm_clamp(acos_albe.y, 0.0f, underestimate_pi());
		float gamma = glm_clamp(acosf_x(cos_gamma), 0.0f, underestimate_pi());
		surface_area = ((alpha + beta) + gamma) - SSB_PI_F;
		if (!(surface_area >= 0)) surface_area = 0;
	} else {
		// degenerate branches (spherical-tri.cpp:74-122): only alpha / cos_alpha are consumed downstream
		surface_area = 0;
		if (sin_a > 0) {
			if (sin_b > 0) {
				if (sin_c > 0) { alpha = cos_alpha = nan; }
				else { cos_alpha = 1; alpha = SSB_PI_F * 0.5f; }
			} else {
				if (sin_c > 0) { cos_alpha = 1; alpha = SSB_PI_F * 0.5f; }
				else { alpha = cos_alpha = nan; }
			}
		} else {
			if (sin_b > 0 && sin_c > 0) { cos_alpha = glm_clamp(numer0 / denom0, -1.0f, 1.0f); alpha = acosf_x(cos_alpha); }
			else { alpha = cos_alpha = nan; }
		}
	}
	// geometry.cpp:115: pdf = 1 / area, +inf when the area is +0 (every point coplanar with the light: the whole ceiling of
	// the Cornell box) — written out so that those lanes stay off the reciprocal's slow path
	{
		const bool flat = surface_area == 0.0f;
		float sa = flat ? 1.0f : surface_area;
		asm("" : "+f"(sa));
		const float r = 1.0f / sa;
		pdf = flat ? __int_as_float(0x7f800000) : r;
	}

	// Arvo sampling (random.cpp:101-154)
	float sin_alpha = sinf_x(alpha);
	float q;
	if (sin_alpha > 0) {
		float random_area = r0 * surface_area;
		float phi = random_area - alpha;
		const float2 sc_phi = sincosf_x(phi);
		const float s = sc_phi.x, tt = sc_phi.y;
		float u = tt - cos_alpha;
		float v = s + sin_alpha * cos_c;
		float denom = (v * s + u * tt) * sin_alpha;
		if (denom != 0.0f) q = ((v * tt - u * s) * cos_alpha - v) / denom;
		else q = cos_c;
	} else {
		q = cos_double_cold(b * r0);  // ::cos(double), random.cpp:135
	}
	q = glm_clamp(q, -1.0f, 1.0f);
	float fx, fy, fz;
	func_bar(Cx, Cy, Cz, Ax, Ay, Az, fx, fy, fz);
	float sq = sqrtf(1.0f - q * q);
	float Chx = q * Ax + sq * fx, Chy = q * Ay + sq * fy, Chz = q * Az + sq * fz;
	float z = 1.0f - r1 * (1.0f - dot3(Chx, Chy, Chz, Bx, By, Bz));
	z = glm_clamp(z, -1.0f, 1.0f);
	func_bar(Chx, Chy, Chz, Bx, By, Bz, fx, fy, fz);
	float sz = sqrtf(1.0f - z * z);
	wx = z * Bx + sz * fx; wy = z * By + sz * fy; wz = z * Bz + sz * fz;
	return make_float4(wx, wy, wz, pdf);
}

// ------------------------------------------------------------------ TMA bulk copy of the blob into shared memory
__device__ __forceinline__ void stage_blob(unsigned char* smem, const unsigned char* gmem, uint32_t bytes,
                                           unsigned long long* bar) {
	const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(bar);
	const uint32_t dst_addr = (uint32_t)__cvta_generic_to_shared(smem);
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
		asm volatile(
			"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_addr),
			"l"(gmem), "r"(bytes), "r"(bar_addr)
			: "memory");
	}
	// all threads wait for phase 0 to complete
	uint32_t done = 0;
	while (!done) {
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(done)
			: "r"(bar_addr)
			: "memory");
	}
}

// ---- per-thread software pipeline of path records (cp.async / LDGSTS, 16 bytes each, L1 bypassed): the records of the
// NEXT grid-stride iteration are copied into the thread's own shared-memory slot while the current iteration computes.
// ncu on the unpipelined kernels: ~20 % (intersect) and ~25 % (shade) of all warp time was `long_scoreboard` at the
// record loads at the top of each iteration (two dependent round trips in the shade stage: order[] -> records).  With the
// pipeline that stall is gone (2.0 -> 0.5 cycles per issue in the intersect stage); the frame gains only ~0.5 %, the
// stages being bound by dependent-issue latency and the phase barriers rather than by that wait (profiles/README.md).
// Each thread reads back only what its own copies wrote, so cp.async.wait_group is the only synchronisation needed.
#ifndef SSB_PIPELINE
#define SSB_PIPELINE 1
#endif
#ifndef SSB_PIPELINE_ISECT
#define SSB_PIPELINE_ISECT SSB_PIPELINE
#endif
#ifndef SSB_PIPELINE_SHADE
#define SSB_PIPELINE_SHADE SSB_PIPELINE
#endif
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- whole-sector accesses (sm_100: 256-bit LDG/STG).  A path record is one 32-byte sector; written as two 16-byte
// stores it reaches the L2 as two PARTIAL sector writes, and the L2 completes a partially written sector with a DRAM read
// (ncu, r5c: the intersect stage read 50 B per ray for 32 B of ray record, the shade stage 212 B per vertex for ~90 B).
// One 32-byte store per record carries the full byte mask: no fill.
__device__ __forceinline__ void st_sector(float4* p, const float4 a, const float4 b) {
	asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}
__device__ __forceinline__ void ld_sector(const float4* p, float4& a, float4& b) {
	asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p) : "memory");
}

// launch bounds, measured (profiles/r5e_ab.txt, r5p_ab.txt): intersect 1024 x 1 beats 512 x 2 by 1 % and 256 x 4 by 2 % (the same
// 32 warps per SM in fewer CTAs: fewer blob stagings and histogram flushes per launch); shade 256 x 3 beats 384 x 2 by 0.4 %,
// 192 x 4 and 128 x 6 by 4 %
#ifndef SSB_INTERSECT_THREADS
#define SSB_INTERSECT_THREADS 1024
#endif
#ifndef SSB_INTERSECT_MIN_BLOCKS
#define SSB_INTERSECT_MIN_BLOCKS 1
#endif
#ifndef SSB_SHADE_THREADS
#define SSB_SHADE_THREADS 256
#endif
#ifndef SSB_SHADE_MIN_BLOCKS
#define SSB_SHADE_MIN_BLOCKS 3
#endif

// image row of local row `lr` of the pass rectangle: consecutive rows, or the lr-th row of this GPU's interleaved bands
__device__ __forceinline__ uint32_t image_row(const KParams& P, uint32_t lr) {
	if (P.band_n <= 1u) return P.y0 + lr;
	return ((lr / P.band_h) * P.band_n + P.band_i) * P.band_h + lr % P.band_h;
}
// Scene::intersect as the kernels run it: the filtered scan, or (ssb_options.scan_mode, uniform) the reference's own loop
// Scene::intersect as the kernels run it: the filtered scan, or — template parameter LIST, chosen by ssb_options.scan_mode —
// the reference's own loop.  (A run-time branch on a kernel parameter cost 1 % of the frame even when never taken: the
// second callee's registers and code; profiles/r5g_ab.txt.  Hence separate instantiations.)
template <bool LIST>
__device__ __forceinline__ void scene_query(const SceneView& S, float eps, int ignore, Hit& hit,
                                            float ox, float oy, float oz, float dx, float dy, float dz) {
	if (LIST) scene_intersect_listscan_noinline(S, eps, ignore, hit, ox, oy, oz, dx, dy, dz);
	else scene_intersect(S, eps, ignore, hit, ox, oy, oz, dx, dy, dz);
}

// the record pipelines live in dynamic shared memory right behind the blob (host side: ssb_pipe_offset / launch sizes in ssb_render)
__device__ __forceinline__ uint32_t pipe_offset(const SceneView& S) { return (S.hdr()->total_bytes + 127u) & ~127u; }
#define SSB_INTERSECT_PIPE_BYTES (2u * 2u * SSB_INTERSECT_THREADS * 16u)
#define SSB_SHADE_PIPE_BYTES (2u * 4u * SSB_SHADE_THREADS * 16u)

__device__ __forceinline__ void stage_scene(const KParams& P, unsigned long long* bar) {
	const DevHeader* gh = reinterpret_cast<const DevHeader*>(P.blob);
	stage_blob(ssb_smem, P.blob, __ldg(&gh->total_bytes), bar);
}

// ---- stage 1 of a bounce: the closest-hit query of every live path at depth P.depth (renderer.cpp:163).
// FIRST: depth 0 — the path is created here (Renderer::_render_sample prologue, renderer.cpp:103-138) and its state
// written.  A miss ends the path (L() returns 0).  Hits are recorded and counted per hit quad, so that the shading
// stage can run with all lanes of a warp on the same quad / material.
template <bool FIRST, bool LIST>
__global__ void __launch_bounds__(SSB_INTERSECT_THREADS, SSB_INTERSECT_MIN_BLOCKS)
ssb_intersect_kernel(const __grid_constant__ KParams P) {
	__shared__ __align__(8) unsigned long long blob_bar;
	stage_scene(P, &blob_bar);
	const SceneView S;
	const unsigned full = 0xffffffffu;
	const uint32_t npix_rect = P.rect_w * P.rect_h;
	const int depth = (int)P.depth;
	const uint32_t n_in = FIRST ? (uint32_t)P.total_work : P.counts[depth];
	const int pin = depth & 1;
	const uint32_t nthreads = gridDim.x * blockDim.x;
	const uint32_t n_round = (n_in + 31u) & ~31u;  // warps iterate together (match/ballot below)
	uint32_t* bins = P.bin_count + (size_t)depth * SSB_MAX_QUADS;
	// hits per quad are counted in shared memory and flushed once per CTA: the global counters are only a handful of
	// addresses (ncu on the first split build: the scatter kernel spent 0.77 ms waiting on 19 contended atomics)
	__shared__ uint32_t s_bins[SSB_MAX_QUADS];
	for (uint32_t q = threadIdx.x; q < SSB_MAX_QUADS; q += blockDim.x) s_bins[q] = 0;
	__syncthreads();
#if SSB_PIPELINE_ISECT
	// [stage][record: recA.0, recA.1][thread]  (the sample id, needed only by the few paths that miss, is fetched on demand:
	// prefetching recR for every path moved 32 B per query for nothing)
	// (in DYNAMIC shared memory behind the blob: launch bounds above 384 threads would exceed the 48 KB of static shared memory)
	float4 (*s_pipe)[2][SSB_INTERSECT_THREADS] = reinterpret_cast<float4 (*)[2][SSB_INTERSECT_THREADS]>(ssb_smem + pipe_offset(S));
	auto prefetch = [&](uint32_t it, int stage) {
		if (!FIRST && it < n_in) {
			cp_async16(&s_pipe[stage][0][threadIdx.x], &P.recA[pin][2 * (size_t)it]);
			cp_async16(&s_pipe[stage][1][threadIdx.x], &P.recA[pin][2 * (size_t)it + 1]);
		}
		cp_async_commit();
	};
	if (!FIRST) prefetch(blockIdx.x * blockDim.x + threadIdx.x, 0);
	int stage = 0;
#endif

	for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < n_round; item += nthreads) {
		const bool valid = item < n_in;
		uint32_t hq = 0xffffffffu;
#if SSB_PIPELINE_ISECT
		const int cur = stage;  // the slot of this iteration stays intact until the next iteration's prefetch
		if (!FIRST) {
			prefetch(item + nthreads, stage ^ 1);
			cp_async_wait<1>();
			stage ^= 1;
		}
#endif
		if (valid) {
			float ox, oy, oz, dx, dy, dz;
			int ignore = -1;
			if (FIRST) {
				const uint32_t id = item;
				uint32_t kk = id / npix_rect;
				uint32_t pr = id - kk * npix_rect;
				uint32_t pi = P.x0 + pr % P.rect_w, pj = image_row(P, pr / P.rect_w);
				uint32_t k = P.sample_begin + kk;
				unsigned long long sample_index = (unsigned long long)k * ((unsigned long long)P.width * P.height) +
				                                  ((unsigned long long)pj * P.width + pi);
				Rng rng;
				rng.state = mix64(P.seed ^ mix64(sample_index));
				rng.inc = mix64(rng.state) | 1ull;
				double sub_y = rand_1d(rng);  // g++ evaluates dvec2(rand_1d(),rand_1d()) right-to-left
				double sub_x = rand_1d(rng);
				double st_x = ((double)pi + sub_x) / (double)P.width;
				double st_y = ((double)pj + sub_y) / (double)P.height;
				double ndc_x = st_x * 2.0 - 1.0, ndc_y = st_y * 2.0 - 1.0;
				double pt[4];
#pragma unroll
				for (int r = 0; r < 4; ++r)
					pt[r] = (P.pv_inv[0 + r] * ndc_x + P.pv_inv[4 + r] * ndc_y) + (P.pv_inv[8 + r] * 0.0 + P.pv_inv[12 + r] * 1.0);
				double w = pt[3];
				double ddx = pt[0] / w - (double)P.cam_pos[0], ddy = pt[1] / w - (double)P.cam_pos[1], ddz = pt[2] / w - (double)P.cam_pos[2];
				double inv = 1.0 / sqrt((ddx * ddx + ddy * ddy) + ddz * ddz);
				ox = P.cam_pos[0]; oy = P.cam_pos[1]; oz = P.cam_pos[2];
				dx = (float)(ddx * inv); dy = (float)(ddy * inv); dz = (float)(ddz * inv);
				// hero wavelength (renderer.cpp:134-143): drawn in spectral mode only
				const float lambda_0 = (P.render_mode == SSB_RENDER_RGB) ? 0.0f : P.lambda_min + rand_1f(rng) * P.lambda_step;
				if (!P.flat_field) P.ff[id] = dot3(dx, dy, dz, P.cam_dir[0], P.cam_dir[1], P.cam_dir[2]);
				// (the camera ray is consumed right here; only MaterialMirror::interact_bsdf looks at it again — 32 B per sample saved otherwise)
				if (P.has_mirror) st_sector(&P.recA[0][2 * (size_t)item], make_float4(ox, oy, oz, __int_as_float(-1)), make_float4(dx, dy, dz, lambda_0));
				st_sector(&P.recR[0][2 * (size_t)item],
				          make_float4(__uint_as_float((uint32_t)rng.state), __uint_as_float((uint32_t)(rng.state >> 32)),
				                      __uint_as_float((uint32_t)rng.inc), __uint_as_float((uint32_t)(rng.inc >> 32))),
				          make_float4(__uint_as_float(id), lambda_0, 0.f, 0.f));
			} else {
#if SSB_PIPELINE_ISECT
				const float4 a = s_pipe[cur][0][threadIdx.x], b = s_pipe[cur][1][threadIdx.x];
#else
				const float4 a = P.recA[pin][2 * (size_t)item], b = P.recA[pin][2 * (size_t)item + 1];
#endif
				ox = a.x; oy = a.y; oz = a.z; ignore = __float_as_int(a.w);
				dx = b.x; dy = b.y; dz = b.z;
			}
			Hit hit;
			scene_query<LIST>(S, P.eps, ignore, hit, ox, oy, oz, dx, dy, dz);
			if (hit.quad >= 0) {
				hq = (uint32_t)hit.quad | ((uint32_t)hit.tri << 31);
				// hit position (Ray::at, stdafx.hpp:219): origin of the shadow ray and of the next path ray
				const float d_ = hit.dist;
				st_sector(&P.recH[2 * (size_t)item], make_float4(ox + d_ * dx, oy + d_ * dy, oz + d_ * dz, __uint_as_float(hq)),
				          make_float4(hit.bx, hit.by, hit.bz, 0.f));
			} else {
				// miss: L() returns 0 (renderer.cpp:161-163 with no hit); hit_anything only if an earlier depth hit
				const float4 r1 = P.recR[pin][2 * (size_t)item + 1];
				const uint32_t id = __float_as_uint(r1.x);
				st_sector(&P.leaf[2 * (size_t)id], make_float4(0.f, 0.f, 0.f, 0.f), make_float4(r1.y, __int_as_float(depth | (FIRST ? 0 : (1 << 16))), 0.f, 0.f));
			}
			P.hit_q[item] = hq;
		}
		// count hits per quad: one atomic per distinct quad per warp
		const bool is_hit = hq != 0xffffffffu;
		const unsigned hmask = __ballot_sync(full, is_hit);
		if (is_hit) {
			const uint32_t q = hq & 0x7fffffffu;
			const unsigned peers = __match_any_sync(hmask, q);
			if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_bins[q], (uint32_t)__popc(peers));
		}
	}
	__syncthreads();
	for (uint32_t q = threadIdx.x; q < SSB_MAX_QUADS; q += blockDim.x)
		if (s_bins[q]) atomicAdd(&bins[q], s_bins[q]);
}

// ---- counting-sort scatter: order[] = queue positions of the hit paths, grouped by hit quad.  Every CTA first turns the
// per-quad hit counts of the depth (complete: the intersect launch has finished) into bin offsets with a block-wide
// exclusive scan (<= SSB_MAX_QUADS = blockDim counters: cheaper than a separate one-thread launch per depth, which cost
// a launch gap + ~5 us nine times per frame), then takes tiles of 4 x blockDim queue entries: it counts the tile's hits
// per quad in shared memory, reserves one contiguous range per quad with a single global atomic on the quad's cursor
// (zeroed with the other counters at the start of the pass), and places the entries through shared-memory cursors.
#ifndef SSB_SCATTER_PER_THREAD
#define SSB_SCATTER_PER_THREAD 4
#endif
__global__ void __launch_bounds__(SSB_MAX_QUADS) ssb_bin_scatter_kernel(const __grid_constant__ KParams P, uint32_t first_depth, uint32_t nquads) {
	__shared__ uint32_t s_cnt[SSB_MAX_QUADS];
	__shared__ uint32_t s_base[SSB_MAX_QUADS];
	__shared__ uint32_t s_off[SSB_MAX_QUADS];
	__shared__ uint32_t s_warp[SSB_MAX_QUADS / 32];
	const int depth = (int)P.depth;
	const uint32_t n_in = first_depth ? (uint32_t)P.total_work : P.counts[depth];
	uint32_t* cur = P.bin_cursor + (size_t)depth * SSB_MAX_QUADS;
	{
		const uint32_t* cnt = P.bin_count + (size_t)depth * SSB_MAX_QUADS;
		const uint32_t q = threadIdx.x, lane = q & 31u;
		const uint32_t v = q < nquads ? cnt[q] : 0u;
		uint32_t incl = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
			if (lane >= (uint32_t)d) incl += t;
		}
		if (lane == 31u) s_warp[q >> 5] = incl;
		__syncthreads();
		uint32_t before = 0;
		for (uint32_t w = 0; w < (q >> 5); ++w) before += s_warp[w];
		s_off[q] = before + incl - v;
		if (blockIdx.x == 0 && q == blockDim.x - 1) P.nhits[depth] = before + incl;  // total hits = length of `order`
		__syncthreads();
	}
	const uint32_t tile = blockDim.x * SSB_SCATTER_PER_THREAD;
	for (uint32_t t0 = blockIdx.x * tile; t0 < n_in; t0 += gridDim.x * tile) {
		for (uint32_t q = threadIdx.x; q < nquads; q += blockDim.x) s_cnt[q] = 0;
		__syncthreads();
		uint32_t hq[SSB_SCATTER_PER_THREAD];
#pragma unroll
		for (int k = 0; k < SSB_SCATTER_PER_THREAD; ++k) {
			const uint32_t item = t0 + k * blockDim.x + threadIdx.x;
			hq[k] = item < n_in ? P.hit_q[item] : 0xffffffffu;
			if (hq[k] != 0xffffffffu) atomicAdd(&s_cnt[hq[k] & 0x7fffffffu], 1u);
		}
		__syncthreads();
		for (uint32_t q = threadIdx.x; q < nquads; q += blockDim.x) {
			const uint32_t c = s_cnt[q];
			s_base[q] = c ? s_off[q] + atomicAdd(&cur[q], c) : 0u;
			s_cnt[q] = 0;
		}
		__syncthreads();
#pragma unroll
		for (int k = 0; k < SSB_SCATTER_PER_THREAD; ++k) {
			if (hq[k] != 0xffffffffu) {
				const uint32_t q = hq[k] & 0x7fffffffu;
				P.order[s_base[q] + atomicAdd(&s_cnt[q], 1u)] = t0 + k * blockDim.x + threadIdx.x;
			}
		}
		__syncthreads();
	}
}

// ---- stage 2 of a bounce: everything the reference's lambda L does after the closest hit (renderer.cpp:165-251)
// for the paths that hit something, visited grouped by hit quad: emission, albedo, light sample + shadow query,
// BSDF sample, fold record, and the compacted state of the continuing paths.
// The body is written as phases separated by CTA barriers (SSB_SHADE_SYNC): the kernel is ~45 KB of mostly
// straight-line code, and ncu showed `no_instruction` (instruction fetch) as its top stall with every warp streaming
// the code on its own; the barriers keep the 8 warps of a CTA inside the same code region so that they share fetched
// lines.  The barriers are outside all data-dependent control flow (the loop trip count is uniform per CTA).
#ifndef SSB_SHADE_SYNC
#define SSB_SHADE_SYNC 1
#endif
#if SSB_SHADE_SYNC
#define SSB_PHASE_BARRIER() __syncthreads()
#else
#define SSB_PHASE_BARRIER() ((void)0)
#endif
// Which of the four barriers of an iteration are kept (bit k = the barrier after phase k).  Measured (profiles/
// r3c_ab_pipeline_unroll_barriers.txt): all four 896.8 Msamples/s; only the one before the long light-sampling phase and the
// one at the end of the iteration (mask 9) 910.7; without the barrier after the shadow query (mask 11) 907; only the
// end-of-iteration one (8) 905; only the one after light sampling (2) 869; none 841.  The barrier after the shadow query
// was where warps waited longest (ncu: 10 % of all warp time) — that phase's length varies with the exact tests a warp runs.
// Round 2, after the record traffic was cut (whole-sector records): the end-of-iteration barrier alone (8) is now the best,
// 990 vs 986 with mask 9 and 913 with none (profiles/r5d_ab.txt).
#ifndef SSB_SHADE_SYNC_MASK
#define SSB_SHADE_SYNC_MASK 8
#endif
#define SSB_PHASE_BARRIER_AT(k) do { if ((SSB_SHADE_SYNC_MASK >> (k)) & 1) SSB_PHASE_BARRIER(); } while (0)
template <bool FIRST, int UPS, bool LIST>
__global__ void __launch_bounds__(SSB_SHADE_THREADS, SSB_SHADE_MIN_BLOCKS)
ssb_shade_kernel(const __grid_constant__ KParams P) {
	__shared__ __align__(8) unsigned long long blob_bar;
	stage_scene(P, &blob_bar);
	const SceneView S;

	const unsigned full = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const float eps = P.eps;
	const bool els = P.els != 0;
	const int depth = (int)P.depth;
	const uint32_t n_in = P.nhits[depth];
	const int pin = depth & 1, pout = pin ^ 1;
	const uint32_t nthreads = gridDim.x * blockDim.x;
	const uint32_t n_round = (n_in + blockDim.x - 1u) / blockDim.x * blockDim.x;  // uniform trip count per CTA (barriers)
	const bool more_depth = (uint32_t)depth + 1u < P.max_depth;
	const bool light_phase = more_depth && els && (!P.indirect_only || !FIRST);

#if SSB_PIPELINE_SHADE
	// order[] is read two iterations ahead (a register), the records it points at one iteration ahead (cp.async into the
	// thread's slot: [stage][recH.0, recH.1, recR.0, recR.1][thread])
	float4 (*s_pipe)[4][SSB_SHADE_THREADS] = reinterpret_cast<float4 (*)[4][SSB_SHADE_THREADS]>(ssb_smem + pipe_offset(S));  // dynamic shared memory, behind the blob
	auto prefetch = [&](uint32_t it, bool ok, int stage) {
		if (ok) {
			cp_async16(&s_pipe[stage][0][threadIdx.x], &P.recH[2 * (size_t)it]);
			cp_async16(&s_pipe[stage][1][threadIdx.x], &P.recH[2 * (size_t)it + 1]);
			cp_async16(&s_pipe[stage][2][threadIdx.x], &P.recR[pin][2 * (size_t)it]);
			cp_async16(&s_pipe[stage][3][threadIdx.x], &P.recR[pin][2 * (size_t)it + 1]);
		}
		cp_async_commit();
	};
	int stage = 0;
	uint32_t item_cur, item_nxt;
	{
		const uint32_t s0 = blockIdx.x * blockDim.x + threadIdx.x;
		item_cur = s0 < n_in ? P.order[s0] : 0u;
		item_nxt = (s0 < n_in && nthreads < n_in - s0) ? P.order[s0 + nthreads] : 0u;
		prefetch(item_cur, s0 < n_in, 0);
	}
#endif
	for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_round; slot += nthreads) {
		const bool valid = slot < n_in;
#if SSB_PIPELINE_SHADE
		const int cur = stage;
		uint32_t item_nn = 0u;
		{
			const uint32_t left = valid ? n_in - slot : 0u;  // slots from this one to the end of the queue
			prefetch(item_nxt, nthreads < left, stage ^ 1);
			if (nthreads < left && nthreads < left - nthreads) item_nn = P.order[slot + 2u * nthreads];
			cp_async_wait<1>();
			stage ^= 1;
		}
#endif
		bool cont = false;  // path continues to depth+1
		float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 1, lambda_0 = 0;
		int ignore = -1;
		uint32_t id = 0, item = 0;
		Rng rng; rng.state = 0; rng.inc = 1;
		float hx = 0, hy = 0, hz = 0, nx = 0, ny = 0, nz = 1;
		int cur_quad = 0;
		uint32_t mat_kind = SSB_MATERIAL_LAMBERT;
		Hero local, f_s;
		local.v[0] = local.v[1] = local.v[2] = local.v[3] = 0.0f;
		f_s = local;

		// ---- phase 0: gather the path, emission, albedo (renderer.cpp:165-175; material.cpp:120-143)
		if (valid) {
#if SSB_PIPELINE_SHADE
			item = item_cur;
			const float4 h0 = s_pipe[cur][0][threadIdx.x], h1 = s_pipe[cur][1][threadIdx.x];
			const float4 r0 = s_pipe[cur][2][threadIdx.x], r1 = s_pipe[cur][3][threadIdx.x];
#else
			item = P.order[slot];
			const float4 h0 = P.recH[2 * (size_t)item], h1 = P.recH[2 * (size_t)item + 1];
			const float4 r0 = P.recR[pin][2 * (size_t)item], r1 = P.recR[pin][2 * (size_t)item + 1];
#endif
			hx = h0.x; hy = h0.y; hz = h0.z;  // hit position
			const uint32_t hq = __float_as_uint(h0.w);
			id = __float_as_uint(r1.x);
			lambda_0 = r1.y;
			rng.state = ((unsigned long long)__float_as_uint(r0.y) << 32) | __float_as_uint(r0.x);
			rng.inc = ((unsigned long long)__float_as_uint(r0.w) << 32) | __float_as_uint(r0.z);
			cur_quad = (int)(hq & 0x7fffffffu);
			const ssb_quad& quad = S.quads()[cur_quad];
			const ssb_tri& tri = quad.tri[hq >> 31];
			const DevMaterial& m = S.materials()[quad.material];
			mat_kind = m.kind;
			nx = tri.normal[0]; ny = tri.normal[1]; nz = tri.normal[2];
			const float st_x = (h1.x * tri.v[0].st[0] + h1.y * tri.v[1].st[0]) + h1.z * tri.v[2].st[0];
			const float st_y = (h1.x * tri.v[0].st[1] + h1.y * tri.v[1].st[1]) + h1.z * tri.v[2].st[1];
			// emission: last_was_delta is true only for the camera ray (the reference recurses with `false`, :248)
			if (!els || (FIRST && !P.indirect_only)) {
				Hero e = material_emission<UPS>(P, S, m, lambda_0);
#pragma unroll
				for (int c = 0; c < 4; ++c) local.v[c] = local.v[c] + e.v[c];
			}
			if (more_depth) {
				// albedo lookup (the reference does it twice with identical arguments, material.cpp:120-143)
				f_s = material_albedo<UPS>(P, S, m, st_x, st_y, lambda_0);
				if (mat_kind == SSB_MATERIAL_LAMBERT) {
#pragma unroll
					for (int c = 0; c < 4; ++c) f_s.v[c] = div_const_or_zero(f_s.v[c], [](float x) { return x / SSB_PI_F; });
				}
			}
		}
		SSB_PHASE_BARRIER_AT(0);

		// ---- phase 1: light sample (renderer.cpp:182-191; Scene::get_rand_toward_light, scene.cpp:417-431)
		float sx = 0, sy = 0, sz = 1, pdf = 1.0f, l_ndl = 0.0f;
		int light_quad = -1;
		if (valid && light_phase) {
			uint32_t li = rand_choice(rng, S.hdr()->nlights);
			light_quad = (int)S.lights()[li];
			const int lt = (rand_1f(rng) <= 0.5f) ? 0 : 1;
			float r0 = rand_1f(rng);
			float r1 = rand_1f(rng);
			const float4 ls = sample_spherical_triangle(light_quad, lt, hx, hy, hz, r0, r1);
			sx = ls.x; sy = ls.y; sz = ls.z; pdf = ls.w;
			pdf *= 0.5f;
			{  // pdf /= nlights (scene.cpp:428-430); inf / n = inf without the division's slow path
				const bool pinf = pdf == __int_as_float(0x7f800000);
				float pn = pinf ? 1.0f : pdf;
				asm("" : "+f"(pn));
				pn /= (float)S.hdr()->nlights;
				pdf = pinf ? pdf : pn;
			}
			l_ndl = dot3(sx, sy, sz, nx, ny, nz);
		}
		SSB_PHASE_BARRIER_AT(1);

		// ---- phase 2: shadow query + direct contribution (renderer.cpp:192-218)
		if (valid && light_phase && l_ndl > 0.0f) {
			Hit hs;
			scene_query<LIST>(S, eps, cur_quad, hs, hx, hy, hz, sx, sy, sz);
			if (hs.quad == light_quad) {
				const DevMaterial& lm = S.materials()[S.quads()[light_quad].material];
				Hero emitted = material_emission<UPS>(P, S, lm, lambda_0);
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					float fe = (mat_kind == SSB_MATERIAL_LAMBERT) ? f_s.v[c] : 0.0f;  // MaterialMirror::evaluate_bsdf = 0
					local.v[c] = local.v[c] + div_or_zero((emitted.v[c] * l_ndl) * fe, pdf);
				}
			}
		}
		SSB_PHASE_BARRIER_AT(2);

		// ---- phase 3: interact_bsdf + recursion decision + records (renderer.cpp:222-251)
		if (valid) {
			Hero rad;  // value this L() call returns if the path ends here
			rad.v[0] = rad.v[1] = rad.v[2] = rad.v[3] = 0.0f;
			int nrec = depth;  // fold records written by shallower depths
			if (more_depth) {
				float wix, wiy, wiz, pdf_w_i;
				if (mat_kind == SSB_MATERIAL_LAMBERT) {
					float cx, cy, cz;  // Math::rand_coshemi (random.cpp:29-49)
					do {
						float angle = rand_1f(rng) * (2.0f * SSB_PI_F);
						const float2 sc_angle = sincosf_x(angle);
						const float si = sc_angle.x, co = sc_angle.y;
						float radius_sq = rand_1f(rng);
						float radius = sqrtf(radius_sq);
						cx = radius * co; cy = sqrtf(1.0f - radius_sq); cz = radius * si;
						pdf_w_i = cy;
					} while (pdf_w_i <= eps);
					pdf_w_i *= 1.0f / SSB_PI_F;
					// Math::get_rotated_to / get_basis (math-helpers.hpp:14-38)
					float sign = copysignf(1.0f, nz);
					float a = -1.0f / (sign + nz);
					float b = nx * ny * a;
					float bxx = 1.0f + sign * nx * nx * a, bxy = sign * b, bxz = -sign * nx;
					float bzx = b, bzy = sign + ny * ny * a, bzz = -ny;
					wix = (cx * bxx + cy * nx) + cz * bzx;
					wiy = (cx * bxy + cy * ny) + cz * bzy;
					wiz = (cx * bxz + cy * nz) + cz * bzz;
				} else {
					// MaterialMirror::interact_bsdf (material.cpp:154-167): reflect(w_o = -ray.dir, N)
					const float4 din = P.recA[pin][2 * (size_t)item + 1];
					float vx = -din.x, vy = -din.y, vz = -din.z;
					float d2 = 2.0f * dot3(vx, vy, vz, nx, ny, nz);
					wix = -vx + d2 * nx; wiy = -vy + d2 * ny; wiz = -vz + d2 * nz;
					pdf_w_i = __int_as_float(0x7f800000);
				}
				bool recurse = false;
				float n_dot_l = 0.0f;
				float ff = (f_s.v[0] * f_s.v[0] + f_s.v[1] * f_s.v[1]) + (f_s.v[2] * f_s.v[2] + f_s.v[3] * f_s.v[3]);
				if (ff > 0.0f) {
					if (isfinite(pdf_w_i)) n_dot_l = dot3(wix, wiy, wiz, nx, ny, nz);
					else { n_dot_l = 1.0f; pdf_w_i = 1.0f; }
					recurse = n_dot_l > 0.0f;
				}
				if (recurse) {
					float4* rec = P.stk + 4 * ((size_t)depth * P.total_work + id);
					st_sector(rec, make_float4(local.v[0], local.v[1], local.v[2], local.v[3]), make_float4(f_s.v[0], f_s.v[1], f_s.v[2], f_s.v[3]));
					st_sector(rec + 2, make_float4(n_dot_l, pdf_w_i, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f));
					nrec = depth + 1;
					// Dead-work skip (result-identical): with explicit light sampling the L() call at the last depth
					// can add neither emission (last_was_delta == false) nor children: it returns 0, hit_anything is
					// already set and no random numbers are drawn there.  rad stays 0.
					if (!(els && (uint32_t)depth + 2u >= P.max_depth)) {
						cont = true;
						ox = hx; oy = hy; oz = hz; dx = wix; dy = wiy; dz = wiz; ignore = cur_quad;
					}
				} else {
					rad = local;
				}
			} else {
				rad = local;
			}
			// ---- the path ends here: the deepest L() call returns `rad`, `nrec` records wait to be folded
			if (!cont) {
				st_sector(&P.leaf[2 * (size_t)id], make_float4(rad.v[0], rad.v[1], rad.v[2], rad.v[3]), make_float4(lambda_0, __int_as_float(nrec | (1 << 16)), 0.f, 0.f));
			}
		}

		// ---- compaction: surviving paths go to dense positions of the next depth's queue
		const unsigned mask = __ballot_sync(full, cont);
		if (mask) {
			unsigned base = 0;
			const int leader = __ffs(mask) - 1;
			if (lane == leader) base = atomicAdd(&P.counts[depth + 1], (uint32_t)__popc(mask));
			base = __shfl_sync(full, base, leader);
			if (cont) {
				const uint32_t o = base + __popc(mask & ((1u << lane) - 1u));
				st_sector(&P.recA[pout][2 * (size_t)o], make_float4(ox, oy, oz, __int_as_float(ignore)), make_float4(dx, dy, dz, lambda_0));
				st_sector(&P.recR[pout][2 * (size_t)o],
				          make_float4(__uint_as_float((uint32_t)rng.state), __uint_as_float((uint32_t)(rng.state >> 32)),
				                      __uint_as_float((uint32_t)rng.inc), __uint_as_float((uint32_t)(rng.inc >> 32))),
				          make_float4(__uint_as_float(id), lambda_0, 0.f, 0.f));
			}
		}
#if SSB_PIPELINE_SHADE
		item_cur = item_nxt; item_nxt = item_nn;
#endif
		SSB_PHASE_BARRIER_AT(3);
	}
}


// Unwind the recursion of one sample and convert it to the value the reference's _render_sample returns:
//   radiance = local + ((child * n.l) * f_s) / pdf        (renderer.cpp:216,248)
//   XYZ = specradflux_to_ciexyz(radiance, lambda_0)        (color.hpp:115-139)      [RGB mode: the l-RGB flux itself]
// One thread per sample (ncu on the one-thread-per-pixel form: a single wave of 262 k threads, `long_scoreboard` 6.2 —
// latency bound at half the DRAM rate); lanes = neighbouring pixels of the same sample index, so every record array is
// read coalesced.  The float4 (X,Y,Z,hit) goes to `samples`, which aliases path-state memory that is dead by now.
#ifndef SSB_FOLD_THREADS
#define SSB_FOLD_THREADS 256
#endif
__global__ void __launch_bounds__(SSB_FOLD_THREADS) ssb_fold_kernel(const __grid_constant__ KParams P) {
	const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (id >= P.total_work) return;
	const DevHeader* hdr = reinterpret_cast<const DevHeader*>(P.blob);
	const float* pool = reinterpret_cast<const float*>(P.blob + hdr->off_pool);
	const DevSpectrum sx = hdr->xbar, sy = hdr->ybar, sz = hdr->zbar;
	float4 lf, mt;
	ld_sector(&P.leaf[2 * id], lf, mt);
	const float lambda_0 = mt.x;
	const int info = __float_as_int(mt.y);
	const int nrec = info & 0xffff;
	float r0 = lf.x, r1 = lf.y, r2 = lf.z, r3 = lf.w;
	for (int d = nrec - 1; d >= 0; --d) {
		const float4* rec = P.stk + 4 * ((size_t)d * P.total_work + id);
		float4 lo, f;
		ld_sector(rec, lo, f);
		const float4 np = rec[2];
		// Exact shortcut: the child radiance is +0 in all four channels (miss / skipped last depth — the common
		// case at the deepest record) and n.l, pdf are positive finite, f_s non-negative finite: then
		// ((+0*n.l)*f_s)/pdf is +0 and local + (+0) == local bit for bit (local is never -0: it is a sum that
		// starts at +0).  Saves four IEEE divisions that would take the zero-numerator slow path.
		if (__float_as_uint(r0) == 0u && __float_as_uint(r1) == 0u && __float_as_uint(r2) == 0u && __float_as_uint(r3) == 0u &&
		    np.x > 0.0f && np.x < __int_as_float(0x7f800000) && np.y > 0.0f && np.y < __int_as_float(0x7f800000) &&
		    f.x >= 0.0f && f.y >= 0.0f && f.z >= 0.0f && f.w >= 0.0f &&
		    f.x < __int_as_float(0x7f800000) && f.y < __int_as_float(0x7f800000) && f.z < __int_as_float(0x7f800000) && f.w < __int_as_float(0x7f800000)) {
			r0 = lo.x; r1 = lo.y; r2 = lo.z; r3 = lo.w;
			continue;
		}
		r0 = lo.x + ((r0 * np.x) * f.x) / np.y;
		r1 = lo.y + ((r1 * np.x) * f.y) / np.y;
		r2 = lo.z + ((r2 * np.x) * f.z) / np.y;
		r3 = lo.w + ((r3 * np.x) * f.w) / np.y;
	}
	if (!P.flat_field) {
		const float s = P.ff[id];
		r0 *= s; r1 *= s; r2 *= s; r3 *= s;
	}
	const float hitf = (info >> 16) ? 1.0f : 0.0f;
	if (P.render_mode == SSB_RENDER_RGB) {  // renderer.cpp:274-275: the sample is (l-RGB flux, hit)
		P.samples[id] = make_float4(r0, r1, r2, hitf);
		return;
	}
	float rad[4] = { r0, r1, r2, r3 };
	float X = 0.0f, Y = 0.0f, Z = 0.0f;
#if SSB_SHARED_GRID
	const bool shared_grid = spec_same_grid(sx, sy, sz);  // uniform; true for the reference's observer tables
#else
	const bool shared_grid = false;
#endif
#pragma unroll
	for (int c = 0; c < 4; ++c) {
		const float lambda = lambda_0 + (float)c * P.lambda_step;
		float xb, yb, zb;
		if (shared_grid) spec_sample3(pool, sx, sy, sz, lambda, xb, yb, zb);
		else { xb = spec_sample(pool, sx, lambda); yb = spec_sample(pool, sy, lambda); zb = spec_sample(pool, sz, lambda); }
		X += (xb * rad[c]) * P.lambda_step;
		Y += (yb * rad[c]) * P.lambda_step;
		Z += (zb * rad[c]) * P.lambda_step;
	}
	P.samples[id] = make_float4(X, Y, Z, hitf);
}

// Renderer::_render_pixel's sample loop (renderer.cpp:292-295 / 301-303): per pixel, IN SAMPLE ORDER,
//   avg += double4(sample * 0.001f)      [RGB mode: avg += double4(sample)]
// One thread per pixel of the pass rectangle; lanes = neighbouring pixels (coalesced float4 reads).
__global__ void __launch_bounds__(128) ssb_accumulate_kernel(const __grid_constant__ KParams P) {
	const uint32_t npix_rect = P.rect_w * P.rect_h;
	const uint32_t pr = blockIdx.x * blockDim.x + threadIdx.x;
	if (pr >= npix_rect) return;
	const uint32_t pi = P.x0 + pr % P.rect_w, pj = image_row(P, pr / P.rect_w);
	double* a = P.accum + 4 * ((size_t)pj * P.width + pi);
	double a0 = a[0], a1 = a[1], a2 = a[2], a3 = a[3];
	const bool rgb = P.render_mode == SSB_RENDER_RGB;
#pragma unroll 8
	for (uint32_t kk = 0; kk < P.nsamp; ++kk) {
		const float4 s = P.samples[(size_t)kk * npix_rect + pr];
		if (rgb) { a0 += (double)s.x; a1 += (double)s.y; a2 += (double)s.z; a3 += (double)s.w; }
		else { a0 += (double)(s.x * 0.001f); a1 += (double)(s.y * 0.001f); a2 += (double)(s.z * 0.001f); a3 += (double)(s.w * 0.001f); }
	}
	a[0] = a0; a[1] = a1; a[2] = a2; a[3] = a3;
}

// ssb_accum_merge: dst (+)= src over the pixels src rendered.  mode 0: dst += src (sample shards of the whole frame);
// mode 1: the pixel rectangle [x0,x1) x [y0,y1) is copied; mode 2: the rows of src's interleaved bands are copied.
// `src` may be a peer GPU's memory (direct NVLink loads) or a staging copy on this device.  One thread per pixel, the four
// doubles of a pixel as two 16-byte accesses.
__global__ void __launch_bounds__(256) ssb_accum_merge_kernel(double* __restrict__ dst, const double* __restrict__ src, uint32_t width, uint32_t height,
                                                              uint32_t mode, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1,
                                                              uint32_t band_h, uint32_t band_n, uint32_t band_i) {
	const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= (size_t)width * height) return;
	const uint32_t i = (uint32_t)(p % width), j = (uint32_t)(p / width);
	const double2* s2 = reinterpret_cast<const double2*>(src) + 2 * p;
	double2* d2 = reinterpret_cast<double2*>(dst) + 2 * p;
	if (mode == 0u) {
		const double2 a = s2[0], b = s2[1];
		double2 c = d2[0], d = d2[1];
		c.x += a.x; c.y += a.y; d.x += b.x; d.y += b.y;
		d2[0] = c; d2[1] = d;
		return;
	}
	const bool mine = mode == 1u ? (i >= x0 && i < x1 && j >= y0 && j < y1) : (i >= x0 && i < x1 && (j / band_h) % band_n == band_i);
	if (mine) { d2[0] = s2[0]; d2[1] = s2[1]; }
}

// ssb_debug_intersect: closest-hit queries of caller-supplied rays through the render kernels' scene_intersect
__global__ void __launch_bounds__(256) ssb_debug_intersect_kernel(const __grid_constant__ KParams P, const float* __restrict__ rays, const int32_t* __restrict__ ignore,
                                                                  float* __restrict__ out, size_t n) {
	__shared__ __align__(8) unsigned long long blob_bar;
	stage_scene(P, &blob_bar);
	const SceneView S;
	for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (size_t)gridDim.x * blockDim.x) {
		Hit hit;
		if (P.scan_list) scene_query<true>(S, P.eps, ignore ? ignore[r] : -1, hit, rays[6 * r], rays[6 * r + 1], rays[6 * r + 2], rays[6 * r + 3], rays[6 * r + 4], rays[6 * r + 5]);
		else scene_query<false>(S, P.eps, ignore ? ignore[r] : -1, hit, rays[6 * r], rays[6 * r + 1], rays[6 * r + 2], rays[6 * r + 3], rays[6 * r + 4], rays[6 * r + 5]);
		out[6 * r] = __int_as_float(hit.quad); out[6 * r + 1] = __int_as_float(hit.tri); out[6 * r + 2] = hit.dist;
		out[6 * r + 3] = hit.bx; out[6 * r + 4] = hit.by; out[6 * r + 5] = hit.bz;
	}
}

// avg *= 1000/spp; framebuffer = (ciexyz_to_srgb(float3(avg)), float(avg.a)) (renderer.cpp:296-298, color.cpp:237-257)
__global__ void ssb_resolve_kernel(const double* __restrict__ accum, double* __restrict__ xyza, float4* __restrict__ srgba,
                                   uint32_t npix, double scale, double spp_rgb, uint32_t upsampling, float d65_rad_Y,
                                   float m0, float m1, float m2, float m3, float m4, float m5, float m6, float m7, float m8) {
	uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= npix) return;
	double avg[4];
#pragma unroll
	for (int c = 0; c < 4; ++c) {
		// spp_rgb > 0 selects RENDER_MODE_RGB: avg /= double(spp) (renderer.cpp:304) instead of avg *= 1000/spp (:296)
		const double a = accum[4 * (size_t)p + c];
		avg[c] = spp_rgb > 0.0 ? a / spp_rgb : a * scale;
	}
	if (xyza) {
#pragma unroll
		for (int c = 0; c < 4; ++c) xyza[4 * (size_t)p + c] = avg[c];
	}
	if (srgba) {
		float x = (float)avg[0], y = (float)avg[1], z = (float)avg[2];
		float lr, lg, lb;
		if (spp_rgb > 0.0) {  // Color::lrgb_to_srgb(lRGB_F32(avg)), renderer.cpp:306
			lr = x; lg = y; lb = z;
		} else if (upsampling == SSB_UPSAMPLE_MENG) {
			x = x / d65_rad_Y; y = y / d65_rad_Y; z = z / d65_rad_Y;
			lr = (3.24156456f * x + -1.53766524f * y) + -0.49870224f * z;
			lg = (-0.96920119f * x + 1.87588535f * y) + 0.04155324f * z;
			lb = (0.05562416f * x + -0.20395525f * y) + 1.05685902f * z;
		} else {
			lr = (m0 * x + m3 * y) + m6 * z;
			lg = (m1 * x + m4 * y) + m7 * z;
			lb = (m2 * x + m5 * y) + m8 * z;
		}
		float4 o;
		o.x = lr < 0.0031308f ? 12.92f * lr : 1.055f * ssbm::powf_exact(lr, 1.0f / 2.4f) - 0.055f;
		o.y = lg < 0.0031308f ? 12.92f * lg : 1.055f * ssbm::powf_exact(lg, 1.0f / 2.4f) - 0.055f;
		o.z = lb < 0.0031308f ? 12.92f * lb : 1.055f * ssbm::powf_exact(lb, 1.0f / 2.4f) - 0.055f;
		o.w = (float)avg[3];
		srgba[p] = o;
	}
}

// sRGB_ReflectanceTexture keeps RGB8 scanlines (material.hpp:20-21); the device copy is RGBA8 so that a texel is
// one aligned 32-bit load
__global__ void ssb_repack_rgb8_kernel(const unsigned char* __restrict__ rgb, uchar4* __restrict__ rgba, size_t n) {
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	rgba[i] = make_uchar4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], 255);
}

// ssb_options.prebaked_textures: the first step of Color::lrgb_to_specrefl's JH build (texel -> sRGB LUT -> rgb2spec_fetch,
// material.cpp:51-55 + color.cpp:218-220) once per texel; the 32-bit coefficient texture the reference's comment
// (color.cpp:204-216,222-223) describes.  Same device function as the per-lookup path: the same floats.
__global__ void __launch_bounds__(256) ssb_bake_jh_kernel(const uchar4* __restrict__ rgba, float4* __restrict__ coef, size_t n,
                                                          const unsigned char* __restrict__ blob, const float* __restrict__ jh_scale,
                                                          const float* __restrict__ jh_data, uint32_t jh_res) {
	const float* lut = reinterpret_cast<const DevHeader*>(blob)->srgb_lut;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const uchar4 px = rgba[i];
		float c[3];
		jh_fetch(jh_scale, jh_data, (int)jh_res, __ldg(lut + px.x), __ldg(lut + px.y), __ldg(lut + px.z), c);
		coef[i] = make_float4(c[0], c[1], c[2], 0.0f);
	}
}

__global__ void ssb_eval_math_kernel(uint32_t fn, const float* __restrict__ x, float arg, float* __restrict__ out, size_t n) {
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float v = x[i], r;
	switch (fn) {
		case 0: r = ssbm::sinf_exact(v); break;
		case 1: r = ssbm::cosf_exact(v); break;
		case 2: r = ssbm::acosf_exact(v); break;
		case 4: { float s, c; ssbm::sincosf_exact(v, &s, &c); r = s; break; }  // the paired form used by the kernels
		case 5: { float s, c; ssbm::sincosf_exact(v, &s, &c); r = c; break; }
		case 6: r = acosf2_x(v, __uint_as_float(__float_as_uint(v) * 2654435761u)).x; break;  // the paired acosf, either lane
		case 7: r = acosf2_x(__uint_as_float(__float_as_uint(v) * 2654435761u), v).y; break;
		default: r = ssbm::powf_exact(v, arg); break;
	}
	out[i] = r;
}

// acosf_exact2 against acosf_exact over ALL 2^32 arguments in either lane (the other lane gets a permutation of the
// argument): counts lanes whose bits differ (two NaNs count as equal)
__global__ void ssb_acos_pair_check_kernel(unsigned long long* mismatches, float* examples, uint32_t max_examples) {
	unsigned long long bad = 0;
	const unsigned long long total = 1ull << 32, stride = (unsigned long long)gridDim.x * blockDim.x;
	for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
		const float x0 = __uint_as_float((uint32_t)i), x1 = __uint_as_float((uint32_t)i * 2654435761u + 12345u);
		const float2 got = acosf2_x(x0, x1);  // the non-inlined function the shade stage calls
		const float w0 = ssbm::acosf_exact(x0), w1 = ssbm::acosf_exact(x1);
		const bool b0 = __float_as_uint(got.x) != __float_as_uint(w0) && !(got.x != got.x && w0 != w0);
		const bool b1 = __float_as_uint(got.y) != __float_as_uint(w1) && !(got.y != got.y && w1 != w1);
		bad += (b0 ? 1u : 0u) + (b1 ? 1u : 0u);
		if ((b0 || b1) && examples) {  // a few offending arguments, for the failure message: (argument, paired result, scalar result)
			const unsigned long long k = atomicAdd(mismatches + 1, 1ull);
			if (k < max_examples) { examples[3 * k] = b0 ? x0 : x1; examples[3 * k + 1] = b0 ? got.x : got.y; examples[3 * k + 2] = b0 ? w0 : w1; }
		}
	}
	if (bad) atomicAdd(mismatches, bad);
}

}  // namespace ssbk
