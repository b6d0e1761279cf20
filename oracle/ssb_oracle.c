/* ssb_oracle.c — TEST INFRASTRUCTURE ONLY.  Not part of the product; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * A plain-C, CPU restatement of the reference's per-pixel Monte-Carlo hot path
 * (geometrian/simple-spectral, src/renderer.cpp:103-308 and everything it reaches), written
 * against the same flat POD inputs as the product's C ABI (include/ssb200.h) so that the CUDA
 * path and this oracle are fed identical bytes.  Each function cites the reference file:line
 * it follows.  Operation order is part of the contract: build with
 * `gcc -O2 -ffp-contract=off` and no -march so that nothing is fused or vectorised differently
 * from the reference's own `-O2` build; libm's sinf/cosf/acosf/powf are the reference's.
 *
 * Pinning: this file is checked bit-for-bit against the real reference compiled here
 * (oracle/_ref/simple_spectral_*_hooked, see oracle/build_ref.py) at matched per-sample seeds:
 * tests/test_oracle_golden.py (live, when oracle/_ref exists) and the committed fixtures under
 * tests/golden/ (made by tests/golden/make_golden.py); and, beyond the reference's three scenes,
 * on random scenes pushed through the real reference (tests/test_oracle_random_scenes.py: live
 * runs of the hooked binaries with SSB_SCENE_QUADS, four of them committed as fixtures).
 *
 * Third-party arithmetic the reference inherits and that is NOT under /root/reference:
 *   GLM (unpinned)           — restated as GLM 0.9.9 scalar semantics (oracle/glm_shim)
 *   libstdc++ 13 <random>    — uniform_real_distribution / uniform_int_distribution mappings,
 *                              restated in rand_1f / rand_1d / rand_choice below
 *   glibc 2.39 libm          — called directly
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/ssb200.h"
#include "ssb_oracle.h"

typedef struct { float x, y, z; } v3;
typedef struct { float x, y; } v2;
typedef struct { float v[4]; } hero; /* _Spectrum::HeroSample, spectrum.hpp:17 */

/* ------------------------------------------------------------------ GLM scalar semantics */
static inline v3 v3_make(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_scale(v3 a, float s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline v3 v3_neg(v3 a) { return v3_make(-a.x, -a.y, -a.z); }
static inline float v3_dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float v3_get(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
static inline float glm_min(float a, float b) { return (b < a) ? b : a; }
static inline float glm_max(float a, float b) { return (a < b) ? b : a; }
static inline float glm_clamp(float x, float lo, float hi) { return glm_min(glm_max(x, lo), hi); }
static inline v3 v3_normalize(v3 a) { return v3_scale(a, 1.0f / sqrtf(v3_dot(a, a))); }

static inline hero hero_splat(float s) { hero h = { { s, s, s, s } }; return h; }
static inline hero hero_add(hero a, hero b) { hero r; for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] + b.v[i]; return r; }
static inline hero hero_mul(hero a, hero b) { hero r; for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] * b.v[i]; return r; }
static inline hero hero_scale(hero a, float s) { hero r; for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] * s; return r; }
static inline hero hero_div(hero a, float s) { hero r; for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] / s; return r; }
static inline float hero_dot(hero a, hero b) {
	return (a.v[0] * b.v[0] + a.v[1] * b.v[1]) + (a.v[2] * b.v[2] + a.v[3] * b.v[3]);
}

/* ------------------------------------------------------------------ RNG (util/random.hpp:16-78) */
typedef struct { uint64_t state, inc; } rng_t;

static inline uint32_t rng_next(rng_t* r) { /* PCG32 XSH-RR, random.hpp:53-59 */
	uint64_t s = r->state;
	uint32_t xorshifted = (uint32_t)(((s >> 18u) ^ s) >> 27u);
	int rot = (int)(s >> 59u);
	uint32_t result = (xorshifted >> rot) | (xorshifted << ((-rot) & 31));
	r->state = s * 6364136223846793005ull + r->inc;
	return result;
}
/* std::uniform_real_distribution<float>()(rng), libstdc++ 13 generate_canonical<float,24>:
 * one 32-bit draw, float(u)/2^32, clamped below 1 (random.hpp:68-70). */
static inline float rand_1f(rng_t* r) {
	float sum = (float)rng_next(r);
	float ret = sum / 4294967296.0f;
	if (ret >= 1.0f) ret = nextafterf(1.0f, 0.0f);
	return ret;
}
/* std::uniform_real_distribution<double>: two draws, (u0 + u1*2^32)/2^64 (random.hpp:71-73). */
static inline double rand_1d(rng_t* r) {
	double sum = (double)rng_next(r);
	sum += (double)rng_next(r) * 4294967296.0;
	double ret = sum / 18446744073709551616.0;
	if (ret >= 1.0) ret = nextafter(1.0, 0.0);
	return ret;
}
/* std::uniform_int_distribution<size_t>(0,n-1): Lemire's nearly-divisionless method on a
 * 32-bit generator (libstdc++ 13 bits/uniform_int_dist.h, _S_nd<uint64_t>) (random.hpp:75-78). */
static inline uint32_t rand_choice(rng_t* r, uint32_t n) {
	uint64_t product = (uint64_t)rng_next(r) * (uint64_t)n;
	uint32_t low = (uint32_t)product;
	if (low < n) {
		uint32_t threshold = (uint32_t)(-n) % n;
		while (low < threshold) {
			product = (uint64_t)rng_next(r) * (uint64_t)n;
			low = (uint32_t)product;
		}
	}
	return (uint32_t)(product >> 32);
}

/* per-sample seeding — the scheme spliced into the hooked reference (oracle/ref_hooks.hpp) */
static inline uint64_t mix64(uint64_t z) {
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
static inline void seed_sample(rng_t* r, uint64_t seed, uint64_t sample_index) {
	r->state = mix64(seed ^ mix64(sample_index));
	r->inc = mix64(r->state) | 1ull;
}

/* ------------------------------------------------------------------ spectra (spectrum.cpp:11-67) */
typedef struct {
	const ssb_scene* scene;
	const ssb_color* color;
	const ssb_options* opt;
	float lambda_step;  /* LAMBDA_STEP, stdafx.hpp:289 */
	int nw;             /* SAMPLE_WAVELENGTHS (stdafx.hpp:90): channels nw..3 of a `hero` stay 0 (glm::vec<nw,float>) */
	uint32_t nlights;
	uint32_t lights[SSB_MAX_LIGHTS]; /* Scene::lights, scene.cpp:27-29 */
	ssb_oracle_counters* counters;
} octx;

static float spectrum_sample_linear(const ssb_spectrum* s, float lambda) { /* spectrum.cpp:39-60 */
	float numer = s->high - s->low;               /* spectrum.cpp:22-25 */
	float denom = (float)(s->n - 1);
	float delta_lambda_recip = denom / numer;
	float i = (lambda - s->low) * delta_lambda_recip;
	float i0f = floorf(i);
	float frac = i - i0f;
	int i0 = (int)i0f;
	int i1 = i0 + 1;
	float val0 = (i0 >= 0 && (uint32_t)i0 < s->n) ? s->data[i0] : 0.0f;
	float val1 = (i1 >= 0 && (uint32_t)i1 < s->n) ? s->data[i1] : 0.0f;
	return val0 * (1.0f - frac) + val1 * frac; /* Math::lerp, math-helpers.hpp:10-12 */
}
static float spectrum_sample_nearest(const ssb_spectrum* s, float lambda) { /* spectrum.cpp:29-38 */
	float numer = s->high - s->low;
	float denom = (float)(s->n - 1);
	float delta_lambda_recip = denom / numer;
	float i_f = (lambda - s->low) * delta_lambda_recip;
	i_f = roundf(i_f);
	int i_i = (int)i_f;
	if (i_i >= 0 && (uint32_t)i_i < s->n) return s->data[i_i];
	return 0.0f;
}
static hero spectrum_hero(const octx* c, const ssb_spectrum* s, float lambda_0) { /* spectrum.cpp:61-67 */
	hero result = { { 0.0f, 0.0f, 0.0f, 0.0f } };
	for (int i = 0; i < c->nw; ++i) {
		float lambda = lambda_0 + (float)i * c->lambda_step;
		result.v[i] = (s->filter == SSB_FILTER_NEAREST) ? spectrum_sample_nearest(s, lambda)
		                                                : spectrum_sample_linear(s, lambda);
	}
	return result;
}

/* ------------------------------------------------------------------ colour (util/color.*) */
static inline float srgb_to_lrgb_1(float c) { /* color.hpp:91-97 */
	return c < 0.04045f ? c / 12.92f : powf((c + 0.055f) / 1.055f, 2.4f);
}
static inline float lrgb_to_srgb_1(float c) { /* color.hpp:84-90 */
	return c < 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
}

/* Jakob & Hanika 2019: rgb2spec.c:56-75 */
static int jh_find_interval(const float* values, int size_, float x) {
	int left = 0, last_interval = size_ - 2, size = last_interval;
	while (size > 0) {
		int half = size >> 1, middle = left + half + 1;
		if (values[middle] < x) { left = middle; size -= half + 1; }
		else size = half;
	}
	return left < last_interval ? left : last_interval;
}
/* rgb2spec.c:77-118 (no FMA: the parity build of the reference has no __FMA__) */
static void jh_fetch(const ssb_color* col, const float rgb[3], float out[3]) {
	int i = 0, res = (int)col->jh_res;
	for (int j = 1; j < 3; ++j) if (rgb[j] >= rgb[i]) i = j;
	float z = rgb[i], scale = (res - 1) / z, x = rgb[(i + 1) % 3] * scale, y = rgb[(i + 2) % 3] * scale;
	/* (uint32_t)NaN (black texel: 0*inf) is UB in C; x86-64 gcc yields 0 and so does the GPU's cvt.rzi */
	uint32_t xu = (x != x) ? 0u : (uint32_t)x, yu = (y != y) ? 0u : (uint32_t)y;
	uint32_t xi = xu < (uint32_t)(res - 2) ? xu : (uint32_t)(res - 2);
	uint32_t yi = yu < (uint32_t)(res - 2) ? yu : (uint32_t)(res - 2);
	uint32_t zi = (uint32_t)jh_find_interval(col->jh_scale, res, z);
	uint32_t offset = (((i * res + zi) * res + yi) * res + xi) * 3, dx = 3, dy = 3 * res, dz = 3 * res * res;
	float x1 = x - xi, x0 = 1.f - x1, y1 = y - yi, y0 = 1.f - y1;
	float z1 = (z - col->jh_scale[zi]) / (col->jh_scale[zi + 1] - col->jh_scale[zi]), z0 = 1.f - z1;
	const float* d = col->jh_data;
	for (int j = 0; j < 3; ++j) {
		out[j] = ((d[offset] * x0 + d[offset + dx] * x1) * y0 + (d[offset + dy] * x0 + d[offset + dy + dx] * x1) * y1) * z0 +
		         ((d[offset + dz] * x0 + d[offset + dz + dx] * x1) * y0 + (d[offset + dz + dy] * x0 + d[offset + dz + dy + dx] * x1) * y1) * z1;
		offset++;
	}
}
static float jh_eval_precise(const float coeff[3], float lambda) { /* rgb2spec.c:129-133 */
	float x = (coeff[0] * lambda + coeff[1]) * lambda + coeff[2];
	float y = 1.f / sqrtf(x * x + 1.f);
	return (.5f * x) * y + .5f;
}

/* Meng et al. 2015: spectrum_grid.h:13-137 */
static float meng_xyz_to_p(const ssb_meng_tables* m, float lambda, const float* xyz) {
	float xyY[3], uv[2];
	const float norm = 1.0 / (xyz[0] + xyz[1] + xyz[2]); /* double divide, spectrum_grid.h:19 */
	if (!(norm < 3.402823466e+38f)) return 0.0f;
	xyY[0] = xyz[0] * norm; xyY[1] = xyz[1] * norm; xyY[2] = xyz[1];
	uv[0] = m->xy_to_uv[0] * xyY[0] + m->xy_to_uv[1] * xyY[1] + m->xy_to_uv[2];
	uv[1] = m->xy_to_uv[3] * xyY[0] + m->xy_to_uv[4] * xyY[1] + m->xy_to_uv[5];
	if (uv[0] < 0.0f || uv[0] >= m->grid_w || uv[1] < 0.0f || uv[1] >= m->grid_h) return 0.f;
	int uvi[2] = { (int)uv[0], (int)uv[1] };
	const int cell_idx = uvi[0] + (int)m->grid_w * uvi[1];
	const int32_t* cell = m->grid + 8 * cell_idx;
	const int inside = cell[0], num = cell[1];
	const int32_t* idx = cell + 2;
	const size_t stride = 5 + m->nsamples;
	float p[6];
	const int ns = (int)m->nsamples;
	const float sb = (lambda - m->sample_min) / (m->sample_max - m->sample_min) * (ns - 1);
	const int sb0 = (int)sb;
	const int sb1 = sb + 1 < ns ? (int)(sb + 1) : ns - 1;
	const float sbf = sb - sb0;
	for (int i = 0; i < num; ++i) {
		const float* spectrum = m->points + stride * idx[i] + 5;
		p[i] = spectrum[sb0] * (1.0f - sbf) + spectrum[sb1] * sbf;
	}
	float interpolated_p = 0.0f;
	if (inside) {
		uv[0] -= uvi[0]; uv[1] -= uvi[1];
		interpolated_p = p[0] * (1.0f - uv[0]) * (1.0f - uv[1]) + p[2] * (1.0f - uv[0]) * uv[1] +
		                 p[3] * uv[0] * uv[1] + p[1] * uv[0] * (1.0f - uv[1]);
	} else {
#define MENG_UV(k, c) (m->points[stride * idx[k] + 3 + (c)])
		const float ex = uv[0] - MENG_UV(0, 0), ey = uv[1] - MENG_UV(0, 1);
		float e0x = MENG_UV(1, 0) - MENG_UV(0, 0), e0y = MENG_UV(1, 1) - MENG_UV(0, 1);
		float uu = e0x * ey - ex * e0y;
		for (int i = 0; i < num - 1; i++) {
			float e1x, e1y;
			if (i == num - 2) { e1x = MENG_UV(1, 0) - MENG_UV(0, 0); e1y = MENG_UV(1, 1) - MENG_UV(0, 1); }
			else { e1x = MENG_UV(i + 2, 0) - MENG_UV(0, 0); e1y = MENG_UV(i + 2, 1) - MENG_UV(0, 1); }
			float vv = ex * e1y - e1x * ey;
			const float area = e0x * e1y - e1x * e0y;
			const float u = uu / area, v = vv / area;
			float w = 1.0f - u - v;
			if (u < 0.0 || v < 0.0 || w < 0.0) { uu = -vv; e0x = e1x; e0y = e1y; continue; }
			interpolated_p = p[0] * w + p[i + 1] * v + p[(i == num - 2) ? 1 : (i + 2)] * u;
			break;
		}
#undef MENG_UV
	}
	return interpolated_p / norm;
}

static hero lrgb_to_specrefl(const octx* c, const float lrgb[3], float lambda_0) {
	const ssb_color* col = c->color;
	hero result = { { 0.0f, 0.0f, 0.0f, 0.0f } };
	if (c->opt->upsampling == SSB_UPSAMPLE_OURS) { /* color.cpp:166-173 */
		hero br = spectrum_hero(c, &col->basis_r, lambda_0);
		hero bg = spectrum_hero(c, &col->basis_g, lambda_0);
		hero bb = spectrum_hero(c, &col->basis_b, lambda_0);
		for (int i = 0; i < c->nw; ++i) result.v[i] = (lrgb[0] * br.v[i] + lrgb[1] * bg.v[i]) + lrgb[2] * bb.v[i];
	} else if (c->opt->upsampling == SSB_UPSAMPLE_JH) { /* color.cpp:202-232 */
		float coeffs[3];
		jh_fetch(col, lrgb, coeffs);
		for (int i = 0; i < c->nw; ++i) result.v[i] = jh_eval_precise(coeffs, lambda_0 + (float)i * c->lambda_step);
	} else { /* MENG, color.cpp:174-201: xyz_rel = transpose(M) * 100 * lrgb */
		static const float M[9] = { 0.41231515f, 0.3576f, 0.1805f, 0.2126f, 0.7152f, 0.0722f, 0.01932727f, 0.1192f, 0.95063333f };
		/* glm::transpose(mat3(9 scalars column-major)) * 100.0f: element (row r, col k) = M[r*3+k]*100 */
		float xyz_rel[3];
		for (int r = 0; r < 3; ++r)
			xyz_rel[r] = ((M[r * 3 + 0] * 100.0f) * lrgb[0] + (M[r * 3 + 1] * 100.0f) * lrgb[1]) + (M[r * 3 + 2] * 100.0f) * lrgb[2];
		for (int i = 0; i < c->nw; ++i) result.v[i] = meng_xyz_to_p(col->meng, lambda_0 + (float)i * c->lambda_step, xyz_rel);
	}
	return result;
}

/* sRGB_ReflectanceTexture::sample, material.cpp:45-97 */
static hero texture_sample(const octx* c, const ssb_texture* tex, v2 st, float lambda_0) {
	float uvx = st.x * (float)tex->width, uvy = st.y * (float)tex->height;
	float index_x = uvx, index_y = (float)tex->height - uvy;
	int i = (int)floorf(index_x), j = (int)floorf(index_y);
	int wi = (int)tex->width - 1, hj = (int)tex->height - 1;
	i = (i < 0) ? 0 : i; i = (wi < i) ? wi : i;  /* glm::clamp = min(max(x,lo),hi) */
	j = (j < 0) ? 0 : j; j = (hj < j) ? hj : j;
	const uint8_t* px = tex->rgb8 + 3 * ((size_t)j * tex->width + (size_t)i);
	float srgb[3] = { (float)px[0] * (1.0f / 255.0f), (float)px[1] * (1.0f / 255.0f), (float)px[2] * (1.0f / 255.0f) };
	float lrgb[3] = { srgb_to_lrgb_1(srgb[0]), srgb_to_lrgb_1(srgb[1]), srgb_to_lrgb_1(srgb[2]) };
	if (c->counters) c->counters->texture_lookups++;
	if (c->opt->render_mode == SSB_RENDER_RGB) { hero h = { { lrgb[0], lrgb[1], lrgb[2], 0.0f } }; return h; } /* material.cpp:64-66 */
	return lrgb_to_specrefl(c, lrgb, lambda_0);
}
static hero material_albedo(const octx* c, const ssb_material* m, v2 st, float lambda_0) {
	if (c->opt->render_mode == SSB_RENDER_RGB) { /* the `#else` branches of material.cpp:120-167: RGB triple or texel l-RGB */
		if (m->albedo_mode == SSB_ALBEDO_CONSTANT) { hero h = { { m->albedo_rgb[0], m->albedo_rgb[1], m->albedo_rgb[2], 0.0f } }; return h; }
		return texture_sample(c, &c->scene->textures[m->texture], st, lambda_0);
	}
	if (m->albedo_mode == SSB_ALBEDO_CONSTANT) return spectrum_hero(c, &m->albedo, lambda_0);
	return texture_sample(c, &c->scene->textures[m->texture], st, lambda_0);
}
/* MaterialBase::evaluate_emission, material.hpp:96-104 */
static hero material_emission(const octx* c, const ssb_material* m, float lambda_0) {
	if (c->opt->render_mode == SSB_RENDER_RGB) { hero h = { { m->emission_rgb[0], m->emission_rgb[1], m->emission_rgb[2], 0.0f } }; return h; }
	return spectrum_hero(c, &m->emission, lambda_0);
}

/* Color::specradflux_to_ciexyz(HeroSample, lambda_0), color.hpp:115-139 */
static void specradflux_to_ciexyz(const octx* c, hero flux, float lambda_0, float xyz[3]) {
	const ssb_spectrum* obs[3] = { &c->color->xbar, &c->color->ybar, &c->color->zbar };
	for (int k = 0; k < 3; ++k) {
		hero v = hero_scale(hero_mul(spectrum_hero(c, obs[k], lambda_0), flux), c->lambda_step);
		float acc = 0.0f;
		for (int i = 0; i < c->nw; ++i) acc += v.v[i];
		xyz[k] = acc;
	}
}

/* ------------------------------------------------------------------ geometry (geometry.cpp) */
typedef struct {
	int quad;  /* -1: none (HitRecord::prim == nullptr) */
	v3 normal;
	v2 st;
	float dist;
	int tri;           /* which triangle of the quad, and its barycentrics: not in the reference's HitRecord (it keeps only */
	float bx, by, bz;  /* normal and st); recorded for ssb_oracle_intersect, the yardstick of the device's hit records */
} hitrec_t;
typedef struct { v3 orig, dir; } ray_t;

/* PrimTri::intersect, geometry.cpp:12-101 (Woop / Benthin / Wald watertight test) */
static int tri_intersect(const octx* c, const ssb_tri* tri, const ray_t* ray, hitrec_t* hitrec) {
	const float EPS = c->opt->eps;
	v3 abs_dir = v3_make(fabsf(ray->dir.x), fabsf(ray->dir.y), fabsf(ray->dir.z));
	int kx, ky, kz;
	if (abs_dir.x > abs_dir.y) {
		if (abs_dir.x > abs_dir.z) { kz = 0; kx = 1; ky = 2; } else { kz = 2; kx = 0; ky = 1; }
	} else {
		if (abs_dir.y > abs_dir.z) { kz = 1; kx = 2; ky = 0; } else { kz = 2; kx = 0; ky = 1; }
	}
	if (v3_get(ray->dir, kz) < 0) { int t = kx; kx = ky; ky = t; }
	float Sx = v3_get(ray->dir, kx) / v3_get(ray->dir, kz);
	float Sy = v3_get(ray->dir, ky) / v3_get(ray->dir, kz);
	float Sz = 1.0f / v3_get(ray->dir, kz);
	v3 A = v3_sub(v3_make(tri->v[0].pos[0], tri->v[0].pos[1], tri->v[0].pos[2]), ray->orig);
	v3 B = v3_sub(v3_make(tri->v[1].pos[0], tri->v[1].pos[1], tri->v[1].pos[2]), ray->orig);
	v3 C = v3_sub(v3_make(tri->v[2].pos[0], tri->v[2].pos[1], tri->v[2].pos[2]), ray->orig);
	v3 ABC_kx = v3_make(v3_get(A, kx), v3_get(B, kx), v3_get(C, kx));
	v3 ABC_ky = v3_make(v3_get(A, ky), v3_get(B, ky), v3_get(C, ky));
	v3 ABC_kz = v3_make(v3_get(A, kz), v3_get(B, kz), v3_get(C, kz));
	v3 ABCx = v3_sub(ABC_kx, v3_scale(ABC_kz, Sx));
	v3 ABCy = v3_sub(ABC_ky, v3_scale(ABC_kz, Sy));
	/* UVW = cross(ABCy, ABCx) with glm::cross(x,y) = (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y) */
	float U = ABCy.y * ABCx.z - ABCx.y * ABCy.z;
	float V = ABCy.z * ABCx.x - ABCx.z * ABCy.x;
	float W = ABCy.x * ABCx.y - ABCx.x * ABCy.y;
	if (U != 0.0f && V != 0.0f && W != 0.0f) {
		if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return 0;
	} else {
		double Ud = (double)ABCy.y * (double)ABCx.z - (double)ABCx.y * (double)ABCy.z;
		double Vd = (double)ABCy.z * (double)ABCx.x - (double)ABCx.z * (double)ABCy.x;
		double Wd = (double)ABCy.x * (double)ABCx.y - (double)ABCx.x * (double)ABCy.y;
		if ((Ud < 0.0 || Vd < 0.0 || Wd < 0.0) && (Ud > 0.0 || Vd > 0.0 || Wd > 0.0)) return 0;
		U = (float)Ud; V = (float)Vd; W = (float)Wd;
		if (c->counters) c->counters->double_fallbacks++;
	}
	float det = U + V + W;
	if (fabsf(det) > EPS) { } else return 0;
	v3 ABCz = v3_scale(ABC_kz, Sz);
	float T = U * ABCz.x + V * ABCz.y + W * ABCz.z;
	uint32_t det_u, T_u;
	memcpy(&det_u, &det, 4); memcpy(&T_u, &T, 4);
	if (((det_u & 0x80000000u) ^ (T_u & 0x80000000u)) > 0) return 0;
	float det_recip = 1 / det;
	float dist = T * det_recip;
	if (dist >= EPS && dist < hitrec->dist) {
		float bx = U * det_recip, by = V * det_recip, bz = W * det_recip;
		hitrec->normal = v3_make(tri->normal[0], tri->normal[1], tri->normal[2]);
		hitrec->st.x = bx * tri->v[0].st[0] + by * tri->v[1].st[0] + bz * tri->v[2].st[0];
		hitrec->st.y = bx * tri->v[0].st[1] + by * tri->v[1].st[1] + bz * tri->v[2].st[1];
		hitrec->dist = dist;
		hitrec->bx = bx; hitrec->by = by; hitrec->bz = bz;
		return 1;
	}
	return 0;
}

/* Scene::intersect, scene.cpp:433-445 + PrimQuad::intersect, geometry.cpp:128-139 */
static int scene_intersect(const octx* c, const ray_t* ray, hitrec_t* hitrec, int ignore) {
	hitrec->quad = -1;
	hitrec->dist = INFINITY;
	int hit = 0;
	for (uint32_t q = 0; q < c->scene->nquads; ++q) {
		if ((int)q == ignore) continue;
		const ssb_quad* quad = &c->scene->quads[q];
		if (c->counters) c->counters->tri_tests++;
		int h = tri_intersect(c, &quad->tri[0], ray, hitrec), t = 0;
		if (!h) { if (c->counters) c->counters->tri_tests++; h = tri_intersect(c, &quad->tri[1], ray, hitrec); t = 1; }
		if (h) { hitrec->quad = (int)q; hitrec->tri = t; hit = 1; }
	}
	return hit;
}

/* ------------------------------------------------------------------ spherical triangle (util/spherical-tri.cpp:18-124) */
typedef struct {
	v3 A, B, C;
	float a, b, c, sin_a, sin_b, sin_c, cos_a, cos_b, cos_c;
	float alpha, beta, gamma, cos_alpha, cos_beta, cos_gamma;
	float surface_area;
} sphtri_t;

static inline float underestimate_pi(void) { uint32_t u = 0x40490FDAu; float r; memcpy(&r, &u, 4); return r; }

static void sphtri_init(sphtri_t* t, v3 A, v3 B, v3 C) {
	const float PI_F = 3.14159265358979323846f;
	t->A = A; t->B = B; t->C = C;
	t->cos_a = glm_clamp(v3_dot(B, C), -1.0f, 1.0f);
	t->cos_b = glm_clamp(v3_dot(A, C), -1.0f, 1.0f);
	t->cos_c = glm_clamp(v3_dot(A, B), -1.0f, 1.0f);
	t->a = acosf(t->cos_a); t->b = acosf(t->cos_b); t->c = acosf(t->cos_c);
	t->a = glm_clamp(t->a, 0.0f, underestimate_pi());
	t->b = glm_clamp(t->b, 0.0f, underestimate_pi());
	t->c = glm_clamp(t->c, 0.0f, underestimate_pi());
	t->sin_a = sinf(t->a); t->sin_b = sinf(t->b); t->sin_c = sinf(t->c);
	float numer0 = t->cos_a - t->cos_b * t->cos_c;
	float numer1 = t->cos_b - t->cos_c * t->cos_a;
	float numer2 = t->cos_c - t->cos_a * t->cos_b;
	float denom0 = t->sin_b * t->sin_c, denom1 = t->sin_c * t->sin_a, denom2 = t->sin_a * t->sin_b;
	if (denom0 > 0 && denom1 > 0 && denom2 > 0) {
		t->cos_alpha = glm_clamp(numer0 / denom0, -1.0f, 1.0f);
		t->cos_beta = glm_clamp(numer1 / denom1, -1.0f, 1.0f);
		t->cos_gamma = glm_clamp(numer2 / denom2, -1.0f, 1.0f);
		t->alpha = glm_clamp(acosf(t->cos_alpha), 0.0f, underestimate_pi());
		t->beta = glm_clamp(acosf(t->cos_beta), 0.0f, underestimate_pi());
		t->gamma = glm_clamp(acosf(t->cos_gamma), 0.0f, underestimate_pi());
		t->surface_area = t->alpha + t->beta + t->gamma - PI_F;
		if (t->surface_area >= 0) { } else t->surface_area = 0;
	} else {
		t->surface_area = 0;
		int degenerate = 0;
		if (t->sin_a > 0) {
			if (t->sin_b > 0) {
				if (t->sin_c > 0) degenerate = 1;
				else {
					t->cos_alpha = t->cos_beta = 1; t->alpha = t->beta = PI_F * 0.5f;
					t->cos_gamma = glm_clamp(numer2 / denom2, -1.0f, 1.0f); t->gamma = acosf(t->cos_gamma);
				}
			} else {
				if (t->sin_c > 0) {
					t->cos_alpha = t->cos_gamma = 1; t->alpha = t->gamma = PI_F * 0.5f;
					t->cos_beta = glm_clamp(numer1 / denom1, -1.0f, 1.0f); t->beta = acosf(t->cos_beta);
				} else degenerate = 1;
			}
		} else {
			if (t->sin_b > 0) {
				if (t->sin_c > 0) {
					t->cos_beta = t->cos_gamma = 1; t->beta = t->gamma = PI_F * 0.5f;
					t->cos_alpha = glm_clamp(numer0 / denom0, -1.0f, 1.0f); t->alpha = acosf(t->cos_alpha);
				} else degenerate = 1;
			} else degenerate = 1;
		}
		if (degenerate)
			t->cos_alpha = t->cos_beta = t->cos_gamma = t->alpha = t->beta = t->gamma = NAN;
	}
}

/* func_bar lambda, random.cpp:139-144 */
static v3 func_bar(v3 x, v3 y) {
	v3 dir = v3_sub(x, v3_scale(y, v3_dot(x, y)));
	float lensq = v3_dot(dir, dir);
	if (lensq == 0.0f) return v3_make(0, 0, 0);
	return v3_scale(dir, 1.0f / sqrtf(lensq));
}
/* Math::rand_toward_sphericaltri, random.cpp:101-154 (Arvo 1995) */
static v3 rand_toward_sphericaltri(rng_t* rng, const sphtri_t* tri) {
	float r0 = rand_1f(rng);
	float r1 = rand_1f(rng);
	float sin_alpha = sinf(tri->alpha);
	float q;
	if (sin_alpha > 0) {
		float random_area = r0 * tri->surface_area;
		float phi = random_area - tri->alpha;
		float s = sinf(phi), t = cosf(phi);
		float u = t - tri->cos_alpha;
		float v = s + sin_alpha * tri->cos_c;
		float denom = (v * s + u * t) * sin_alpha;
		if (denom != 0.0f) q = ((v * t - u * s) * tri->cos_alpha - v) / denom;
		else q = tri->cos_c;
	} else {
		q = (float)cos((double)(tri->b * r0)); /* unqualified cos(float) binds ::cos(double), random.cpp:135 */
	}
	q = glm_clamp(q, -1.0f, 1.0f);
	v3 C_hat = v3_add(v3_scale(tri->A, q), v3_scale(func_bar(tri->C, tri->A), sqrtf(1 - q * q)));
	float z = 1.0f - r1 * (1.0f - v3_dot(C_hat, tri->B));
	z = glm_clamp(z, -1.0f, 1.0f);
	return v3_add(v3_scale(tri->B, z), v3_scale(func_bar(C_hat, tri->B), sqrtf(1 - z * z)));
}

/* Scene::get_rand_toward_light (scene.cpp:417-431) -> PrimQuad::get_rand_toward (geometry.cpp:141-145)
 * -> PrimTri::get_rand_toward (geometry.cpp:103-116) */
static void get_rand_toward_light(const octx* c, rng_t* rng, v3 from, v3* dir, int* light, float* pdf) {
	uint32_t li = rand_choice(rng, c->nlights);
	*light = (int)c->lights[li];
	const ssb_quad* quad = &c->scene->quads[*light];
	const ssb_tri* tri = (rand_1f(rng) <= 0.5f) ? &quad->tri[0] : &quad->tri[1];
	sphtri_t st;
	sphtri_init(&st,
		v3_normalize(v3_sub(v3_make(tri->v[0].pos[0], tri->v[0].pos[1], tri->v[0].pos[2]), from)),
		v3_normalize(v3_sub(v3_make(tri->v[1].pos[0], tri->v[1].pos[1], tri->v[1].pos[2]), from)),
		v3_normalize(v3_sub(v3_make(tri->v[2].pos[0], tri->v[2].pos[1], tri->v[2].pos[2]), from)));
	*dir = rand_toward_sphericaltri(rng, &st);
	*pdf = 1.0f / st.surface_area;
	*pdf *= 0.5f;
	*pdf /= (float)c->nlights;
}

/* Math::rand_coshemi, random.cpp:29-49 */
static v3 rand_coshemi(const octx* c, rng_t* rng, float* pdf) {
	const float PI_F = 3.14159265358979323846f;
	v3 result;
	do {
		float angle = rand_1f(rng) * (2.0f * PI_F);
		float co = cosf(angle), si = sinf(angle);
		float radius_sq = rand_1f(rng);
		float radius = sqrtf(radius_sq);
		result = v3_make(radius * co, sqrtf(1 - radius_sq), radius * si);
		*pdf = result.y;
	} while (*pdf <= c->opt->eps);
	*pdf *= 1.0f / PI_F;
	return result;
}
/* Math::get_basis + get_rotated_to, math-helpers.hpp:14-38 (Duff et al. 2017) */
static v3 get_rotated_to(v3 dir, v3 n) {
	float sign = copysignf(1.0f, n.z);
	float a = -1.0f / (sign + n.z);
	float b = n.x * n.y * a;
	v3 bx = v3_make(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
	v3 bz = v3_make(b, sign + n.y * n.y * a, -n.y);
	return v3_add(v3_add(v3_scale(bx, dir.x), v3_scale(n, dir.y)), v3_scale(bz, dir.z));
}
/* Math::reflect, math-helpers.hpp:40-42 */
static v3 reflect(v3 vec, v3 n) { return v3_add(v3_neg(vec), v3_scale(n, 2.0f * v3_dot(vec, n))); }

/* ------------------------------------------------------------------ integrator (renderer.cpp:103-277) */
typedef struct {
	const octx* c;
	rng_t* rng;
	float lambda_0;
	int hit_anything;
} path_t;

/* the lambda `L`, renderer.cpp:147-255 */
static hero L(path_t* p, const ray_t* ray, int last_was_delta, unsigned depth, int ignore) {
	const octx* c = p->c;
	const ssb_options* opt = c->opt;
	const float PI_F = 3.14159265358979323846f;
	hero radiance = hero_splat(0.0f);
	hitrec_t hitrec;
	if (c->counters) c->counters->closest_queries++;
	if (scene_intersect(c, ray, &hitrec, ignore)) {
		p->hit_anything = 1;
		const ssb_quad* quad = &c->scene->quads[hitrec.quad];
		const ssb_material* mat = &c->scene->materials[quad->material];
		int els = opt->explicit_light_sampling != 0;
		if (!els || (last_was_delta && (!opt->indirect_only || depth > 0u))) { /* renderer.cpp:167-175 */
			radiance = hero_add(radiance, material_emission(c, mat, p->lambda_0));
		}
		if (depth + 1u < opt->max_depth) {
			v3 hit_pos = v3_add(ray->orig, v3_scale(ray->dir, hitrec.dist)); /* Ray::at, stdafx.hpp:219 */
			if (els && (!opt->indirect_only || depth > 0u)) { /* renderer.cpp:182-220 */
				v3 shad_dir; int light; float shad_pdf;
				get_rand_toward_light(c, p->rng, hit_pos, &shad_dir, &light, &shad_pdf);
				float n_dot_l = v3_dot(shad_dir, hitrec.normal);
				if (n_dot_l > 0.0f) {
					ray_t ray_shad = { hit_pos, shad_dir };
					hitrec_t hs;
					if (c->counters) c->counters->shadow_queries++;
					scene_intersect(c, &ray_shad, &hs, hitrec.quad);
					if (hs.quad == light) {
						const ssb_material* lm = &c->scene->materials[c->scene->quads[light].material];
						hero emitted = material_emission(c, lm, p->lambda_0);
						hero f_s; /* evaluate_bsdf, material.cpp:120-129 / 146-153 */
						if (mat->kind == SSB_MATERIAL_LAMBERT) f_s = hero_div(material_albedo(c, mat, hitrec.st, p->lambda_0), PI_F);
						else f_s = hero_splat(0.0f);
						radiance = hero_add(radiance, hero_div(hero_mul(hero_scale(emitted, n_dot_l), f_s), shad_pdf));
						if (c->counters) c->counters->unshadowed++;
					}
				}
			}
			/* interact_bsdf, material.cpp:130-143 / 154-167 */
			v3 w_i; float pdf_w_i; hero f_s;
			if (mat->kind == SSB_MATERIAL_LAMBERT) {
				w_i = rand_coshemi(c, p->rng, &pdf_w_i);
				w_i = get_rotated_to(w_i, hitrec.normal);
				f_s = hero_div(material_albedo(c, mat, hitrec.st, p->lambda_0), PI_F);
			} else {
				w_i = reflect(v3_neg(ray->dir), hitrec.normal);
				pdf_w_i = INFINITY;
				f_s = material_albedo(c, mat, hitrec.st, p->lambda_0);
			}
			if (c->counters) c->counters->bsdf_samples++;
			if (hero_dot(f_s, f_s) > 0.0f) { /* renderer.cpp:231 */
				float n_dot_l;
				if (isfinite(pdf_w_i)) n_dot_l = v3_dot(w_i, hitrec.normal);
				else { n_dot_l = 1.0f; pdf_w_i = 1.0f; }
				if (n_dot_l > 0.0f) {
					ray_t ray_next = { hit_pos, w_i };
					hero child = L(p, &ray_next, 0, depth + 1u, hitrec.quad); /* `false`, renderer.cpp:248 */
					radiance = hero_add(radiance, hero_div(hero_mul(hero_scale(child, n_dot_l), f_s), pdf_w_i));
				}
			}
		}
	}
	return radiance;
}

/* Renderer::_render_sample, renderer.cpp:103-277 */
static void render_sample(const octx* c, rng_t* rng, uint32_t i, uint32_t j, float out[4]) {
	const ssb_options* opt = c->opt;
	const ssb_camera* cam = &c->scene->camera;
	/* glm::dvec2 subpixel(rand_1d(rng),rand_1d(rng)): g++ evaluates the arguments right-to-left,
	 * so .y receives the first pair of draws (renderer.cpp:113; verified against oracle/_ref) */
	double sub_y = rand_1d(rng);
	double sub_x = rand_1d(rng);
	double st_x = ((double)i + sub_x) / (double)opt->width;
	double st_y = ((double)j + sub_y) / (double)opt->height;
	double ndc_x = st_x * 2.0 - 1.0, ndc_y = st_y * 2.0 - 1.0;
	/* point = PV_inv * dvec4(ndc,0,1): (m0*v0 + m1*v1) + (m2*v2 + m3*v3), renderer.cpp:129 */
	double point[4];
	for (int r = 0; r < 4; ++r)
		point[r] = (cam->pv_inv[0 + r] * ndc_x + cam->pv_inv[4 + r] * ndc_y) + (cam->pv_inv[8 + r] * 0.0 + cam->pv_inv[12 + r] * 1.0);
	double w = point[3];
	for (int r = 0; r < 4; ++r) point[r] /= w; /* renderer.cpp:130 */
	double dx = point[0] - (double)cam->pos[0], dy = point[1] - (double)cam->pos[1], dz = point[2] - (double)cam->pos[2];
	double inv = 1.0 / sqrt((dx * dx + dy * dy) + dz * dz); /* glm::normalize(dvec3) */
	ray_t ray_camera;
	ray_camera.orig = v3_make(cam->pos[0], cam->pos[1], cam->pos[2]);
	ray_camera.dir = v3_make((float)(dx * inv), (float)(dy * inv), (float)(dz * inv));

	/* renderer.cpp:134-143: the hero wavelength is drawn in spectral mode only */
	const int rgb = opt->render_mode == SSB_RENDER_RGB;
	float lambda_0 = rgb ? 0.0f : opt->lambda_min + rand_1f(rng) * c->lambda_step;

	path_t p = { c, rng, lambda_0, 0 };
	hero pixel_rad_est = L(&p, &ray_camera, 1, 0u, -1);
	hero pixel_flux_est = pixel_rad_est;
	if (!opt->flat_field_correction) /* renderer.cpp:262-266 */
		pixel_flux_est = hero_scale(pixel_rad_est, v3_dot(ray_camera.dir, v3_make(cam->dir[0], cam->dir[1], cam->dir[2])));
	if (rgb) { out[0] = pixel_flux_est.v[0]; out[1] = pixel_flux_est.v[1]; out[2] = pixel_flux_est.v[2]; } /* renderer.cpp:274-275 */
	else specradflux_to_ciexyz(c, pixel_flux_est, lambda_0, out);
	out[3] = p.hit_anything ? 1.0f : 0.0f;
	if (c->counters) c->counters->samples++;
}

static int prepare(octx* c, const ssb_scene* scene, const ssb_color* color, const ssb_options* opt) {
	memset(c, 0, sizeof(*c));
	if (!scene || !opt) return SSB_ERR_ARG;
	if (opt->render_mode > SSB_RENDER_RGB) return SSB_ERR_UNSUPPORTED;
	if (!color && opt->render_mode != SSB_RENDER_RGB) return SSB_ERR_ARG; /* RGB mode needs no colour tables */
	if (opt->width == 0 || opt->height == 0 || opt->spp == 0) return SSB_ERR_ARG;
	if (opt->render_mode != SSB_RENDER_RGB && (opt->upsampling < SSB_UPSAMPLE_OURS || opt->upsampling > SSB_UPSAMPLE_JH)) return SSB_ERR_UNSUPPORTED;
	c->scene = scene; c->color = color; c->opt = opt;
	c->nw = opt->n_wavelengths ? (int)opt->n_wavelengths : 4;
	if (c->nw < 2 || c->nw > 4) return SSB_ERR_UNSUPPORTED; /* glm::vec<N,float>: the reference compiles for N = 2, 3, 4 */
	c->lambda_step = (opt->lambda_max - opt->lambda_min) / (float)c->nw; /* stdafx.hpp:289 */
	for (uint32_t q = 0; q < scene->nquads; ++q)
		if (scene->quads[q].is_light) { if (c->nlights >= SSB_MAX_LIGHTS) return SSB_ERR_UNSUPPORTED; c->lights[c->nlights++] = q; }
	if (opt->explicit_light_sampling && c->nlights == 0) return SSB_ERR_ARG; /* assert(!lights.empty()), scene.cpp:30 */
	return SSB_OK;
}

/* Scene::intersect (scene.cpp:433-445) for n caller-supplied rays: rays6 = n x {origin, direction}, ignore = n quad
 * indices or -1 (or NULL), out6 = n x {quad (int bits, -1: miss), tri (int bits), dist, barycentrics U,V,W * det_recip}.
 * The yardstick of tests/test_gpu_isect_fuzz.py for the device's filtered scan. */
int ssb_oracle_intersect(const ssb_scene* scene, const float* rays6, const int32_t* ignore, float eps, float* out6, size_t n) {
	if (!scene || !rays6 || !out6) return SSB_ERR_ARG;
	ssb_options opt;
	memset(&opt, 0, sizeof(opt));
	opt.eps = eps;
	octx c;
	memset(&c, 0, sizeof(c));
	c.scene = scene; c.opt = &opt;
	#pragma omp parallel for schedule(static)
	for (long long r = 0; r < (long long)n; ++r) {
		ray_t ray;
		ray.orig = v3_make(rays6[6 * r], rays6[6 * r + 1], rays6[6 * r + 2]);
		ray.dir = v3_make(rays6[6 * r + 3], rays6[6 * r + 4], rays6[6 * r + 5]);
		hitrec_t h;
		memset(&h, 0, sizeof(h));
		scene_intersect(&c, &ray, &h, ignore ? ignore[r] : -1);
		int32_t q = h.quad, t = h.quad >= 0 ? h.tri : 0;
		memcpy(out6 + 6 * r, &q, 4); memcpy(out6 + 6 * r + 1, &t, 4);
		out6[6 * r + 2] = h.dist;
		out6[6 * r + 3] = h.quad >= 0 ? h.bx : 0.0f; out6[6 * r + 4] = h.quad >= 0 ? h.by : 0.0f; out6[6 * r + 5] = h.quad >= 0 ? h.bz : 0.0f;
	}
	return SSB_OK;
}

/* Renderer::_render_pixel's sample loop, renderer.cpp:292-295: accum += double4(sample * 0.001f) */
int ssb_oracle_render(const ssb_scene* scene, const ssb_color* color, const ssb_options* opt,
                      double* accum, float* samples_out, ssb_oracle_counters* counters) {
	octx base;
	int rc = prepare(&base, scene, color, opt);
	if (rc != SSB_OK) return rc;
	uint32_t x1 = opt->x1 ? opt->x1 : opt->width, y1 = opt->y1 ? opt->y1 : opt->height;
	uint32_t s1 = opt->sample_end ? opt->sample_end : opt->spp;
	if (x1 > opt->width || y1 > opt->height || opt->x0 > x1 || opt->y0 > y1 || opt->sample_begin > s1) return SSB_ERR_ARG;
	uint32_t ns = s1 - opt->sample_begin;
	ssb_oracle_counters total;
	memset(&total, 0, sizeof(total));
	#pragma omp parallel
	{
		ssb_oracle_counters local;
		memset(&local, 0, sizeof(local));
		octx c = base;
		c.counters = counters ? &local : NULL;
		#pragma omp for schedule(dynamic, 1)
		for (uint32_t j = opt->y0; j < y1; ++j) {
			/* ssb_options.band_*: only the rows of this share of the interleaved row bands (include/ssb200.h) */
			if (opt->band_count > 1 && (opt->band_height == 0 || (j / opt->band_height) % opt->band_count != opt->band_index)) continue;
			for (uint32_t i = opt->x0; i < x1; ++i) {
				size_t pixel = (size_t)j * opt->width + i;
				double* avg = accum ? accum + 4 * pixel : NULL;
				for (uint32_t k = opt->sample_begin; k < s1; ++k) {
					rng_t rng;
					uint64_t index = (uint64_t)k * ((uint64_t)opt->width * opt->height) + (uint64_t)pixel;
					seed_sample(&rng, opt->seed, index);
					float s[4];
					render_sample(&c, &rng, i, j, s);
					if (samples_out) memcpy(samples_out + 4 * (pixel * ns + (k - opt->sample_begin)), s, sizeof(s));
					if (avg) {
						if (opt->render_mode == SSB_RENDER_RGB) for (int ch = 0; ch < 4; ++ch) avg[ch] += (double)s[ch]; /* renderer.cpp:301-303 */
						else for (int ch = 0; ch < 4; ++ch) avg[ch] += (double)(s[ch] * 0.001f);
					}
				}
			}
		}
		if (counters) {
			#pragma omp critical
			{
				uint64_t* t = (uint64_t*)&total; const uint64_t* l = (const uint64_t*)&local;
				for (size_t n = 0; n < sizeof(total) / sizeof(uint64_t); ++n) t[n] += l[n];
			}
		}
	}
	if (counters) *counters = total;
	return SSB_OK;
}

/* renderer.cpp:296-298 + Color::ciexyz_to_srgb, color.cpp:237-257 */
int ssb_oracle_resolve(const ssb_color* color, const ssb_options* opt, const double* accum, double* xyza, float* srgba) {
	if (!opt || !accum) return SSB_ERR_ARG;
	const int rgb = opt->render_mode == SSB_RENDER_RGB;
	if (!color && !rgb) return SSB_ERR_ARG;
	size_t npix = (size_t)opt->width * opt->height;
	double scale = 1000.0 / (double)opt->spp;
	for (size_t p = 0; p < npix; ++p) {
		double avg[4];
		if (rgb) for (int ch = 0; ch < 4; ++ch) avg[ch] = accum[4 * p + ch] / (double)opt->spp; /* renderer.cpp:304 */
		else for (int ch = 0; ch < 4; ++ch) avg[ch] = accum[4 * p + ch] * scale;
		if (xyza) memcpy(xyza + 4 * p, avg, sizeof(avg));
		if (srgba && rgb) { /* Color::lrgb_to_srgb(lRGB_F32(avg)), renderer.cpp:306 */
			for (int ch = 0; ch < 3; ++ch) srgba[4 * p + ch] = lrgb_to_srgb_1((float)avg[ch]);
			srgba[4 * p + 3] = (float)avg[3];
		} else if (srgba) {
			float xyz[3] = { (float)avg[0], (float)avg[1], (float)avg[2] };
			float lrgb[3];
			if (opt->upsampling == SSB_UPSAMPLE_MENG) { /* color.cpp:243-254 */
				static const float Mi[9] = { 3.24156456f, -1.53766524f, -0.49870224f, -0.96920119f, 1.87588535f, 0.04155324f, 0.05562416f, -0.20395525f, 1.05685902f };
				float rel[3] = { xyz[0] / color->d65_rad_Y, xyz[1] / color->d65_rad_Y, xyz[2] / color->d65_rad_Y };
				for (int r = 0; r < 3; ++r) lrgb[r] = (Mi[r * 3 + 0] * rel[0] + Mi[r * 3 + 1] * rel[1]) + Mi[r * 3 + 2] * rel[2];
			} else { /* matr_xyz_to_lrgb * xyz, column-major */
				const float* m = color->xyz_to_lrgb;
				for (int r = 0; r < 3; ++r) lrgb[r] = (m[0 + r] * xyz[0] + m[3 + r] * xyz[1]) + m[6 + r] * xyz[2];
			}
			for (int ch = 0; ch < 3; ++ch) srgba[4 * p + ch] = lrgb_to_srgb_1(lrgb[ch]);
			srgba[4 * p + 3] = (float)avg[3];
		}
	}
	return SSB_OK;
}

/* libm probes used by the device-maths bit-exactness tests */
void ssb_oracle_eval_math(uint32_t fn, const float* x, float arg, float* out, size_t n) {
	for (size_t i = 0; i < n; ++i) {
		switch (fn) {
			case 0: out[i] = sinf(x[i]); break;
			case 1: out[i] = cosf(x[i]); break;
			case 2: out[i] = acosf(x[i]); break;
			default: out[i] = powf(x[i], arg); break;
		}
	}
}
