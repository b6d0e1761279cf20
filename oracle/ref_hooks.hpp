// TEST INFRASTRUCTURE ONLY — not part of the product.
//
// Hooks that oracle/build_ref.py splices into a *temporary copy* of the reference's
// renderer.cpp when it builds the "hooked" oracle binaries under oracle/_ref/.
// They change nothing unless an SSB_* environment variable is set, in which case they
//   SSB_SEED=<u64>          re-seed the reference's PCG32 before every sample from
//                           (seed, sample index) — the matched-seed scheme the CUDA
//                           path uses (the pristine reference has no reproducible
//                           seeding: one RNG stream per worker thread over dynamically
//                           scheduled tiles, renderer.cpp:323-379);
//   SSB_DUMP_XYZA=<path>    write the per-pixel double XYZA accumulator
//                           (renderer.cpp:292-296, never stored by the reference);
//   SSB_DUMP_SAMPLES=<path> write every sample's float4 (X,Y,Z,hit);
//   SSB_DUMP_TABLES=<path>  write Color::data, the camera and the flattened scene;
//   SSB_THREADS=<n>         override the worker-thread count;
//   SSB_SCENE_QUADS=<path>  replace the scene's quads by the ones in a file (random-scene tests).
// This file is ours (no reference code); it is #included after the reference's own
// headers, so it sees Scene / PrimQuad / MaterialLambertian / Color::data.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace ssb_hooks {

inline uint64_t mix64(uint64_t z) {
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

struct State {
	bool seeded = false;
	uint64_t seed = 0;
	const char* dump_xyza = nullptr;
	const char* dump_samples = nullptr;
	const char* dump_tables = nullptr;
	std::vector<double> xyza;
	std::vector<float> samples;
	size_t w = 0, h = 0, spp = 0;
	State() {
		if (const char* s = std::getenv("SSB_SEED")) { seeded = true; seed = std::strtoull(s, nullptr, 10); }
		dump_xyza = std::getenv("SSB_DUMP_XYZA");
		dump_samples = std::getenv("SSB_DUMP_SAMPLES");
		dump_tables = std::getenv("SSB_DUMP_TABLES");
	}
};
inline State& st() { static State s; return s; }

inline size_t threads(size_t hw) {
	if (const char* s = std::getenv("SSB_THREADS")) { long n = std::atol(s); if (n > 0) return static_cast<size_t>(n); }
	return hw;
}

inline void begin(size_t w, size_t h, size_t spp) {
	State& s = st();
	s.w = w; s.h = h; s.spp = spp;
	if (s.dump_xyza) s.xyza.assign(w * h * 4, 0.0);
	if (s.dump_samples) s.samples.assign(w * h * spp * 4, 0.0f);
}

// sample index = k*(W*H) + (j*W + i): the first n samples of a pixel do not depend on spp.
template <class RNG> inline void seed_sample(RNG& rng, size_t i, size_t j, size_t k) {
	State& s = st();
	if (!s.seeded) return;
	uint64_t index = static_cast<uint64_t>(k) * (static_cast<uint64_t>(s.w) * s.h) + (static_cast<uint64_t>(j) * s.w + i);
	uint64_t state = mix64(s.seed ^ mix64(index));
	uint64_t inc = mix64(state) | 1ull;
	rng.seed(state, inc);
}
template <class V4> inline void record_sample(size_t i, size_t j, size_t k, V4 const& v) {
	State& s = st();
	if (!s.dump_samples) return;
	float* dst = &s.samples[((j * s.w + i) * s.spp + k) * 4];
	dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2]; dst[3] = v[3];
}
template <class DV4> inline void record_pixel(size_t i, size_t j, DV4 const& avg) {
	State& s = st();
	if (!s.dump_xyza) return;
	double* dst = &s.xyza[(j * s.w + i) * 4];
	dst[0] = avg[0]; dst[1] = avg[1]; dst[2] = avg[2]; dst[3] = avg[3];
}
inline void finish() {
	State& s = st();
	if (s.dump_xyza) {
		FILE* f = std::fopen(s.dump_xyza, "wb");
		if (f) { std::fwrite(s.xyza.data(), sizeof(double), s.xyza.size(), f); std::fclose(f); }
	}
	if (s.dump_samples) {
		FILE* f = std::fopen(s.dump_samples, "wb");
		if (f) { std::fwrite(s.samples.data(), sizeof(float), s.samples.size(), f); std::fclose(f); }
	}
}

// ---- table dump: records "REC <name> <f32|f64|u32|u8> <count>\n" + raw little-endian payload
inline void rec(FILE* f, std::string const& name, const char* ty, size_t n, const void* p, size_t elsize) {
	std::fprintf(f, "REC %s %s %zu\n", name.c_str(), ty, n);
	if (n) std::fwrite(p, elsize, n, f);
}
inline void rec_f32(FILE* f, std::string const& n, const float* p, size_t c) { rec(f, n, "f32", c, p, 4); }
inline void rec_f64(FILE* f, std::string const& n, const double* p, size_t c) { rec(f, n, "f64", c, p, 8); }
inline void rec_u32(FILE* f, std::string const& n, const uint32_t* p, size_t c) { rec(f, n, "u32", c, p, 4); }

template <class M3> inline void rec_mat3(FILE* f, std::string const& name, M3 const& m) {
	float v[9]; for (size_t c = 0; c < 3; ++c) for (size_t r = 0; r < 3; ++r) v[c * 3 + r] = m[c][r];
	rec_f32(f, name, v, 9);
}
template <class M4> inline void rec_dmat4(FILE* f, std::string const& name, M4 const& m) {
	double v[16]; for (size_t c = 0; c < 4; ++c) for (size_t r = 0; r < 4; ++r) v[c * 4 + r] = m[c][r];
	rec_f64(f, name, v, 16);
}
#ifdef RENDER_MODE_SPECTRAL
inline void rec_spectrum(FILE* f, std::string const& name, _Spectrum const& s) {
	rec_f32(f, name + ".data", s._data.data(), s._data.size());
	float lh[2] = { s._low, s._high };
	rec_f32(f, name + ".lowhigh", lh, 2);
}
#endif

// SSB_SCENE_QUADS=<path>: replace the primitive list of the scene just built by quads read from a file — the way the
// tests put RANDOM geometry (arbitrary lights, slivers, non-planar and degenerate quads) through the real reference.
// File: u32 n, then per quad u32 material (index into the ORIGINAL scene's materials in first-use order over its primitive
// list, the order dump_tables() writes them in) and 4 x (pos.xyz, st.xy) f32 for v00, v10, v11, v01.  The quads are built by
// the reference's own PrimQuad constructor (normals, is_light from the material), the light list as Scene::_init() does.
inline void replace_scene(Scene* scene) {
	const char* path = std::getenv("SSB_SCENE_QUADS");
	if (!path) return;
	FILE* f = std::fopen(path, "rb");
	if (!f) { std::fprintf(stderr, "SSB_SCENE_QUADS: cannot open %s\n", path); std::exit(3); }
	std::vector<MaterialBase*> mats;
	for (PrimBase* prim : scene->primitives) {
		size_t m = 0;
		for (; m < mats.size(); ++m) if (mats[m] == prim->material) break;
		if (m == mats.size()) mats.push_back(prim->material);
	}
	uint32_t n = 0;
	if (std::fread(&n, 4, 1, f) != 1) std::exit(3);
	std::vector<PrimBase*> prims;
	for (uint32_t q = 0; q < n; ++q) {
		uint32_t m = 0; float v[20];
		if (std::fread(&m, 4, 1, f) != 1 || std::fread(v, 4, 20, f) != 20 || m >= mats.size()) { std::fprintf(stderr, "SSB_SCENE_QUADS: bad record %u\n", q); std::exit(3); }
		Vertex vert[4];
		for (size_t k = 0; k < 4; ++k) { vert[k].pos = Pos(v[5 * k], v[5 * k + 1], v[5 * k + 2]); vert[k].st = ST(v[5 * k + 3], v[5 * k + 4]); }
		prims.push_back(new PrimQuad(mats[m], vert[0], vert[1], vert[2], vert[3]));
	}
	std::fclose(f);
	scene->primitives = prims;  // (the old primitives leak: test binary)
	scene->lights.clear();
	for (PrimBase* prim : scene->primitives) if (prim->is_light) scene->lights.emplace_back(prim);
	if (scene->lights.empty()) { std::fprintf(stderr, "SSB_SCENE_QUADS: no light in the scene\n"); std::exit(3); }
}

inline void dump_tables(Scene* scene) {
	State& s = st();
	if (!s.dump_tables) return;
	FILE* f = std::fopen(s.dump_tables, "wb");
	if (!f) return;
#ifdef RENDER_MODE_SPECTRAL
	// colour data (util/color.hpp:22-69)
	rec_spectrum(f, "color.xbar", Color::data->std_obs_xbar);
	rec_spectrum(f, "color.ybar", Color::data->std_obs_ybar);
	rec_spectrum(f, "color.zbar", Color::data->std_obs_zbar);
	rec_spectrum(f, "color.D65_orig", Color::data->D65_orig);
	rec_spectrum(f, "color.D65_rad", Color::data->D65_rad);
	rec_f32(f, "color.D65_orig_XYZ", &Color::data->D65_orig_XYZ[0], 3);
	rec_f32(f, "color.D65_rad_XYZ", &Color::data->D65_rad_XYZ[0], 3);
	#if defined RENDER_MODE_SPECTRAL_OURS
	rec_spectrum(f, "color.basis_r", Color::data->basis_bt709.r);
	rec_spectrum(f, "color.basis_g", Color::data->basis_bt709.g);
	rec_spectrum(f, "color.basis_b", Color::data->basis_bt709.b);
	#endif
	rec_mat3(f, "color.matr_lrgb_to_xyz", Color::data->matr_lrgb_to_xyz);
	rec_mat3(f, "color.matr_xyz_to_lrgb", Color::data->matr_xyz_to_lrgb);
	float lam[3] = { LAMBDA_MIN, LAMBDA_MAX, LAMBDA_STEP };
	rec_f32(f, "config.lambda_min_max_step", lam, 3);
#endif
	// camera (scene.hpp:16-33)
	rec_f32(f, "camera.pos", &scene->camera.pos[0], 3);
	rec_f32(f, "camera.dir", &scene->camera.dir[0], 3);
	rec_f32(f, "camera.up", &scene->camera.up[0], 3);
	rec_dmat4(f, "camera.matr_P", scene->camera.matr_P);
	rec_dmat4(f, "camera.matr_V", scene->camera.matr_V);
	rec_dmat4(f, "camera.matr_PV_inv", scene->camera.matr_PV_inv);
	// materials, in first-use order over the primitive list
	std::vector<MaterialBase const*> mats;
	std::vector<uint32_t> prim_mat, prim_light;
	std::vector<float> quads;  // per quad: 2 tris x (3 x (pos3, st2) + normal3) = 36 floats
	for (PrimBase const* prim : scene->primitives) {
		PrimQuad const* q = static_cast<PrimQuad const*>(prim);
		size_t m = 0;
		for (; m < mats.size(); ++m) if (mats[m] == q->material) break;
		if (m == mats.size()) mats.push_back(q->material);
		prim_mat.push_back(static_cast<uint32_t>(m));
		prim_light.push_back(q->is_light ? 1u : 0u);
		PrimTri const* tris[2] = { &q->tri0, &q->tri1 };
		for (PrimTri const* t : tris) {
			for (size_t v = 0; v < 3; ++v) {
				quads.push_back(t->verts[v].pos.x); quads.push_back(t->verts[v].pos.y); quads.push_back(t->verts[v].pos.z);
				quads.push_back(t->verts[v].st.x); quads.push_back(t->verts[v].st.y);
			}
			quads.push_back(t->normal.x); quads.push_back(t->normal.y); quads.push_back(t->normal.z);
		}
	}
	rec_f32(f, "scene.quads", quads.data(), quads.size());
	rec_u32(f, "scene.quad_material", prim_mat.data(), prim_mat.size());
	rec_u32(f, "scene.quad_is_light", prim_light.data(), prim_light.size());
	std::vector<uint32_t> lights;
	for (PrimBase const* l : scene->lights)
		for (size_t p = 0; p < scene->primitives.size(); ++p)
			if (scene->primitives[p] == l) lights.push_back(static_cast<uint32_t>(p));
	rec_u32(f, "scene.lights", lights.data(), lights.size());
	for (size_t m = 0; m < mats.size(); ++m) {
		std::string base = "material." + std::to_string(m);
		MaterialSimpleAlbedoBase const* mat = static_cast<MaterialSimpleAlbedoBase const*>(mats[m]);
		uint32_t kind[2] = { dynamic_cast<MaterialLambertian const*>(mats[m]) ? 0u : 1u, mat->mode == MaterialSimpleAlbedoBase::CONSTANT ? 0u : 1u };
		rec_u32(f, base + ".kind_mode", kind, 2);
#ifdef RENDER_MODE_SPECTRAL
		rec_spectrum(f, base + ".emission", mat->emission);
		if (mat->mode == MaterialSimpleAlbedoBase::CONSTANT) rec_spectrum(f, base + ".albedo", *mat->albedo.constant);
#else
		rec_f32(f, base + ".emission_rgb", &mat->emission[0], 3);
		if (mat->mode == MaterialSimpleAlbedoBase::CONSTANT) rec_f32(f, base + ".albedo_rgb", &mat->albedo.constant[0], 3);
#endif
		else {
			uint32_t res[2] = { static_cast<uint32_t>(mat->albedo.texture->res[0]), static_cast<uint32_t>(mat->albedo.texture->res[1]) };
			rec_u32(f, base + ".texture_res", res, 2);
		}
	}
	std::fclose(f);
}

}  // namespace ssb_hooks
