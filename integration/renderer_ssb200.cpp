// renderer_ssb200.cpp — the file a maintainer of geometrian/simple-spectral ADDS to the reference tree to render through
// libssb200.so.  It is compiled INTO THE REAL REFERENCE by oracle/build_ref.py (tool "ssb200": scratch
// copy of /root/reference/src + this file, Renderer::render_start / render_wait of renderer.cpp compiled out, linked
// against simple-spectral_b200/libssb200.so) and run by tests/test_gpu_boundary.py: the image file the reference then
// writes must equal, byte for byte, what the reference's own CPU loop writes at the same per-sample seeds.
//
// Everything above the seam stays the reference's own code: main.cpp's command line, Color::init, Scene::get_new_*,
// Renderer::Options, Framebuffer::save.  What is replaced is Renderer::render_start() / _render_threadwork /
// _render_pixel / _render_sample (renderer.cpp:103-422): flatten the pointer graph once, one call, save.
//
// Needs read access to a few private members (_Spectrum::_data/_low/_high, sRGB_ReflectanceTexture::_data): the
// maintainer adds `friend class Renderer;`; the scratch build turns `private:` into `public:` in those two headers,
// as the parity hooks already do.
#include "renderer.hpp"

#include "geometry.hpp"
#include "material.hpp"
#include "scene.hpp"
#include "util/color.hpp"

#include <cstdlib>
#include <cstring>
#include <map>

extern "C" {
#include "ssb200.h"
#ifdef RENDER_MODE_SPECTRAL_JH
#include "jakob-and-hanika-2019/rgb2spec.h"  // struct _RGB2Spec { res, scale, data } (as util/color.cpp:7 includes it)
#endif
}

#ifndef RENDER_MODE_SPECTRAL
#error "this stub covers the spectral build (the RGB build fills ssb_material.albedo_rgb / emission_rgb instead; INTEGRATION.md)"
#endif
#if !defined RENDER_MODE_SPECTRAL_OURS && !defined RENDER_MODE_SPECTRAL_JH
#error "Meng et al.'s grid is compiled into the reference (spectrum_grid.h): pass it as ssb_meng_tables (tools/gen_meng_tables.c shows the layout)"
#endif

namespace {
ssb_spectrum flat(_Spectrum const& s) {
	ssb_spectrum f;
	f.data = s._data.data(); f.n = static_cast<uint32_t>(s._data.size()); f.low = s._low; f.high = s._high; f.filter = SSB_FILTER_LINEAR;
	return f;
}
}  // namespace

void Renderer::render_start() {  // replaces renderer.cpp:396-422
	_time_start = std::chrono::steady_clock::now();
	_num_rendering = 0u;

	// 1. Scene -> ssb_scene (list order preserved: it defines tie-breaks and the identity used by `ignore`)
	std::vector<ssb_quad> quads;
	std::vector<ssb_material> mats;
	std::vector<ssb_texture> texs;
	std::map<MaterialBase const*, uint32_t> mat_index;
	for (PrimBase const* prim : scene->primitives) {
		PrimQuad const* q = static_cast<PrimQuad const*>(prim);
		auto it = mat_index.find(q->material);
		if (it == mat_index.end()) {
			auto const* m = static_cast<MaterialSimpleAlbedoBase const*>(q->material);
			ssb_material fm;
			std::memset(&fm, 0, sizeof(fm));
			fm.kind = dynamic_cast<MaterialLambertian const*>(q->material) ? SSB_MATERIAL_LAMBERT : SSB_MATERIAL_MIRROR;
			fm.emission = flat(m->emission);
			if (m->mode == MaterialSimpleAlbedoBase::CONSTANT) {
				fm.albedo_mode = SSB_ALBEDO_CONSTANT;
				fm.albedo = flat(*m->albedo.constant);
			} else {
				fm.albedo_mode = SSB_ALBEDO_TEXTURE;
				fm.texture = static_cast<uint32_t>(texs.size());
				ssb_texture t;
				t.rgb8 = reinterpret_cast<uint8_t const*>(m->albedo.texture->_data);
				t.width = static_cast<uint32_t>(m->albedo.texture->res[0]); t.height = static_cast<uint32_t>(m->albedo.texture->res[1]);
				texs.push_back(t);
			}
			it = mat_index.emplace(q->material, static_cast<uint32_t>(mats.size())).first;
			mats.push_back(fm);
		}
		ssb_quad fq;
		std::memset(&fq, 0, sizeof(fq));
		PrimTri const* tris[2] = { &q->tri0, &q->tri1 };
		for (int t = 0; t < 2; ++t) {
			for (int v = 0; v < 3; ++v) {
				std::memcpy(fq.tri[t].v[v].pos, &tris[t]->verts[v].pos[0], 12);
				std::memcpy(fq.tri[t].v[v].st, &tris[t]->verts[v].st[0], 8);
			}
			std::memcpy(fq.tri[t].normal, &tris[t]->normal[0], 12);
		}
		fq.material = it->second;
		fq.is_light = q->is_light ? 1u : 0u;
		quads.push_back(fq);
	}
	ssb_scene fs;
	std::memset(&fs, 0, sizeof(fs));
	for (size_t c = 0; c < 4; ++c) for (size_t r = 0; r < 4; ++r) fs.camera.pv_inv[c * 4 + r] = scene->camera.matr_PV_inv[c][r];  // column-major
	std::memcpy(fs.camera.pos, &scene->camera.pos[0], 12);
	std::memcpy(fs.camera.dir, &scene->camera.dir[0], 12);
	fs.quads = quads.data(); fs.nquads = static_cast<uint32_t>(quads.size());
	fs.materials = mats.data(); fs.nmaterials = static_cast<uint32_t>(mats.size());
	fs.textures = texs.data(); fs.ntextures = static_cast<uint32_t>(texs.size());

	// 2. Color::data -> ssb_color
	ssb_color fc;
	std::memset(&fc, 0, sizeof(fc));
	fc.xbar = flat(Color::data->std_obs_xbar); fc.ybar = flat(Color::data->std_obs_ybar); fc.zbar = flat(Color::data->std_obs_zbar);
#if defined RENDER_MODE_SPECTRAL_OURS
	fc.basis_r = flat(Color::data->basis_bt709.r); fc.basis_g = flat(Color::data->basis_bt709.g); fc.basis_b = flat(Color::data->basis_bt709.b);
#elif defined RENDER_MODE_SPECTRAL_JH
	fc.jh_scale = Color::data->model_jh2019->scale; fc.jh_data = Color::data->model_jh2019->data; fc.jh_res = Color::data->model_jh2019->res;
#endif
	for (size_t c = 0; c < 3; ++c) for (size_t r = 0; r < 3; ++r) fc.xyz_to_lrgb[c * 3 + r] = Color::data->matr_xyz_to_lrgb[c][r];
	fc.d65_rad_Y = Color::data->D65_rad_XYZ.y;

	// 3. options: Renderer::Options + the compile-time macros of stdafx.hpp
	ssb_options o;
	ssb_default_options(&o, static_cast<uint32_t>(options.res[0]), static_cast<uint32_t>(options.res[1]), static_cast<uint32_t>(options.spp));
	o.indirect_only = options.indirect_only ? 1u : 0u;
	o.upsampling = RENDER_MODE_SPECTRAL_ALGNUM;
	o.lambda_min = LAMBDA_MIN; o.lambda_max = LAMBDA_MAX;
	o.max_depth = MAX_DEPTH; o.eps = EPS;
	o.n_wavelengths = static_cast<uint32_t>(SAMPLE_WAVELENGTHS);
#ifdef EXPLICIT_LIGHT_SAMPLING
	o.explicit_light_sampling = 1;
#else
	o.explicit_light_sampling = 0;
#endif
#ifdef FLAT_FIELD_CORRECTION
	o.flat_field_correction = 1;
#else
	o.flat_field_correction = 0;
#endif
	if (const char* s = std::getenv("SSB_SEED")) o.seed = std::strtoull(s, nullptr, 10);  // the per-sample seeding of the parity hooks

	// 4. render straight into the reference's framebuffer (float4 sRGBA, row 0 = bottom), then what the last worker thread does
	ssb_ctx* ctx = nullptr;
	int device = 0;
	if (const char* s = std::getenv("SSB_DEVICE")) device = std::atoi(s);
	int rc = ssb_create(device, &ctx);
	if (rc == SSB_OK) rc = ssb_upload_scene(ctx, &fs);
	if (rc == SSB_OK) rc = ssb_upload_color(ctx, &fc);
	if (rc == SSB_OK) rc = ssb_render_frame(ctx, &o, /*xyza=*/nullptr, /*srgba=*/reinterpret_cast<float*>(&framebuffer(0, 0)));
	if (rc != SSB_OK) {
		fprintf(stderr, "ssb200: %s\n", ssb_last_error());
		if (ctx) ssb_destroy(ctx);
		throw rc;  // the reference's `throw int` convention (-1 data / CUDA, -2 argument, -3 unsupported)
	}
	ssb_destroy(ctx);
	_print_progress();
	framebuffer.save(options.output_path);  // renderer.cpp:388-394
}
void Renderer::render_wait() {}  // replaces renderer.cpp:423-430: the call above is synchronous
