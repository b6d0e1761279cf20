// host_capi.cpp — C wrappers (include/ssb200_host.h) over the C++ host layer, for ctypes / cgo-style bindings.
#include <cstring>
#include <string>

#include "../../../include/ssb200_host.h"
#include "ssb_host.hpp"

using namespace ssbh;

struct ssbh_color { ColorData data; ssb_color flat; };
struct ssbh_scene { Scene data; };
struct ssbh_renderer { Renderer* r; };

namespace {
thread_local std::string g_err;
template <class F> int guard(F&& f) {
	try { f(); return SSB_OK; }
	catch (Error const& e) { g_err = e.message; return e.code; }
	catch (std::exception const& e) { g_err = e.what(); return SSB_ERR_DATA; }
}
}  // namespace

extern "C" {

const char* ssbh_last_error(void) { return g_err.c_str(); }

int ssbh_color_init(const char* data_root, int observer, uint32_t upsampling, ssbh_color** out) {
	if (!data_root || !out) { g_err = "ssbh_color_init: NULL argument"; return SSB_ERR_ARG; }
	*out = nullptr;
	return guard([&] {
		ssbh_color* c = new ssbh_color{ color_init(data_root, observer, upsampling), ssb_color{} };
		c->flat = c->data.flat();
		*out = c;
	});
}
const ssb_color* ssbh_color_flat(const ssbh_color* c) { return c ? &c->flat : nullptr; }
void ssbh_color_free(ssbh_color* c) { delete c; }
int ssbh_color_query(const ssbh_color* c, float* lambda_min_max, float* d65_orig_xyz, float* d65_rad_xyz, float* lrgb_to_xyz9, float* xyz_to_lrgb9) {
	if (!c) { g_err = "ssbh_color_query: NULL"; return SSB_ERR_ARG; }
	if (lambda_min_max) { lambda_min_max[0] = c->data.lambda_min; lambda_min_max[1] = c->data.lambda_max; }
	if (d65_orig_xyz) std::memcpy(d65_orig_xyz, c->data.D65_orig_XYZ, 12);
	if (d65_rad_xyz) std::memcpy(d65_rad_xyz, c->data.D65_rad_XYZ, 12);
	if (lrgb_to_xyz9) std::memcpy(lrgb_to_xyz9, c->data.matr_lrgb_to_xyz, 36);
	if (xyz_to_lrgb9) std::memcpy(xyz_to_lrgb9, c->data.matr_xyz_to_lrgb, 36);
	return SSB_OK;
}
int ssbh_color_spectrum(const ssbh_color* c, const char* name, ssb_spectrum* out) {
	if (!c || !name || !out) { g_err = "ssbh_color_spectrum: NULL"; return SSB_ERR_ARG; }
	std::string n = name;
	Spectrum const* s = n == "D65_orig" ? &c->data.D65_orig : n == "D65_rad" ? &c->data.D65_rad : n == "xbar" ? &c->data.std_obs_xbar :
	                    n == "ybar" ? &c->data.std_obs_ybar : n == "zbar" ? &c->data.std_obs_zbar : n == "basis_r" ? &c->data.basis_r :
	                    n == "basis_g" ? &c->data.basis_g : n == "basis_b" ? &c->data.basis_b : nullptr;
	if (!s) { g_err = "unknown spectrum name"; return SSB_ERR_ARG; }
	*out = s->flat();
	return SSB_OK;
}

int ssbh_color_round_trip_srgb(const ssbh_color* c, const float* srgb3, float* out3) {
	if (!c || !srgb3 || !out3) { g_err = "ssbh_color_round_trip_srgb: NULL argument"; return SSB_ERR_ARG; }
	return guard([&] { c->data.round_trip_srgb(srgb3, out3); });
}
int ssbh_color_round_trip_running_max(const ssbh_color* c, uint32_t r_begin, uint32_t r_end, float start_max, float* running_max, uint32_t threads) {
	if (!c || !running_max || r_begin > r_end || r_end > 256) { g_err = "ssbh_color_round_trip_running_max: bad argument"; return SSB_ERR_ARG; }
	return guard([&] { c->data.round_trip_running_max(r_begin, r_end, start_max, running_max, threads); });
}

int ssbh_scene_new(const char* name, const char* data_root, const ssbh_color* color, int explicit_light_sampling, ssbh_scene** out) {
	if (!name || !data_root || !color || !out) { g_err = "ssbh_scene_new: NULL argument"; return SSB_ERR_ARG; }
	*out = nullptr;
	return guard([&] {
		ssbh_scene* s = new ssbh_scene{ scene_new(name, data_root, color->data, explicit_light_sampling != 0) };
		s->data.flatten();  // pointers must refer to the heap copy
		*out = s;
	});
}
const ssb_scene* ssbh_scene_flat(const ssbh_scene* s) { return s ? &s->data.flat : nullptr; }
void ssbh_scene_free(ssbh_scene* s) { delete s; }
int ssbh_scene_camera(const ssbh_scene* s, double* matr_P, double* matr_V, double* matr_PV_inv) {
	if (!s) { g_err = "ssbh_scene_camera: NULL"; return SSB_ERR_ARG; }
	if (matr_P) std::memcpy(matr_P, s->data.camera.matr_P, 128);
	if (matr_V) std::memcpy(matr_V, s->data.camera.matr_V, 128);
	if (matr_PV_inv) std::memcpy(matr_PV_inv, s->data.camera.matr_PV_inv, 128);
	return SSB_OK;
}

int ssbh_load_png_rgb8(const char* path, uint8_t** rgb8, uint32_t* width, uint32_t* height) {
	if (!path || !rgb8 || !width || !height) { g_err = "ssbh_load_png_rgb8: NULL argument"; return SSB_ERR_ARG; }
	return guard([&] {
		Texture t = load_png_rgb8(path);
		uint8_t* p = static_cast<uint8_t*>(malloc(t.rgb8.size()));
		std::memcpy(p, t.rgb8.data(), t.rgb8.size());
		*rgb8 = p; *width = t.width; *height = t.height;
	});
}
void ssbh_free(void* p) { free(p); }

int ssbh_save_image(const char* path, const float* srgba, uint32_t width, uint32_t height) {
	if (!path || !srgba) { g_err = "ssbh_save_image: NULL argument"; return SSB_ERR_ARG; }
	return guard([&] {
		Framebuffer fb;
		fb.res[0] = width; fb.res[1] = height;
		fb.pixels.assign(srgba, srgba + static_cast<size_t>(width) * height * 4);
		fb.save(path);
	});
}

int ssbh_renderer_new(const ssbh_renderer_options* o, ssbh_renderer** out) {
	if (!o || !out || !o->scene_name) { g_err = "ssbh_renderer_new: NULL argument"; return SSB_ERR_ARG; }
	*out = nullptr;
	return guard([&] {
		RendererOptions r;
		r.scene_name = o->scene_name;
		r.res[0] = o->width; r.res[1] = o->height; r.spp = o->spp;
		r.indirect_only = o->indirect_only != 0;
		r.output_path = o->output_path ? o->output_path : "";
		r.observer = o->observer; r.upsampling = o->upsampling;
		r.explicit_light_sampling = o->explicit_light_sampling != 0;
		r.max_depth = o->max_depth; r.flat_field_correction = o->flat_field_correction != 0;
		r.seed = o->seed; r.device = o->device;
		r.render_mode = o->render_mode;
		r.n_wavelengths = o->n_wavelengths ? o->n_wavelengths : 4u;
		r.data_root = o->data_root ? o->data_root : ".";
		r.prebaked_textures = o->prebaked_textures != 0;
		r.progressive = o->progressive != 0;
		if (o->ndevices) {
			if (!o->devices) throw Error{ SSB_ERR_ARG, "ssbh_renderer_new: devices is NULL" };
			r.devices.assign(o->devices, o->devices + o->ndevices);
		}
		if (o->shard > SSBH_SHARD_SAMPLES) throw Error{ SSB_ERR_UNSUPPORTED, "ssbh_renderer_new: unknown shard mode" };
		r.shard = o->shard;
		if (o->band_height) r.band_height = o->band_height;
		*out = new ssbh_renderer{ new Renderer(r) };
	});
}
int ssbh_renderer_render(ssbh_renderer* r) {
	if (!r) { g_err = "ssbh_renderer_render: NULL"; return SSB_ERR_ARG; }
	return guard([&] { r->r->render_start(); r->r->render_wait(); });
}
int ssbh_renderer_start(ssbh_renderer* r) {
	if (!r) { g_err = "ssbh_renderer_start: NULL"; return SSB_ERR_ARG; }
	return guard([&] { r->r->render_start(); });
}
void ssbh_renderer_stop(ssbh_renderer* r) { if (r) r->r->render_stop(); }
int ssbh_renderer_wait(ssbh_renderer* r) {
	if (!r) { g_err = "ssbh_renderer_wait: NULL"; return SSB_ERR_ARG; }
	return guard([&] { r->r->render_wait(); });
}
int ssbh_renderer_is_rendering(const ssbh_renderer* r) { return (r && r->r->is_rendering()) ? 1 : 0; }
uint32_t ssbh_renderer_snapshot(const ssbh_renderer* r, float* srgba) {
	if (!r) return 0;
	if (!srgba) return r->r->samples_done();
	std::vector<float> tmp;
	const uint32_t done = r->r->snapshot(tmp);
	std::memcpy(srgba, tmp.data(), tmp.size() * sizeof(float));
	return done;
}
const float* ssbh_renderer_framebuffer(const ssbh_renderer* r) { return r ? r->r->framebuffer.pixels.data() : nullptr; }
const double* ssbh_renderer_xyza(const ssbh_renderer* r) { return (r && !r->r->xyza.empty()) ? r->r->xyza.data() : nullptr; }
int ssbh_renderer_stats(const ssbh_renderer* r, ssb_stats* out) {
	if (!r || !out) { g_err = "ssbh_renderer_stats: NULL"; return SSB_ERR_ARG; }
	*out = r->r->last_stats;
	return SSB_OK;
}
void ssbh_renderer_free(ssbh_renderer* r) { if (r) { delete r->r; delete r; } }

}  // extern "C"
