// ssb_host.hpp — host side of the drop-in: a C++ mirror of the reference's Scene / Material /
// Color / Framebuffer / Renderer surface for the hot path (the reference is C++17, so the host
// layer above the C ABI is C++ too).  It loads the same data/*.csv spectra, texture and tables
// (cwd-relative "data/..." paths under a data root, as the reference does), builds the same
// hard-coded scenes, flattens them to the POD structs of include/ssb200.h and drives the CUDA
// path through the C ABI.  Init-time arithmetic (Color::init, camera matrices) follows the
// reference operation by operation, in GLM 0.9.9 scalar semantics, so that the uploaded tables
// are bit-identical to what the reference computes (checked against dumps of the real reference:
// tests/test_host_layer.py).
#pragma once

#include <atomic>
#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/ssb200.h"

namespace ssbh {

// reference: `throw int` codes carried with a message
struct Error {
	int code;
	std::string message;
};

// ---- _Spectrum (src/spectrum.hpp:12-70, spectrum.cpp:11-173)
struct Spectrum {
	std::vector<float> data;
	float low = 0, high = 0;
	float delta_lambda = 0, delta_lambda_recip = 0;
	uint32_t filter = SSB_FILTER_LINEAR;

	Spectrum() = default;
	Spectrum(float value, float lambda_min, float lambda_max);  // constant over [LAMBDA_MIN,LAMBDA_MAX]
	Spectrum(std::vector<float> const& data, float low, float high);

	float sample_nearest(float lambda) const;
	float sample_linear(float lambda) const;
	Spectrum operator*(float sc) const;
	Spectrum operator*(Spectrum const& other) const;  // spectrum.cpp:74-96 (sample-wise, on the common range)
	Spectrum operator+(Spectrum const& other) const;  // spectrum.cpp:97-118
	static float integrate(Spectrum const& spec);
	static float integrate(Spectrum const& spec0, Spectrum const& spec1);
	ssb_spectrum flat() const;
};

std::vector<std::vector<float>> load_spectral_data(std::string const& csv_path);  // spectrum.cpp:177-213

// ---- PNG texture (material.cpp:10-29; the reference decodes with lodepng, we with zlib)
struct Texture {
	uint32_t width = 0, height = 0;
	std::vector<uint8_t> rgb8;  // scanlines top-to-bottom
};
Texture load_png_rgb8(std::string const& path);

// ---- Color::data (src/util/color.hpp:22-69, color.cpp:72-155)
struct ColorData {
	int observer = 1931;       // CIE_OBSERVER
	uint32_t upsampling = SSB_UPSAMPLE_OURS;
	float lambda_min = 380, lambda_max = 780;
	Spectrum std_obs_xbar, std_obs_ybar, std_obs_zbar;
	Spectrum D65_orig, D65_rad;
	float D65_orig_XYZ[3] = { 0, 0, 0 }, D65_rad_XYZ[3] = { 0, 0, 0 };
	Spectrum basis_r, basis_g, basis_b;
	float matr_lrgb_to_xyz[9] = { 0 }, matr_xyz_to_lrgb[9] = { 0 };  // column-major
	// Jakob-Hanika model (rgb2spec.c:16-43)
	uint32_t jh_res = 0;
	std::vector<float> jh_scale, jh_data;
	// Meng et al. tables (serialised by tools/gen_meng_tables.c)
	ssb_meng_tables meng{};
	std::vector<int32_t> meng_grid;
	std::vector<float> meng_points;
	bool have_meng = false;

	ssb_color flat() const;  // pointers into this object

	// Color::round_trip_lrgb / round_trip_srgb (color.cpp:259-294, OURS only): colour -> spectrum by the basis ->
	// D65 reflected off it -> XYZ by integration against the observer -> colour.  The reference's self-test
	// (main.cpp:184-264) documents its maximum error over all 2^24 sRGB values: 1.851469e-5.
	void round_trip_lrgb(const float lrgb[3], float out[3]) const;
	void round_trip_srgb(const float srgb[3], float out[3]) const;
	// the self-test's loop for red levels [r_begin, r_end): running_max[k] = the maximum error after level r_begin + k,
	// starting from `start_max` (main.cpp:246-262); green levels are spread over `threads` threads (a maximum does not
	// depend on the order)
	void round_trip_running_max(uint32_t r_begin, uint32_t r_end, float start_max, float* running_max, unsigned threads) const;
};
// Color::init(): observer 1931|2006, upsampling SSB_UPSAMPLE_*; data_root contains "data/"
ColorData color_init(std::string const& data_root, int observer, uint32_t upsampling);

// ---- materials / scene (src/material.hpp, scene.hpp, scene.cpp)
struct Material {
	std::string name;
	uint32_t kind = SSB_MATERIAL_LAMBERT;
	uint32_t albedo_mode = SSB_ALBEDO_CONSTANT;
	Spectrum albedo;
	int texture = -1;
	Spectrum emission;
	// the same material in the reference's RENDER_MODE_RGB build (the `#else` branches of scene.cpp:48-101,297-343)
	float albedo_rgb[3] = { 1, 1, 1 };    // Albedo(): RGB_Reflectance(1.0f), material.hpp:131
	float emission_rgb[3] = { 0, 0, 0 };  // MaterialBase(): emission(0.0f)
	bool is_emissive() const;  // material.cpp:100-106
};

struct Camera {  // scene.hpp:16-33
	float pos[3], dir[3], up[3];
	uint32_t res[2];
	float near_, far_, vfov_deg;
	double matr_P[16], matr_V[16], matr_PV_inv[16];
};

struct Scene {
	std::string name;
	Camera camera{};
	std::vector<Material> materials;  // first-use order over the primitive list
	std::vector<Texture> textures;
	std::vector<ssb_quad> quads;      // insertion order = tie-break / ignore identity
	std::vector<uint32_t> lights;

	// flat view (pointers into this object; rebuilt by flatten())
	std::vector<ssb_material> flat_materials;
	std::vector<ssb_texture> flat_textures;
	ssb_scene flat{};
	void flatten();
};
// Scene::get_new_cornell / get_new_cornell_srgb / get_new_plane_srgb (scene.cpp:32-415); unknown name -> Error{-3}
Scene scene_new(std::string const& name, std::string const& data_root, ColorData const& color, bool explicit_light_sampling);

// ---- Framebuffer (src/framebuffer.hpp:8-43, framebuffer.cpp)
struct Framebuffer {
	uint32_t res[2] = { 0, 0 };
	std::vector<float> pixels;  // sRGBA float4, row 0 = bottom
	void reset(uint32_t w, uint32_t h);  // checkerboard init (framebuffer.cpp:15-33)
	void save(std::string const& path) const;  // .csv / .hdr / .pfm / else PNG (framebuffer.cpp:39-176)
};

// ---- Renderer (src/renderer.hpp:13-82): same Options, render_start/wait/stop/is_rendering
struct RendererOptions {
	std::string scene_name;
	uint32_t res[2] = { 0, 0 };
	uint32_t spp = 0;
	bool indirect_only = false;
	std::string output_path;
	// the reference's compile-time configuration (stdafx.hpp:44-90), runtime here
	int observer = 1931;
	uint32_t upsampling = SSB_UPSAMPLE_OURS;
	bool explicit_light_sampling = true;
	uint32_t max_depth = 10;
	bool flat_field_correction = true;
	uint32_t render_mode = SSB_RENDER_SPECTRAL;  // SSB_RENDER_RGB = the RENDER_MODE_RGB build
	uint32_t n_wavelengths = 4;                  // SAMPLE_WAVELENGTHS (stdafx.hpp:90): 2, 3 or 4
	uint64_t seed = 1;
	int device = 0;
	// Multi-GPU (the reference's Renderer spreads its tiles over every core of the box, renderer.cpp:396-430): render on all
	// of `devices` (empty: just `device`; a device may be listed more than once), one host thread and one context per
	// entry, the frame split as `shard` says, the f64 accumulators merged on the first device (ssb_accum_merge: peer
	// loads over NVLink) and resolved there.
	std::vector<int> devices;
	enum Shard : uint32_t {
		SHARD_TILES = 0,    // interleaved bands of band_height rows (ssb_options.band_*): disjoint pixels, bit-identical to one GPU
		SHARD_SAMPLES = 1,  // every GPU renders all pixels for a slice of the sample range: partial sums added in device order
		                    // (differs from one GPU only by the order of the f64 additions, ~1e-16 relative)
	};
	uint32_t shard = SHARD_TILES;
	uint32_t band_height = 8;  // rows per band: the reference's tile edge (framebuffer.hpp:14-21 tiles are 8 x 8)
	std::string data_root = ".";
	// JH only: texel -> coefficient pre-process once per texture (ssb_options.prebaked_textures; color.cpp:204-216)
	bool prebaked_textures = false;
	// Progressive preview (what the reference's window shows while its tiles fill in, main.cpp:316-323): the frame is
	// rendered in sample slices [0,1) [1,2) [2,4) [4,8) ... and `framebuffer` is refreshed after each one with the
	// average of the samples so far.  The final image does not depend on the slicing (samples are accumulated in order).
	bool progressive = false;
};

class Renderer {
public:
	RendererOptions const options;
	Framebuffer framebuffer;  // while is_rendering(): read it through snapshot()
	ColorData color;
	Scene scene;
	std::vector<double> xyza;  // per-pixel double XYZA (the reference's local `avg`, renderer.cpp:292-296)

	explicit Renderer(RendererOptions const& options);
	~Renderer();
	Renderer(Renderer const&) = delete;
	Renderer& operator=(Renderer const&) = delete;

	// Same life cycle as the reference (renderer.hpp:71-81, renderer.cpp:396-430): render_start() returns at once —
	// ONE worker thread feeds the GPU where the reference spawns one thread per core —, render_stop() asks it to
	// end after the slice in flight, render_wait() joins it, rethrows its error, and saves the image as the
	// reference's last worker does (renderer.cpp:388-394; an aborted render is saved with what it has).
	void render_start();
	void render_stop() { continue_ = false; }
	void render_wait();
	bool is_rendering() const { return rendering_; }
	uint32_t samples_done() const { return done_spp_; }  // spp behind the current framebuffer contents
	uint32_t snapshot(std::vector<float>& srgba) const;  // thread-safe copy of the framebuffer; returns samples_done()
	ssb_options make_options() const;
	ssb_stats last_stats{};  // of the whole frame (summed over slices)

	size_t device_count() const { return ctxs_.size(); }

private:
	void work();
	void render_slice(ssb_options const& slice, ssb_stats& sum);
	std::vector<ssb_ctx*> ctxs_;   // one per entry of options.devices
	std::vector<int> devices_;
	ssb_ctx* sum_ctx_ = nullptr;   // multi-GPU only: merge target + resolve, on the first device
	ssb_ctx* ctx_ = nullptr;       // the context that resolves (ctxs_[0], or sum_ctx_)
	std::thread worker_;
	mutable std::mutex fb_mutex_;
	std::atomic<bool> continue_{ true }, rendering_{ false };
	std::atomic<uint32_t> done_spp_{ 0 };
	bool rendered_ = false, failed_ = false;
	Error error_{ 0, "" };
};

}  // namespace ssbh
