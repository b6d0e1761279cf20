// host_renderer.cpp — Renderer façade (reference src/renderer.hpp:13-82, renderer.cpp:12-51,396-430): same
// Options / render_start / render_wait / framebuffer / scene surface; the per-pixel Monte-Carlo loop
// (renderer.cpp:103-395) is replaced by one call into the CUDA path through the C ABI.
#include <algorithm>
#include <chrono>
#include <cstdio>

#include "ssb_host.hpp"

namespace ssbh {

Renderer::Renderer(RendererOptions const& opts) : options(opts) {
	framebuffer.reset(options.res[0], options.res[1]);
	color = color_init(options.data_root, options.observer, options.upsampling);  // Color::init(), main.cpp:181
	scene = scene_new(options.scene_name, options.data_root, color, options.explicit_light_sampling);
	scene.flatten();  // the flat view holds pointers into the Scene object: rebuild it for this copy
	if (options.scene_name == "plane-srgb" && options.explicit_light_sampling)  // renderer.cpp:27-31
		std::fprintf(stderr, "Warning: Plane converges much faster without explicit light sampling!\n");
	if (options.scene_name != "plane-srgb" && !options.explicit_light_sampling)  // renderer.cpp:18-26
		std::fprintf(stderr, "Warning: Cornell converges much faster with explicit light sampling!\n");
	int rc = ssb_create(options.device, &ctx_);
	if (rc != SSB_OK) throw Error{ rc, ssb_last_error() };
	ssb_color fc = color.flat();
	if ((rc = ssb_upload_color(ctx_, &fc)) != SSB_OK || (rc = ssb_upload_scene(ctx_, &scene.flat)) != SSB_OK) {
		std::string msg = ssb_last_error();
		ssb_destroy(ctx_);
		ctx_ = nullptr;
		throw Error{ rc, msg };
	}
}
Renderer::~Renderer() {
	continue_ = false;
	if (worker_.joinable()) worker_.join();
	if (ctx_) ssb_destroy(ctx_);
}

ssb_options Renderer::make_options() const {
	ssb_options o;
	ssb_default_options(&o, options.res[0], options.res[1], options.spp);
	o.indirect_only = options.indirect_only ? 1u : 0u;
	o.upsampling = options.upsampling;
	o.lambda_min = color.lambda_min; o.lambda_max = color.lambda_max;
	o.max_depth = options.max_depth;
	o.explicit_light_sampling = options.explicit_light_sampling ? 1u : 0u;
	o.flat_field_correction = options.flat_field_correction ? 1u : 0u;
	o.render_mode = options.render_mode;
	o.n_wavelengths = options.n_wavelengths;
	o.seed = options.seed;
	o.prebaked_textures = options.prebaked_textures ? 1u : 0u;
	return o;
}

void Renderer::render_start() {
	if (worker_.joinable()) throw Error{ SSB_ERR_ARG, "render_start: a render is already in flight (call render_wait first)" };
	continue_ = true; rendering_ = true; done_spp_ = 0;
	rendered_ = false; failed_ = false;
	xyza.assign(static_cast<size_t>(options.res[0]) * options.res[1] * 4, 0.0);
	worker_ = std::thread(&Renderer::work, this);
}

// The worker: the whole frame in one device call, or — progressive — in sample slices of doubling size, the
// framebuffer refreshed after each slice with the average so far (`ssb_resolve` with spp = samples done).  Sample
// ranges accumulate in sample order into the context's f64 accumulator (ssb_render does not clear it for
// sample_begin > 0), so the finished frame is bit-identical to the single-call one.
void Renderer::work() {
	auto t0 = std::chrono::steady_clock::now();
	std::printf("\rRender started                               ");
	std::fflush(stdout);
	ssb_stats sum{};
	try {
		ssb_options o = make_options();
		std::vector<float> preview(framebuffer.pixels.size());
		uint32_t done = 0;
		while (done < o.spp && continue_) {
			ssb_options slice = o;
			slice.sample_begin = done;
			slice.sample_end = options.progressive ? std::min(o.spp, done == 0 ? 1u : 2u * done) : o.spp;
			int rc = ssb_render(ctx_, &slice);
			if (rc != SSB_OK) throw Error{ rc, ssb_last_error() };
			done = slice.sample_end;
			ssb_options avg = o;
			avg.spp = done;  // avg = accum * 1000/done (renderer.cpp:296 with the samples so far)
			const bool last = done == o.spp || !continue_;
			if ((rc = ssb_resolve(ctx_, &avg, last ? xyza.data() : nullptr, preview.data())) != SSB_OK) throw Error{ rc, ssb_last_error() };
			ssb_stats st{};
			ssb_get_stats(ctx_, &st);
			sum.samples += st.samples; sum.device_ms += st.device_ms; sum.trace_ms += st.trace_ms; sum.launches += st.launches;
			{
				std::lock_guard<std::mutex> lock(fb_mutex_);
				framebuffer.pixels.swap(preview);
				done_spp_ = done;
			}
			preview.resize(framebuffer.pixels.size());
			if (options.progressive && !last) {
				std::printf("\rRender progress: %u of %u samples per pixel      ", done, o.spp);
				std::fflush(stdout);
			}
		}
		rendered_ = done > 0;
	} catch (Error const& e) {
		error_ = e; failed_ = true;
	}
	last_stats = sum;
	double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (!failed_)
		std::printf("\rRender %s in %02d:%02d:%06.3f             \n", continue_ ? "completed" : "aborted  ", (int)(secs / 3600), (int)(secs / 60) % 60, secs - 60.0 * (int)(secs / 60));  // renderer.cpp:93-99
	rendering_ = false;
}

void Renderer::render_wait() {
	if (worker_.joinable()) worker_.join();
	if (failed_) { failed_ = false; throw error_; }
	if (rendered_ && !options.output_path.empty()) framebuffer.save(options.output_path);  // renderer.cpp:388-394
	rendered_ = false;  // saved once per render, like the reference's last worker thread
}

uint32_t Renderer::snapshot(std::vector<float>& srgba) const {
	std::lock_guard<std::mutex> lock(fb_mutex_);
	srgba = framebuffer.pixels;
	return done_spp_;
}

}  // namespace ssbh
