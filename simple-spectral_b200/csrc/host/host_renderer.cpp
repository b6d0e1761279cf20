// host_renderer.cpp — Renderer façade (reference src/renderer.hpp:13-82, renderer.cpp:12-51,396-430): same
// Options / render_start / render_wait / framebuffer / scene surface; the per-pixel Monte-Carlo loop
// (renderer.cpp:103-395) is replaced by one call into the CUDA path through the C ABI.
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
Renderer::~Renderer() { if (ctx_) ssb_destroy(ctx_); }

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
	return o;
}

void Renderer::render_start() {
	ssb_options o = make_options();
	xyza.assign(static_cast<size_t>(o.width) * o.height * 4, 0.0);
	auto t0 = std::chrono::steady_clock::now();
	std::printf("\rRender started                               ");
	int rc = ssb_render_frame(ctx_, &o, xyza.data(), framebuffer.pixels.data());
	if (rc != SSB_OK) throw Error{ rc, ssb_last_error() };
	ssb_get_stats(ctx_, &last_stats);
	double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	std::printf("\rRender completed in %02d:%02d:%06.3f             \n", (int)(secs / 3600), (int)(secs / 60) % 60, secs - 60.0 * (int)(secs / 60));  // renderer.cpp:93-99
	rendered_ = true;
}

void Renderer::render_wait() {
	if (rendered_ && !options.output_path.empty()) framebuffer.save(options.output_path);  // renderer.cpp:388-394
}

}  // namespace ssbh
