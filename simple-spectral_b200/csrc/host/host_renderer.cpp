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
	devices_ = options.devices.empty() ? std::vector<int>{ options.device } : options.devices;
	ctxs_.assign(devices_.size(), nullptr);
	// one thread per device: context creation and the texture upload (48 MiB) of the GPUs overlap
	std::vector<Error> errs(devices_.size(), Error{ 0, "" });
	ssb_color fc = color.flat();
	auto setup = [&](size_t d) {
		int rc = ssb_create(devices_[d], &ctxs_[d]);
		if (rc == SSB_OK) rc = ssb_upload_color(ctxs_[d], &fc);
		if (rc == SSB_OK) rc = ssb_upload_scene(ctxs_[d], &scene.flat);
		if (rc != SSB_OK) errs[d] = Error{ rc, ssb_last_error() };
	};
	if (devices_.size() == 1) setup(0);
	else {
		std::vector<std::thread> th;
		for (size_t d = 0; d < devices_.size(); ++d) th.emplace_back(setup, d);
		for (auto& t : th) t.join();
	}
	Error bad{ 0, "" };
	for (auto const& e : errs) if (e.code != 0 && bad.code == 0) bad = e;
	if (bad.code == 0 && devices_.size() > 1) {
		int rc = ssb_create(devices_[0], &sum_ctx_);
		if (rc == SSB_OK) rc = ssb_upload_color(sum_ctx_, &fc);  // resolve needs the XYZ -> l-RGB matrix
		if (rc != SSB_OK) bad = Error{ rc, ssb_last_error() };
	}
	if (bad.code != 0) {
		for (ssb_ctx* c : ctxs_) if (c) ssb_destroy(c);
		if (sum_ctx_) ssb_destroy(sum_ctx_);
		ctxs_.clear(); sum_ctx_ = nullptr;
		throw bad;
	}
	ctx_ = sum_ctx_ ? sum_ctx_ : ctxs_[0];
}
Renderer::~Renderer() {
	continue_ = false;
	if (worker_.joinable()) worker_.join();
	for (ssb_ctx* c : ctxs_) if (c) ssb_destroy(c);
	if (sum_ctx_) ssb_destroy(sum_ctx_);
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
// One slice of the sample range on every device.  Single GPU: one ssb_render.  Several: each device renders its share
// (its interleaved row bands of the slice's samples, or its part of the slice's sample range) from its own host thread —
// the ~30 kernel launches of a pass are enqueued concurrently, not device after device — into its own accumulator.
void Renderer::render_slice(ssb_options const& slice, ssb_stats& sum) {
	const size_t nd = ctxs_.size();
	if (nd == 1) {
		int rc = ssb_render(ctxs_[0], &slice);
		if (rc != SSB_OK) throw Error{ rc, ssb_last_error() };
		ssb_stats st{};
		ssb_get_stats(ctxs_[0], &st);
		sum.samples += st.samples; sum.device_ms += st.device_ms; sum.trace_ms += st.trace_ms; sum.launches += st.launches;
		return;
	}
	std::vector<ssb_options> part(nd, slice);
	std::vector<char> active(nd, 1);
	const uint32_t s0 = slice.sample_begin, s1 = slice.sample_end ? slice.sample_end : slice.spp;
	for (size_t d = 0; d < nd; ++d) {
		part[d].keep_accumulator = 1;  // the accumulators are cleared once per frame (work()): later slices add to them
		if (options.shard == RendererOptions::SHARD_SAMPLES) {
			const uint32_t b = s0 + (uint32_t)((uint64_t)(s1 - s0) * d / nd), e = s0 + (uint32_t)((uint64_t)(s1 - s0) * (d + 1) / nd);
			part[d].sample_begin = b; part[d].sample_end = e;
			active[d] = e > b;
			if (e == 0) active[d] = 0;  // (sample_end == 0 would mean "all")
		} else {
			part[d].band_height = options.band_height; part[d].band_count = (uint32_t)nd; part[d].band_index = (uint32_t)d;
			active[d] = (uint64_t)d * options.band_height < slice.height;
		}
	}
	std::vector<Error> errs(nd, Error{ 0, "" });
	std::vector<std::thread> th;
	for (size_t d = 0; d < nd; ++d) {
		if (!active[d]) continue;
		th.emplace_back([&, d] {
			int rc = ssb_render(ctxs_[d], &part[d]);
			if (rc != SSB_OK) errs[d] = Error{ rc, ssb_last_error() };
		});
	}
	for (auto& t : th) t.join();
	for (auto const& e : errs) if (e.code != 0) throw e;
	double ms = 0, trace = 0;
	for (size_t d = 0; d < nd; ++d) {
		if (!active[d]) continue;
		ssb_stats st{};
		ssb_get_stats(ctxs_[d], &st);  // (waits for that device: the GPUs run concurrently, the slice takes the longest of them)
		sum.samples += st.samples; sum.launches += st.launches;
		ms = std::max(ms, st.device_ms); trace = std::max(trace, st.trace_ms);
	}
	sum.device_ms += ms; sum.trace_ms += trace;
}

// The worker: the whole frame in one device call, or — progressive — in sample slices of doubling size, the
// framebuffer refreshed after each slice with the average so far (`ssb_resolve` with spp = samples done).  Sample
// ranges accumulate in sample order into the context's f64 accumulator (ssb_render does not clear it for
// sample_begin > 0), so the finished frame is bit-identical to the single-call one.  With several GPUs the per-device
// accumulators are merged into a separate accumulator on the first device before each resolve (ssb_accum_merge).
void Renderer::work() {
	auto t0 = std::chrono::steady_clock::now();
	std::printf("\rRender started                               ");
	std::fflush(stdout);
	ssb_stats sum{};
	try {
		ssb_options o = make_options();
		std::vector<float> preview(framebuffer.pixels.size());
		const size_t nd = ctxs_.size();
		if (nd > 1) {  // every device needs a cleared accumulator of the frame's size before the first (partial) render
			for (ssb_ctx* c : ctxs_) {
				ssb_options empty = o;  // an empty pixel rectangle: allocates and clears the accumulator, traces nothing
				empty.x0 = o.width; empty.x1 = o.width;
				int rc = ssb_render(c, &empty);
				if (rc != SSB_OK) throw Error{ rc, ssb_last_error() };
			}
		}
		uint32_t done = 0;
		bool go = continue_;
		while (done < o.spp && go) {
			ssb_options slice = o;
			slice.sample_begin = done;
			slice.sample_end = options.progressive ? std::min(o.spp, done == 0 ? 1u : 2u * done) : o.spp;
			render_slice(slice, sum);
			done = slice.sample_end;
			go = continue_;  // read ONCE per slice: `last` and the loop exit below must agree (a render_stop() landing in
			                 // between would otherwise end the loop after a non-last resolve that skipped the XYZA copy)
			const bool last = done == o.spp || !go;
			ssb_options avg = o;
			avg.spp = done;  // avg = accum * 1000/done (renderer.cpp:296 with the samples so far)
			int rc;
			if (nd > 1) {
				if ((rc = ssb_clear(sum_ctx_)) != SSB_OK) throw Error{ rc, ssb_last_error() };
				for (size_t d = 0; d < nd; ++d) {
					ssb_options how = o;  // what device d holds: its bands of the frame, or partial sums of all pixels
					if (options.shard != RendererOptions::SHARD_SAMPLES) { how.band_height = options.band_height; how.band_count = (uint32_t)nd; how.band_index = (uint32_t)d; }
					if ((rc = ssb_accum_merge(sum_ctx_, ctxs_[d], &how)) != SSB_OK) throw Error{ rc, ssb_last_error() };
				}
				sum.launches += (uint32_t)nd;
			}
			if ((rc = ssb_resolve(ctx_, &avg, last ? xyza.data() : nullptr, preview.data())) != SSB_OK) throw Error{ rc, ssb_last_error() };
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
	} catch (std::exception const& e) {  // bad_alloc & co. must reach render_wait(), not std::terminate
		error_ = Error{ SSB_ERR_DATA, e.what() }; failed_ = true;
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
