// main.cpp — command-line front end with the reference's own flags (reference src/main.cpp:33-162,164-361):
//   --scene=<name>/-s  --width=<w>/-w  --height=<h>/-h  --samples=<n>/-spp  --output=<path>/-o  [--indirect-only/-io]
// so that `simple_spectral_b200 --scene=cornell-srgb -w=512 -h=512 -spp=64 --output=out.png` is a drop-in for the
// reference binary, rendering on the GPU through the Renderer façade.  The reference's compile-time variants are
// extra, optional flags here: --variant=ours1931|ours2006|meng|jh|rgb  --seed=<n>  --device=<n>  --data-root=<dir>
// (default data root: the current directory, like the reference's cwd-relative "data/..." paths), --prebake (Jakob-Hanika
// coefficient textures), and --progressive [--preview=<path>]: the headless stand-in for the reference's window
// (main.cpp:313-327) — the frame refines in sample slices and the preview file is rewritten after each one.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../host/ssb_host.hpp"

namespace {

void print_usage() {
	std::printf(
		"simple_spectral_b200: the simple-spectral renderer's hot path on a B200\n"
		"  Required arguments:\n"
		"    --scene=<name>/-s=<name>            cornell | cornell-srgb | plane-srgb\n"
		"    --width=<w>/-w=<w>  --height=<h>/-h=<h>  --samples=<n>/-spp=<n>\n"
		"    --output=<path>/-o=<path>           .png / .pfm / .hdr / .csv by extension\n"
		"  Optional arguments:\n"
		"    --indirect-only/-io\n"
		"    --variant=ours1931|ours2006|meng|jh|rgb   (the reference's compile-time modes)\n"
		"    --wavelengths=2|3|4                       (SAMPLE_WAVELENGTHS, default 4)\n"
		"    --seed=<n>  --device=<n>  --data-root=<dir containing data/>\n"
		"    --devices=<a-b | a,b,c | all> [--shard=tiles|samples]   (several GPUs: interleaved row bands, bit-identical\n"
		"                                              to one GPU, or sample ranges)\n"
		"    --prebake                                 (variant jh: texel -> coefficient textures, once)\n"
		"    --progressive [--preview=<path>]          (refine in sample slices; rewrite <path> after each)\n");
}

struct Args {
	std::vector<std::string> rest;
	// main.cpp:61-79: "name=value" or bare "name"; throws -2 when absent
	std::string get(std::string const& name, std::string const& shortname) {
		for (auto it = rest.begin(); it != rest.end(); ++it) {
			size_t eq = it->find('=');
			if (eq != std::string::npos) {
				std::string key = it->substr(0, eq);
				if (key == name || key == shortname) { std::string v = it->substr(eq + 1); rest.erase(it); return v; }
			} else if (*it == name || *it == shortname) { rest.erase(it); return name; }
		}
		throw -2;
	}
	std::string req(std::string const& name, std::string const& shortname) {
		try { return get(name, shortname); }
		catch (int) { std::fprintf(stderr, "Required argument `%s`/`%s` not found!\n", name.c_str(), shortname.c_str()); throw; }
	}
};

unsigned to_pos(std::string const& s) {  // util/string.hpp:55-59
	size_t i = 0;
	int v = std::stoi(s, &i);
	if (i != s.size()) throw -1;
	if (v <= 0) throw -2;
	return static_cast<unsigned>(v);
}

}  // namespace

int main(int argc, char* argv[]) {
	ssbh::RendererOptions o;
	std::string preview_path;
	try {
		Args a;
		for (int i = 1; i < argc; ++i) a.rest.emplace_back(argv[i]);
		o.scene_name = a.req("--scene", "-s");
		if (o.scene_name != "cornell" && o.scene_name != "cornell-srgb" && o.scene_name != "plane-srgb") {
			std::fprintf(stderr, "Unrecognized scene \"%s\"!  (Supported scenes: \"cornell\", \"cornell-srgb\", \"plane-srgb\")\n", o.scene_name.c_str());
			throw -3;
		}
		try { o.res[0] = to_pos(a.req("--width", "-w")); o.res[1] = to_pos(a.req("--height", "-h")); }
		catch (int) { std::fprintf(stderr, "Invalid width or height!\n"); throw; }
		catch (std::exception const&) { std::fprintf(stderr, "Invalid width or height!\n"); throw -1; }
		try { o.spp = to_pos(a.req("--samples", "-spp")); }
		catch (int) { std::fprintf(stderr, "Invalid number of samples!\n"); throw; }
		catch (std::exception const&) { std::fprintf(stderr, "Invalid number of samples!\n"); throw -1; }
		try {
			std::string v = a.get("--indirect-only", "-io");
			if (v != "--indirect-only") { std::fprintf(stderr, "`--indirect-only`/`-io` does not take a value!\n"); throw -1; }
			o.indirect_only = true;
		} catch (int code) { if (code != -2) throw; o.indirect_only = false; }
		o.output_path = a.req("--output", "-o");
		try {
			std::string v = a.get("--variant", "--variant");
			if (v == "ours1931") { o.observer = 1931; o.upsampling = SSB_UPSAMPLE_OURS; }
			else if (v == "ours2006") { o.observer = 2006; o.upsampling = SSB_UPSAMPLE_OURS; }
			else if (v == "meng") { o.observer = 1931; o.upsampling = SSB_UPSAMPLE_MENG; }
			else if (v == "jh") { o.observer = 1931; o.upsampling = SSB_UPSAMPLE_JH; }
			else if (v == "rgb") { o.render_mode = SSB_RENDER_RGB; }  // the reference's RENDER_MODE_RGB build
			else { std::fprintf(stderr, "Unknown variant \"%s\"\n", v.c_str()); throw -3; }
		} catch (int code) { if (code != -2) throw; }
		try {  // SAMPLE_WAVELENGTHS (stdafx.hpp:90)
			int n = std::atoi(a.get("--wavelengths", "--wavelengths").c_str());
			if (n < 2 || n > 4) { std::fprintf(stderr, "Invalid number of wavelengths (2, 3 or 4)!\n"); throw -1; }
			o.n_wavelengths = static_cast<uint32_t>(n);
		} catch (int code) { if (code != -2) throw; }
		try { o.seed = std::strtoull(a.get("--seed", "--seed").c_str(), nullptr, 10); } catch (int code) { if (code != -2) throw; }
		try { o.device = std::atoi(a.get("--device", "--device").c_str()); } catch (int code) { if (code != -2) throw; }
		try {  // --devices=0-7 | --devices=0,2,5 | --devices=all: render on several GPUs of the box
			std::string v = a.get("--devices", "--devices");
			if (v == "all") {
				int n = 0;
				if (ssb_device_count(&n) != SSB_OK) { std::fprintf(stderr, "%s\n", ssb_last_error()); throw -1; }
				for (int d = 0; d < n; ++d) o.devices.push_back(d);
			} else {
				size_t pos = 0;
				while (pos <= v.size()) {
					size_t comma = v.find(',', pos);
					std::string item = v.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
					size_t dash = item.find('-');
					try {
						if (dash == std::string::npos) o.devices.push_back(std::stoi(item));
						else { int lo = std::stoi(item.substr(0, dash)), hi = std::stoi(item.substr(dash + 1)); if (hi < lo || hi - lo > 63) throw -1; for (int d = lo; d <= hi; ++d) o.devices.push_back(d); }
					} catch (...) { std::fprintf(stderr, "Invalid device list \"%s\"!\n", v.c_str()); throw -1; }
					if (comma == std::string::npos) break;
					pos = comma + 1;
				}
			}
		} catch (int code) { if (code != -2) throw; }
		try {
			std::string v = a.get("--shard", "--shard");
			if (v == "tiles") o.shard = ssbh::RendererOptions::SHARD_TILES;
			else if (v == "samples") o.shard = ssbh::RendererOptions::SHARD_SAMPLES;
			else { std::fprintf(stderr, "Unknown shard mode \"%s\" (tiles | samples)\n", v.c_str()); throw -1; }
		} catch (int code) { if (code != -2) throw; }
		try { o.data_root = a.get("--data-root", "--data-root"); } catch (int code) { if (code != -2) throw; }
		try { a.get("--prebake", "--prebake"); o.prebaked_textures = true; } catch (int code) { if (code != -2) throw; }
		try { a.get("--progressive", "--progressive"); o.progressive = true; } catch (int code) { if (code != -2) throw; }
		try { preview_path = a.get("--preview", "--preview"); o.progressive = true; } catch (int code) { if (code != -2) throw; }
		if (!a.rest.empty()) {
			std::fprintf(stderr, "Warning: ignoring extraneous argument(s):\n");
			for (auto const& s : a.rest) std::fprintf(stderr, "  \"%s\"\n", s.c_str());
		}
	} catch (int) {
		print_usage();
		return -1;
	}
	try {
		ssbh::Renderer renderer(o);
		renderer.render_start();
		// the reference's display loop (main.cpp:316-323), headless: poll, and write each new preview to a file
		if (!preview_path.empty()) {
			ssbh::Framebuffer shot;
			shot.res[0] = o.res[0]; shot.res[1] = o.res[1];
			uint32_t shown = 0;
			while (renderer.is_rendering()) {
				if (renderer.samples_done() != shown) { shown = renderer.snapshot(shot.pixels); shot.save(preview_path); }
				std::this_thread::sleep_for(std::chrono::milliseconds(2));
			}
		}
		renderer.render_wait();
		if (!preview_path.empty()) {  // the last slices may have finished between two polls: the preview ends on the final image
			ssbh::Framebuffer shot;
			shot.res[0] = o.res[0]; shot.res[1] = o.res[1];
			renderer.snapshot(shot.pixels);
			shot.save(preview_path);
		}
		if (renderer.last_stats.device_ms > 0)
			std::printf("%.3f Mpath-samples/s on the device%s (%llu samples, %.3f ms)\n",
			            renderer.last_stats.samples / renderer.last_stats.device_ms / 1e3, renderer.device_count() > 1 ? "s" : "",
			            (unsigned long long)renderer.last_stats.samples, renderer.last_stats.device_ms);
	} catch (ssbh::Error const& e) {
		std::fprintf(stderr, "%s\n", e.message.c_str());
		return e.code;
	} catch (std::exception const& e) {  // bad_alloc, length_error, ...: the reference would abort; report and fail
		std::fprintf(stderr, "%s\n", e.what());
		return -1;
	}
	return 0;
}
