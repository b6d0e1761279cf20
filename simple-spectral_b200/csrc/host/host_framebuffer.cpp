// host_framebuffer.cpp — Framebuffer (reference src/framebuffer.cpp): float4 sRGBA store, bottom-to-top, and
// the writers selected by file extension: .csv / .hdr (Radiance RGBE) / .pfm / anything else = PNG.
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "ssb_host.hpp"

namespace ssbh {

void save_png_rgba8(std::string const& path, const unsigned char* rgba, uint32_t w, uint32_t h);

namespace {
bool endswith(std::string const& s, std::string const& e) { return s.size() >= e.size() && s.compare(s.size() - e.size(), e.size(), e) == 0; }
float srgb_to_lrgb_1(float c) { return c < 0.04045f ? c / 12.92f : std::pow((c + 0.055f) / 1.055f, 2.4f); }  // color.hpp:91-97
}  // namespace

void Framebuffer::reset(uint32_t w, uint32_t h) {  // framebuffer.cpp:9-33 (TILE_SIZE 8 checkerboard, alpha 1)
	res[0] = w; res[1] = h;
	pixels.assign(static_cast<size_t>(w) * h * 4, 0.0f);
	for (uint32_t j = 0; j < h; ++j)
		for (uint32_t i = 0; i < w; ++i) {
			float v = (((i / 8) ^ (j / 8)) % 2 == 0) ? 0.7f : 0.3f;
			float* p = &pixels[(static_cast<size_t>(j) * w + i) * 4];
			p[0] = p[1] = p[2] = v; p[3] = 1.0f;
		}
}

void Framebuffer::save(std::string const& path) const {
	const uint32_t W = res[0], H = res[1];
	auto px = [&](uint32_t i, uint32_t j) { return &pixels[(static_cast<size_t>(j) * W + i) * 4]; };
	if (endswith(path, ".csv")) {  // framebuffer.cpp:40-63 (rows bottom-to-top as stored)
		FILE* f = std::fopen(path.c_str(), "wb");
		if (!f) throw Error{ -1, "Could not open \"" + path + "\"" };
		for (uint32_t j = 0; j < H; ++j)
			for (uint32_t i = 0;; ++i) {
				const float* s = px(i, j);
				std::fprintf(f, "%g,%g,%g", (double)srgb_to_lrgb_1(s[0]), (double)srgb_to_lrgb_1(s[1]), (double)srgb_to_lrgb_1(s[2]));
				if (i < W - 1) std::fputc(',', f);
				else { std::fputc('\n', f); break; }
			}
		std::fclose(f);
	} else if (endswith(path, ".hdr")) {  // framebuffer.cpp:64-112
		FILE* f = std::fopen(path.c_str(), "wb");
		if (!f) throw Error{ -1, "Could not open \"" + path + "\"" };
		std::fprintf(f, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0\nSOFTWARE=simple-spectral\n\n-Y %zu +X %zu\n", (size_t)H, (size_t)W);
		for (uint32_t j = 0; j < H; ++j)
			for (uint32_t i = 0; i < W; ++i) {
				const float* s = px(i, H - 1 - j);
				float l[3] = { srgb_to_lrgb_1(s[0]), srgb_to_lrgb_1(s[1]), srgb_to_lrgb_1(s[2]) };
				float v = std::max(l[0], std::max(l[1], l[2]));
				if (v < 1.0e-32f) { uint32_t zero = 0u; std::fwrite(&zero, 4, 1, f); }
				else {
					int e;
					v = std::frexp(v, &e) * 256.0f / v;
					e += 128;
					unsigned char d[4];
					for (int c = 0; c < 3; ++c) {
						int q = static_cast<int>(std::round(std::round(l[c] * v)));
						d[c] = static_cast<unsigned char>(std::min(std::max(q, 0), 255));
					}
					d[3] = static_cast<unsigned char>(e);
					std::fwrite(d, 1, 4, f);
				}
			}
		std::fclose(f);
	} else if (endswith(path, ".pfm")) {  // framebuffer.cpp:113-139
		FILE* f = std::fopen(path.c_str(), "wb");
		if (!f) throw Error{ -1, "Could not open \"" + path + "\"" };
		std::fprintf(f, "PF\n%zu %zu\n-1.0\n", (size_t)W, (size_t)H);
		for (uint32_t j = 0; j < H; ++j)
			for (uint32_t i = 0; i < W; ++i) {
				const float* s = px(i, H - 1 - j);
				float l[3] = { srgb_to_lrgb_1(s[0]), srgb_to_lrgb_1(s[1]), srgb_to_lrgb_1(s[2]) };
				std::fwrite(l, sizeof(float), 3, f);
			}
		std::fclose(f);
	} else {  // PNG, framebuffer.cpp:140-175: clip(255*srgba, 0, 255), round, vertical flip
		std::vector<unsigned char> out(static_cast<size_t>(W) * H * 4);
		for (uint32_t j = 0; j < H; ++j)
			for (uint32_t i = 0; i < W; ++i) {
				const float* s = px(i, j);
				unsigned char* d = &out[(static_cast<size_t>(H - 1 - j) * W + i) * 4];
				for (int c = 0; c < 4; ++c) {
					float v = 255.0f * s[c];
					v = (v < 0.0f) ? 0.0f : v;      // glm::clamp = min(max(x,lo),hi)
					v = (255.0f < v) ? 255.0f : v;
					d[c] = static_cast<unsigned char>(std::round(v));
				}
			}
		save_png_rgba8(path, out.data(), W, H);
	}
}

}  // namespace ssbh
