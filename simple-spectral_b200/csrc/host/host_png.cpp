// host_png.cpp — PNG read/write for the texture (material.cpp:10-29) and the output image
// (framebuffer.cpp:140-175).  The reference uses its vendored lodepng; this is an independent, minimal
// implementation on top of zlib (inflate/deflate + CRC): 8-bit gray / RGB / palette / gray+alpha / RGBA,
// non-interlaced, converted to RGB8 on load; RGBA8 on save.
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ssb_host.hpp"

namespace ssbh {

namespace {

uint32_t be32(const unsigned char* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | uint32_t(p[3]); }
void put_be32(std::vector<unsigned char>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }

int paeth(int a, int b, int c) {
	int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
	if (pa <= pb && pa <= pc) return a;
	if (pb <= pc) return b;
	return c;
}

}  // namespace

Texture load_png_rgb8(std::string const& path) {
	auto fail = [&]() -> Texture { throw Error{ -1, "Could not load texture \"" + path + "\"" }; };  // material.cpp:15-18
	FILE* f = std::fopen(path.c_str(), "rb");
	if (!f) return fail();
	std::fseek(f, 0, SEEK_END);
	long n = std::ftell(f);
	std::fseek(f, 0, SEEK_SET);
	std::vector<unsigned char> b(static_cast<size_t>(n > 0 ? n : 0));
	size_t got = b.empty() ? 0 : std::fread(b.data(), 1, b.size(), f);
	std::fclose(f);
	static const unsigned char sig[8] = { 137, 80, 78, 71, 13, 10, 26, 10 };
	if (got != b.size() || b.size() < 33 || std::memcmp(b.data(), sig, 8) != 0) return fail();
	uint32_t w = 0, h = 0;
	int depth = 0, ctype = 0, interlace = 0;
	std::vector<unsigned char> idat, palette;
	size_t pos = 8;
	bool end = false;
	while (!end && pos + 12 <= b.size()) {
		uint32_t len = be32(&b[pos]);
		const unsigned char* type = &b[pos + 4];
		const unsigned char* data = &b[pos + 8];
		if (pos + 12 + static_cast<size_t>(len) > b.size()) return fail();
		if (!std::memcmp(type, "IHDR", 4)) {
			if (len < 13) return fail();
			w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
		} else if (!std::memcmp(type, "PLTE", 4)) palette.assign(data, data + len);
		else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
		else if (!std::memcmp(type, "IEND", 4)) end = true;
		pos += 12 + static_cast<size_t>(len);
	}
	int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
	// Supported subset: 8 bits per channel, non-interlaced (what the reference's scenes ship; lodepng also reads 16-bit and
	// Adam7 files).  The dimensions are bounded so that the size products below cannot overflow or exhaust memory on a
	// crafted header.
	if (w == 0 || h == 0 || w > 65536u || h > 65536u || depth != 8 || channels == 0 || interlace != 0) return fail();
	size_t stride = static_cast<size_t>(w) * channels;
	std::vector<unsigned char> raw((stride + 1) * h);
	uLongf rawlen = static_cast<uLongf>(raw.size());
	if (uncompress(raw.data(), &rawlen, idat.data(), static_cast<uLong>(idat.size())) != Z_OK || rawlen != raw.size()) return fail();
	// undo the scanline filters in place
	std::vector<unsigned char> img(stride * h);
	for (uint32_t y = 0; y < h; ++y) {
		const unsigned char* src = &raw[(stride + 1) * y];
		unsigned char* dst = &img[stride * y];
		const unsigned char* up = y ? &img[stride * (y - 1)] : nullptr;
		int ft = src[0];
		for (size_t x = 0; x < stride; ++x) {
			int a = x >= static_cast<size_t>(channels) ? dst[x - channels] : 0;
			int bb = up ? up[x] : 0;
			int c = (up && x >= static_cast<size_t>(channels)) ? up[x - channels] : 0;
			int v = src[1 + x];
			switch (ft) {
				case 0: break;
				case 1: v += a; break;
				case 2: v += bb; break;
				case 3: v += (a + bb) >> 1; break;
				case 4: v += paeth(a, bb, c); break;
				default: return fail();
			}
			dst[x] = static_cast<unsigned char>(v);
		}
	}
	Texture t;
	t.width = w; t.height = h;
	t.rgb8.resize(static_cast<size_t>(w) * h * 3);
	for (size_t i = 0; i < static_cast<size_t>(w) * h; ++i) {
		unsigned char r, g, bl;
		const unsigned char* p = &img[i * channels];
		if (ctype == 2 || ctype == 6) { r = p[0]; g = p[1]; bl = p[2]; }
		else if (ctype == 3) { size_t k = static_cast<size_t>(p[0]) * 3; if (k + 3 > palette.size()) return fail(); r = palette[k]; g = palette[k + 1]; bl = palette[k + 2]; }
		else { r = g = bl = p[0]; }
		t.rgb8[3 * i] = r; t.rgb8[3 * i + 1] = g; t.rgb8[3 * i + 2] = bl;
	}
	return t;
}

// lodepng::encode(path, rgba, w, h, LCT_RGBA) equivalent (framebuffer.cpp:167-172)
void save_png_rgba8(std::string const& path, const unsigned char* rgba, uint32_t w, uint32_t h) {
	size_t stride = static_cast<size_t>(w) * 4;
	std::vector<unsigned char> raw((stride + 1) * h);
	for (uint32_t y = 0; y < h; ++y) {
		raw[(stride + 1) * y] = 0;
		std::memcpy(&raw[(stride + 1) * y + 1], rgba + stride * y, stride);
	}
	uLongf clen = compressBound(static_cast<uLong>(raw.size()));
	std::vector<unsigned char> comp(clen);
	if (compress2(comp.data(), &clen, raw.data(), static_cast<uLong>(raw.size()), 6) != Z_OK) throw Error{ -1, "PNG compression failed" };
	std::vector<unsigned char> out = { 137, 80, 78, 71, 13, 10, 26, 10 };
	auto chunk = [&](const char* type, const unsigned char* data, size_t len) {
		put_be32(out, static_cast<uint32_t>(len));
		size_t start = out.size();
		out.insert(out.end(), type, type + 4);
		out.insert(out.end(), data, data + len);
		put_be32(out, static_cast<uint32_t>(crc32(0L, &out[start], static_cast<uInt>(len + 4))));
	};
	unsigned char ihdr[13];
	ihdr[0] = w >> 24; ihdr[1] = w >> 16; ihdr[2] = w >> 8; ihdr[3] = w;
	ihdr[4] = h >> 24; ihdr[5] = h >> 16; ihdr[6] = h >> 8; ihdr[7] = h;
	ihdr[8] = 8; ihdr[9] = 6; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
	chunk("IHDR", ihdr, 13);
	chunk("IDAT", comp.data(), clen);
	chunk("IEND", nullptr, 0);
	FILE* f = std::fopen(path.c_str(), "wb");
	if (!f) throw Error{ -1, "Could not open \"" + path + "\" for writing" };
	std::fwrite(out.data(), 1, out.size(), f);
	std::fclose(f);
}

}  // namespace ssbh
