// host_color.cpp — Color::init (reference src/util/color.cpp:26-155): observer, D65 (radiometric scaling via
// Planck), basis / JH / Meng tables, RGB<->XYZ matrices.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "host_math.hpp"
#include "ssb_host.hpp"

namespace ssbh {

namespace {

// Constants:: (stdafx.hpp:186-204), each rounded to float from the long-double literal
const float kH = static_cast<float>(6.62607015e-34L);
const float kC = static_cast<float>(299792458.0L);
const float kKB = static_cast<float>(1.38064852e-23L);

float planck(float lambda_nm, float temp) {  // color.cpp:49-66
	float lambda_m = lambda_nm * 1.0e-9f;
	float c_1L = 2.0f * kH * kC * kC;
	float c_2 = kH * kC / kKB;
	float numer = c_1L;
	float denom = std::pow(lambda_m, 5.0f) * (std::exp(c_2 / (lambda_m * temp)) - 1.0f);
	float value = numer / denom;
	return value * 1.0e-9f;
}

void specradflux_to_ciexyz(ColorData const& d, Spectrum const& flux, float out[3]) {  // color.hpp:104-109
	out[0] = Spectrum::integrate(flux, d.std_obs_xbar);
	out[1] = Spectrum::integrate(flux, d.std_obs_ybar);
	out[2] = Spectrum::integrate(flux, d.std_obs_zbar);
}

M3 calc_matr_rgb_to_xyz(float xr, float yr, float xg, float yg, float xb, float yb, F3 XYZ_W) {  // color.cpp:26-46
	F3 x_rgb = f3(xr, xg, xb), y_rgb = f3(yr, yg, yb);
	F3 X_rgb = x_rgb / y_rgb;
	F3 Y_rgb = f3(1, 1, 1);
	F3 Z_rgb = (f3(1, 1, 1) - x_rgb - y_rgb) / y_rgb;
	F3 S_rgb = inverse(transpose(m3_cols(X_rgb, Y_rgb, Z_rgb))) * XYZ_W;
	return transpose(m3_cols(S_rgb * X_rgb, S_rgb * Y_rgb, S_rgb * Z_rgb));
}

std::vector<unsigned char> read_file(std::string const& path) {
	FILE* f = std::fopen(path.c_str(), "rb");
	if (!f) throw Error{ -1, "Could not open required file \"" + path + "\"!" };
	std::fseek(f, 0, SEEK_END);
	long n = std::ftell(f);
	std::fseek(f, 0, SEEK_SET);
	std::vector<unsigned char> b(static_cast<size_t>(n));
	size_t got = n ? std::fread(b.data(), 1, b.size(), f) : 0;
	std::fclose(f);
	if (got != b.size()) throw Error{ -1, "Short read on \"" + path + "\"" };
	return b;
}

}  // namespace

ColorData color_init(std::string const& data_root, int observer, uint32_t upsampling) {
	if (observer != 1931 && observer != 2006) throw Error{ -3, "CIE_OBSERVER must be 1931 or 2006" };
	if (upsampling < SSB_UPSAMPLE_OURS || upsampling > SSB_UPSAMPLE_JH) throw Error{ -3, "unknown upsampling mode" };
	if (upsampling != SSB_UPSAMPLE_OURS && observer != 1931)  // stdafx.hpp:106-108
		throw Error{ -3, "Only our algorithm currently implements support for the newest CIE standard observer!" };
	std::string const base = data_root + "/data/";
	ColorData d;
	d.observer = observer;
	d.upsampling = upsampling;
	if (observer == 1931) { d.lambda_min = 380.0f; d.lambda_max = 780.0f; }  // stdafx.hpp:115-121
	else { d.lambda_min = 390.0f; d.lambda_max = 830.0f; }

	{  // observer (color.cpp:77-99)
		bool o31 = observer == 1931;
		auto tmp = load_spectral_data(base + (o31 ? "cie1931-xyzbar-380+5+780.csv" : "cie2006-xyzbar-390+1+830.csv"));
		if (tmp.size() != 3) throw Error{ -1, "Invalid data in file!" };
		float lo = o31 ? 380.0f : 390.0f, hi = o31 ? 780.0f : 830.0f;
		d.std_obs_xbar = Spectrum(tmp[0], lo, hi);
		d.std_obs_ybar = Spectrum(tmp[1], lo, hi);
		d.std_obs_zbar = Spectrum(tmp[2], lo, hi);
	}
	{  // D65 (color.cpp:102-120)
		auto tmp = load_spectral_data(base + "d65-300+5+780.csv");
		if (tmp.size() != 1) throw Error{ -1, "Invalid data in file!" };
		d.D65_orig = Spectrum(tmp[0], 300.0f, 780.0f);
		specradflux_to_ciexyz(d, d.D65_orig, d.D65_orig_XYZ);
		float temp_d65 = 6500.0f;
		temp_d65 *= (kH * kC / kKB) / 1.438e-2f;
		float scalar = 0.00001f * planck(560.0f, temp_d65);
		d.D65_rad = d.D65_orig * scalar;
		specradflux_to_ciexyz(d, d.D65_rad, d.D65_rad_XYZ);
	}
	if (upsampling == SSB_UPSAMPLE_OURS) {  // color.cpp:122-141
		bool o31 = observer == 1931;
		auto tmp = load_spectral_data(base + (o31 ? "cie1931-basis-bt709-380+5+780.csv" : "cie2006-basis-bt709-390+1+780.csv"));
		if (tmp.size() != 3) throw Error{ -1, "Invalid data in file!" };
		float lo = o31 ? 380.0f : 390.0f, hi = 780.0f;
		d.basis_r = Spectrum(tmp[0], lo, hi);
		d.basis_g = Spectrum(tmp[1], lo, hi);
		d.basis_b = Spectrum(tmp[2], lo, hi);
	} else if (upsampling == SSB_UPSAMPLE_JH) {  // rgb2spec_load, rgb2spec.c:10-47
		auto b = read_file(base + "jakob-and-hanika-2019-srgb.coeff");
		if (b.size() < 8 || std::memcmp(b.data(), "SPEC", 4) != 0) throw Error{ -1, "Invalid JH coefficient file" };
		uint32_t res;
		std::memcpy(&res, b.data() + 4, 4);
		size_t nd = static_cast<size_t>(res) * res * res * 3 * 3;
		if (b.size() < 8 + 4 * (res + nd)) throw Error{ -1, "Truncated JH coefficient file" };
		d.jh_res = res;
		d.jh_scale.resize(res);
		d.jh_data.resize(nd);
		std::memcpy(d.jh_scale.data(), b.data() + 8, 4 * static_cast<size_t>(res));
		std::memcpy(d.jh_data.data(), b.data() + 8 + 4 * static_cast<size_t>(res), 4 * nd);
	} else {  // Meng tables (tools/gen_meng_tables.c format)
		auto b = read_file(base + "meng-et-al-2015-tables.bin");
		if (b.size() < 56 || std::memcmp(b.data(), "SSBMENG1", 8) != 0) throw Error{ -1, "Invalid Meng table file" };
		uint32_t hdr[4];
		std::memcpy(hdr, b.data() + 8, 16);
		d.meng.grid_w = hdr[0]; d.meng.grid_h = hdr[1]; d.meng.npoints = hdr[2]; d.meng.nsamples = hdr[3];
		std::memcpy(d.meng.xy_to_uv, b.data() + 24, 24);
		std::memcpy(&d.meng.sample_min, b.data() + 48, 4);
		std::memcpy(&d.meng.sample_max, b.data() + 52, 4);
		size_t ng = static_cast<size_t>(hdr[0]) * hdr[1] * 8, np = static_cast<size_t>(hdr[2]) * (5 + hdr[3]);
		if (b.size() < 56 + 4 * (ng + np)) throw Error{ -1, "Truncated Meng table file" };
		d.meng_grid.resize(ng);
		d.meng_points.resize(np);
		std::memcpy(d.meng_grid.data(), b.data() + 56, 4 * ng);
		std::memcpy(d.meng_points.data(), b.data() + 56 + 4 * ng, 4 * np);
		d.have_meng = true;
	}
	{  // color.cpp:146-154
		M3 M = calc_matr_rgb_to_xyz(0.64f, 0.33f, 0.30f, 0.60f, 0.15f, 0.06f, f3(d.D65_rad_XYZ[0], d.D65_rad_XYZ[1], d.D65_rad_XYZ[2]));
		M3 Mi = inverse(M);
		std::memcpy(d.matr_lrgb_to_xyz, M.m, sizeof(M.m));
		std::memcpy(d.matr_xyz_to_lrgb, Mi.m, sizeof(Mi.m));
	}
	return d;
}

ssb_color ColorData::flat() const {
	ssb_color c{};
	c.xbar = std_obs_xbar.flat(); c.ybar = std_obs_ybar.flat(); c.zbar = std_obs_zbar.flat();
	if (!basis_r.data.empty()) { c.basis_r = basis_r.flat(); c.basis_g = basis_g.flat(); c.basis_b = basis_b.flat(); }
	std::memcpy(c.xyz_to_lrgb, matr_xyz_to_lrgb, sizeof(c.xyz_to_lrgb));
	c.d65_rad_Y = D65_rad_XYZ[1];
	if (jh_res) { c.jh_scale = jh_scale.data(); c.jh_data = jh_data.data(); c.jh_res = jh_res; }
	if (have_meng) {
		ssb_meng_tables* m = const_cast<ssb_meng_tables*>(&meng);
		m->grid = meng_grid.data();
		m->points = meng_points.data();
		c.meng = m;
	}
	return c;
}

// ---- the reference's round-trip self-test (color.cpp:259-294, main.cpp:184-264)
namespace {
float srgb_to_lrgb_c(float c) { return c < 0.04045f ? c / 12.92f : std::pow((c + 0.055f) / 1.055f, 2.4f); }            // color.hpp:91-97
float lrgb_to_srgb_c(float c) { return c < 0.0031308f ? 12.92f * c : 1.055f * std::pow(c, 1.0f / 2.4f) - 0.055f; }    // color.hpp:84-90
}  // namespace

void ColorData::round_trip_lrgb(const float lrgb[3], float out[3]) const {
	if (basis_r.data.empty()) throw Error{ -3, "round_trip_lrgb needs the basis spectra (RENDER_MODE_SPECTRAL_OURS)" };
	Spectrum reflectance = basis_r * lrgb[0] + basis_g * lrgb[1] + basis_b * lrgb[2];
	Spectrum flux = D65_rad * reflectance;  // FLAT_FIELD_CORRECTION: flux = radiance (color.cpp:275-279)
	float xyz[3] = { Spectrum::integrate(flux, std_obs_xbar), Spectrum::integrate(flux, std_obs_ybar), Spectrum::integrate(flux, std_obs_zbar) };  // color.hpp:106-111
	const float* m = matr_xyz_to_lrgb;  // column-major; glm mat3 * vec3
	for (int r = 0; r < 3; ++r) out[r] = m[0 + r] * xyz[0] + m[3 + r] * xyz[1] + m[6 + r] * xyz[2];
}
void ColorData::round_trip_srgb(const float srgb[3], float out[3]) const {
	float lin[3] = { srgb_to_lrgb_c(srgb[0]), srgb_to_lrgb_c(srgb[1]), srgb_to_lrgb_c(srgb[2]) }, back[3];
	round_trip_lrgb(lin, back);
	for (int c = 0; c < 3; ++c) out[c] = lrgb_to_srgb_c(back[c]);
}
void ColorData::round_trip_running_max(uint32_t r_begin, uint32_t r_end, float start_max, float* running_max, unsigned threads) const {
	if (threads == 0) threads = 1;
	float max_error = start_max;
	for (uint32_t r = r_begin; r < r_end; ++r) {
		std::vector<float> part(threads, 0.0f);
		std::vector<Error> errors(threads, Error{ 0, "" });
		std::vector<std::thread> pool;
		for (unsigned t = 0; t < threads; ++t)
			pool.emplace_back([&, t] {
				try {
					for (uint32_t g = t; g <= 255; g += threads)
						for (uint32_t b = 0; b <= 255; ++b) {
							// sRGB_F32(u8 r, u8 g, u8 b) * (1.0f/255.0f)   (main.cpp:250-255)
							float in[3] = { static_cast<float>(r) * (1.0f / 255.0f), static_cast<float>(g) * (1.0f / 255.0f), static_cast<float>(b) * (1.0f / 255.0f) }, out[3];
							round_trip_srgb(in, out);
							for (int c = 0; c < 3; ++c) part[t] = std::max(part[t], std::fabs(out[c] - in[c]));
						}
				} catch (Error const& e) { errors[t] = e; }
			});
		for (auto& th : pool) th.join();
		for (unsigned t = 0; t < threads; ++t) {
			if (errors[t].code != 0) throw errors[t];
			max_error = std::max(max_error, part[t]);
		}
		running_max[r - r_begin] = max_error;
	}
}

}  // namespace ssbh
