// host_spectrum.cpp — _Spectrum and the CSV loader (reference src/spectrum.cpp).
#include <algorithm>
#include <cmath>
#include <fstream>
#include <set>
#include <sstream>

#include "ssb_host.hpp"

namespace ssbh {

Spectrum::Spectrum(float value, float lambda_min, float lambda_max)  // spectrum.cpp:11-13
	: Spectrum(std::vector<float>(2, value), lambda_min, lambda_max) {}

Spectrum::Spectrum(std::vector<float> const& d, float lo, float hi) : data(d), low(lo), high(hi) {  // spectrum.cpp:14-26
	if (d.size() < 2) throw Error{ -1, "Must have at-least two elements in sampled spectrum!" };
	float numer = high - low;
	float denom = static_cast<float>(d.size() - 1);
	delta_lambda = numer / denom;
	delta_lambda_recip = denom / numer;
}

float Spectrum::sample_nearest(float lambda) const {  // spectrum.cpp:29-38
	float i_f = (lambda - low) * delta_lambda_recip;
	i_f = std::round(i_f);
	int i_i = static_cast<int>(i_f);
	if (i_i >= 0 && static_cast<size_t>(i_i) < data.size()) return data[static_cast<size_t>(i_i)];
	return 0.0f;
}
float Spectrum::sample_linear(float lambda) const {  // spectrum.cpp:39-60
	float i = (lambda - low) * delta_lambda_recip;
	float i0f = std::floor(i);
	float frac = i - i0f;
	int i0 = static_cast<int>(i0f), i1 = i0 + 1;
	float val0 = (i0 >= 0 && static_cast<size_t>(i0) < data.size()) ? data[static_cast<size_t>(i0)] : 0.0f;
	float val1 = (i1 >= 0 && static_cast<size_t>(i1) < data.size()) ? data[static_cast<size_t>(i1)] : 0.0f;
	return val0 * (1.0f - frac) + val1 * frac;
}
Spectrum Spectrum::operator*(float sc) const {  // spectrum.cpp:69-73
	Spectrum r = *this;
	for (float& f : r.data) f *= sc;
	return r;
}
namespace {
template <class Op> Spectrum combine(Spectrum const& a, Spectrum const& b, Op op) {  // spectrum.cpp:74-118
	float lo = std::max(a.low, b.low), hi = std::min(a.high, b.high);
	// the reference asserts equal steps and aligned ranges ("other cases are valid, but not implemented")
	if (a.delta_lambda != b.delta_lambda || std::fmod(a.low - lo, a.delta_lambda) != 0.0f || std::fmod(b.low - lo, b.delta_lambda) != 0.0f ||
	    std::fmod(a.high - hi, a.delta_lambda) != 0.0f || std::fmod(b.high - hi, b.delta_lambda) != 0.0f)
		throw Error{ -3, "Spectrum arithmetic on spectra with different or misaligned sample grids is not implemented!" };
	std::vector<float> data(static_cast<size_t>((hi - lo) / a.delta_lambda + 1));
	for (size_t i = 0; i < data.size(); ++i) {
		float lambda = lo + a.delta_lambda * static_cast<float>(i);
		data[i] = op(a.sample_nearest(lambda), b.sample_nearest(lambda));
	}
	return Spectrum(data, lo, hi);
}
}  // namespace
Spectrum Spectrum::operator*(Spectrum const& other) const { return combine(*this, other, [](float x, float y) { return x * y; }); }
Spectrum Spectrum::operator+(Spectrum const& other) const { return combine(*this, other, [](float x, float y) { return x + y; }); }
float Spectrum::integrate(Spectrum const& spec) {  // spectrum.cpp:120-133
	float result = 0.0f;
	for (float v : spec.data) result += v;
	result *= spec.delta_lambda;
	return result;
}
float Spectrum::integrate(Spectrum const& spec0, Spectrum const& spec1) {  // spectrum.cpp:134-173
	float lo = std::max(spec0.low - spec0.delta_lambda, spec1.low - spec1.delta_lambda);
	float hi = std::min(spec0.high + spec0.delta_lambda, spec1.high + spec1.delta_lambda);
	std::set<float> pts;
	auto add = [&](Spectrum const& s) {
		float sample = s.low - s.delta_lambda;
		while (sample < lo) sample += s.delta_lambda;
		while (sample <= hi) { pts.insert(sample); sample += s.delta_lambda; }
	};
	add(spec0);
	add(spec1);
	std::vector<float> p(pts.begin(), pts.end());
	float result = 0.0f;
	for (size_t i = 0; i + 1 < p.size(); ++i) {
		float l0 = p[i], l1 = p[i + 1];
		float vallow = spec0.sample_linear(l0) * spec1.sample_linear(l0);
		float valhigh = spec0.sample_linear(l1) * spec1.sample_linear(l1);
		result += 0.5f * (vallow + valhigh) * (l1 - l0);
	}
	return result;
}
ssb_spectrum Spectrum::flat() const {
	ssb_spectrum s{};
	s.data = data.empty() ? nullptr : data.data();
	s.n = static_cast<uint32_t>(data.size());
	s.low = low; s.high = high; s.filter = filter;
	return s;
}

std::vector<std::vector<float>> load_spectral_data(std::string const& csv_path) {  // spectrum.cpp:177-213
	std::ifstream file(csv_path);
	if (!file.good()) throw Error{ -1, "Could not open required file \"" + csv_path + "\"!" };
	std::vector<std::vector<float>> data;
	std::string line;
	while (std::getline(file, line)) {
		std::istringstream ss(line);
		for (size_t i = 0;; ++i) {
			float f;
			if (ss >> f) {
				if (i == data.size()) data.emplace_back();
				data[i].push_back(f);
			} else {
				throw Error{ -2, "Expected number when parsing file!" };
			}
			char c;
			if (!(ss >> c)) break;
		}
	}
	for (size_t i = 1; i < data.size(); ++i)
		if (data[i].size() != data[0].size()) throw Error{ -3, "Data dimension mismatch in file!" };
	return data;
}

}  // namespace ssbh
