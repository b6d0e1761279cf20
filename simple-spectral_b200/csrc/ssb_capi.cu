// ssb_capi.cu — C ABI (include/ssb200.h) of the B200 spectral path-tracing integrator.
// Host side: deep-copies the caller's flat scene / colour tables, packs them into the device
// "blob" the trace kernel stages into shared memory, owns the device buffers and the stream, and
// launches the kernels of ssb_kernels.cuh.  There is NO CPU fallback: every entry point that
// computes fails with SSB_ERR_DATA when CUDA is unavailable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ssb200.h"
#include "ssb_kernels.cuh"

using namespace ssbk;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_last_error = buf;
	return code;
}

#define SSB_CUDA(expr)                                                                                  \
	do {                                                                                                \
		cudaError_t err__ = (expr);                                                                     \
		if (err__ != cudaSuccess)                                                                       \
			return fail(SSB_ERR_DATA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(err__), __FILE__, __LINE__, #expr); \
	} while (0)

struct HostSpectrum {
	std::vector<float> data;
	float low = 0, high = 0;
	uint32_t filter = 0;
	bool present = false;
};
struct HostMaterial {
	uint32_t kind = 0, albedo_mode = 0, texture = 0;
	HostSpectrum albedo, emission;
	float albedo_rgb[3] = { 1, 1, 1 }, emission_rgb[3] = { 0, 0, 0 };  // RENDER_MODE_RGB constants
};

int copy_spectrum(const ssb_spectrum& s, HostSpectrum& out, const char* what, bool required) {
	out = HostSpectrum();
	if (!s.data || s.n == 0) {
		if (required) return fail(SSB_ERR_ARG, "%s: missing spectrum", what);
		return SSB_OK;
	}
	if (s.n < 2) return fail(SSB_ERR_DATA, "%s: must have at-least two elements in sampled spectrum", what);  // spectrum.cpp:17-20
	if (!(s.high > s.low)) return fail(SSB_ERR_ARG, "%s: high must exceed low", what);
	if (s.filter > SSB_FILTER_NEAREST) return fail(SSB_ERR_ARG, "%s: unknown filter", what);
	out.data.assign(s.data, s.data + s.n);
	out.low = s.low; out.high = s.high; out.filter = s.filter; out.present = true;
	return SSB_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct ssb_ctx;
namespace {
// Host -> device copy ordered on the context's stream, complete on return.  (Not cudaMemcpy: for pageable host memory
// that returns once the data is STAGED, "the DMA to final destination may not have completed", and the context's
// stream is non-blocking, i.e. not ordered after the legacy default stream cudaMemcpy works on.  A kernel launched
// right after could read the tail of the previous contents — observed on the B200 in ssb_debug_eval_math: the last
// partial block of a batch evaluated the previous batch's arguments, profiles/r4d_diag_eval_math_race.txt.)
cudaError_t copy_to_device(ssb_ctx* c, void* dst, const void* src, size_t bytes);
}  // namespace

struct ssb_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;      // the stream every kernel / copy of this context is issued on
	cudaStream_t own_stream = nullptr;  // created by ssb_create; `stream` may be replaced by ssb_set_stream
	cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
	// ssb_upload_scene_async: texture copies run on their own stream and overlap the first kernels of the next render
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t ev_tex_ready = nullptr, ev_tex_free = nullptr;  // copy finished / last render that read the textures finished
	cudaEvent_t ev_accum_ready = nullptr;  // ssb_accum_merge: the source context's accumulator is complete
	double* d_peer_staging = nullptr;      // ssb_accum_merge without peer access: the source accumulator, copied over
	size_t peer_staging_count = 0;
	bool tex_pending = false, tex_in_use = false;
	std::vector<cudaEvent_t> ev_pass;   // (t0,t1) pairs around each trace-kernel launch of the last ssb_render
	uint32_t passes = 0;
	bool stats_pending = false;
	int sm_count = 0;

	// host copies of the uploaded tables
	bool have_scene = false, have_color = false, blob_dirty = true;
	ssb_camera camera{};
	std::vector<ssb_quad> quads;
	std::vector<HostMaterial> materials;
	std::vector<uint32_t> lights;
	HostSpectrum xbar, ybar, zbar, basis_r, basis_g, basis_b;
	float xyz_to_lrgb[9] = { 0 };
	float d65_rad_Y = 1.0f;
	uint32_t jh_res = 0;
	ssb_meng_tables meng{};
	bool have_meng = false;

	// device
	std::vector<uchar4*> d_textures;
	std::vector<uint32_t> tex_w, tex_h;
	// ssb_options.prebaked_textures: per-texture Jakob-Hanika coefficient texels, baked on first use by ssb_render
	std::vector<float4*> d_tex_coef;
	std::vector<size_t> tex_coef_n;   // texels each buffer was allocated for
	bool coef_valid = false;          // the buffers hold the coefficients of the current textures and JH tables
	unsigned char* d_rgb_staging = nullptr;
	size_t rgb_staging_capacity = 0;
	unsigned char* d_blob = nullptr;
	size_t blob_bytes = 0, blob_capacity = 0;
	float* d_jh_scale = nullptr;
	float* d_jh_data = nullptr;
	int32_t* d_meng_grid = nullptr;
	float* d_meng_points = nullptr;
	double* d_accum = nullptr;
	uint32_t accum_w = 0, accum_h = 0;
	// wavefront buffers of one pass (see KParams)
	unsigned char* d_wave = nullptr;
	size_t wave_bytes = 0;
	float4* d_samples = nullptr;      // per-sample (X,Y,Z,hit) of the last pass: alias into d_wave (see ssb_render)
	size_t samples_capacity = 0;      // in float4
	uint32_t* d_counts = nullptr;     // queue lengths per depth
	double* d_xyza = nullptr;
	float4* d_srgba = nullptr;
	size_t resolve_capacity = 0;  // pixels

	ssb_stats stats{};
	// cached launch configuration (see ssb_render)
	const void* occ_key_kernel = nullptr;
	size_t occ_key_smem = 0;
	int occ[4] = { 0, 0, 0, 0 };
};

namespace {

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (device, kernel) shared by every context of the process:
// it is only ever RAISED here, so that a context with small tables cannot lower it under one with large tables.
cudaError_t raise_dynamic_smem_limit(int device, const void* kernel, size_t bytes) {
	static std::mutex mu;
	static std::map<std::pair<int, const void*>, size_t> limit;
	std::lock_guard<std::mutex> lock(mu);
	size_t& cur = limit[std::make_pair(device, kernel)];
	if (bytes <= cur) return cudaSuccess;
	cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
	if (e == cudaSuccess) cur = bytes;
	return e;
}

cudaError_t copy_to_device(ssb_ctx* c, void* dst, const void* src, size_t bytes) {
	cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream);
	return e != cudaSuccess ? e : cudaStreamSynchronize(c->stream);
}

[[maybe_unused]] void free_textures(ssb_ctx* c) {
	for (uchar4* p : c->d_textures) cudaFree(p);
	c->d_textures.clear(); c->tex_w.clear(); c->tex_h.clear();
	for (float4* p : c->d_tex_coef) cudaFree(p);
	c->d_tex_coef.clear(); c->tex_coef_n.clear(); c->coef_valid = false;
}

DevSpectrum pack_spectrum(const HostSpectrum& s, std::vector<float>& pool) {
	DevSpectrum d{};
	if (!s.present) { d.offset = 0; d.n_filter = 0; d.low = 0; d.recip = 0; return d; }
	d.offset = (uint32_t)pool.size();
	pool.insert(pool.end(), s.data.begin(), s.data.end());
	uint32_t n = (uint32_t)s.data.size();
	d.n_filter = n | (s.filter == SSB_FILTER_NEAREST ? 0x80000000u : 0u);
	d.low = s.low;
	float numer = s.high - s.low;        // spectrum.cpp:22-25
	float denom = (float)(n - 1);
	d.recip = denom / numer;
	return d;
}

// Build the shared-memory image: [DevHeader][quads][materials][lights][textures][float pool]
int build_blob(ssb_ctx* c) {
	if (!c->have_scene) return fail(SSB_ERR_ARG, "ssb_render: no scene uploaded");
	// (colour tables may be absent: RGB mode needs none; ssb_render checks what the chosen mode needs)
	std::vector<float> pool;
	DevHeader hdr{};
	hdr.nquads = (uint32_t)c->quads.size();
	hdr.nmaterials = (uint32_t)c->materials.size();
	hdr.nlights = (uint32_t)c->lights.size();
	hdr.ntextures = (uint32_t)c->d_textures.size();
	hdr.xbar = pack_spectrum(c->xbar, pool);
	hdr.ybar = pack_spectrum(c->ybar, pool);
	hdr.zbar = pack_spectrum(c->zbar, pool);
	hdr.basis_r = pack_spectrum(c->basis_r, pool);
	hdr.basis_g = pack_spectrum(c->basis_g, pool);
	hdr.basis_b = pack_spectrum(c->basis_b, pool);
	for (int v = 0; v < 256; ++v) {  // Color::srgb_to_lrgb (color.hpp:91-97) of texel/255 (material.cpp:51-55)
		float srgb = (float)v * (1.0f / 255.0f);
		hdr.srgb_lut[v] = srgb < 0.04045f ? srgb / 12.92f : powf((srgb + 0.055f) / 1.055f, 2.4f);
	}
	std::vector<DevMaterial> mats(c->materials.size());
	for (size_t m = 0; m < mats.size(); ++m) {
		const HostMaterial& hm = c->materials[m];
		mats[m].kind = hm.kind; mats[m].albedo_mode = hm.albedo_mode; mats[m].texture = hm.texture; mats[m].pad = 0;
		mats[m].albedo = pack_spectrum(hm.albedo, pool);
		mats[m].emission = pack_spectrum(hm.emission, pool);
		for (int k = 0; k < 3; ++k) { mats[m].albedo_rgb[k] = hm.albedo_rgb[k]; mats[m].emission_rgb[k] = hm.emission_rgb[k]; }
		mats[m].albedo_rgb[3] = mats[m].emission_rgb[3] = 0.0f;
	}
	std::vector<DevTexture> texs(c->d_textures.size());
	for (size_t t = 0; t < texs.size(); ++t) {
		texs[t].rgba = c->d_textures[t]; texs[t].width = c->tex_w[t]; texs[t].height = c->tex_h[t];
		texs[t].coef = t < c->d_tex_coef.size() ? c->d_tex_coef[t] : nullptr;
	}

	size_t off = align_up(sizeof(DevHeader), 16);
	hdr.off_quads = (uint32_t)off; off = align_up(off + c->quads.size() * sizeof(ssb_quad), 16);
	hdr.off_materials = (uint32_t)off; off = align_up(off + mats.size() * sizeof(DevMaterial), 16);
	hdr.off_lights = (uint32_t)off; off = align_up(off + c->lights.size() * sizeof(uint32_t), 16);
	hdr.off_textures = (uint32_t)off; off = align_up(off + texs.size() * sizeof(DevTexture), 16);
	hdr.off_pool = (uint32_t)off; off = align_up(off + pool.size() * sizeof(float), 16);
	// conservative filter tables of scene_intersect (ssb_blob.hpp / ssb_isect.cuh)
	const FilterTables ft = build_filter_tables(c->quads.data(), c->quads.size(), c->camera.pos);
	ft.fill_header(hdr);
	off = align_up(off, 128);
	hdr.off_fpairs = (uint32_t)off; off = align_up(off + ft.pairs.size() * sizeof(float), 16);
	hdr.off_planes = (uint32_t)off; off = align_up(off + ft.planes.size() * sizeof(float), 16);
	hdr.off_entry_quad = (uint32_t)off; off = align_up(off + ft.entry_quad.size() * sizeof(uint32_t), 16);
	hdr.off_quad_mask = (uint32_t)off; off = align_up(off + ft.quad_mask.size() * sizeof(uint32_t), 16);
	hdr.off_chunks = (uint32_t)off; off = align_up(off + ft.chunks.size() * sizeof(uint32_t), 16);
	hdr.total_bytes = (uint32_t)off;
	if (off > 160 * 1024) return fail(SSB_ERR_UNSUPPORTED, "scene tables (%zu bytes) exceed the shared-memory budget", off);

	std::vector<unsigned char> blob(off, 0);
	memcpy(blob.data(), &hdr, sizeof(hdr));
	if (!c->quads.empty()) memcpy(blob.data() + hdr.off_quads, c->quads.data(), c->quads.size() * sizeof(ssb_quad));
	if (!mats.empty()) memcpy(blob.data() + hdr.off_materials, mats.data(), mats.size() * sizeof(DevMaterial));
	if (!c->lights.empty()) memcpy(blob.data() + hdr.off_lights, c->lights.data(), c->lights.size() * sizeof(uint32_t));
	if (!texs.empty()) memcpy(blob.data() + hdr.off_textures, texs.data(), texs.size() * sizeof(DevTexture));
	if (!pool.empty()) memcpy(blob.data() + hdr.off_pool, pool.data(), pool.size() * sizeof(float));
	if (!ft.pairs.empty()) memcpy(blob.data() + hdr.off_fpairs, ft.pairs.data(), ft.pairs.size() * sizeof(float));
	if (!ft.planes.empty()) memcpy(blob.data() + hdr.off_planes, ft.planes.data(), ft.planes.size() * sizeof(float));
	if (!ft.entry_quad.empty()) memcpy(blob.data() + hdr.off_entry_quad, ft.entry_quad.data(), ft.entry_quad.size() * sizeof(uint32_t));
	if (!ft.quad_mask.empty()) memcpy(blob.data() + hdr.off_quad_mask, ft.quad_mask.data(), ft.quad_mask.size() * sizeof(uint32_t));
	if (!ft.chunks.empty()) memcpy(blob.data() + hdr.off_chunks, ft.chunks.data(), ft.chunks.size() * sizeof(uint32_t));

	if (off > c->blob_capacity) {
		if (c->d_blob) cudaFree(c->d_blob);
		c->d_blob = nullptr; c->blob_capacity = 0;
		SSB_CUDA(cudaMalloc(&c->d_blob, off));
		c->blob_capacity = off;
	}
	SSB_CUDA(cudaMemcpyAsync(c->d_blob, blob.data(), off, cudaMemcpyHostToDevice, c->stream));
	SSB_CUDA(cudaStreamSynchronize(c->stream));  // `blob` is a pageable temporary
	c->blob_bytes = off;
	c->blob_dirty = false;
	return SSB_OK;
}

int ensure_accum(ssb_ctx* c, uint32_t w, uint32_t h) {
	if (c->d_accum && c->accum_w == w && c->accum_h == h) return SSB_OK;
	if (c->d_accum) cudaFree(c->d_accum);
	c->d_accum = nullptr;
	SSB_CUDA(cudaMalloc(&c->d_accum, (size_t)w * h * 4 * sizeof(double)));
	SSB_CUDA(cudaMemsetAsync(c->d_accum, 0, (size_t)w * h * 4 * sizeof(double), c->stream));
	c->accum_w = w; c->accum_h = h;
	return SSB_OK;
}

int validate_options(const ssb_options* o, uint32_t& x1, uint32_t& y1, uint32_t& s1) {
	if (!o) return fail(SSB_ERR_ARG, "options is NULL");
	if (o->width == 0 || o->height == 0 || o->spp == 0) return fail(SSB_ERR_ARG, "width, height and spp must be positive");
	x1 = o->x1 ? o->x1 : o->width; y1 = o->y1 ? o->y1 : o->height; s1 = o->sample_end ? o->sample_end : o->spp;
	if (x1 > o->width || y1 > o->height || o->x0 > x1 || o->y0 > y1 || o->sample_begin > s1)
		return fail(SSB_ERR_ARG, "work subset out of range");
	if (o->render_mode > SSB_RENDER_RGB) return fail(SSB_ERR_UNSUPPORTED, "unknown render mode %u", o->render_mode);
	const bool rgb = o->render_mode == SSB_RENDER_RGB;  // upsampling / wavelength range are not used in RGB mode
	if (!rgb && (o->upsampling < SSB_UPSAMPLE_OURS || o->upsampling > SSB_UPSAMPLE_JH)) return fail(SSB_ERR_UNSUPPORTED, "unknown upsampling mode %u", o->upsampling);
	if (o->max_depth == 0 || o->max_depth > SSB_MAX_DEPTH) return fail(SSB_ERR_UNSUPPORTED, "max_depth must be in [1,%u]", SSB_MAX_DEPTH);
	if (!rgb && !(o->lambda_max > o->lambda_min)) return fail(SSB_ERR_ARG, "lambda_max must exceed lambda_min");
	if (o->n_wavelengths != 0 && (o->n_wavelengths < 2 || o->n_wavelengths > 4))  // glm::vec<N,float> of the reference: N = 2, 3, 4
		return fail(SSB_ERR_UNSUPPORTED, "n_wavelengths must be 0 (= 4), 2, 3 or 4");
	if (o->band_count > 1) {
		if (o->band_height == 0 || o->band_index >= o->band_count) return fail(SSB_ERR_ARG, "row bands: band_height must be positive and band_index < band_count");
		if (o->y0 != 0 || y1 != o->height) return fail(SSB_ERR_ARG, "row bands cover the whole image height: y0/y1 must be 0");
	}
	if (o->scan_mode > SSB_SCAN_LIST) return fail(SSB_ERR_UNSUPPORTED, "unknown scan mode %u", o->scan_mode);
	return SSB_OK;
}

// rows j < height with (j / band_h) % band_n == band_i
uint32_t band_rows(uint32_t height, uint32_t band_h, uint32_t band_n, uint32_t band_i) {
	uint32_t rows = 0;
	for (uint32_t b = band_i; (unsigned long long)b * band_h < height; b += band_n)
		rows += std::min<unsigned long long>(band_h, height - (unsigned long long)b * band_h);
	return rows;
}

const size_t kWaveBudgetBytes = (size_t)24 << 30;  // device memory for the path state + fold records of one pass (of 180 GB)
// counters, cleared once per pass: queue lengths [D+2] | hits per depth [D] | per-depth per-quad counts and cursors
const size_t kCountsOff = 0, kNhitsOff = SSB_MAX_DEPTH + 2, kBinCountOff = kNhitsOff + SSB_MAX_DEPTH,
             kBinCursorOff = kBinCountOff + (size_t)SSB_MAX_DEPTH * SSB_MAX_QUADS,
             kCounterWords = kBinCursorOff + (size_t)SSB_MAX_DEPTH * SSB_MAX_QUADS;

}  // namespace

extern "C" {

uint32_t ssb_abi_version(void) { return SSB_ABI_VERSION; }
const char* ssb_last_error(void) { return g_last_error.c_str(); }

void ssb_default_options(ssb_options* opt, uint32_t width, uint32_t height, uint32_t spp) {
	if (!opt) return;
	memset(opt, 0, sizeof(*opt));
	opt->width = width; opt->height = height; opt->spp = spp;
	opt->upsampling = SSB_UPSAMPLE_OURS;       // RENDER_MODE_SPECTRAL_ALGNUM 1 (stdafx.hpp:66)
	opt->lambda_min = 380.0f; opt->lambda_max = 780.0f;  // CIE 1931 (stdafx.hpp:115-117)
	opt->max_depth = 10;                       // MAX_DEPTH (stdafx.hpp:47)
	opt->explicit_light_sampling = 1;          // EXPLICIT_LIGHT_SAMPLING (stdafx.hpp:44)
	opt->flat_field_correction = 1;            // FLAT_FIELD_CORRECTION (stdafx.hpp:55)
	opt->eps = 0.001f;                         // EPS (stdafx.hpp:58)
	opt->seed = 1;
}

int ssb_device_count(int* count) {
	if (!count) return fail(SSB_ERR_ARG, "ssb_device_count: NULL argument");
	*count = 0;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) return fail(SSB_ERR_DATA, "no CUDA device available (%s); this library has no CPU path", cudaGetErrorString(e));
	*count = n;
	return SSB_OK;
}

int ssb_create(int device, ssb_ctx** out) {
	if (!out) return fail(SSB_ERR_ARG, "ssb_create: out is NULL");
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
		return fail(SSB_ERR_DATA, "ssb_create: no CUDA device available (%s); this library has no CPU path", cudaGetErrorString(e));
	if (device < 0 || device >= count) return fail(SSB_ERR_ARG, "ssb_create: device %d out of range [0,%d)", device, count);
	SSB_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop{};
	SSB_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10)
		return fail(SSB_ERR_UNSUPPORTED, "ssb_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
	ssb_ctx* c = new ssb_ctx();
	c->device = device;
	c->sm_count = prop.multiProcessorCount;
	// every failure from here on releases what was created so far (ssb_destroy accepts a partially built context)
	cudaError_t e2 = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
	c->stream = c->own_stream;
	if (e2 == cudaSuccess) e2 = cudaEventCreate(&c->ev_begin);
	if (e2 == cudaSuccess) e2 = cudaEventCreate(&c->ev_end);
	if (e2 == cudaSuccess) e2 = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
	if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&c->ev_tex_ready, cudaEventDisableTiming);
	if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&c->ev_tex_free, cudaEventDisableTiming);
	if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&c->ev_accum_ready, cudaEventDisableTiming);
	if (e2 == cudaSuccess) e2 = cudaMalloc(&c->d_counts, kCounterWords * sizeof(uint32_t));
	if (e2 != cudaSuccess) {
		ssb_destroy(c);
		return fail(SSB_ERR_DATA, "ssb_create: CUDA error %s while creating the context of device %d", cudaGetErrorString(e2), device);
	}
	*out = c;
	return SSB_OK;
}

void ssb_destroy(ssb_ctx* c) {
	if (!c) return;
	cudaSetDevice(c->device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
	free_textures(c);
	cudaFree(c->d_blob); cudaFree(c->d_jh_scale); cudaFree(c->d_jh_data); cudaFree(c->d_meng_grid); cudaFree(c->d_meng_points);
	cudaFree(c->d_accum); cudaFree(c->d_counts); cudaFree(c->d_wave); cudaFree(c->d_xyza); cudaFree(c->d_srgba);
	cudaFree(c->d_rgb_staging);
	cudaFree(c->d_peer_staging);
	if (c->ev_begin) cudaEventDestroy(c->ev_begin);
	if (c->ev_end) cudaEventDestroy(c->ev_end);
	if (c->ev_tex_ready) cudaEventDestroy(c->ev_tex_ready);
	if (c->ev_tex_free) cudaEventDestroy(c->ev_tex_free);
	if (c->ev_accum_ready) cudaEventDestroy(c->ev_accum_ready);
	for (cudaEvent_t e : c->ev_pass) cudaEventDestroy(e);
	if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
	if (c->own_stream) cudaStreamDestroy(c->own_stream);
	delete c;
}

static int upload_scene_impl(ssb_ctx* c, const ssb_scene* scene, bool async);
int ssb_upload_scene(ssb_ctx* c, const ssb_scene* scene) { return upload_scene_impl(c, scene, false); }
int ssb_upload_scene_async(ssb_ctx* c, const ssb_scene* scene) { return upload_scene_impl(c, scene, true); }

static int upload_scene_impl(ssb_ctx* c, const ssb_scene* scene, bool async) {
	if (!c || !scene) return fail(SSB_ERR_ARG, "ssb_upload_scene: NULL argument");
	if (scene->nquads == 0 || !scene->quads) return fail(SSB_ERR_ARG, "ssb_upload_scene: empty primitive list");
	if (scene->nquads > SSB_MAX_QUADS) return fail(SSB_ERR_UNSUPPORTED, "ssb_upload_scene: %u quads exceed SSB_MAX_QUADS=%u", scene->nquads, SSB_MAX_QUADS);
	if (scene->nmaterials == 0 || !scene->materials || scene->nmaterials > SSB_MAX_MATERIALS)
		return fail(SSB_ERR_ARG, "ssb_upload_scene: bad material list");
	if (scene->ntextures && !scene->textures) return fail(SSB_ERR_ARG, "ssb_upload_scene: textures is NULL");
	SSB_CUDA(cudaSetDevice(c->device));
	std::vector<HostMaterial> mats(scene->nmaterials);
	for (uint32_t m = 0; m < scene->nmaterials; ++m) {
		const ssb_material& sm = scene->materials[m];
		if (sm.kind > SSB_MATERIAL_MIRROR || sm.albedo_mode > SSB_ALBEDO_TEXTURE) return fail(SSB_ERR_ARG, "material %u: bad kind/mode", m);
		mats[m].kind = sm.kind; mats[m].albedo_mode = sm.albedo_mode; mats[m].texture = sm.texture;
		int rc;
		// the spectra may be absent in a scene meant for RGB mode only; ssb_render checks what its mode needs
		if (sm.albedo_mode == SSB_ALBEDO_CONSTANT) { if ((rc = copy_spectrum(sm.albedo, mats[m].albedo, "material albedo", false)) != SSB_OK) return rc; }
		else if (sm.texture >= scene->ntextures) return fail(SSB_ERR_ARG, "material %u: texture index out of range", m);
		if ((rc = copy_spectrum(sm.emission, mats[m].emission, "material emission", false)) != SSB_OK) return rc;
		for (int k = 0; k < 3; ++k) { mats[m].albedo_rgb[k] = sm.albedo_rgb[k]; mats[m].emission_rgb[k] = sm.emission_rgb[k]; }
	}
	std::vector<uint32_t> lights;
	for (uint32_t q = 0; q < scene->nquads; ++q) {
		if (scene->quads[q].material >= scene->nmaterials) return fail(SSB_ERR_ARG, "quad %u: material index out of range", q);
		if (scene->quads[q].is_light) lights.push_back(q);  // Scene::_init (scene.cpp:26-29)
	}
	if (lights.size() > SSB_MAX_LIGHTS) return fail(SSB_ERR_UNSUPPORTED, "more than %u lights", SSB_MAX_LIGHTS);
	// textures: RGB8 -> RGBA8 on the device.  Synchronous form: on the render stream, returns when the caller's buffers
	// are no longer needed.  Asynchronous form: on the copy stream, ordered after the last render that read the old
	// texels and before the first shade kernel of the next render (which waits on ev_tex_ready, see ssb_render).
	cudaStream_t up = c->stream;
	if (async) {
		up = c->copy_stream;
		if (c->tex_in_use) SSB_CUDA(cudaStreamWaitEvent(up, c->ev_tex_free, 0));
	} else {
		SSB_CUDA(cudaStreamSynchronize(c->copy_stream));
		SSB_CUDA(cudaStreamSynchronize(c->stream));
		c->tex_pending = false;
	}
	// ---- textures, transactionally: (1) everything that can be refused is checked, and every buffer the new scene needs is
	// allocated, BEFORE the context is touched — a NULL texel pointer or an out-of-memory leaves the previous scene intact;
	// (2) only then are texels enqueued into (re-used or new) buffers; a CUDA error past that point invalidates the scene
	// (have_scene = false, all texture buffers released) instead of leaving materials that index a half-built table.
	for (uint32_t t = 0; t < scene->ntextures; ++t) {
		const ssb_texture& tx = scene->textures[t];
		if (!tx.rgb8 || tx.width == 0 || tx.height == 0) return fail(SSB_ERR_DATA, "texture %u: could not load texture", t);  // material.cpp:15-18
	}
	std::vector<uchar4*> new_tex(scene->ntextures, nullptr);
	std::vector<char> reused(scene->ntextures, 0);
	size_t staging_need = 0;
	for (uint32_t t = 0; t < scene->ntextures; ++t) staging_need = std::max(staging_need, (size_t)3 * scene->textures[t].width * scene->textures[t].height);
	unsigned char* new_staging = nullptr;
	auto release_new = [&]() {
		for (uint32_t t = 0; t < scene->ntextures; ++t) if (new_tex[t] && !reused[t]) cudaFree(new_tex[t]);
		if (new_staging) cudaFree(new_staging);
	};
	for (uint32_t t = 0; t < scene->ntextures; ++t) {
		const ssb_texture& tx = scene->textures[t];
		if (t < c->d_textures.size() && c->tex_w[t] == tx.width && c->tex_h[t] == tx.height) { new_tex[t] = c->d_textures[t]; reused[t] = 1; continue; }
		cudaError_t e = cudaMalloc(&new_tex[t], (size_t)tx.width * tx.height * sizeof(uchar4));
		if (e != cudaSuccess) { new_tex[t] = nullptr; release_new(); return fail(SSB_ERR_DATA, "ssb_upload_scene: texture %u: %s", t, cudaGetErrorString(e)); }
	}
	if (staging_need > c->rgb_staging_capacity) {
		cudaError_t e = cudaMalloc(&new_staging, staging_need);
		if (e != cudaSuccess) { new_staging = nullptr; release_new(); return fail(SSB_ERR_DATA, "ssb_upload_scene: staging buffer: %s", cudaGetErrorString(e)); }
	}
	// ---- commit
	if (new_staging) {
		cudaStreamSynchronize(up);  // (a previous asynchronous upload may still read the old staging buffer)
		cudaFree(c->d_rgb_staging);
		c->d_rgb_staging = new_staging; c->rgb_staging_capacity = staging_need;
	}
	for (size_t t = 0; t < c->d_textures.size(); ++t)
		if (!(t < scene->ntextures && reused[t])) cudaFree(c->d_textures[t]);  // (the synchronisation / event wait above covers their last reader)
	c->d_textures.assign(new_tex.begin(), new_tex.end());
	c->tex_w.resize(scene->ntextures); c->tex_h.resize(scene->ntextures);
	cudaError_t copy_err = cudaSuccess;
	for (uint32_t t = 0; t < scene->ntextures && copy_err == cudaSuccess; ++t) {
		const ssb_texture& tx = scene->textures[t];
		const size_t n = (size_t)tx.width * tx.height;
		c->tex_w[t] = tx.width; c->tex_h[t] = tx.height;
		// one H2D copy of the caller's RGB8 scanlines (fast when the caller's buffer is pinned), repack on the device
		copy_err = cudaMemcpyAsync(c->d_rgb_staging, tx.rgb8, 3 * n, cudaMemcpyHostToDevice, up);
		if (copy_err == cudaSuccess) {
			ssb_repack_rgb8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, up>>>(c->d_rgb_staging, c->d_textures[t], n);
			copy_err = cudaGetLastError();
		}
	}
	if (copy_err != cudaSuccess) {
		cudaStreamSynchronize(up);
		free_textures(c);
		c->have_scene = false; c->blob_dirty = true; c->tex_pending = false;
		return fail(SSB_ERR_DATA, "ssb_upload_scene: texel upload failed (%s); the context holds no scene now", cudaGetErrorString(copy_err));
	}
	if (async) {
		SSB_CUDA(cudaEventRecord(c->ev_tex_ready, up));
		c->tex_pending = scene->ntextures != 0;
	} else {
		SSB_CUDA(cudaStreamSynchronize(c->stream));  // the caller's host buffers may be released after return
	}
	c->camera = scene->camera;
	c->quads.assign(scene->quads, scene->quads + scene->nquads);
	c->materials.swap(mats);
	c->lights.swap(lights);
	c->have_scene = true; c->blob_dirty = true;
	c->coef_valid = false;  // new texels: re-baked by the next ssb_render that asks for prebaked textures
	return SSB_OK;
}

int ssb_upload_color(ssb_ctx* c, const ssb_color* color) {
	if (!c || !color) return fail(SSB_ERR_ARG, "ssb_upload_color: NULL argument");
	SSB_CUDA(cudaSetDevice(c->device));
	int rc;
	if ((rc = copy_spectrum(color->xbar, c->xbar, "xbar", true)) != SSB_OK) return rc;
	if ((rc = copy_spectrum(color->ybar, c->ybar, "ybar", true)) != SSB_OK) return rc;
	if ((rc = copy_spectrum(color->zbar, c->zbar, "zbar", true)) != SSB_OK) return rc;
	if ((rc = copy_spectrum(color->basis_r, c->basis_r, "basis_r", false)) != SSB_OK) return rc;
	if ((rc = copy_spectrum(color->basis_g, c->basis_g, "basis_g", false)) != SSB_OK) return rc;
	if ((rc = copy_spectrum(color->basis_b, c->basis_b, "basis_b", false)) != SSB_OK) return rc;
	memcpy(c->xyz_to_lrgb, color->xyz_to_lrgb, sizeof(c->xyz_to_lrgb));
	c->d65_rad_Y = color->d65_rad_Y;
	SSB_CUDA(cudaStreamSynchronize(c->stream));  // no render may still read the tables that are replaced below
	// Jakob-Hanika (9.4 MB) and Meng (69 KB) tables: the device buffers are kept across calls and only re-allocated when
	// their size changes (a caller that re-uploads its colour tables every frame paid a cudaFree + cudaMalloc of 9.4 MB
	// per call: the end-to-end leg of the JH configuration)
	if (color->jh_res) {
		if (!color->jh_scale || !color->jh_data) return fail(SSB_ERR_ARG, "ssb_upload_color: JH tables missing");
		size_t res = color->jh_res, nd = 3 * res * res * res * 3;
		if (c->jh_res != color->jh_res || !c->d_jh_scale || !c->d_jh_data) {
			cudaFree(c->d_jh_scale); cudaFree(c->d_jh_data); c->d_jh_scale = c->d_jh_data = nullptr; c->jh_res = 0;
			SSB_CUDA(cudaMalloc(&c->d_jh_scale, res * sizeof(float)));
			SSB_CUDA(cudaMalloc(&c->d_jh_data, nd * sizeof(float)));
		}
		c->jh_res = 0;  // (invalid until both copies have landed)
		SSB_CUDA(copy_to_device(c, c->d_jh_scale, color->jh_scale, res * sizeof(float)));
		SSB_CUDA(copy_to_device(c, c->d_jh_data, color->jh_data, nd * sizeof(float)));
		c->jh_res = color->jh_res;
	} else {
		cudaFree(c->d_jh_scale); cudaFree(c->d_jh_data); c->d_jh_scale = c->d_jh_data = nullptr; c->jh_res = 0;
	}
	if (color->meng) {
		const ssb_meng_tables& m = *color->meng;
		if (!m.grid || !m.points || m.grid_w == 0 || m.grid_h == 0 || m.npoints == 0 || m.nsamples < 2) return fail(SSB_ERR_ARG, "ssb_upload_color: bad Meng tables");
		size_t ng = (size_t)m.grid_w * m.grid_h * 8, np = (size_t)m.npoints * (5 + m.nsamples);
		const bool same = c->have_meng && c->meng.grid_w == m.grid_w && c->meng.grid_h == m.grid_h && c->meng.npoints == m.npoints && c->meng.nsamples == m.nsamples;
		c->have_meng = false;
		if (!same) {
			cudaFree(c->d_meng_grid); cudaFree(c->d_meng_points); c->d_meng_grid = nullptr; c->d_meng_points = nullptr;
			SSB_CUDA(cudaMalloc(&c->d_meng_grid, ng * sizeof(int32_t)));
			SSB_CUDA(cudaMalloc(&c->d_meng_points, np * sizeof(float)));
		}
		SSB_CUDA(copy_to_device(c, c->d_meng_grid, m.grid, ng * sizeof(int32_t)));
		SSB_CUDA(copy_to_device(c, c->d_meng_points, m.points, np * sizeof(float)));
		c->meng = m; c->meng.grid = nullptr; c->meng.points = nullptr;
		c->have_meng = true;
	} else {
		cudaFree(c->d_meng_grid); cudaFree(c->d_meng_points); c->d_meng_grid = nullptr; c->d_meng_points = nullptr; c->have_meng = false;
	}
	c->have_color = true; c->blob_dirty = true;
	c->coef_valid = false;  // the JH tables may have changed
	return SSB_OK;
}

int ssb_clear(ssb_ctx* c) {
	if (!c) return fail(SSB_ERR_ARG, "ssb_clear: NULL context");
	SSB_CUDA(cudaSetDevice(c->device));
	if (c->d_accum) SSB_CUDA(cudaMemsetAsync(c->d_accum, 0, (size_t)c->accum_w * c->accum_h * 4 * sizeof(double), c->stream));
	return SSB_OK;
}

int ssb_render(ssb_ctx* c, const ssb_options* o) {
	if (!c) return fail(SSB_ERR_ARG, "ssb_render: NULL context");
	uint32_t x1, y1, s1;
	int rc = validate_options(o, x1, y1, s1);
	if (rc != SSB_OK) return rc;
	SSB_CUDA(cudaSetDevice(c->device));
	if (c->blob_dirty && (rc = build_blob(c)) != SSB_OK) return rc;
	if (o->explicit_light_sampling && c->lights.empty()) return fail(SSB_ERR_ARG, "explicit light sampling needs at least one light (scene.cpp:30)");
	const bool rgb = o->render_mode == SSB_RENDER_RGB;
	if (!rgb) {
		if (!c->have_color) return fail(SSB_ERR_ARG, "ssb_render: no colour tables uploaded");
		for (size_t m = 0; m < c->materials.size(); ++m) {
			const HostMaterial& hm = c->materials[m];
			if (!hm.emission.present || (hm.albedo_mode == SSB_ALBEDO_CONSTANT && !hm.albedo.present))
				return fail(SSB_ERR_ARG, "material %zu: missing spectrum (the scene was uploaded with RGB constants only)", m);
		}
	}
	if (!rgb && o->upsampling == SSB_UPSAMPLE_OURS && !c->basis_r.present) {
		for (const HostMaterial& m : c->materials)
			if (m.albedo_mode == SSB_ALBEDO_TEXTURE) return fail(SSB_ERR_ARG, "OURS upsampling needs the basis spectra");
	}
	if (!rgb && o->upsampling == SSB_UPSAMPLE_JH && !c->jh_res) return fail(SSB_ERR_ARG, "JH upsampling needs the coefficient tables");
	if (!rgb && o->upsampling == SSB_UPSAMPLE_MENG && !c->have_meng) return fail(SSB_ERR_ARG, "MENG upsampling needs the grid tables");
	// prebaked JH coefficient textures: (re)allocate before the blob is built (it carries the pointers), bake after
	// (the bake kernel reads the sRGB LUT from the blob)
	const bool prebaked = !rgb && o->upsampling == SSB_UPSAMPLE_JH && o->prebaked_textures && !c->d_textures.empty();
	uint32_t bake_launches = 0;
	if (prebaked && !c->coef_valid) {
		while (c->d_tex_coef.size() > c->d_textures.size()) {  // the scene was replaced by one with fewer textures
			SSB_CUDA(cudaStreamSynchronize(c->stream));
			cudaFree(c->d_tex_coef.back()); c->d_tex_coef.pop_back(); c->tex_coef_n.pop_back();
			c->blob_dirty = true;
		}
		c->d_tex_coef.resize(c->d_textures.size(), nullptr);
		c->tex_coef_n.resize(c->d_textures.size(), 0);
		for (size_t t = 0; t < c->d_textures.size(); ++t) {
			const size_t n = (size_t)c->tex_w[t] * c->tex_h[t];
			if (c->d_tex_coef[t] && c->tex_coef_n[t] == n) continue;
			SSB_CUDA(cudaStreamSynchronize(c->stream));  // an earlier render may still read the old buffer
			cudaFree(c->d_tex_coef[t]); c->d_tex_coef[t] = nullptr; c->tex_coef_n[t] = 0;
			SSB_CUDA(cudaMalloc(&c->d_tex_coef[t], n * sizeof(float4)));
			c->tex_coef_n[t] = n;
			c->blob_dirty = true;
		}
		if (c->blob_dirty && (rc = build_blob(c)) != SSB_OK) return rc;
		// After an ssb_upload_scene_async the bake follows the texel copies on the COPY stream, so that it still overlaps
		// the camera-ray stage; the first shade stage then waits for texels and coefficients together (ev_tex_ready is
		// recorded again behind the bake).  That stream is already ordered after the last render that read the old
		// coefficients (the upload made it wait for ev_tex_free), and the blob it reads was copied synchronously.
		cudaStream_t bake_stream = c->tex_pending ? c->copy_stream : c->stream;
		for (size_t t = 0; t < c->d_textures.size(); ++t) {
			const size_t n = c->tex_coef_n[t];
			const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)c->sm_count * 32);
			ssb_bake_jh_kernel<<<grid, 256, 0, bake_stream>>>(c->d_textures[t], c->d_tex_coef[t], n, c->d_blob, c->d_jh_scale, c->d_jh_data, c->jh_res);
			SSB_CUDA(cudaGetLastError());
			++bake_launches;
		}
		if (c->tex_pending) SSB_CUDA(cudaEventRecord(c->ev_tex_ready, bake_stream));
		c->coef_valid = true;
	}
	if ((rc = ensure_accum(c, o->width, o->height)) != SSB_OK) return rc;
	if (o->sample_begin == 0 && !o->keep_accumulator) SSB_CUDA(cudaMemsetAsync(c->d_accum, 0, (size_t)o->width * o->height * 4 * sizeof(double), c->stream));

	const bool banded = o->band_count > 1;
	const uint32_t rect_w = x1 - o->x0, rect_h = banded ? band_rows(o->height, o->band_height, o->band_count, o->band_index) : y1 - o->y0;
	const size_t npix_rect = (size_t)rect_w * rect_h;
	const uint32_t nsamp_total = s1 - o->sample_begin;
	c->stats = ssb_stats{}; c->stats_pending = false; c->passes = 0;
	if (npix_rect == 0 || nsamp_total == 0) return SSB_OK;

	// ---- size one pass: N = npix_rect * chunk samples share the wavefront buffers
	const uint32_t nrec_depths = o->max_depth > 1 ? o->max_depth - 1 : 1;
	const size_t bytes_per_sample = 2 * (32 + 32) + 32 + (size_t)nrec_depths * 64 + 32 + 4 + (4 + 4);
	size_t budget = kWaveBudgetBytes;
	if (const char* e = getenv("SSB_WAVE_BUDGET_MB")) {  // tests force multi-pass rendering with a tiny budget
		long mb = atol(e);
		if (mb > 0) budget = (size_t)mb << 20;
	}
	const size_t max_samples = std::min<size_t>(std::max<size_t>(budget / bytes_per_sample, npix_rect), (size_t)1 << 31);
	if (npix_rect > max_samples) return fail(SSB_ERR_UNSUPPORTED, "pixel rectangle too large for one pass");
	const uint32_t chunk = (uint32_t)std::min<size_t>(nsamp_total, std::max<size_t>(1, max_samples / npix_rect));
	const size_t N = npix_rect * chunk;
	auto up = [](size_t v) { return (v + 255) / 256 * 256; };
	size_t off = 0;
	const size_t o_a0 = off; off += up(N * 32); const size_t o_a1 = off; off += up(N * 32);
	const size_t o_r0 = off; off += up(N * 32); const size_t o_r1 = off; off += up(N * 32);
	const size_t o_h = off; off += up(N * 32);
	const size_t o_stk = off; off += up(N * nrec_depths * 64);
	const size_t o_leaf = off; off += up(N * 32);
	const size_t o_ff = off; off += up(N * 4);
	const size_t o_hq = off; off += up(N * 4);
	const size_t o_ord = off; off += up(N * 4);
	if (off > c->wave_bytes) {
		if (c->d_wave) cudaFree(c->d_wave);
		c->d_wave = nullptr; c->wave_bytes = 0;
		SSB_CUDA(cudaMalloc(&c->d_wave, off));
		c->wave_bytes = off;
	}
	c->d_samples = reinterpret_cast<float4*>(c->d_wave + o_a0);  // dead ray records of the pass are reused (N x 16 B <= N x 32 B)
	c->samples_capacity = N;

	KParams P{};
	P.blob = c->d_blob;
	unsigned char* wv = c->d_wave;
	P.recA[0] = reinterpret_cast<float4*>(wv + o_a0); P.recA[1] = reinterpret_cast<float4*>(wv + o_a1);
	P.recR[0] = reinterpret_cast<float4*>(wv + o_r0); P.recR[1] = reinterpret_cast<float4*>(wv + o_r1);
	P.recH = reinterpret_cast<float4*>(wv + o_h);
	P.stk = reinterpret_cast<float4*>(wv + o_stk);
	P.leaf = reinterpret_cast<float4*>(wv + o_leaf);
	P.ff = reinterpret_cast<float*>(wv + o_ff);
	P.counts = c->d_counts + kCountsOff;
	P.nhits = c->d_counts + kNhitsOff;
	P.bin_count = c->d_counts + kBinCountOff;
	P.bin_cursor = c->d_counts + kBinCursorOff;
	P.hit_q = reinterpret_cast<uint32_t*>(wv + o_hq);
	P.order = reinterpret_cast<uint32_t*>(wv + o_ord);
	P.samples = c->d_samples;
	P.accum = c->d_accum;
	P.width = o->width; P.height = o->height; P.x0 = o->x0; P.y0 = o->y0; P.rect_w = rect_w; P.rect_h = rect_h;
	P.indirect_only = o->indirect_only; P.upsampling = o->upsampling; P.max_depth = o->max_depth;
	P.els = o->explicit_light_sampling; P.flat_field = o->flat_field_correction;
	P.render_mode = o->render_mode;
	P.has_mirror = 0;
	for (const HostMaterial& m : c->materials) if (m.kind == SSB_MATERIAL_MIRROR) P.has_mirror = 1;
	P.band_h = banded ? o->band_height : 1u; P.band_n = banded ? o->band_count : 1u; P.band_i = banded ? o->band_index : 0u;
	P.eps = o->eps; P.lambda_min = o->lambda_min;
	P.n_wavelengths = o->n_wavelengths ? o->n_wavelengths : 4u;  // SAMPLE_WAVELENGTHS (stdafx.hpp:90)
	P.lambda_step = (o->lambda_max - o->lambda_min) / (float)P.n_wavelengths;  // LAMBDA_STEP (stdafx.hpp:289)
	P.seed = o->seed;
	memcpy(P.pv_inv, c->camera.pv_inv, sizeof(P.pv_inv));
	memcpy(P.cam_pos, c->camera.pos, sizeof(P.cam_pos));
	memcpy(P.cam_dir, c->camera.dir, sizeof(P.cam_dir));
	P.jh_scale = c->d_jh_scale; P.jh_data = c->d_jh_data; P.jh_res = c->jh_res;
	P.jh_prebaked = prebaked ? 1u : 0u;
	P.meng_grid = c->d_meng_grid; P.meng_points = c->d_meng_points;
	P.meng_grid_w = c->meng.grid_w; P.meng_grid_h = c->meng.grid_h; P.meng_npoints = c->meng.npoints; P.meng_nsamples = c->meng.nsamples;
	memcpy(P.meng_xy_to_uv, c->meng.xy_to_uv, sizeof(P.meng_xy_to_uv));
	P.meng_sample_min = c->meng.sample_min; P.meng_sample_max = c->meng.sample_max;

	const size_t smem = c->blob_bytes;  // the scene tables; the two big kernels add their record pipelines behind them
	const size_t smem_i = align_up(smem, 128) + SSB_INTERSECT_PIPE_BYTES, smem_s = align_up(smem, 128) + SSB_SHADE_PIPE_BYTES;
	typedef void (*kfn)(const KParams);
	kfn k_shade_first = nullptr, k_shade_next = nullptr, k_isect_first = nullptr, k_isect_next = nullptr;
	// one instantiation per upsampling mode keeps the instruction footprint small; LIST = ssb_options.scan_mode
#define SSB_PICK(UPS, LIST)                                                                                         \
	do {                                                                                                            \
		k_shade_first = ssb_shade_kernel<true, UPS, LIST>; k_shade_next = ssb_shade_kernel<false, UPS, LIST>;       \
		k_isect_first = ssb_intersect_kernel<true, LIST>; k_isect_next = ssb_intersect_kernel<false, LIST>;          \
	} while (0)
	const bool list = o->scan_mode == SSB_SCAN_LIST;
	switch (rgb ? (uint32_t)SSB_UPS_RGB : o->upsampling) {
		case SSB_UPS_RGB: if (list) SSB_PICK(SSB_UPS_RGB, true); else SSB_PICK(SSB_UPS_RGB, false); break;
		case SSB_UPSAMPLE_OURS: if (list) SSB_PICK(SSB_UPSAMPLE_OURS, true); else SSB_PICK(SSB_UPSAMPLE_OURS, false); break;
		case SSB_UPSAMPLE_JH: if (list) SSB_PICK(SSB_UPSAMPLE_JH, true); else SSB_PICK(SSB_UPSAMPLE_JH, false); break;
		default: if (list) SSB_PICK(SSB_UPSAMPLE_MENG, true); else SSB_PICK(SSB_UPSAMPLE_MENG, false); break;
	}
#undef SSB_PICK
	// dynamic shared memory opt-in + resident CTAs per SM: queried once per (kernel set, table size) and kept in the context
	// (four attribute calls and four occupancy queries per ssb_render were a measurable part of a 2 ms strong-scaling slice)
	int occ_sf = 0, occ_sn = 0, occ_if = 0, occ_in = 0;
	if (c->occ_key_kernel != (const void*)k_shade_next || c->occ_key_smem != smem) {
		for (kfn k : { k_shade_first, k_shade_next }) SSB_CUDA(raise_dynamic_smem_limit(c->device, (const void*)k, smem_s));
		for (kfn k : { k_isect_first, k_isect_next }) SSB_CUDA(raise_dynamic_smem_limit(c->device, (const void*)k, smem_i));
		SSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ[0], k_shade_first, SSB_SHADE_THREADS, smem_s));
		SSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ[1], k_shade_next, SSB_SHADE_THREADS, smem_s));
		SSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ[2], k_isect_first, SSB_INTERSECT_THREADS, smem_i));
		SSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ[3], k_isect_next, SSB_INTERSECT_THREADS, smem_i));
		c->occ_key_kernel = (const void*)k_shade_next; c->occ_key_smem = smem;
	}
	occ_sf = c->occ[0]; occ_sn = c->occ[1]; occ_if = c->occ[2]; occ_in = c->occ[3];
	if (occ_sf < 1 || occ_sn < 1 || occ_if < 1 || occ_in < 1) return fail(SSB_ERR_UNSUPPORTED, "kernels do not fit on an SM with %zu bytes of tables", smem);
	const uint32_t nquads = (uint32_t)c->quads.size();

	SSB_CUDA(cudaEventRecord(c->ev_begin, c->stream));
	uint32_t launches = bake_launches, passes = 0;
	for (uint32_t k0 = 0; k0 < nsamp_total; k0 += chunk) {
		const uint32_t ns = std::min(chunk, nsamp_total - k0);
		P.sample_begin = o->sample_begin + k0;
		P.nsamp = ns;
		P.total_work = (unsigned long long)npix_rect * ns;
		SSB_CUDA(cudaMemsetAsync(c->d_counts, 0, kCounterWords * sizeof(uint32_t), c->stream));
		while (c->ev_pass.size() < 2 * (size_t)(passes + 1)) { cudaEvent_t e; SSB_CUDA(cudaEventCreate(&e)); c->ev_pass.push_back(e); }
		SSB_CUDA(cudaEventRecord(c->ev_pass[2 * passes], c->stream));
		// per path depth: closest-hit queries -> counting sort by hit quad -> shading; persistent grids (SMs x resident
		// CTAs, queue lengths are read on the device), never more CTAs than the first queue needs
		for (uint32_t d = 0; d < o->max_depth; ++d) {
			// with explicit light sampling the last depth is never entered (dead-work skip in the shade kernel)
			if (d > 0 && o->explicit_light_sampling && d + 1 >= o->max_depth) break;
			P.depth = d;
			const unsigned long long want_i = (P.total_work + SSB_INTERSECT_THREADS - 1) / SSB_INTERSECT_THREADS;
			const unsigned long long want_s = (P.total_work + SSB_SHADE_THREADS - 1) / SSB_SHADE_THREADS;
			const unsigned grid_i = (unsigned)std::min<unsigned long long>((unsigned long long)c->sm_count * (d == 0 ? occ_if : occ_in), want_i);
			const unsigned grid_s = (unsigned)std::min<unsigned long long>((unsigned long long)c->sm_count * (d == 0 ? occ_sf : occ_sn), want_s);
			(d == 0 ? k_isect_first : k_isect_next)<<<grid_i, SSB_INTERSECT_THREADS, smem_i, c->stream>>>(P);
			SSB_CUDA(cudaGetLastError());
			if (c->tex_pending) {  // texels are first read by the shade stage: the camera-ray queries overlap the upload
				SSB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_tex_ready, 0));
				c->tex_pending = false;
			}
			const unsigned grid_b = (unsigned)std::min<unsigned long long>((unsigned long long)c->sm_count * 8, (P.total_work + SSB_MAX_QUADS * SSB_SCATTER_PER_THREAD - 1) / (SSB_MAX_QUADS * SSB_SCATTER_PER_THREAD));
			ssb_bin_scatter_kernel<<<grid_b, SSB_MAX_QUADS, 0, c->stream>>>(P, d == 0 ? 1u : 0u, nquads);
			SSB_CUDA(cudaGetLastError());
			(d == 0 ? k_shade_first : k_shade_next)<<<grid_s, SSB_SHADE_THREADS, smem_s, c->stream>>>(P);
			SSB_CUDA(cudaGetLastError());
			launches += 3;
		}
		SSB_CUDA(cudaEventRecord(c->ev_pass[2 * passes + 1], c->stream));
		ssb_fold_kernel<<<(unsigned)((P.total_work + SSB_FOLD_THREADS - 1) / SSB_FOLD_THREADS), SSB_FOLD_THREADS, 0, c->stream>>>(P);
		SSB_CUDA(cudaGetLastError());
		ssb_accumulate_kernel<<<(unsigned)((npix_rect + 127) / 128), 128, 0, c->stream>>>(P);
		SSB_CUDA(cudaGetLastError());
		launches += 2; passes += 1;
	}
	SSB_CUDA(cudaEventRecord(c->ev_end, c->stream));
	SSB_CUDA(cudaEventRecord(c->ev_tex_free, c->stream));
	c->tex_in_use = true;
	// asynchronous: the event times are read lazily by ssb_get_stats()
	c->stats.samples = (uint64_t)npix_rect * nsamp_total;
	c->stats.launches = launches;
	c->passes = passes;
	c->stats_pending = true;
	return SSB_OK;
}

int ssb_read_accum(ssb_ctx* c, double* dst) {
	if (!c || !dst) return fail(SSB_ERR_ARG, "ssb_read_accum: NULL argument");
	if (!c->d_accum) return fail(SSB_ERR_ARG, "ssb_read_accum: nothing rendered yet");
	SSB_CUDA(cudaSetDevice(c->device));
	SSB_CUDA(cudaMemcpyAsync(dst, c->d_accum, (size_t)c->accum_w * c->accum_h * 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	SSB_CUDA(cudaStreamSynchronize(c->stream));
	return SSB_OK;
}
int ssb_write_accum(ssb_ctx* c, const double* src) {
	if (!c || !src) return fail(SSB_ERR_ARG, "ssb_write_accum: NULL argument");
	if (!c->d_accum) return fail(SSB_ERR_ARG, "ssb_write_accum: no accumulator yet");
	SSB_CUDA(cudaSetDevice(c->device));
	SSB_CUDA(cudaMemcpyAsync(c->d_accum, src, (size_t)c->accum_w * c->accum_h * 4 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	SSB_CUDA(cudaStreamSynchronize(c->stream));
	return SSB_OK;
}
int ssb_accum_device(ssb_ctx* c, double** dptr, size_t* count) {
	if (!c || !dptr) return fail(SSB_ERR_ARG, "ssb_accum_device: NULL argument");
	if (!c->d_accum) return fail(SSB_ERR_ARG, "ssb_accum_device: nothing rendered yet");
	SSB_CUDA(cudaSetDevice(c->device));
	SSB_CUDA(cudaStreamSynchronize(c->stream));
	*dptr = c->d_accum;
	if (count) *count = (size_t)c->accum_w * c->accum_h * 4;
	return SSB_OK;
}

int ssb_accum_merge(ssb_ctx* dst, ssb_ctx* src, const ssb_options* so) {
	if (!dst || !src || !so) return fail(SSB_ERR_ARG, "ssb_accum_merge: NULL argument");
	if (dst == src) return fail(SSB_ERR_ARG, "ssb_accum_merge: dst and src are the same context");
	if (!src->d_accum) return fail(SSB_ERR_ARG, "ssb_accum_merge: src has rendered nothing");
	uint32_t x1, y1, s1;
	int rc = validate_options(so, x1, y1, s1);
	if (rc != SSB_OK) return rc;
	if (src->accum_w != so->width || src->accum_h != so->height) return fail(SSB_ERR_ARG, "ssb_accum_merge: src_opt is not what src rendered (resolution)");
	SSB_CUDA(cudaSetDevice(dst->device));
	if ((rc = ensure_accum(dst, so->width, so->height)) != SSB_OK) return rc;
	const size_t count = (size_t)so->width * so->height * 4;
	// src's accumulator is complete once everything enqueued on its stream so far has run
	SSB_CUDA(cudaSetDevice(src->device));
	SSB_CUDA(cudaEventRecord(src->ev_accum_ready, src->stream));
	SSB_CUDA(cudaSetDevice(dst->device));
	SSB_CUDA(cudaStreamWaitEvent(dst->stream, src->ev_accum_ready, 0));
	const double* sp = src->d_accum;
	if (src->device != dst->device) {
		int can = 0;
		SSB_CUDA(cudaDeviceCanAccessPeer(&can, dst->device, src->device));
		if (can) {  // direct loads from the peer's HBM over NVLink
			cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
			if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); e = cudaSuccess; }
			if (e != cudaSuccess) can = 0;
		}
		if (!can) {  // no peer mapping (PCIe topology, MIG, ...): copy into a staging buffer on dst's device
			if (dst->peer_staging_count < count) {
				SSB_CUDA(cudaStreamSynchronize(dst->stream));
				cudaFree(dst->d_peer_staging); dst->d_peer_staging = nullptr; dst->peer_staging_count = 0;
				SSB_CUDA(cudaMalloc(&dst->d_peer_staging, count * sizeof(double)));
				dst->peer_staging_count = count;
			}
			SSB_CUDA(cudaMemcpyPeerAsync(dst->d_peer_staging, dst->device, src->d_accum, src->device, count * sizeof(double), dst->stream));
			sp = dst->d_peer_staging;
		}
	}
	const bool banded = so->band_count > 1;
	const bool subset = banded || so->x0 != 0 || x1 != so->width || so->y0 != 0 || y1 != so->height;
	const uint32_t mode = banded ? 2u : (subset ? 1u : 0u);
	const size_t npix = (size_t)so->width * so->height;
	ssb_accum_merge_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, dst->stream>>>(dst->d_accum, sp, so->width, so->height, mode, so->x0, so->y0, x1, y1,
	                                                                            banded ? so->band_height : 1u, banded ? so->band_count : 1u, banded ? so->band_index : 0u);
	SSB_CUDA(cudaGetLastError());
	// src must not overwrite its accumulator before dst has read it
	SSB_CUDA(cudaEventRecord(dst->ev_accum_ready, dst->stream));
	SSB_CUDA(cudaSetDevice(src->device));
	SSB_CUDA(cudaStreamWaitEvent(src->stream, dst->ev_accum_ready, 0));
	return SSB_OK;
}

static int resolve_impl(ssb_ctx* c, const ssb_options* o, double* xyza_host, float* srgba_host, bool device_only);

int ssb_resolve(ssb_ctx* c, const ssb_options* o, double* xyza_host, float* srgba_host) {
	return resolve_impl(c, o, xyza_host, srgba_host, false);
}
int ssb_resolve_device(ssb_ctx* c, const ssb_options* o, double** xyza_dev, float** srgba_dev) {
	int rc = resolve_impl(c, o, nullptr, nullptr, true);
	if (rc != SSB_OK) return rc;
	if (xyza_dev) *xyza_dev = c->d_xyza;
	if (srgba_dev) *srgba_dev = reinterpret_cast<float*>(c->d_srgba);
	return SSB_OK;
}
int ssb_set_stream(ssb_ctx* c, void* cuda_stream) {
	if (!c) return fail(SSB_ERR_ARG, "ssb_set_stream: NULL context");
	SSB_CUDA(cudaSetDevice(c->device));
	SSB_CUDA(cudaStreamSynchronize(c->stream));
	c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->own_stream;
	return SSB_OK;
}

static int resolve_impl(ssb_ctx* c, const ssb_options* o, double* xyza_host, float* srgba_host, bool device_only) {
	if (!c || !o) return fail(SSB_ERR_ARG, "ssb_resolve: NULL argument");
	if (!c->d_accum || c->accum_w != o->width || c->accum_h != o->height) return fail(SSB_ERR_ARG, "ssb_resolve: no accumulator for this resolution");
	if (o->spp == 0) return fail(SSB_ERR_ARG, "ssb_resolve: spp must be positive");
	SSB_CUDA(cudaSetDevice(c->device));
	size_t npix = (size_t)o->width * o->height;
	if (npix > c->resolve_capacity) {
		cudaFree(c->d_xyza); cudaFree(c->d_srgba); c->d_xyza = nullptr; c->d_srgba = nullptr; c->resolve_capacity = 0;
		SSB_CUDA(cudaMalloc(&c->d_xyza, npix * 4 * sizeof(double)));
		SSB_CUDA(cudaMalloc(&c->d_srgba, npix * sizeof(float4)));
		c->resolve_capacity = npix;
	}
	const float* m = c->xyz_to_lrgb;
	double scale = 1000.0 / (double)o->spp;  // renderer.cpp:296
	unsigned grid = (unsigned)((npix + 127) / 128);
	ssb_resolve_kernel<<<grid, 128, 0, c->stream>>>(c->d_accum, (device_only || xyza_host) ? c->d_xyza : nullptr, (device_only || srgba_host) ? c->d_srgba : nullptr,
	                                                (uint32_t)npix, scale, o->render_mode == SSB_RENDER_RGB ? (double)o->spp : 0.0, o->upsampling, c->d65_rad_Y,
	                                                m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8]);
	SSB_CUDA(cudaGetLastError());
	c->stats.launches += 1;
	if (device_only) return SSB_OK;
	if (xyza_host) SSB_CUDA(cudaMemcpyAsync(xyza_host, c->d_xyza, npix * 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	if (srgba_host) SSB_CUDA(cudaMemcpyAsync(srgba_host, c->d_srgba, npix * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
	SSB_CUDA(cudaStreamSynchronize(c->stream));
	return SSB_OK;
}

int ssb_render_frame(ssb_ctx* c, const ssb_options* o, double* xyza_host, float* srgba_host) {
	if (!c || !o) return fail(SSB_ERR_ARG, "ssb_render_frame: NULL argument");
	ssb_options full = *o;
	full.sample_begin = 0; full.sample_end = 0;
	int rc = ssb_render(c, &full);
	if (rc != SSB_OK) return rc;
	return ssb_resolve(c, &full, xyza_host, srgba_host);
}

int ssb_get_stats(ssb_ctx* c, ssb_stats* out) {
	if (!c || !out) return fail(SSB_ERR_ARG, "ssb_get_stats: NULL argument");
	if (c->stats_pending) {
		SSB_CUDA(cudaSetDevice(c->device));
		SSB_CUDA(cudaEventSynchronize(c->ev_end));
		float ms = 0;
		SSB_CUDA(cudaEventElapsedTime(&ms, c->ev_begin, c->ev_end));
		c->stats.device_ms = ms;
		double trace = 0;
		for (uint32_t p = 0; p < c->passes; ++p) { float t = 0; SSB_CUDA(cudaEventElapsedTime(&t, c->ev_pass[2 * p], c->ev_pass[2 * p + 1])); trace += t; }
		c->stats.trace_ms = trace;
		c->stats_pending = false;
	}
	*out = c->stats;
	return SSB_OK;
}
int ssb_synchronize(ssb_ctx* c) {
	if (!c) return fail(SSB_ERR_ARG, "ssb_synchronize: NULL context");
	SSB_CUDA(cudaSetDevice(c->device));
	SSB_CUDA(cudaStreamSynchronize(c->copy_stream));  // an ssb_upload_scene_async that no render has consumed yet
	SSB_CUDA(cudaStreamSynchronize(c->stream));
	return SSB_OK;
}

int ssb_debug_eval_math(ssb_ctx* c, uint32_t fn, const float* x_host, float arg, float* out_host, size_t n) {
	if (!c || !x_host || !out_host) return fail(SSB_ERR_ARG, "ssb_debug_eval_math: NULL argument");
	if (n == 0) return SSB_OK;
	SSB_CUDA(cudaSetDevice(c->device));
	if (fn == 8) {  // out[0] = lanes (of 2 x 2^32) in which the paired acosf differs from the scalar one, evaluated on the device;
		            // out[1..] = up to (n-1)/3 offending (argument, paired result, scalar result) triples
		unsigned long long* dbad = nullptr;
		float* dex = nullptr;
		const uint32_t nex = (uint32_t)((n - 1) / 3);
		SSB_CUDA(cudaMalloc(&dbad, 2 * sizeof(unsigned long long)));
		SSB_CUDA(cudaMemsetAsync(dbad, 0, 2 * sizeof(unsigned long long), c->stream));
		if (nex) { SSB_CUDA(cudaMalloc(&dex, 3 * nex * sizeof(float))); SSB_CUDA(cudaMemsetAsync(dex, 0, 3 * nex * sizeof(float), c->stream)); }
		ssb_acos_pair_check_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(dbad, dex, nex);
		unsigned long long bad = ~0ull;
		cudaError_t e = cudaStreamSynchronize(c->stream);
		if (e == cudaSuccess) e = cudaMemcpy(&bad, dbad, sizeof(bad), cudaMemcpyDeviceToHost);
		if (e == cudaSuccess && nex) e = cudaMemcpy(out_host + 1, dex, 3 * nex * sizeof(float), cudaMemcpyDeviceToHost);
		cudaFree(dbad); cudaFree(dex);
		if (e != cudaSuccess) return fail(SSB_ERR_DATA, "ssb_debug_eval_math: %s", cudaGetErrorString(e));
		out_host[0] = (float)bad;
		return SSB_OK;
	}
	float *dx = nullptr, *dout = nullptr;
	SSB_CUDA(cudaMalloc(&dx, n * sizeof(float)));
	SSB_CUDA(cudaMalloc(&dout, n * sizeof(float)));
	SSB_CUDA(copy_to_device(c, dx, x_host, n * sizeof(float)));
	ssb_eval_math_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(fn, dx, arg, dout, n);
	cudaError_t e = cudaStreamSynchronize(c->stream);
	if (e == cudaSuccess) e = cudaMemcpy(out_host, dout, n * sizeof(float), cudaMemcpyDeviceToHost);
	cudaFree(dx); cudaFree(dout);
	if (e != cudaSuccess) return fail(SSB_ERR_DATA, "ssb_debug_eval_math: %s", cudaGetErrorString(e));
	return SSB_OK;
}

int ssb_debug_intersect(ssb_ctx* c, const float* rays, const int32_t* ignore, uint32_t scan_mode, float eps, float* out, size_t n) {
	if (!c || !rays || !out) return fail(SSB_ERR_ARG, "ssb_debug_intersect: NULL argument");
	if (scan_mode > SSB_SCAN_LIST) return fail(SSB_ERR_UNSUPPORTED, "unknown scan mode %u", scan_mode);
	if (n == 0) return SSB_OK;
	SSB_CUDA(cudaSetDevice(c->device));
	int rc;
	if (c->blob_dirty && (rc = build_blob(c)) != SSB_OK) return rc;
	float *d_rays = nullptr, *d_out = nullptr;
	int32_t* d_ign = nullptr;
	cudaError_t e = cudaMalloc(&d_rays, n * 6 * sizeof(float));
	if (e == cudaSuccess) e = cudaMalloc(&d_out, n * 6 * sizeof(float));
	if (e == cudaSuccess && ignore) e = cudaMalloc(&d_ign, n * sizeof(int32_t));
	if (e == cudaSuccess) e = copy_to_device(c, d_rays, rays, n * 6 * sizeof(float));
	if (e == cudaSuccess && ignore) e = copy_to_device(c, d_ign, ignore, n * sizeof(int32_t));
	if (e == cudaSuccess) {
		KParams P{};
		P.blob = c->d_blob; P.eps = eps; P.scan_list = scan_mode == SSB_SCAN_LIST ? 1u : 0u;
		e = raise_dynamic_smem_limit(c->device, (const void*)ssb_debug_intersect_kernel, c->blob_bytes);
		if (e == cudaSuccess) {
			const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)c->sm_count * 4);
			ssb_debug_intersect_kernel<<<grid, 256, c->blob_bytes, c->stream>>>(P, d_rays, d_ign, d_out, n);
			e = cudaGetLastError();
		}
	}
	if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, n * 6 * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	cudaFree(d_rays); cudaFree(d_out); cudaFree(d_ign);
	if (e != cudaSuccess) return fail(SSB_ERR_DATA, "ssb_debug_intersect: %s", cudaGetErrorString(e));
	return SSB_OK;
}

int ssb_debug_trace_samples(ssb_ctx* c, const ssb_options* o, uint32_t px, uint32_t py, float* out_host) {
	if (!c || !o || !out_host) return fail(SSB_ERR_ARG, "ssb_debug_trace_samples: NULL argument");
	ssb_options one = *o;
	one.x0 = px; one.x1 = px + 1; one.y0 = py; one.y1 = py + 1;
	// keep the accumulator untouched: render into the sample buffer only by using a scratch accumulator pass
	uint32_t x1, y1, s1;
	int rc = validate_options(&one, x1, y1, s1);
	if (rc != SSB_OK) return rc;
	std::vector<double> saved;
	bool had = c->d_accum && c->accum_w == o->width && c->accum_h == o->height;
	if (had) { saved.resize((size_t)o->width * o->height * 4); if ((rc = ssb_read_accum(c, saved.data())) != SSB_OK) return rc; }
	uint32_t begin = one.sample_begin;
	rc = ssb_render(c, &one);  // the per-sample values of the (last) pass stay in d_samples until the next render
	if (rc != SSB_OK) return rc;
	uint32_t ns = s1 - begin;
	if ((size_t)ns > c->samples_capacity) return fail(SSB_ERR_UNSUPPORTED, "too many samples for one pass");
	SSB_CUDA(cudaStreamSynchronize(c->stream));
	SSB_CUDA(cudaMemcpy(out_host, c->d_samples, (size_t)ns * sizeof(float4), cudaMemcpyDeviceToHost));
	if (had) rc = ssb_write_accum(c, saved.data());
	return rc;
}

}  // extern "C"
