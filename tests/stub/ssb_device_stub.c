/* ssb_device_stub.c — TEST INFRASTRUCTURE ONLY: the device-layer entry points of include/ssb200.h that the C++ host
 * layer (simple-spectral_b200/csrc/host) calls, answered by the CPU oracle (oracle/ssb_oracle.c) instead of the GPU.
 * tests/test_host_renderer_stub.py links it with the host layer's sources so that the Renderer facade's own logic
 * (worker thread, progressive sample slices, abort, error propagation, image saving) is exercised on a machine
 * without a GPU.  It is never part of libssb200.so; the product has no CPU path.
 *
 * Unlike the real library the stub copies only the top-level structs and BORROWS the arrays they point to (the
 * Renderer keeps its Scene / ColorData alive for its whole life). */
#include <stdlib.h>
#include <string.h>

#include "../../include/ssb200.h"
#include "../../oracle/ssb_oracle.h"

struct ssb_ctx {
	ssb_scene scene;
	ssb_color color;
	double* accum;
	uint32_t w, h;
	ssb_stats stats;
	uint32_t renders; /* ssb_render calls so far (the test reads it through ssb_stats.reserved) */
};

static const char* g_err = "";

uint32_t ssb_abi_version(void) { return SSB_ABI_VERSION; }
const char* ssb_last_error(void) { return g_err; }

void ssb_default_options(ssb_options* o, uint32_t width, uint32_t height, uint32_t spp) { /* stdafx.hpp:44-90 */
	memset(o, 0, sizeof(*o));
	o->width = width; o->height = height; o->spp = spp;
	o->upsampling = SSB_UPSAMPLE_OURS;
	o->lambda_min = 380.0f; o->lambda_max = 780.0f;
	o->max_depth = 10; o->explicit_light_sampling = 1; o->flat_field_correction = 1;
	o->eps = 0.001f; o->seed = 1;
}

#define STUB_DEVICES 4 /* the stub pretends to be a four-GPU box */
int ssb_device_count(int* count) { *count = STUB_DEVICES; return SSB_OK; }
int ssb_create(int device, ssb_ctx** out) {
	if (device < 0 || device >= STUB_DEVICES) { g_err = "stub: device out of range"; return SSB_ERR_ARG; }
	*out = (ssb_ctx*)calloc(1, sizeof(ssb_ctx));
	return SSB_OK;
}
void ssb_destroy(ssb_ctx* c) { if (c) { free(c->accum); free(c); } }
int ssb_upload_scene(ssb_ctx* c, const ssb_scene* s) { c->scene = *s; return SSB_OK; }
int ssb_upload_color(ssb_ctx* c, const ssb_color* col) { c->color = *col; return SSB_OK; }

int ssb_render(ssb_ctx* c, const ssb_options* o) {
	if (o->spp == 0x7fffffffu) { g_err = "stub: injected render failure"; return SSB_ERR_DATA; }
	if (!c->accum || c->w != o->width || c->h != o->height) {
		free(c->accum);
		c->accum = (double*)calloc((size_t)o->width * o->height * 4, sizeof(double));
		c->w = o->width; c->h = o->height;
	}
	if (o->sample_begin == 0 && !o->keep_accumulator) memset(c->accum, 0, (size_t)c->w * c->h * 4 * sizeof(double));
	int rc = ssb_oracle_render(&c->scene, &c->color, o, c->accum, NULL, NULL);
	if (rc != SSB_OK) { g_err = "stub: oracle render failed"; return rc; }
	uint32_t s1 = o->sample_end ? o->sample_end : o->spp;
	memset(&c->stats, 0, sizeof(c->stats));
	{ /* pixels of the requested subset (rectangle, or this share of the interleaved row bands) */
		uint32_t x1 = o->x1 ? o->x1 : o->width, y1 = o->y1 ? o->y1 : o->height;
		uint64_t rows = 0;
		for (uint32_t j = o->y0; j < y1; ++j)
			if (o->band_count <= 1 || (j / o->band_height) % o->band_count == o->band_index) ++rows;
		c->stats.samples = rows * (x1 - o->x0) * (s1 - o->sample_begin);
	}
	c->stats.device_ms = 1.0; c->stats.trace_ms = 1.0; c->stats.launches = 1;
	c->stats.reserved = ++c->renders;
	return SSB_OK;
}

int ssb_resolve(ssb_ctx* c, const ssb_options* o, double* xyza, float* srgba) {
	if (!c->accum) { g_err = "stub: nothing rendered"; return SSB_ERR_ARG; }
	return ssb_oracle_resolve(&c->color, o, c->accum, xyza, srgba);
}

int ssb_get_stats(ssb_ctx* c, ssb_stats* out) { *out = c->stats; return SSB_OK; }

int ssb_clear(ssb_ctx* c) {
	if (c->accum) memset(c->accum, 0, (size_t)c->w * c->h * 4 * sizeof(double));
	return SSB_OK;
}

/* same rule as the library: a pixel subset is copied, a whole-frame share (sample range) is added */
int ssb_accum_merge(ssb_ctx* dst, ssb_ctx* src, const ssb_options* o) {
	if (!src->accum || src->w != o->width || src->h != o->height) { g_err = "stub: merge source mismatch"; return SSB_ERR_ARG; }
	if (!dst->accum || dst->w != o->width || dst->h != o->height) {
		free(dst->accum);
		dst->accum = (double*)calloc((size_t)o->width * o->height * 4, sizeof(double));
		dst->w = o->width; dst->h = o->height;
	}
	uint32_t x1 = o->x1 ? o->x1 : o->width, y1 = o->y1 ? o->y1 : o->height;
	int banded = o->band_count > 1;
	int subset = banded || o->x0 != 0 || x1 != o->width || o->y0 != 0 || y1 != o->height;
	for (uint32_t j = 0; j < o->height; ++j)
		for (uint32_t i = 0; i < o->width; ++i) {
			size_t p = 4 * ((size_t)j * o->width + i);
			if (!subset) { for (int k = 0; k < 4; ++k) dst->accum[p + k] += src->accum[p + k]; continue; }
			int mine = banded ? (i >= o->x0 && i < x1 && (j / o->band_height) % o->band_count == o->band_index) : (i >= o->x0 && i < x1 && j >= o->y0 && j < y1);
			if (mine) memcpy(dst->accum + p, src->accum + p, 4 * sizeof(double));
		}
	return SSB_OK;
}
