// TEST TOOL (not part of the product): exhaustive check of the libm restatements in
// simple-spectral_b200/csrc/ssb_math.cuh against the host libm the reference links (glibc), over ALL 2^32 float bit
// patterns per function.  The header compiles as plain C++ (its SSB_HD functions are `inline` on the host and use the
// same fma()/operation order as on the device, where the toolkit's fma() is the IEEE fused operation as well).
//   g++ -O2 -fopenmp -ffp-contract=off -I simple-spectral_b200/csrc tools/check_math_exhaustive.cpp -o /tmp/check_math && /tmp/check_math
// Prints the mismatch count per function (expected: 0 everywhere); exit code 1 on any mismatch.
// The device build of the same header is compared against libm on random samples by tests/test_gpu_parity.py.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "ssb_math.cuh"

static inline float f_of(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t u_of(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline bool same(float a, float b) { return u_of(a) == u_of(b) || (a != a && b != b); }

template <class F, class G> static unsigned long long sweep(const char* name, F ours, G ref) {
	unsigned long long bad = 0;
	uint32_t first = 0;
	bool have = false;
#pragma omp parallel for schedule(static) reduction(+ : bad)
	for (long long i = 0; i < (1ll << 32); ++i) {
		const float x = f_of((uint32_t)i);
		if (!same(ours(x), ref(x))) {
			++bad;
#pragma omp critical
			if (!have) { have = true; first = (uint32_t)i; }
		}
	}
	std::printf("%-28s mismatches: %llu", name, bad);
	if (bad) std::printf("   (first at bits 0x%08x = %g)", first, (double)f_of(first));
	std::printf("\n");
	std::fflush(stdout);
	return bad;
}

int main() {
	unsigned long long bad = 0;
	bad += sweep("sinf", [](float x) { return ssbm::sinf_exact(x); }, [](float x) { return ::sinf(x); });
	bad += sweep("cosf", [](float x) { return ssbm::cosf_exact(x); }, [](float x) { return ::cosf(x); });
	bad += sweep("sincosf (sin)", [](float x) { float s, c; ssbm::sincosf_exact(x, &s, &c); return s; }, [](float x) { return ::sinf(x); });
	bad += sweep("sincosf (cos)", [](float x) { float s, c; ssbm::sincosf_exact(x, &s, &c); return c; }, [](float x) { return ::cosf(x); });
	bad += sweep("acosf", [](float x) { return ssbm::acosf_exact(x); }, [](float x) { return ::acosf(x); });
	bad += sweep("powf(x, 2.4f)", [](float x) { return ssbm::powf_exact(x, 2.4f); }, [](float x) { return ::powf(x, 2.4f); });
	bad += sweep("powf(x, 1/2.4f)", [](float x) { return ssbm::powf_exact(x, 1.0f / 2.4f); }, [](float x) { return ::powf(x, 1.0f / 2.4f); });
	std::printf(bad ? "FAILED\n" : "all functions bit-identical to libm over all 2^32 inputs\n");
	return bad ? 1 : 0;
}
