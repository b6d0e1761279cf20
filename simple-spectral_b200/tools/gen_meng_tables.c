/* gen_meng_tables.c — build-time tool.  The reference compiles Meng et al.'s spectral grid into its
 * binary as static C tables (src/meng-et-al.-2015/spectra_xyz_5nm_380_780_0.97.h, third-party data
 * vendored by the reference).  We do not copy that header: this tool is compiled against it where it
 * lies (-I/root/reference/src/meng-et-al.-2015) and serialises the tables to a small binary file that
 * the host layer loads at run time (ssb_meng_tables in include/ssb200.h):
 *   char magic[8] = "SSBMENG1"; uint32 grid_w, grid_h, npoints, nsamples; float xy_to_uv[6];
 *   float sample_min, sample_max; int32 grid[grid_w*grid_h*8]; float points[npoints*(5+nsamples)]
 *   (per point: xystar[2], 0, uv[2], spectrum[nsamples]).
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "spectra_xyz_5nm_380_780_0.97.h"

int main(int argc, char** argv) {
	if (argc != 2) { fprintf(stderr, "usage: %s out.bin\n", argv[0]); return 2; }
	FILE* f = fopen(argv[1], "wb");
	if (!f) { perror("fopen"); return 1; }
	uint32_t ncells = sizeof(spectrum_grid) / sizeof(spectrum_grid[0]);
	uint32_t npoints = sizeof(spectrum_data_points) / sizeof(spectrum_data_points[0]);
	uint32_t hdr[4] = { (uint32_t)spectrum_grid_width, (uint32_t)spectrum_grid_height, npoints, (uint32_t)spectrum_num_samples };
	if (ncells != hdr[0] * hdr[1]) { fprintf(stderr, "unexpected grid size\n"); return 1; }
	fwrite("SSBMENG1", 1, 8, f);
	fwrite(hdr, 4, 4, f);
	fwrite(spectrum_mat_xy_to_uv, 4, 6, f);
	float mm[2] = { spectrum_sample_min, spectrum_sample_max };
	fwrite(mm, 4, 2, f);
	for (uint32_t c = 0; c < ncells; ++c) {
		int32_t rec[8] = { spectrum_grid[c].inside, spectrum_grid[c].num_points };
		for (int i = 0; i < 6; ++i) rec[2 + i] = spectrum_grid[c].idx[i];
		fwrite(rec, 4, 8, f);
	}
	for (uint32_t p = 0; p < npoints; ++p) {
		float head[5] = { spectrum_data_points[p].xystar[0], spectrum_data_points[p].xystar[1], 0.0f,
		                  spectrum_data_points[p].uv[0], spectrum_data_points[p].uv[1] };
		fwrite(head, 4, 5, f);
		fwrite(spectrum_data_points[p].spectrum, 4, hdr[3], f);
	}
	fclose(f);
	return 0;
}
