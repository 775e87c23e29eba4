// line.cu -- TEMPORARY stubs (replaced by the LSD/LBD kernels).
#include "common.cuh"
struct sdpl_line { int dummy; };
extern "C" {
int sdpl_line_create(sdpl_line**, int, int, float, int, float, int, int) { sdpl::set_last_error("line extractor not built yet"); return SDPL_ERR_UNSUPPORTED; }
void sdpl_line_destroy(sdpl_line*) {}
int sdpl_line_levels(const sdpl_line*) { return 0; }
int sdpl_line_tables(const sdpl_line*, float*, float*, float*, float*) { return SDPL_ERR_UNSUPPORTED; }
int sdpl_line_extract(sdpl_line*, const uint8_t*, int, int, int, sdpl_keyline*, uint8_t*, int, int*) { return SDPL_ERR_UNSUPPORTED; }
int sdpl_line_extract_batch(sdpl_line*, const uint8_t*, int, int, int, int, size_t, sdpl_keyline*, uint8_t*, int, int*) { return SDPL_ERR_UNSUPPORTED; }
int sdpl_line_extract_batch_dev(sdpl_line*, const uint8_t*, int, int, int, int, size_t, sdpl_keyline*, uint8_t*, int, int*, int) { return SDPL_ERR_UNSUPPORTED; }
int sdpl_line_lbd_compute(sdpl_line*, const uint8_t*, int, int, int, const sdpl_keyline*, int, uint8_t*) { return SDPL_ERR_UNSUPPORTED; }
int sdpl_line_lsd_segments(sdpl_line*, int, int, float*, int, int*) { return SDPL_ERR_UNSUPPORTED; }
int sdpl_line_last_launches(const sdpl_line*) { return 0; }
int sdpl_line_set_stream(sdpl_line*, void*) { return SDPL_ERR_UNSUPPORTED; }
int sdpl_line_set_profiling(sdpl_line*, int) { return SDPL_ERR_UNSUPPORTED; }
int sdpl_line_stage_times(sdpl_line*, float*, const char**, int*, int) { return 0; }
}
