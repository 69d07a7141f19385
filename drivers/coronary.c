/* Coronary tree, one inlet / one main outlet / three sub-outlets: drop-in for coronary_cfd/coronary.cu
 * (main: cor:1053-1163).  Reads ./geo.txt (y fastest, cor:45-56); the reference does not ship that file. */
#include "common.h"

int main(int argc, char **argv) {
    lbm_case_desc d;
    lbm_handle h = NULL;
    int repeat = 300000, time_save = 5000; /* cor:19 */
    lbm_case_defaults(LBM_CASE_GEO_OPENINGS, &d);
    d.storage = LBM_STORE_SPARSE_AA; /* --storage overrides */
    if (parse_common(argc, argv, &d, &repeat, &time_save)) return 2;
    CHECK(h, lbm_create(&d, &h));
    CHECK(h, lbm_set_output_format(h, g_out_format));
    int64_t nlattice = 0;
    CHECK(h, lbm_geo_pre(h));
    CHECK(h, lbm_index_transform(h, &nlattice));
    CHECK(h, lbm_initialize(h));
    CHECK(h, lbm_run_fixed(h, repeat, time_save, 1)); /* cor:1100-1132 */
    lbm_destroy(h);
    return 0;
}
