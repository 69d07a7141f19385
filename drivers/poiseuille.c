/* Poiseuille flow in a circular pipe: drop-in for Poiseulle_flow/Poiseulle.cu (main: pos:940-1056). */
#include "common.h"

int main(int argc, char **argv) {
    lbm_case_desc d;
    lbm_handle h = NULL;
    int max_it = 10000, time_save = 500; /* pos:944 */
    lbm_case_defaults(LBM_CASE_POISEUILLE, &d);
    d.storage = LBM_STORE_SPARSE_AA; /* --storage overrides */
    if (parse_common(argc, argv, &d, &max_it, &time_save)) return 2;
    CHECK(h, lbm_create(&d, &h));
    CHECK(h, lbm_set_output_format(h, g_out_format));
    int64_t nlattice = 0;
    CHECK(h, lbm_geo_pre(h));                    /* geo_pre();         pos:950 */
    CHECK(h, lbm_index_transform(h, &nlattice)); /* index_transform(); pos:951 */
    CHECK(h, lbm_initialize(h));                 /* initialize();      pos:975 */
    int32_t its = 0;
    double res = 0;
    CHECK(h, lbm_run_converge(h, max_it, 1e-6, 50, time_save, 1, &its, &res)); /* pos:986-1019 */
    lbm_destroy(h);
    return 0;
}
