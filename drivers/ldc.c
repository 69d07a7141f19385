/* Lid-driven cavity: drop-in for Lid_driven_cavity/ldc.cu (main: ldc.cu:612-717).
 * Writes ./out/lid_<k>.vtk every 500 iterations and ./out/CONVERGENCE.log. */
#include "common.h"

int main(int argc, char **argv) {
    lbm_case_desc d;
    lbm_handle h = NULL;
    int max_it = 10000, time_save = 500; /* ldc.cu:615 */
    lbm_case_defaults(LBM_CASE_LDC, &d);
    d.storage = LBM_STORE_SPARSE_AA; /* --storage overrides */
    if (parse_common(argc, argv, &d, &max_it, &time_save)) return 2;
    CHECK(h, lbm_create(&d, &h));
    CHECK(h, lbm_set_output_format(h, g_out_format));
    int64_t nlattice = 0;
    CHECK(h, lbm_geo_pre(h));                      /* geo_pre();    ldc.cu:645 */
    CHECK(h, lbm_index_transform(h, &nlattice));   /* (ldc stores the whole box, ldc.cu:54) */
    CHECK(h, lbm_initialize(h));                   /* initialize(); ldc.cu:646 */
    int32_t its = 0;
    double res = 0;
    /* while(k<=max_it&&tol_count<=stag_max){ update; boundary_stream; calc_vel_square; reduce; ... }  ldc.cu:653-685 */
    CHECK(h, lbm_run_converge(h, max_it, 1e-6, 50, time_save, 1, &its, &res));
    lbm_destroy(h);
    return 0;
}
