/* Arbitrary geometry, velocity inlet / pressure outlet: drop-in for bifurcation/bifurcation.cu
 * (main: bif:1177-1326).  Reads ./geo.txt and ./bc.txt, writes ./out/bif_<t>.vtk, ./out/CONVERGENCE.log. */
#include "common.h"

int main(int argc, char **argv) {
    lbm_case_desc d;
    lbm_handle h = NULL;
    int repeat = 4400, time_save = 4400; /* REPEAT, time_save: bif:19 */
    lbm_case_defaults(LBM_CASE_GEO_Y_INOUT, &d);
    if (parse_common(argc, argv, &d, &repeat, &time_save)) return 2;
    CHECK(h, lbm_create(&d, &h));
    CHECK(h, lbm_set_output_format(h, g_out_format));
    int64_t nlattice = 0;
    CHECK(h, lbm_geo_pre(h));                    /* geo_pre();   bif:1187 (labels, -1 marking) */
    CHECK(h, lbm_index_transform(h, &nlattice)); /*              bif:241-252 */
    CHECK(h, lbm_read_vel(h));                   /* read_vel();  bif:1223 */
    CHECK(h, lbm_initialize(h));                 /* initialize();bif:1224 */
    CHECK(h, lbm_run_fixed(h, repeat, time_save, 1)); /* for(i=0;i<=REPEAT;i++){...}  bif:1246-1274 */
    lbm_destroy(h);
    return 0;
}
