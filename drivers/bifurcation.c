/* Arbitrary geometry, velocity inlet / pressure outlet: drop-in for bifurcation/bifurcation.cu
 * (main: bif:1177-1326).  Reads ./geo.txt and ./bc.txt, writes ./out/bif_<t>.vtk, ./out/CONVERGENCE.log
 * and ./meas1.txt. */
#include "common.h"

/* write_once(): bif:1055-1075 -- u_y then u_x (lattice units) of every (y,x) of the plane z = NZ/2,
 * space separated, ostream default format (= %g).  The reference indexes h_uy[-1] for nodes that
 * are not stored; 0 is written for those. */
static int write_once(lbm_handle h, const lbm_case_desc *d, int64_t nlattice) {
    const size_t cells = (size_t)d->nx * d->ny * d->nz, rsz = d->precision == LBM_F64 ? 8 : 4;
    int32_t *index = (int32_t *)malloc(cells * sizeof(int32_t));
    char *ux = (char *)malloc((size_t)nlattice * rsz), *uy = (char *)malloc((size_t)nlattice * rsz);
    FILE *f = fopen("./meas1.txt", "w");
    int rc = -1;
    if (index && ux && uy && f && !lbm_get_index(h, index) && !lbm_get_fields(h, NULL, ux, uy, NULL, NULL, NULL)) {
        const int z = d->nz / 2;
        for (int pass = 0; pass < 2; pass++) {
            const char *src = pass == 0 ? uy : ux;
            for (int y = 0; y < d->ny; y++)
                for (int x = 0; x < d->nx; x++) {
                    const int32_t idx = index[(size_t)x + (size_t)d->nx * ((size_t)y + (size_t)d->ny * z)];
                    double v = 0.0;
                    if (idx >= 0) v = rsz == 8 ? ((const double *)src)[idx] : (double)((const float *)src)[idx];
                    fprintf(f, "%g ", rsz == 8 ? v : (double)(float)v);
                }
        }
        rc = 0;
    }
    if (f) fclose(f);
    free(index), free(ux), free(uy);
    return rc;
}

int main(int argc, char **argv) {
    lbm_case_desc d;
    lbm_handle h = NULL;
    int repeat = 4400, time_save = 4400; /* REPEAT, time_save: bif:19 */
    lbm_case_defaults(LBM_CASE_GEO_Y_INOUT, &d);
    d.storage = LBM_STORE_SPARSE_AA; /* --storage overrides */
    if (parse_common(argc, argv, &d, &repeat, &time_save)) return 2;
    CHECK(h, lbm_create(&d, &h));
    CHECK(h, lbm_set_output_format(h, g_out_format));
    int64_t nlattice = 0;
    CHECK(h, lbm_geo_pre(h));                    /* geo_pre();   bif:1187 (labels, -1 marking) */
    CHECK(h, lbm_index_transform(h, &nlattice)); /*              bif:241-252 */
    CHECK(h, lbm_read_vel(h));                   /* read_vel();  bif:1223 */
    CHECK(h, lbm_initialize(h));                 /* initialize();bif:1224 */
    CHECK(h, lbm_run_fixed(h, repeat, time_save, 1)); /* for(i=0;i<=REPEAT;i++){...}  bif:1246-1274 */
    if (write_once(h, &d, nlattice)) {               /* write_once();                 bif:1277 */
        fprintf(stderr, "cannot write ./meas1.txt: %s\n", lbm_last_error(h));
        lbm_destroy(h);
        return 1;
    }
    lbm_destroy(h);
    return 0;
}
