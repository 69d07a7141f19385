/* Lid-driven cavity on SEVERAL GPUs: the main() of Lid_driven_cavity/ldc.cu (ldc.cu:612-717) with the box
 * cut into z-slabs, one per device (the reference is single-GPU; lbm_create_distributed is the multi-GPU
 * form of its malloc/cudaMalloc block).  Same files as the single-GPU driver: ./out/lid_<k>.vtk every 500
 * iterations -- ONE file per save, byte-identical to the single-domain writer -- and ./out/CONVERGENCE.log;
 * the residual S is the sum of the slabs' shares (ldc.cu:660-668).
 *   ldc_mgpu [--slabs P] [--devices a,b,...] [--n N | --dims NX NY NZ] [--f64] [--steps K] [--save S] ...
 * Without --devices slab r runs on device r modulo the device count. */
#include "common.h"

#define GCHECK(g, call)                                                                  \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ != LBM_OK) {                                                             \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, lbm_group_last_error(g)); \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

int main(int argc, char **argv) {
    lbm_case_desc d;
    lbm_group g = NULL;
    int max_it = 10000, time_save = 500, nslabs = 2, ndev = 0; /* ldc.cu:615 */
    int32_t devices[64];
    /* strip the two options of this driver, hand the rest to the common parser */
    char *rest[64];
    int nrest = 0;
    rest[nrest++] = argv[0];
    for (int i = 1; i < argc && nrest < 63; i++) {
        if (!strcmp(argv[i], "--slabs") && i + 1 < argc) nslabs = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--devices") && i + 1 < argc) {
            for (char *t = strtok(argv[++i], ","); t && ndev < 64; t = strtok(NULL, ",")) devices[ndev++] = atoi(t);
        } else rest[nrest++] = argv[i];
    }
    lbm_case_defaults(LBM_CASE_LDC, &d);
    d.storage = LBM_STORE_DENSE_AA; /* one population buffer, streamed in place */
    if (parse_common(nrest, rest, &d, &max_it, &time_save)) return 2;
    if (ndev && ndev != nslabs) {
        fprintf(stderr, "--devices lists %d devices for %d slabs\n", ndev, nslabs);
        return 2;
    }
    GCHECK(g, lbm_create_distributed(&d, nslabs, ndev ? devices : NULL, &g));
    GCHECK(g, lbm_group_set_output_format(g, g_out_format));
    int64_t nlattice = 0;
    GCHECK(g, lbm_group_setup(g, NULL, NULL, NULL, &nlattice)); /* geo_pre(); initialize(); on every slab */
    int32_t its = 0;
    double res = 0;
    GCHECK(g, lbm_group_run_converge(g, max_it, 1e-6, 50, time_save, 1, &its, &res));
    lbm_group_destroy(g);
    return 0;
}
