/* Shared helpers of the four case drivers.  Each driver reads like the reference program's
 * main() (bifurcation.cu:1177-1326): geo_pre(); index_transform(); read_vel(); initialize();
 * the time loop with periodic outputSave(); -- but every step runs in liblbm_b200.so. */
#ifndef LBM_DRIVER_COMMON_H
#define LBM_DRIVER_COMMON_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/lbm_b200.h"

#define CHECK(h, call)                                                                  \
    do {                                                                                \
        int rc_ = (call);                                                               \
        if (rc_ != LBM_OK) {                                                            \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, lbm_last_error(h));     \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

/* optional overrides; the reference has none (argc/argv unused, ldc.cu:612):
 *   --n N | --dims NX NY NZ | --f64 | --strict | --steps K | --save S | --tau T | --out DIR | --device D
 *   --storage ab|aa|sparse|sparse_aa   (default of the drivers: sparse_aa, one population buffer over the fluid nodes) */
static int g_out_format = LBM_OUT_ASCII_VTK; /* --binary: legacy-VTK BINARY dumps instead of the reference's ASCII */
static int parse_common(int argc, char **argv, lbm_case_desc *d, int *steps, int *save) {
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--n") && i + 1 < argc) {
            d->nx = d->ny = d->nz = atoi(argv[++i]);
            d->z_begin = 0, d->z_end = d->nz;
        } else if (!strcmp(argv[i], "--dims") && i + 3 < argc) {
            d->nx = atoi(argv[++i]), d->ny = atoi(argv[++i]), d->nz = atoi(argv[++i]);
            d->z_begin = 0, d->z_end = d->nz;
        } else if (!strcmp(argv[i], "--f64")) d->precision = LBM_F64;
        else if (!strcmp(argv[i], "--strict")) d->math = LBM_MATH_STRICT;
        else if (!strcmp(argv[i], "--binary")) g_out_format = LBM_OUT_BINARY_VTK;
        else if (!strcmp(argv[i], "--storage") && i + 1 < argc) {
            const char *v = argv[++i];
            d->storage = !strcmp(v, "ab") ? LBM_STORE_DENSE_AB : !strcmp(v, "aa") ? LBM_STORE_DENSE_AA
                         : !strcmp(v, "sparse") ? LBM_STORE_SPARSE_AB : LBM_STORE_SPARSE_AA;
        }
        else if (!strcmp(argv[i], "--steps") && i + 1 < argc) *steps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--save") && i + 1 < argc) *save = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--tau") && i + 1 < argc) d->tau = atof(argv[++i]);
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) d->device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--out") && i + 1 < argc) snprintf(d->out_dir, sizeof d->out_dir, "%s", argv[++i]);
        else {
            fprintf(stderr, "unknown option %s\n", argv[i]);
            return 1;
        }
    }
    return 0;
}
#endif
