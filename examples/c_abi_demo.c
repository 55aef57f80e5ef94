/* A plain C99 client of the C-ABI (include/smrfb.h): one IDW handle, one detrended distribute call with the clamp fused, the
 * precipitation phase consumer.  This is what a cgo / JNI / Cython binding of the reference would wrap; the Python classes in
 * smrf_b200/ do the same through ctypes.
 *     gcc -std=c99 -Iinclude examples/c_abi_demo.c -Lsmrf_b200 -lsmrfb -Wl,-rpath,$PWD/smrf_b200 -lm -o c_abi_demo && ./c_abi_demo
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "smrfb.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        int rc_ = (call);                                                        \
        if (rc_ != SMRFB_OK) {                                                   \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, smrfb_last_error());   \
            return 1;                                                            \
        }                                                                        \
    } while (0)

int main(void) {
    enum { NY = 64, NX = 96, S = 5, T = 3 };
    static double x[NX], y[NY], dem[NY * NX], out[T * NY * NX], pcs[NY * NX], rho[NY * NX], pp[NY * NX];
    const double mx[S] = {310.0, 1500.0, 2600.0, 700.0, 2100.0}, my[S] = {1700.0, 400.0, 1200.0, 900.0, 1850.0};
    const double mz[S] = {1800.0, 2100.0, 1950.0, 2300.0, 1750.0};
    double data[T * S], pv[T * 2];
    int i, j, t, ndev = 0;
    smrfb_handle_t h = NULL;
    for (j = 0; j < NX; j++) x[j] = 30.0 * j;
    for (i = 0; i < NY; i++) y[i] = 30.0 * (NY - 1 - i);                 /* descending, as in a topo file */
    for (i = 0; i < NY; i++)
        for (j = 0; j < NX; j++) dem[i * NX + j] = 1900.0 + 200.0 * sin(x[j] / 700.0) * cos(y[i] / 500.0);
    for (t = 0; t < T; t++)
        for (j = 0; j < S; j++) data[t * S + j] = 12.0 - 0.0065 * mz[j] + t;   /* air temperature with a lapse rate */
    data[1 * S + 2] = NAN;                                                /* a missing station at step 1 */
    CHECK(smrfb_device_count(&ndev));
    printf("libsmrfb %d, %d CUDA device(s)\n", smrfb_version(), ndev);
    CHECK(smrfb_idw_create(S, mx, my, mz, NY, NX, 0, x, y, dem, 2.0, 0, &h));
    /* detrendedIDW(data, flag = -1) for T steps, clamped to [-73, 47] in the kernel's store (air_temp.py:73-86) */
    CHECK(smrfb_idw_distribute(h, data, T, 1, -1, -73.0, 47.0, NULL, pv, out));
    for (t = 0; t < T; t++) printf("step %d: slope %.6f intercept %.3f  out[0] %.4f\n", t, pv[2 * t], pv[2 * t + 1], out[t * NY * NX]);
    /* percent snow / new-snow density from that temperature and a uniform 2 mm of precipitation (envphys/snow.py:37-75) */
    for (i = 0; i < NY * NX; i++) pp[i] = 2.0;
    CHECK(smrfb_snow_phase(out, pp, NY * NX, 0, pcs, rho, 0));
    printf("susong1999 at cell 0: %.2f snow, %.0f kg/m^3\n", pcs[0], rho[0]);
    CHECK(smrfb_destroy(h));
    return 0;
}
