/* The header is plain C and the library is usable without Python or torch: compiled with gcc and linked
 * against liblec_b200.so by tests/test_host_helpers.py.  Prints one line per check; exit code 0 = all passed.
 * With a GPU (argv[1] = "gpu") it also runs a tiny fixed-box case through lec_run_host and through
 * lec_run_host_raw on the same data stored upside-down, and compares the two. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lec_b200.h"

#define NLON 16
#define NLAT 9
#define NLEV 4
#define NT 3

static int fail(const char *what) { printf("FAIL %s\n", what); return 1; }

int main(int argc, char **argv) {
  int with_gpu = argc > 1 && strcmp(argv[1], "gpu") == 0;
  printf("version %s\n", lec_version());
  if (sizeof(lec_step) != 56) return fail("sizeof(lec_step)");
  if (LEC_NTERMS != 16 || LEC_NLEVEL_TERMS != 19) return fail("term counts");

  double lat[NLAT], lon[NLON], rlat[NLAT], rlon[NLON], coslat[NLAT], plev[NLEV] = {20000, 50000, 85000, 100000};
  for (int j = 0; j < NLAT; ++j) { lat[j] = -40.0 + 2.5 * j; rlat[j] = lat[j] * M_PI / 180.0; coslat[j] = cos(rlat[j]); }
  for (int i = 0; i < NLON; ++i) { lon[i] = -60.0 + 2.5 * i; rlon[i] = lon[i] * M_PI / 180.0; }
  if (lec_nearest_index(lat, NLAT, -31.25) != 4) return fail("nearest tie -> larger coordinate");   /* -30.0 */
  double a[NLEV], b[NLEV], c[NLEV];
  if (lec_gradient_coefs(plev, NLEV, a, b, c) != LEC_OK) return fail("gradient_coefs");
  if (fabs(c[0] - 1.0 / 30000.0) > 1e-18 || a[0] != 0.0) return fail("one-sided first level");
  if (lec_gradient_coefs(plev, 1, a, b, c) == LEC_OK) return fail("gradient of one point must fail");

  lec_grid_desc d;
  memset(&d, 0, sizeof d);
  d.nlon = NLON; d.nlat = NLAT; d.nlev = NLEV;
  d.lon_deg = lon; d.lat_deg = lat; d.rlon = rlon; d.rlat = rlat; d.coslat = coslat; d.plev = plev;
  d.dtype = LEC_F64; d.math = LEC_MATH_AUTO;
  for (int f = 0; f < 5; ++f) d.field_scale[f] = 1.0;
  d.max_steps = NT; d.max_box_rows = 0; d.device = 0;
  lec_handle *h = NULL;
  int rc = lec_create(&h, &d);
  if (!with_gpu) {
    if (rc == LEC_OK) printf("note: a CUDA device is present\n");
    else if (rc != LEC_ERR_CUDA) return fail("lec_create without a GPU must return LEC_ERR_CUDA");
    else printf("no GPU: lec_create -> %s (%s)\n", lec_strerror(rc), lec_last_error(h));
    lec_destroy(h);
    printf("OK\n");
    return 0;
  }
  if (rc != LEC_OK) return fail(lec_last_error(h));

  const size_t n = (size_t)NT * NLEV * NLAT * NLON;
  double *F[5], *R[5];
  for (int f = 0; f < 5; ++f) {
    F[f] = malloc(n * sizeof(double)); R[f] = malloc(n * sizeof(double));
    for (int t = 0; t < NT; ++t) for (int k = 0; k < NLEV; ++k) for (int j = 0; j < NLAT; ++j) for (int i = 0; i < NLON; ++i) {
      const double x = sin(0.4 * i + 0.3 * f + 0.2 * t) * cos(0.5 * j + 0.1 * k) + 0.05 * k;
      const double v = f == 0 ? 250.0 + 40.0 * plev[k] / 1e5 + 3.0 * x : f == 4 ? 9.80665 * (16000.0 * (1 - plev[k] / 1.05e5) + 30.0 * x) : 5.0 * x * (f == 3 ? 0.05 : 1.0);
      F[f][(((size_t)t * NLEV + k) * NLAT + j) * NLON + i] = v;
      R[f][(((size_t)t * NLEV + k) * NLAT + (NLAT - 1 - j)) * NLON + i] = v;      /* file stores north -> south */
    }
  }
  lec_step st[NT];
  memset(st, 0, sizeof st);
  for (int t = 0; t < NT; ++t) {
    st[t].slot = t; st[t].slot_m = t > 0 ? t - 1 : 0; st[t].slot_p = t < NT - 1 ? t + 1 : NT - 1;
    st[t].i0 = 1; st[t].i1 = NLON - 2; st[t].j0 = 1; st[t].j1 = NLAT - 2;
    const double dt = 21600.0;
    if (t == 0) { st[t].ct_0 = -1 / dt; st[t].ct_p = 1 / dt; }
    else if (t == NT - 1) { st[t].ct_m = -1 / dt; st[t].ct_0 = 1 / dt; }
    else { st[t].ct_m = -0.5 / dt; st[t].ct_p = 0.5 / dt; }
  }
  double terms[NT][LEC_NTERMS], terms_raw[NT][LEC_NTERMS];
  int32_t flags[NT];
  const void *fp[5] = {F[0], F[1], F[2], F[3], F[4]}, *rp[5] = {R[0], R[1], R[2], R[3], R[4]};
  rc = lec_run_host(h, fp, NT, st, NT, &terms[0][0], NULL, flags);
  if (rc != LEC_OK) return fail(lec_last_error(h));
  int32_t lon_map[NLON], lat_map[NLAT], lev_map[NLEV], rec[NT] = {0, 1, 2};
  for (int i = 0; i < NLON; ++i) lon_map[i] = i;
  for (int j = 0; j < NLAT; ++j) lat_map[j] = NLAT - 1 - j;
  for (int k = 0; k < NLEV; ++k) lev_map[k] = k;
  lec_raw_desc rd;
  memset(&rd, 0, sizeof rd);
  rd.dtype = LEC_RAW_F64; rd.nlon = NLON; rd.nlat = NLAT; rd.nlev = NLEV;
  rd.lon_map = lon_map; rd.lat_map = lat_map; rd.lev_map = lev_map;
  rc = lec_run_host_raw(h, &rd, rp, NT, rec, NT, st, NT, &terms_raw[0][0], NULL, NULL);
  if (rc != LEC_OK) return fail(lec_last_error(h));
  if (memcmp(terms, terms_raw, sizeof terms) != 0) return fail("raw-record path differs from the prepared path");
  for (int t = 0; t < NT; ++t) if (flags[t] != 0 || !(terms[t][LEC_AZ] > 0) || !(terms[t][LEC_KE] > 0)) return fail("terms");
  int64_t bytes[2];
  lec_last_transfer(h, bytes);
  printf("Az %.6e Ke %.6e, %lld launches, last call moved %lld B to the device\n", terms[1][LEC_AZ], terms[1][LEC_KE],
         (long long)lec_launch_count(h), (long long)bytes[0]);
  lec_destroy(h);
  printf("OK\n");
  return 0;
}
