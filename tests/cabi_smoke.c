/* Plain-C consumer of include/sats.h: proves the header is C (not C++), that libsats links from C, and that the host
 * half works without a GPU while the search half refuses to run (no CPU fallback).  Built and run by
 * tests/test_cabi_host.py::test_c_consumer_links_and_runs. */
#include <stdio.h>
#include <string.h>
#include "sats.h"

int main(int argc, char **argv)
{
  sats_db *db = NULL, *queries = NULL;
  sats_searcher *s = NULL;
  sats_params prm;
  char dbfile[256];
  int flags[3];
  const char *input =
      "some.db\nT T F\n"
      "QRY1        3\ne  \nOT e  \nLE RT xa \n 0.000 \n 4.501  0.000 \n11.662 10.386  1.000 \n";
  if (argc < 2) return 2;
  if (sats_db_read_packed(argv[1], &db) != SATS_OK) { fprintf(stderr, "%s\n", sats_last_error()); return 3; }
  if (sats_input_parse(input, strlen(input), dbfile, sizeof dbfile, flags, &queries) != SATS_OK) return 4;
  if (strcmp(dbfile, "some.db") || !flags[0] || !flags[1] || flags[2] || sats_db_count(queries) != 1) return 5;
  if (sats_db_order(queries, 0) != 3 || strcmp(sats_db_name(queries, 0), "QRY1")) return 6;
  sats_params_default(&prm);
  if (prm.restarts != SATS_DEFAULT_MAXSTART || prm.seed != SATS_REF_SEED) return 7;
  printf("%d %d %.6f %.6f\n", sats_db_count(db), sats_db_max_order(db), sats_norm2(54, 8, 8),
         sats_z_gumbel((int)sats_norm2(54, 8, 8), sats_gumbel_a, sats_gumbel_b));
  {
    /* round-2 surface from plain C: query construction and the integer cut-off tables */
    char code[3];
    uint32_t cut[19];
    double d[3], c[3], omega = 0.0;
    const double helix[12] = {2.3, 0.0, 0.0, -0.4, 2.27, 1.5, -2.16, -0.79, 3.0, 1.15, -1.99, 4.5};
    const double c1[3] = {0, 0, 0}, d1[3] = {1, 0, 0}, c2[3] = {0, 0, 7}, d2[3] = {0, 1, 0};
    sats_db *built = NULL;
    const uint8_t types[2] = {1, 0};
    const int32_t nres[2] = {4, 2};
    const double traces[18] = {2.3, 0.0, 0.0, -0.4, 2.27, 1.5, -2.16, -0.79, 3.0, 1.15, -1.99, 4.5, 10, 0, 0, 10, 0, 3.3};
    if (sats_tabcode_from_angle(0.1, code) != SATS_OK || strcmp(code, "PD")) return 11;
    if (sats_tabcode_from_angle(4.0, code) != SATS_ERR_ARG) return 12;
    if (sats_relative_angle(c1, d1, c2, d2, &omega) != 0 || omega == 0.0) return 13;
    if (sats_fit_axis(1, 4, helix, d, c) != 0 || sats_fit_axis(1, 2, helix, d, c) != 1) return 14;
    if (sats_build_structure_from_ca("built", 2, types, nres, traces, &built) != SATS_OK || sats_db_order(built, 0) != 2) return 15;
    sats_db_free(built);
    if (sats_pick_boundaries(19, cut) != SATS_OK || cut[0] != 0 || cut[1] <= (1u << 27) || sats_seed_cutoff() != 0x7fffffc0u) return 16;
  }
  if (sats_device_count() == 0) {
    int rc = sats_searcher_create(db, 0, 0, 1, &s);
    if (rc != SATS_ERR_CUDA || s != NULL) return 8;      /* must fail loudly, never fall back */
    printf("no-gpu: %s\n", sats_last_error());
  } else {
    int32_t scores[4096];
    if (sats_searcher_create(db, 0, 0, 1, &s) != SATS_OK) return 9;
    prm.restarts = 32;
    if (sats_db_count(db) > 4096 || sats_search(s, queries, 0, 1, &prm, 0, scores, NULL) != SATS_OK) return 10;
    printf("gpu: first score %d\n", scores[0]);
    sats_searcher_free(s);
  }
  sats_db_free(db);
  sats_db_free(queries);
  return 0;
}
