/*
 * examples/c_host.c — the C ABI of libabm.so used from plain C: the v0.1 model (RANDOM_WALK, 100 sheep on a 100 x 100
 * non-periodic grid, reference scripts/v0.1.jl) stepped for 100 ticks with the two aggregations `run!` collects
 * (adata = [(sheep, count)], mdata = [count_grass]).
 *
 *   gcc -std=c99 -Wall -Iinclude examples/c_host.c -L animal-movement-abm_b200 -labm -Wl,-rpath,$PWD/animal-movement-abm_b200 -o c_host
 *
 * Needs a B200: abm_create fails with ABM_E_CUDA elsewhere (there is no CPU path).
 */
#include <stdio.h>
#include <stdlib.h>

#include "abm.h"

#define CHECK(call) do { int rc_ = (call); if (rc_ != ABM_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, abm_last_error(h)); return 1; } } while (0)

int main(void) {
  enum { NX = 100, NY = 100, N = 100 };
  abm_config cfg;
  abm_handle* h = NULL;
  abm_config_default(&cfg);
  cfg.nx = NX; cfg.ny = NY; cfg.periodic = 0; cfg.walk_type = ABM_RANDOM_WALK; cfg.seed = 321;
  CHECK(abm_create(&cfg, &h));

  /* grass: every cell grown (countdown = regrowth_time); sheep on a diagonal with energy 5 */
  uint8_t* grown = malloc((size_t)NX * NY);
  int64_t* countdown = malloc(sizeof(int64_t) * NX * NY);
  for (int c = 0; c < NX * NY; ++c) { grown[c] = 1; countdown[c] = cfg.regrowth_time; }
  CHECK(abm_set_grid(h, grown, countdown, NULL, NULL));
  int64_t id[N], x[N], y[N];
  double energy[N];
  for (int i = 0; i < N; ++i) { id[i] = i + 1; x[i] = 1 + i % NX; y[i] = 1 + (7 * i) % NY; energy[i] = 5.0; }
  CHECK(abm_set_agents(h, N, id, x, y, energy, N + 1));

  for (int t = 0; t <= 100; t += 10) {
    double sheep = 0, grass = 0;
    CHECK(abm_reduce(h, ABM_REDUCE_COUNT_AGENTS, &sheep));
    CHECK(abm_reduce(h, ABM_REDUCE_COUNT_GROWN, &grass));
    printf("tick %3d  sheep %5.0f  grown cells %6.0f\n", t, sheep, grass);
    if (t < 100) CHECK(abm_step(h, 10));
  }
  abm_destroy(h);
  free(grown); free(countdown);
  return 0;
}
