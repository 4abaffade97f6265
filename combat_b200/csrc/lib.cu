#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

long long g_combat_launches = 0;
char g_combat_err[256] = "";

void combat_set_err(const char* what, cudaError_t e) {
  snprintf(g_combat_err, sizeof(g_combat_err), "%s: %s", what, cudaGetErrorString(e));
}

extern "C" int combat_version(void) { return 100; }
extern "C" long long combat_launch_count(void) { return g_combat_launches; }
extern "C" const char* combat_last_error(void) { return g_combat_err; }

// Programmatic dependent launch: OPT-IN (COMBAT_PDL=1, read once).  Measured on B200 inside the CUDA-graph-replayed step (batch
// 512, 349 launches): 12.24 ms per step with it, 11.64 ms without (profiles/r02_experiments.md) -- the early-resident CTAs of
// the next kernel cost more than the launch gap they hide; all GPU tests pass either way.
bool combat_pdl_enabled() {
  static int on = -1;
  if (on < 0) on = getenv("COMBAT_PDL") ? 1 : 0;
  return on != 0;
}
