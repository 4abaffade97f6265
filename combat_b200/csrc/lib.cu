#include <stdio.h>
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
