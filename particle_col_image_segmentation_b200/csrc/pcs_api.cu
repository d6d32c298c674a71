// Library-wide entry points: version and error reporting.
#include <stdio.h>
#include <string.h>

#include "pcs_common.cuh"

#include "pcs.h"

static thread_local char g_err[512] = "";

extern "C" void pcs_set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

int pcs_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return PCS_OK;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return PCS_ERR_CUDA;
}

// ---------------------------------------------------------------- launch accounting
// Every kernel launch in the library goes through PCS_LAUNCH: it is counted, and while
// profiling is enabled it is bracketed by CUDA events on the launching stream.
#define PCS_PROF_MAX 8192
struct ProfRec {
  const char* name;
  cudaEvent_t a, b;
};
static ProfRec g_recs[PCS_PROF_MAX];
static int g_nrec = 0, g_nevents = 0;
static bool g_prof = false, g_open = false;
static unsigned long long g_launches = 0;

void pcs_prof_begin(const char* name, cudaStream_t st) {
  ++g_launches;
  g_open = false;
  if (!g_prof || g_nrec >= PCS_PROF_MAX) return;
  ProfRec& r = g_recs[g_nrec];
  if (g_nrec >= g_nevents) {
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    ++g_nevents;
  }
  r.name = name;
  cudaEventRecord(r.a, st);
  g_open = true;
}

void pcs_prof_end(cudaStream_t st) {
  if (!g_open) return;
  cudaEventRecord(g_recs[g_nrec].b, st);
  ++g_nrec;
  g_open = false;
}

extern "C" {

int pcs_version(void) { return PCS_VERSION; }

uint64_t pcs_kernel_launches(void) { return g_launches; }

int pcs_profile_enable(int on) {
  g_prof = on != 0;
  if (on) g_nrec = 0;
  return PCS_OK;
}

// Aggregates the recorded launches by kernel name.  names: n_max slots of 64 chars.
// Returns the number of distinct kernels (synchronises on the recorded events).
int pcs_profile_collect(char* names, double* total_ms, int32_t* launches, int n_max) {
  int n = 0;
  for (int i = 0; i < g_nrec; ++i) {
    float ms = 0.f;
    if (cudaEventSynchronize(g_recs[i].b) != cudaSuccess) continue;
    if (cudaEventElapsedTime(&ms, g_recs[i].a, g_recs[i].b) != cudaSuccess) continue;
    int j = 0;
    for (; j < n; ++j)
      if (strncmp(names + 64 * j, g_recs[i].name, 63) == 0) break;
    if (j == n) {
      if (n >= n_max) continue;
      strncpy(names + 64 * n, g_recs[i].name, 63);
      names[64 * n + 63] = 0;
      total_ms[n] = 0.0;
      launches[n] = 0;
      ++n;
    }
    total_ms[j] += ms;
    launches[j] += 1;
  }
  g_nrec = 0;
  return n;
}

const char* pcs_last_error_string(void) { return g_err; }

int pcs_device_sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
  return n;
}

}  // extern "C"
