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

extern "C" {

int pcs_version(void) { return PCS_VERSION; }

const char* pcs_last_error_string(void) { return g_err; }

int pcs_device_sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
  return n;
}

}  // extern "C"
