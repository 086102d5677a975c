// Peer-mapped device buffers for ranks on one NVLink / NVSwitch node: plain cudaMalloc blocks exported and imported with CUDA IPC
// handles (the host layer exchanges the 64 handle bytes through torch.distributed).  Used by the SVGD workspace so that the exact
// distributed median reads the other ranks' window tables and radix histograms directly (svgd_state.cuh) instead of launching
// collectives.  The reference has no counterpart: it is a single-process program.
#include "common.cuh"
#include <string.h>

using namespace bode;

extern "C" int bode_peer_alloc(size_t bytes, void** out) {
  BODE_REQUIRE(out && bytes > 0, "bad args");
  void* p = nullptr;
  BODE_CUDA(cudaMalloc(&p, bytes));
  BODE_CUDA(cudaMemset(p, 0, bytes));
  *out = p;
  return BODE_OK;
}

extern "C" int bode_peer_free(void* ptr) {
  if (ptr) BODE_CUDA(cudaFree(ptr));
  return BODE_OK;
}

/* handle: 64 bytes (cudaIpcMemHandle_t) */
extern "C" int bode_peer_export(void* ptr, void* handle) {
  BODE_REQUIRE(ptr && handle, "null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  BODE_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle, &h, sizeof(h));
  return BODE_OK;
}

extern "C" int bode_peer_import(const void* handle, void** out) {
  BODE_REQUIRE(handle && out, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  BODE_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *out = p;
  return BODE_OK;
}

extern "C" int bode_peer_release(void* imported) {
  if (imported) BODE_CUDA(cudaIpcCloseMemHandle(imported));
  return BODE_OK;
}
