// lvba_b200.cu — the single translation unit of liblvba_b200.so (sm_100a).
// Kernels live in the .cuh files next to this one; this file adds the path-independent entry points.
#include <memory>

#include "lidar_api.cuh"
#include "visual_api.cuh"
#include "voxel_api.cuh"
#include "depth_api.cuh"
#include "track_api.cuh"
#include "anchor_api.cuh"

extern "C" {

int lvba_version(void) { return LVBA_B200_VERSION; }
int lvba_device_count(void) { return lvba::device_count(); }
const char* lvba_last_error(void) { return lvba::last_error_ref().c_str(); }
int lvba_release_cached_memory(void) { lvba::device_pool().clear(); return LVBA_OK; }
const char* lvba_status_string(int status) {
  switch (status) {
    case LVBA_OK: return "ok";
    case LVBA_ERR_INVALID_ARG: return "invalid argument";
    case LVBA_ERR_NO_DEVICE: return "no CUDA device (no CPU fallback)";
    case LVBA_ERR_CUDA: return "CUDA runtime error";
    case LVBA_ERR_UNSUPPORTED: return "unsupported problem shape";
    case LVBA_ERR_NUMERIC: return "numerical failure";
    case LVBA_ERR_COMM: return "communication (NCCL) error";
    case LVBA_ERR_NOMEM: return "out of memory";
    default: return "unknown status";
  }
}

// ---------------------------------------------------------------- multi-GPU
int lvba_comm_unique_id(void* id_out) {
  if (!id_out) return lvba::fail(LVBA_ERR_INVALID_ARG, "null argument");
  lvba::Comm& c = lvba::comm();
  LVBA_TRY(c.load());
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == LVBA_NCCL_ID_BYTES, "ncclUniqueId size");
  ncclResult_t r = c.GetUniqueId(&id);
  if (r != ncclSuccess) return lvba::fail(LVBA_ERR_COMM, "ncclGetUniqueId: %s", c.GetErrorString(r));
  memcpy(id_out, &id, sizeof id);
  return LVBA_OK;
}

int lvba_comm_init(int32_t n_ranks, int32_t rank, const void* id, int32_t device) {
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return lvba::fail(LVBA_ERR_INVALID_ARG, "bad rank %d / %d", rank, n_ranks);
  lvba::Comm& c = lvba::comm();
  if (c.comm) return lvba::fail(LVBA_ERR_INVALID_ARG, "communicator already initialised");
  c.n_ranks = n_ranks; c.rank = rank;
  if (n_ranks == 1) return LVBA_OK;
  if (!id) return lvba::fail(LVBA_ERR_INVALID_ARG, "null unique id");
  LVBA_TRY(lvba::select_device(device));
  LVBA_TRY(c.load());
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof uid);
  ncclResult_t r = c.CommInitRank(&c.comm, n_ranks, uid, rank);
  if (r != ncclSuccess) { c.comm = nullptr; c.n_ranks = 1; c.rank = 0; return lvba::fail(LVBA_ERR_COMM, "ncclCommInitRank: %s", c.GetErrorString(r)); }
  return LVBA_OK;
}

int lvba_comm_destroy(void) {
  lvba::Comm& c = lvba::comm();
  if (c.comm) { c.CommDestroy(c.comm); c.comm = nullptr; }
  c.n_ranks = 1; c.rank = 0;
  return LVBA_OK;
}

int lvba_comm_info(int32_t* n_ranks, int32_t* rank) {
  lvba::Comm& c = lvba::comm();
  if (n_ranks) *n_ranks = c.n_ranks;
  if (rank) *rank = c.rank;
  return LVBA_OK;
}

int32_t lvba_shard_owner(int32_t min_pose, int32_t n_rows, int32_t n_ranks) { return lvba::shard_owner(min_pose, n_rows, n_ranks); }

}  // extern "C"
