// comm.cuh — multi-GPU plumbing: one process per GPU, NCCL over NVLink, loaded lazily with
// dlopen("libnccl.so.2") so the single-GPU library has no NCCL link dependency and, inside a
// torch.distributed process, shares the NCCL copy that process already loaded.
// The reference has no communication backend at all; the sharding rule is §8(e).
#pragma once
#include <dlfcn.h>
#include <nccl.h>   // types only; every symbol is resolved through dlsym

#ifndef LVBA_RUNTIME_PRELUDE
#error "include runtime.cuh (it pulls comm.cuh in after its error-reporting prelude)"
#endif

namespace lvba {

struct Comm {
  int n_ranks = 1, rank = 0;
  void* lib = nullptr;
  ncclComm_t comm = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  int64_t bytes_sent = 0;         // payload this rank handed to NCCL since the last reset (the bench reports it per LM pass)
  const char* (*GetErrorString)(ncclResult_t) = nullptr;

  bool active() const { return comm != nullptr && n_ranks > 1; }

  int load() {
    if (lib) return LVBA_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return fail(LVBA_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
#define LVBA_SYM(field, name)                                                            \
  *(void**)(&field) = dlsym(lib, name);                                                  \
  if (!field) return fail(LVBA_ERR_COMM, "libnccl is missing symbol %s", name);
    LVBA_SYM(GetUniqueId, "ncclGetUniqueId")
    LVBA_SYM(CommInitRank, "ncclCommInitRank")
    LVBA_SYM(AllReduce, "ncclAllReduce")
    LVBA_SYM(AllGather, "ncclAllGather")
    LVBA_SYM(Send, "ncclSend")
    LVBA_SYM(Recv, "ncclRecv")
    LVBA_SYM(GroupStart, "ncclGroupStart")
    LVBA_SYM(GroupEnd, "ncclGroupEnd")
    LVBA_SYM(CommDestroy, "ncclCommDestroy")
    LVBA_SYM(GetErrorString, "ncclGetErrorString")
#undef LVBA_SYM
    return LVBA_OK;
  }
  int allreduce_sum(double* buf, size_t count, cudaStream_t s) {
    if (!active() || count == 0) return LVBA_OK;
    ncclResult_t r = AllReduce(buf, buf, count, ncclDouble, ncclSum, comm, s);
    if (r != ncclSuccess) return fail(LVBA_ERR_COMM, "ncclAllReduce: %s", GetErrorString(r));
    bytes_sent += (int64_t)count * 8;
    return LVBA_OK;
  }
  int allreduce_max_int(int* buf, size_t count, cudaStream_t s) {
    if (!active() || count == 0) return LVBA_OK;
    ncclResult_t r = AllReduce(buf, buf, count, ncclInt32, ncclMax, comm, s);
    if (r != ncclSuccess) return fail(LVBA_ERR_COMM, "ncclAllReduce(int, max): %s", GetErrorString(r));
    bytes_sent += (int64_t)count * 4;
    return LVBA_OK;
  }
  // in place: rank r's contribution sits at buf + r * count
  int allgather_inplace(double* buf, size_t count, cudaStream_t s) {
    if (!active() || count == 0) return LVBA_OK;
    ncclResult_t r = AllGather(buf + (size_t)rank * count, buf, count, ncclDouble, comm, s);
    if (r != ncclSuccess) return fail(LVBA_ERR_COMM, "ncclAllGather: %s", GetErrorString(r));
    bytes_sent += (int64_t)count * 8;
    return LVBA_OK;
  }
  // neighbour exchange along the chain of ranks: send `n_send` doubles to rank+1 (if any), receive `n_recv` from rank-1 (if any)
  int shift_right(const double* send, size_t n_send, double* recv, size_t n_recv, cudaStream_t s) {
    if (!active()) return LVBA_OK;
    ncclResult_t r = GroupStart();
    if (r == ncclSuccess && rank + 1 < n_ranks && n_send > 0) { r = Send(send, n_send, ncclDouble, rank + 1, comm, s); bytes_sent += (int64_t)n_send * 8; }
    if (r == ncclSuccess && rank > 0 && n_recv > 0) r = Recv(recv, n_recv, ncclDouble, rank - 1, comm, s);
    const ncclResult_t r2 = GroupEnd();
    if (r != ncclSuccess || r2 != ncclSuccess) return fail(LVBA_ERR_COMM, "ncclSend/Recv: %s", GetErrorString(r != ncclSuccess ? r : r2));
    return LVBA_OK;
  }
};

inline Comm& comm() {
  static Comm c;
  return c;
}

// contiguous pose-block-row ownership: rows [p*n/P, (p+1)*n/P) belong to rank p
inline int shard_owner(int min_pose, int n_rows, int n_ranks) {
  if (n_ranks <= 1 || n_rows <= 0) return 0;
  long long o = ((long long)min_pose * n_ranks) / n_rows;
  if (o < 0) o = 0;
  if (o >= n_ranks) o = n_ranks - 1;
  // make the rule exact with respect to the floor boundaries
  while (o + 1 < n_ranks && (long long)min_pose >= ((long long)(o + 1) * n_rows) / n_ranks) ++o;
  while (o > 0 && (long long)min_pose < ((long long)o * n_rows) / n_ranks) --o;
  return (int)o;
}

}  // namespace lvba
