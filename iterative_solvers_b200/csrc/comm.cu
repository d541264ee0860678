#include "comm.h"

#include <dlfcn.h>
#include <nccl.h>  // declarations only: every NCCL symbol is resolved with dlsym

#include <cstddef>
#include <cstring>

namespace b200cg {

namespace {
struct Api {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
Api g_api;

bool load_api(std::string* err) {
  if (g_api.lib) return true;
  // "libnccl.so.2" first: inside a torch process this resolves to the copy torch already loaded
  const char* names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
  void* lib = nullptr;
  for (const char* n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) {
    *err = std::string("cannot load NCCL (needed for world > 1): ") + dlerror();
    return false;
  }
#define SYM(field, name)                                                   \
  g_api.field = reinterpret_cast<decltype(g_api.field)>(dlsym(lib, name)); \
  if (!g_api.field) {                                                      \
    *err = std::string("NCCL symbol missing: ") + name;                    \
    return false;                                                          \
  }
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(AllGather, "ncclAllGather");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_api.lib = lib;
  return true;
}

bool check(ncclResult_t r, const char* what, std::string* err) {
  if (r == ncclSuccess) return true;
  *err = std::string(what) + " failed: " + (g_api.GetErrorString ? g_api.GetErrorString(r) : "?");
  return false;
}
}  // namespace

bool comm_unique_id(void* id128, std::string* err) {
  if (!load_api(err)) return false;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (!check(g_api.GetUniqueId(&id), "ncclGetUniqueId", err)) return false;
  memcpy(id128, &id, sizeof(id));
  return true;
}

bool comm_init(Comm* c, const void* id128, int rank, int world, cudaStream_t s, std::string* err) {
  (void)s;
  if (!load_api(err)) return false;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  if (!check(g_api.CommInitRank(&comm, world, id, rank), "ncclCommInitRank", err)) return false;
  c->comm = comm;
  c->rank = rank;
  c->world = world;
  return true;
}

void comm_destroy(Comm* c) {
  if (c->comm && g_api.CommDestroy) g_api.CommDestroy(static_cast<ncclComm_t>(c->comm));
  c->comm = nullptr;
}

bool comm_halo(Comm* c, const double* first_owned, const double* last_owned, double* halo_below, double* halo_above,
               int count, cudaStream_t s, std::string* err) {
  ncclComm_t comm = static_cast<ncclComm_t>(c->comm);
  if (!check(g_api.GroupStart(), "ncclGroupStart", err)) return false;
  bool ok = true;
  if (c->rank > 0) {  // neighbour below owns the rows under ours
    ok = ok && check(g_api.Send(first_owned, count, ncclDouble, c->rank - 1, comm, s), "ncclSend", err);
    ok = ok && check(g_api.Recv(halo_below, count, ncclDouble, c->rank - 1, comm, s), "ncclRecv", err);
  }
  if (c->rank < c->world - 1) {
    ok = ok && check(g_api.Send(last_owned, count, ncclDouble, c->rank + 1, comm, s), "ncclSend", err);
    ok = ok && check(g_api.Recv(halo_above, count, ncclDouble, c->rank + 1, comm, s), "ncclRecv", err);
  }
  ncclResult_t r = g_api.GroupEnd();
  return ok && check(r, "ncclGroupEnd", err);
}

bool comm_halo2(Comm* c, double* v0, double* v1, int yrows, int pitch, cudaStream_t s, std::string* err) {
  ncclComm_t comm = static_cast<ncclComm_t>(c->comm);
  if (!check(g_api.GroupStart(), "ncclGroupStart", err)) return false;
  bool ok = true;
  double* vs[2] = {v0, v1};
  for (double* v : vs) {
    double* first_owned = v + (size_t)1 * pitch;
    double* last_owned = v + (size_t)(yrows - 2) * pitch;
    double* halo_below = v;
    double* halo_above = v + (size_t)(yrows - 1) * pitch;
    if (c->rank > 0) {
      ok = ok && check(g_api.Send(first_owned, pitch, ncclDouble, c->rank - 1, comm, s), "ncclSend", err);
      ok = ok && check(g_api.Recv(halo_below, pitch, ncclDouble, c->rank - 1, comm, s), "ncclRecv", err);
    }
    if (c->rank < c->world - 1) {
      ok = ok && check(g_api.Send(last_owned, pitch, ncclDouble, c->rank + 1, comm, s), "ncclSend", err);
      ok = ok && check(g_api.Recv(halo_above, pitch, ncclDouble, c->rank + 1, comm, s), "ncclRecv", err);
    }
  }
  ncclResult_t r = g_api.GroupEnd();
  return ok && check(r, "ncclGroupEnd", err);
}

bool comm_allreduce_state(Comm* c, DevState* st, bool with_max, cudaStream_t s, std::string* err) {
  ncclComm_t comm = static_cast<ncclComm_t>(c->comm);
  double* sums = reinterpret_cast<double*>(reinterpret_cast<char*>(st) + offsetof(DevState, loc_s));
  double* maxs = reinterpret_cast<double*>(reinterpret_cast<char*>(st) + offsetof(DevState, loc_m));
  if (!check(g_api.GroupStart(), "ncclGroupStart", err)) return false;
  bool ok = check(g_api.AllReduce(sums, sums, 4, ncclDouble, ncclSum, comm, s), "ncclAllReduce(sum)", err);
  if (ok && with_max) ok = check(g_api.AllReduce(maxs, maxs, 4, ncclDouble, ncclMax, comm, s), "ncclAllReduce(max)", err);
  ncclResult_t r = g_api.GroupEnd();
  return ok && check(r, "ncclGroupEnd", err);
}

bool comm_allgather_bytes(Comm* c, const void* send_dev, void* recv_dev, size_t bytes, cudaStream_t s, std::string* err) {
  ncclComm_t comm = static_cast<ncclComm_t>(c->comm);
  return check(g_api.AllGather(send_dev, recv_dev, bytes, ncclUint8, comm, s), "ncclAllGather", err);
}

bool comm_all_agree(Comm* c, bool mine, bool* all, cudaStream_t s, std::string* err) {
  ncclComm_t comm = static_cast<ncclComm_t>(c->comm);
  int* d = nullptr;
  if (cudaMalloc(&d, sizeof(int)) != cudaSuccess) {
    *err = "cudaMalloc failed in comm_all_agree";
    return false;
  }
  int h = mine ? 1 : 0;
  cudaMemcpyAsync(d, &h, sizeof(int), cudaMemcpyHostToDevice, s);
  bool ok = check(g_api.AllReduce(d, d, 1, ncclInt32, ncclMin, comm, s), "ncclAllReduce(min)", err);
  cudaMemcpyAsync(&h, d, sizeof(int), cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
  cudaFree(d);
  *all = h != 0;
  return ok;
}

}  // namespace b200cg
