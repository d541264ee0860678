// Row-slab sharding plumbing: one process per GPU, NCCL resolved at run time (dlopen), so that the single-GPU
// path has no NCCL dependency and the library loads on machines without it.
// Per iteration: all-reduce of the dot-phase sums, all-reduce of the update-phase sums and maxima,
// one-row halo exchange of r and p with the slab neighbours - all captured in the iteration graph.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "common.cuh"

namespace b200cg {

struct Comm {
  void* lib = nullptr;
  void* comm = nullptr;  // ncclComm_t
  int rank = 0, world = 1;
};

bool comm_unique_id(void* id128, std::string* err);
bool comm_init(Comm* c, const void* id128, int rank, int world, cudaStream_t s, std::string* err);
void comm_destroy(Comm* c);
// exchange one row (count doubles) with both neighbours: send first/last owned rows, receive into the halo rows
bool comm_halo(Comm* c, const double* first_owned, const double* last_owned, double* halo_below, double* halo_above,
               int count, cudaStream_t s, std::string* err);
// the same for two vectors (r and p) in one NCCL group = one launch
bool comm_halo2(Comm* c, double* v0, double* v1, int yrows, int pitch, cudaStream_t s, std::string* err);
// in-place all-reduce of st->loc_s (sum, 4 doubles) and optionally st->loc_m (max, 4 doubles)
bool comm_allreduce_state(Comm* c, DevState* st, bool with_max, cudaStream_t s, std::string* err);

// all-gather of `bytes` per rank (device buffers); used once, to swap CUDA IPC handles
bool comm_allgather_bytes(Comm* c, const void* send_dev, void* recv_dev, size_t bytes, cudaStream_t s, std::string* err);
// all ranks learn whether every rank said yes (min-reduction of a flag; host values, blocking)
bool comm_all_agree(Comm* c, bool mine, bool* all, cudaStream_t s, std::string* err);

}  // namespace b200cg
