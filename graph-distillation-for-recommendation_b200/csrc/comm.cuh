// comm.cuh — internal view of the communicator (comm.cu) for the distributed drivers (lloyd.cu).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>   // types and prototypes only: the functions are bound at run time (dlopen), never linked
#include <stdint.h>
#include "common.cuh"

struct gdr_comm {
  ncclComm_t nccl;
  int rank, world, device;
};

// Symmetric buffer: the same allocation on every rank of the communicator, each rank holding peer-mapped (CUDA IPC over
// NVLink) pointers to all copies, plus one epoch flag per source rank for the stream-ordered barrier.
struct gdr_symm {
  gdr_comm* comm;
  char* local;
  char* peer[GDR_MAX_RANKS];       // peer[rank] == local
  int64_t bytes;                   // usable bytes (the flag page lies behind them)
  uint32_t* flag_local;            // [GDR_MAX_RANKS] epochs written by the peers
  uint32_t* flag_peer[GDR_MAX_RANKS];
  uint32_t epoch;
};

namespace gdr {

struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId;
  decltype(&ncclCommInitRank) CommInitRank;
  decltype(&ncclCommDestroy) CommDestroy;
  decltype(&ncclAllReduce) AllReduce;
  decltype(&ncclAllGather) AllGather;
  decltype(&ncclSend) Send;
  decltype(&ncclRecv) Recv;
  decltype(&ncclGroupStart) GroupStart;
  decltype(&ncclGroupEnd) GroupEnd;
  decltype(&ncclGetErrorString) GetErrorString;
  decltype(&ncclGetVersion) GetVersion;
};

const NcclApi* nccl_api();   // nullptr (+ error string) when libnccl cannot be loaded

// [sums f32 | ints i32] summed over the ranks as ONE grouped NCCL operation (no-op for world 1 / null comm)
int comm_allreduce_lloyd(gdr_comm* c, float* sums, int64_t n_floats, int32_t* ints, int64_t n_ints, cudaStream_t s);
int comm_allreduce_f64(gdr_comm* c, double* buf, int64_t n, int op_max, cudaStream_t s);
int comm_allgather(gdr_comm* c, const void* send, void* recv, int64_t bytes_per_rank, cudaStream_t s);

}  // namespace gdr
