// Multi-device plumbing of a handle created with cfg.num_devices > 1: one worker thread per GPU (engine.h Worker) and the
// library's own NCCL communicator for the weight broadcast of model_load.  libnccl.so.2 is loaded at run time (dlopen), so
// the library still loads -- and single-device handles still work -- on a machine without NCCL.
#pragma once
#include <cuda_runtime.h>

#include <condition_variable>
#include <mutex>
#include <string>

namespace infur {

struct Latch {
  std::mutex m;
  std::condition_variable cv;
  int n;
  explicit Latch(int count) : n(count) {}
  void count_down() { std::lock_guard<std::mutex> lk(m); if (--n == 0) cv.notify_all(); }
  void wait() { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [&] { return n <= 0; }); }
};

// ncclCommInitAll over `devices`; comms[i] receives the communicator of devices[i] (opaque).
bool nccl_comm_init_all(void** comms, int n, const int* devices, std::string* err);
void nccl_comm_destroy_all(void** comms, int n);
// In-place ncclBroadcast of `bytes` from rank 0's buffer to every rank's buffer, each on its own stream; returns after every
// stream has completed the transfer.
bool nccl_broadcast_all(void** comms, int n, const int* devices, void* const* bufs, size_t bytes, const cudaStream_t* streams, std::string* err);

}  // namespace infur
