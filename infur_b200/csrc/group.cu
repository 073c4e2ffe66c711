// Worker threads and the run-time NCCL binding of multi-device handles (group.h).
#include "group.h"

#include <dlfcn.h>
#include <nccl.h>

#include "engine.h"

namespace infur {

void Worker::start() {
  th = std::thread([this] {
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return stop || !q.empty(); });
        if (q.empty()) return;   // stop requested and the queue is drained
        f = std::move(q.front());
        q.pop_front();
      }
      f();
    }
  });
}

void Worker::post(std::function<void()> f) {
  {
    std::lock_guard<std::mutex> lk(m);
    q.push_back(std::move(f));
  }
  cv.notify_one();
}

void Worker::run_sync(const std::function<void()>& f) {
  Latch latch(1);
  post([&] { f(); latch.count_down(); });
  latch.wait();
}

void Worker::shutdown() {
  {
    std::lock_guard<std::mutex> lk(m);
    stop = true;
  }
  cv.notify_all();
  if (th.joinable()) th.join();
}

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // a process that already holds NCCL (e.g. PyTorch's bundled copy) gets that copy back: same SONAME
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) { api.why = std::string("libnccl.so.2 could not be loaded: ") + dlerror(); return; }
    auto sym = [&](const char* name) { void* p = dlsym(api.lib, name); if (!p && api.why.empty()) api.why = std::string("libnccl lacks ") + name; return p; };
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return &api;
}

}  // namespace

bool nccl_comm_init_all(void** comms, int n, const int* devices, std::string* err) {
  NcclApi* a = nccl_api();
  if (!a->why.empty()) { if (err) *err = "multi-device handles need NCCL for the weight broadcast: " + a->why; return false; }
  ncclComm_t cs[INFUR_B200_MAX_DEVICES];
  const ncclResult_t r = a->CommInitAll(cs, n, devices);
  if (r != ncclSuccess) { if (err) *err = std::string("ncclCommInitAll: ") + a->GetErrorString(r); return false; }
  for (int i = 0; i < n; ++i) comms[i] = cs[i];
  return true;
}

void nccl_comm_destroy_all(void** comms, int n) {
  NcclApi* a = nccl_api();
  if (!a->CommDestroy) return;
  for (int i = 0; i < n; ++i)
    if (comms[i]) { a->CommDestroy(reinterpret_cast<ncclComm_t>(comms[i])); comms[i] = nullptr; }
}

bool nccl_broadcast_all(void** comms, int n, const int* devices, void* const* bufs, size_t bytes, const cudaStream_t* streams, std::string* err) {
  NcclApi* a = nccl_api();
  if (!a->why.empty()) { if (err) *err = a->why; return false; }
  ncclResult_t r = a->GroupStart();
  for (int i = 0; i < n && r == ncclSuccess; ++i)
    r = a->Broadcast(bufs[i], bufs[i], bytes, ncclUint8, 0, reinterpret_cast<ncclComm_t>(comms[i]), streams[i]);
  const ncclResult_t r2 = a->GroupEnd();
  if (r == ncclSuccess) r = r2;
  if (r != ncclSuccess) { if (err) *err = std::string("ncclBroadcast: ") + a->GetErrorString(r); return false; }
  for (int i = 0; i < n; ++i) {
    cudaSetDevice(devices[i]);
    const cudaError_t e = cudaStreamSynchronize(streams[i]);
    if (e != cudaSuccess) { if (err) *err = std::string("after ncclBroadcast: ") + cudaGetErrorString(e); return false; }
  }
  return true;
}

}  // namespace infur
