// Plugin registry, library entry points and process-wide services (logger, device counters, NCCL
// communicators).  Replaces P/api/InferPlugin.cpp:55-171 (initLibNvInferPlugins, creator registration
// under a mutex) and P/common/plugin.cpp:13-142 (NCCL dtype map, communicator map).
// TensorRT's own getPluginRegistry() lives in libnvinfer; without TensorRT this library provides it.
#include <dlfcn.h>

#include <map>
#include <memory>
#include <set>

#include "pluginBase.h"

using namespace nvinfer1;

namespace tb {
namespace plugins {

std::vector<IPluginCreator*>& all_creators();

// ---- logging -----------------------------------------------------------------------------------------
static ILogger* g_logger = nullptr;
void set_logger(ILogger* l) { g_logger = l; }
void log_msg(ILogger::Severity sev, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (g_logger) {
    g_logger->log(sev, buf);
  } else if (sev <= ILogger::Severity::kWARNING) {
    fprintf(stderr, "[trtllm_b200][%s] %s\n", sev <= ILogger::Severity::kERROR ? "E" : "W", buf);
  }
}

void Fields::report_unused(const char* plugin) const {
  for (size_t i = 0; i < used_.size(); ++i)
    if (!used_[i]) log_msg(ILogger::Severity::kERROR, "%s: unused plugin field '%s'", plugin, fc_->fields[i].name);
}

// ---- device counters ------------------------------------------------------------------------------------
// Growing never frees the smaller buffer: a CUDA graph captured earlier may still hold its address (its kernels
// keep using the old, still valid, self-resetting counters); everything is freed when the plugin is destroyed.
int* DeviceCounters::get(size_t bytes) {
  if (ptr_ && bytes_ >= bytes) return static_cast<int*>(ptr_);
  if (ptr_) retired_.push_back(ptr_);
  ptr_ = nullptr;
  if (cudaMalloc(&ptr_, bytes) != cudaSuccess) throw PluginError("cudaMalloc of plugin counters failed");
  if (cudaMemset(ptr_, 0, bytes) != cudaSuccess) throw PluginError("cudaMemset of plugin counters failed");
  bytes_ = bytes;
  return static_cast<int*>(ptr_);
}
void DeviceCounters::release() {
  if (ptr_) cudaFree(ptr_);
  for (void* p : retired_) cudaFree(p);
  retired_.clear();
  ptr_ = nullptr;
  bytes_ = 0;
}

// ---- registry ---------------------------------------------------------------------------------------------
class Registry : public IPluginRegistry {
 public:
  bool registerCreator(IPluginCreator& c, const char* const ns) noexcept override {
    std::lock_guard<std::mutex> g(mu_);
    const std::string key = make_key(ns, c.getPluginName(), c.getPluginVersion());
    if (by_key_.count(key)) return false;
    c.setPluginNamespace(ns);
    by_key_[key] = &c;
    list_.push_back(&c);
    return true;
  }
  IPluginCreator* const* getPluginCreatorList(int32_t* const n) const noexcept override {
    std::lock_guard<std::mutex> g(mu_);
    if (n) *n = (int32_t) list_.size();
    return list_.data();
  }
  IPluginCreator* getPluginCreator(const char* const name, const char* const version, const char* const ns) noexcept override {
    std::lock_guard<std::mutex> g(mu_);
    auto it = by_key_.find(make_key(ns, name, version));
    return it == by_key_.end() ? nullptr : it->second;
  }
  bool deregisterCreator(const IPluginCreator& c) noexcept override {
    std::lock_guard<std::mutex> g(mu_);
    for (auto it = by_key_.begin(); it != by_key_.end(); ++it)
      if (it->second == &c) {
        for (size_t i = 0; i < list_.size(); ++i)
          if (list_[i] == &c) { list_.erase(list_.begin() + i); break; }
        by_key_.erase(it);
        return true;
      }
    return false;
  }

 private:
  static std::string make_key(const char* ns, const char* name, const char* version) {
    return std::string(ns ? ns : "") + "::" + (name ? name : "") + " version " + (version ? version : "");
  }
  mutable std::mutex mu_;
  std::map<std::string, IPluginCreator*> by_key_;
  std::vector<IPluginCreator*> list_;
};

// ---- NCCL through dlopen: no link-time dependency; the AllReduce / AllGather plugins fail loudly without it ----
namespace {
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId { char internal[128]; };
struct NcclApi {
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // RTLD_NOLOAD first: share the instance torch already mapped (same soname) instead of loading a second NCCL
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.GetUniqueId = (decltype(api.GetUniqueId)) dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank)) dlsym(h, "ncclCommInitRank");
    api.AllReduce = (decltype(api.AllReduce)) dlsym(h, "ncclAllReduce");
    api.AllGather = (decltype(api.AllGather)) dlsym(h, "ncclAllGather");
    api.CommDestroy = (decltype(api.CommDestroy)) dlsym(h, "ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString)) dlsym(h, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.AllGather;
  });
  return api;
}
constexpr int kNcclFloat16 = 6, kNcclSum = 0;   // ncclDataType_t::ncclFloat16 / ncclRedOp_t::ncclSum (nccl.h)
std::mutex g_comm_mu;
}  // namespace

struct CommHandle {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
};
static std::map<std::set<int32_t>, std::unique_ptr<CommHandle>>& comm_map() {
  static std::map<std::set<int32_t>, std::unique_ptr<CommHandle>> m;
  return m;
}
CommHandle* find_comm(const std::vector<int32_t>& group) {
  std::lock_guard<std::mutex> g(g_comm_mu);
  auto it = comm_map().find(std::set<int32_t>(group.begin(), group.end()));
  return it == comm_map().end() ? nullptr : it->second.get();
}
int comm_size(const CommHandle* c) { return c->nranks; }
int comm_allreduce_half(CommHandle* c, const void* in, void* out, size_t count, cudaStream_t stream) {
  return nccl().AllReduce(in, out, count, kNcclFloat16, kNcclSum, c->comm, stream);
}
int comm_allgather_half(CommHandle* c, const void* in, void* out, size_t count_per_rank, cudaStream_t stream) {
  return nccl().AllGather(in, out, count_per_rank, kNcclFloat16, c->comm, stream);
}

static Registry& registry() {
  static Registry r;
  return r;
}

}  // namespace plugins
}  // namespace tb

using namespace tb::plugins;

extern "C" {

#ifndef TB_HAVE_TENSORRT   // with TensorRT present its own registry (libnvinfer) is the one plugins register into
IPluginRegistry* getPluginRegistry() noexcept { return &registry(); }
int32_t getInferLibVersion() noexcept { return NV_TENSORRT_VERSION; }
#endif

// P/api/InferPlugin.cpp:149-171: register every creator once under `libNamespace`.
bool initLibNvInferPlugins(void* logger, const char* libNamespace) {
  if (logger) set_logger(static_cast<ILogger*>(logger));
  const char* ns = libNamespace ? libNamespace : kNamespace;
  IPluginRegistry* reg = getPluginRegistry();
  for (IPluginCreator* c : all_creators())
    if (!reg->getPluginCreator(c->getPluginName(), c->getPluginVersion(), ns)) reg->registerCreator(*c, ns);
  return true;
}

// ---- communicator bootstrap (replaces the MPI exchange of P/ncclPlugin/allreducePlugin.cpp:128-167) ------------
int tb_comm_unique_id(void* out128) {
  if (!nccl().ok) return -20;
  ncclUniqueId id;
  const int rc = nccl().GetUniqueId(&id);
  if (rc == 0) std::memcpy(out128, &id, sizeof(id));
  return rc;
}
int tb_comm_init(const void* unique_id128, const int32_t* group, int group_size, int rank_in_group) {
  if (!nccl().ok) return -20;
  ncclUniqueId id;
  std::memcpy(&id, unique_id128, sizeof(id));
  auto h = std::make_unique<CommHandle>();
  h->nranks = group_size;
  h->rank = rank_in_group;
  const int rc = nccl().CommInitRank(&h->comm, group_size, id, rank_in_group);
  if (rc != 0) return rc;
  std::lock_guard<std::mutex> g(g_comm_mu);
  comm_map()[std::set<int32_t>(group, group + group_size)] = std::move(h);
  return 0;
}
}
