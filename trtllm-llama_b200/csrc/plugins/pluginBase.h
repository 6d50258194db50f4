// Shared machinery of the plugin library: registry, field parsing, POD (de)serialisation, a base
// class that supplies the IPluginV2DynamicExt boilerplate, and a generic creator.
//
// Replaces P/common/plugin.{h,cpp} (PluginFieldParser P/common/plugin.cpp:187-260, read/write
// helpers P/common/plugin.h:60-100, BasePlugin / BaseCreator P/common/plugin.h:30-58) and the
// registration half of P/api/InferPlugin.cpp:55-171.  Different design: plugins declare a
// table of typed fields once; creation, getFieldNames and the unused-field check are derived from it.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/trtllm_b200.h"
#include "NvInferMin.h"

namespace tb {
namespace plugins {

constexpr const char* kNamespace = "tensorrt_llm";
constexpr const char* kVersion = "1";

void log_msg(nvinfer1::ILogger::Severity sev, const char* fmt, ...);
void set_logger(nvinfer1::ILogger* logger);

struct PluginError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
#define TBP_REQUIRE(cond, msg)                                                        \
  do {                                                                                \
    if (!(cond)) throw ::tb::plugins::PluginError(std::string(msg) + " [" #cond "]"); \
  } while (0)

// ---- in-order POD blobs (the reference's write()/read(), P/common/plugin.h:60-100) -------------
struct Writer {
  char* p;
  template <class T> void put(const T& v) { std::memcpy(p, &v, sizeof(T)); p += sizeof(T); }
};
struct Reader {
  const char* p;
  const char* end;
  template <class T> T get() {
    if (p + sizeof(T) > end) throw PluginError("serialised plugin blob too short");
    T v;
    std::memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  void finish() const { if (p != end) throw PluginError("serialised plugin blob has trailing bytes"); }
};

// ---- typed access to a PluginFieldCollection ------------------------------------------------------
class Fields {
 public:
  explicit Fields(const nvinfer1::PluginFieldCollection* fc) : fc_(fc), used_(fc ? fc->nbFields : 0, false) {}
  // scalar of type T stored as the field's declared PluginFieldType; returns false when absent
  template <class T> bool scalar(const char* name, T& out) {
    const nvinfer1::PluginField* f = find(name);
    if (!f || !f->data) return false;
    out = convert<T>(*f, 0);
    return true;
  }
  template <class T> T required(const char* name) {
    T v{};
    if (!scalar(name, v)) throw PluginError(std::string("missing plugin field: ") + name);
    return v;
  }
  template <class T> T optional(const char* name, T dflt) {
    T v = dflt;
    scalar(name, v);
    return v;
  }
  std::vector<int32_t> int_list(const char* name) {
    std::vector<int32_t> r;
    const nvinfer1::PluginField* f = find(name);
    if (f && f->data)
      for (int i = 0; i < f->length; ++i) r.push_back(convert<int32_t>(*f, i));
    return r;
  }
  // the reference logs every field the plugin did not consume (P/common/plugin.cpp:196-207)
  void report_unused(const char* plugin) const;

 private:
  const nvinfer1::PluginField* find(const char* name) {
    if (!fc_) return nullptr;
    for (int i = 0; i < fc_->nbFields; ++i)
      if (fc_->fields[i].name && std::strcmp(fc_->fields[i].name, name) == 0) {
        used_[i] = true;
        return &fc_->fields[i];
      }
    return nullptr;
  }
  template <class T> static T convert(const nvinfer1::PluginField& f, int i) {
    using FT = nvinfer1::PluginFieldType;
    switch (f.type) {
      case FT::kINT8: return (T) static_cast<const int8_t*>(f.data)[i];
      case FT::kINT16: return (T) static_cast<const int16_t*>(f.data)[i];
      case FT::kINT32: return (T) static_cast<const int32_t*>(f.data)[i];
      case FT::kFLOAT32: return (T) static_cast<const float*>(f.data)[i];
      case FT::kFLOAT64: return (T) static_cast<const double*>(f.data)[i];
      default: throw PluginError(std::string("unsupported PluginFieldType for field ") + f.name);
    }
  }
  const nvinfer1::PluginFieldCollection* fc_;
  std::vector<bool> used_;
};

// ---- dims helpers -----------------------------------------------------------------------------------
inline int64_t volume(const nvinfer1::Dims& d) {
  int64_t v = 1;
  for (int i = 0; i < d.nbDims; ++i) v *= d.d[i];
  return v;
}
inline int64_t rows_of(const nvinfer1::Dims& d) { return d.nbDims ? volume(d) / d.d[d.nbDims - 1] : 1; }
inline int last_dim(const nvinfer1::Dims& d) { return d.d[d.nbDims - 1]; }
// workspace carving at the reference's 128-byte alignment (P/common/plugin.cpp:11,144-185)
inline size_t align128(size_t n) { return (n + 127) & ~(size_t) 127; }
inline void* carve(void*& cursor, size_t bytes) {
  uintptr_t a = (reinterpret_cast<uintptr_t>(cursor) + 127) & ~(uintptr_t) 127;
  cursor = reinterpret_cast<void*>(a + bytes);
  return reinterpret_cast<void*>(a);
}

// Zeroed-once device counters owned by a plugin instance (split-K / split-L arrival counters):
// a TensorRT workspace is never initialised, so the self-resetting counters cannot live there.
class DeviceCounters {
 public:
  DeviceCounters() = default;
  DeviceCounters(const DeviceCounters&) {}                       // a copy (plugin clone) owns its own, lazily
  DeviceCounters& operator=(const DeviceCounters&) { release(); return *this; }
  ~DeviceCounters() { release(); }
  int* get(size_t bytes);
  void release();

 private:
  void* ptr_ = nullptr;
  size_t bytes_ = 0;
  std::vector<void*> retired_;   // outgrown buffers, kept alive for graphs captured with them
};

// ---- base plugin -----------------------------------------------------------------------------------------
class BasePlugin : public nvinfer1::IPluginV2DynamicExt {
 public:
  const char* getPluginVersion() const noexcept override { return kVersion; }
  int32_t initialize() noexcept override { return 0; }
  void terminate() noexcept override {}
  void destroy() noexcept override { delete this; }
  void setPluginNamespace(const char* ns) noexcept override { ns_ = ns ? ns : ""; }
  const char* getPluginNamespace() const noexcept override { return ns_.c_str(); }
  void configurePlugin(const nvinfer1::DynamicPluginTensorDesc*, int32_t, const nvinfer1::DynamicPluginTensorDesc*,
                       int32_t) noexcept override {}
  using nvinfer1::IPluginV2DynamicExt::getWorkspaceSize;
  size_t getWorkspaceSize(const nvinfer1::PluginTensorDesc*, int32_t, const nvinfer1::PluginTensorDesc*,
                          int32_t) const noexcept override { return 0; }

 protected:
  std::string ns_ = kNamespace;
};

// enqueue bodies throw PluginError / return kernel codes; this maps both onto the noexcept int contract
template <class F> int guarded(const char* what, F&& f) noexcept {
  try {
    const int rc = f();
    if (rc != 0) log_msg(nvinfer1::ILogger::Severity::kERROR, "%s: kernel launch failed with code %d", what, rc);
    return rc;
  } catch (const std::exception& e) {
    log_msg(nvinfer1::ILogger::Severity::kERROR, "%s: %s", what, e.what());
    return -1;
  }
}

// ---- generic creator ------------------------------------------------------------------------------------
// P must provide: static const char* type_name(); static const std::vector<nvinfer1::PluginField>& field_table();
//                 P(Fields&) ; P(Reader&)
template <class P>
class Creator : public nvinfer1::IPluginCreator {
 public:
  Creator() {
    fc_.nbFields = (int32_t) P::field_table().size();
    fc_.fields = P::field_table().data();
  }
  const char* getPluginName() const noexcept override { return P::type_name(); }
  const char* getPluginVersion() const noexcept override { return kVersion; }
  const nvinfer1::PluginFieldCollection* getFieldNames() noexcept override { return &fc_; }
  nvinfer1::IPluginV2* createPlugin(const char* /*name*/, const nvinfer1::PluginFieldCollection* fc) noexcept override {
    try {
      Fields f(fc);
      P* p = new P(f);
      f.report_unused(P::type_name());
      p->setPluginNamespace(ns_.c_str());
      return p;
    } catch (const std::exception& e) {
      // creators never throw: log and return nullptr (P/gptAttentionPlugin/gptAttentionPlugin.cpp:487-510)
      log_msg(nvinfer1::ILogger::Severity::kERROR, "%s::createPlugin: %s", P::type_name(), e.what());
      return nullptr;
    }
  }
  nvinfer1::IPluginV2* deserializePlugin(const char* /*name*/, const void* data, size_t len) noexcept override {
    try {
      Reader r{static_cast<const char*>(data), static_cast<const char*>(data) + len};
      P* p = new P(r);
      r.finish();
      p->setPluginNamespace(ns_.c_str());
      return p;
    } catch (const std::exception& e) {
      log_msg(nvinfer1::ILogger::Severity::kERROR, "%s::deserializePlugin: %s", P::type_name(), e.what());
      return nullptr;
    }
  }
  void setPluginNamespace(const char* ns) noexcept override { ns_ = ns ? ns : ""; }
  const char* getPluginNamespace() const noexcept override { return ns_.c_str(); }

 private:
  nvinfer1::PluginFieldCollection fc_{};
  std::string ns_ = kNamespace;
};

inline nvinfer1::PluginField field_decl(const char* name, nvinfer1::PluginFieldType t) {
  return nvinfer1::PluginField(name, nullptr, t, 1);
}

// data-type ids used by the reference's `type_id` fields are nvinfer1::DataType values
// (T/tensorrt_llm/functional.py:2884-2886 `int(str_dtype_to_trt(dtype))`)
inline bool is_half(int32_t type_id) { return type_id == (int32_t) nvinfer1::DataType::kHALF; }

// ---- NCCL communicators (P/common/plugin.cpp:20-24 keeps a process-global map keyed by the rank set) ----
struct CommHandle;
CommHandle* find_comm(const std::vector<int32_t>& group);
int comm_allreduce_half(CommHandle* c, const void* in, void* out, size_t count, cudaStream_t stream);
int comm_allgather_half(CommHandle* c, const void* in, void* out, size_t count_per_rank, cudaStream_t stream);
int comm_size(const CommHandle* c);

}  // namespace plugins
}  // namespace tb
