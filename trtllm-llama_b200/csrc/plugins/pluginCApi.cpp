// C view of the plugin interface (include/trtllm_b200_plugin.h): forwards to the virtual calls
// TensorRT would make, with a concrete IExprBuilder for getOutputDimensions.
#include <deque>

#include "../../../include/trtllm_b200_plugin.h"
#include "pluginBase.h"

using namespace nvinfer1;
using namespace tb::plugins;

extern "C" bool initLibNvInferPlugins(void* logger, const char* libNamespace);

namespace {

static_assert(sizeof(tbp_field) == sizeof(PluginField), "tbp_field must mirror nvinfer1::PluginField");
static_assert(sizeof(tbp_dims) == sizeof(Dims), "tbp_dims must mirror nvinfer1::Dims");
static_assert(sizeof(tbp_tensor_desc) == sizeof(PluginTensorDesc), "tbp_tensor_desc must mirror nvinfer1::PluginTensorDesc");

// every expression is a constant: shapes are concrete by the time the runtime asks
class ConstExpr : public IDimensionExpr {
 public:
  explicit ConstExpr(int32_t v) : v_(v) {}
  bool isConstant() const noexcept override { return true; }
  int32_t getConstantValue() const noexcept override { return v_; }

 private:
  int32_t v_;
};
class ConstExprBuilder : public IExprBuilder {
 public:
  const IDimensionExpr* constant(int32_t v) noexcept override {
    pool_.emplace_back(v);
    return &pool_.back();
  }
  const IDimensionExpr* operation(DimensionOperation op, const IDimensionExpr& a, const IDimensionExpr& b) noexcept override {
    const int32_t x = a.getConstantValue(), y = b.getConstantValue();
    switch (op) {
      case DimensionOperation::kSUM: return constant(x + y);
      case DimensionOperation::kPROD: return constant(x * y);
      case DimensionOperation::kMAX: return constant(x > y ? x : y);
      case DimensionOperation::kMIN: return constant(x < y ? x : y);
      case DimensionOperation::kSUB: return constant(x - y);
      case DimensionOperation::kEQUAL: return constant(x == y);
      case DimensionOperation::kLESS: return constant(x < y);
      case DimensionOperation::kFLOOR_DIV: return constant(y ? x / y : 0);
      case DimensionOperation::kCEIL_DIV: return constant(y ? (x + y - 1) / y : 0);
    }
    return constant(0);
  }

 private:
  std::deque<ConstExpr> pool_;
};

IPluginV2DynamicExt* P(tbp_plugin* p) { return reinterpret_cast<IPluginV2DynamicExt*>(p); }
const IPluginV2DynamicExt* P(const tbp_plugin* p) { return reinterpret_cast<const IPluginV2DynamicExt*>(p); }
IPluginCreator* creator(const char* name, const char* version, const char* ns) {
  return getPluginRegistry()->getPluginCreator(name, version ? version : kVersion, ns ? ns : kNamespace);
}
}  // namespace

extern "C" {

int tbp_init(const char* ns) { return initLibNvInferPlugins(nullptr, ns ? ns : kNamespace) ? 0 : -1; }
int tbp_num_creators(void) {
  int32_t n = 0;
  getPluginRegistry()->getPluginCreatorList(&n);
  return n;
}
const char* tbp_creator_name(int i) {
  int32_t n = 0;
  IPluginCreator* const* l = getPluginRegistry()->getPluginCreatorList(&n);
  return (i >= 0 && i < n) ? l[i]->getPluginName() : nullptr;
}
int tbp_creator_fields(const char* name, const char** names, int max_names) {
  IPluginCreator* c = creator(name, nullptr, nullptr);
  if (!c) return -1;
  const PluginFieldCollection* fc = c->getFieldNames();
  for (int i = 0; names && i < fc->nbFields && i < max_names; ++i) names[i] = fc->fields[i].name;
  return fc->nbFields;
}
tbp_plugin* tbp_create(const char* name, const char* version, const char* ns, const tbp_field* fields, int nb) {
  IPluginCreator* c = creator(name, version, ns);
  if (!c) {
    log_msg(ILogger::Severity::kERROR, "no plugin creator %s version %s in namespace %s", name, version ? version : kVersion,
            ns ? ns : kNamespace);
    return nullptr;
  }
  PluginFieldCollection fc{nb, reinterpret_cast<const PluginField*>(fields)};
  return reinterpret_cast<tbp_plugin*>(static_cast<IPluginV2DynamicExt*>(c->createPlugin(name, &fc)));
}
tbp_plugin* tbp_deserialize(const char* name, const char* version, const char* ns, const void* data, size_t len) {
  IPluginCreator* c = creator(name, version, ns);
  if (!c) return nullptr;
  return reinterpret_cast<tbp_plugin*>(static_cast<IPluginV2DynamicExt*>(c->deserializePlugin(name, data, len)));
}
tbp_plugin* tbp_clone(const tbp_plugin* p) { return reinterpret_cast<tbp_plugin*>(P(p)->clone()); }
void tbp_destroy(tbp_plugin* p) { if (p) P(p)->destroy(); }
const char* tbp_type(const tbp_plugin* p) { return P(p)->getPluginType(); }
const char* tbp_version(const tbp_plugin* p) { return P(p)->getPluginVersion(); }
const char* tbp_namespace(const tbp_plugin* p) { return P(p)->getPluginNamespace(); }
size_t tbp_serialization_size(const tbp_plugin* p) { return P(p)->getSerializationSize(); }
int tbp_serialize(const tbp_plugin* p, void* buf) { P(p)->serialize(buf); return 0; }
int tbp_nb_outputs(const tbp_plugin* p) { return P(p)->getNbOutputs(); }
int tbp_output_dims(tbp_plugin* p, int idx, const tbp_dims* inputs, int nb_inputs, tbp_dims* out) {
  ConstExprBuilder eb;
  std::vector<DimsExprs> in(nb_inputs);
  for (int i = 0; i < nb_inputs; ++i) {
    in[i].nbDims = inputs[i].nb_dims;
    for (int j = 0; j < inputs[i].nb_dims; ++j) in[i].d[j] = eb.constant(inputs[i].d[j]);
  }
  const DimsExprs r = P(p)->getOutputDimensions(idx, in.data(), nb_inputs, eb);
  out->nb_dims = r.nbDims;
  for (int j = 0; j < r.nbDims; ++j) out->d[j] = r.d[j]->getConstantValue();
  return 0;
}
int tbp_output_dtype(const tbp_plugin* p, int idx, const int32_t* types, int nb) {
  return (int) P(p)->getOutputDataType(idx, reinterpret_cast<const DataType*>(types), nb);
}
int tbp_supports_format(tbp_plugin* p, int pos, const tbp_tensor_desc* io, int nb_in, int nb_out) {
  return P(p)->supportsFormatCombination(pos, reinterpret_cast<const PluginTensorDesc*>(io), nb_in, nb_out) ? 1 : 0;
}
size_t tbp_workspace_size(const tbp_plugin* p, const tbp_tensor_desc* in, int nb_in, const tbp_tensor_desc* out, int nb_out) {
  return P(p)->getWorkspaceSize(reinterpret_cast<const PluginTensorDesc*>(in), nb_in,
                                reinterpret_cast<const PluginTensorDesc*>(out), nb_out);
}
int tbp_initialize(tbp_plugin* p) { return P(p)->initialize(); }
int tbp_enqueue(tbp_plugin* p, const tbp_tensor_desc* id, const tbp_tensor_desc* od, const void* const* inputs,
                void* const* outputs, void* workspace, tb_stream_t stream) {
  return P(p)->enqueue(reinterpret_cast<const PluginTensorDesc*>(id), reinterpret_cast<const PluginTensorDesc*>(od), inputs,
                       outputs, workspace, reinterpret_cast<cudaStream_t>(stream));
}
}
