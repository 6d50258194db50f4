// Clean-room subset of the TensorRT plugin interface (namespace nvinfer1) — only what the
// reference's plugins for this path use: IPluginV2DynamicExt, IPluginCreator,
// IPluginRegistry, PluginField(Collection), Dims / DimsExprs / IExprBuilder, PluginTensorDesc.
//
// TensorRT is not installed in this image (no NvInfer.h, no libnvinfer).  When the real headers are
// available, build with -DTB_HAVE_TENSORRT -I<TensorRT>/include and this file forwards to them; the
// plugin sources compile unchanged against either.  Method names, argument meaning and declaration
// order follow the public TensorRT 8.6/9.0 API (NvInferRuntimeCommon.h / NvInferRuntime.h) as used by
// the reference: P/gptAttentionPlugin/gptAttentionPlugin.h:52-88, P/gptAttentionCommon/gptAttentionCommon.h:36-76.
#pragma once
#ifdef TB_HAVE_TENSORRT
#include <NvInferRuntime.h>
#else
#include <cstddef>
#include <cstdint>
#include <cuda_runtime_api.h>

#define NV_TENSORRT_MAJOR 9
#define NV_TENSORRT_MINOR 0
#define NV_TENSORRT_PATCH 0
#define NV_TENSORRT_VERSION ((NV_TENSORRT_MAJOR * 1000) + (NV_TENSORRT_MINOR * 100) + NV_TENSORRT_PATCH)

struct cudnnContext;
struct cublasContext;

namespace nvinfer1 {

using AsciiChar = char;

enum class DataType : int32_t { kFLOAT = 0, kHALF = 1, kINT8 = 2, kINT32 = 3, kBOOL = 4, kUINT8 = 5, kFP8 = 6, kBF16 = 7, kINT64 = 8 };
enum class TensorFormat : int32_t { kLINEAR = 0 };
using PluginFormat = TensorFormat;

class Dims32 {
 public:
  static constexpr int32_t MAX_DIMS{8};
  int32_t nbDims;
  int32_t d[MAX_DIMS];
};
using Dims = Dims32;

enum class PluginFieldType : int32_t { kFLOAT16 = 0, kFLOAT32 = 1, kFLOAT64 = 2, kINT8 = 3, kINT16 = 4, kINT32 = 5, kCHAR = 6, kDIMS = 7, kUNKNOWN = 8 };

class PluginField {
 public:
  AsciiChar const* name;
  void const* data;
  PluginFieldType type;
  int32_t length;
  PluginField(AsciiChar const* const name_ = nullptr, void const* const data_ = nullptr,
              PluginFieldType const type_ = PluginFieldType::kUNKNOWN, int32_t const length_ = 0) noexcept
      : name(name_), data(data_), type(type_), length(length_) {}
};

struct PluginFieldCollection {
  int32_t nbFields;
  PluginField const* fields;
};

struct PluginTensorDesc {
  Dims dims;
  DataType type;
  TensorFormat format;
  float scale;
};

struct DynamicPluginTensorDesc {
  PluginTensorDesc desc;
  Dims min;
  Dims max;
};

enum class DimensionOperation : int32_t { kSUM = 0, kPROD = 1, kMAX = 2, kMIN = 3, kSUB = 4, kEQUAL = 5, kLESS = 6, kFLOOR_DIV = 7, kCEIL_DIV = 8 };

class IDimensionExpr {
 public:
  virtual bool isConstant() const noexcept = 0;
  virtual int32_t getConstantValue() const noexcept = 0;

 protected:
  virtual ~IDimensionExpr() noexcept = default;
};

class IExprBuilder {
 public:
  virtual IDimensionExpr const* constant(int32_t value) noexcept = 0;
  virtual IDimensionExpr const* operation(DimensionOperation op, IDimensionExpr const& first,
                                          IDimensionExpr const& second) noexcept = 0;

 protected:
  virtual ~IExprBuilder() noexcept = default;
};

class DimsExprs {
 public:
  int32_t nbDims;
  IDimensionExpr const* d[Dims::MAX_DIMS];
};

class IGpuAllocator;

class ILogger {
 public:
  enum class Severity : int32_t { kINTERNAL_ERROR = 0, kERROR = 1, kWARNING = 2, kINFO = 3, kVERBOSE = 4 };
  virtual void log(Severity severity, AsciiChar const* msg) noexcept = 0;
  virtual ~ILogger() = default;
};

class IPluginV2 {
 public:
  virtual int32_t getTensorRTVersion() const noexcept { return NV_TENSORRT_VERSION; }
  virtual AsciiChar const* getPluginType() const noexcept = 0;
  virtual AsciiChar const* getPluginVersion() const noexcept = 0;
  virtual int32_t getNbOutputs() const noexcept = 0;
  virtual Dims getOutputDimensions(int32_t index, Dims const* inputs, int32_t nbInputDims) noexcept = 0;
  virtual bool supportsFormat(DataType type, PluginFormat format) const noexcept = 0;
  virtual void configureWithFormat(Dims const* inputDims, int32_t nbInputs, Dims const* outputDims, int32_t nbOutputs,
                                   DataType type, PluginFormat format, int32_t maxBatchSize) noexcept = 0;
  virtual int32_t initialize() noexcept = 0;
  virtual void terminate() noexcept = 0;
  virtual size_t getWorkspaceSize(int32_t maxBatchSize) const noexcept = 0;
  virtual int32_t enqueue(int32_t batchSize, void const* const* inputs, void* const* outputs, void* workspace,
                          cudaStream_t stream) noexcept = 0;
  virtual size_t getSerializationSize() const noexcept = 0;
  virtual void serialize(void* buffer) const noexcept = 0;
  virtual void destroy() noexcept = 0;
  virtual IPluginV2* clone() const noexcept = 0;
  virtual void setPluginNamespace(AsciiChar const* pluginNamespace) noexcept = 0;
  virtual AsciiChar const* getPluginNamespace() const noexcept = 0;

 protected:
  IPluginV2() = default;
  virtual ~IPluginV2() noexcept = default;
};

class IPluginV2Ext : public IPluginV2 {
 public:
  virtual DataType getOutputDataType(int32_t index, DataType const* inputTypes, int32_t nbInputs) const noexcept = 0;
  virtual bool isOutputBroadcastAcrossBatch(int32_t outputIndex, bool const* inputIsBroadcasted,
                                            int32_t nbInputs) const noexcept = 0;
  virtual bool canBroadcastInputAcrossBatch(int32_t inputIndex) const noexcept = 0;
  virtual void configurePlugin(Dims const* inputDims, int32_t nbInputs, Dims const* outputDims, int32_t nbOutputs,
                               DataType const* inputTypes, DataType const* outputTypes, bool const* inputIsBroadcast,
                               bool const* outputIsBroadcast, PluginFormat floatFormat, int32_t maxBatchSize) noexcept = 0;
  virtual void attachToContext(cudnnContext*, cublasContext*, IGpuAllocator*) noexcept {}
  virtual void detachFromContext() noexcept {}
  IPluginV2Ext* clone() const noexcept override = 0;

 protected:
  void configureWithFormat(Dims const*, int32_t, Dims const*, int32_t, DataType, PluginFormat, int32_t) noexcept override {}
};

class IPluginV2DynamicExt : public IPluginV2Ext {
 public:
  IPluginV2DynamicExt* clone() const noexcept override = 0;
  virtual DimsExprs getOutputDimensions(int32_t outputIndex, DimsExprs const* inputs, int32_t nbInputs,
                                        IExprBuilder& exprBuilder) noexcept = 0;
  static constexpr int32_t kFORMAT_COMBINATION_LIMIT = 100;
  virtual bool supportsFormatCombination(int32_t pos, PluginTensorDesc const* inOut, int32_t nbInputs,
                                         int32_t nbOutputs) noexcept = 0;
  virtual void configurePlugin(DynamicPluginTensorDesc const* in, int32_t nbInputs, DynamicPluginTensorDesc const* out,
                               int32_t nbOutputs) noexcept = 0;
  virtual size_t getWorkspaceSize(PluginTensorDesc const* inputs, int32_t nbInputs, PluginTensorDesc const* outputs,
                                  int32_t nbOutputs) const noexcept = 0;
  virtual int32_t enqueue(PluginTensorDesc const* inputDesc, PluginTensorDesc const* outputDesc,
                          void const* const* inputs, void* const* outputs, void* workspace,
                          cudaStream_t stream) noexcept = 0;

 protected:
  // the implicit-batch entry points of the base classes are not used by dynamic-shape plugins
  Dims getOutputDimensions(int32_t, Dims const*, int32_t) noexcept override { return Dims{-1, {}}; }
  bool isOutputBroadcastAcrossBatch(int32_t, bool const*, int32_t) const noexcept override { return false; }
  bool canBroadcastInputAcrossBatch(int32_t) const noexcept override { return true; }
  bool supportsFormat(DataType, PluginFormat) const noexcept override { return false; }
  void configurePlugin(Dims const*, int32_t, Dims const*, int32_t, DataType const*, DataType const*, bool const*,
                       bool const*, PluginFormat, int32_t) noexcept override {}
  size_t getWorkspaceSize(int32_t) const noexcept override { return 0; }
  int32_t enqueue(int32_t, void const* const*, void* const*, void*, cudaStream_t) noexcept override { return 1; }
};

class IPluginCreator {
 public:
  virtual int32_t getTensorRTVersion() const noexcept { return NV_TENSORRT_VERSION; }
  virtual AsciiChar const* getPluginName() const noexcept = 0;
  virtual AsciiChar const* getPluginVersion() const noexcept = 0;
  virtual PluginFieldCollection const* getFieldNames() noexcept = 0;
  virtual IPluginV2* createPlugin(AsciiChar const* name, PluginFieldCollection const* fc) noexcept = 0;
  virtual IPluginV2* deserializePlugin(AsciiChar const* name, void const* serialData, size_t serialLength) noexcept = 0;
  virtual void setPluginNamespace(AsciiChar const* pluginNamespace) noexcept = 0;
  virtual AsciiChar const* getPluginNamespace() const noexcept = 0;
  IPluginCreator() = default;
  virtual ~IPluginCreator() = default;
};

class IPluginRegistry {
 public:
  virtual bool registerCreator(IPluginCreator& creator, AsciiChar const* const pluginNamespace) noexcept = 0;
  virtual IPluginCreator* const* getPluginCreatorList(int32_t* const numCreators) const noexcept = 0;
  virtual IPluginCreator* getPluginCreator(AsciiChar const* const pluginName, AsciiChar const* const pluginVersion,
                                           AsciiChar const* const pluginNamespace = "") noexcept = 0;
  virtual bool deregisterCreator(IPluginCreator const& creator) noexcept = 0;

 protected:
  virtual ~IPluginRegistry() noexcept = default;
};

}  // namespace nvinfer1

extern "C" nvinfer1::IPluginRegistry* getPluginRegistry() noexcept;
#endif  // TB_HAVE_TENSORRT
