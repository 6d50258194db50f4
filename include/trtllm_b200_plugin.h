/*
 * trtllm_b200_plugin.h — C view of the plugin library (the same shared object, which also exports the
 * reference's own entry points `initLibNvInferPlugins`, `getPluginRegistry`, `getInferLibVersion`:
 * P/api/InferPlugin.cpp:149-171, P/exports.map:19-31).
 *
 * TensorRT drives plugins through C++ virtual calls (IPluginCreator / IPluginV2DynamicExt).  These
 * functions expose exactly those calls over plain C so that a host without TensorRT (the Python
 * tests, the engine in trtllm_b200_runtime.h, a cgo/ctypes binding) can play TensorRT's role:
 * look a creator up by (name, version, namespace), create or deserialise a plugin, query shapes and
 * types, size the workspace, enqueue.  Struct layouts equal nvinfer1::PluginField / Dims /
 * PluginTensorDesc so the arrays are passed through unchanged.
 */
#ifndef TRTLLM_B200_PLUGIN_H
#define TRTLLM_B200_PLUGIN_H

#include <stddef.h>
#include <stdint.h>

#include "trtllm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tbp_plugin tbp_plugin; /* opaque nvinfer1::IPluginV2DynamicExt* */

/* == nvinfer1::PluginField; type: nvinfer1::PluginFieldType (3 int8, 5 int32, 1 float32) */
typedef struct { const char* name; const void* data; int32_t type; int32_t length; } tbp_field;
/* == nvinfer1::Dims */
typedef struct { int32_t nb_dims; int32_t d[8]; } tbp_dims;
/* == nvinfer1::PluginTensorDesc; type: nvinfer1::DataType (0 float, 1 half, 2 int8, 3 int32); format 0 = linear */
typedef struct { tbp_dims dims; int32_t type; int32_t format; float scale; } tbp_tensor_desc;

/* initLibNvInferPlugins(NULL, ns) — T/tensorrt_llm/plugin/plugin.py:7-22 calls it with "tensorrt_llm". */
int tbp_init(const char* lib_namespace);
int tbp_num_creators(void);
const char* tbp_creator_name(int index);
/* number of declared fields; names[i] (if names != NULL) receives up to max_names field names */
int tbp_creator_fields(const char* name, const char** names, int max_names);

/* IPluginCreator::createPlugin / deserializePlugin; NULL on error (the error is logged). */
tbp_plugin* tbp_create(const char* name, const char* version, const char* ns, const tbp_field* fields, int nb_fields);
tbp_plugin* tbp_deserialize(const char* name, const char* version, const char* ns, const void* data, size_t length);
tbp_plugin* tbp_clone(const tbp_plugin* p);
void tbp_destroy(tbp_plugin* p);

const char* tbp_type(const tbp_plugin* p);
const char* tbp_version(const tbp_plugin* p);
const char* tbp_namespace(const tbp_plugin* p);
size_t tbp_serialization_size(const tbp_plugin* p);
int tbp_serialize(const tbp_plugin* p, void* buffer);
int tbp_nb_outputs(const tbp_plugin* p);
/* getOutputDimensions evaluated on concrete input dims */
int tbp_output_dims(tbp_plugin* p, int output_index, const tbp_dims* inputs, int nb_inputs, tbp_dims* out);
int tbp_output_dtype(const tbp_plugin* p, int output_index, const int32_t* input_types, int nb_inputs);
int tbp_supports_format(tbp_plugin* p, int pos, const tbp_tensor_desc* in_out, int nb_inputs, int nb_outputs);
size_t tbp_workspace_size(const tbp_plugin* p, const tbp_tensor_desc* inputs, int nb_inputs,
                          const tbp_tensor_desc* outputs, int nb_outputs);
int tbp_initialize(tbp_plugin* p);
/* IPluginV2DynamicExt::enqueue: device pointers except where the plugin documents a host tensor
 * (GPTAttention input 3, past_key_value_length). */
int tbp_enqueue(tbp_plugin* p, const tbp_tensor_desc* input_desc, const tbp_tensor_desc* output_desc,
                const void* const* inputs, void* const* outputs, void* workspace, tb_stream_t stream);

/* NCCL communicator bootstrap for the AllReduce / AllGather plugins (replaces the MPI exchange of
 * P/ncclPlugin/allreducePlugin.cpp:128-167): rank 0 of a group creates the id, the host distributes the
 * 128 bytes (torch.distributed / a file), every rank of the group calls tb_comm_init. */
int tb_comm_unique_id(void* out128);
int tb_comm_init(const void* unique_id128, const int32_t* group, int group_size, int rank_in_group);

#ifdef __cplusplus
}
#endif
#endif /* TRTLLM_B200_PLUGIN_H */
