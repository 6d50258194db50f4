// internal: every kernel TU sees the public C ABI it implements
#pragma once
#include <cuda_runtime.h>
#include "../../include/trtllm_b200.h"
