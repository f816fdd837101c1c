// TEST INFRASTRUCTURE (tests/cusim): stand-in for <cuda.h> -- the driver types wgb_api.cpp names.
#pragma once
typedef int CUresult;
enum { CUDA_SUCCESS = 0, CUDA_ERROR_INVALID_VALUE = 1, CUDA_ERROR_NOT_FOUND = 500, CUDA_ERROR_LAUNCH_FAILED = 719 };
typedef struct cusimModule* CUmodule;
typedef struct cusimFunction* CUfunction;
typedef struct cusimStream* CUstream;

// tensor maps (cuTensorMapEncodeTiled): the model keeps base / extents / pitch / box in the opaque words
typedef unsigned int cuuint32_t;
typedef unsigned long long cuuint64_t;
typedef struct __attribute__((aligned(64))) { cuuint64_t opaque[16]; } CUtensorMap;
typedef enum { CU_TENSOR_MAP_DATA_TYPE_UINT32 = 4 } CUtensorMapDataType;
typedef enum { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 } CUtensorMapInterleave;
typedef enum { CU_TENSOR_MAP_SWIZZLE_NONE = 0 } CUtensorMapSwizzle;
typedef enum { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0 } CUtensorMapL2promotion;
typedef enum { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 } CUtensorMapFloatOOBfill;
