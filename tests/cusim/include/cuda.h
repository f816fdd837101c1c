// TEST INFRASTRUCTURE (tests/cusim): stand-in for <cuda.h> -- the driver types wgb_api.cpp names.
#pragma once
typedef int CUresult;
enum { CUDA_SUCCESS = 0, CUDA_ERROR_INVALID_VALUE = 1, CUDA_ERROR_NOT_FOUND = 500, CUDA_ERROR_LAUNCH_FAILED = 719 };
typedef struct cusimModule* CUmodule;
typedef struct cusimFunction* CUfunction;
typedef struct cusimStream* CUstream;
