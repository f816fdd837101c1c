// TEST INFRASTRUCTURE (tests/cusim): stand-in for <nvrtc.h>.  The "compiler" is g++ over the pipeline translation unit
// with tests/cusim/cusim_device.h force-included; the "cubin" it returns is the path of the shared object it built.
#pragma once
#include <stddef.h>
typedef int nvrtcResult;
enum { NVRTC_SUCCESS = 0, NVRTC_ERROR_COMPILATION = 6 };
typedef struct cusimProgram* nvrtcProgram;
extern "C" {
nvrtcResult nvrtcCreateProgram(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
nvrtcResult nvrtcDestroyProgram(nvrtcProgram*);
nvrtcResult nvrtcCompileProgram(nvrtcProgram, int, const char* const*);
nvrtcResult nvrtcGetCUBINSize(nvrtcProgram, size_t*);
nvrtcResult nvrtcGetCUBIN(nvrtcProgram, char*);
nvrtcResult nvrtcGetProgramLogSize(nvrtcProgram, size_t*);
nvrtcResult nvrtcGetProgramLog(nvrtcProgram, char*);
const char* nvrtcGetErrorString(nvrtcResult);
}
