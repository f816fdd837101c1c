// TEST INFRASTRUCTURE (tests/cusim): stand-in for <cuda_runtime.h> -- the subset wgb_api.cpp uses, implemented on host
// memory by ../cusim_host.cpp.  Streams execute at issue (every call is synchronous), events are timestamps.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
typedef struct cusimStream* cudaStream_t;
typedef struct cusimEvent* cudaEvent_t;
typedef unsigned long long cudaTextureObject_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEnableDefault = 0, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; size_t totalGlobalMem; };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct uchar4 { unsigned char x, y, z, w; };
struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
struct cudaChannelFormatDesc { int x, y, z, w, f; };
template <class T> inline cudaChannelFormatDesc cudaCreateChannelDesc() { return cudaChannelFormatDesc{(int)sizeof(T) * 2, (int)sizeof(T) * 2, (int)sizeof(T) * 2, (int)sizeof(T) * 2, 1}; }
enum cudaResourceType { cudaResourceTypeArray = 0, cudaResourceTypeLinear = 2 };
enum cudaTextureReadMode { cudaReadModeElementType = 0 };
enum cudaTextureFilterMode { cudaFilterModePoint = 0 };
enum cudaTextureAddressMode { cudaAddressModeWrap = 0, cudaAddressModeClamp = 1 };
struct cudaResourceDesc { cudaResourceType resType; struct { struct { void* devPtr; cudaChannelFormatDesc desc; size_t sizeInBytes; } linear; } res; };
struct cudaTextureDesc { cudaTextureAddressMode addressMode[3]; cudaTextureFilterMode filterMode; cudaTextureReadMode readMode; int normalizedCoords; };

extern "C" {
const char* cudaGetErrorString(cudaError_t);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDeviceCount(int*);
cudaError_t cudaGetDevice(int*);
cudaError_t cudaSetDevice(int);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp*, int);
cudaError_t cudaMalloc(void**, size_t);
cudaError_t cudaFree(void*);
cudaError_t cudaMallocHost(void**, size_t);
cudaError_t cudaFreeHost(void*);
cudaError_t cudaMemcpy(void*, const void*, size_t, cudaMemcpyKind);
cudaError_t cudaMemcpyAsync(void*, const void*, size_t, cudaMemcpyKind, cudaStream_t);
cudaError_t cudaMemcpy2DAsync(void*, size_t, const void*, size_t, size_t, size_t, cudaMemcpyKind, cudaStream_t);
cudaError_t cudaMemsetAsync(void*, int, size_t, cudaStream_t);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t*, unsigned);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t*, unsigned, int);
cudaError_t cudaDeviceGetStreamPriorityRange(int*, int*);
cudaError_t cudaStreamDestroy(cudaStream_t);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned);
cudaError_t cudaEventCreate(cudaEvent_t*);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t*, unsigned);
cudaError_t cudaEventDestroy(cudaEvent_t);
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t);
cudaError_t cudaEventSynchronize(cudaEvent_t);
cudaError_t cudaEventQuery(cudaEvent_t);
cudaError_t cudaEventElapsedTime(float*, cudaEvent_t, cudaEvent_t);
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes*, const void*);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*);
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned);
cudaError_t cudaIpcCloseMemHandle(void*);
cudaError_t cudaCreateTextureObject(cudaTextureObject_t*, const cudaResourceDesc*, const cudaTextureDesc*, const void*);
cudaError_t cudaDestroyTextureObject(cudaTextureObject_t);
cudaError_t cudaGetDriverEntryPoint(const char*, void**, unsigned long long, cudaDriverEntryPointQueryResult*);
}
