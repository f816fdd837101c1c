// TEST INFRASTRUCTURE (tests/cusim): wgb_api.cpp dlopen()s libnvrtc by name; in the model build those two calls are
// routed to the stand-in compiler of cusim_host.cpp.
#pragma once
#include_next <dlfcn.h>
extern "C" void* cusim_dlopen(const char* name, int flags);
extern "C" void* cusim_dlsym(void* lib, const char* name);
#define dlopen cusim_dlopen
#define dlsym cusim_dlsym
