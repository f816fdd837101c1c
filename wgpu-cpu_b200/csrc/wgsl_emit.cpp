// wgsl_emit.cpp -- WGSL -> CUDA C++ emitter (placeholder until the front end lands; see wgb_api.cpp).
#include <stdexcept>
#include <string>
#include <cstdint>

std::string wgb_emit_wgsl(const std::string& wgsl, uint32_t stage, const std::string& entry_point) {
    (void)wgsl; (void)stage;
    throw std::runtime_error("WGSL front end not built yet: supply emitted CUDA for entry point '" + entry_point + "'");
}
