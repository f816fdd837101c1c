// wgpu_b200.hpp -- C++ host side above the C ABI (include/wgpu_b200.h).
//
// The reference is a Rust crate whose objects are `Clone`-over-`Arc` handles with wgpu's method names
// (wgpu-cpu/src/{instance,adapter,device,buffer,texture,command}.rs, render_pass/mod.rs).  There is no Rust
// toolchain in this build image, so the compiled-language host mirrors that surface in C++: one RAII handle
// class per backend object (copy = wgb_retain, destroy = wgb_release, like Arc clone / drop), the same method
// names and argument meaning as the wgpu calls the reference implements, and errors raised as exceptions
// where the reference panics (device.rs:46-48, state.rs:243).  Header only; link with -lwgpu_b200.
// examples/hello_mesh.cpp is the reference's examples/hello_mesh.rs flow written against it.
#pragma once

#include "wgpu_b200.h"

#include <array>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace wgb {

struct Error : std::runtime_error {
    wgb_status status;
    Error(wgb_status s, const char* msg) : std::runtime_error(msg ? msg : "wgpu-b200 error"), status(s) {}
};
inline void check(wgb_status s) {
    if (s != 0) throw Error(s, wgb_last_error());
}

// Arc-like handle: copying retains, destruction releases
template <class H>
class Handle {
  public:
    Handle() = default;
    explicit Handle(H h) : h_(h) {}
    Handle(const Handle& o) : h_(o.h_) { if (h_) wgb_retain(reinterpret_cast<wgb_object>(h_)); }
    Handle(Handle&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    Handle& operator=(Handle o) noexcept { std::swap(h_, o.h_); return *this; }
    ~Handle() { if (h_) wgb_release(reinterpret_cast<wgb_object>(h_)); }
    H get() const { return h_; }
    explicit operator bool() const { return h_ != nullptr; }

  private:
    H h_ = nullptr;
};

struct TextureView : Handle<wgb_texture_view> { using Handle::Handle; };
struct Sampler : Handle<wgb_sampler> { using Handle::Handle; };
struct ShaderModule : Handle<wgb_shader_module> { using Handle::Handle; };
struct BindGroupLayout : Handle<wgb_bind_group_layout> { using Handle::Handle; };
struct PipelineLayout : Handle<wgb_pipeline_layout> { using Handle::Handle; };
struct BindGroup : Handle<wgb_bind_group> { using Handle::Handle; };
struct CommandBuffer : Handle<wgb_command_buffer> { using Handle::Handle; };

struct RenderPipeline : Handle<wgb_render_pipeline> {
    using Handle::Handle;
    std::string get_source() const {       // the generated CUDA translation unit (diagnostics)
        char* p = nullptr;
        check(wgb_render_pipeline_get_source(get(), &p));
        std::string s(p);
        wgb_free(p);
        return s;
    }
};

struct Buffer : Handle<wgb_buffer> {
    using Handle::Handle;
    // BufferInterface::map_async / get_mapped_range / unmap (buffer.rs:112-172)
    void map_async(uint32_t mode, uint64_t offset = 0, uint64_t size = WGB_WHOLE_SIZE) const { check(wgb_buffer_map_async(get(), mode, offset, size, nullptr, nullptr)); }
    void* get_mapped_range(uint64_t offset = 0, uint64_t size = WGB_WHOLE_SIZE) const {
        void* p = nullptr;
        check(wgb_buffer_get_mapped_range(get(), offset, size, &p));
        return p;
    }
    void unmap() const { check(wgb_buffer_unmap(get())); }
};

struct Texture : Handle<wgb_texture> {
    using Handle::Handle;
    wgb_texture_descriptor desc{};
    TextureView create_view(uint32_t base_array_layer = 0) const {                      // texture.rs:52-75
        wgb_texture_view_descriptor d{base_array_layer, 0};
        wgb_texture_view v = nullptr;
        check(wgb_texture_create_view(get(), &d, &v));
        return TextureView(v);
    }
    // what wgpu_cpu::image::rgba_texture_image / dump_texture observe (lib.rs:111-173)
    std::vector<uint8_t> read() const {
        const uint32_t bpp = desc.format == WGB_TEXTURE_FORMAT_R8_UNORM ? 1 : desc.format == WGB_TEXTURE_FORMAT_RG8_UNORM ? 2 : 4;
        std::vector<uint8_t> out((size_t)desc.width * desc.height * desc.depth_or_array_layers * bpp);
        check(wgb_texture_read(get(), out.data(), out.size()));
        return out;
    }
    // not waited for: `dst` is page-locked and holds the texels after Device::wait_readbacks()
    void read_pinned_async(void* dst, uint64_t size) const { check(wgb_texture_read_pinned_async(get(), dst, size)); }
    void dump_png(const std::string& path) const { check(wgb_texture_dump_png(get(), path.c_str())); }
};

struct VertexBufferLayout {
    uint64_t array_stride = 0;
    uint32_t step_mode = WGB_VERTEX_STEP_MODE_VERTEX;
    std::vector<wgb_vertex_attribute> attributes;     // {format, offset, shader_location}
};
struct DepthStencilState {
    uint32_t format = WGB_TEXTURE_FORMAT_DEPTH32_FLOAT;
    bool depth_write_enabled = true;
    uint32_t depth_compare = WGB_COMPARE_LESS;
};
// wgpu::RenderPipelineDescriptor, the fields the reference reads (pipeline.rs:57-76)
struct RenderPipelineDescriptor {
    ShaderModule vertex_module, fragment_module;      // fragment_module empty -> vertex stage only
    std::string vertex_entry_point = "vs_main", fragment_entry_point = "fs_main";
    std::vector<VertexBufferLayout> vertex_buffers;
    uint32_t topology = WGB_TOPOLOGY_TRIANGLE_LIST, strip_index_format = WGB_INDEX_FORMAT_NONE;
    uint32_t front_face = WGB_FRONT_FACE_CCW, cull_mode = WGB_CULL_MODE_NONE;
    bool has_depth_stencil = false;
    DepthStencilState depth_stencil;
    std::vector<wgb_color_target_state> targets;
};
inline wgb_color_target_state color_target(uint32_t format) {
    wgb_color_target_state t{};
    t.format = format; t.has_blend = 0; t.write_mask = WGB_COLOR_WRITE_ALL;
    t.blend_color = {WGB_BLEND_FACTOR_ONE, WGB_BLEND_FACTOR_ZERO, WGB_BLEND_OPERATION_ADD};
    t.blend_alpha = t.blend_color;
    return t;
}

struct ColorAttachment {
    TextureView view;
    bool clear = true;                                 // LoadOp::Clear(clear_value) / LoadOp::Load
    std::array<double, 4> clear_value{0, 0, 0, 1};
};
struct DepthAttachment {
    TextureView view;
    bool clear = true;
    float clear_value = 1.0f;
};

// RenderPassInterface (render_pass/mod.rs:73-176); end() also runs from the destructor, like Drop (mod.rs:325-329)
class RenderPass {
  public:
    explicit RenderPass(wgb_render_pass p) : p_(p) {}
    RenderPass(RenderPass&& o) noexcept : p_(o.p_) { o.p_ = nullptr; }
    RenderPass(const RenderPass&) = delete;
    ~RenderPass() { if (p_) { wgb_render_pass_end(p_); wgb_release(reinterpret_cast<wgb_object>(p_)); } }
    void set_pipeline(const RenderPipeline& p) { check(wgb_render_pass_set_pipeline(p_, p.get())); }
    void set_bind_group(uint32_t index, const BindGroup& g, const std::vector<uint32_t>& offsets = {}) {
        check(wgb_render_pass_set_bind_group(p_, index, g.get(), offsets.empty() ? nullptr : offsets.data(), (uint32_t)offsets.size()));
    }
    void set_index_buffer(const Buffer& b, uint32_t format, uint64_t offset = 0, uint64_t size = WGB_WHOLE_SIZE) { check(wgb_render_pass_set_index_buffer(p_, b.get(), format, offset, size)); }
    void set_vertex_buffer(uint32_t slot, const Buffer& b, uint64_t offset = 0, uint64_t size = WGB_WHOLE_SIZE) { check(wgb_render_pass_set_vertex_buffer(p_, slot, b.get(), offset, size)); }
    void set_viewport(float x, float y, float w, float h, float min_depth = 0.0f, float max_depth = 1.0f) { check(wgb_render_pass_set_viewport(p_, x, y, w, h, min_depth, max_depth)); }
    void set_scissor_rect(uint32_t x, uint32_t y, uint32_t w, uint32_t h) { check(wgb_render_pass_set_scissor_rect(p_, x, y, w, h)); }
    void set_blend_constant(const std::array<double, 4>& c) { check(wgb_render_pass_set_blend_constant(p_, c.data())); }
    void draw(uint32_t first_vertex, uint32_t vertex_count, uint32_t first_instance = 0, uint32_t instance_count = 1) {
        check(wgb_render_pass_draw(p_, first_vertex, vertex_count, first_instance, instance_count));
    }
    void draw_indexed(uint32_t first_index, uint32_t index_count, int32_t base_vertex = 0, uint32_t first_instance = 0, uint32_t instance_count = 1) {
        check(wgb_render_pass_draw_indexed(p_, first_index, index_count, base_vertex, first_instance, instance_count));
    }
    void end() { check(wgb_render_pass_end(p_)); }

  private:
    wgb_render_pass p_;
};

struct CommandEncoder : Handle<wgb_command_encoder> {
    using Handle::Handle;
    RenderPass begin_render_pass(const std::vector<ColorAttachment>& colors, const DepthAttachment* depth = nullptr) const {   // command.rs:82-87
        std::vector<wgb_color_attachment> ca(colors.size());
        for (size_t i = 0; i < colors.size(); i++) {
            ca[i].view = colors[i].view.get();
            ca[i].load_op = colors[i].clear ? WGB_LOAD_OP_CLEAR : WGB_LOAD_OP_LOAD;
            ca[i].store_op = WGB_STORE_OP_STORE;
            std::memcpy(ca[i].clear_value, colors[i].clear_value.data(), sizeof(ca[i].clear_value));
        }
        wgb_depth_stencil_attachment da{};
        if (depth) {
            da.view = depth->view.get(); da.has_depth_ops = 1;
            da.depth_load_op = depth->clear ? WGB_LOAD_OP_CLEAR : WGB_LOAD_OP_LOAD; da.depth_store_op = WGB_STORE_OP_STORE;
            da.depth_clear_value = depth->clear_value; da.has_stencil_ops = 0;
        }
        wgb_render_pass_descriptor d{(uint32_t)ca.size(), ca.data(), depth ? &da : nullptr};
        wgb_render_pass p = nullptr;
        check(wgb_command_encoder_begin_render_pass(get(), &d, &p));
        return RenderPass(p);
    }
    void copy_texture_to_buffer(const Texture& src, const Buffer& dst, uint32_t bytes_per_row, uint32_t width, uint32_t height) const {
        wgb_texel_copy_texture_info s{src.get(), 0, 0, 0};
        wgb_texel_copy_buffer_info b{dst.get(), 0, bytes_per_row, height};
        check(wgb_command_encoder_copy_texture_to_buffer(get(), &s, &b, width, height));
    }
    CommandBuffer finish() const {                                                        // command.rs:89-98
        wgb_command_buffer cb = nullptr;
        check(wgb_command_encoder_finish(get(), &cb));
        return CommandBuffer(cb);
    }
};

struct Queue : Handle<wgb_queue> {
    using Handle::Handle;
    void write_buffer(const Buffer& b, uint64_t offset, const void* data, uint64_t size) const { check(wgb_queue_write_buffer(get(), b.get(), offset, data, size)); }   // device.rs:332-344
    // zero-copy from page-locked memory; `data` stays valid until wait_uploads()
    void write_buffer_pinned_async(const Buffer& b, uint64_t offset, const void* data, uint64_t size) const { check(wgb_queue_write_buffer_pinned_async(get(), b.get(), offset, data, size)); }
    void wait_uploads() const { check(wgb_queue_wait_uploads(get())); }
    void write_texture(const Texture& t, const void* data, uint64_t size, uint32_t bytes_per_row = 0) const {                                                           // device.rs:371-434
        check(wgb_queue_write_texture(get(), t.get(), 0, 0, data, size, bytes_per_row, t.desc.width, t.desc.height));
    }
    uint64_t submit(const std::vector<CommandBuffer>& cbs) const {                        // device.rs:436-462
        std::vector<wgb_command_buffer> raw;
        for (const auto& c : cbs) raw.push_back(c.get());
        uint64_t index = 0;
        check(wgb_queue_submit(get(), raw.data(), (uint32_t)raw.size(), &index));
        return index;
    }
};

struct Device : Handle<wgb_device> {
    using Handle::Handle;
    Buffer create_buffer(uint64_t size, uint32_t usage, bool mapped_at_creation = false) const {   // device.rs:150-160
        wgb_buffer_descriptor d{size, usage, mapped_at_creation ? 1u : 0u};
        wgb_buffer b = nullptr;
        check(wgb_device_create_buffer(get(), &d, &b));
        return Buffer(b);
    }
    // wgpu::util::DeviceExt::create_buffer_init: created mapped, filled, unmapped (hello_mesh.rs:111-157)
    Buffer create_buffer_init(const void* data, uint64_t size, uint32_t usage) const {
        Buffer b = create_buffer(size, usage, true);
        std::memcpy(b.get_mapped_range(), data, size);
        b.unmap();
        return b;
    }
    Texture create_texture(uint32_t width, uint32_t height, uint32_t format, uint32_t layers = 1) const {   // device.rs:162-175
        Texture t;
        wgb_texture_descriptor d{width, height, layers, 1, 1, format, 0};
        wgb_texture h = nullptr;
        check(wgb_device_create_texture(get(), &d, &h));
        t = Texture(h);
        t.desc = d;
        return t;
    }
    Sampler create_sampler(uint32_t address_u = WGB_ADDRESS_MODE_CLAMP_TO_EDGE, uint32_t address_v = WGB_ADDRESS_MODE_CLAMP_TO_EDGE) const {
        wgb_sampler_descriptor d{address_u, address_v, address_u, WGB_FILTER_MODE_NEAREST, WGB_FILTER_MODE_NEAREST, WGB_FILTER_MODE_NEAREST};
        wgb_sampler s = nullptr;
        check(wgb_device_create_sampler(get(), &d, &s));
        return Sampler(s);
    }
    ShaderModule create_shader_module(const std::string& wgsl) const {                   // device.rs:88-100
        wgb_shader_module_descriptor d{wgsl.c_str(), 0, nullptr};
        wgb_shader_module m = nullptr;
        check(wgb_device_create_shader_module(get(), &d, &m));
        return ShaderModule(m);
    }
    BindGroup create_bind_group(const std::vector<wgb_bind_group_entry>& entries, const BindGroupLayout& layout = BindGroupLayout()) const {   // device.rs:109-116
        wgb_bind_group g = nullptr;
        check(wgb_device_create_bind_group(get(), layout.get(), entries.data(), (uint32_t)entries.size(), &g));
        return BindGroup(g);
    }
    RenderPipeline create_render_pipeline(const RenderPipelineDescriptor& desc) const {   // device.rs:129-134
        std::vector<wgb_vertex_buffer_layout> vbs(desc.vertex_buffers.size());
        for (size_t i = 0; i < vbs.size(); i++) {
            const auto& v = desc.vertex_buffers[i];
            vbs[i] = {v.array_stride, v.step_mode, (uint32_t)v.attributes.size(), v.attributes.data()};
        }
        wgb_render_pipeline_descriptor d{};
        d.vertex_module = desc.vertex_module.get(); d.vertex_entry_point = desc.vertex_entry_point.c_str();
        d.vertex_buffer_count = (uint32_t)vbs.size(); d.vertex_buffers = vbs.data();
        d.topology = desc.topology; d.strip_index_format = desc.strip_index_format;
        d.front_face = desc.front_face; d.cull_mode = desc.cull_mode; d.polygon_mode = WGB_POLYGON_MODE_FILL;
        d.has_depth_stencil = desc.has_depth_stencil ? 1u : 0u;
        d.depth_format = desc.depth_stencil.format; d.depth_write_enabled = desc.depth_stencil.depth_write_enabled ? 1u : 0u;
        d.depth_compare = desc.depth_stencil.depth_compare;
        d.multisample_count = 1;
        d.fragment_module = desc.fragment_module.get(); d.fragment_entry_point = desc.fragment_entry_point.c_str();
        d.target_count = (uint32_t)desc.targets.size(); d.targets = desc.targets.data();
        wgb_render_pipeline p = nullptr;
        check(wgb_device_create_render_pipeline(get(), &d, &p));
        return RenderPipeline(p);
    }
    CommandEncoder create_command_encoder() const {                                       // device.rs:190-196
        wgb_command_encoder e = nullptr;
        check(wgb_device_create_command_encoder(get(), &e));
        return CommandEncoder(e);
    }
    // PollType::Wait { submission_index, timeout } (device.rs:237-295)
    int32_t poll_wait(uint64_t submission_index = WGB_SUBMISSION_ANY, uint64_t timeout_ns = 0) const {
        int32_t out = 0;
        check(wgb_device_poll(get(), 1, submission_index, timeout_ns, &out));
        return out;
    }
    void wait_readbacks() const { check(wgb_device_wait_readbacks(get())); }
    wgb_pass_stats last_pass_stats() const {
        wgb_pass_stats s{};
        check(wgb_device_get_last_pass_stats(get(), &s));
        return s;
    }
};

// SurfaceInterface + SurfaceOutputDetailInterface (surface.rs:51-198); the window is a host pixel sink (wgpu_b200.h)
struct Surface : Handle<wgb_surface> {
    using Handle::Handle;
    wgb_surface_capabilities get_capabilities(wgb_adapter adapter) const {               // surface.rs:52-72
        wgb_surface_capabilities c{};
        check(wgb_surface_get_capabilities(get(), adapter, &c));
        return c;
    }
    void configure(const Device& device, uint32_t width, uint32_t height, uint32_t format = WGB_TEXTURE_FORMAT_BGRA8_UNORM,
                   uint32_t usage = WGB_TEXTURE_USAGE_RENDER_ATTACHMENT) {               // surface.rs:74-114
        wgb_surface_configuration c{usage, format, width, height, WGB_PRESENT_MODE_IMMEDIATE, WGB_COMPOSITE_ALPHA_MODE_OPAQUE, 0, nullptr};
        check(wgb_surface_configure(get(), device.get(), &c));
        width_ = width; height_ = height; format_ = format;
    }
    Texture get_current_texture() const {                                                // surface.rs:116-146
        wgb_texture t = nullptr;
        uint32_t status = 0;
        check(wgb_surface_get_current_texture(get(), &t, &status));
        Texture tex(t);
        tex.desc.width = width_; tex.desc.height = height_; tex.desc.depth_or_array_layers = 1;
        tex.desc.mip_level_count = 1; tex.desc.sample_count = 1; tex.desc.format = format_;
        return tex;
    }
    void present() const { check(wgb_surface_present(get())); }                          // surface.rs:168-193
    void texture_discard() const { check(wgb_surface_texture_discard(get())); }          // surface.rs:195-197
    // the window's pixels as the last present left them (width * height * 4 bytes, BGRA), and the presents so far
    std::pair<const uint8_t*, uint64_t> window_buffer() const {
        const void* p = nullptr;
        uint64_t n = 0, k = 0;
        check(wgb_surface_get_window_buffer(get(), &p, &n, &k));
        return {static_cast<const uint8_t*>(p), k};
    }

  private:
    uint32_t width_ = 0, height_ = 0, format_ = WGB_TEXTURE_FORMAT_BGRA8_UNORM;
};

struct Adapter : Handle<wgb_adapter> {
    using Handle::Handle;
    bool is_surface_supported(const Surface& surface) const {                            // adapter.rs:44-54
        int32_t out = 0;
        check(wgb_adapter_is_surface_supported(get(), surface.get(), &out));
        return out != 0;
    }
    wgb_adapter_info get_info() const {                                                  // adapter.rs:60-75
        wgb_adapter_info i{};
        check(wgb_adapter_get_info(get(), &i));
        return i;
    }
    // AdapterInterface::request_device (adapter.rs:24-42): features = 0 renders exactly what the reference renders
    std::pair<Device, Queue> request_device(int32_t cuda_device = 0, uint32_t band_rank = 0, uint32_t band_count = 1, uint32_t features = 0) const {
        wgb_device_descriptor d{cuda_device, band_rank, band_count, features};
        wgb_device dev = nullptr;
        wgb_queue q = nullptr;
        check(wgb_adapter_request_device(get(), &d, &dev, &q));
        return {Device(dev), Queue(q)};
    }
};

struct Instance : Handle<wgb_instance> {
    using Handle::Handle;
    Adapter request_adapter() const {                                                    // instance.rs:72-102
        wgb_adapter a = nullptr;
        check(wgb_instance_request_adapter(get(), &a));
        return Adapter(a);
    }
    // InstanceInterface::create_surface (instance.rs:51-70): `on_present(user_data, pixels, w, h, bytes_per_row)` stands
    // where softbuffer's Buffer::present does
    Surface create_surface(wgb_present_callback on_present = nullptr, void* user_data = nullptr) const {
        wgb_surface_target t{on_present, user_data};
        wgb_surface s = nullptr;
        check(wgb_instance_create_surface(get(), &t, &s));
        return Surface(s);
    }
};
// wgpu_cpu::instance(Config) (lib.rs:22-27)
inline Instance instance() {
    wgb_instance i = nullptr;
    check(wgb_create_instance(nullptr, &i));
    return Instance(i);
}

}  // namespace wgb
