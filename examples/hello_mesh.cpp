// The reference's examples/hello_mesh.rs (wgpu-cpu/examples/hello_mesh.rs:111-339) written against the C++ host
// (include/wgpu_b200.hpp): instance -> adapter -> device, shader module, buffers through create_buffer_init, render
// pipeline (front face Cw, cull Back, Depth32Float Less + write, Rgba8UnormSrgb target), one encoder, one render
// pass (clear black / 1.0), draw_indexed, submit, poll(Wait), dump_texture.
//
//   hello_mesh <shader.wgsl> <vertices.bin> <indices.bin> <uniform.bin> <width> <height> <out-prefix> [frames]
//
// With `frames` the example takes the reference's *windowed* flow instead (hello_mesh.rs:236-250, 380-410, 476-490):
// create_surface, get_capabilities, formats[0] as the pipeline's target format, configure, and per frame
// get_current_texture -> render pass -> submit -> present.  The window is the library's host pixel sink: the present
// callback keeps the frame, and the last one is written to <out-prefix>.window (raw BGRA texels).
//
// vertices: {pos vec4f, colour vec4f} (32 bytes, hello_mesh.rs:558-571); indices: u32; uniform: the 64-byte camera
// matrix.  Writes <out-prefix>.png (dump_texture), <out-prefix>.rgba and <out-prefix>.depth (raw texels, compared
// bit for bit against the CPU oracle by tests/test_cpp_host_gpu.py).
#include "wgpu_b200.hpp"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>

static std::vector<char> read_file(const char* path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", path); std::exit(2); }
    return std::vector<char>(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}
static void write_file(const std::string& path, const void* data, size_t size) {
    std::ofstream f(path, std::ios::binary);
    f.write(static_cast<const char*>(data), (std::streamsize)size);
}

struct Window { std::vector<uint8_t> pixels; uint32_t width = 0, height = 0, presents = 0; };
static void on_present(void* user_data, const void* pixels, uint32_t width, uint32_t height, uint32_t bytes_per_row) {
    Window* w = static_cast<Window*>(user_data);
    w->width = width; w->height = height; w->presents++;
    w->pixels.assign(static_cast<const uint8_t*>(pixels), static_cast<const uint8_t*>(pixels) + (size_t)bytes_per_row * height);
}

int main(int argc, char** argv) {
    if (argc != 8 && argc != 9) { std::fprintf(stderr, "usage: %s shader.wgsl vertices.bin indices.bin uniform.bin width height out-prefix\n", argv[0]); return 2; }
    try {
        const std::vector<char> wgsl = read_file(argv[1]), vertices = read_file(argv[2]), indices = read_file(argv[3]), uniform = read_file(argv[4]);
        const uint32_t width = (uint32_t)std::atoi(argv[5]), height = (uint32_t)std::atoi(argv[6]);
        const std::string out = argv[7];
        const int frames = argc == 9 ? std::atoi(argv[8]) : 0;

        wgb::Instance instance = wgb::instance();                                  // wgpu_cpu::instance(Default::default())
        wgb::Adapter adapter = instance.request_adapter();
        std::cout << "adapter: " << adapter.get_info().name << "\n";
        auto [device, queue] = adapter.request_device(0);

        wgb::ShaderModule shader = device.create_shader_module(std::string(wgsl.begin(), wgsl.end()));
        wgb::Buffer vertex_buffer = device.create_buffer_init(vertices.data(), vertices.size(), WGB_BUFFER_USAGE_VERTEX);
        wgb::Buffer index_buffer = device.create_buffer_init(indices.data(), indices.size(), WGB_BUFFER_USAGE_INDEX);
        wgb::Buffer camera_buffer = device.create_buffer_init(uniform.data(), uniform.size(), WGB_BUFFER_USAGE_UNIFORM | WGB_BUFFER_USAGE_COPY_DST);

        wgb_bind_group_entry camera_entry{};
        camera_entry.binding = 0; camera_entry.kind = WGB_BINDING_BUFFER; camera_entry.buffer = camera_buffer.get();
        camera_entry.offset = 0; camera_entry.size = WGB_WHOLE_SIZE;
        wgb::BindGroup bind_group = device.create_bind_group({camera_entry});

        wgb::RenderPipelineDescriptor pd;
        pd.vertex_module = shader; pd.fragment_module = shader;
        wgb::VertexBufferLayout layout;
        layout.array_stride = 32;
        layout.attributes = {{WGB_VERTEX_FORMAT_FLOAT32X4, 0, 0}, {WGB_VERTEX_FORMAT_FLOAT32X4, 16, 1}};
        pd.vertex_buffers = {layout};
        pd.front_face = WGB_FRONT_FACE_CW; pd.cull_mode = WGB_CULL_MODE_BACK;       // hello_mesh.rs:193-199
        pd.has_depth_stencil = true;                                               // Depth32Float, Less, write (hello_mesh.rs:200-206)
        // off-screen: Rgba8UnormSrgb (hello_mesh.rs:263); on a surface: the first format it offers (hello_mesh.rs:240-241)
        Window window;
        wgb::Surface surface;
        uint32_t target_format = WGB_TEXTURE_FORMAT_RGBA8_UNORM_SRGB;
        if (frames > 0) {
            surface = instance.create_surface(on_present, &window);
            if (!adapter.is_surface_supported(surface)) { std::fprintf(stderr, "surface not supported\n"); return 1; }
            target_format = surface.get_capabilities(adapter.get()).formats[0];
            surface.configure(device, width, height, target_format);
        }
        pd.targets = {wgb::color_target(target_format)};
        wgb::RenderPipeline pipeline = device.create_render_pipeline(pd);

        wgb::Texture depth = device.create_texture(width, height, WGB_TEXTURE_FORMAT_DEPTH32_FLOAT);
        for (int frame = 0; frame < frames; frame++) {                              // the redraw handler (hello_mesh.rs:380-410)
            wgb::Texture frame_texture = surface.get_current_texture();
            wgb::CommandEncoder encoder = device.create_command_encoder();
            {
                wgb::DepthAttachment da{depth.create_view(), true, 1.0f};
                wgb::RenderPass pass = encoder.begin_render_pass({wgb::ColorAttachment{frame_texture.create_view(), true, {0.0, 0.0, 0.0, 1.0}}}, &da);
                pass.set_pipeline(pipeline);
                pass.set_bind_group(0, bind_group);
                pass.set_index_buffer(index_buffer, WGB_INDEX_FORMAT_UINT32);
                pass.set_vertex_buffer(0, vertex_buffer);
                pass.draw_indexed(0, (uint32_t)(indices.size() / 4));
            }
            queue.submit({encoder.finish()});
            surface.present();                                                     // waits for the frame, fills the window
        }
        if (frames > 0) {
            std::cout << "presented " << window.presents << " frames of " << window.width << "x" << window.height << "\n";
            write_file(out + ".window", window.pixels.data(), window.pixels.size());
            return 0;
        }
        wgb::Texture target = device.create_texture(width, height, WGB_TEXTURE_FORMAT_RGBA8_UNORM_SRGB);

        wgb::CommandEncoder encoder = device.create_command_encoder();
        {
            wgb::DepthAttachment da{depth.create_view(), true, 1.0f};
            wgb::RenderPass pass = encoder.begin_render_pass({wgb::ColorAttachment{target.create_view(), true, {0.0, 0.0, 0.0, 1.0}}}, &da);
            pass.set_pipeline(pipeline);
            pass.set_bind_group(0, bind_group);
            pass.set_index_buffer(index_buffer, WGB_INDEX_FORMAT_UINT32);
            pass.set_vertex_buffer(0, vertex_buffer);
            pass.draw_indexed(0, (uint32_t)(indices.size() / 4));
        }   // the pass ends when it is dropped (render_pass/mod.rs:325-329)
        const uint64_t submission = queue.submit({encoder.finish()});
        device.poll_wait(submission);

        const wgb_pass_stats st = device.last_pass_stats();
        std::cout << "primitives " << st.primitives << ", fragments " << st.fragments << ", shaded " << st.shaded << ", device "
                  << st.total_ms << " ms in " << st.kernel_launches << " kernel launches\n";
        target.dump_png(out + ".png");                                              // wgpu_cpu::dump_texture (lib.rs:111-158)
        const std::vector<uint8_t> rgba = target.read(), z = depth.read();
        write_file(out + ".rgba", rgba.data(), rgba.size());
        write_file(out + ".depth", z.data(), z.size());
        return 0;
    } catch (const wgb::Error& e) {
        std::fprintf(stderr, "wgpu-b200 error %d: %s\n", e.status, e.what());
        return 1;
    }
}
