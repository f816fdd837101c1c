// The reference's examples/hello_shader.rs (wgpu-cpu/examples/hello_shader.rs:123-290) written against the C++ host
// (include/wgpu_b200.hpp): a pipeline without vertex buffers whose vertex stage derives position and colour from
// vertex_index (front face Ccw, no culling, Depth32Float Less + write), `--vertices` many of them drawn as a triangle
// list, colour and depth dumped.
//
//   hello_shader <shader.wgsl> <vertices> <width> <height> <out-prefix> [frames]
//
// With `frames`: the windowed flow of hello_shader.rs (:168-185, 330-350, 366-380) -- create_surface, get_capabilities,
// formats[0], configure, and per frame get_current_texture -> pass -> submit -> present -- against the library's headless
// surface; the last presented frame goes to <out-prefix>.window (raw BGRA texels).
//
// Writes <out-prefix>.png and <out-prefix>.depth.png (dump_texture of both attachments, hello_shader.rs --output /
// --output-depth), and the raw <out-prefix>.rgba / .depth that tests/test_cpp_host_gpu.py compares with the oracle.
#include "wgpu_b200.hpp"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>
#include <vector>

struct Window { std::vector<uint8_t> pixels; uint32_t presents = 0; };
static void on_present(void* user_data, const void* pixels, uint32_t, uint32_t height, uint32_t bytes_per_row) {
    Window* w = static_cast<Window*>(user_data);
    w->presents++;
    w->pixels.assign(static_cast<const uint8_t*>(pixels), static_cast<const uint8_t*>(pixels) + (size_t)bytes_per_row * height);
}

int main(int argc, char** argv) {
    if (argc != 6 && argc != 7) { std::fprintf(stderr, "usage: %s shader.wgsl vertices width height out-prefix\n", argv[0]); return 2; }
    try {
        std::ifstream f(argv[1], std::ios::binary);
        if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
        const std::string wgsl((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        const uint32_t vertices = (uint32_t)std::atoi(argv[2]), width = (uint32_t)std::atoi(argv[3]), height = (uint32_t)std::atoi(argv[4]);
        const std::string out = argv[5];
        const int frames = argc == 7 ? std::atoi(argv[6]) : 0;

        wgb::Instance instance = wgb::instance();
        wgb::Adapter adapter = instance.request_adapter();
        auto [device, queue] = adapter.request_device(0);
        Window window;
        wgb::Surface surface;
        uint32_t target_format = WGB_TEXTURE_FORMAT_RGBA8_UNORM_SRGB;
        if (frames > 0) {
            surface = instance.create_surface(on_present, &window);
            target_format = surface.get_capabilities(adapter.get()).formats[0];      // hello_shader.rs:172-173
            surface.configure(device, width, height, target_format);
        }
        wgb::ShaderModule shader = device.create_shader_module(wgsl);
        wgb::RenderPipelineDescriptor pd;
        pd.vertex_module = shader; pd.fragment_module = shader;
        pd.front_face = WGB_FRONT_FACE_CCW; pd.cull_mode = WGB_CULL_MODE_NONE;      // hello_shader.rs:134-139
        pd.has_depth_stencil = true;                                               // Depth32Float, Less, write (hello_shader.rs:143-149)
        pd.targets = {wgb::color_target(target_format)};
        wgb::RenderPipeline pipeline = device.create_render_pipeline(pd);
        wgb::Texture depth = device.create_texture(width, height, WGB_TEXTURE_FORMAT_DEPTH32_FLOAT);
        for (int frame = 0; frame < frames; frame++) {                              // the redraw handler (hello_shader.rs:330-350)
            wgb::Texture frame_texture = surface.get_current_texture();
            wgb::CommandEncoder encoder = device.create_command_encoder();
            {
                wgb::DepthAttachment da{depth.create_view(), true, 1.0f};
                wgb::RenderPass pass = encoder.begin_render_pass({wgb::ColorAttachment{frame_texture.create_view(), true, {0.0, 0.0, 0.0, 1.0}}}, &da);
                pass.set_pipeline(pipeline);
                pass.draw(0, vertices);
            }
            queue.submit({encoder.finish()});
            surface.present();
        }
        if (frames > 0) {
            std::cout << "presented " << window.presents << " frames\n";
            std::ofstream(out + ".window", std::ios::binary).write(reinterpret_cast<const char*>(window.pixels.data()), (std::streamsize)window.pixels.size());
            return 0;
        }
        wgb::Texture target = device.create_texture(width, height, WGB_TEXTURE_FORMAT_RGBA8_UNORM_SRGB);

        wgb::CommandEncoder encoder = device.create_command_encoder();
        {
            wgb::DepthAttachment da{depth.create_view(), true, 1.0f};
            wgb::RenderPass pass = encoder.begin_render_pass({wgb::ColorAttachment{target.create_view(), true, {0.0, 0.0, 0.0, 1.0}}}, &da);
            pass.set_pipeline(pipeline);
            pass.draw(0, vertices);
        }
        device.poll_wait(queue.submit({encoder.finish()}));
        std::cout << "primitives " << device.last_pass_stats().primitives << "\n";
        target.dump_png(out + ".png");
        depth.dump_png(out + ".depth.png");
        const std::vector<uint8_t> rgba = target.read(), z = depth.read();
        std::ofstream(out + ".rgba", std::ios::binary).write(reinterpret_cast<const char*>(rgba.data()), (std::streamsize)rgba.size());
        std::ofstream(out + ".depth", std::ios::binary).write(reinterpret_cast<const char*>(z.data()), (std::streamsize)z.size());
        return 0;
    } catch (const wgb::Error& e) {
        std::fprintf(stderr, "wgpu-b200 error %d: %s\n", e.status, e.what());
        return 1;
    }
}
