// The reference's examples/hello_shader.rs (wgpu-cpu/examples/hello_shader.rs:123-290) written against the C++ host
// (include/wgpu_b200.hpp): a pipeline without vertex buffers whose vertex stage derives position and colour from
// vertex_index (front face Ccw, no culling, Depth32Float Less + write), `--vertices` many of them drawn as a triangle
// list, colour and depth dumped.
//
//   hello_shader <shader.wgsl> <vertices> <width> <height> <out-prefix>
//
// Writes <out-prefix>.png and <out-prefix>.depth.png (dump_texture of both attachments, hello_shader.rs --output /
// --output-depth), and the raw <out-prefix>.rgba / .depth that tests/test_cpp_host_gpu.py compares with the oracle.
#include "wgpu_b200.hpp"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>

int main(int argc, char** argv) {
    if (argc != 6) { std::fprintf(stderr, "usage: %s shader.wgsl vertices width height out-prefix\n", argv[0]); return 2; }
    try {
        std::ifstream f(argv[1], std::ios::binary);
        if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
        const std::string wgsl((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        const uint32_t vertices = (uint32_t)std::atoi(argv[2]), width = (uint32_t)std::atoi(argv[3]), height = (uint32_t)std::atoi(argv[4]);
        const std::string out = argv[5];

        auto [device, queue] = wgb::instance().request_adapter().request_device(0);
        wgb::ShaderModule shader = device.create_shader_module(wgsl);
        wgb::RenderPipelineDescriptor pd;
        pd.vertex_module = shader; pd.fragment_module = shader;
        pd.front_face = WGB_FRONT_FACE_CCW; pd.cull_mode = WGB_CULL_MODE_NONE;      // hello_shader.rs:134-139
        pd.has_depth_stencil = true;                                               // Depth32Float, Less, write (hello_shader.rs:143-149)
        pd.targets = {wgb::color_target(WGB_TEXTURE_FORMAT_RGBA8_UNORM_SRGB)};
        wgb::RenderPipeline pipeline = device.create_render_pipeline(pd);
        wgb::Texture target = device.create_texture(width, height, WGB_TEXTURE_FORMAT_RGBA8_UNORM_SRGB);
        wgb::Texture depth = device.create_texture(width, height, WGB_TEXTURE_FORMAT_DEPTH32_FLOAT);

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
