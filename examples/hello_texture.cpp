// The reference's examples/hello_texture.rs (wgpu-cpu/examples/hello_texture.rs:180-240, 647-677) written against the C++
// host (include/wgpu_b200.hpp): the hello_mesh flow plus a texture created from image bytes (create_texture_with_data ->
// Queue::write_texture, hello_texture.rs:184-206), a Repeat / Nearest sampler (hello_texture.rs:209-217) and a second
// bind group {texture view, sampler} for the fragment stage's textureSample.
//
//   hello_texture <shader.wgsl> <vertices.bin> <indices.bin> <uniform.bin> <texture.rgba> <tex-width> <tex-height> <width> <height> <out-prefix> [frames]
//
// With `frames`: the windowed flow of hello_texture.rs (:328-345, the redraw handler) against the library's headless surface
// -- create_surface, formats[0], configure, and per frame get_current_texture -> pass -> submit -> present; the last frame
// the window received goes to <out-prefix>.window (raw BGRA texels).
//
// vertices: {pos vec4f, uv vec2f} (24 bytes, hello_texture.rs:651-659); indices: u32; uniform: the 64-byte camera matrix;
// texture: tex-width x tex-height RGBA8 texels.  Writes <out-prefix>.png, .rgba and .depth like hello_mesh.
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

struct Window { std::vector<uint8_t> pixels; uint32_t presents = 0; };
static void on_present(void* user_data, const void* pixels, uint32_t, uint32_t height, uint32_t bytes_per_row) {
    Window* w = static_cast<Window*>(user_data);
    w->presents++;
    w->pixels.assign(static_cast<const uint8_t*>(pixels), static_cast<const uint8_t*>(pixels) + (size_t)bytes_per_row * height);
}

int main(int argc, char** argv) {
    if (argc != 11 && argc != 12) {
        std::fprintf(stderr, "usage: %s shader.wgsl vertices.bin indices.bin uniform.bin texture.rgba tex-width tex-height width height out-prefix\n", argv[0]);
        return 2;
    }
    try {
        const std::vector<char> wgsl = read_file(argv[1]), vertices = read_file(argv[2]), indices = read_file(argv[3]), uniform = read_file(argv[4]),
                                texels = read_file(argv[5]);
        const uint32_t tex_w = (uint32_t)std::atoi(argv[6]), tex_h = (uint32_t)std::atoi(argv[7]);
        const uint32_t width = (uint32_t)std::atoi(argv[8]), height = (uint32_t)std::atoi(argv[9]);
        const std::string out = argv[10];
        const int frames = argc == 12 ? std::atoi(argv[11]) : 0;
        if (texels.size() != (size_t)tex_w * tex_h * 4) { std::fprintf(stderr, "texture file does not hold %u x %u RGBA8 texels\n", tex_w, tex_h); return 2; }

        wgb::Instance instance = wgb::instance();
        wgb::Adapter adapter = instance.request_adapter();
        auto [device, queue] = adapter.request_device(0);

        wgb::ShaderModule shader = device.create_shader_module(std::string(wgsl.begin(), wgsl.end()));
        wgb::Buffer vertex_buffer = device.create_buffer_init(vertices.data(), vertices.size(), WGB_BUFFER_USAGE_VERTEX);
        wgb::Buffer index_buffer = device.create_buffer_init(indices.data(), indices.size(), WGB_BUFFER_USAGE_INDEX);
        wgb::Buffer camera_buffer = device.create_buffer_init(uniform.data(), uniform.size(), WGB_BUFFER_USAGE_UNIFORM | WGB_BUFFER_USAGE_COPY_DST);

        // create_texture_with_data (hello_texture.rs:184-206): create, then Queue::write_texture with the image bytes
        wgb::Texture image = device.create_texture(tex_w, tex_h, WGB_TEXTURE_FORMAT_RGBA8_UNORM_SRGB);
        queue.write_texture(image, texels.data(), texels.size(), tex_w * 4);
        wgb::TextureView image_view = image.create_view();
        wgb::Sampler sampler = device.create_sampler(WGB_ADDRESS_MODE_REPEAT, WGB_ADDRESS_MODE_REPEAT);

        wgb_bind_group_entry camera_entry{};
        camera_entry.binding = 0; camera_entry.kind = WGB_BINDING_BUFFER; camera_entry.buffer = camera_buffer.get();
        camera_entry.offset = 0; camera_entry.size = WGB_WHOLE_SIZE;
        wgb::BindGroup camera_group = device.create_bind_group({camera_entry});
        wgb_bind_group_entry view_entry{}, sampler_entry{};
        view_entry.binding = 0; view_entry.kind = WGB_BINDING_TEXTURE_VIEW; view_entry.texture_view = image_view.get(); view_entry.size = WGB_WHOLE_SIZE;
        sampler_entry.binding = 1; sampler_entry.kind = WGB_BINDING_SAMPLER; sampler_entry.sampler = sampler.get(); sampler_entry.size = WGB_WHOLE_SIZE;
        wgb::BindGroup texture_group = device.create_bind_group({view_entry, sampler_entry});

        wgb::RenderPipelineDescriptor pd;
        pd.vertex_module = shader; pd.fragment_module = shader;
        wgb::VertexBufferLayout layout;
        layout.array_stride = 24;
        layout.attributes = {{WGB_VERTEX_FORMAT_FLOAT32X4, 0, 0}, {WGB_VERTEX_FORMAT_FLOAT32X2, 16, 1}};
        pd.vertex_buffers = {layout};
        pd.front_face = WGB_FRONT_FACE_CW; pd.cull_mode = WGB_CULL_MODE_BACK;
        pd.has_depth_stencil = true;
        Window window;
        wgb::Surface surface;
        uint32_t target_format = WGB_TEXTURE_FORMAT_RGBA8_UNORM_SRGB;
        if (frames > 0) {
            surface = instance.create_surface(on_present, &window);
            target_format = surface.get_capabilities(adapter.get()).formats[0];      // hello_texture.rs:332-333
            surface.configure(device, width, height, target_format);
        }
        pd.targets = {wgb::color_target(target_format)};
        wgb::RenderPipeline pipeline = device.create_render_pipeline(pd);

        wgb::Texture depth = device.create_texture(width, height, WGB_TEXTURE_FORMAT_DEPTH32_FLOAT);
        for (int frame = 0; frame < frames; frame++) {
            wgb::Texture frame_texture = surface.get_current_texture();
            wgb::CommandEncoder encoder = device.create_command_encoder();
            {
                wgb::DepthAttachment da{depth.create_view(), true, 1.0f};
                wgb::RenderPass pass = encoder.begin_render_pass({wgb::ColorAttachment{frame_texture.create_view(), true, {0.0, 0.0, 0.0, 1.0}}}, &da);
                pass.set_pipeline(pipeline);
                pass.set_bind_group(0, camera_group);
                pass.set_bind_group(1, texture_group);
                pass.set_index_buffer(index_buffer, WGB_INDEX_FORMAT_UINT32);
                pass.set_vertex_buffer(0, vertex_buffer);
                pass.draw_indexed(0, (uint32_t)(indices.size() / 4));
            }
            queue.submit({encoder.finish()});
            surface.present();
        }
        if (frames > 0) {
            std::cout << "presented " << window.presents << " frames\n";
            write_file(out + ".window", window.pixels.data(), window.pixels.size());
            return 0;
        }
        wgb::Texture target = device.create_texture(width, height, WGB_TEXTURE_FORMAT_RGBA8_UNORM_SRGB);

        wgb::CommandEncoder encoder = device.create_command_encoder();
        {
            wgb::DepthAttachment da{depth.create_view(), true, 1.0f};
            wgb::RenderPass pass = encoder.begin_render_pass({wgb::ColorAttachment{target.create_view(), true, {0.0, 0.0, 0.0, 1.0}}}, &da);
            pass.set_pipeline(pipeline);
            pass.set_bind_group(0, camera_group);
            pass.set_bind_group(1, texture_group);
            pass.set_index_buffer(index_buffer, WGB_INDEX_FORMAT_UINT32);
            pass.set_vertex_buffer(0, vertex_buffer);
            pass.draw_indexed(0, (uint32_t)(indices.size() / 4));
        }
        device.poll_wait(queue.submit({encoder.finish()}));

        const wgb_pass_stats st = device.last_pass_stats();
        std::cout << "primitives " << st.primitives << ", fragments " << st.fragments << ", shaded " << st.shaded << "\n";
        target.dump_png(out + ".png");
        const std::vector<uint8_t> rgba = target.read(), z = depth.read();
        write_file(out + ".rgba", rgba.data(), rgba.size());
        write_file(out + ".depth", z.data(), z.size());
        return 0;
    } catch (const wgb::Error& e) {
        std::fprintf(stderr, "wgpu-b200 error %d: %s\n", e.status, e.what());
        return 1;
    }
}
