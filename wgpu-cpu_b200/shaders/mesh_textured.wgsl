// Scene program of configs C2 and C5 (textured bunny): camera matrix times object-space position, texture
// coordinates passed through, one nearest-neighbour textureSample in the fragment stage.  It computes what the
// reference's hello_texture example shader computes; the text is this repository's own.

struct Camera {
    matrix: mat4x4f,
}
@group(0) @binding(0) var<uniform> camera: Camera;
@group(1) @binding(0) var albedo: texture_2d<f32>;
@group(1) @binding(1) var albedo_sampler: sampler;

struct Interstage {
    @builtin(position) clip: vec4f,
    @location(0) @interpolate(linear, sample) uv: vec2f,
}

@vertex
fn vs_main(@location(0) object_position: vec4f, @location(1) uv: vec2f) -> Interstage {
    return Interstage(camera.matrix * object_position, uv);
}

@fragment
fn fs_main(frag: Interstage) -> @location(0) vec4f {
    return textureSample(albedo, albedo_sampler, frag.uv);
}
