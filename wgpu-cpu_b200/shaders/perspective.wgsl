// Parity-test program for what the reference leaves todo!() (fragment.rs:343-345): a varying with WGSL's default
// interpolation -- perspective-correct -- next to a linear and a flat one, so that the three can be told apart.

struct Camera {
    matrix: mat4x4f,
}
@group(0) @binding(0) var<uniform> camera: Camera;

struct Interstage {
    @builtin(position) clip: vec4f,
    @location(0) corrected: vec4f,
    @location(1) @interpolate(linear) screen_space: vec2f,
    @location(2) @interpolate(flat) provoking: f32,
}

@vertex
fn vs_main(@location(0) object_position: vec4f, @location(1) tint: vec4f) -> Interstage {
    return Interstage(camera.matrix * object_position, tint, vec2f(tint.x, tint.y), tint.z);
}

@fragment
fn fs_main(frag: Interstage) -> @location(0) vec4f {
    return vec4f(frag.corrected.x, frag.corrected.y * 0.5 + frag.screen_space.y * 0.5, frag.screen_space.x, frag.provoking);
}
