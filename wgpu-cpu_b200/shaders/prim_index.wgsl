// Parity-test program: the fragment-stage builtins the example programs do not read.  primitive_index is the position of
// the primitive in its instance's assembly order, counted before culling and clipping (state.rs:541); the draw path has
// no multisampling, so sample_index is 0 and sample_mask all ones (fragment.rs:127-165).

struct Camera {
    matrix: mat4x4f,
}
@group(0) @binding(0) var<uniform> camera: Camera;

@vertex
fn vs_main(@location(0) object_position: vec4f, @location(1) tint: vec4f) -> @builtin(position) vec4f {
    return camera.matrix * object_position;
}

@fragment
fn fs_main(@builtin(primitive_index) primitive: u32, @builtin(sample_index) sample: u32, @builtin(sample_mask) mask: u32) -> @location(0) vec4f {
    let low = f32(primitive & 255u) / 255.0;
    let high = f32((primitive >> 8u) & 255u) / 255.0;
    return vec4f(low, high, f32(sample) + f32(mask & 1u) * 0.5, 1.0);
}
