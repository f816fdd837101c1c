// Parity-test program: a depth-only pass.  The fragment stage has no outputs, so the late depth test -- it runs at the
// first @location output (fragment.rs:457-488) -- never happens; @early_depth_test(force) makes the early test
// (fragment.rs:166-194) the one that tests and writes depth.

struct Camera {
    matrix: mat4x4f,
}
@group(0) @binding(0) var<uniform> camera: Camera;

@vertex
fn vs_main(@location(0) object_position: vec4f, @location(1) tint: vec4f) -> @builtin(position) vec4f {
    return camera.matrix * object_position;
}

@fragment
@early_depth_test(force)
fn fs_main() {
}
