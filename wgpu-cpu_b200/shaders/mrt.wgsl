// Parity-test program: one fragment stage, three colour attachments.  The outputs are declared out of location order;
// the draw path visits them in declaration order and the late depth test runs once, at the first of them
// (fragment.rs:457-488).

struct Camera {
    matrix: mat4x4f,
}
@group(0) @binding(0) var<uniform> camera: Camera;

struct Interstage {
    @builtin(position) clip: vec4f,
    @location(0) @interpolate(linear, sample) tint: vec4f,
}

struct Targets {
    @location(0) first: vec4f,
    @location(2) third: vec4f,
    @location(1) second: vec4f,
}

@vertex
fn vs_main(@location(0) object_position: vec4f, @location(1) tint: vec4f) -> Interstage {
    return Interstage(camera.matrix * object_position, tint);
}

@fragment
fn fs_main(frag: Interstage) -> Targets {
    let third = vec4f(frag.clip.z, frag.tint.x * frag.tint.y, 0.25, 1.0);
    let second = vec4f(frag.tint.z, frag.tint.y, frag.tint.x, 0.5);
    return Targets(frag.tint, third, second);
}
