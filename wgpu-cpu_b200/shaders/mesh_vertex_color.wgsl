// Scene program of configs C1 (teapot) and C3 (synthetic mesh): the clip position is the camera matrix times the
// object-space position, and the per-vertex colour rides along to the fragment stage unchanged.  It computes what
// the reference's hello_mesh example shader computes (one mat4x4 * vec4 and a pass-through); the text is this
// repository's own.

struct Camera {
    matrix: mat4x4f,
}
@group(0) @binding(0) var<uniform> camera: Camera;

struct Interstage {
    @builtin(position) clip: vec4f,
    @location(0) @interpolate(linear, sample) tint: vec4f,
}

@vertex
fn vs_main(@location(0) object_position: vec4f, @location(1) tint: vec4f) -> Interstage {
    return Interstage(camera.matrix * object_position, tint);
}

@fragment
fn fs_main(frag: Interstage) -> @location(0) vec4f {
    return frag.tint;
}
