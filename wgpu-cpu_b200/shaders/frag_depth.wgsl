// Parity-test program: the fragment stage overrides the rasteriser's depth through @builtin(frag_depth), which is
// declared ahead of the colour output so that the late depth test -- it runs at the first @location output -- uses it.

struct Camera {
    matrix: mat4x4f,
}
@group(0) @binding(0) var<uniform> camera: Camera;

struct Interstage {
    @builtin(position) clip: vec4f,
    @location(0) @interpolate(linear, sample) tint: vec4f,
}

struct Shaded {
    @builtin(frag_depth) depth: f32,
    @location(0) tint: vec4f,
}

@vertex
fn vs_main(@location(0) object_position: vec4f, @location(1) tint: vec4f) -> Interstage {
    return Interstage(camera.matrix * object_position, tint);
}

@fragment
fn fs_main(frag: Interstage) -> Shaded {
    return Shaded(1.0 - frag.clip.z * frag.tint.x, frag.tint);
}
