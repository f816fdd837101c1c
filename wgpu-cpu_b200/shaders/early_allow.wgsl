// Parity-test program: @early_depth_test(less_equal) -- naga's EarlyDepthTest::Allow -- in front of a fragment stage that discards and writes frag_depth.
// The depth test -- and the depth write -- happen before the stage runs, with the rasteriser's depth
// (fragment.rs:166-194): a discarded fragment has already left its depth behind, and the frag_depth it returns
// is then tested once more by the late test, against the depth the early test has just stored.

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
@early_depth_test(less_equal)
fn fs_main(frag: Interstage) -> Shaded {
    if (frag.tint.y > 0.6) {
        discard;
    }
    return Shaded(frag.clip.z * 0.5, frag.tint);
}
