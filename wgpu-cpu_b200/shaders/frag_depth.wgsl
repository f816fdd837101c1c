// Parity-test shader: the fragment stage replaces the rasteriser's depth through
// @builtin(frag_depth), declared before the colour output so that the late depth test
// (which runs at the first @location output) sees it.

struct VertexInput {
    @location(0)
    vertex_position: vec4f,
    @location(1)
    vertex_color: vec4f,
}

struct VertexOutput {
    @builtin(position)
    position: vec4f,

    @location(0)
    @interpolate(linear, sample)
    color: vec4f,
}

struct Camera {
    matrix: mat4x4f,
}

@group(0)
@binding(0)
var<uniform> camera: Camera;

@vertex
fn vs_main(input: VertexInput) -> VertexOutput {
    let position = camera.matrix * input.vertex_position;
    return VertexOutput(position, input.vertex_color);
}

struct FragmentOutput {
    @builtin(frag_depth)
    depth: f32,

    @location(0)
    color: vec4f,
}

@fragment
fn fs_main(input: VertexOutput) -> FragmentOutput {
    let depth = 1.0 - input.position.z * input.color.x;
    return FragmentOutput(depth, input.color);
}
