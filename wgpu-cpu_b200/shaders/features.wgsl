// Parity-test shader exercising the parts of the draw path the example shaders do not:
// instance_index, a flat integer varying, front_facing and discard.  Same operation
// subset as procedural.wgsl.

struct VertexInput {
    @builtin(vertex_index)
    vertex_index: u32,
    @builtin(instance_index)
    instance_index: u32,

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

    @location(1)
    @interpolate(flat)
    tag: u32,
}

struct Params {
    matrix: mat4x4f,
    instance_offset: vec4f,
}

@group(0)
@binding(0)
var<uniform> params: Params;

@vertex
fn vs_main(input: VertexInput) -> VertexOutput {
    let shift = f32(input.instance_index);
    let p = params.matrix * input.vertex_position;
    let position = vec4f(
        p.x + params.instance_offset.x * shift,
        p.y + params.instance_offset.y * shift,
        p.z + params.instance_offset.z * shift,
        p.w,
    );

    return VertexOutput(
        position,
        input.vertex_color,
        input.vertex_index + input.instance_index * 1000u,
    );
}

struct FragmentInput {
    @builtin(position)
    position: vec4f,
    @builtin(front_facing)
    front_facing: bool,

    @location(0)
    @interpolate(linear, sample)
    color: vec4f,

    @location(1)
    @interpolate(flat)
    tag: u32,
}

@fragment
fn fs_main(input: FragmentInput) -> @location(0) vec4f {
    // the interpolated position is extrapolated for fragments outside their triangle (the
    // reference's scan-line coverage has no inside test), so bound it before the trapping cast
    let lx = select(input.position.x, 0.0, input.position.x < 0.0);
    let ly = select(input.position.y, 0.0, input.position.y < 0.0);
    let px = u32(select(lx, 4096.0, lx > 4096.0));
    let py = u32(select(ly, 4096.0, ly > 4096.0));
    if ((px / 4u + py / 4u) % 3u == 0u) {
        discard;
    }

    let t = f32(input.tag % 7u) / 7.0;
    let facing = f32(input.front_facing);
    return vec4f(input.color.x, input.color.y * t, input.color.z * facing, 1.0);
}
