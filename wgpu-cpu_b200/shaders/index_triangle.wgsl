// Three corners derived from the vertex index alone (no vertex buffer): corner k sits at (k - 1, -1 or +1) and
// carries one primary colour.  It computes what the reference's colored_triangle test shader / hello_shader
// example compute (integer arithmetic on the index, then casts); the text is this repository's own.

struct Corner {
    @builtin(position) clip: vec4f,
    @location(0) @interpolate(linear, sample) tint: vec4f,
}

@vertex
fn vs_main(@builtin(vertex_index) index: u32) -> Corner {
    let k = index % 3;
    let clip = vec4f(f32(i32(k) - 1), f32(i32(k & 1u) * 2 - 1), 0.0, 1.0);
    let tint = vec4f(f32(k == 0), f32(k == 1), f32(k == 2), 1.0);
    return Corner(clip, tint);
}

@fragment
fn fs_main(corner: Corner) -> @location(0) vec4f {
    return corner.tint;
}
