// Config C4: full-screen ALU-heavy fragment shader, 64 fixed iterations, no early exit.
// Restricted to the operations wgpu-cpu's shader JIT implements (SURVEY.md 2.3): scalar
// + - * /, comparisons, select, casts, `for`, member access.  No math builtins, no
// swizzles, no splats.  12 flops per iteration (4 + 3 + 3 + compare + accumulate).

struct VertexInput {
    @builtin(vertex_index)
    vertex_index: u32,
}

struct VertexOutput {
    @builtin(position)
    position: vec4f,
}

@vertex
fn vs_main(input: VertexInput) -> VertexOutput {
    let vi = input.vertex_index % 6u;

    // two triangles covering NDC [-1,1]^2: (0,0) (1,0) (0,1) / (0,1) (1,0) (1,1)
    let ux = f32(vi == 1u || vi == 4u || vi == 5u);
    let uy = f32(vi == 2u || vi == 3u || vi == 5u);
    let position = vec4f(ux * 2.0 - 1.0, uy * 2.0 - 1.0, 0.0, 1.0);

    return VertexOutput(position);
}

@fragment
fn fs_main(input: VertexOutput) -> @location(0) vec4f {
    let cx = input.position.x / 2560.0 - 2.0;
    let cy = input.position.y / 2160.0 - 1.0;

    var zx = 0.0;
    var zy = 0.0;
    var acc = 0.0;

    for (var i = 0; i < 64; i++) {
        let nx = zx * zx - zy * zy + cx;
        let ny = 2.0 * zx * zy + cy;
        let r2 = nx * nx + ny * ny;
        let escaped = r2 > 4.0;
        zx = select(nx, zx, escaped);
        zy = select(ny, zy, escaped);
        acc = acc + select(1.0, 0.0, escaped);
    }

    let t = acc / 64.0;
    return vec4f(t, t * t, 1.0 - t, 1.0);
}
