"""Render a `scenes.Scene` through the wgpu-style host API, the way hello_mesh.rs drives wgpu
(wgpu-cpu/examples/hello_mesh.rs:111-157, 174-224, 263-339): create resources, one command
encoder, one render pass, submit, poll(Wait), read the attachments back."""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import api, shaders
from .scenes import Scene


@dataclass
class Frame:
    color: np.ndarray
    depth: Optional[np.ndarray]
    coverage: Optional[np.ndarray]
    stats: dict
    extra_colors: Optional[list] = None     # further colour attachments, in location order


class SceneRenderer:
    """Holds the device-resident resources of one scene so that frames can be re-submitted."""

    def __init__(self, device: api.Device, queue: api.Queue, scene: Scene, use_emitted: bool = False,
                 target: Optional[api.Texture] = None, targets: Optional[list] = None, wgsl: Optional[str] = None):
        """`target`: render into this texture instead of creating one (e.g. the presenter's colour target
        imported over CUDA IPC, so that the tile kernel stores its band straight into peer memory).
        `targets`: several colour targets to alternate between (`encode(which)`): consecutive frames of a multi-GPU run
        go to different presenter targets, so that a rank that is ahead never stores into the frame being read."""
        self.device, self.queue, self.scene = device, queue, scene
        s = scene
        if use_emitted:
            module = device.create_shader_module(None, [
                (api.STAGE_VERTEX, "vs_main", shaders.emitted(s.shader, "vs")),
                (api.STAGE_FRAGMENT, "fs_main", shaders.emitted(s.shader, "fs"))])
        elif wgsl is not None:          # the application's own WGSL text instead of this repository's text for the scene's program
            module = device.create_shader_module(wgsl)
        else:
            module = device.create_shader_module(shaders.wgsl(s.shader))
        self.module = module
        self.vertex_buffers = [device.create_buffer_init(vb, api.BUFFER_USAGE["VERTEX"]) for vb in s.vertex_buffers]
        self.index_buffer = None
        if s.index_data is not None:
            self.index_buffer = device.create_buffer_init(s.index_data, api.BUFFER_USAGE["INDEX"])
            self.index_format = "uint16" if s.index_data.dtype == np.uint16 else "uint32"
        self.resources = {}
        groups = {}
        for (g, b), res in sorted(s.bindings.items()):
            if res[0] == "buffer":
                buf = device.create_buffer_init(res[1], api.BUFFER_USAGE["UNIFORM"] | api.BUFFER_USAGE["COPY_DST"])
                self.resources[(g, b)] = buf
                groups.setdefault(g, []).append({"binding": b, "buffer": buf})
            elif res[0] == "texture":
                img = res[1]
                tex = device.create_texture_with_data(queue, img.shape[1], img.shape[0], res[2], img)
                view = tex.create_view()
                self.resources[(g, b)] = (tex, view)
                groups.setdefault(g, []).append({"binding": b, "texture_view": view})
            elif res[0] == "sampler":
                smp = device.create_sampler(address_mode_u=res[1], address_mode_v=res[2], address_mode_w=res[1])
                self.resources[(g, b)] = smp
                groups.setdefault(g, []).append({"binding": b, "sampler": smp})
        self.bind_groups = {}
        for g, entries in groups.items():
            layout = None
            if s.dynamic_bindings and g in s.dynamic_bindings:       # bindings that take a dynamic offset need a layout that says so
                layout = device.create_bind_group_layout(
                    [(e["binding"], 3, 1 if "buffer" in e else 2 if "texture_view" in e else 3,
                      1 if e["binding"] in s.dynamic_bindings[g] else 0) for e in entries])
            self.bind_groups[g] = device.create_bind_group(layout, entries)
        depth_state = None
        if s.depth_compare is not None:
            depth_state = {"format": "depth32float", "depth_write_enabled": s.depth_write, "depth_compare": s.depth_compare}
        self.pipeline = device.create_render_pipeline(
            vertex_module=module, fragment_module=module,
            vertex_buffers=[{"array_stride": l.stride, "step_mode": l.step_mode,
                             "attributes": [(a.format, a.offset, a.location) for a in l.attributes]} for l in s.vertex_layouts],
            topology=s.topology, strip_index_format=s.strip_index_format, front_face=s.front_face, cull_mode=s.cull_mode,
            depth_stencil=depth_state,
            targets=[s.color_format if s.color_write_mask == 15 and not s.blend
                     else {"format": s.color_format, "write_mask": s.color_write_mask, "blend": s.blend}] +
                    [fmt for fmt, _ in (s.extra_targets or [])],
            multisample_count=getattr(s, "multisample_count", 1))      # carried and never applied, like the reference's (pipeline.rs:98)
        self.extra_targets = [device.create_texture(s.width, s.height, fmt) for fmt, _ in (s.extra_targets or [])]
        self.extra_views = [t.create_view() for t in self.extra_targets]
        if targets:
            self.targets = list(targets)
        else:
            self.targets = [target if target is not None else device.create_texture(s.width, s.height, s.color_format)]
        self.target_views = [t.create_view() for t in self.targets]
        self.target, self.target_view = self.targets[0], self.target_views[0]
        self.depth_texture = self.depth_view = None
        if s.has_depth:
            self.depth_texture = device.create_texture(s.width, s.height, "depth32float")
            self.depth_view = self.depth_texture.create_view()
        if s.initial_color is not None:
            queue.write_texture(self.target, s.initial_color)
        if s.initial_depth is not None and self.depth_texture is not None:
            queue.write_texture(self.depth_texture, np.ascontiguousarray(s.initial_depth, dtype=np.float32))

    def encode(self, which: int = 0) -> api.CommandBuffer:
        s = self.scene
        enc = self.device.create_command_encoder()
        color = {"view": self.target_views[which % len(self.target_views)], "load": ("clear", s.clear_color) if s.clear_color is not None else "load"}
        depth = None
        if s.has_depth:
            depth = {"view": self.depth_view, "depth_load": ("clear", s.clear_depth) if s.clear_depth is not None else "load",
                     "depth_store": "discard"}
        colors = [color] + [{"view": v, "load": ("clear", clear) if clear is not None else "load"}
                            for v, (_, clear) in zip(self.extra_views, s.extra_targets or [])]
        with enc.begin_render_pass(colors, depth) as rp:
            self.record_into(rp)
        return enc.finish()

    def record_into(self, rp) -> None:
        """This scene's pipeline, bindings, state and draws into an open render pass (several scenes can share one)."""
        s = self.scene
        rp.set_pipeline(self.pipeline)
        for g, bg in self.bind_groups.items():
            rp.set_bind_group(g, bg, (s.dynamic_offsets or {}).get(g, ()))
        if self.index_buffer is not None:
            rp.set_index_buffer(self.index_buffer, self.index_format)
        for i, vb in enumerate(self.vertex_buffers):
            rp.set_vertex_buffer(i, vb)
        if s.blend:
            rp.set_blend_constant(s.blend_constant)
        if s.viewport is not None:
            rp.set_viewport(*s.viewport)
        if s.scissor is not None:
            rp.set_scissor_rect(*s.scissor)
        for d in s.draws:
            if d.indexed:
                rp.draw_indexed(range(d.first, d.first + d.count), d.base_vertex,
                                range(d.first_instance, d.first_instance + d.instance_count))
            else:
                rp.draw(range(d.first, d.first + d.count), range(d.first_instance, d.first_instance + d.instance_count))

    def submit(self, command_buffer: Optional[api.CommandBuffer] = None) -> int:
        idx = self.queue.submit([command_buffer if command_buffer is not None else self.encode()])
        return idx

    def render(self, command_buffer: Optional[api.CommandBuffer] = None) -> dict:
        """Submit (a command buffer recorded earlier with encode(), or a fresh one) and wait for it."""
        idx = self.submit(command_buffer)
        self.device.poll(True, idx)
        return self.device.last_pass_stats()

    def read(self, want_coverage: bool = False) -> Frame:
        s = self.scene
        color = self.target.read()
        depth = self.depth_texture.read() if self.depth_texture is not None else None
        cov = self.device.read_coverage(s.width, s.height) if want_coverage else None
        return Frame(color, depth, cov, self.device.last_pass_stats(), [t.read() for t in self.extra_targets])


def render_scene(device: api.Device, queue: api.Queue, scene: Scene, want_coverage: bool = True,
                 use_emitted: bool = False, wgsl: Optional[str] = None) -> Frame:
    device.set_coverage_capture(want_coverage)
    r = SceneRenderer(device, queue, scene, use_emitted=use_emitted, wgsl=wgsl)
    r.render()
    f = r.read(want_coverage)
    device.set_coverage_capture(False)
    return f
