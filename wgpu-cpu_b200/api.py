"""Host-side mirror of the wgpu surface over the C ABI (include/wgpu_b200.h).

Rust is not available in this image, so the host above the C ABI is Python (ctypes): the classes
and methods follow wgpu's names (Instance.request_adapter, Adapter.request_device,
Device.create_buffer / create_texture / create_shader_module / create_render_pipeline /
create_command_encoder, CommandEncoder.begin_render_pass, RenderPass.set_pipeline / draw_indexed,
Queue.submit / write_buffer / write_texture, Device.poll ...), with the argument meaning and
error behaviour of wgpu-cpu's implementation of them.  Every method is one call
into libwgpu_b200.so; no arithmetic of the draw path lives here and there is no CPU fallback --
if the library or a CUDA device is missing, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwgpu_b200.so")

TOPOLOGY = {"point-list": 0, "line-list": 1, "line-strip": 2, "triangle-list": 3, "triangle-strip": 4}
INDEX_FORMAT = {None: 0, "uint16": 1, "uint32": 2}
FRONT_FACE = {"ccw": 0, "cw": 1}
CULL_MODE = {None: 0, "front": 1, "back": 2}
COMPARE = {"never": 1, "less": 2, "equal": 3, "less-equal": 4, "greater": 5, "not-equal": 6,
           "greater-equal": 7, "always": 8}
TEXTURE_FORMAT = {"rgba8unorm": 0, "rgba8unorm-srgb": 1, "bgra8unorm": 2, "bgra8unorm-srgb": 3, "r8unorm": 4,
                  "rg8unorm": 5, "rgba8snorm": 6, "depth32float": 7}
BYTES_PER_TEXEL = {0: 4, 1: 4, 2: 4, 3: 4, 4: 1, 5: 2, 6: 4, 7: 4}
ADDRESS_MODE = {"clamp-to-edge": 0, "repeat": 1, "mirror-repeat": 2, "clamp-to-border": 3}
FILTER_MODE = {"nearest": 0, "linear": 1}
STEP_MODE = {"vertex": 0, "instance": 1}
VERTEX_FORMAT = {"float32": 0, "float32x2": 1, "float32x3": 2, "float32x4": 3, "uint32": 4, "sint32": 5,
                 "uint32x2": 6, "uint32x3": 7, "uint32x4": 8, "sint32x2": 9, "sint32x3": 10, "sint32x4": 11}
STAGE_VERTEX, STAGE_FRAGMENT = 1, 2
BUFFER_USAGE = {"MAP_READ": 1, "MAP_WRITE": 2, "COPY_SRC": 4, "COPY_DST": 8, "INDEX": 16, "VERTEX": 32, "UNIFORM": 64,
                "STORAGE": 128}
WHOLE_SIZE = 0xFFFFFFFFFFFFFFFF
CUDA_DEVICE_CURRENT = -1
CUDA_DEVICE_COMPILE_ONLY = -2   # shader translation + NVRTC only (host-logic tests on machines without a GPU)

STATUS_NAMES = {1: "validation", 2: "unsupported", 3: "out of memory", 4: "device", 5: "shader", 6: "out of bounds",
                7: "timeout"}


class WgpuError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"wgpu-b200 {STATUS_NAMES.get(status, status)} error: {message}")
        self.status = status


# ---- ctypes mirrors of the descriptor structs ----
class _DeviceDescriptor(C.Structure):
    _fields_ = [("cuda_device", C.c_int32), ("band_rank", C.c_uint32), ("band_count", C.c_uint32), ("features", C.c_uint32)]


FEATURE = {"VIEWPORT_DEPTH_RANGE": 1, "COLOR_WRITE_MASK": 2, "SRGB_ENCODE": 4, "DYNAMIC_OFFSETS": 8, "BLEND": 16}
COLOR_WRITE = {"RED": 1, "GREEN": 2, "BLUE": 4, "ALPHA": 8, "ALL": 15}


class _AdapterInfo(C.Structure):
    _fields_ = [("name", C.c_char * 128), ("device_type", C.c_uint32), ("cuda_device_count", C.c_uint32)]


class _BufferDescriptor(C.Structure):
    _fields_ = [("size", C.c_uint64), ("usage", C.c_uint32), ("mapped_at_creation", C.c_uint32)]


class _TextureDescriptor(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("depth_or_array_layers", C.c_uint32),
                ("mip_level_count", C.c_uint32), ("sample_count", C.c_uint32), ("format", C.c_uint32),
                ("usage", C.c_uint32)]


class _SamplerDescriptor(C.Structure):
    _fields_ = [("address_mode_u", C.c_uint32), ("address_mode_v", C.c_uint32), ("address_mode_w", C.c_uint32),
                ("mag_filter", C.c_uint32), ("min_filter", C.c_uint32), ("mipmap_filter", C.c_uint32)]


class _EmittedEntryPoint(C.Structure):
    _fields_ = [("stage", C.c_uint32), ("entry_point", C.c_char_p), ("cuda_source", C.c_char_p)]


class _ShaderModuleDescriptor(C.Structure):
    _fields_ = [("wgsl", C.c_char_p), ("emitted_count", C.c_uint32), ("emitted", C.POINTER(_EmittedEntryPoint))]


class _BindGroupLayoutEntry(C.Structure):
    _fields_ = [("binding", C.c_uint32), ("visibility", C.c_uint32), ("kind", C.c_uint32), ("has_dynamic_offset", C.c_uint32)]


class _BindGroupEntry(C.Structure):
    _fields_ = [("binding", C.c_uint32), ("kind", C.c_uint32), ("buffer", C.c_void_p), ("offset", C.c_uint64),
                ("size", C.c_uint64), ("texture_view", C.c_void_p), ("sampler", C.c_void_p)]


class _VertexAttribute(C.Structure):
    _fields_ = [("format", C.c_uint32), ("offset", C.c_uint64), ("shader_location", C.c_uint32)]


class _VertexBufferLayout(C.Structure):
    _fields_ = [("array_stride", C.c_uint64), ("step_mode", C.c_uint32), ("attribute_count", C.c_uint32),
                ("attributes", C.POINTER(_VertexAttribute))]


class _BlendComponent(C.Structure):
    _fields_ = [("src_factor", C.c_uint32), ("dst_factor", C.c_uint32), ("operation", C.c_uint32)]


class _ColorTargetState(C.Structure):
    _fields_ = [("format", C.c_uint32), ("has_blend", C.c_uint32), ("write_mask", C.c_uint32),
                ("blend_color", _BlendComponent), ("blend_alpha", _BlendComponent)]


BLEND_FACTOR = {"zero": 0, "one": 1, "src": 2, "one-minus-src": 3, "src-alpha": 4, "one-minus-src-alpha": 5, "dst": 6,
                "one-minus-dst": 7, "dst-alpha": 8, "one-minus-dst-alpha": 9, "src-alpha-saturated": 10, "constant": 11,
                "one-minus-constant": 12}
BLEND_OPERATION = {"add": 0, "subtract": 1, "reverse-subtract": 2, "min": 3, "max": 4}


def _blend_component(c):
    src, dst, op = c
    return _BlendComponent(BLEND_FACTOR[src], BLEND_FACTOR[dst], BLEND_OPERATION[op])


class _RenderPipelineDescriptor(C.Structure):
    _fields_ = [("layout", C.c_void_p), ("vertex_module", C.c_void_p), ("vertex_entry_point", C.c_char_p),
                ("vertex_buffer_count", C.c_uint32), ("vertex_buffers", C.POINTER(_VertexBufferLayout)),
                ("topology", C.c_uint32), ("strip_index_format", C.c_uint32), ("front_face", C.c_uint32),
                ("cull_mode", C.c_uint32), ("polygon_mode", C.c_uint32), ("unclipped_depth", C.c_uint32),
                ("conservative", C.c_uint32), ("has_depth_stencil", C.c_uint32), ("depth_format", C.c_uint32),
                ("depth_write_enabled", C.c_uint32), ("depth_compare", C.c_uint32), ("multisample_count", C.c_uint32),
                ("fragment_module", C.c_void_p), ("fragment_entry_point", C.c_char_p), ("target_count", C.c_uint32),
                ("targets", C.POINTER(_ColorTargetState))]


class _TexelCopyTextureInfo(C.Structure):
    _fields_ = [("texture", C.c_void_p), ("x", C.c_uint32), ("y", C.c_uint32), ("layer", C.c_uint32)]


class _TexelCopyBufferInfo(C.Structure):
    _fields_ = [("buffer", C.c_void_p), ("offset", C.c_uint64), ("bytes_per_row", C.c_uint32), ("rows_per_image", C.c_uint32)]


class _ColorAttachment(C.Structure):
    _fields_ = [("view", C.c_void_p), ("load_op", C.c_uint32), ("store_op", C.c_uint32), ("clear_value", C.c_double * 4)]


class _DepthStencilAttachment(C.Structure):
    _fields_ = [("view", C.c_void_p), ("has_depth_ops", C.c_uint32), ("depth_load_op", C.c_uint32),
                ("depth_store_op", C.c_uint32), ("depth_clear_value", C.c_float), ("has_stencil_ops", C.c_uint32)]


class _RenderPassDescriptor(C.Structure):
    _fields_ = [("color_attachment_count", C.c_uint32), ("color_attachments", C.POINTER(_ColorAttachment)),
                ("depth_stencil_attachment", C.POINTER(_DepthStencilAttachment))]


_PRESENT_CALLBACK = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32)


class _SurfaceTarget(C.Structure):
    _fields_ = [("on_present", _PRESENT_CALLBACK), ("user_data", C.c_void_p)]


class _SurfaceCapabilities(C.Structure):
    _fields_ = [("format_count", C.c_uint32), ("formats", C.c_uint32 * 4),
                ("present_mode_count", C.c_uint32), ("present_modes", C.c_uint32 * 4),
                ("alpha_mode_count", C.c_uint32), ("alpha_modes", C.c_uint32 * 4), ("usages", C.c_uint32)]


class _SurfaceConfiguration(C.Structure):
    _fields_ = [("usage", C.c_uint32), ("format", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32),
                ("present_mode", C.c_uint32), ("alpha_mode", C.c_uint32),
                ("view_format_count", C.c_uint32), ("view_formats", C.POINTER(C.c_uint32))]


TEXTURE_USAGE = {"copy-src": 1, "copy-dst": 2, "texture-binding": 4, "storage-binding": 8, "render-attachment": 16}
PRESENT_MODE = {"auto-vsync": 0, "auto-no-vsync": 1, "fifo": 2, "fifo-relaxed": 3, "immediate": 4, "mailbox": 5}
COMPOSITE_ALPHA_MODE = {"auto": 0, "opaque": 1}


class PassStats(C.Structure):
    _fields_ = [("primitives", C.c_uint64), ("fragments", C.c_uint64), ("shaded", C.c_uint64), ("bin_pairs", C.c_uint64),
                ("big_primitives", C.c_uint64), ("clipped_primitives", C.c_uint64), ("clip_records", C.c_uint64),
                ("draws", C.c_uint32), ("kernel_launches", C.c_uint32), ("geometry_ms", C.c_float), ("tile_ms", C.c_float),
                ("total_ms", C.c_float), ("replays", C.c_uint32), ("hiz_culled", C.c_uint64)]

    def as_dict(self):
        return {f[0]: getattr(self, f[0]) for f in self._fields_}


_lib = None


def load_library():
    """Load libwgpu_b200.so; raises if it has not been built (there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` first; "
                           "the B200 backend has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.wgb_last_error.restype = C.c_char_p
    lib.wgb_version.restype = C.c_char_p
    lib.wgb_release.argtypes = [C.c_void_p]
    lib.wgb_retain.argtypes = [C.c_void_p]
    lib.wgb_free.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _check(status: int):
    if status != 0:
        raise WgpuError(status, load_library().wgb_last_error().decode(errors="replace"))


class _Handle:
    def __init__(self, handle):
        self._h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle

    def __del__(self):
        try:
            if self._h and _lib is not None:
                _lib.wgb_release(self._h)
                self._h = None
        except Exception:
            pass


def instance() -> "Instance":
    """wgpu_cpu::instance(Config) (wgpu-cpu/src/lib.rs:22-27)."""
    lib = load_library()
    h = C.c_void_p()
    _check(lib.wgb_create_instance(None, C.byref(h)))
    return Instance(h)


class Instance(_Handle):
    def request_adapter(self) -> "Adapter":
        h = C.c_void_p()
        _check(_lib.wgb_instance_request_adapter(self._h, C.byref(h)))
        return Adapter(h)

    def create_surface(self, on_present=None) -> "Surface":
        """InstanceInterface::create_surface (instance.rs:51-70).  The window is a host pixel sink: `on_present(pixels)` is
        called by Surface.present with the frame as an (H, W, 4) uint8 view of the window buffer (BGRA bytes)."""
        surface = Surface(None)
        target = None
        if on_present is not None:
            def _cb(_user, pixels, width, height, bytes_per_row):
                view = np.ctypeslib.as_array(C.cast(pixels, C.POINTER(C.c_uint8)), shape=(height, bytes_per_row))
                on_present(view[:, :width * 4].reshape(height, width, 4))
            surface._callback = _PRESENT_CALLBACK(_cb)        # kept alive as long as the surface
            target = C.byref(_SurfaceTarget(surface._callback, None))
        h = C.c_void_p()
        _check(_lib.wgb_instance_create_surface(self._h, target, C.byref(h)))
        surface._h = h
        return surface


class Adapter(_Handle):
    def get_info(self) -> dict:
        info = _AdapterInfo()
        _check(_lib.wgb_adapter_get_info(self._h, C.byref(info)))
        return {"name": info.name.decode(), "device_type": info.device_type, "cuda_device_count": info.cuda_device_count}

    def is_surface_supported(self, surface: "Surface") -> bool:
        out = C.c_int32()
        _check(_lib.wgb_adapter_is_surface_supported(self._h, surface._h, C.byref(out)))
        return bool(out.value)

    def request_device(self, cuda_device: int = CUDA_DEVICE_CURRENT, band_rank: int = 0, band_count: int = 1, features: int = 0):
        """`features`: FEATURE[...] bits -- WebGPU behaviour the reference accepts and ignores; 0 = parity mode."""
        desc = _DeviceDescriptor(cuda_device, band_rank, band_count, features)
        d, q = C.c_void_p(), C.c_void_p()
        _check(_lib.wgb_adapter_request_device(self._h, C.byref(desc), C.byref(d), C.byref(q)))
        dev = Device(d)
        queue = Queue(q)
        queue.device = dev
        return dev, queue


class Surface(_Handle):
    """SurfaceInterface + SurfaceOutputDetailInterface (surface.rs:51-198)."""
    _callback = None

    def get_capabilities(self, adapter: "Adapter") -> dict:
        caps = _SurfaceCapabilities()
        _check(_lib.wgb_surface_get_capabilities(self._h, adapter._h, C.byref(caps)))
        inv = lambda table, values, n: [next(k for k, v in table.items() if v == values[i]) for i in range(n)]   # noqa: E731
        return {"formats": inv(TEXTURE_FORMAT, caps.formats, caps.format_count),
                "present_modes": inv(PRESENT_MODE, caps.present_modes, caps.present_mode_count),
                "alpha_modes": inv(COMPOSITE_ALPHA_MODE, caps.alpha_modes, caps.alpha_mode_count),
                "usages": caps.usages}

    def configure(self, device: "Device", width: int, height: int, format: str = "bgra8unorm", usage: int = 16,
                  present_mode: str = "immediate", alpha_mode: str = "opaque", view_formats: Sequence[str] = ()):
        vf = (C.c_uint32 * max(len(view_formats), 1))(*[TEXTURE_FORMAT[f] for f in view_formats])
        cfg = _SurfaceConfiguration(usage, TEXTURE_FORMAT[format], width, height, PRESENT_MODE[present_mode],
                                    COMPOSITE_ALPHA_MODE[alpha_mode], len(view_formats), C.cast(vf, C.POINTER(C.c_uint32)))
        _check(_lib.wgb_surface_configure(self._h, device._h, C.byref(cfg)))
        self._extent = (width, height, format)

    def get_current_texture(self) -> "Texture":
        h, status = C.c_void_p(), C.c_uint32()
        _check(_lib.wgb_surface_get_current_texture(self._h, C.byref(h), C.byref(status)))
        t = Texture(h)
        t.width, t.height, t.format = self._extent
        t.layers = 1
        t.status = status.value
        return t

    def present(self):
        _check(_lib.wgb_surface_present(self._h))

    def texture_discard(self):
        _check(_lib.wgb_surface_texture_discard(self._h))

    def window_buffer(self):
        """(pixels as an (H, W, 4) uint8 view of the window buffer -- valid until the next configure --, presents so far)"""
        p, n, k = C.c_void_p(), C.c_uint64(), C.c_uint64()
        _check(_lib.wgb_surface_get_window_buffer(self._h, C.byref(p), C.byref(n), C.byref(k)))
        w, h, _ = self._extent
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(h, w, 4)), k.value


class Buffer(_Handle):
    size = 0

    def device_pointer(self):
        p, n = C.c_uint64(), C.c_uint64()
        _check(_lib.wgb_buffer_device_pointer(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def get_mapped_range(self, offset=0, size=WHOLE_SIZE) -> np.ndarray:
        """Mapped bytes as a writable numpy view (valid until unmap)."""
        p = C.c_void_p()
        _check(_lib.wgb_buffer_get_mapped_range(self._h, C.c_uint64(offset), C.c_uint64(size), C.byref(p)))
        n = self.size - offset if size == WHOLE_SIZE else size
        if n == 0:
            return np.zeros(0, dtype=np.uint8)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n,))

    def map_async(self, mode: str, offset=0, size=WHOLE_SIZE):
        _check(_lib.wgb_buffer_map_async(self._h, 1 if mode == "read" else 2, C.c_uint64(offset), C.c_uint64(size), None, None))

    def unmap(self):
        _check(_lib.wgb_buffer_unmap(self._h))


class Texture(_Handle):
    width = height = layers = 0
    format = "rgba8unorm"

    def create_view(self, base_array_layer: int = 0) -> "TextureView":
        h = C.c_void_p()
        if base_array_layer:
            class _D(C.Structure):
                _fields_ = [("base_array_layer", C.c_uint32), ("reserved", C.c_uint32)]
            d = _D(base_array_layer, 0)
            _check(_lib.wgb_texture_create_view(self._h, C.byref(d), C.byref(h)))
        else:
            _check(_lib.wgb_texture_create_view(self._h, None, C.byref(h)))
        v = TextureView(h)
        v.texture = self
        return v

    def read(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Texel bytes, row-major: what wgpu_cpu::dump_texture observes (lib.rs:111-173).  `out`: a caller-owned
        uint8 destination of the texture's size (e.g. pinned host memory, so the copy is one DMA)."""
        bpp = BYTES_PER_TEXEL[TEXTURE_FORMAT[self.format]]
        shape = (self.layers, self.height, self.width, bpp)
        if out is None:
            out = np.empty(shape, dtype=np.uint8)
        else:
            if out.dtype != np.uint8 or out.size != int(np.prod(shape)) or not out.flags["C_CONTIGUOUS"]:
                raise ValueError("out must be a C-contiguous uint8 array of the texture's byte size")
            out = out.reshape(shape)
        _check(_lib.wgb_texture_read(self._h, out.ctypes.data_as(C.c_void_p), C.c_uint64(out.nbytes)))
        if self.format == "depth32float":
            out = out.view(np.float32).reshape(self.layers, self.height, self.width)
        return out if self.layers > 1 else out[0]      # array textures come back with the layer axis first

    def read_pinned_async(self, out: np.ndarray):
        """The read-back that is not waited for: `out` (uint8, page-locked, the texture's byte size) holds the texels
        after Device.wait_readbacks(); submissions made in between run while the copy does."""
        if out.dtype != np.uint8 or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a C-contiguous uint8 array")
        _check(_lib.wgb_texture_read_pinned_async(self._h, out.ctypes.data_as(C.c_void_p), C.c_uint64(out.nbytes)))

    def dump_png(self, path: str):
        """wgpu_cpu::dump_texture (lib.rs:111-158)."""
        _check(_lib.wgb_texture_dump_png(self._h, os.fsencode(path)))

    def export_ipc(self) -> bytes:
        """cudaIpcMemHandle_t of the texel storage (peer-memory presenter, see wgpu_b200.h)."""
        h = (C.c_uint8 * 64)()
        _check(_lib.wgb_texture_export_ipc(self._h, h))
        return bytes(h)

    def device_pointer(self):
        p, n = C.c_uint64(), C.c_uint64()
        _check(_lib.wgb_texture_device_pointer(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value


class TextureView(_Handle):
    texture: Texture = None


class Sampler(_Handle):
    pass


class ShaderModule(_Handle):
    pass


class BindGroupLayout(_Handle):
    pass


class PipelineLayout(_Handle):
    pass


class BindGroup(_Handle):
    _keep = ()


class RenderPipeline(_Handle):
    _keep = ()

    def get_source(self) -> str:
        p = C.c_void_p()
        _check(_lib.wgb_render_pipeline_get_source(self._h, C.byref(p)))
        s = C.string_at(p).decode()
        _lib.wgb_free(p)
        return s


class CommandBuffer(_Handle):
    pass


class RenderPass(_Handle):
    def set_pipeline(self, pipeline: RenderPipeline):
        _check(_lib.wgb_render_pass_set_pipeline(self._h, pipeline._h))

    def set_bind_group(self, index: int, bind_group: Optional[BindGroup], offsets: Sequence[int] = ()):
        n = len(offsets)
        arr = (C.c_uint32 * max(n, 1))(*offsets)
        _check(_lib.wgb_render_pass_set_bind_group(self._h, index, bind_group._h if bind_group else None, arr if n else None, n))

    def set_index_buffer(self, buffer: Buffer, index_format: str, offset=0, size=WHOLE_SIZE):
        _check(_lib.wgb_render_pass_set_index_buffer(self._h, buffer._h, INDEX_FORMAT[index_format], C.c_uint64(offset), C.c_uint64(size)))

    def set_vertex_buffer(self, slot: int, buffer: Buffer, offset=0, size=WHOLE_SIZE):
        _check(_lib.wgb_render_pass_set_vertex_buffer(self._h, slot, buffer._h, C.c_uint64(offset), C.c_uint64(size)))

    def set_viewport(self, x, y, w, h, min_depth=0.0, max_depth=1.0):
        _check(_lib.wgb_render_pass_set_viewport(self._h, C.c_float(x), C.c_float(y), C.c_float(w), C.c_float(h),
                                                 C.c_float(min_depth), C.c_float(max_depth)))

    def set_scissor_rect(self, x, y, w, h):
        _check(_lib.wgb_render_pass_set_scissor_rect(self._h, x, y, w, h))

    def set_blend_constant(self, color):
        _check(_lib.wgb_render_pass_set_blend_constant(self._h, (C.c_double * 4)(*color)))

    def set_stencil_reference(self, reference: int):
        _check(_lib.wgb_render_pass_set_stencil_reference(self._h, reference))

    def draw(self, vertices: range, instances: range = range(0, 1)):
        _check(_lib.wgb_render_pass_draw(self._h, vertices.start, len(vertices), instances.start, len(instances)))

    def draw_indexed(self, indices: range, base_vertex: int = 0, instances: range = range(0, 1)):
        _check(_lib.wgb_render_pass_draw_indexed(self._h, indices.start, len(indices), C.c_int32(base_vertex),
                                                 instances.start, len(instances)))

    def end(self):
        _check(_lib.wgb_render_pass_end(self._h))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.end()


class CommandEncoder(_Handle):
    def begin_render_pass(self, color_attachments, depth_stencil_attachment=None) -> RenderPass:
        """color_attachments: [{view, load: ("clear", (r,g,b,a)) | "load", store: "store"|"discard"}];
        depth_stencil_attachment: {view, depth_load: ("clear", v) | "load" | None, depth_store}"""
        n = len(color_attachments)
        cas = (_ColorAttachment * max(n, 1))()
        keep = []
        for i, a in enumerate(color_attachments):
            cas[i].view = a["view"]._h if a.get("view") is not None else None
            keep.append(a.get("view"))
            load = a.get("load", ("clear", (0, 0, 0, 0)))
            if isinstance(load, tuple):
                cas[i].load_op = 0
                for k in range(4):
                    cas[i].clear_value[k] = float(load[1][k])
            else:
                cas[i].load_op = 1
            cas[i].store_op = 0 if a.get("store", "store") == "store" else 1
        desc = _RenderPassDescriptor(n, cas, None)
        if depth_stencil_attachment is not None:
            d = _DepthStencilAttachment()
            d.view = depth_stencil_attachment["view"]._h
            keep.append(depth_stencil_attachment["view"])
            dl = depth_stencil_attachment.get("depth_load", ("clear", 1.0))
            d.has_depth_ops = 0 if dl is None else 1
            if isinstance(dl, tuple):
                d.depth_load_op = 0
                d.depth_clear_value = float(dl[1])
            else:
                d.depth_load_op = 1
            d.depth_store_op = 0 if depth_stencil_attachment.get("depth_store", "store") == "store" else 1
            d.has_stencil_ops = 1 if depth_stencil_attachment.get("stencil_ops") else 0
            desc.depth_stencil_attachment = C.pointer(d)
        h = C.c_void_p()
        _check(_lib.wgb_command_encoder_begin_render_pass(self._h, C.byref(desc), C.byref(h)))
        return RenderPass(h)

    def copy_buffer_to_buffer(self, source: Buffer, source_offset: int, destination: Buffer, destination_offset: int, size: int = WHOLE_SIZE):
        _check(_lib.wgb_command_encoder_copy_buffer_to_buffer(self._h, source._h, C.c_uint64(source_offset), destination._h,
                                                              C.c_uint64(destination_offset), C.c_uint64(size)))

    def copy_buffer_to_texture(self, source: Buffer, destination: Texture, size=None, offset=0, bytes_per_row=0, origin=(0, 0), layer=0):
        w, h = size if size is not None else (destination.width, destination.height)
        src = _TexelCopyBufferInfo(source._h, offset, bytes_per_row, 0)
        dst = _TexelCopyTextureInfo(destination._h, origin[0], origin[1], layer)
        _check(_lib.wgb_command_encoder_copy_buffer_to_texture(self._h, C.byref(src), C.byref(dst), w, h))

    def copy_texture_to_buffer(self, source: Texture, destination: Buffer, size=None, offset=0, bytes_per_row=0, origin=(0, 0), layer=0):
        w, h = size if size is not None else (source.width, source.height)
        src = _TexelCopyTextureInfo(source._h, origin[0], origin[1], layer)
        dst = _TexelCopyBufferInfo(destination._h, offset, bytes_per_row, 0)
        _check(_lib.wgb_command_encoder_copy_texture_to_buffer(self._h, C.byref(src), C.byref(dst), w, h))

    def copy_texture_to_texture(self, source: Texture, destination: Texture, size=None, src_origin=(0, 0), dst_origin=(0, 0)):
        w, h = size if size is not None else (source.width, source.height)
        src = _TexelCopyTextureInfo(source._h, src_origin[0], src_origin[1], 0)
        dst = _TexelCopyTextureInfo(destination._h, dst_origin[0], dst_origin[1], 0)
        _check(_lib.wgb_command_encoder_copy_texture_to_texture(self._h, C.byref(src), C.byref(dst), w, h))

    def clear_buffer(self, buffer: Buffer, offset: int = 0, size: int = WHOLE_SIZE):
        _check(_lib.wgb_command_encoder_clear_buffer(self._h, buffer._h, C.c_uint64(offset), C.c_uint64(size)))

    def clear_texture(self, texture: Texture):
        _check(_lib.wgb_command_encoder_clear_texture(self._h, texture._h))

    def finish(self) -> CommandBuffer:
        h = C.c_void_p()
        _check(_lib.wgb_command_encoder_finish(self._h, C.byref(h)))
        return CommandBuffer(h)


class Queue(_Handle):
    device: "Device" = None

    def write_buffer(self, buffer: Buffer, offset: int, data):
        a = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        _check(_lib.wgb_queue_write_buffer(self._h, buffer._h, C.c_uint64(offset), a.ctypes.data_as(C.c_void_p), C.c_uint64(a.nbytes)))

    def write_buffer_pinned_async(self, buffer: Buffer, offset: int, data):
        """Zero-copy upload from page-locked memory on the copy stream; `data` must stay valid and unchanged until
        wait_uploads() (or a poll(wait) of a later submission that uses the buffer)."""
        a = data if isinstance(data, np.ndarray) and data.flags["C_CONTIGUOUS"] else np.ascontiguousarray(data)
        a = a.view(np.uint8).reshape(-1)
        _check(_lib.wgb_queue_write_buffer_pinned_async(self._h, buffer._h, C.c_uint64(offset), a.ctypes.data_as(C.c_void_p), C.c_uint64(a.nbytes)))

    def wait_uploads(self):
        _check(_lib.wgb_queue_wait_uploads(self._h))

    def write_texture(self, texture: Texture, data, bytes_per_row: int = 0, origin=(0, 0), size=None):
        a = np.ascontiguousarray(data)
        w, h = size if size is not None else (texture.width, texture.height)
        raw = a.view(np.uint8).reshape(-1)
        _check(_lib.wgb_queue_write_texture(self._h, texture._h, origin[0], origin[1], raw.ctypes.data_as(C.c_void_p),
                                            C.c_uint64(raw.nbytes), bytes_per_row, w, h))

    def submit(self, command_buffers: Sequence[CommandBuffer]) -> int:
        n = len(command_buffers)
        arr = (C.c_void_p * max(n, 1))(*[cb._h for cb in command_buffers])
        idx = C.c_uint64()
        _check(_lib.wgb_queue_submit(self._h, arr, n, C.byref(idx)))
        return idx.value


class Device(_Handle):
    def poll(self, wait: bool = True, submission_index: Optional[int] = None, timeout_ns: int = 0) -> int:
        """PollType::Wait{submission_index, timeout} / Poll (device.rs:237-295); returns WGB_POLL_*."""
        out = C.c_int32()
        idx = WHOLE_SIZE if submission_index is None else submission_index
        _check(_lib.wgb_device_poll(self._h, 1 if wait else 0, C.c_uint64(idx), C.c_uint64(timeout_ns), C.byref(out)))
        return out.value

    def wait_readbacks(self):
        """Waits for every Texture.read_pinned_async started on this device."""
        _check(_lib.wgb_device_wait_readbacks(self._h))

    def create_buffer(self, size: int, usage: int = 0, mapped_at_creation: bool = False) -> Buffer:
        desc = _BufferDescriptor(size, usage, 1 if mapped_at_creation else 0)
        h = C.c_void_p()
        _check(_lib.wgb_device_create_buffer(self._h, C.byref(desc), C.byref(h)))
        b = Buffer(h)
        b.size = size
        return b

    def create_buffer_init(self, contents, usage: int = 0) -> Buffer:
        """wgpu::util::DeviceExt::create_buffer_init: create mapped, copy, unmap (hello_mesh.rs:540-544)."""
        a = np.ascontiguousarray(contents).view(np.uint8).reshape(-1)
        b = self.create_buffer(a.nbytes, usage, mapped_at_creation=True)
        if a.nbytes:
            b.get_mapped_range()[:] = a
        b.unmap()
        return b

    def create_texture(self, width: int, height: int, format: str, layers: int = 1, usage: int = 0,
                       mip_level_count: int = 1, sample_count: int = 1) -> Texture:
        desc = _TextureDescriptor(width, height, layers, mip_level_count, sample_count, TEXTURE_FORMAT[format], usage)
        h = C.c_void_p()
        _check(_lib.wgb_device_create_texture(self._h, C.byref(desc), C.byref(h)))
        t = Texture(h)
        t.width, t.height, t.layers, t.format = width, height, layers, format
        return t

    def create_texture_with_data(self, queue: Queue, width, height, format, data) -> Texture:
        """wgpu::util::DeviceExt::create_texture_with_data -> Queue::write_texture (hello_texture.rs:184-201)."""
        t = self.create_texture(width, height, format)
        queue.write_texture(t, data)
        return t

    def import_texture_ipc(self, handle: bytes, width: int, height: int, format: str, layers: int = 1) -> Texture:
        """Map a texture another process exported with Texture.export_ipc (its device must be an NVLink peer)."""
        desc = _TextureDescriptor(width, height, layers, 1, 1, TEXTURE_FORMAT[format], 0)
        h = C.c_void_p()
        buf = (C.c_uint8 * 64)(*handle)
        _check(_lib.wgb_device_import_texture_ipc(self._h, buf, C.byref(desc), C.byref(h)))
        t = Texture(h)
        t.width, t.height, t.layers, t.format = width, height, layers, format
        return t

    def create_sampler(self, address_mode_u="clamp-to-edge", address_mode_v="clamp-to-edge", address_mode_w="clamp-to-edge",
                       mag_filter="nearest", min_filter="nearest", mipmap_filter="nearest") -> Sampler:
        desc = _SamplerDescriptor(ADDRESS_MODE[address_mode_u], ADDRESS_MODE[address_mode_v], ADDRESS_MODE[address_mode_w],
                                  FILTER_MODE[mag_filter], FILTER_MODE[min_filter], FILTER_MODE[mipmap_filter])
        h = C.c_void_p()
        _check(_lib.wgb_device_create_sampler(self._h, C.byref(desc), C.byref(h)))
        return Sampler(h)

    def create_shader_module(self, wgsl: Optional[str] = None, emitted: Sequence[tuple] = ()) -> ShaderModule:
        """wgsl: WGSL source; emitted: [(stage, entry_point, cuda_source)] pre-translated entry points."""
        n = len(emitted)
        arr = (_EmittedEntryPoint * max(n, 1))()
        for i, (stage, entry, src) in enumerate(emitted):
            arr[i] = _EmittedEntryPoint(stage, entry.encode(), src.encode())
        desc = _ShaderModuleDescriptor(wgsl.encode() if wgsl is not None else None, n, arr)
        h = C.c_void_p()
        _check(_lib.wgb_device_create_shader_module(self._h, C.byref(desc), C.byref(h)))
        return ShaderModule(h)

    def create_bind_group_layout(self, entries: Sequence[tuple] = ()) -> BindGroupLayout:
        n = len(entries)
        arr = (_BindGroupLayoutEntry * max(n, 1))(*[_BindGroupLayoutEntry(*e) for e in entries])
        h = C.c_void_p()
        _check(_lib.wgb_device_create_bind_group_layout(self._h, arr, n, C.byref(h)))
        return BindGroupLayout(h)

    def create_pipeline_layout(self, bind_group_layouts: Sequence[BindGroupLayout] = ()) -> PipelineLayout:
        n = len(bind_group_layouts)
        arr = (C.c_void_p * max(n, 1))(*[l._h for l in bind_group_layouts])
        h = C.c_void_p()
        _check(_lib.wgb_device_create_pipeline_layout(self._h, arr, n, C.byref(h)))
        return PipelineLayout(h)

    def create_bind_group(self, layout: Optional[BindGroupLayout], entries: Sequence[dict]) -> BindGroup:
        """entries: {binding, buffer[, offset, size]} | {binding, texture_view} | {binding, sampler}"""
        n = len(entries)
        arr = (_BindGroupEntry * max(n, 1))()
        keep = []
        for i, e in enumerate(entries):
            arr[i].binding = e["binding"]
            if "buffer" in e:
                arr[i].kind = 1
                arr[i].buffer = e["buffer"]._h
                arr[i].offset = e.get("offset", 0)
                arr[i].size = e.get("size", WHOLE_SIZE)
                keep.append(e["buffer"])
            elif "texture_view" in e:
                arr[i].kind = 2
                arr[i].texture_view = e["texture_view"]._h
                keep.append(e["texture_view"])
            else:
                arr[i].kind = 3
                arr[i].sampler = e["sampler"]._h
                keep.append(e["sampler"])
        h = C.c_void_p()
        _check(_lib.wgb_device_create_bind_group(self._h, layout._h if layout else None, arr, n, C.byref(h)))
        g = BindGroup(h)
        g._keep = tuple(keep)
        return g

    def create_render_pipeline(self, *, vertex_module: ShaderModule, vertex_entry_point="vs_main", vertex_buffers=(),
                               topology="triangle-list", strip_index_format=None, front_face="ccw", cull_mode=None,
                               depth_stencil: Optional[dict] = None, fragment_module: Optional[ShaderModule] = None,
                               fragment_entry_point="fs_main", targets=(), layout: Optional[PipelineLayout] = None,
                               polygon_mode: int = 0, multisample_count: int = 1) -> RenderPipeline:
        """vertex_buffers: [{array_stride, step_mode, attributes: [(format, offset, shader_location)]}];
        depth_stencil: {format, depth_write_enabled, depth_compare}; targets: [format | {format, blend, write_mask}]"""
        nvb = len(vertex_buffers)
        vbs = (_VertexBufferLayout * max(nvb, 1))()
        keep = []
        for i, vb in enumerate(vertex_buffers):
            attrs = vb["attributes"]
            aa = (_VertexAttribute * max(len(attrs), 1))(*[_VertexAttribute(VERTEX_FORMAT[f], off, loc) for f, off, loc in attrs])
            keep.append(aa)
            vbs[i] = _VertexBufferLayout(vb["array_stride"], STEP_MODE[vb.get("step_mode", "vertex")], len(attrs), aa)
        nt = len(targets)
        ts = (_ColorTargetState * max(nt, 1))()
        for i, t in enumerate(targets):
            if isinstance(t, dict):
                # blend: {"color": (src_factor, dst_factor, operation), "alpha": (...)} as in wgpu::BlendState
                blend = t.get("blend")
                ts[i] = _ColorTargetState(TEXTURE_FORMAT[t["format"]], 1 if blend else 0, t.get("write_mask", 15),
                                          _blend_component(blend["color"]) if blend else _BlendComponent(1, 0, 0),
                                          _blend_component(blend["alpha"]) if blend else _BlendComponent(1, 0, 0))
            else:
                ts[i] = _ColorTargetState(TEXTURE_FORMAT[t], 0, 15, _BlendComponent(1, 0, 0), _BlendComponent(1, 0, 0))
        d = _RenderPipelineDescriptor()
        d.layout = layout._h if layout else None
        d.vertex_module = vertex_module._h
        d.vertex_entry_point = vertex_entry_point.encode()
        d.vertex_buffer_count, d.vertex_buffers = nvb, vbs
        d.topology, d.strip_index_format = TOPOLOGY[topology], INDEX_FORMAT[strip_index_format]
        d.front_face, d.cull_mode, d.polygon_mode = FRONT_FACE[front_face], CULL_MODE[cull_mode], polygon_mode
        if depth_stencil is not None:
            d.has_depth_stencil = 1
            d.depth_format = TEXTURE_FORMAT[depth_stencil.get("format", "depth32float")]
            d.depth_write_enabled = 1 if depth_stencil.get("depth_write_enabled", True) else 0
            d.depth_compare = COMPARE[depth_stencil.get("depth_compare", "less")]
        d.multisample_count = multisample_count
        if fragment_module is not None:
            d.fragment_module = fragment_module._h
            d.fragment_entry_point = fragment_entry_point.encode()
        d.target_count, d.targets = nt, ts
        h = C.c_void_p()
        _check(_lib.wgb_device_create_render_pipeline(self._h, C.byref(d), C.byref(h)))
        p = RenderPipeline(h)
        p._keep = (vertex_module, fragment_module, layout)
        return p

    def create_command_encoder(self) -> CommandEncoder:
        h = C.c_void_p()
        _check(_lib.wgb_device_create_command_encoder(self._h, C.byref(h)))
        return CommandEncoder(h)

    # ---- measurement / multi-GPU ----
    def last_pass_stats(self) -> dict:
        s = PassStats()
        _check(_lib.wgb_device_get_last_pass_stats(self._h, C.byref(s)))
        return s.as_dict()

    def set_coverage_capture(self, enabled: bool):
        _check(_lib.wgb_device_set_coverage_capture(self._h, 1 if enabled else 0))

    def read_coverage(self, width: int, height: int) -> np.ndarray:
        out = np.empty((height, width), dtype=np.uint32)
        _check(_lib.wgb_device_read_coverage(self._h, out.ctypes.data_as(C.c_void_p), C.c_uint64(width * height)))
        return out

    def set_band(self, rank: int, count: int):
        _check(_lib.wgb_device_set_band(self._h, rank, count))

    def band_rows(self, height: int):
        a, b = C.c_uint32(), C.c_uint32()
        _check(_lib.wgb_device_get_band_rows(self._h, height, C.byref(a), C.byref(b)))
        return a.value, b.value

    def timer_begin(self):
        _check(_lib.wgb_device_timer_begin(self._h))

    def timer_end(self) -> float:
        ms = C.c_float()
        _check(_lib.wgb_device_timer_end(self._h, C.byref(ms)))
        return ms.value

    def stream(self) -> int:
        p = C.c_void_p()
        _check(_lib.wgb_device_get_stream(self._h, C.byref(p)))
        return p.value or 0


def translate_wgsl(wgsl: str, stage: int, entry_point: str) -> str:
    """The WGSL -> CUDA C++ emitter on its own (no device needed)."""
    lib = load_library()
    p = C.c_void_p()
    _check(lib.wgb_translate_wgsl(wgsl.encode(), stage, entry_point.encode(), C.byref(p)))
    s = C.string_at(p).decode()
    lib.wgb_free(p)
    return s


def dump_texture(texture: Texture, path: str):
    """`wgpu_cpu::dump_texture(&texture, path, None)` (lib.rs:111-158): what the reference's examples and tests call to
    look at a render target."""
    texture.dump_png(path)


def write_png(path: str, pixels: np.ndarray):
    """The library's PNG writer on a host array of shape (h, w), (h, w, 3) or (h, w, 4), uint8."""
    a = np.ascontiguousarray(pixels, dtype=np.uint8)
    ch = 1 if a.ndim == 2 else a.shape[2]
    load_library()
    _check(_lib.wgb_write_png(os.fsencode(path), a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[0], ch))
