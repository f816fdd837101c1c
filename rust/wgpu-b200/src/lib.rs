//! wgpu custom backend for NVIDIA B200: every live method of the `wgpu::custom::*Interface` traits that
//! jgraef/wgpu-cpu implements is a thin call into `libwgpu_b200.so` (`include/wgpu_b200.h`);
//! the methods wgpu-cpu leaves `todo!()` stay unimplemented here, except the command-encoder copies and clears,
//! which the library provides.  An application switches backends by changing one constructor:
//!
//! ```ignore
//! let instance = wgpu_b200::instance(wgpu_b200::Config::default());   // was: wgpu_cpu::instance(Default::default())
//! ```
//!
//! NOT COMPILED IN THE BUILD IMAGE (no rustc / cargo there).  The trait method signatures follow the reference's
//! `impl` blocks (wgpu-cpu/src/{instance,adapter,device,buffer,texture,command}.rs, render_pass/mod.rs); the bodies
//! are this crate's own.  Error behaviour follows the reference: hot-path errors panic (device.rs:46-48).

use std::ffi::{CStr, CString};
use std::fmt;
use std::ops::Range;
use std::pin::Pin;
use std::ptr::{self, NonNull};
use std::sync::Arc;

use wgpu::custom::*;
use wgpu_b200_sys as sys;

// ---------------------------------------------------------------------------------------------------------------
// handles and errors
// ---------------------------------------------------------------------------------------------------------------

/// Owning reference to a library object.  The library counts references (`wgb_retain` / `wgb_release`), which gives
/// the `Clone`-over-`Arc` semantics of the reference's objects (buffer.rs:22-29): commands keep what they use alive.
pub struct Handle<T>(NonNull<T>);

impl<T> Handle<T> {
    /// Takes over the reference returned by a `wgb_*create*` call.
    unsafe fn adopt(raw: *mut T) -> Self {
        Handle(NonNull::new(raw).expect("libwgpu_b200 returned a null handle with status OK"))
    }
    fn raw(&self) -> *mut T {
        self.0.as_ptr()
    }
}
impl<T> Clone for Handle<T> {
    fn clone(&self) -> Self {
        unsafe { sys::wgb_retain(self.0.as_ptr().cast()) };
        Handle(self.0)
    }
}
impl<T> Drop for Handle<T> {
    fn drop(&mut self) {
        unsafe { sys::wgb_release(self.0.as_ptr().cast()) }
    }
}
impl<T> fmt::Debug for Handle<T> {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        write!(f, "Handle({:p})", self.0)
    }
}
// every entry point of the library is thread safe (include/wgpu_b200.h, "Threading")
unsafe impl<T> Send for Handle<T> {}
unsafe impl<T> Sync for Handle<T> {}

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::wgb_last_error()) }.to_string_lossy().into_owned()
}

/// The reference has no error returns on this path: it unwraps and panics (device.rs:97,133; state.rs:243).
#[track_caller]
fn check(status: sys::wgb_status) {
    if status != sys::WGB_OK {
        panic!("wgpu-b200 error {status}: {}", last_error());
    }
}

macro_rules! create {
    ($ty:ty, |$out:ident| $call:expr) => {{
        let mut $out: *mut $ty = ptr::null_mut();
        check(unsafe { $call });
        unsafe { Handle::<$ty>::adopt($out) }
    }};
}

// ---------------------------------------------------------------------------------------------------------------
// configuration and the one public constructor (wgpu-cpu/src/lib.rs:22-27)
// ---------------------------------------------------------------------------------------------------------------

/// What `wgpu_cpu::Config` is to the reference.  `features == 0` renders exactly what wgpu-cpu renders (blend states,
/// write masks, sRGB targets, viewport depth range and dynamic offsets accepted and ignored); each `FEATURE_*` bit
/// gives one of them its WebGPU meaning.
#[derive(Clone, Copy, Debug)]
pub struct Config {
    /// CUDA device ordinal, or -1 for the current device.
    pub cuda_device: i32,
    /// Sort-first multi-GPU: this process renders band `band_rank` of `band_count` (RANK / WORLD_SIZE).
    pub band_rank: u32,
    pub band_count: u32,
    pub features: u32,
}
impl Default for Config {
    fn default() -> Self {
        Config { cuda_device: -1, band_rank: 0, band_count: 1, features: 0 }
    }
}
pub const FEATURE_VIEWPORT_DEPTH_RANGE: u32 = sys::WGB_FEATURE_VIEWPORT_DEPTH_RANGE;
pub const FEATURE_COLOR_WRITE_MASK: u32 = sys::WGB_FEATURE_COLOR_WRITE_MASK;
pub const FEATURE_SRGB_ENCODE: u32 = sys::WGB_FEATURE_SRGB_ENCODE;
pub const FEATURE_DYNAMIC_OFFSETS: u32 = sys::WGB_FEATURE_DYNAMIC_OFFSETS;
pub const FEATURE_BLEND: u32 = sys::WGB_FEATURE_BLEND;

pub fn instance(config: Config) -> wgpu::Instance {
    let cfg = sys::wgb_instance_config { reserved: 0 };
    let handle = create!(sys::wgb_instance_t, |out| sys::wgb_create_instance(&cfg, &mut out));
    wgpu::Instance::from_custom(Instance { handle, config })
}

// ---------------------------------------------------------------------------------------------------------------
// instance / adapter  (instance.rs:43-121, adapter.rs:23-105)
// ---------------------------------------------------------------------------------------------------------------

#[derive(Debug)]
pub struct Instance {
    handle: Handle<sys::wgb_instance_t>,
    config: Config,
}

impl Instance {
    fn adapter(&self) -> DispatchAdapter {
        let handle = create!(sys::wgb_adapter_t, |out| sys::wgb_instance_request_adapter(self.handle.raw(), &mut out));
        DispatchAdapter::custom(Adapter { handle, config: self.config })
    }
}

impl InstanceInterface for Instance {
    fn new(_desc: wgpu::InstanceDescriptor) -> Self
    where
        Self: Sized,
    {
        unreachable!("wgpu_b200::instance(Config) is the constructor");
    }

    unsafe fn create_surface(&self, target: wgpu::SurfaceTargetUnsafe) -> Result<DispatchSurface, wgpu::CreateSurfaceError> {
        // instance.rs:51-70: with the `softbuffer` feature the window's pixel buffer is the sink the library presents
        // into (surface::Surface below); without it there are no surfaces, as in the reference
        #[cfg(feature = "softbuffer")]
        {
            surface::Surface::new(self.handle.clone(), target)
                .map(DispatchSurface::custom)
                .map_err(|error| wgpu::CreateSurfaceError::custom(error))
        }
        #[cfg(not(feature = "softbuffer"))]
        {
            let _ = target;
            Err(wgpu::CreateSurfaceError::custom("wgpu-b200 compiled without softbuffer feature, so no surfaces are supported".to_owned()))
        }
    }

    fn request_adapter(&self, _options: &wgpu::RequestAdapterOptions<'_, '_>) -> Pin<Box<dyn RequestAdapterFuture>> {
        // every surface of this crate is supported by its one adapter (adapter.rs:44-54)
        let result = Ok(self.adapter());
        Box::pin(async move { result })
    }

    fn poll_all_devices(&self, _force_wait: bool) -> bool {
        true
    }

    fn enumerate_adapters(&self, _backends: wgpu::Backends) -> Pin<Box<dyn EnumerateAdapterFuture>> {
        let adapters = vec![self.adapter()];
        Box::pin(async move { adapters })
    }

    fn wgsl_language_features(&self) -> wgpu::WgslLanguageFeatures {
        wgpu::WgslLanguageFeatures::empty()
    }
}

#[derive(Debug)]
pub struct Adapter {
    handle: Handle<sys::wgb_adapter_t>,
    config: Config,
}

impl AdapterInterface for Adapter {
    fn request_device(&self, desc: &wgpu::DeviceDescriptor<'_>) -> Pin<Box<dyn RequestDeviceFuture>> {
        let dd = sys::wgb_device_descriptor {
            cuda_device: self.config.cuda_device,
            band_rank: self.config.band_rank,
            band_count: self.config.band_count,
            features: self.config.features,
        };
        let (features, limits) = (desc.required_features, desc.required_limits.clone());
        let mut device: sys::wgb_device = ptr::null_mut();
        let mut queue: sys::wgb_queue = ptr::null_mut();
        let status = unsafe { sys::wgb_adapter_request_device(self.handle.raw(), &dd, &mut device, &mut queue) };
        let result = if status == sys::WGB_OK {
            let device = Device { handle: unsafe { Handle::adopt(device) }, features, limits };
            let queue = Queue { handle: unsafe { Handle::adopt(queue) } };
            Ok((DispatchDevice::custom(device), DispatchQueue::custom(queue)))
        } else {
            // no CUDA device, or not an sm_100a one: the library refuses, there is no CPU fallback
            Err(wgpu::RequestDeviceError::custom(last_error()))
        };
        Box::pin(async move { result })
    }

    fn is_surface_supported(&self, surface: &DispatchSurface) -> bool {
        #![allow(unused)]
        let mut supported = false;
        #[cfg(feature = "softbuffer")]
        {
            supported = surface.as_custom::<surface::Surface>().is_some();
        }
        supported
    }

    fn features(&self) -> wgpu::Features {
        wgpu::Features::default()
    }

    fn limits(&self) -> wgpu::Limits {
        wgpu::Limits::downlevel_defaults()
    }

    fn downlevel_capabilities(&self) -> wgpu::DownlevelCapabilities {
        wgpu::DownlevelCapabilities::default()
    }

    fn get_info(&self) -> wgpu::AdapterInfo {
        let mut info: sys::wgb_adapter_info = unsafe { std::mem::zeroed() };
        check(unsafe { sys::wgb_adapter_get_info(self.handle.raw(), &mut info) });
        let name = unsafe { CStr::from_ptr(info.name.as_ptr()) }.to_string_lossy().into_owned();
        wgpu::AdapterInfo {
            name,
            vendor: 0x10DE,
            device: 0,
            device_type: wgpu::DeviceType::DiscreteGpu,
            driver: "libwgpu_b200".to_owned(),
            driver_info: unsafe { CStr::from_ptr(sys::wgb_version()) }.to_string_lossy().into_owned(),
            backend: wgpu::Backend::Noop,
            ..wgpu::AdapterInfo::new(wgpu::DeviceType::DiscreteGpu, wgpu::Backend::Noop)
        }
    }

    fn get_texture_format_features(&self, format: wgpu::TextureFormat) -> wgpu::TextureFormatFeatures {
        let mut allowed_usages = wgpu::TextureUsages::COPY_SRC | wgpu::TextureUsages::COPY_DST;
        if convert::texture_format(format).is_some() {
            allowed_usages |= wgpu::TextureUsages::TEXTURE_BINDING | wgpu::TextureUsages::RENDER_ATTACHMENT;
        }
        wgpu::TextureFormatFeatures { allowed_usages, flags: wgpu::TextureFormatFeatureFlags::empty() }
    }

    fn get_presentation_timestamp(&self) -> wgpu::PresentationTimestamp {
        wgpu::PresentationTimestamp::INVALID_TIMESTAMP
    }

    fn cooperative_matrix_properties(&self) -> Vec<wgpu::wgt::CooperativeMatrixProperties> {
        Vec::new()
    }
}

// ---------------------------------------------------------------------------------------------------------------
// wgpu enums -> the library's constants
// ---------------------------------------------------------------------------------------------------------------

mod convert {
    use super::sys;

    pub fn texture_format(f: wgpu::TextureFormat) -> Option<u32> {
        use wgpu::TextureFormat as F;
        Some(match f {
            F::Rgba8Unorm => sys::WGB_TEXTURE_FORMAT_RGBA8_UNORM,
            F::Rgba8UnormSrgb => sys::WGB_TEXTURE_FORMAT_RGBA8_UNORM_SRGB,
            F::Bgra8Unorm => sys::WGB_TEXTURE_FORMAT_BGRA8_UNORM,
            F::Bgra8UnormSrgb => sys::WGB_TEXTURE_FORMAT_BGRA8_UNORM_SRGB,
            F::R8Unorm => sys::WGB_TEXTURE_FORMAT_R8_UNORM,
            F::Rg8Unorm => sys::WGB_TEXTURE_FORMAT_RG8_UNORM,
            F::Rgba8Snorm => sys::WGB_TEXTURE_FORMAT_RGBA8_SNORM,
            F::Depth32Float => sys::WGB_TEXTURE_FORMAT_DEPTH32_FLOAT,
            _ => return None,
        })
    }
    pub fn topology(t: wgpu::PrimitiveTopology) -> u32 {
        use wgpu::PrimitiveTopology as T;
        match t {
            T::PointList => sys::WGB_TOPOLOGY_POINT_LIST,
            T::LineList => sys::WGB_TOPOLOGY_LINE_LIST,
            T::LineStrip => sys::WGB_TOPOLOGY_LINE_STRIP,
            T::TriangleList => sys::WGB_TOPOLOGY_TRIANGLE_LIST,
            T::TriangleStrip => sys::WGB_TOPOLOGY_TRIANGLE_STRIP,
        }
    }
    pub fn index_format(f: Option<wgpu::IndexFormat>) -> u32 {
        match f {
            None => sys::WGB_INDEX_FORMAT_NONE,
            Some(wgpu::IndexFormat::Uint16) => sys::WGB_INDEX_FORMAT_UINT16,
            Some(wgpu::IndexFormat::Uint32) => sys::WGB_INDEX_FORMAT_UINT32,
        }
    }
    pub fn compare(c: wgpu::CompareFunction) -> u32 {
        use wgpu::CompareFunction as C;
        match c {
            C::Never => sys::WGB_COMPARE_NEVER,
            C::Less => sys::WGB_COMPARE_LESS,
            C::Equal => sys::WGB_COMPARE_EQUAL,
            C::LessEqual => sys::WGB_COMPARE_LESS_EQUAL,
            C::Greater => sys::WGB_COMPARE_GREATER,
            C::NotEqual => sys::WGB_COMPARE_NOT_EQUAL,
            C::GreaterEqual => sys::WGB_COMPARE_GREATER_EQUAL,
            C::Always => sys::WGB_COMPARE_ALWAYS,
        }
    }
    pub fn address_mode(m: wgpu::AddressMode) -> u32 {
        match m {
            wgpu::AddressMode::ClampToEdge => sys::WGB_ADDRESS_MODE_CLAMP_TO_EDGE,
            wgpu::AddressMode::Repeat => sys::WGB_ADDRESS_MODE_REPEAT,
            wgpu::AddressMode::MirrorRepeat => sys::WGB_ADDRESS_MODE_MIRROR_REPEAT,
            wgpu::AddressMode::ClampToBorder => sys::WGB_ADDRESS_MODE_CLAMP_TO_BORDER,
        }
    }
    pub fn filter(m: wgpu::FilterMode) -> u32 {
        match m {
            wgpu::FilterMode::Nearest => sys::WGB_FILTER_MODE_NEAREST,
            wgpu::FilterMode::Linear => sys::WGB_FILTER_MODE_LINEAR,
        }
    }
    pub fn vertex_format(f: wgpu::VertexFormat) -> u32 {
        use wgpu::VertexFormat as V;
        match f {
            V::Float32 => sys::WGB_VERTEX_FORMAT_FLOAT32,
            V::Float32x2 => sys::WGB_VERTEX_FORMAT_FLOAT32X2,
            V::Float32x3 => sys::WGB_VERTEX_FORMAT_FLOAT32X3,
            V::Float32x4 => sys::WGB_VERTEX_FORMAT_FLOAT32X4,
            V::Uint32 => sys::WGB_VERTEX_FORMAT_UINT32,
            V::Sint32 => sys::WGB_VERTEX_FORMAT_SINT32,
            V::Uint32x2 => sys::WGB_VERTEX_FORMAT_UINT32X2,
            V::Uint32x3 => sys::WGB_VERTEX_FORMAT_UINT32X3,
            V::Uint32x4 => sys::WGB_VERTEX_FORMAT_UINT32X4,
            V::Sint32x2 => sys::WGB_VERTEX_FORMAT_SINT32X2,
            V::Sint32x3 => sys::WGB_VERTEX_FORMAT_SINT32X3,
            V::Sint32x4 => sys::WGB_VERTEX_FORMAT_SINT32X4,
            other => panic!("wgpu-b200: vertex format {other:?} is not supported"),
        }
    }
    pub fn blend_factor(f: wgpu::BlendFactor) -> u32 {
        use wgpu::BlendFactor as B;
        match f {
            B::Zero => sys::WGB_BLEND_FACTOR_ZERO,
            B::One => sys::WGB_BLEND_FACTOR_ONE,
            B::Src => sys::WGB_BLEND_FACTOR_SRC,
            B::OneMinusSrc => sys::WGB_BLEND_FACTOR_ONE_MINUS_SRC,
            B::SrcAlpha => sys::WGB_BLEND_FACTOR_SRC_ALPHA,
            B::OneMinusSrcAlpha => sys::WGB_BLEND_FACTOR_ONE_MINUS_SRC_ALPHA,
            B::Dst => sys::WGB_BLEND_FACTOR_DST,
            B::OneMinusDst => sys::WGB_BLEND_FACTOR_ONE_MINUS_DST,
            B::DstAlpha => sys::WGB_BLEND_FACTOR_DST_ALPHA,
            B::OneMinusDstAlpha => sys::WGB_BLEND_FACTOR_ONE_MINUS_DST_ALPHA,
            B::SrcAlphaSaturated => sys::WGB_BLEND_FACTOR_SRC_ALPHA_SATURATED,
            B::Constant => sys::WGB_BLEND_FACTOR_CONSTANT,
            B::OneMinusConstant => sys::WGB_BLEND_FACTOR_ONE_MINUS_CONSTANT,
            other => panic!("wgpu-b200: blend factor {other:?} (dual-source blending) is not supported"),
        }
    }
    pub fn blend_component(c: &wgpu::BlendComponent) -> sys::wgb_blend_component {
        use wgpu::BlendOperation as O;
        sys::wgb_blend_component {
            src_factor: blend_factor(c.src_factor),
            dst_factor: blend_factor(c.dst_factor),
            operation: match c.operation {
                O::Add => sys::WGB_BLEND_OPERATION_ADD,
                O::Subtract => sys::WGB_BLEND_OPERATION_SUBTRACT,
                O::ReverseSubtract => sys::WGB_BLEND_OPERATION_REVERSE_SUBTRACT,
                O::Min => sys::WGB_BLEND_OPERATION_MIN,
                O::Max => sys::WGB_BLEND_OPERATION_MAX,
            },
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// device  (device.rs:82-308)
// ---------------------------------------------------------------------------------------------------------------

#[derive(Debug)]
pub struct Device {
    handle: Handle<sys::wgb_device_t>,
    features: wgpu::Features,
    limits: wgpu::Limits,
}

fn custom<'a, T: 'static>(object: Option<&'a T>, what: &str) -> &'a T {
    object.unwrap_or_else(|| panic!("wgpu-b200: {what} was created by another backend"))
}

impl DeviceInterface for Device {
    fn features(&self) -> wgpu::Features {
        self.features
    }

    fn limits(&self) -> wgpu::Limits {
        self.limits.clone()
    }

    fn create_shader_module(&self, desc: wgpu::ShaderModuleDescriptor<'_>, _checks: wgpu::ShaderRuntimeChecks) -> DispatchShaderModule {
        // WGSL only, like the reference (shader.rs:35-38).  The library parses, validates and translates the module
        // to CUDA C++ when a pipeline names its entry points; errors surface there (device.rs:133 unwraps too).
        let wgpu::ShaderSource::Wgsl(source) = desc.source else { panic!("wgpu-b200: only WGSL shader sources are supported") };
        let wgsl = CString::new(source.as_bytes()).expect("WGSL source contains a NUL byte");
        let d = sys::wgb_shader_module_descriptor { wgsl: wgsl.as_ptr(), emitted_count: 0, emitted: ptr::null() };
        let handle = create!(sys::wgb_shader_module_t, |out| sys::wgb_device_create_shader_module(self.handle.raw(), &d, &mut out));
        DispatchShaderModule::custom(ShaderModule { handle })
    }

    unsafe fn create_shader_module_passthrough(&self, _desc: &wgpu::ShaderModuleDescriptorPassthrough<'_>) -> DispatchShaderModule {
        todo!()
    }

    fn create_bind_group_layout(&self, desc: &wgpu::BindGroupLayoutDescriptor<'_>) -> DispatchBindGroupLayout {
        let entries: Vec<sys::wgb_bind_group_layout_entry> = desc
            .entries
            .iter()
            .map(|e| {
                let (kind, dynamic) = match e.ty {
                    wgpu::BindingType::Buffer { has_dynamic_offset, .. } => (sys::WGB_BINDING_BUFFER, has_dynamic_offset),
                    wgpu::BindingType::Texture { .. } => (sys::WGB_BINDING_TEXTURE_VIEW, false),
                    wgpu::BindingType::Sampler(_) => (sys::WGB_BINDING_SAMPLER, false),
                    other => panic!("wgpu-b200: binding type {other:?} is not supported"),
                };
                let mut visibility = 0;
                if e.visibility.contains(wgpu::ShaderStages::VERTEX) {
                    visibility |= sys::WGB_SHADER_STAGE_VERTEX;
                }
                if e.visibility.contains(wgpu::ShaderStages::FRAGMENT) {
                    visibility |= sys::WGB_SHADER_STAGE_FRAGMENT;
                }
                sys::wgb_bind_group_layout_entry { binding: e.binding, visibility, kind, has_dynamic_offset: dynamic as u32 }
            })
            .collect();
        let handle = create!(sys::wgb_bind_group_layout_t, |out| sys::wgb_device_create_bind_group_layout(
            self.handle.raw(), entries.as_ptr(), entries.len() as u32, &mut out));
        DispatchBindGroupLayout::custom(BindGroupLayout { handle })
    }

    fn create_bind_group(&self, desc: &wgpu::BindGroupDescriptor<'_>) -> DispatchBindGroup {
        let layout = custom(desc.layout.as_custom::<BindGroupLayout>(), "bind group layout");
        let entries: Vec<sys::wgb_bind_group_entry> = desc
            .entries
            .iter()
            .map(|e| {
                let mut out = sys::wgb_bind_group_entry {
                    binding: e.binding,
                    kind: 0,
                    buffer: ptr::null_mut(),
                    offset: 0,
                    size: sys::WGB_WHOLE_SIZE,
                    texture_view: ptr::null_mut(),
                    sampler: ptr::null_mut(),
                };
                match &e.resource {
                    wgpu::BindingResource::Buffer(b) => {
                        out.kind = sys::WGB_BINDING_BUFFER;
                        out.buffer = custom(b.buffer.as_custom::<Buffer>(), "buffer").handle.raw();
                        out.offset = b.offset;
                        out.size = b.size.map_or(sys::WGB_WHOLE_SIZE, |s| s.get());
                    }
                    wgpu::BindingResource::TextureView(v) => {
                        out.kind = sys::WGB_BINDING_TEXTURE_VIEW;
                        out.texture_view = custom(v.as_custom::<TextureView>(), "texture view").handle.raw();
                    }
                    wgpu::BindingResource::Sampler(s) => {
                        out.kind = sys::WGB_BINDING_SAMPLER;
                        out.sampler = custom(s.as_custom::<Sampler>(), "sampler").handle.raw();
                    }
                    _ => todo!("binding arrays, acceleration structures and external textures (bind_group.rs:100-110)"),
                }
                out
            })
            .collect();
        let handle = create!(sys::wgb_bind_group_t, |out| sys::wgb_device_create_bind_group(
            self.handle.raw(), layout.handle.raw(), entries.as_ptr(), entries.len() as u32, &mut out));
        DispatchBindGroup::custom(BindGroup { handle })
    }

    fn create_pipeline_layout(&self, desc: &wgpu::PipelineLayoutDescriptor<'_>) -> DispatchPipelineLayout {
        let groups: Vec<BindGroupLayout> =
            desc.bind_group_layouts.iter().map(|l| custom(l.as_custom::<BindGroupLayout>(), "bind group layout").clone()).collect();
        let raw: Vec<sys::wgb_bind_group_layout> = groups.iter().map(|l| l.handle.raw()).collect();
        let handle = create!(sys::wgb_pipeline_layout_t, |out| sys::wgb_device_create_pipeline_layout(
            self.handle.raw(), raw.as_ptr(), raw.len() as u32, &mut out));
        DispatchPipelineLayout::custom(PipelineLayout { handle, groups: Arc::new(groups) })
    }

    fn create_render_pipeline(&self, desc: &wgpu::RenderPipelineDescriptor<'_>) -> DispatchRenderPipeline {
        let layout = desc.layout.map(|l| custom(l.as_custom::<PipelineLayout>(), "pipeline layout").clone());
        let vs_module = custom(desc.vertex.module.as_custom::<ShaderModule>(), "shader module");
        let vs_entry = desc.vertex.entry_point.map(|e| CString::new(e).unwrap());
        let attributes: Vec<Vec<sys::wgb_vertex_attribute>> = desc
            .vertex
            .buffers
            .iter()
            .map(|b| {
                b.attributes
                    .iter()
                    .map(|a| sys::wgb_vertex_attribute { format: convert::vertex_format(a.format), offset: a.offset, shader_location: a.shader_location })
                    .collect()
            })
            .collect();
        let buffers: Vec<sys::wgb_vertex_buffer_layout> = desc
            .vertex
            .buffers
            .iter()
            .zip(&attributes)
            .map(|(b, a)| sys::wgb_vertex_buffer_layout {
                array_stride: b.array_stride,
                step_mode: match b.step_mode {
                    wgpu::VertexStepMode::Vertex => sys::WGB_VERTEX_STEP_MODE_VERTEX,
                    wgpu::VertexStepMode::Instance => sys::WGB_VERTEX_STEP_MODE_INSTANCE,
                },
                attribute_count: a.len() as u32,
                attributes: a.as_ptr(),
            })
            .collect();
        let fs_module = desc.fragment.as_ref().map(|f| custom(f.module.as_custom::<ShaderModule>(), "shader module"));
        let fs_entry = desc.fragment.as_ref().and_then(|f| f.entry_point).map(|e| CString::new(e).unwrap());
        let targets: Vec<sys::wgb_color_target_state> = desc
            .fragment
            .as_ref()
            .map(|f| {
                f.targets
                    .iter()
                    .map(|t| {
                        let t = t.as_ref().expect("wgpu-b200: empty colour target slots are not supported");
                        let none = wgpu::BlendComponent::REPLACE;
                        sys::wgb_color_target_state {
                            format: convert::texture_format(t.format).unwrap_or_else(|| panic!("Unsupported texture format: {:?}", t.format)),
                            has_blend: t.blend.is_some() as u32,
                            write_mask: t.write_mask.bits(),
                            blend_color: convert::blend_component(t.blend.as_ref().map_or(&none, |b| &b.color)),
                            blend_alpha: convert::blend_component(t.blend.as_ref().map_or(&none, |b| &b.alpha)),
                        }
                    })
                    .collect()
            })
            .unwrap_or_default();
        let p = &desc.primitive;
        let ds = desc.depth_stencil.as_ref();
        let d = sys::wgb_render_pipeline_descriptor {
            layout: layout.as_ref().map_or(ptr::null_mut(), |l| l.handle.raw()),
            vertex_module: vs_module.handle.raw(),
            vertex_entry_point: vs_entry.as_ref().map_or(ptr::null(), |e| e.as_ptr()),
            vertex_buffer_count: buffers.len() as u32,
            vertex_buffers: buffers.as_ptr(),
            topology: convert::topology(p.topology),
            strip_index_format: convert::index_format(p.strip_index_format),
            front_face: match p.front_face {
                wgpu::FrontFace::Ccw => sys::WGB_FRONT_FACE_CCW,
                wgpu::FrontFace::Cw => sys::WGB_FRONT_FACE_CW,
            },
            cull_mode: match p.cull_mode {
                None => sys::WGB_CULL_MODE_NONE,
                Some(wgpu::Face::Front) => sys::WGB_CULL_MODE_FRONT,
                Some(wgpu::Face::Back) => sys::WGB_CULL_MODE_BACK,
            },
            polygon_mode: match p.polygon_mode {
                wgpu::PolygonMode::Fill => sys::WGB_POLYGON_MODE_FILL,
                wgpu::PolygonMode::Line => sys::WGB_POLYGON_MODE_LINE,
                wgpu::PolygonMode::Point => sys::WGB_POLYGON_MODE_POINT,
            },
            unclipped_depth: p.unclipped_depth as u32,
            conservative: p.conservative as u32,
            has_depth_stencil: ds.is_some() as u32,
            depth_format: ds.map_or(0, |d| convert::texture_format(d.format).unwrap_or_else(|| panic!("Unsupported depth format: {:?}", d.format))),
            depth_write_enabled: ds.map_or(0, |d| d.depth_write_enabled as u32),
            depth_compare: ds.map_or(sys::WGB_COMPARE_ALWAYS, |d| convert::compare(d.depth_compare)),
            multisample_count: desc.multisample.count,
            fragment_module: fs_module.map_or(ptr::null_mut(), |m| m.handle.raw()),
            fragment_entry_point: fs_entry.as_ref().map_or(ptr::null(), |e| e.as_ptr()),
            target_count: targets.len() as u32,
            targets: targets.as_ptr(),
        };
        // translates both entry points to CUDA C++ and compiles the pipeline's kernels with NVRTC for sm_100a
        let handle = create!(sys::wgb_render_pipeline_t, |out| sys::wgb_device_create_render_pipeline(self.handle.raw(), &d, &mut out));
        DispatchRenderPipeline::custom(RenderPipeline { handle, layout })
    }

    fn create_mesh_pipeline(&self, _desc: &wgpu::MeshPipelineDescriptor<'_>) -> DispatchRenderPipeline {
        todo!()
    }

    fn create_compute_pipeline(&self, _desc: &wgpu::ComputePipelineDescriptor<'_>) -> DispatchComputePipeline {
        todo!("the reference has no compute stage (device.rs:143-148)")
    }

    unsafe fn create_pipeline_cache(&self, _desc: &wgpu::PipelineCacheDescriptor<'_>) -> DispatchPipelineCache {
        todo!()
    }

    fn create_buffer(&self, desc: &wgpu::BufferDescriptor<'_>) -> DispatchBuffer {
        // the reference maps every new buffer for writing whatever the descriptor says (device.rs:157-159); here
        // `mapped_at_creation` is honoured, which is what `create_buffer_init` relies on either way
        let d = sys::wgb_buffer_descriptor { size: desc.size, usage: desc.usage.bits(), mapped_at_creation: desc.mapped_at_creation as u32 };
        let handle = create!(sys::wgb_buffer_t, |out| sys::wgb_device_create_buffer(self.handle.raw(), &d, &mut out));
        DispatchBuffer::custom(Buffer { handle, size: desc.size })
    }

    fn create_texture(&self, desc: &wgpu::TextureDescriptor<'_>) -> DispatchTexture {
        let format = convert::texture_format(desc.format).unwrap_or_else(|| panic!("Unsupported texture format: {:?}", desc.format));
        let d = sys::wgb_texture_descriptor {
            width: desc.size.width,
            height: desc.size.height,
            depth_or_array_layers: desc.size.depth_or_array_layers,
            mip_level_count: desc.mip_level_count,
            sample_count: desc.sample_count,
            format,
            usage: desc.usage.bits(),
        };
        let handle = create!(sys::wgb_texture_t, |out| sys::wgb_device_create_texture(self.handle.raw(), &d, &mut out));
        DispatchTexture::custom(Texture { handle, size: desc.size, format: desc.format })
    }

    fn create_external_texture(&self, _desc: &wgpu::ExternalTextureDescriptor<'_>, _planes: &[&wgpu::TextureView]) -> DispatchExternalTexture {
        todo!()
    }

    fn create_blas(&self, _desc: &wgpu::CreateBlasDescriptor<'_>, _sizes: wgpu::BlasGeometrySizeDescriptors) -> (Option<u64>, DispatchBlas) {
        todo!()
    }

    fn create_tlas(&self, _desc: &wgpu::CreateTlasDescriptor<'_>) -> DispatchTlas {
        todo!()
    }

    fn create_sampler(&self, desc: &wgpu::SamplerDescriptor<'_>) -> DispatchSampler {
        let d = sys::wgb_sampler_descriptor {
            address_mode_u: convert::address_mode(desc.address_mode_u),
            address_mode_v: convert::address_mode(desc.address_mode_v),
            address_mode_w: convert::address_mode(desc.address_mode_w),
            mag_filter: convert::filter(desc.mag_filter),
            min_filter: convert::filter(desc.min_filter),
            mipmap_filter: match desc.mipmap_filter {
                wgpu::MipmapFilterMode::Nearest => sys::WGB_FILTER_MODE_NEAREST,
                wgpu::MipmapFilterMode::Linear => sys::WGB_FILTER_MODE_LINEAR,
            },
        };
        let handle = create!(sys::wgb_sampler_t, |out| sys::wgb_device_create_sampler(self.handle.raw(), &d, &mut out));
        DispatchSampler::custom(Sampler { handle })
    }

    fn create_query_set(&self, _desc: &wgpu::QuerySetDescriptor<'_>) -> DispatchQuerySet {
        todo!()
    }

    fn create_command_encoder(&self, _desc: &wgpu::CommandEncoderDescriptor<'_>) -> DispatchCommandEncoder {
        let handle = create!(sys::wgb_command_encoder_t, |out| sys::wgb_device_create_command_encoder(self.handle.raw(), &mut out));
        DispatchCommandEncoder::custom(CommandEncoder { handle })
    }

    fn create_render_bundle_encoder(&self, _desc: &wgpu::RenderBundleEncoderDescriptor<'_>) -> DispatchRenderBundleEncoder {
        todo!()
    }

    fn set_device_lost_callback(&self, _callback: BoxDeviceLostCallback) {}

    fn on_uncaptured_error(&self, _handler: Arc<dyn wgpu::UncapturedErrorHandler>) {
        todo!()
    }

    fn push_error_scope(&self, _filter: wgpu::ErrorFilter) -> u32 {
        todo!()
    }

    fn pop_error_scope(&self, _index: u32) -> Pin<Box<dyn PopErrorScopeFuture>> {
        todo!()
    }

    unsafe fn start_graphics_debugger_capture(&self) {
        todo!()
    }

    unsafe fn stop_graphics_debugger_capture(&self) {
        todo!()
    }

    fn poll(&self, poll_type: wgpu::wgt::PollType<u64>) -> Result<wgpu::PollStatus, wgpu::PollError> {
        // device.rs:237-295.  An error raised while a submission executed surfaces here, where the reference's
        // engine-thread panic is observed (device.rs:498-503).
        let (wait, index, timeout_ns) = match poll_type {
            wgpu::wgt::PollType::Poll => (0, sys::WGB_SUBMISSION_ANY, 0),
            wgpu::wgt::PollType::Wait { submission_index, timeout } => (
                1,
                submission_index.unwrap_or(sys::WGB_SUBMISSION_ANY),
                timeout.map_or(u64::MAX, |t| t.as_nanos().min((u64::MAX - 1) as u128) as u64),
            ),
        };
        let mut outcome = sys::WGB_POLL_OK;
        check(unsafe { sys::wgb_device_poll(self.handle.raw(), wait, index, timeout_ns, &mut outcome) });
        match (wait, outcome) {
            (_, sys::WGB_POLL_TIMEOUT) => Err(wgpu::PollError::Timeout),
            (0, _) => Ok(wgpu::PollStatus::Poll),
            (_, sys::WGB_POLL_QUEUE_EMPTY) => Ok(wgpu::PollStatus::QueueEmpty),
            _ => Ok(wgpu::PollStatus::WaitSucceeded),
        }
    }

    fn get_internal_counters(&self) -> wgpu::InternalCounters {
        todo!()
    }

    fn generate_allocator_report(&self) -> Option<wgpu::AllocatorReport> {
        todo!()
    }

    fn destroy(&self) {}
}

impl Device {
    /// Counters and device timings of the last executed render pass (`wgb_pass_stats`).
    pub fn last_pass_stats(device: &wgpu::Device) -> sys::wgb_pass_stats {
        let device = custom(device.as_custom::<Device>(), "device");
        let mut stats: sys::wgb_pass_stats = unsafe { std::mem::zeroed() };
        check(unsafe { sys::wgb_device_get_last_pass_stats(device.handle.raw(), &mut stats) });
        stats
    }
}

// ---------------------------------------------------------------------------------------------------------------
// queue  (device.rs:331-478)
// ---------------------------------------------------------------------------------------------------------------

#[derive(Debug)]
pub struct Queue {
    handle: Handle<sys::wgb_queue_t>,
}

impl QueueInterface for Queue {
    fn write_buffer(&self, buffer: &DispatchBuffer, offset: wgpu::BufferAddress, data: &[u8]) {
        let buffer = custom(buffer.as_custom::<Buffer>(), "buffer");
        check(unsafe { sys::wgb_queue_write_buffer(self.handle.raw(), buffer.handle.raw(), offset, data.as_ptr().cast(), data.len() as u64) });
    }

    fn create_staging_buffer(&self, _size: wgpu::BufferSize) -> Option<DispatchQueueWriteBuffer> {
        todo!()
    }

    fn validate_write_buffer(&self, _buffer: &DispatchBuffer, _offset: wgpu::BufferAddress, _size: wgpu::BufferSize) -> Option<()> {
        todo!()
    }

    fn write_staging_buffer(&self, _buffer: &DispatchBuffer, _offset: wgpu::BufferAddress, _staging: &DispatchQueueWriteBuffer) {
        todo!()
    }

    fn write_texture(&self, texture: wgpu::TexelCopyTextureInfo<'_>, data: &[u8], layout: wgpu::TexelCopyBufferLayout, size: wgpu::Extent3d) {
        if texture.mip_level != 0 {
            todo!("write to mip_level: {}", texture.mip_level);
        }
        assert!(size.depth_or_array_layers == 1, "wgpu-b200: write_texture writes one layer at a time");
        let dst = custom(texture.texture.as_custom::<Texture>(), "texture");
        let offset = usize::try_from(layout.offset).expect("source offset overflow");
        let data = &data[offset..];
        // unlike the reference (device.rs:387-388, which copies the whole texture), the rectangle is honoured
        let bytes_per_row = layout.bytes_per_row.unwrap_or(data.len() as u32);
        check(unsafe {
            sys::wgb_queue_write_texture(self.handle.raw(), dst.handle.raw(), texture.origin.x, texture.origin.y, data.as_ptr().cast(),
                                         data.len() as u64, bytes_per_row, size.width, size.height)
        });
    }

    fn submit(&self, command_buffers: &mut dyn Iterator<Item = DispatchCommandBuffer>) -> u64 {
        let owned: Vec<DispatchCommandBuffer> = command_buffers.collect();
        let raw: Vec<sys::wgb_command_buffer> =
            owned.iter().map(|c| custom(c.as_custom::<CommandBuffer>(), "command buffer").handle.raw()).collect();
        let mut index = 0u64;
        check(unsafe { sys::wgb_queue_submit(self.handle.raw(), raw.as_ptr(), raw.len() as u32, &mut index) });
        index
    }

    fn get_timestamp_period(&self) -> f32 {
        todo!()
    }

    fn on_submitted_work_done(&self, _callback: BoxSubmittedWorkDoneCallback) {
        todo!()
    }

    fn compact_blas(&self, _blas: &DispatchBlas) -> (Option<u64>, DispatchBlas) {
        todo!()
    }
}

// ---------------------------------------------------------------------------------------------------------------
// buffers  (buffer.rs:111-180, 441-456)
// ---------------------------------------------------------------------------------------------------------------

#[derive(Clone, Debug)]
pub struct Buffer {
    handle: Handle<sys::wgb_buffer_t>,
    size: u64,
}

impl BufferInterface for Buffer {
    fn map_async(&self, mode: wgpu::MapMode, range: Range<wgpu::BufferAddress>, callback: BufferMapCallback) {
        let mode = match mode {
            wgpu::MapMode::Read => sys::WGB_MAP_MODE_READ,
            wgpu::MapMode::Write => sys::WGB_MAP_MODE_WRITE,
        };
        // the library maps before it returns (device work on the buffer is waited for, a read copies the bytes back),
        // so the callback runs here -- the reference drops it uncalled in this branch (buffer.rs:127-129)
        let status = unsafe { sys::wgb_buffer_map_async(self.handle.raw(), mode, range.start, range.end - range.start, None, ptr::null_mut()) };
        check(status);
        callback(Ok(()));
    }

    #[track_caller]
    fn get_mapped_range(&self, range: Range<wgpu::BufferAddress>) -> DispatchBufferMappedRange {
        let end = range.end.min(self.size);
        let mut p: *mut std::os::raw::c_void = ptr::null_mut();
        check(unsafe { sys::wgb_buffer_get_mapped_range(self.handle.raw(), range.start, end - range.start, &mut p) });
        DispatchBufferMappedRange::custom(BufferMappedRange { _buffer: self.clone(), ptr: p.cast(), len: (end - range.start) as usize })
    }

    fn unmap(&self) {
        check(unsafe { sys::wgb_buffer_unmap(self.handle.raw()) });
    }

    fn destroy(&self) {}
}

/// A view of the buffer's host staging copy; it stays valid until `unmap`, and holds the buffer alive.
#[derive(Debug)]
pub struct BufferMappedRange {
    _buffer: Buffer,
    ptr: *mut u8,
    len: usize,
}
unsafe impl Send for BufferMappedRange {}
unsafe impl Sync for BufferMappedRange {}

impl BufferMappedRangeInterface for BufferMappedRange {
    fn slice(&self) -> &[u8] {
        unsafe { std::slice::from_raw_parts(self.ptr, self.len) }
    }

    fn slice_mut(&mut self) -> &mut [u8] {
        unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// textures, samplers, shader modules, layouts, bind groups, pipelines
// ---------------------------------------------------------------------------------------------------------------

#[derive(Clone, Debug)]
pub struct Texture {
    handle: Handle<sys::wgb_texture_t>,
    size: wgpu::Extent3d,
    format: wgpu::TextureFormat,
}

impl TextureInterface for Texture {
    fn create_view(&self, desc: &wgpu::TextureViewDescriptor<'_>) -> DispatchTextureView {
        let d = sys::wgb_texture_view_descriptor { base_array_layer: desc.base_array_layer, reserved: 0 };
        let handle = create!(sys::wgb_texture_view_t, |out| sys::wgb_texture_create_view(self.handle.raw(), &d, &mut out));
        DispatchTextureView::custom(TextureView { handle })
    }

    fn destroy(&self) {}
}

#[derive(Clone, Debug)]
pub struct TextureView {
    handle: Handle<sys::wgb_texture_view_t>,
}
impl TextureViewInterface for TextureView {}

#[derive(Clone, Debug)]
pub struct Sampler {
    handle: Handle<sys::wgb_sampler_t>,
}
impl SamplerInterface for Sampler {}

#[derive(Clone, Debug)]
pub struct ShaderModule {
    handle: Handle<sys::wgb_shader_module_t>,
}
impl ShaderModuleInterface for ShaderModule {
    fn get_compilation_info(&self) -> Pin<Box<dyn ShaderCompilationInfoFuture>> {
        Box::pin(async { wgpu::CompilationInfo { messages: Vec::new() } })
    }
}

#[derive(Clone, Debug)]
pub struct BindGroupLayout {
    handle: Handle<sys::wgb_bind_group_layout_t>,
}
impl BindGroupLayoutInterface for BindGroupLayout {}

#[derive(Clone, Debug)]
pub struct PipelineLayout {
    handle: Handle<sys::wgb_pipeline_layout_t>,
    groups: Arc<Vec<BindGroupLayout>>,
}
impl PipelineLayoutInterface for PipelineLayout {}

#[derive(Clone, Debug)]
pub struct BindGroup {
    handle: Handle<sys::wgb_bind_group_t>,
}
impl BindGroupInterface for BindGroup {}

#[derive(Clone, Debug)]
pub struct RenderPipeline {
    handle: Handle<sys::wgb_render_pipeline_t>,
    layout: Option<PipelineLayout>,
}
impl RenderPipelineInterface for RenderPipeline {
    fn get_bind_group_layout(&self, index: u32) -> DispatchBindGroupLayout {
        match &self.layout {
            Some(layout) => DispatchBindGroupLayout::custom(layout.groups[index as usize].clone()),
            None => todo!("layout: None (pipeline.rs:85-87)"),
        }
    }
}

impl RenderPipeline {
    /// The CUDA translation unit NVRTC compiled for this pipeline (state `#define`s + emitted shaders + kernels).
    pub fn cuda_source(pipeline: &wgpu::RenderPipeline) -> String {
        let pipeline = custom(pipeline.as_custom::<RenderPipeline>(), "render pipeline");
        let mut text: *mut std::os::raw::c_char = ptr::null_mut();
        check(unsafe { sys::wgb_render_pipeline_get_source(pipeline.handle.raw(), &mut text) });
        let out = unsafe { CStr::from_ptr(text) }.to_string_lossy().into_owned();
        unsafe { sys::wgb_free(text.cast()) };
        out
    }
}

// ---------------------------------------------------------------------------------------------------------------
// command encoder  (command.rs:36-171; the copies and clears are todo!() there and live here)
// ---------------------------------------------------------------------------------------------------------------

#[derive(Clone, Debug)]
pub struct CommandEncoder {
    handle: Handle<sys::wgb_command_encoder_t>,
}

fn texture_info(info: &wgpu::TexelCopyTextureInfo<'_>) -> sys::wgb_texel_copy_texture_info {
    assert!(info.mip_level == 0, "wgpu-b200: textures have one mip level");
    sys::wgb_texel_copy_texture_info {
        texture: custom(info.texture.as_custom::<Texture>(), "texture").handle.raw(),
        x: info.origin.x,
        y: info.origin.y,
        layer: info.origin.z,
    }
}
fn buffer_info(info: &wgpu::TexelCopyBufferInfo<'_>) -> sys::wgb_texel_copy_buffer_info {
    sys::wgb_texel_copy_buffer_info {
        buffer: custom(info.buffer.as_custom::<Buffer>(), "buffer").handle.raw(),
        offset: info.layout.offset,
        bytes_per_row: info.layout.bytes_per_row.unwrap_or(0),
        rows_per_image: info.layout.rows_per_image.unwrap_or(0),
    }
}

impl CommandEncoderInterface for CommandEncoder {
    fn copy_buffer_to_buffer(&self, source: &DispatchBuffer, source_offset: wgpu::BufferAddress, destination: &DispatchBuffer,
                             destination_offset: wgpu::BufferAddress, copy_size: Option<wgpu::BufferAddress>) {
        let (src, dst) = (custom(source.as_custom::<Buffer>(), "buffer"), custom(destination.as_custom::<Buffer>(), "buffer"));
        check(unsafe {
            sys::wgb_command_encoder_copy_buffer_to_buffer(self.handle.raw(), src.handle.raw(), source_offset, dst.handle.raw(),
                                                           destination_offset, copy_size.unwrap_or(sys::WGB_WHOLE_SIZE))
        });
    }

    fn copy_buffer_to_texture(&self, source: wgpu::TexelCopyBufferInfo<'_>, destination: wgpu::TexelCopyTextureInfo<'_>, copy_size: wgpu::Extent3d) {
        let (s, d) = (buffer_info(&source), texture_info(&destination));
        check(unsafe { sys::wgb_command_encoder_copy_buffer_to_texture(self.handle.raw(), &s, &d, copy_size.width, copy_size.height) });
    }

    fn copy_texture_to_buffer(&self, source: wgpu::TexelCopyTextureInfo<'_>, destination: wgpu::TexelCopyBufferInfo<'_>, copy_size: wgpu::Extent3d) {
        let (s, d) = (texture_info(&source), buffer_info(&destination));
        check(unsafe { sys::wgb_command_encoder_copy_texture_to_buffer(self.handle.raw(), &s, &d, copy_size.width, copy_size.height) });
    }

    fn copy_texture_to_texture(&self, source: wgpu::TexelCopyTextureInfo<'_>, destination: wgpu::TexelCopyTextureInfo<'_>, copy_size: wgpu::Extent3d) {
        let (s, d) = (texture_info(&source), texture_info(&destination));
        check(unsafe { sys::wgb_command_encoder_copy_texture_to_texture(self.handle.raw(), &s, &d, copy_size.width, copy_size.height) });
    }

    fn begin_compute_pass(&self, _desc: &wgpu::ComputePassDescriptor<'_>) -> DispatchComputePass {
        todo!("the reference has no compute stage (command.rs:75-80)")
    }

    fn begin_render_pass(&self, desc: &wgpu::RenderPassDescriptor<'_>) -> DispatchRenderPass {
        let colors: Vec<sys::wgb_color_attachment> = desc
            .color_attachments
            .iter()
            .map(|a| {
                let a = a.as_ref().expect("wgpu-b200: empty colour attachment slots are not supported");
                assert!(a.resolve_target.is_none(), "resolve targets are not supported (fragment.rs:525-527 todo!)");
                let (load_op, clear_value) = match a.ops.load {
                    wgpu::LoadOp::Clear(c) => (sys::WGB_LOAD_OP_CLEAR, [c.r, c.g, c.b, c.a]),
                    wgpu::LoadOp::Load => (sys::WGB_LOAD_OP_LOAD, [0.0; 4]),
                    _ => todo!("LoadOp::DontCare"),
                };
                sys::wgb_color_attachment {
                    view: custom(a.view.as_custom::<TextureView>(), "texture view").handle.raw(),
                    load_op,
                    store_op: match a.ops.store {
                        wgpu::StoreOp::Store => sys::WGB_STORE_OP_STORE,
                        wgpu::StoreOp::Discard => sys::WGB_STORE_OP_DISCARD,
                    },
                    clear_value,
                }
            })
            .collect();
        let depth = desc.depth_stencil_attachment.as_ref().map(|a| {
            let (load, clear) = match a.depth_ops.as_ref().map(|o| o.load) {
                Some(wgpu::LoadOp::Clear(v)) => (sys::WGB_LOAD_OP_CLEAR, v),
                _ => (sys::WGB_LOAD_OP_LOAD, 0.0),
            };
            sys::wgb_depth_stencil_attachment {
                view: custom(a.view.as_custom::<TextureView>(), "texture view").handle.raw(),
                has_depth_ops: a.depth_ops.is_some() as u32,
                depth_load_op: load,
                depth_store_op: match a.depth_ops.as_ref().map(|o| o.store) {
                    Some(wgpu::StoreOp::Discard) => sys::WGB_STORE_OP_DISCARD,
                    _ => sys::WGB_STORE_OP_STORE,
                },
                depth_clear_value: clear,
                has_stencil_ops: a.stencil_ops.is_some() as u32,
            }
        });
        let d = sys::wgb_render_pass_descriptor {
            color_attachment_count: colors.len() as u32,
            color_attachments: colors.as_ptr(),
            depth_stencil_attachment: depth.as_ref().map_or(ptr::null(), |d| d as *const _),
        };
        let handle = create!(sys::wgb_render_pass_t, |out| sys::wgb_command_encoder_begin_render_pass(self.handle.raw(), &d, &mut out));
        DispatchRenderPass::custom(RenderPass { handle, ended: false })
    }

    fn finish(&mut self) -> DispatchCommandBuffer {
        let handle = create!(sys::wgb_command_buffer_t, |out| sys::wgb_command_encoder_finish(self.handle.raw(), &mut out));
        DispatchCommandBuffer::custom(CommandBuffer { handle })
    }

    fn clear_texture(&self, texture: &DispatchTexture, _range: &wgpu::ImageSubresourceRange) {
        let texture = custom(texture.as_custom::<Texture>(), "texture");
        check(unsafe { sys::wgb_command_encoder_clear_texture(self.handle.raw(), texture.handle.raw()) });
    }

    fn clear_buffer(&self, buffer: &DispatchBuffer, offset: wgpu::BufferAddress, size: Option<wgpu::BufferAddress>) {
        let buffer = custom(buffer.as_custom::<Buffer>(), "buffer");
        check(unsafe { sys::wgb_command_encoder_clear_buffer(self.handle.raw(), buffer.handle.raw(), offset, size.unwrap_or(sys::WGB_WHOLE_SIZE)) });
    }

    fn insert_debug_marker(&self, _label: &str) {}
    fn push_debug_group(&self, _label: &str) {}
    fn pop_debug_group(&self) {}

    fn write_timestamp(&self, _query_set: &DispatchQuerySet, _query_index: u32) {
        todo!()
    }

    fn resolve_query_set(&self, _query_set: &DispatchQuerySet, _first_query: u32, _query_count: u32, _destination: &DispatchBuffer,
                         _destination_offset: wgpu::BufferAddress) {
        todo!()
    }

    fn mark_acceleration_structures_built<'a>(&self, _blas: &mut dyn Iterator<Item = &'a wgpu::Blas>, _tlas: &mut dyn Iterator<Item = &'a wgpu::Tlas>) {
        todo!()
    }

    fn build_acceleration_structures<'a>(&self, _blas: &mut dyn Iterator<Item = &'a wgpu::BlasBuildEntry<'a>>,
                                         _tlas: &mut dyn Iterator<Item = &'a wgpu::Tlas>) {
        todo!()
    }

    fn transition_resources<'a>(&mut self, _buffers: &mut dyn Iterator<Item = wgpu::wgt::BufferTransition<&'a DispatchBuffer>>,
                                _textures: &mut dyn Iterator<Item = wgpu::wgt::TextureTransition<&'a DispatchTexture>>) {
        // nothing to do: submissions execute in order on one CUDA stream per device
    }
}

#[derive(Debug)]
pub struct CommandBuffer {
    handle: Handle<sys::wgb_command_buffer_t>,
}
impl CommandBufferInterface for CommandBuffer {}

// ---------------------------------------------------------------------------------------------------------------
// render pass  (render_pass/mod.rs:73-329): every live method records through the library
// ---------------------------------------------------------------------------------------------------------------

#[derive(Debug)]
pub struct RenderPass {
    handle: Handle<sys::wgb_render_pass_t>,
    ended: bool,
}

impl RenderPassInterface for RenderPass {
    fn set_pipeline(&mut self, pipeline: &DispatchRenderPipeline) {
        let pipeline = custom(pipeline.as_custom::<RenderPipeline>(), "render pipeline");
        check(unsafe { sys::wgb_render_pass_set_pipeline(self.handle.raw(), pipeline.handle.raw()) });
    }

    fn set_bind_group(&mut self, index: u32, bind_group: Option<&DispatchBindGroup>, offsets: &[wgpu::DynamicOffset]) {
        let group = bind_group.map_or(ptr::null_mut(), |g| custom(g.as_custom::<BindGroup>(), "bind group").handle.raw());
        check(unsafe { sys::wgb_render_pass_set_bind_group(self.handle.raw(), index, group, offsets.as_ptr(), offsets.len() as u32) });
    }

    fn set_index_buffer(&mut self, buffer: &DispatchBuffer, index_format: wgpu::IndexFormat, offset: wgpu::BufferAddress, size: Option<wgpu::BufferSize>) {
        let buffer = custom(buffer.as_custom::<Buffer>(), "buffer");
        check(unsafe {
            sys::wgb_render_pass_set_index_buffer(self.handle.raw(), buffer.handle.raw(), convert::index_format(Some(index_format)), offset,
                                                  size.map_or(sys::WGB_WHOLE_SIZE, |s| s.get()))
        });
    }

    fn set_vertex_buffer(&mut self, slot: u32, buffer: &DispatchBuffer, offset: wgpu::BufferAddress, size: Option<wgpu::BufferSize>) {
        let buffer = custom(buffer.as_custom::<Buffer>(), "buffer");
        check(unsafe {
            sys::wgb_render_pass_set_vertex_buffer(self.handle.raw(), slot, buffer.handle.raw(), offset, size.map_or(sys::WGB_WHOLE_SIZE, |s| s.get()))
        });
    }

    fn set_immediates(&mut self, _offset: u32, _data: &[u8]) {
        todo!()
    }

    fn set_blend_constant(&mut self, color: wgpu::Color) {
        let c = [color.r, color.g, color.b, color.a];
        check(unsafe { sys::wgb_render_pass_set_blend_constant(self.handle.raw(), c.as_ptr()) });
    }

    fn set_scissor_rect(&mut self, x: u32, y: u32, width: u32, height: u32) {
        check(unsafe { sys::wgb_render_pass_set_scissor_rect(self.handle.raw(), x, y, width, height) });
    }

    fn set_viewport(&mut self, x: f32, y: f32, width: f32, height: f32, min_depth: f32, max_depth: f32) {
        check(unsafe { sys::wgb_render_pass_set_viewport(self.handle.raw(), x, y, width, height, min_depth, max_depth) });
    }

    fn set_stencil_reference(&mut self, reference: u32) {
        check(unsafe { sys::wgb_render_pass_set_stencil_reference(self.handle.raw(), reference) });
    }

    fn draw(&mut self, vertices: Range<u32>, instances: Range<u32>) {
        check(unsafe {
            sys::wgb_render_pass_draw(self.handle.raw(), vertices.start, vertices.end - vertices.start, instances.start, instances.end - instances.start)
        });
    }

    fn draw_indexed(&mut self, indices: Range<u32>, base_vertex: i32, instances: Range<u32>) {
        check(unsafe {
            sys::wgb_render_pass_draw_indexed(self.handle.raw(), indices.start, indices.end - indices.start, base_vertex, instances.start,
                                              instances.end - instances.start)
        });
    }

    fn draw_mesh_tasks(&mut self, _x: u32, _y: u32, _z: u32) {
        todo!()
    }
    fn draw_indirect(&mut self, _buffer: &DispatchBuffer, _offset: wgpu::BufferAddress) {
        todo!()
    }
    fn draw_indexed_indirect(&mut self, _buffer: &DispatchBuffer, _offset: wgpu::BufferAddress) {
        todo!()
    }
    fn draw_mesh_tasks_indirect(&mut self, _buffer: &DispatchBuffer, _offset: wgpu::BufferAddress) {
        todo!()
    }
    fn multi_draw_indirect(&mut self, _buffer: &DispatchBuffer, _offset: wgpu::BufferAddress, _count: u32) {
        todo!()
    }
    fn multi_draw_indexed_indirect(&mut self, _buffer: &DispatchBuffer, _offset: wgpu::BufferAddress, _count: u32) {
        todo!()
    }
    fn multi_draw_indirect_count(&mut self, _buffer: &DispatchBuffer, _offset: wgpu::BufferAddress, _count_buffer: &DispatchBuffer,
                                 _count_offset: wgpu::BufferAddress, _max_count: u32) {
        todo!()
    }
    fn multi_draw_mesh_tasks_indirect(&mut self, _buffer: &DispatchBuffer, _offset: wgpu::BufferAddress, _count: u32) {
        todo!()
    }
    fn multi_draw_indexed_indirect_count(&mut self, _buffer: &DispatchBuffer, _offset: wgpu::BufferAddress, _count_buffer: &DispatchBuffer,
                                         _count_offset: wgpu::BufferAddress, _max_count: u32) {
        todo!()
    }
    fn multi_draw_mesh_tasks_indirect_count(&mut self, _buffer: &DispatchBuffer, _offset: wgpu::BufferAddress, _count_buffer: &DispatchBuffer,
                                            _count_offset: wgpu::BufferAddress, _max_count: u32) {
        todo!()
    }

    fn insert_debug_marker(&mut self, _label: &str) {}
    fn push_debug_group(&mut self, _label: &str) {}
    fn pop_debug_group(&mut self) {}

    fn write_timestamp(&mut self, _query_set: &DispatchQuerySet, _query_index: u32) {
        todo!()
    }
    fn begin_occlusion_query(&mut self, _query_index: u32) {
        todo!()
    }
    fn end_occlusion_query(&mut self) {
        todo!()
    }
    fn begin_pipeline_statistics_query(&mut self, _query_set: &DispatchQuerySet, _query_index: u32) {
        todo!()
    }
    fn end_pipeline_statistics_query(&mut self) {
        todo!()
    }
    fn execute_bundles(&mut self, _bundles: &mut dyn Iterator<Item = &DispatchRenderBundle>) {
        todo!()
    }

    fn end(&mut self) {
        if !std::mem::replace(&mut self.ended, true) {
            check(unsafe { sys::wgb_render_pass_end(self.handle.raw()) });
        }
    }
}

impl Drop for RenderPass {
    fn drop(&mut self) {
        // wgpu ends a pass by dropping it (render_pass/mod.rs:325-329)
        RenderPassInterface::end(self)
    }
}

// ---------------------------------------------------------------------------------------------------------------
// read-back helpers  (wgpu-cpu/src/lib.rs:111-173)
// ---------------------------------------------------------------------------------------------------------------

/// The texture's texels, row-major without padding (a device-to-host copy ordered after everything submitted).
pub fn read_texture(texture: &wgpu::Texture) -> Vec<u8> {
    let size = texture.size();
    let bpp = texture.format().block_copy_size(None).expect("texture format has no fixed texel size") as usize;
    let t = custom(texture.as_custom::<Texture>(), "texture");
    let mut out = vec![0u8; size.width as usize * size.height as usize * size.depth_or_array_layers as usize * bpp];
    check(unsafe { sys::wgb_texture_read(t.handle.raw(), out.as_mut_ptr().cast(), out.len() as u64) });
    out
}

/// `wgpu_cpu::dump_texture` for PNG output: written by the library itself (Rgba8 / Bgra8 targets as RGBA, Depth32Float
/// as 8-bit grey, like lib.rs:129-158).
pub fn dump_texture(texture: &wgpu::Texture, path: impl AsRef<std::path::Path>) {
    let t = custom(texture.as_custom::<Texture>(), "texture");
    let path = CString::new(path.as_ref().to_string_lossy().as_bytes()).expect("path contains a NUL byte");
    check(unsafe { sys::wgb_texture_dump_png(t.handle.raw(), path.as_ptr()) });
}

/// `wgpu_cpu::image::rgba_texture_image` (lib.rs:160-173); a copy rather than a view, the texels live in HBM.
#[cfg(feature = "image")]
pub fn rgba_texture_image(texture: &wgpu::Texture) -> image::ImageBuffer<image::Rgba<u8>, Vec<u8>> {
    assert!(matches!(texture.format(), wgpu::TextureFormat::Rgba8Unorm | wgpu::TextureFormat::Rgba8UnormSrgb));
    let size = texture.size();
    image::ImageBuffer::from_raw(size.width, size.height, read_texture(texture)).expect("texel buffer matches the texture size")
}

// ---------------------------------------------------------------------------------------------------------------
// multi-GPU presenter: the presenting rank exports its colour target, the others render into it
// ---------------------------------------------------------------------------------------------------------------

pub fn export_texture_ipc(texture: &wgpu::Texture) -> [u8; sys::WGB_IPC_HANDLE_SIZE] {
    let t = custom(texture.as_custom::<Texture>(), "texture");
    let mut handle = [0u8; sys::WGB_IPC_HANDLE_SIZE];
    check(unsafe { sys::wgb_texture_export_ipc(t.handle.raw(), handle.as_mut_ptr()) });
    handle
}

/// First and one-past-last pixel row of this device's band of a `height`-row target.
pub fn band_rows(device: &wgpu::Device, height: u32) -> Range<u32> {
    let d = custom(device.as_custom::<Device>(), "device");
    let (mut a, mut b) = (0u32, 0u32);
    check(unsafe { sys::wgb_device_get_band_rows(d.handle.raw(), height, &mut a, &mut b) });
    a..b
}

// ---------------------------------------------------------------------------------------------------------------
// surface / present  (surface.rs:24-198)
// ---------------------------------------------------------------------------------------------------------------

/// The library's surface (`wgb_surface_*`) presents into a host pixel sink; here the sink is the softbuffer window the
/// reference presents into.  `present` = wait for the frame, one device-to-host copy into the library's page-locked
/// window buffer, then the `on_present` callback below copies those bytes into softbuffer's buffer and presents it.
#[cfg(feature = "softbuffer")]
pub mod surface {
    use std::ffi::c_void;
    use std::num::NonZero;
    use std::ptr;
    use std::sync::Arc;

    use parking_lot::Mutex;
    use wgpu::custom::*;
    use wgpu_b200_sys as sys;

    use super::{check, Device, Handle, Texture};

    struct Window {
        _context: softbuffer::Context<Display>,
        surface: softbuffer::Surface<Display, WindowHandle>,
    }

    #[derive(Debug)]
    pub struct Surface {
        handle: Handle<sys::wgb_surface_t>,
        _instance: Handle<sys::wgb_instance_t>,
        #[allow(dead_code)]
        window: Arc<Mutex<Window>>,      // `user_data` of the present callback points into this allocation
        extent: Mutex<Option<(wgpu::Extent3d, wgpu::TextureFormat)>>,
    }

    impl std::fmt::Debug for Window {
        fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
            f.write_str("Window")
        }
    }

    /// `SurfaceOutputDetailInterface::present` of the reference (surface.rs:185-192): bytes as they are, then present.
    unsafe extern "C" fn on_present(user_data: *mut c_void, pixels: *const c_void, width: u32, height: u32, _bytes_per_row: u32) {
        let window = &*(user_data as *const Mutex<Window>);
        let mut window = window.lock();
        let mut target = window.surface.buffer_mut().unwrap();
        let source = std::slice::from_raw_parts(pixels as *const u8, width as usize * height as usize * 4);
        let bytes: &mut [u8] = bytemuck::cast_slice_mut(&mut *target);
        bytes.copy_from_slice(source);
        target.present().unwrap();
    }

    impl Surface {
        pub(super) fn new(instance: Handle<sys::wgb_instance_t>, target: wgpu::SurfaceTargetUnsafe) -> Result<Self, String> {
            match target {
                wgpu::SurfaceTargetUnsafe::RawHandle { raw_display_handle, raw_window_handle } => {
                    let context = softbuffer::Context::new(Display::from(raw_display_handle)).map_err(|e| e.to_string())?;
                    let surface = softbuffer::Surface::new(&context, WindowHandle::from(raw_window_handle)).map_err(|e| e.to_string())?;
                    let window = Arc::new(Mutex::new(Window { _context: context, surface }));
                    let sink = sys::wgb_surface_target { on_present: Some(on_present), user_data: Arc::as_ptr(&window) as *mut c_void };
                    let handle = create_surface(&instance, &sink);
                    Ok(Surface { handle, _instance: instance, window, extent: Mutex::new(None) })
                }
                _ => Err("Surface not supported".to_owned()),
            }
        }
    }

    fn create_surface(instance: &Handle<sys::wgb_instance_t>, sink: &sys::wgb_surface_target) -> Handle<sys::wgb_surface_t> {
        let mut out: sys::wgb_surface = ptr::null_mut();
        check(unsafe { sys::wgb_instance_create_surface(instance.raw(), sink, &mut out) });
        unsafe { Handle::adopt(out) }
    }

    impl SurfaceInterface for Surface {
        fn get_capabilities(&self, adapter: &DispatchAdapter) -> wgpu::SurfaceCapabilities {
            let adapter = adapter.as_custom::<super::Adapter>().unwrap();
            let mut caps: sys::wgb_surface_capabilities = unsafe { std::mem::zeroed() };
            check(unsafe { sys::wgb_surface_get_capabilities(self.handle.raw(), adapter.handle.raw(), &mut caps) });
            // the library answers what the reference answers (surface.rs:52-72)
            assert_eq!((caps.formats[0], caps.present_modes[0], caps.alpha_modes[0]),
                       (sys::WGB_TEXTURE_FORMAT_BGRA8_UNORM, sys::WGB_PRESENT_MODE_IMMEDIATE, sys::WGB_COMPOSITE_ALPHA_MODE_OPAQUE));
            wgpu::SurfaceCapabilities {
                formats: vec![wgpu::TextureFormat::Bgra8Unorm],
                present_modes: vec![wgpu::PresentMode::Immediate],
                alpha_modes: vec![wgpu::CompositeAlphaMode::Opaque],
                usages: wgpu::TextureUsages::from_bits_truncate(caps.usages),
            }
        }

        fn configure(&self, device: &DispatchDevice, config: &wgpu::SurfaceConfiguration) {
            let device = device.as_custom::<Device>().unwrap();
            let view_formats: Vec<u32> = config.view_formats.iter().map(|f| super::convert::texture_format(*f).unwrap_or(u32::MAX)).collect();
            let c = sys::wgb_surface_configuration {
                usage: config.usage.bits(),
                format: super::convert::texture_format(config.format).unwrap_or(u32::MAX),     // not Bgra8Unorm: the library refuses
                width: config.width,
                height: config.height,
                present_mode: sys::WGB_PRESENT_MODE_IMMEDIATE,
                alpha_mode: sys::WGB_COMPOSITE_ALPHA_MODE_OPAQUE,
                view_format_count: view_formats.len() as u32,
                view_formats: view_formats.as_ptr(),
            };
            check(unsafe { sys::wgb_surface_configure(self.handle.raw(), device.handle.raw(), &c) });     // panics like the reference's unwrap
            self.window.lock().surface
                .resize(NonZero::new(config.width).expect("Surface width must not be zero"),
                        NonZero::new(config.height).expect("Surface height must not be zero"))
                .unwrap();
            *self.extent.lock() = Some((wgpu::Extent3d { width: config.width, height: config.height, depth_or_array_layers: 1 }, config.format));
        }

        fn get_current_texture(&self) -> (Option<DispatchTexture>, wgpu::SurfaceStatus, DispatchSurfaceOutputDetail) {
            let mut texture: sys::wgb_texture = ptr::null_mut();
            let mut status = 0u32;
            check(unsafe { sys::wgb_surface_get_current_texture(self.handle.raw(), &mut texture, &mut status) });
            let (size, format) = self.extent.lock().expect("Surface not configured yet");
            let texture = Texture { handle: unsafe { Handle::adopt(texture) }, size, format };
            (Some(DispatchTexture::custom(texture)), wgpu::SurfaceStatus::Good,
             DispatchSurfaceOutputDetail::custom(SurfaceOutputDetail { handle: self.handle.clone() }))
        }
    }

    #[derive(Debug)]
    pub struct SurfaceOutputDetail {
        handle: Handle<sys::wgb_surface_t>,
    }

    impl SurfaceOutputDetailInterface for SurfaceOutputDetail {
        fn present(&self) {
            check(unsafe { sys::wgb_surface_present(self.handle.raw()) });
        }

        fn texture_discard(&self) {
            check(unsafe { sys::wgb_surface_texture_discard(self.handle.raw()) });
        }
    }

    // raw-window-handle plumbing, as the reference has it (surface.rs:200-232)
    struct Display(wgpu::rwh::DisplayHandle<'static>);
    impl From<wgpu::rwh::RawDisplayHandle> for Display {
        fn from(value: wgpu::rwh::RawDisplayHandle) -> Self {
            Self(unsafe { wgpu::rwh::DisplayHandle::borrow_raw(value) })
        }
    }
    impl wgpu::rwh::HasDisplayHandle for Display {
        fn display_handle(&self) -> Result<wgpu::rwh::DisplayHandle<'_>, wgpu::rwh::HandleError> {
            Ok(self.0)
        }
    }
    unsafe impl Send for Display {}
    unsafe impl Sync for Display {}

    struct WindowHandle(wgpu::rwh::WindowHandle<'static>);
    impl From<wgpu::rwh::RawWindowHandle> for WindowHandle {
        fn from(value: wgpu::rwh::RawWindowHandle) -> Self {
            Self(unsafe { wgpu::rwh::WindowHandle::borrow_raw(value) })
        }
    }
    impl wgpu::rwh::HasWindowHandle for WindowHandle {
        fn window_handle(&self) -> Result<wgpu::rwh::WindowHandle<'_>, wgpu::rwh::HandleError> {
            Ok(self.0)
        }
    }
    unsafe impl Send for WindowHandle {}
}
