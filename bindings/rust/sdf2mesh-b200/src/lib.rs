//! Safe Rust surface over `libsdf2mesh_b200.so` that keeps the names of WilstonOreo/sdf2mesh, so that
//! `run()` (src/bin/sdf2mesh/main.rs:177-364) changes only where it touched wgpu.
//!
//! NOT COMPILED in the build image (no rustc / cargo there).  Everything here is a thin layer over
//! `ffi.rs`; the same C ABI is what the tested C++ CLI and Python host call.
pub mod ffi;

use std::ffi::{CStr, CString};
use std::os::raw::c_int;
use std::path::Path;
use std::ptr;

/// shadertoy.rs:70-80, without the variants that carried naga / reqwest types: the messages are the
/// front-end's, with line numbers.
#[derive(Debug)]
pub enum ShaderProcessingError {
    ShaderError(String),
    ParseErrors(String),
    ValidationError(String),
    /// Error when the SDF is missing in the shader
    MissingSdf(String),
}

/// Everything that was an `unwrap()` / `expect()` panic around wgpu in the reference.
#[derive(Debug)]
pub enum Error {
    Shader(ShaderProcessingError),
    Io(String),
    Nvrtc(String),
    Cuda(String),
    NoDevice(String),
    Other(c_int, String),
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::s2m_last_error()).to_string_lossy().into_owned() }
}

fn check(status: c_int) -> Result<(), Error> {
    use ShaderProcessingError::*;
    match status {
        ffi::S2M_OK => Ok(()),
        ffi::S2M_ERR_PARSE => Err(Error::Shader(ParseErrors(last_error()))),
        ffi::S2M_ERR_VALIDATION => Err(Error::Shader(ValidationError(last_error()))),
        ffi::S2M_ERR_MISSING_SDF => Err(Error::Shader(MissingSdf(last_error()))),
        ffi::S2M_ERR_SHADER => Err(Error::Shader(ShaderError(last_error()))),
        ffi::S2M_ERR_IO => Err(Error::Io(last_error())),
        ffi::S2M_ERR_NVRTC => Err(Error::Nvrtc(last_error())),
        ffi::S2M_ERR_CUDA | ffi::S2M_ERR_OOM => Err(Error::Cuda(last_error())),
        ffi::S2M_ERR_NO_DEVICE => Err(Error::NoDevice(last_error())),
        s => Err(Error::Other(s, last_error())),
    }
}

fn cpath(p: impl AsRef<Path>) -> CString {
    CString::new(p.as_ref().to_string_lossy().as_bytes()).expect("path contains a NUL byte")
}

/// shader.rs:35-40
pub struct Sdf3DShader(*mut ffi::s2m_shader);

impl Sdf3DShader {
    /// shader.rs:44 -- infallible: an unreadable file is logged and yields an empty source
    pub fn from_path(path: impl AsRef<Path>) -> Self {
        let mut h = ptr::null_mut();
        let p = cpath(path);
        let st = unsafe { ffi::s2m_shader_from_path(p.as_ptr(), &mut h) };
        assert!(st == ffi::S2M_OK && !h.is_null(), "{}", last_error());
        let s = Self(h);
        for line in s.log().lines() {
            log::info!("{line}");
        }
        s
    }

    /// shader.rs:73
    pub fn from_glsl_fragment_shader(path: impl AsRef<Path>, sdf: &str) -> Result<Self, ShaderProcessingError> {
        let mut h = ptr::null_mut();
        let (p, f) = (cpath(path), CString::new(sdf).unwrap());
        match check(unsafe { ffi::s2m_shader_from_glsl_fragment_shader(p.as_ptr(), f.as_ptr(), &mut h) }) {
            Ok(()) => Ok(Self(h)),
            Err(Error::Shader(e)) => Err(e),
            Err(e) => Err(ShaderProcessingError::ShaderError(format!("{e:?}"))),
        }
    }

    /// shader.rs:110 once the caller has fetched the code (`shadertoy::Shader::fetch_code_from_last_pass`)
    pub fn from_shadertoy_source(code: &str, sdf: &str) -> Result<Self, ShaderProcessingError> {
        let mut h = ptr::null_mut();
        let f = CString::new(sdf).unwrap();
        match check(unsafe { ffi::s2m_shader_from_shadertoy_source(code.as_ptr() as *const _, code.len(), f.as_ptr(), &mut h) }) {
            Ok(()) => Ok(Self(h)),
            Err(Error::Shader(e)) => Err(e),
            Err(e) => Err(ShaderProcessingError::ShaderError(format!("{e:?}"))),
        }
    }

    /// shader.rs:110 `from_shadertoy_api` after the HTTP GET (`shadertoy::Shader::from_api`, shadertoy.rs:126-131):
    /// `body` is the API response; an `{"Error": ...}` response becomes `ShaderError`, as in the reference.
    pub fn from_shadertoy_response(body: &str, sdf: &str) -> Result<Self, ShaderProcessingError> {
        let mut h = ptr::null_mut();
        let f = CString::new(sdf).unwrap();
        match check(unsafe { ffi::s2m_shader_from_shadertoy_response(body.as_ptr() as *const _, body.len(), f.as_ptr(), &mut h) }) {
            Ok(()) => Ok(Self(h)),
            Err(Error::Shader(e)) => Err(e),
            Err(e) => Err(ShaderProcessingError::ShaderError(format!("{e:?}"))),
        }
    }

    /// shader.rs:155
    pub fn add_to_source(&mut self, source: &str) {
        let c = CString::new(source).expect("source contains a NUL byte");
        check(unsafe { ffi::s2m_shader_add_to_source(self.0, c.as_ptr()) }).expect("add_to_source");
    }

    /// shader.rs:206
    pub fn write_to_file(&self, path: impl AsRef<Path>) -> std::io::Result<()> {
        let p = cpath(path);
        check(unsafe { ffi::s2m_shader_write_to_file(self.0, p.as_ptr()) }).map_err(|e| std::io::Error::new(std::io::ErrorKind::Other, format!("{e:?}")))
    }

    /// Front-end + NVRTC without a device: the cubins for sm_100a, to be loaded with `Context::instantiate`.
    pub fn compile(&self) -> Result<Module, Error> {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::s2m_module_compile(ptr::null_mut(), self.0, 0, &mut h) })?;
        Ok(Module(h))
    }

    pub fn source(&self) -> String {
        unsafe { CStr::from_ptr(ffi::s2m_shader_source(self.0)).to_string_lossy().into_owned() }
    }

    pub fn log(&self) -> String {
        unsafe { CStr::from_ptr(ffi::s2m_shader_log(self.0)).to_string_lossy().into_owned() }
    }
}

impl Drop for Sdf3DShader {
    fn drop(&mut self) {
        unsafe { ffi::s2m_shader_free(self.0) }
    }
}

/// What `wgpu::Instance` / `Adapter` / `Device` / `Queue` were (main.rs:180-196): one B200.
/// One host thread per context; a context is not re-entrant.
pub struct Context(*mut ffi::s2m_ctx);

impl Context {
    pub fn new(device_ordinal: i32) -> Result<Self, Error> {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::s2m_ctx_create(device_ordinal, &mut h) })?;
        Ok(Self(h))
    }

    /// shader.rs:220 `create_shader_module` + main.rs:283 `create_compute_pipeline`:
    /// front-end -> CUDA C++ -> NVRTC (sm_100a) -> loaded module
    pub fn create_shader_module(&self, shader: &Sdf3DShader) -> Result<Module, Error> {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::s2m_module_compile(self.0, shader.0, 0, &mut h) })?;
        Ok(Module(h))
    }

    /// Loads a module compiled without a device (`Sdf3DShader::compile`) or for another GPU into this
    /// context: compile once while the contexts come up, instantiate per GPU.
    pub fn instantiate(&self, compiled: &Module) -> Result<Module, Error> {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::s2m_module_instantiate(compiled.0, self.0, &mut h) })?;
        Ok(Module(h))
    }

    /// main.rs:298-356 (slice loop) + mesh.rs:229-331 (VertexList, quads) in one call
    pub fn mesh_run(&self, module: &Module, params: &MeshParams) -> Result<MeshResult, Error> {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::s2m_mesh_run(self.0, module.0, &params.0, &mut h) })?;
        MeshResult::new(h)
    }

    /// z-slab form for one process per GPU: K1..K4a now, quads after the count exchange
    pub fn mesh_begin(&self, module: &Module, params: &MeshParams) -> Result<MeshResult, Error> {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::s2m_mesh_begin(self.0, module.0, &params.0, &mut h) })?;
        MeshResult::new(h)
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ffi::s2m_ctx_destroy(self.0) }
    }
}

pub struct Module(*mut ffi::s2m_module);

impl Module {
    pub fn log(&self) -> String {
        unsafe { CStr::from_ptr(ffi::s2m_module_log(self.0)).to_string_lossy().into_owned() }
    }
    /// K1 evaluates corner pairs in packed f32x2 arithmetic (FFMA2 / FMUL2)
    pub fn is_packed(&self) -> bool {
        unsafe { ffi::s2m_module_is_packed(self.0) != 0 }
    }
}

impl Drop for Module {
    fn drop(&mut self) {
        unsafe { ffi::s2m_module_free(self.0) }
    }
}

/// `AppState` (main.rs:28-33) without `dims.w`: there is no per-slice loop any more
#[derive(Clone, Copy, Debug)]
pub struct MeshParams(pub ffi::s2m_mesh_params);

impl MeshParams {
    /// `AppState::from(&Arguments)` (main.rs:139-175); the bool says that the resolution was rounded
    pub fn from_cli(resolution: Option<u32>, bounds: Option<f32>) -> (Self, bool) {
        let mut p = ffi::s2m_mesh_params::default();
        let mut rounded = 0;
        let st = unsafe { ffi::s2m_params_from_cli(resolution.unwrap_or(0), bounds.unwrap_or(0.0), &mut p, &mut rounded) };
        assert_eq!(st, ffi::S2M_OK, "{}", last_error());
        (Self(p), rounded != 0)
    }
    pub fn with_flags(mut self, flags: u32) -> Self {
        self.0.flags |= flags;
        self
    }
    pub fn with_z_slab(mut self, z_begin: u32, z_end: u32) -> Self {
        self.0.z_begin = z_begin;
        self.0.z_end = z_end;
        self
    }
}

// ---- the reference's own geometry types (lib.rs:120-211), reduced to what the mesh needs
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct Vec3D {
    pub x: f32,
    pub y: f32,
    pub z: f32,
}
#[derive(Clone, Copy, Debug, Default)]
pub struct Vertex {
    pub pos: Vec3D,
    pub normal: Vec3D,
}
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct Triangle<T: Copy>(pub T, pub T, pub T);
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct Quad<T: Copy>(pub T, pub T, pub T, pub T);

impl<T: Copy> Quad<T> {
    /// lib.rs:199-204 (the quads of a MeshResult are already swapped)
    pub fn make_triangles(&self) -> (Triangle<T>, Triangle<T>) {
        (Triangle(self.2, self.1, self.0), Triangle(self.0, self.3, self.2))
    }
}

/// mesh.rs:143-147
#[derive(Default)]
pub struct TriangleMesh {
    pub vertices: Vec<Vertex>,
    pub triangle_indices: Vec<Triangle<u32>>,
}

/// Vertices and quads of one run (or one z-slab), resident in pinned host memory owned by the library.
pub struct MeshResult {
    h: *mut ffi::s2m_result,
    info: ffi::s2m_result_info,
}

impl MeshResult {
    fn new(h: *mut ffi::s2m_result) -> Result<Self, Error> {
        let mut r = Self { h, info: unsafe { std::mem::zeroed() } };
        r.refresh()?;
        Ok(r)
    }
    fn refresh(&mut self) -> Result<(), Error> {
        check(unsafe { ffi::s2m_result_get(self.h, &mut self.info) })
    }
    /// second half of the z-slab form: `global_vertex_base` = vertices of all lower slabs
    pub fn finish(&mut self, global_vertex_base: i64) -> Result<(), Error> {
        check(unsafe { ffi::s2m_mesh_finish(self.h, global_vertex_base) })?;
        self.refresh()
    }
    pub fn len(&self) -> usize {
        self.info.n_vertices as usize
    }
    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }
    pub fn info(&self) -> &ffi::s2m_result_info {
        &self.info
    }
    pub fn positions(&self) -> &[f32] {
        if self.info.positions.is_null() { &[] } else { unsafe { std::slice::from_raw_parts(self.info.positions, 3 * self.len()) } }
    }
    pub fn normals(&self) -> &[f32] {
        if self.info.normals.is_null() { &[] } else { unsafe { std::slice::from_raw_parts(self.info.normals, 3 * self.len()) } }
    }
    /// x | y<<16 | label<<32 (mesh.rs:224-226), ascending = the reference's vertex order
    pub fn cell_keys(&self) -> &[u64] {
        if self.info.cell_keys.is_null() { &[] } else { unsafe { std::slice::from_raw_parts(self.info.cell_keys, self.len()) } }
    }
    /// `VertexList::fetch_vertices` (mesh.rs:259)
    pub fn fetch_vertices(&self) -> Vec<Vertex> {
        let (p, n) = (self.positions(), self.normals());
        (0..self.len())
            .map(|i| Vertex {
                pos: Vec3D { x: p[3 * i], y: p[3 * i + 1], z: p[3 * i + 2] },
                normal: if n.is_empty() { Vec3D::default() } else { Vec3D { x: n[3 * i], y: n[3 * i + 1], z: n[3 * i + 2] } },
            })
            .collect()
    }
    /// valid quads, after `Quad::swap`, in the order `fetch_triangle_indices` visits them (mesh.rs:280-320)
    pub fn quads(&self) -> Vec<Quad<u64>> {
        let n = self.info.n_quads as usize;
        if !self.info.quads32.is_null() {
            let q = unsafe { std::slice::from_raw_parts(self.info.quads32, 4 * n) };
            q.chunks_exact(4).map(|c| Quad(c[0] as u64, c[1] as u64, c[2] as u64, c[3] as u64)).collect()
        } else if !self.info.quads.is_null() {
            let q = unsafe { std::slice::from_raw_parts(self.info.quads, 4 * n) };
            q.chunks_exact(4).map(|c| Quad(c[0], c[1], c[2], c[3])).collect()
        } else {
            Vec::new()
        }
    }
    /// `VertexList::fetch_triangle_indices` (mesh.rs:267); panics above u32::MAX vertices like the reference's index type
    pub fn fetch_triangle_indices(&self) -> Vec<Triangle<u32>> {
        let mut out = Vec::with_capacity(2 * self.info.n_quads as usize);
        for q in self.quads() {
            let q = Quad(u32::try_from(q.0).unwrap(), u32::try_from(q.1).unwrap(), u32::try_from(q.2).unwrap(), u32::try_from(q.3).unwrap());
            let (a, b) = q.make_triangles();
            out.push(a);
            out.push(b);
        }
        out
    }
    /// the reference's "Invalid quad" warnings (mesh.rs:270-278); needs S2M_MESH_KEEP_INVALID for the records
    pub fn invalid_quads(&self) -> Vec<Quad<u32>> {
        if self.info.invalid_records.is_null() {
            return Vec::new();
        }
        let r = unsafe { std::slice::from_raw_parts(self.info.invalid_records, 6 * self.info.n_invalid_records as usize) };
        r.chunks_exact(6).map(|c| Quad(c[2] as u32, c[3] as u32, c[4] as u32, c[5] as u32)).collect() // u64::MAX -> u32::MAX
    }
    /// `TriangleMesh::write_to_file` (mesh.rs:182): .stl / .ply by extension, byte-identical text
    pub fn write_to_file(&self, path: impl AsRef<Path>) -> std::io::Result<()> {
        let p = cpath(path);
        check(unsafe { ffi::s2m_result_write_mesh(self.h, p.as_ptr()) }).map_err(|e| std::io::Error::new(std::io::ErrorKind::Other, format!("{e:?}")))
    }
    pub fn write_stl_binary(&self, path: impl AsRef<Path>) -> std::io::Result<()> {
        let p = cpath(path);
        check(unsafe { ffi::s2m_result_write_stl_binary(self.h, p.as_ptr()) }).map_err(|e| std::io::Error::new(std::io::ErrorKind::Other, format!("{e:?}")))
    }
}

impl From<&MeshResult> for TriangleMesh {
    /// mesh.rs:334
    fn from(r: &MeshResult) -> Self {
        TriangleMesh { vertices: r.fetch_vertices(), triangle_indices: r.fetch_triangle_indices() }
    }
}

impl Drop for MeshResult {
    fn drop(&mut self) {
        unsafe { ffi::s2m_result_free(self.h) }
    }
}
