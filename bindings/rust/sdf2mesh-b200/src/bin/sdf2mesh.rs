//! `run()` of /root/reference/src/bin/sdf2mesh/main.rs:177-364 on top of libsdf2mesh_b200.so: same flags, same
//! log lines, the block from `wgpu::Instance::default()` to `TriangleMesh::from(vertex_items)` replaced.
//! NOT COMPILED in the build image (no rustc / cargo there); `cli/sdf2mesh.cpp` is the built and tested twin.
use sdf2mesh_b200::{ffi, Context, MeshParams, Sdf3DShader};

#[derive(Default, Debug)]
struct Arguments {
    sdf: Option<String>,          // -i, --sdf
    shadertoy_id: Option<String>, // --shadertoy-id   (needs the reference's reqwest client; not linked here)
    shadertoy_sdf: String,        // --shadertoy-sdf, default "sdf"
    glsl: Option<String>,         // --glsl
    glsl_sdf: String,             // --glsl-sdf, default "sdf"
    mesh: String,                 // -0, --mesh
    debug_wgsl: Option<String>,   // --debug-wgsl
    debug_png: Option<String>,    // --debug-png (accepted; the per-slice textures it dumped no longer exist)
    resolution: Option<u32>,      // -r, --resolution
    bounds: Option<f32>,          // -b, --bounds
}

fn parse() -> Arguments {
    let mut a = Arguments { shadertoy_sdf: "sdf".into(), glsl_sdf: "sdf".into(), ..Default::default() };
    let mut it = std::env::args().skip(1);
    while let Some(flag) = it.next() {
        let mut value = || it.next().unwrap_or_else(|| panic!("{flag} needs a value"));
        match flag.as_str() {
            "-i" | "--sdf" => a.sdf = Some(value()),
            "--shadertoy-id" => a.shadertoy_id = Some(value()),
            "--shadertoy-sdf" => a.shadertoy_sdf = value(),
            "--glsl" => a.glsl = Some(value()),
            "--glsl-sdf" => a.glsl_sdf = value(),
            "-0" | "--mesh" => a.mesh = value(),
            "--debug-wgsl" => a.debug_wgsl = Some(value()),
            "--debug-png" => a.debug_png = Some(value()),
            "-r" | "--resolution" => a.resolution = Some(value().parse().expect("resolution")),
            "-b" | "--bounds" => a.bounds = Some(value().parse().expect("bounds")),
            other => panic!("unknown argument {other}"),
        }
    }
    assert!(!a.mesh.is_empty(), "the following required arguments were not provided: --mesh <MESH>");
    a
}

fn main() {
    env_logger::init();
    let args = parse();

    let (params, rounded) = MeshParams::from_cli(args.resolution, args.bounds);
    if rounded {
        log::warn!("Resolution should be a power of 2 (actual resolution : {})", params.0.dims[0]);
    }
    let params = params.with_flags(ffi::S2M_MESH_KEEP_INVALID);

    let ctx = Context::new(0).unwrap();

    let shader = if let Some(sdf) = &args.sdf {
        log::info!("Reading SDF from file {sdf}...");
        Sdf3DShader::from_path(sdf)
    } else if let Some(glsl) = &args.glsl {
        log::info!("Reading GLSL from file {glsl}...");
        Sdf3DShader::from_glsl_fragment_shader(glsl, &args.glsl_sdf).unwrap()
    } else if args.shadertoy_id.is_some() {
        // keep `shadertoy::Shader::from_api(id).await` for the fetch, then
        // Sdf3DShader::from_shadertoy_source(&shader.fetch_code_from_last_pass(), &args.shadertoy_sdf)
        panic!("--shadertoy-id needs the reference's REST client");
    } else {
        panic!("no input: --sdf, --glsl or --shadertoy-id");
    };
    if let Some(path) = &args.debug_wgsl {
        shader.write_to_file(path).unwrap();
    }
    if args.debug_png.is_some() {
        log::warn!("--debug-png: there are no per-slice textures to dump any more");
    }

    let module = ctx.create_shader_module(&shader).unwrap();
    let result = ctx.mesh_run(&module, &params).unwrap();
    log::info!("Mesh has {} vertices.", result.len());
    for quad in result.invalid_quads() {
        log::warn!("Invalid quad: {:?}. Mesh will not be water-tight!", quad);
    }
    if let Err(err) = result.write_to_file(&args.mesh) {
        log::error!("Could not write mesh to {}!", err);
    }
}
