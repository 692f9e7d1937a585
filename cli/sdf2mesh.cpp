// sdf2mesh -- command-line front end with the reference's flag surface
// (/root/reference/src/bin/sdf2mesh/main.rs:95-137 `Arguments`, :177-364 `run`, :366-376 `main`),
// driving libsdf2mesh_b200.so through its C ABI.  Same flags, same defaults, same input priority
// (shadertoy > sdf > glsl, main.rs:200-221), same power-of-two rounding with a warning, same log
// lines ("Reading SDF from ...", module names, "Mesh has N vertices.", "Mesh written to ...").
// Additions: --device, --gpus, --all-slices, --watertight, --exact-dense, --no-normals, --binary-stl, --stats,
// --shadertoy-file, --debug-cuda.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../include/sdf2mesh_b200.h"

namespace {

void log_line(const char* level, const std::string& msg) {
  using namespace std::chrono;
  const auto now = system_clock::now();
  const std::time_t t = system_clock::to_time_t(now);
  const long ns = (long)(duration_cast<nanoseconds>(now.time_since_epoch()).count() % 1000000000LL);
  char buf[64];
  std::tm tm{};
  gmtime_r(&t, &tm);
  strftime(buf, sizeof buf, "%Y-%m-%dT%H:%M:%S", &tm);
  fprintf(stderr, "[%s.%09ldZ %-5s sdf2mesh] %s\n", buf, ns, level, msg.c_str());  // env_logger, nanosecond timestamps (main.rs:370-373)
}
void info(const std::string& m) { log_line("INFO", m); }
void warn(const std::string& m) { log_line("WARN", m); }
void error(const std::string& m) { log_line("ERROR", m); }

[[noreturn]] void die(const std::string& what) {
  // the reference unwrap()s / expect()s here and panics
  error(what + ": " + s2m_last_error());
  exit(101);
}

void usage(FILE* f) {
  fputs(
      "sdf2mesh\n\n"
      "Usage: sdf2mesh [OPTIONS] --mesh <MESH>\n\n"
      "Options:\n"
      "  -i, --sdf <SDF>                      Input SDF file\n"
      "      --shadertoy-id <SHADERTOY_ID>    Input ShaderToy shader ID\n"
      "      --shadertoy-sdf <SHADERTOY_SDF>  ShaderToy SDF name [default: sdf]\n"
      "      --shadertoy-file <FILE>          ShaderToy image-pass code saved to a file (offline stand-in for --shadertoy-id)\n"
      "      --glsl <GLSL>                    Input GLSL fragment shader\n"
      "      --glsl-sdf <GLSL_SDF>            GLSL SDF name [default: sdf]\n"
      "  -0, --mesh <MESH>                    Output mesh file (supports STL and PLY output)\n"
      "      --debug-wgsl <DEBUG_WGSL>        Output WGSL file for debugging\n"
      "      --debug-png <DEBUG_PNG>          Write PNG images for debugging (not available: there are no per-slice textures)\n"
      "  -r, --resolution <RESOLUTION>        Grid resolution. Default: 256\n"
      "  -b, --bounds <BOUNDS>                Size of bounding box. Default: 2\n"
      "      --device <N>                     CUDA device ordinal [default: 0]\n"
      "      --gpus <N>                       Split the grid into N z-slabs, one GPU (and host thread) each [default: 1]\n"
      "      --no-nccl                        With --gpus: exchange the slab vertex counts through host memory, not ncclAllGather\n"
      "      --all-slices                     Mesh every z-slice (the reference never reads back the last one)\n"
      "      --watertight                     --all-slices plus consistent cell corners: no quads lost to rounding\n"
      "      --exact-dense                    Evaluate all 8 corners of every cell, like the reference (slow)\n"
      "      --no-normals                     Skip vertex normals (only PLY output uses them)\n"
      "      --binary-stl                     Write binary instead of ASCII STL\n"
      "      --debug-cuda <FILE>              Write the generated CUDA C++ for debugging\n"
      "      --stats                          Print per-kernel timings\n"
      "  -h, --help                           Print help\n"
      "  -V, --version                        Print version\n",
      f);
}

struct Args {
  std::string sdf, shadertoy_id, shadertoy_file, shadertoy_sdf = "sdf", glsl, glsl_sdf = "sdf", mesh, debug_wgsl, debug_png, debug_cuda;
  unsigned resolution = 0;
  float bounds = 0.0f;
  int device = 0, gpus = 1;
  bool no_nccl = false;
  bool watertight = false, all_slices = false, exact_dense = false, no_normals = false, binary_stl = false, stats = false;
};

bool take_value(int argc, char** argv, int& i, const char* longf, const char* shortf, std::string* out) {
  const std::string a = argv[i];
  const std::string l = std::string("--") + longf;
  if (a == l || (shortf && a == shortf)) {
    if (i + 1 >= argc) { fprintf(stderr, "error: a value is required for '%s <%s>' but none was supplied\n", l.c_str(), longf); exit(2); }
    *out = argv[++i];
    return true;
  }
  if (a.compare(0, l.size() + 1, l + "=") == 0) { *out = a.substr(l.size() + 1); return true; }
  if (shortf && a.size() > 2 && a.compare(0, 2, shortf) == 0) { *out = a.substr(2); return true; }
  return false;
}


// builds the shader from the arguments (input priority: shadertoy > sdf > glsl, main.rs:200-221)
s2m_shader* load_shader(const Args& a, bool quiet) {
  s2m_shader* shader = nullptr;
  if (!a.shadertoy_id.empty()) {
    if (!quiet) info("Reading SDF from ShaderToy (shader ID " + a.shadertoy_id + ")");
    error("the ShaderToy REST fetch is not part of this build (no network access); save the API response "
          "(https://www.shadertoy.com/api/v1/shaders/" + a.shadertoy_id + "?key=<your key>) or the shader's code to a file and use --shadertoy-file");
    exit(101);
  } else if (!a.shadertoy_file.empty()) {
    if (!quiet) info("Reading SDF from ShaderToy code in " + a.shadertoy_file + "...");
    FILE* f = fopen(a.shadertoy_file.c_str(), "rb");
    if (!f) { error("cannot open " + a.shadertoy_file); exit(101); }
    std::string code;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) code.append(buf, n);
    fclose(f);
    // either the image-pass GLSL itself or a saved API response (curl 'https://www.shadertoy.com/api/v1/shaders/ID?key=...' > file)
    size_t first = 0;
    while (first < code.size() && (code[first] == ' ' || code[first] == '\n' || code[first] == '\r' || code[first] == '\t')) ++first;
    const bool is_json = first < code.size() && code[first] == '{';
    if (is_json ? s2m_shader_from_shadertoy_response(code.data(), code.size(), a.shadertoy_sdf.c_str(), &shader)
                : s2m_shader_from_shadertoy_source(code.data(), code.size(), a.shadertoy_sdf.c_str(), &shader))
      die("cannot convert ShaderToy shader");
  } else if (!a.sdf.empty()) {
    if (!quiet) info("Reading SDF from " + a.sdf + "...");
    if (s2m_shader_from_path(a.sdf.c_str(), &shader)) die("cannot read SDF");
  } else if (!a.glsl.empty()) {
    if (!quiet) info("Reading SDF from GLSL fragment shader " + a.glsl + "...");
    if (s2m_shader_from_glsl_fragment_shader(a.glsl.c_str(), a.glsl_sdf.c_str(), &shader)) die("cannot convert GLSL shader");
  } else {
    if (s2m_shader_from_source("", 0, S2M_SRC_SDF3D, nullptr, nullptr, &shader)) die("empty shader");  // Sdf3DShader::default()
  }
  if (!quiet) {  // forward what the reference logs while assembling the source
    std::string lg = s2m_shader_log(shader);
    size_t p = 0;
    while (p < lg.size()) {
      size_t e = lg.find('\n', p);
      if (e == std::string::npos) e = lg.size();
      const std::string line = lg.substr(p, e - p);
      if (line.compare(0, 5, "INFO ") == 0) info(line.substr(5));
      else if (line.compare(0, 6, "ERROR ") == 0) error(line.substr(6));
      else if (!line.empty()) info(line);
      p = e + 1;
    }
  }
  return shader;
}

// --gpus N: the library's one-call form (s2m_multi, csrc/multi.cpp): one context and host thread per GPU, contiguous
// z-slabs balanced with the cost probe, vertex counts exchanged by one ncclAllGather, one output file written from all
// parts.  (The multi-process form of the same protocol is sdf2mesh_b200/distributed.py.)  --no-nccl exchanges the counts
// through host memory.
int run_multi_gpu(const Args& a, s2m_mesh_params params) {
  const int n = a.gpus;
  s2m_shader* shader = load_shader(a, false);
  if (!a.debug_wgsl.empty() && s2m_shader_write_to_file(shader, a.debug_wgsl.c_str())) die("cannot write --debug-wgsl file");
  const auto t0 = std::chrono::steady_clock::now();
  // the N contexts (and the NCCL communicators) come up on another thread while this one lowers and NVRTC-compiles the SDF
  s2m_multi* mc = nullptr;
  int mc_status = 0;
  std::string mc_error;
  std::vector<int> ordinals;
  for (int g = 0; g < n; ++g) ordinals.push_back(a.device + g);
  std::thread mc_thread([&] {
    mc_status = s2m_multi_create(ordinals.data(), n, a.no_nccl ? S2M_MULTI_NO_NCCL : 0u, &mc);
    if (mc_status) mc_error = s2m_last_error();   // the error slot is per thread
  });
  s2m_module* compiled = nullptr;
  const int compile_status = s2m_module_compile(nullptr, shader, 0, &compiled);
  const std::string compile_error = compile_status ? s2m_last_error() : "";
  mc_thread.join();
  if (mc_status) { error("no usable CUDA devices: " + mc_error); return 101; }
  if (compile_status) { error("shader module creation failed: " + compile_error); s2m_multi_destroy(mc); return 101; }
  int nccl_version = 0;
  const bool with_nccl = s2m_multi_uses_nccl(mc, &nccl_version) != 0;
  info("CUDA contexts set up on " + std::to_string(n) + " GPUs" + (with_nccl ? " (NCCL " + std::to_string(nccl_version) + ")." : "."));
  std::vector<s2m_result*> res((size_t)n, nullptr);
  if (s2m_multi_mesh_run(mc, compiled, &params, res.data())) { error(std::string("meshing failed: ") + s2m_last_error()); s2m_multi_destroy(mc); return 101; }
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  std::vector<uint32_t> bounds((size_t)n + 1, 0);
  s2m_multi_get_partition(mc, bounds.data());
  s2m_multi_timings mt;
  s2m_multi_last_timings(mc, &mt);
  uint64_t nv = 0, nq = 0, ninv = 0;
  for (int g = 0; g < n; ++g) {
    s2m_result_info ri;
    s2m_result_get(res[(size_t)g], &ri);
    nv += ri.n_vertices; nq += ri.n_quads; ninv += ri.n_invalid_quads;
    if (a.stats)
      fprintf(stderr, "stats: GPU %d slices [%u, %u): %llu vertices, %llu quads, device %.3f ms | begin %.3f, count exchange %.3f, finish %.3f ms\n", a.device + g,
              bounds[(size_t)g], bounds[(size_t)g + 1], (unsigned long long)ri.n_vertices, (unsigned long long)ri.n_quads, ri.timings.device_ms,
              mt.begin_ms[g], mt.exchange_ms[g], mt.finish_ms[g]);
  }
  info("Mesh has " + std::to_string(nv) + " vertices.");
  if (ninv) warn(std::to_string(ninv) + " invalid quads. Mesh will not be water-tight!");
  if (a.stats) fprintf(stderr, "stats: %llu quads; meshing on %d GPUs %.3f ms; contexts + JIT + meshing %.1f ms wall\n", (unsigned long long)nq, n, mt.wall_ms, ms);
  if (s2m_write_mesh_parts(res.data(), n, a.mesh.c_str(), a.binary_stl ? 1 : 0)) error(std::string("Could not write mesh to ") + s2m_last_error() + "!");
  info("Mesh written to " + a.mesh);
  for (int g = 0; g < n; ++g) s2m_result_free(res[(size_t)g]);
  s2m_module_free(compiled);
  s2m_multi_destroy(mc);
  s2m_shader_free(shader);
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  Args a;
  bool have_mesh = false;
  for (int i = 1; i < argc; ++i) {
    std::string v;
    const std::string s = argv[i];
    if (s == "-h" || s == "--help") { usage(stdout); return 0; }
    if (s == "-V" || s == "--version") { printf("sdf2mesh (%s)\n", s2m_version()); return 0; }
    if (take_value(argc, argv, i, "sdf", "-i", &a.sdf)) continue;
    if (take_value(argc, argv, i, "shadertoy-id", nullptr, &a.shadertoy_id)) continue;
    if (take_value(argc, argv, i, "shadertoy-sdf", nullptr, &a.shadertoy_sdf)) continue;
    if (take_value(argc, argv, i, "shadertoy-file", nullptr, &a.shadertoy_file)) continue;
    if (take_value(argc, argv, i, "glsl", nullptr, &a.glsl)) continue;
    if (take_value(argc, argv, i, "glsl-sdf", nullptr, &a.glsl_sdf)) continue;
    if (take_value(argc, argv, i, "mesh", "-0", &a.mesh)) { have_mesh = true; continue; }
    if (take_value(argc, argv, i, "debug-wgsl", nullptr, &a.debug_wgsl)) continue;
    if (take_value(argc, argv, i, "debug-png", nullptr, &a.debug_png)) continue;
    if (take_value(argc, argv, i, "debug-cuda", nullptr, &a.debug_cuda)) continue;
    if (take_value(argc, argv, i, "resolution", "-r", &v)) {
      char* end = nullptr;
      const long long n = strtoll(v.c_str(), &end, 10);
      if (!end || *end || n < 0 || n > 0xffffffffLL) { fprintf(stderr, "error: invalid value '%s' for '--resolution <RESOLUTION>'\n", v.c_str()); return 2; }
      a.resolution = (unsigned)n;
      continue;
    }
    if (take_value(argc, argv, i, "bounds", "-b", &v)) {
      char* end = nullptr;
      a.bounds = strtof(v.c_str(), &end);
      if (!end || *end) { fprintf(stderr, "error: invalid value '%s' for '--bounds <BOUNDS>'\n", v.c_str()); return 2; }
      continue;
    }
    if (take_value(argc, argv, i, "device", nullptr, &v)) { a.device = atoi(v.c_str()); continue; }
    if (take_value(argc, argv, i, "gpus", nullptr, &v)) { a.gpus = std::max(1, atoi(v.c_str())); continue; }
    if (s == "--all-slices") { a.all_slices = true; continue; }
    if (s == "--watertight") { a.watertight = true; continue; }
    if (s == "--exact-dense") { a.exact_dense = true; continue; }
    if (s == "--no-normals") { a.no_normals = true; continue; }
    if (s == "--binary-stl") { a.binary_stl = true; continue; }
    if (s == "--stats") { a.stats = true; continue; }
    if (s == "--no-nccl") { a.no_nccl = true; continue; }
    fprintf(stderr, "error: unexpected argument '%s' found\n\n", s.c_str());
    usage(stderr);
    return 2;
  }
  if (!have_mesh) {
    fputs("error: the following required arguments were not provided:\n  --mesh <MESH>\n\n", stderr);
    usage(stderr);
    return 2;
  }

  // AppState::from(&args)  (main.rs:139-175)
  s2m_mesh_params params;
  int rounded = 0;
  if (s2m_params_from_cli(a.resolution, a.bounds, &params, &rounded)) die("bad arguments");
  if (rounded) warn("Resolution should be a power of 2 (actual resolution : " + std::to_string(params.dims[0]) + ")");
  if (a.all_slices) params.flags |= S2M_MESH_ALL_SLICES;
  if (a.watertight) params.flags |= S2M_MESH_ALL_SLICES | S2M_MESH_CONSISTENT_CORNERS;
  if (a.exact_dense) params.flags |= S2M_MESH_EXACT_DENSE;
  if (a.no_normals) params.flags |= S2M_MESH_NO_NORMALS;
  params.flags |= S2M_MESH_KEEP_INVALID;  // so that the reference's per-quad warnings can be printed
  if (a.stats) params.flags |= S2M_MESH_TIMINGS;

  if (a.gpus > 1) return run_multi_gpu(a, params);

  // The CUDA context comes up on its own thread while this one reads, lowers and NVRTC-compiles the
  // shader (neither needs the device); the cubins are loaded once both are there.
  const auto t_start = std::chrono::steady_clock::now();
  s2m_ctx* ctx = nullptr;
  int ctx_status = 0;
  std::string ctx_error;
  std::thread ctx_thread([&] {
    ctx_status = s2m_ctx_create(a.device, &ctx);  // request_adapter/request_device .unwrap()
    if (ctx_status) ctx_error = s2m_last_error();  // the error slot is per thread
  });

  s2m_shader* shader = load_shader(a, false);
  if (!a.debug_wgsl.empty() && s2m_shader_write_to_file(shader, a.debug_wgsl.c_str())) { ctx_thread.join(); die("cannot write --debug-wgsl file"); }
  if (!a.debug_png.empty()) warn("--debug-png is ignored: this engine has no per-slice textures to dump");

  s2m_module* compiled = nullptr;
  const int compile_status = s2m_module_compile(nullptr, shader, 0, &compiled);
  ctx_thread.join();
  if (ctx_status) { error("no usable CUDA device: " + ctx_error); exit(101); }
  if (compile_status) die("shader module creation failed");
  s2m_module* module = nullptr;
  if (s2m_module_instantiate(compiled, ctx, &module)) die("shader module creation failed");
  s2m_module_free(compiled);
  const double setup_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
  if (!a.debug_cuda.empty()) {
    FILE* f = fopen(a.debug_cuda.c_str(), "wb");
    if (f) { fputs(s2m_module_cuda_source(module), f); fclose(f); }
  }
  info("CUDA context set up.");

  s2m_result* result = nullptr;
  if (s2m_mesh_run(ctx, module, &params, &result)) die("meshing failed");
  s2m_result_info ri;
  s2m_result_get(result, &ri);
  info("Mesh has " + std::to_string(ri.n_vertices) + " vertices.");
  {  // mesh.rs:276: one warning per invalid quad, with u32::MAX for a missing corner
    const uint64_t shown = ri.n_invalid_records < 64 ? ri.n_invalid_records : 64;
    for (uint64_t i = 0; i < shown; ++i) {
      const uint64_t* q = ri.invalid_records + 6 * i + 2;
      std::string s = "Invalid quad: Quad(";
      for (int k = 0; k < 4; ++k) s += (k ? ", " : "") + (q[k] == UINT64_MAX ? std::string("4294967295") : std::to_string(q[k]));
      warn(s + "). Mesh will not be water-tight!");
    }
    if (ri.n_invalid_quads > shown) warn("... " + std::to_string(ri.n_invalid_quads - shown) + " more invalid quads. Mesh will not be water-tight!");
  }
  if (a.stats) {
    const s2m_timings& t = ri.timings;
    const double vox = (double)params.dims[0] * params.dims[1] * params.dims[2];
    fprintf(stderr, "stats: %llu candidates, %llu vertices, %llu quads | front-end %.1f ms, NVRTC %.1f ms | K1 %.3f K2 %.3f K3 %.3f K4a %.3f K4b %.3f d2h %.3f total %.3f ms (%u launches, %u chunks), mesh phase on the host clock %.3f ms | %.2f Gvoxel/s\n",
            (unsigned long long)ri.n_candidates, (unsigned long long)ri.n_vertices, (unsigned long long)ri.n_quads,
            s2m_module_compile_ms(module, 0), s2m_module_compile_ms(module, 1), t.k1_slab_ms, t.k2_classify_ms, t.k3_compact_ms,
            t.k4_vertices_ms, t.k4_quads_ms, t.d2h_ms, t.total_ms, t.launches, t.chunks, t.host_wall_ms, vox / (t.total_ms * 1e-3) / 1e9);
  }
  int st;
  const std::string& m = a.mesh;
  const bool is_stl = m.size() >= 4 && strcasecmp(m.c_str() + m.size() - 4, ".stl") == 0;
  const auto t_write = std::chrono::steady_clock::now();
  if (a.binary_stl && is_stl) st = s2m_result_write_stl_binary(result, m.c_str());
  else st = s2m_result_write_mesh(result, m.c_str());
  if (st) error(std::string("Could not write mesh to ") + s2m_last_error() + "!");  // main.rs:359-361 (and it carries on)
  info("Mesh written to " + a.mesh);
  if (a.stats)
    fprintf(stderr, "stats: context + shader + NVRTC + module load %.1f ms wall (context creation overlaps the compile) | file written in %.1f ms\n",
            setup_ms, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_write).count());

  s2m_result_free(result);
  s2m_module_free(module);
  s2m_shader_free(shader);
  s2m_ctx_destroy(ctx);
  return 0;
}
