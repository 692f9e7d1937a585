// shader_api.cpp -- Sdf3DShader: source assembly for .sdf3d / GLSL inputs and the WGSL text
// munging helpers.  Mirrors (behaviour, not code):
//   /root/reference/src/shader.rs:44-67   from_path + module table
//   /root/reference/src/shader.rs:73-104  from_glsl_fragment_shader
//   /root/reference/src/shader.rs:155-216 add_to_source / shader_source_input / write_to_file
//   /root/reference/src/shadertoy.rs:199-352 WgslShaderCode::{remove_function, has_function,
//                                            rename_function, remove_line, add_line}
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <sstream>
#include <string>
#include <vector>

#include "common.h"
#include "frontend/frontend.h"

using s2m_internal::fail;

namespace {

// ---- built-in modules (shader.rs:12-20, :50-62).  These are this project's own WGSL text for
// the three reference libraries (same functions, same operation order).  They are what
// --debug-wgsl shows; for compilation the functions are linked from s2m_sdf3d_lib.h instead.
const char* kModSdfOp = R"WGSL(// sdf::op -- smooth boolean operators on distances
fn sdf_op_smooth_union(d1: f32, d2: f32, k: f32) -> f32 {
    let h = clamp(0.5 + 0.5 * (d2 - d1) / k, 0.0, 1.0);
    return mix(d2, d1, h) - k * h * (1.0 - h);
}
fn sdf_op_smooth_intersection(d1: f32, d2: f32, k: f32) -> f32 {
    let h = clamp(0.5 - 0.5 * (d2 - d1) / k, 0.0, 1.0);
    return mix(d2, d1, h) + k * h * (1.0 - h);
}
fn sdf_op_smooth_subtraction(d1: f32, d2: f32, k: f32) -> f32 {
    let h = clamp(0.5 - 0.5 * (d2 + d1) / k, 0.0, 1.0);
    return mix(d2, -d1, h) + k * h * (1.0 - h);
}
)WGSL";

const char* kModNormal = R"WGSL(// sdf3d::normal -- tetrahedral 4-tap gradient of sdf3d
fn sdf3d_normal(p: vec3<f32>, eps: f32) -> vec3<f32> {
    let v1 = vec3( 1.0, -1.0, -1.0);
    let v2 = vec3(-1.0, -1.0,  1.0);
    let v3 = vec3(-1.0,  1.0, -1.0);
    let v4 = vec3( 1.0,  1.0,  1.0);
    return v1 * sdf3d(p + v1 * eps) + v2 * sdf3d(p + v2 * eps) + v3 * sdf3d(p + v3 * eps) + v4 * sdf3d(p + v4 * eps);
}
)WGSL";

const char* kModPrimitives = R"WGSL(// sdf3d::primitives -- box, cylinder, capsule, sphere, torus
fn sdf3d_box(p: vec3f, b: vec3f) -> f32 {
    let q = abs(p) - 0.5 * b;
    return length(max(q, vec3f(0.0, 0.0, 0.0))) + min(max(q.x, max(q.y, q.z)), 0.0);
}
fn sdf3d_cylinder(p: vec3f, h: f32, r: f32) -> f32 {
    let d: vec2f = abs(vec2(length(p.xz), p.y)) - vec2(r, h);
    return min(max(d.x, d.y), 0.0) + length(max(d, vec2f()));
}
fn sdf3d_capsule(p: vec3f, a: vec3f, b: vec3f, r: f32) -> f32 {
    let pa = p - a;
    let ba = b - a;
    let h = clamp(dot(pa, ba) / dot(ba, ba), 0.0, 1.0);
    return length(pa - ba * h) - r;
}
fn sdf3d_sphere(p: vec3f, s: f32) -> f32 {
    return length(p) - s;
}
fn sdf3d_torus(p: vec3f, t: vec2f) -> f32 {
    let q = vec2(length(p.xz) - t.x, p.y);
    return length(q) - t.y;
}
)WGSL";

const char* const kPrimitiveFns[] = {"sdf3d_box", "sdf3d_cylinder", "sdf3d_capsule", "sdf3d_sphere", "sdf3d_torus"};
const char* const kOpFns[] = {"sdf_op_smooth_union", "sdf_op_smooth_intersection", "sdf_op_smooth_subtraction"};

std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) ++a;
  while (b > a && isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}
bool starts_with(const std::string& s, const std::string& p) { return s.compare(0, p.size(), p) == 0; }
std::string replace_first(std::string s, const std::string& what, const std::string& with) {
  size_t i = s.find(what);
  if (i != std::string::npos) s.replace(i, what.size(), with);
  return s;
}
std::string strip_chars(const std::string& s, const char* chars) {
  std::string o;
  for (char c : s) if (!strchr(chars, c)) o += c;
  return o;
}
std::vector<std::string> split_lines(const std::string& text) {  // like Rust str::lines / BufRead::lines
  std::vector<std::string> out;
  size_t i = 0;
  while (i < text.size()) {
    size_t j = text.find('\n', i);
    std::string line = text.substr(i, j == std::string::npos ? std::string::npos : j - i);
    if (!line.empty() && line.back() == '\r') line.pop_back();
    out.push_back(line);
    if (j == std::string::npos) break;
    i = j + 1;
  }
  return out;
}
bool read_file(const std::string& path, std::string* out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::ostringstream ss;
  ss << f.rdbuf();
  *out = ss.str();
  if (out->compare(0, 3, "\xef\xbb\xbf") == 0) out->erase(0, 3);  // a UTF-8 byte-order mark would hide a `use ...;` on line 1
  return true;
}

void add_builtin(s2m_shader* sh, const char* const* names, size_t n) {
  for (size_t i = 0; i < n; ++i) sh->builtin_functions.push_back(names[i]);
}

// The reference pastes the bytes of its three library files (shader.rs:12-20 include_str!).  This project does not carry
// those files: the module texts above are its own wording of the same functions, and calls to them are linked against
// s2m_sdf3d_lib.h.  A user who has the reference's files and needs --debug-wgsl to be BYTE-identical to the reference's
// dump points S2M_WGSL_MODULE_DIR at the directory that holds sdf3d_primitives.wgsl, sdf_op.wgsl
// and sdf3d_normal.wgsl: their bytes are then pasted verbatim and compiled through the front-end like any user code (same
// operation order, so the same bits: tests/test_frontend.py::test_reference_module_files_give_the_reference_dump).
bool module_file_text(const char* file, std::string* out) {
  const char* dir = getenv("S2M_WGSL_MODULE_DIR");
  if (!dir || !*dir) return false;
  std::ifstream f(std::string(dir) + "/" + file, std::ios::binary);   // verbatim: no BOM stripping, no newline changes
  if (!f) return false;
  std::ostringstream ss;
  ss << f.rdbuf();
  *out = ss.str();
  return true;
}
std::string normal_module_text(s2m_shader* sh) {
  std::string t;
  if (module_file_text("sdf3d_normal.wgsl", &t)) return t;
  if (sh) sh->builtin_functions.push_back("sdf3d_normal");
  return kModNormal;
}

// shader.rs:50-62 module table
bool emit_module(s2m_shader* sh, const std::string& name, std::string* w) {
  auto prim = [&] { std::string t; if (module_file_text("sdf3d_primitives.wgsl", &t)) *w += t; else { *w += kModPrimitives; add_builtin(sh, kPrimitiveFns, 5); } };
  auto op = [&] { std::string t; if (module_file_text("sdf_op.wgsl", &t)) *w += t; else { *w += kModSdfOp; add_builtin(sh, kOpFns, 3); } };
  auto nrm = [&] { *w += normal_module_text(sh); };
  if (name == "sdf::*" || name == "sdf::op") { op(); return true; }
  if (name == "sdf3d::normal") { nrm(); return true; }
  if (name == "sdf3d::primitives") { prim(); return true; }
  if (name == "sdf3d::*") { prim(); nrm(); return true; }
  return false;
}

// shader.rs:159-203 shader_source_input
void process_sdf3d_text(s2m_shader* sh, const std::string& text, const std::string& self_path,
                        const std::string& include_dir, std::string* w, int depth);

void process_sdf3d_file(s2m_shader* sh, const std::string& path, const std::string& include_dir, std::string* w, int depth) {
  std::string text;
  if (!read_file(path, &text)) {
    sh->log += "ERROR Could not include \"" + path + "\": " + strerror(errno) + "\n";  // shader.rs:197-199
    return;
  }
  process_sdf3d_text(sh, text, path, include_dir, w, depth);
}

void process_sdf3d_text(s2m_shader* sh, const std::string& text, const std::string& self_path,
                        const std::string& include_dir, std::string* w, int depth) {
  for (const std::string& line : split_lines(text)) {
    const std::string t = trim(line);
    if (!t.empty() && t.back() == ';') {
      if (starts_with(t, "use")) {
        const std::string name = trim(strip_chars(replace_first(t, "use", ""), "\";"));
        sh->log += "INFO " + name + "\n";  // shader.rs:176
        emit_module(sh, name, w);          // unknown module: line silently dropped
        continue;
      }
      if (starts_with(t, "include")) {
        const std::string file = trim(strip_chars(replace_first(t, "include", ""), "\";"));
        if (file != self_path) {
          if (depth > 64) { sh->log += "ERROR include depth exceeded at \"" + file + "\"\n"; continue; }
          std::string full = file;
          if (!include_dir.empty() && !file.empty() && file[0] != '/') full = include_dir + "/" + file;
          process_sdf3d_file(sh, full, include_dir, w, depth + 1);
        }
        continue;
      }
    }
    *w += line;
    *w += "\n";
  }
}

char* dup_string(const std::string& s) {
  char* p = (char*)malloc(s.size() + 1);
  if (p) memcpy(p, s.c_str(), s.size() + 1);
  return p;
}

// shadertoy.rs:251-293
int remove_function(const std::string& wgsl, const std::string& prefix, std::string* out, std::string* err) {
  std::string nw;
  bool in_fn = false, found = false;
  long braces = 0;
  for (const std::string& raw : split_lines(wgsl)) {
    const std::string line = trim(raw);
    if (starts_with(line, prefix)) { in_fn = true; found = true; }
    if (in_fn)
      for (char ch : line) { if (ch == '{') ++braces; else if (ch == '}') --braces; }
    if (braces == 0) {
      if (!in_fn) nw += line + "\n";
      in_fn = false;
    }
  }
  if (!found) { *err = "Function " + prefix + " not found in shader"; return S2M_ERR_SHADER; }
  *out = nw;
  return S2M_OK;
}
// shadertoy.rs:295-314
bool has_function(const std::string& wgsl, const std::string& name) {
  const std::string pat = "fn " + name + "(";
  for (const std::string& raw : split_lines(wgsl))
    if (starts_with(trim(raw), pat)) return true;
  return false;
}
// shadertoy.rs:316-352
int rename_function(const std::string& wgsl, const std::string& old_name, const std::string& new_name,
                    std::string* out, std::string* err) {
  std::string nw;
  bool in_fn = false, found = false;
  const std::string pat = "fn " + old_name + "(";
  for (const std::string& raw : split_lines(wgsl)) {
    const std::string line = trim(raw);
    if (starts_with(line, pat)) {
      in_fn = true; found = true;
      nw += replace_first(line, old_name, new_name);  // (the reference appends no newline here)
    } else {
      nw += line + "\n";
    }
    if (in_fn && starts_with(line, "}")) in_fn = false;
  }
  if (!found) { *err = "Function `" + old_name + "` not found in shader"; return S2M_ERR_SHADER; }
  *out = nw;
  return S2M_OK;
}
// shadertoy.rs:221-230
std::string remove_line(const std::string& wgsl, const std::string& what) {
  std::string s;
  for (const std::string& line : split_lines(wgsl))
    if (trim(line) != trim(what)) { s += line; s += "\n"; }
  return s;
}

// the front-end's note about a function it left out ("// left out: fn NAME -- reason"), for the error text
std::string why_left_out(const std::string& wgsl, const std::string& fn) {
  const std::string tag = "// left out: fn " + fn + " -- ";
  const size_t at = wgsl.find(tag);
  if (at == std::string::npos) return "";
  const size_t end = wgsl.find('\n', at);
  return wgsl.substr(at + tag.size(), end == std::string::npos ? std::string::npos : end - at - tag.size());
}

int build_from_glsl(const std::string& glsl, const std::string& sdf, s2m_shader** out) {
  std::string wgsl, err;
  int st = s2m_frontend::glsl_to_wgsl(glsl, &wgsl, &err);  // shadertoy.rs:199 WgslShaderCode::from_glsl
  if (st) return fail(st, err);
  if ((st = remove_function(wgsl, "fn main_1(", &wgsl, &err))) return fail(st, err);  // shader.rs:84
  if ((st = remove_function(wgsl, "fn main(", &wgsl, &err))) return fail(st, err);    // :85
  wgsl = remove_line(wgsl, "@fragment");                                               // :86
  bool normal_is_builtin = true;
  { std::string t; if (module_file_text("sdf3d_normal.wgsl", &t)) { wgsl += t; normal_is_builtin = false; } else wgsl += kModNormal; }
  wgsl += "\n";                                                                        // :87 add_line(include_str!("sdf3d_normal.wgsl"))
  if (has_function(wgsl, sdf)) {                                                       // :89-98
    if (!has_function(wgsl, "sdf3d")) wgsl += "fn sdf3d(p: vec3<f32>) -> f32 { return " + sdf + "(p); }\n";
  } else {
    const std::string why = why_left_out(wgsl, sdf);
    if (!why.empty()) return fail(S2M_ERR_UNSUPPORTED, "SDF function `" + sdf + "`: " + why);
    return fail(S2M_ERR_MISSING_SDF, "Missing SDF function `" + sdf + "` in shader");
  }
  s2m_shader* sh = new s2m_shader();
  sh->kind = S2M_SRC_WGSL;  // from here on it is ordinary assembled WGSL, as in the reference
  sh->source = wgsl;
  sh->sdf_name = sdf;
  sh->glsl = glsl;
  if (normal_is_builtin) sh->builtin_functions.push_back("sdf3d_normal");
  *out = sh;
  return S2M_OK;
}

// Shader::default_uniform_block (shadertoy.rs:141-152): the ShaderToy inputs, all read as zero
const char* kShaderToyUniforms =
    "layout(binding=0) uniform vec3 iResolution;\n"
    "layout(binding=0) uniform float iTime;\n"
    "layout(binding=0) uniform float iTimeDelta;\n"
    "layout(binding=0) uniform int iFrame;\n"
    "layout(binding=0) uniform vec4 iChannelTime;\n"
    "layout(binding=0) uniform vec4 iMouse;\n"
    "layout(binding=0) uniform vec4 iDate;\n"
    "layout(binding=0) uniform float iSampleRate;\n";

// Sdf3DShader::from_shadertoy_api (shader.rs:110-144) minus the REST fetch: `code` is the text of
// the shader's last render pass (shadertoy.rs:154-167 generate_wgsl_shader_code wraps it).
int build_from_shadertoy(const std::string& code, const std::string& sdf, s2m_shader** out) {
  const std::string glsl = std::string("#version 450 core\n") + kShaderToyUniforms + code + "\n void main() {}";
  std::string wgsl, err;
  int st = s2m_frontend::glsl_to_wgsl(glsl, &wgsl, &err);
  if (st) return fail(st, err);
  if ((st = remove_function(wgsl, "fn main_1(", &wgsl, &err))) return fail(st, err);
  if ((st = remove_function(wgsl, "fn main(", &wgsl, &err))) return fail(st, err);
  if ((st = remove_function(wgsl, "fn mainImage(", &wgsl, &err))) return fail(st, err);  // shader.rs:123
  wgsl = remove_line(wgsl, "@fragment");
  if (has_function(wgsl, sdf)) {
    if (!has_function(wgsl, "sdf3d")) wgsl += "fn sdf3d(p: vec3<f32>) -> f32 { return " + sdf + "(p); }\n";
  } else {
    const std::string why = why_left_out(wgsl, sdf);
    if (!why.empty()) return fail(S2M_ERR_UNSUPPORTED, "SDF function `" + sdf + "`: " + why);
    return fail(S2M_ERR_MISSING_SDF, "Missing SDF function `" + sdf + "` in shader");
  }
  bool normal_is_builtin = true;
  { std::string t; if (module_file_text("sdf3d_normal.wgsl", &t)) { wgsl += t; normal_is_builtin = false; } else wgsl += kModNormal; }
  wgsl += "\n";  // shader.rs:139
  s2m_shader* sh = new s2m_shader();
  sh->kind = S2M_SRC_WGSL;
  sh->source = wgsl;
  sh->sdf_name = sdf;
  sh->glsl = glsl;
  if (normal_is_builtin) sh->builtin_functions.push_back("sdf3d_normal");
  *out = sh;
  return S2M_OK;
}

}  // namespace

namespace {
// --- just enough JSON for a ShaderToy API response (shadertoy.rs:5-68, :119-123) -----------------------
struct Json {
  const std::string& s;
  size_t i = 0;
  bool ok = true;
  explicit Json(const std::string& text) : s(text) {}
  void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\r' || s[i] == '\t')) ++i; }
  bool eat(char c) { ws(); if (i < s.size() && s[i] == c) { ++i; return true; } return false; }
  static void utf8(unsigned cp, std::string* out) {
    if (cp < 0x80) *out += (char)cp;
    else if (cp < 0x800) { *out += (char)(0xC0 | (cp >> 6)); *out += (char)(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { *out += (char)(0xE0 | (cp >> 12)); *out += (char)(0x80 | ((cp >> 6) & 0x3F)); *out += (char)(0x80 | (cp & 0x3F)); }
    else { *out += (char)(0xF0 | (cp >> 18)); *out += (char)(0x80 | ((cp >> 12) & 0x3F)); *out += (char)(0x80 | ((cp >> 6) & 0x3F)); *out += (char)(0x80 | (cp & 0x3F)); }
  }
  bool hex4(unsigned* v) {
    if (i + 4 > s.size()) return false;
    *v = 0;
    for (int k = 0; k < 4; ++k) {
      const char c = s[i++];
      *v = *v * 16 + (c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : 99);
      if ((*v & 0xFF) >= 99 && !(c >= '0' && c <= '9') && !(c >= 'a' && c <= 'f') && !(c >= 'A' && c <= 'F')) return false;
    }
    return true;
  }
  bool string(std::string* out) {
    if (!eat('"')) return ok = false;
    while (i < s.size() && s[i] != '"') {
      char c = s[i++];
      if (c != '\\') { if (out) *out += c; continue; }
      if (i >= s.size()) return ok = false;
      c = s[i++];
      std::string piece;
      switch (c) {
        case 'n': piece = "\n"; break; case 't': piece = "\t"; break; case 'r': piece = "\r"; break;
        case 'b': piece = "\b"; break; case 'f': piece = "\f"; break;
        case 'u': {
          unsigned cp = 0;
          if (!hex4(&cp)) return ok = false;
          if (cp >= 0xD800 && cp < 0xDC00 && i + 1 < s.size() && s[i] == '\\' && s[i + 1] == 'u') {  // surrogate pair
            i += 2;
            unsigned lo = 0;
            if (!hex4(&lo)) return ok = false;
            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
          }
          utf8(cp, &piece);
          break;
        }
        default: piece = std::string(1, c);  // \" \\ \/
      }
      if (out) *out += piece;
    }
    if (i >= s.size()) return ok = false;
    ++i;
    return true;
  }
  // skips any value; for an object calls member(key) before each value -- it returns true if it consumed the value
  typedef std::function<bool(const std::string&)> Member;
  static bool skip_member(const std::string&) { return false; }
  bool value(const Member& member) {
    ws();
    if (i >= s.size()) return ok = false;
    if (s[i] == '"') return string(nullptr);
    if (s[i] == '{') {
      ++i;
      if (eat('}')) return true;
      do {
        std::string key;
        if (!string(&key) || !eat(':')) return ok = false;
        if (!member(key) && !value(Member(skip_member))) return false;
      } while (ok && eat(','));
      return ok && eat('}') ? true : (ok = false);
    }
    if (s[i] == '[') {
      ++i;
      if (eat(']')) return true;
      do { if (!value(member)) return false; } while (eat(','));
      return eat(']') ? true : (ok = false);
    }
    const size_t start = i;
    while (i < s.size() && s[i] != ',' && s[i] != '}' && s[i] != ']' && s[i] != ' ' && s[i] != '\n' && s[i] != '\r' && s[i] != '\t') ++i;
    return i > start ? true : (ok = false);
  }
};
}  // namespace

// Sdf3DShader::from_shadertoy_api (shader.rs:110-144) after the HTTP GET the host makes (shadertoy.rs:126-131:
// https://www.shadertoy.com/api/v1/shaders/{id}?key=...): `body` is the response, {"Shader": {..., "renderpass":
// [{"code": ...}, ...]}} or {"Error": "..."}.  The code of all passes is concatenated, as fetch_code_from_last_pass does.
extern "C" int s2m_shader_from_shadertoy_response(const char* body, size_t len, const char* sdf_name, s2m_shader** out) {
  if (!body || !out) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  *out = nullptr;
  const std::string text(body, len);
  Json j(text);
  std::string code, api_error, name, username;
  bool saw_shader = false, saw_error = false;
  // top level: {"Shader": {...}} | {"Error": "..."}
  std::function<bool(const std::string&)> in_pass = [&](const std::string& key) {
    if (key == "code") { std::string c; if (j.string(&c)) code += c; return true; }
    return false;
  };
  std::function<bool(const std::string&)> in_info = [&](const std::string& key) {
    if (key == "name") { j.string(&name); return true; }
    if (key == "username") { j.string(&username); return true; }
    return false;
  };
  std::function<bool(const std::string&)> in_shader = [&](const std::string& key) {
    if (key == "renderpass") { j.value(in_pass); return true; }
    if (key == "info") { j.value(in_info); return true; }
    return false;
  };
  std::function<bool(const std::string&)> top = [&](const std::string& key) {
    if (key == "Shader") { saw_shader = true; j.value(in_shader); return true; }
    if (key == "Error") { saw_error = true; j.ws(); if (j.i < text.size() && text[j.i] == '"') j.string(&api_error); else j.value(Json::Member(Json::skip_member)); return true; }
    return false;
  };
  j.ws();
  if (j.i >= text.size() || text[j.i] != '{' || !j.value(top) || !j.ok) return fail(S2M_ERR_SHADER, "ShaderToy response is not valid JSON");
  if (saw_error) return fail(S2M_ERR_SHADER, api_error.empty() ? "ShaderToy API error" : api_error);  // ShaderProcessingError::ShaderError
  if (!saw_shader) return fail(S2M_ERR_SHADER, "not a ShaderToy API response (no `Shader` member)");
  int st = build_from_shadertoy(code, sdf_name && *sdf_name ? sdf_name : "sdf", out);
  if (st == S2M_OK) (*out)->log = "INFO Shader: " + name + "\nINFO Shader author: " + username + "\n" + (*out)->log;  // shader.rs:115-116
  return st;
}

extern "C" int s2m_shader_from_shadertoy_source(const char* code, size_t len, const char* sdf_name, s2m_shader** out) {
  if (!code || !out) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  *out = nullptr;
  return build_from_shadertoy(std::string(code, len), sdf_name && *sdf_name ? sdf_name : "sdf", out);
}

extern "C" int s2m_shader_from_path(const char* path, s2m_shader** out) {
  if (!path || !out) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  s2m_shader* sh = new s2m_shader();
  sh->kind = S2M_SRC_SDF3D;
  process_sdf3d_file(sh, path, "", &sh->source, 0);
  *out = sh;
  return S2M_OK;
}

extern "C" int s2m_shader_from_glsl_fragment_shader(const char* path, const char* sdf_name, s2m_shader** out) {
  if (!path || !out) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  *out = nullptr;
  std::string glsl;
  if (!read_file(path, &glsl)) return fail(S2M_ERR_IO, std::string("cannot open ") + path + ": " + strerror(errno));
  return build_from_glsl(glsl, sdf_name && *sdf_name ? sdf_name : "sdf", out);
}

extern "C" int s2m_shader_from_source(const char* text, size_t len, int kind, const char* sdf_name,
                                      const char* include_dir, s2m_shader** out) {
  if (!text || !out) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  *out = nullptr;
  const std::string src(text, len);
  if (kind == S2M_SRC_GLSL_FRAGMENT) return build_from_glsl(src, sdf_name && *sdf_name ? sdf_name : "sdf", out);
  if (kind != S2M_SRC_SDF3D && kind != S2M_SRC_WGSL && kind != S2M_SRC_CUDA) return fail(S2M_ERR_INVALID_ARG, "unknown source kind");
  s2m_shader* sh = new s2m_shader();
  sh->kind = kind;
  if (kind == S2M_SRC_SDF3D) process_sdf3d_text(sh, src, "", include_dir ? include_dir : "", &sh->source, 0);
  else sh->source = src;
  *out = sh;
  return S2M_OK;
}

extern "C" int s2m_shader_add_to_source(s2m_shader* s, const char* text) {
  if (!s || !text) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  s->source += text;
  return S2M_OK;
}
extern "C" const char* s2m_shader_source(const s2m_shader* s) { return s ? s->source.c_str() : ""; }
extern "C" const char* s2m_shader_log(const s2m_shader* s) { return s ? s->log.c_str() : ""; }
extern "C" int s2m_shader_write_to_file(const s2m_shader* s, const char* path) {
  if (!s || !path) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  FILE* f = fopen(path, "wb");
  if (!f) return fail(S2M_ERR_IO, std::string("cannot create ") + path + ": " + strerror(errno));
  size_t n = fwrite(s->source.data(), 1, s->source.size(), f);
  int rc = fclose(f);
  if (n != s->source.size() || rc) return fail(S2M_ERR_IO, std::string("short write to ") + path);
  return S2M_OK;
}
extern "C" int s2m_shader_lower_to_cuda(const s2m_shader* s, char** cuda_out) {
  if (!s || !cuda_out) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  *cuda_out = nullptr;
  std::string cuda, err;
  if (s->kind == S2M_SRC_CUDA) cuda = s->source;
  else {
    int st = s2m_frontend::lower_to_cuda(*s, &cuda, &err, nullptr);
    if (st) return fail(st, err);
  }
  *cuda_out = dup_string(cuda);
  return *cuda_out ? S2M_OK : fail(S2M_ERR_OOM, "malloc");
}
extern "C" int s2m_shader_lower_to_cuda_packed(const s2m_shader* s, char** cuda_out) {
  if (!s || !cuda_out) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  *cuda_out = nullptr;
  std::string cuda, packed, err;
  if (s->kind != S2M_SRC_CUDA) {
    int st = s2m_frontend::lower_to_cuda(*s, &cuda, &err, &packed);
    if (st) return fail(st, err);
  }
  *cuda_out = dup_string(packed);
  return *cuda_out ? S2M_OK : fail(S2M_ERR_OOM, "malloc");
}
extern "C" void s2m_shader_free(s2m_shader* s) { delete s; }

extern "C" int s2m_glsl_to_wgsl(const char* glsl, char** wgsl_out) {
  if (!glsl || !wgsl_out) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  *wgsl_out = nullptr;
  std::string w, err;
  int st = s2m_frontend::glsl_to_wgsl(glsl, &w, &err);
  if (st) return fail(st, err);
  *wgsl_out = dup_string(w);
  return S2M_OK;
}
extern "C" int s2m_wgsl_remove_function(const char* wgsl, const char* fn_prefix, char** out) {
  if (!wgsl || !fn_prefix || !out) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  *out = nullptr;
  std::string o, err;
  int st = remove_function(wgsl, fn_prefix, &o, &err);
  if (st) return fail(st, err);
  *out = dup_string(o);
  return S2M_OK;
}
extern "C" int s2m_wgsl_has_function(const char* wgsl, const char* fn_name, int* found) {
  if (!wgsl || !fn_name || !found) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  *found = has_function(wgsl, fn_name) ? 1 : 0;
  return S2M_OK;
}
extern "C" int s2m_wgsl_rename_function(const char* wgsl, const char* old_name, const char* new_name, char** out) {
  if (!wgsl || !old_name || !new_name || !out) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  *out = nullptr;
  std::string o, err;
  int st = rename_function(wgsl, old_name, new_name, &o, &err);
  if (st) return fail(st, err);
  *out = dup_string(o);
  return S2M_OK;
}
