"""Sdf3DShader -- host-side mirror of /root/reference/src/shader.rs:35-225.

Same constructor names, argument meaning and error behaviour as the reference; the only change is
what `create_shader_module` returns: a CUDA module (front-end -> CUDA C++ -> NVRTC for sm_100a)
instead of a wgpu::ShaderModule.
"""
import ctypes

from . import _capi
from ._capi import check, lib


class WgslShaderCode:
    """String-level WGSL helpers (/root/reference/src/shadertoy.rs:196-249)."""

    def __init__(self, text: str):
        self.text = text

    @classmethod
    def from_glsl(cls, glsl: str) -> "WgslShaderCode":
        return cls(convert_glsl_to_wgsl(glsl))

    def remove_function(self, function_name: str) -> None:
        out = ctypes.c_void_p()
        check(lib().s2m_wgsl_remove_function(self.text.encode(), function_name.encode(), ctypes.byref(out)))
        self.text = _capi.take_string(out)

    def has_function(self, function_name: str) -> bool:
        found = ctypes.c_int()
        check(lib().s2m_wgsl_has_function(self.text.encode(), function_name.encode(), ctypes.byref(found)))
        return bool(found.value)

    def rename_function(self, old: str, new: str) -> None:
        out = ctypes.c_void_p()
        check(lib().s2m_wgsl_rename_function(self.text.encode(), old.encode(), new.encode(), ctypes.byref(out)))
        self.text = _capi.take_string(out)

    def remove_line(self, line: str) -> None:
        self.text = "".join(l + "\n" for l in self.text.splitlines() if l.strip() != line.strip())

    def add_line(self, line: str) -> None:
        self.text += line + "\n"

    def __str__(self):
        return self.text


def convert_glsl_to_wgsl(glsl: str) -> str:
    """shadertoy.rs:169 convert_glsl_to_wgsl."""
    out = ctypes.c_void_p()
    check(lib().s2m_glsl_to_wgsl(glsl.encode(), ctypes.byref(out)))
    return _capi.take_string(out)


class Sdf3DShader:
    def __init__(self, handle=None):
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().s2m_shader_free(self._h)
            except Exception:
                pass
            self._h = None

    # --- constructors -------------------------------------------------------------------
    @classmethod
    def from_path(cls, path) -> "Sdf3DShader":
        """shader.rs:44.  Like the reference this never raises on IO errors (see .log)."""
        h = ctypes.c_void_p()
        check(lib().s2m_shader_from_path(str(path).encode(), ctypes.byref(h)))
        return cls(h)

    @classmethod
    def from_glsl_fragment_shader(cls, path, sdf: str = "sdf") -> "Sdf3DShader":
        """shader.rs:73.  Raises S2mError(PARSE | VALIDATION | MISSING_SDF | SHADER)."""
        h = ctypes.c_void_p()
        check(lib().s2m_shader_from_glsl_fragment_shader(str(path).encode(), sdf.encode(), ctypes.byref(h)))
        return cls(h)

    @classmethod
    def from_source(cls, text: str, kind: int = _capi.SRC_SDF3D, sdf: str = "sdf", include_dir=None) -> "Sdf3DShader":
        h = ctypes.c_void_p()
        raw = text.encode()
        check(lib().s2m_shader_from_source(raw, len(raw), kind, sdf.encode(),
                                           None if include_dir is None else str(include_dir).encode(), ctypes.byref(h)))
        return cls(h)

    @classmethod
    def from_shadertoy_source(cls, code: str, sdf: str = "sdf") -> "Sdf3DShader":
        """shader.rs:110 from_shadertoy_api without the network fetch: `code` is the image-pass GLSL."""
        h = ctypes.c_void_p()
        raw = code.encode()
        check(lib().s2m_shader_from_shadertoy_source(raw, len(raw), sdf.encode(), ctypes.byref(h)))
        return cls(h)

    SHADERTOY_API_KEY = "rdnjhn"   # shadertoy.rs:1

    @classmethod
    def from_shadertoy_json(cls, response, sdf: str = "sdf") -> "Sdf3DShader":
        """A ShaderToy API response (text or parsed) -> shader, as shader.rs:110-144 does after the fetch:
        `{"Shader": {"info": ..., "renderpass": [{"code": ...}, ...]}}` or `{"Error": "..."}` (shadertoy.rs:119-123).
        Like the reference's fetch_code_from_last_pass (shadertoy.rs:126-132) the code of ALL passes is concatenated."""
        import json
        if isinstance(response, (bytes, bytearray)):
            raw = bytes(response)
        else:
            raw = (response if isinstance(response, str) else json.dumps(response)).encode()
        h = ctypes.c_void_p()
        check(lib().s2m_shader_from_shadertoy_response(raw, len(raw), sdf.encode(), ctypes.byref(h)))
        sh = cls(h)
        lines = dict(l[5:].split(": ", 1) for l in sh.log.splitlines() if l.startswith("INFO Shader") and ": " in l)
        sh.info = {"name": lines.get("Shader", ""), "username": lines.get("Shader author", "")}   # what the reference logs (shader.rs:115-116)
        return sh

    @classmethod
    def from_shadertoy_api(cls, shader_id: str, sdf: str = "sdf", fetch=None) -> "Sdf3DShader":
        """shader.rs:110 from_shadertoy_api: GET https://www.shadertoy.com/api/v1/shaders/{id}?key=... (shadertoy.rs:126-131),
        then from_shadertoy_json.  `fetch(url) -> bytes` replaces the HTTP client (tests, proxies); a failed request is
        S2mError(REQUEST), the reference's ShaderProcessingError::RequestError."""
        url = f"https://www.shadertoy.com/api/v1/shaders/{shader_id}?key={cls.SHADERTOY_API_KEY}"
        try:
            if fetch is None:
                import urllib.request
                with urllib.request.urlopen(url, timeout=30) as r:
                    body = r.read()
            else:
                body = fetch(url)
        except _capi.S2mError:
            raise
        except Exception as e:
            raise _capi.S2mError(_capi.ERR_REQUEST, f"request for {url} failed: {e}") from e
        return cls.from_shadertoy_json(body, sdf)

    # --- reference API ------------------------------------------------------------------
    def add_to_source(self, source: str) -> None:
        check(lib().s2m_shader_add_to_source(self._h, source.encode()))

    def write_to_file(self, path) -> None:
        check(lib().s2m_shader_write_to_file(self._h, str(path).encode()))

    @property
    def source(self) -> str:
        return lib().s2m_shader_source(self._h).decode("utf-8", "replace")

    @property
    def log(self) -> str:
        return lib().s2m_shader_log(self._h).decode("utf-8", "replace")

    def lower_to_cuda(self) -> str:
        out = ctypes.c_void_p()
        check(lib().s2m_shader_lower_to_cuda(self._h, ctypes.byref(out)))
        return _capi.take_string(out)

    def lower_to_cuda_packed(self) -> str:
        """the f32x2 form K1 evaluates two corners at a time with ("" if the shader has none)"""
        out = ctypes.c_void_p()
        check(lib().s2m_shader_lower_to_cuda_packed(self._h, ctypes.byref(out)))
        return _capi.take_string(out)

    def create_shader_module(self, ctx=None, flags: int = 0):
        """shader.rs:220.  ctx=None compiles to a cubin without loading it (no GPU needed)."""
        from .engine import Module
        return Module(self, ctx, flags)
