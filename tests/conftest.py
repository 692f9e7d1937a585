import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the suite always goes through NVRTC: no cubins from (or into) the user's ~/.cache/sdf2mesh_b200 (tests of the cache set their own directory)
os.environ.setdefault("S2M_CACHE_DIR", "")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def built():
    """The in-tree native library and the oracle (built on demand on the CPU box)."""
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "sdf2mesh_b200", "libsdf2mesh_b200.so")) or \
            not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        g.build()
    return True


@pytest.fixture(scope="session", autouse=True)
def packed_crosscheck(built):
    """Every shader a test lowers with Sdf3DShader.lower_to_cuda() is also lowered to its packed f32x2
    form (csrc/s2m_pvec.h); tests.support.host_eval.eval_points then evaluates both and compares them
    lane by lane.  So the whole front-end suite doubles as the packed emitter's suite."""
    import sdf2mesh_b200 as s2m
    from tests.support import host_eval
    orig = s2m.Sdf3DShader.lower_to_cuda

    def lower_and_register(self):
        text = orig(self)
        host_eval.register_packed(text, self.lower_to_cuda_packed())
        return text

    s2m.Sdf3DShader.lower_to_cuda = lower_and_register
    yield
    s2m.Sdf3DShader.lower_to_cuda = orig


@pytest.fixture(scope="session")
def ctx(built):
    import sdf2mesh_b200 as s2m
    c = s2m.Context(0)  # raises S2mError(NO_DEVICE) without a GPU: there is no CPU fallback
    yield c
    c.close()


EXAMPLES = os.path.join(ROOT, "examples")
# name -> (file, kind, oracle sdf id name)
EXAMPLE_INPUTS = {
    "torus": ("torus.sdf3d", "sdf3d", "torus"),
    "martin_cube": ("martin_cube.sdf3d", "sdf3d", "martin_cube"),
    "p_key": ("p_key.sdf3d", "sdf3d", "p_key"),
    "mandelbulb": ("mandelmesh.frag", "glsl", "mandelbulb"),
}


def load_example_shader(name):
    import sdf2mesh_b200 as s2m
    f, kind, _ = EXAMPLE_INPUTS[name]
    path = os.path.join(EXAMPLES, f)
    if kind == "glsl":
        return s2m.Sdf3DShader.from_glsl_fragment_shader(path, "sdf")
    return s2m.Sdf3DShader.from_path(path)
