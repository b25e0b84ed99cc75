"""Import shim so the UNMODIFIED reference (/root/reference, read-only) can be imported in this
container to generate golden vectors.  Third-party packages the reference imports at module scope but
that are absent here are replaced by empty stub modules.  Used ONLY by
oracle/gen_golden.py; nothing that runs on the GPU box imports this."""
import importlib.abc
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get("QV2X_REF", "/root/reference")
_STUBS = {"matplotlib", "icecream", "efficientnet_pytorch", "spconv", "shapely", "timm", "pyquaternion",
          "termcolor", "onnx", "tensorrt", "open3d", "easydict", "tensorboardX", "h5py", "skimage", "lzf",
          "pypcd", "cumm", "fvcore"}


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {})
        setattr(self, name, cls)
        return cls


class _Loader(importlib.abc.Loader):
    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        top = fullname.split(".")[0]
        if top in _STUBS:
            try:
                # prefer the real package when it exists
                for f in sys.meta_path:
                    if f is self:
                        continue
                    spec = f.find_spec(fullname, path, target) if hasattr(f, "find_spec") else None
                    if spec is not None:
                        return spec
            except Exception:
                pass
            return importlib.machinery.ModuleSpec(fullname, _Loader(), is_package=True)
        return None


def install():
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.append(_Finder())
    import icecream

    icecream.ic = lambda *a, **k: None
