"""Load the read-only reference with auto-mocked missing third-party deps (generation-time only)."""
import sys, types, importlib, importlib.abc, importlib.machinery
from unittest import mock

import abc
class _Meta(abc.ABCMeta):
    def __getattr__(cls, n):
        if n.startswith("__") and n.endswith("__"):
            raise AttributeError(n)
        return mock.MagicMock(name=n)

class _AutoMock(types.ModuleType):
    def __getattr__(self, name):
        if name == "__version__":
            return "0.0.0"
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        if name[:1].isupper():
            m = _Meta(name, (), {"__init__": lambda self, *a, **k: None,
                                "__getattr__": lambda self, n: mock.MagicMock(name=n)})
        else:
            m = mock.MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m

def _importer_file():
    f = sys._getframe(1)
    while f is not None:
        fn = f.f_code.co_filename
        if "importlib" not in fn and fn != __file__ and not fn.startswith("<frozen"):
            return fn
        f = f.f_back
    return ""

class Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    mocked = set()
    def find_spec(self, fullname, path, target=None):
        root = fullname.split(".")[0]
        if root in ("mlrun",):
            return None
        if root not in self.mocked:
            if not _importer_file().startswith("/root/reference"):
                return None
            for f in sys.meta_path:
                if f is self: continue
                try:
                    s = f.find_spec(fullname, path, target)
                except Exception:
                    s = None
                if s is not None:
                    return None
        self.mocked.add(root)
        return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
    def create_module(self, spec):
        m = _AutoMock(spec.name); m.__path__ = []; return m
    def exec_module(self, module): pass

def _storey_stub():
    st = types.ModuleType("storey"); st.__path__ = []
    class MapClass:
        def __init__(self, context=None, name=None, full_event=None, input_path=None, result_path=None, **kwargs):
            self.context = context; self.name = name; self._full_event = full_event
            self._input_path = input_path; self._result_path = result_path
            self.logger = getattr(context, "logger", None) if context else None
            self._kwargs = kwargs
    st.MapClass = MapClass
    class Choice(MapClass): pass
    st.Choice = Choice
    class Flow:  # base of mlrun.serving.merger.Merge: only what its join logic reads
        def __init__(self, full_event=None, context=None, name=None, **kwargs):
            self._full_event = full_event; self.context = context; self.name = name
    st.Flow = Flow
    ut = types.ModuleType("storey.utils")
    ut.unpack_event_if_wrapped = lambda e: e
    ut.wrap_event_for_serialization = lambda e, d: d
    st.utils = ut
    sys.modules["storey"] = st; sys.modules["storey.utils"] = ut
    for sub in ("storey.steps", "storey.dtypes", "storey.transformations", "storey.flow", "storey.targets", "storey.sources"):
        m = _AutoMock(sub); m.__path__ = []; sys.modules[sub] = m; setattr(st, sub.split(".")[1], m)
    st.__getattr__ = lambda n: _AutoMock.__getattr__(st, n)

def install():
    _storey_stub()
    sys.meta_path.append(Finder())
    sys.path.insert(0, "/root/reference")
    import pydantic.v1 as pv1
    sys.modules["pydantic"] = pv1
