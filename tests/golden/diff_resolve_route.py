"""Differential check of ModelRouter / VotingEnsemble._resolve_route against the REAL reference classes (needs /root/reference,
so it runs in the build container only, like gen_golden.py): every URL of up to four segments over a vocabulary of model
names, versions and operations x eleven body shapes x four router names -- about 257 000 cases (a seeded 5 % sample of the four-segment URLs), results, exceptions (type and
text) and the log_router side effect compared.  Last run: identical on all cases.

    python -m tests.golden.diff_resolve_route
"""
import sys, itertools, random

random.seed(0)
sys.path.insert(0, '/root/repo')
from tests.golden import _refshim
_refshim.install()
from mlrun.serving import routers as ref
sys.path.insert(0, '/root/repo')
from mlrun_b200.serving import routing as mine

class R(dict):
    pass
routes = {"m1": "R1", "m2": "R2", "m1:v2": "R1v2", "infer": "Rinfer", "ens": "Rens"}

def mk(mod, cls, name):
    o = cls.__new__(cls)
    o.name = name; o.routes = dict(routes); o.url_prefix = "/v2/models"; o.log_router = "unset"
    return o
segs = ["", "m1", "m2", "m3", "versions", "v2", "infer", "predict", "explain", "ens", "x", "metrics"]
urls = [None, "", "/"]
for n in range(0, 5):
    for combo in itertools.product(segs, repeat=n):
        if n > 3 and random.random() > 0.05: continue
        urls.append("/v2/models" + "".join("/" + c for c in combo))
        urls.append("/v2/models/" + "/".join(combo) + "/")
bodies = [None, "str", b"x", [1], {}, {"model": "m2"}, {"operation": "predict"}, {"model": "m3", "operation": "explain"}, {"model": "m1:v2"}, {"operation": None}, {"operation": ""}]
cnt = 0
for clsname, name in (("ModelRouter", None), ("VotingEnsemble", "ens"), ("VotingEnsemble", "vote"), ("VotingEnsemble", "m1")):
    for u in urls:
        for b in bodies:
            a = mk(ref, getattr(ref, clsname), name); c = mk(mine, getattr(mine, clsname), name)
            try: ra = ("ok", a._resolve_route(b, u), a.log_router)
            except Exception as e: ra = ("err", type(e).__name__, str(e))
            try: rc = ("ok", c._resolve_route(b, u), c.log_router)
            except Exception as e: rc = ("err", type(e).__name__, str(e))
            cnt += 1
            if ra != rc:
                print("DIFF", clsname, name, repr(u), b, ra, rc); sys.exit(1)
print("identical on", cnt, "cases")
