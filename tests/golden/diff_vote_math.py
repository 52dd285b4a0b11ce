"""Differential check of the vote arithmetic and the dotted-path getter against the REAL reference (build container only):
VotingEnsemble._majority_vote / _mean_vote on 3 000 random prediction tables (ints, floats, integral floats, mixed; random
weights incl. zeros) and mlrun.utils.helpers.get_in on 20 000 random nested objects x paths.

    python -m tests.golden.diff_vote_math
"""
import sys, random
import numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.golden import _refshim
_refshim.install()
from mlrun.serving import routers as ref
from mlrun.utils import helpers as refh
from mlrun_b200.serving import routing as mine
from mlrun_b200.serving import paths as mpaths
rnd = random.Random(1); rng = np.random.default_rng(1)
def mk(mod):
    o = mod.VotingEnsemble.__new__(mod.VotingEnsemble)
    o.name="ens"; o.vote_type=None; o.vote_flag=False; o.weights=None; o._weights = {}
    return o
n=0
for it in range(3000):
    M = rnd.randint(1,6); B = rnd.randint(1,7)
    kind = rnd.choice(["int","float","mixed","intfloat"])
    if kind=="int": preds = [[rnd.randint(0,4) for _ in range(B)] for _ in range(M)]
    elif kind=="float": preds = [[rnd.uniform(-3,3) for _ in range(B)] for _ in range(M)]
    elif kind=="intfloat": preds = [[float(rnd.randint(0,4)) for _ in range(B)] for _ in range(M)]
    else: preds = [[rnd.choice([rnd.randint(0,3), rnd.uniform(0,3)]) for _ in range(B)] for _ in range(M)]
    w = np.array([rnd.choice([1.0/M, rnd.random(), 0.0, 1.0]) for _ in range(M)])
    for fn in ("_majority_vote", "_mean_vote"):
        a, b = mk(ref), mk(mine)
        try: ra = ("ok", getattr(a, fn)(preds, w))
        except Exception as e: ra = ("err", type(e).__name__)
        try: rb = ("ok", getattr(b, fn)(preds, w))
        except Exception as e: rb = ("err", type(e).__name__)
        n += 1
        if repr(ra) != repr(rb):
            print("DIFF", fn, preds, w, ra, rb); sys.exit(1)
print("vote math identical on", n)
# get_in / update_in
def rand_obj(d=0):
    if d>2 or rnd.random()<0.3: return rnd.choice([1, "x", None, 2.5, [1,2], {}])
    return {rnd.choice("abc"): rand_obj(d+1) for _ in range(rnd.randint(0,3))}
n=0
for it in range(20000):
    obj = rand_obj()
    key = ".".join(rnd.choice("abcd") for _ in range(rnd.randint(1,3)))
    for fname in ("get_in",):
        fa = getattr(refh, fname); fb = getattr(mpaths, fname, None)
        if fb is None: print("no", fname, "in paths:", dir(mpaths)); sys.exit(0)
        try: ra=("ok", fa(obj,key, "DEF"))
        except Exception as e: ra=("err", type(e).__name__)
        try: rb=("ok", fb(obj,key, "DEF"))
        except Exception as e: rb=("err", type(e).__name__)
        n+=1
        if repr(ra)!=repr(rb): print("DIFF get_in", obj, key, ra, rb); sys.exit(1)
print("get_in identical on", n)
