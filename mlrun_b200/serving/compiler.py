"""Graph compiler: lower a whole serving graph into one DevicePlan.

Recognised topologies (everything else raises LoweringError -- the batched path never falls back to
per-event Python):

    flow:   [Imputer | OneHotEncoder | MapValues | DropFeatures]*  ->  model server | router
    router: ModelRouter (single route selected per call) | VotingEnsemble over device model servers

The lowering follows the storey-engine semantics of the steps (see mlrun_b200.lowering) and the vote
semantics of VotingEnsemble (explicit vote_type, or decided from the models' links: classifiers ->
majority, regressors -> mean; the reference infers it from the first request's values, which for
scikit-learn models is the same thing).
"""

import itertools

import numpy as np

from .. import _native as nat
from ..lowering import ColumnProgram, LoweringError
from .device_models import PickleModelServer
from .routing import ModelRouter, VotingEnsemble

_TRANSFORMS = ("Imputer", "OneHotEncoder", "MapValues", "DropFeatures")


class CompiledGraph:
    def __init__(self, plan, program, in_names, responder, tracker=None):
        self.plan = plan
        self.program = program
        self.in_names = list(in_names)
        self.responder = responder  # (model_name, version) used to shape per-event responses
        self.tracker = tracker      # the responder's _ModelLogPusher when model tracking is on (else None)

    def pack_events(self, bodies):
        """feature-dict bodies -> (B, F) float32 rows; a None value is a missing value (NaN), as pd.isna treats it"""
        names = tuple(self.in_names)
        for i, body in enumerate(bodies):
            if tuple(body) != names:
                raise ValueError(f"event {i} does not carry the compiled schema {self.in_names[:4]}...")
        n, width = len(bodies), len(names)
        try:  # one pass over all values (2x the row-by-row assignment); float64 first: the same single rounding to float32
            flat = np.fromiter(itertools.chain.from_iterable(map(dict.values, bodies)), dtype=np.float64, count=n * width)
        except TypeError:  # a None (or a mapping that is not a dict) somewhere
            flat = np.array([np.nan if v is None else v for body in bodies for v in body.values()], dtype=np.float64)
        return flat.astype(np.float32).reshape(n, width)

    def responses(self, out, status, context):
        name, version = self.responder
        head = {"model_name": name}
        tail = {"model_version": version} if version else {}
        rows, flagged = out.tolist(), status.tolist()
        res = []
        for row, bad in zip(rows, flagged):
            if bad:
                res.append(context.Response(body="ValueError: Input X contains NaN or infinity.", content_type="text/plain", status_code=400))
            else:
                res.append({**head, "outputs": row, **tail})
        return res


def _chain(flow):
    """the single start -> ... -> end chain of a flow (no branches)"""
    starts, _final, _resp = flow.check_and_process_graph()
    if len(starts) != 1:
        raise LoweringError("only single-entry flows are lowered")
    chain, cur = [], starts[0]
    while cur is not None:
        chain.append(cur)
        nxt = cur.next or []
        if len(nxt) > 1:
            raise LoweringError(f"step {cur.name} branches; branches are not lowered")
        cur = flow[nxt[0]] if nxt else None
    return chain


def _transform_object(step):
    obj = getattr(step, "_object", None)
    if obj is not None:
        return obj
    raise LoweringError(f"step {step.name} is not initialised (call init_object first)")


def _model_pack(step):
    obj = getattr(step, "_object", None)
    if not isinstance(obj, PickleModelServer):
        raise LoweringError(f"route/step {step.name}: {type(obj).__name__} is not a device model server")
    return obj.packed


def _vote_of(router_obj, packs):
    classifiers = [packing_is_clf(k, p) for k, p in packs]
    if any(classifiers) and not all(classifiers):
        raise LoweringError("an ensemble mixing classifiers and regressors is not lowered")
    vt = router_obj.vote_type
    vt = getattr(vt, "value", vt)
    if vt is None:
        vt = "classification" if all(classifiers) else "regression"
    weights = [router_obj._weights[name] for name in router_obj.routes.keys()]
    return (nat.VOTE_MAJORITY if vt == "classification" else nat.VOTE_MEAN), weights


def packing_is_clf(kind, packed):
    link = packed["link"] if kind == "linear" else packed.link
    return link != nat.LINK_IDENTITY


def compile_graph(graph, in_names=None, route=None):
    """graph: an initialised RootFlowStep or RouterStep.  route: for a ModelRouter, which model to lower"""
    transforms, terminal = [], None
    if graph.kind == "router":
        terminal = graph
    else:
        for step in _chain(graph):
            if step.kind == "router":
                terminal = step
                break
            obj = _transform_object(step)
            if type(obj).__name__ in _TRANSFORMS:
                if terminal is not None:
                    raise LoweringError("transform steps after the model are not lowered")
                transforms.append(obj)
            elif isinstance(obj, PickleModelServer):
                terminal = step
                break
            else:
                raise LoweringError(f"step {step.name} ({type(obj).__name__}) is not lowerable")
        if terminal is not None and terminal.next:
            raise LoweringError("steps after the model / router are not lowered")

    models, vote, responder, tracked = [], None, ("", ""), None
    if terminal is not None and terminal.kind == "router":
        robj = terminal._object
        routes = list(terminal.routes.values())
        if isinstance(robj, VotingEnsemble) and route is None:
            models = [_model_pack(r) for r in routes]
            vote = _vote_of(robj, models)
            responder, tracked = (robj.name, robj.version), robj
        elif isinstance(robj, (ModelRouter, VotingEnsemble)):
            key = route or list(terminal.routes.keys())[0]
            models = [_model_pack(terminal.routes[key])]
            mobj = terminal.routes[key]._object
            responder, tracked = (mobj.name, mobj.version), mobj
        else:
            raise LoweringError(f"router class {type(robj).__name__} is not lowerable")
    elif terminal is not None:
        models = [_model_pack(terminal)]
        responder, tracked = (terminal._object.name, terminal._object.version), terminal._object

    if in_names is None:
        n = _n_inputs(transforms, models)
        in_names = [f"f{i}" for i in range(n)]
    program = ColumnProgram(in_names)
    for t in transforms:
        program.apply(t)
    plan = program.build_plan(models, vote)
    return CompiledGraph(plan, program, in_names, responder, getattr(tracked, "_model_logger", None))


def _n_inputs(transforms, models):
    if transforms:
        raise LoweringError("a graph with feature steps needs the input column names (run_batch(X, names=[...]))")
    widths = []
    for kind, packed in models:
        if kind == "linear":
            widths.append(packed["W"].shape[1])
        else:
            widths.append(getattr(packed, "n_features", None) or int(packed.feature.max()) + 1)
    if len(set(widths)) != 1:
        raise LoweringError(f"the router's models take different input widths {sorted(set(widths))}")
    return widths[0]
