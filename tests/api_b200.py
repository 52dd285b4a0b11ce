"""`api` namespace over the product (mlrun_b200), see tests/scenarios.py"""

import json
import os

from mlrun_b200.api import *  # noqa: F401,F403
from mlrun_b200.api import GraphContext
from mlrun_b200.serving import host


def init_from_spec(spec, namespace):
    os.environ[host.SERVING_SPEC_ENV] = json.dumps(spec)
    context = GraphContext()
    context.is_mock = True
    host.nuclio_init_hook(context, namespace, "serving_v2")
    return context


def validator_step(rules, columns):
    from mlrun_b200.feature_store import transforms as _t
    from mlrun_b200.feature_store.ingest import MinMaxValidator as _MinMax

    return _t.FeaturesetValidator(columns=columns, validators={c: _MinMax(**kw) for c, kw in rules.items()
                                                                 if not columns or c in columns})
