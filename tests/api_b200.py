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
