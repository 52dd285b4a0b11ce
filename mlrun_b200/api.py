"""One flat namespace with the reference's plugin names (what `tests/scenarios.py`, `bench.py` and
`mlrun_b200.synthetic` workloads build graphs with)."""

from .feature_store.transforms import (  # noqa: F401
    DateExtractor,
    DropFeatures,
    FeaturesetValidator,
    Imputer,
    MapValues,
    OneHotEncoder,
    SetEventMetadata,
    _Step as MapClass,
)
from .feature_store.online import FeatureVector, get_feature_vector, register_feature_vector  # noqa: F401
from .serving.merger import Merge  # noqa: F401
from .serving.routing import EnrichmentModelRouter, EnrichmentVotingEnsemble  # noqa: F401
from .serving import (  # noqa: F401
    FeatureRowModelServer,
    FeatureRowVotingEnsemble,
    GraphContext,
    GraphServer,
    LGBMModelServer,
    MockEvent,
    MockTrigger,
    ModelRouter,
    ParallelRun,
    RouterStep,
    SKLearnModelServer,
    TaskStep,
    V2ModelServer,
    VotingEnsemble,
    XGBoostModelServer,
    create_graph_server,
    new_function,
)

NAME = "mlrun_b200"
