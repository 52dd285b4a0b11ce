"""mlrun_b200.serving -- drop-in names of mlrun.serving, backed by the B200 engine"""
from .device_models import (  # noqa: F401
    FeatureRowModelServer,
    FeatureRowVotingEnsemble,
    LGBMModelServer,
    PickleModelServer,
    SKLearnModelServer,
    XGBoostModelServer,
)
from .events import MockEvent, MockTrigger, Response  # noqa: F401
from .function import ServingFunction, ServingRuntime, new_function  # noqa: F401
from .graph import (  # noqa: F401
    ErrorStep,
    FlowStep,
    QueueStep,
    RootFlowStep,
    RouterStep,
    StepKinds,
    TaskStep,
)
from .host import GraphContext, GraphServer, create_graph_server, nuclio_init_hook, v2_serving_handler, v2_serving_init  # noqa: F401
from .model_server import V2ModelServer  # noqa: F401
from .resolve import GraphError, MLRunInvalidArgumentError  # noqa: F401
from .routing import ModelRouter, ParallelRun, ParallelRunnerModes, VotingEnsemble, VotingTypes  # noqa: F401
