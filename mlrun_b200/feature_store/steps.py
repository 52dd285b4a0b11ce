"""drop-in import path: `from mlrun_b200.feature_store.steps import Imputer, OneHotEncoder, ...`"""
from .transforms import (  # noqa: F401
    DateExtractor,
    DropFeatures,
    FeaturesetValidator,
    Imputer,
    MapValues,
    OneHotEncoder,
    SetEventMetadata,
)
from .ingest import MinMaxValidator  # noqa: F401,E402
