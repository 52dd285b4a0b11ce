"""Seeded synthetic workloads for the BASELINE.json configs.

Used by tests/, bench.py and __graft_entry__.smoke().  Pure numpy + scikit-learn (model fitting
only) -- nothing here touches the GPU or the oracle; a workload only *describes* a serving graph
(steps, mappings, fitted models, input matrix) so the same description can be built on any
implementation of the plugin API (`build_server(api)`).
"""

from dataclasses import dataclass, field

import numpy as np


def feature_names(n):
    return [f"f{i}" for i in range(n)]


@dataclass
class Flow3Workload:
    """Imputer -> OneHotEncoder -> {linear model | VotingEnsemble of linear models} (configs[1] / metric)"""

    X: np.ndarray  # (B, F) float32, NaNs in numeric cols, small-int codes (as f32) in categorical cols
    names: list
    num_cols: list
    cat_cols: list
    categories: list  # category values per categorical col (ints)
    impute_mapping: dict  # {name: fill}
    impute_default: object  # int 0: an imputed categorical must stay an int code (see rows_as_dicts)
    onehot_mapping: dict  # {name: [cats]}
    coefs: list  # per model: (F_out,) float64
    intercepts: list  # per model: float
    out_names: list = field(default_factory=list)

    @property
    def n_models(self):
        return len(self.coefs)

    def sklearn_models(self):
        from sklearn.linear_model import LinearRegression

        models = []
        for w, b in zip(self.coefs, self.intercepts):
            m = LinearRegression()
            m.coef_ = np.asarray(w, dtype=np.float64)
            m.intercept_ = float(b)
            m.n_features_in_ = len(w)
            models.append(m)
        return models

    def rows_as_dicts(self, limit=None):
        rows = []
        X = self.X if limit is None else self.X[:limit]
        cat = set(self.cat_cols)
        for r in X:
            row = {}
            for n, v in zip(self.names, r):
                v = float(v)
                # categorical columns carry *integer* codes in dict events: the reference's OneHotEncoder
                # adds a stray "<col>_2.0" key when a float 2.0 matches the int category 2 (steps.py:461-464)
                row[n] = int(v) if (n in cat and v == v and v.is_integer()) else v
            rows.append(row)
        return rows

    def build_graph(self, api, engine="sync", executor="array"):
        fn = api.new_function("flow3", kind="serving")
        graph = fn.set_topology("flow", engine=engine)
        step = graph.to(api.Imputer(mapping=dict(self.impute_mapping), default_value=self.impute_default), name="imputer")
        step = step.to(api.OneHotEncoder(mapping={k: list(v) for k, v in self.onehot_mapping.items()}), name="onehot")
        models = self.sklearn_models()
        if self.n_models == 1:
            step = step.to(api.FeatureRowModelServer(name="linear", model=models[0]), name="linear")
        else:
            step = step.to("*FeatureRowVotingEnsemble", name="ensemble", vote_type="regression", executor_type=executor)
            for i, m in enumerate(models):
                step.add_route(f"m{i + 1}", class_name="FeatureRowModelServer", model=m, model_path="")
        if engine != "sync":
            step.respond()
        return fn

    def build_server(self, api, engine="sync", executor="array", **kw):
        fn = self.build_graph(api, engine, executor)
        ns = {"FeatureRowVotingEnsemble": api.FeatureRowVotingEnsemble,
              "FeatureRowModelServer": api.FeatureRowModelServer}
        return fn.to_mock_server(namespace=ns, **kw)


def flow3_workload(n_rows=4096, n_num=56, n_cat=8, seed=2, n_models=1, cats_per=4, nan_frac=0.05, oov_frac=0.01):
    """§8(d) config 2: cols [0,n_num) ~N(0,1) with 5% NaN; cols [n_num, n_num+n_cat) integer codes
    0..cats_per-1 stored as f32, 1% out-of-vocabulary value 7, 1% NaN (imputed with the default 0).
    Imputer mapping = column mean of the numeric cols; OneHot over all categorical cols; linear
    weights ~N(0,1) (seed+20)."""
    rng = np.random.default_rng(seed)
    F = n_num + n_cat
    X = np.empty((n_rows, F), dtype=np.float32)
    X[:, :n_num] = rng.normal(size=(n_rows, n_num)).astype(np.float32)
    X[:, n_num:] = rng.integers(0, cats_per, size=(n_rows, n_cat)).astype(np.float32)
    mrng = np.random.default_rng(seed * 10 + 1)
    nan_mask = mrng.random((n_rows, n_num)) < nan_frac
    kept = np.where(nan_mask, 0.0, X[:, :n_num].astype(np.float64))
    means = kept.sum(axis=0) / np.maximum((~nan_mask).sum(axis=0), 1)  # column mean of the observed values
    X[:, :n_num][nan_mask] = np.nan
    oov = mrng.random((n_rows, n_cat)) < oov_frac
    X[:, n_num:][oov] = 7.0
    cat_nan = mrng.random((n_rows, n_cat)) < 0.01
    X[:, n_num:][cat_nan] = np.nan

    names = feature_names(F)
    num_cols, cat_cols = names[:n_num], names[n_num:]
    # fills are rounded to f32 so that host dicts / device tables hold identical values
    impute_mapping = {n: float(np.float32(m)) for n, m in zip(num_cols, means)}
    cats = list(range(cats_per))
    onehot_mapping = {n: list(cats) for n in cat_cols}
    out_names = list(num_cols) + [f"{c}_{k}" for c in cat_cols for k in cats]
    wrng = np.random.default_rng(seed + 20)
    coefs = [wrng.normal(size=len(out_names)) for _ in range(n_models)]
    intercepts = [float(wrng.normal()) for _ in range(n_models)]
    return Flow3Workload(X, names, num_cols, cat_cols, [list(cats)] * n_cat, impute_mapping, 0, onehot_mapping,
                         coefs, intercepts, out_names)


@dataclass
class TreeWorkload:
    """ModelRouter / VotingEnsemble over sklearn tree-ensemble scorers (configs[2])"""

    X: np.ndarray
    models: list
    kind: str

    def build_server(self, api, executor="array", **kw):
        fn = api.new_function("trees", kind="serving")
        graph = fn.set_topology("router", api.VotingEnsemble(vote_type=self.kind, executor_type=executor))
        for i, m in enumerate(self.models):
            graph.add_route(f"m{i + 1}", class_name="SKLearnModelServer", model=m, model_path="")
        return fn.to_mock_server(namespace={"SKLearnModelServer": api.SKLearnModelServer}, **kw)


def tree_workload(n_rows=16384, n_feat=128, n_models=4, n_trees=100, depth=6, seed=3, kind="regression",
                  n_fit=2000, max_features=None):
    """§8(d) config 3: X ~N(0,1); models = GradientBoosting{Regressor,Classifier}(n_trees, depth,
    random_state=30+i) fit on synthetic rows y = 2*x0 + sin(x1) + x2*x3 + eps (3-class: terciles)."""
    from sklearn.ensemble import GradientBoostingClassifier, GradientBoostingRegressor

    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n_rows, n_feat)).astype(np.float32)
    frng = np.random.default_rng(seed + 100)
    Xf = frng.normal(size=(n_fit, n_feat)).astype(np.float32)
    y = 2 * Xf[:, 0] + np.sin(Xf[:, 1]) + Xf[:, 2] * Xf[:, 3] + 0.1 * frng.normal(size=n_fit)
    models = []
    for i in range(n_models):
        if kind == "regression":
            m = GradientBoostingRegressor(n_estimators=n_trees, max_depth=depth, random_state=30 + i,
                                          max_features=max_features, subsample=0.8)
            m.fit(Xf, y)
        else:
            labels = np.digitize(y, np.quantile(y, [1 / 3, 2 / 3]))
            m = GradientBoostingClassifier(n_estimators=n_trees, max_depth=depth, random_state=30 + i,
                                           max_features=max_features, subsample=0.8)
            m.fit(Xf, labels)
        models.append(m)
    return TreeWorkload(X, models, kind)


@dataclass
class IngestWorkload:
    """feature-set ingest graph (configs[4]); see mlrun_b200 ingest plan"""

    X: np.ndarray
    names: list
    steps: list


def events_matrix(n_rows, n_feat=64, seed=4):
    rng = np.random.default_rng(seed)
    return rng.normal(size=(n_rows, n_feat)).astype(np.float32)


@dataclass
class IngestWorkload:
    """feature-set ingest of SURVEY.md 8(d) config 5: 256 four-byte input slots per row, six steps"""

    df: object
    f32_cols: list
    cat_cols: list
    counter_cols: list
    means: dict
    mapped_cols: list
    onehot_cols: list
    checked_cols: list

    RANGES = {0: ["-inf", -1.0], 1: [-1.0, 0.0], 2: [0.0, 1.0], 3: [1.0, "inf"]}

    def build_steps(self, api):
        """`api` provides the step classes (mlrun_b200.feature_store.steps / oracle.transforms) + MinMaxValidator"""
        validators = {c: api.MinMaxValidator(severity="info", min=-2.5, max=2.5) for c in self.checked_cols}
        return [
            api.Imputer(mapping=dict(self.means)),
            api.MapValues(mapping={c: {"ranges": dict(self.RANGES)} for c in self.mapped_cols}, with_original_features=True),
            api.OneHotEncoder(mapping={c: list(range(8)) for c in self.onehot_cols}),
            api.DateExtractor(parts=["hour", "day_of_week"], timestamp_col="timestamp"),
            api.DropFeatures(features=list(self.mapped_cols)),
            api.FeaturesetValidator(validators=validators),
        ]

    @property
    def in_bytes_per_row(self):
        return 4 * (len(self.f32_cols) + len(self.cat_cols) + len(self.counter_cols) + 2)

    @property
    def out_bytes_per_row(self):
        n = len(self.mapped_cols) + (len(self.f32_cols) - len(self.mapped_cols)) + (len(self.cat_cols) - len(self.onehot_cols))
        n += 8 * len(self.onehot_cols) + len(self.counter_cols) + 2 + 2
        return 4 * n


def ingest_workload(n_rows=100_000, seed=5, n_f32=192, n_cat=48, n_counter=14, nan_frac=0.05):
    """192 float32 columns ~N(0,1) with 5 % NaN, 48 int32 categorical codes (cardinality 8, 1 % out-of-vocabulary 9),
    14 int32 counters, one datetime64[ns] timestamp (random seconds over 2015..2030).  Imputer fills = column means
    rounded to float32; ranges on the first 16 float columns; one-hot over the first 8 categorical columns;
    validators [-2.5, 2.5] on float columns 16..23."""
    import pandas as pd

    rng = np.random.default_rng(seed)
    data = {}
    f32_cols = [f"x{i}" for i in range(n_f32)]
    means = {}
    for c in f32_cols:
        a = rng.normal(size=n_rows).astype(np.float32)
        means[c] = float(np.float32(a.mean(dtype=np.float64)))
        a[rng.random(n_rows) < nan_frac] = np.nan
        data[c] = a
    cat_cols = [f"c{i}" for i in range(n_cat)]
    for c in cat_cols:
        a = rng.integers(0, 8, size=n_rows).astype(np.int32)
        a[rng.random(n_rows) < 0.01] = 9
        data[c] = a
    counter_cols = [f"n{i}" for i in range(n_counter)]
    for c in counter_cols:
        data[c] = rng.integers(0, 1 << 20, size=n_rows).astype(np.int32)
    secs = rng.integers(1_420_070_400, 1_893_456_000, size=n_rows)  # 2015-01-01 .. 2030-01-01
    data["timestamp"] = (secs * 1_000_000_000).astype("datetime64[ns]")
    df = pd.DataFrame(data)
    return IngestWorkload(df, f32_cols, cat_cols, counter_cols, means, f32_cols[:16], cat_cols[:8], f32_cols[16:24])
