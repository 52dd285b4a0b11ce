"""Real-time feature enrichment backed by a device-resident online table.

Plugin-API mirror of the parts of mlrun.feature_store the enrichment routers touch: `get_feature_vector(uri)
.get_online_feature_service(impute_policy=...)` -> `OnlineVectorService.get(entity_rows, as_list)`
(mlrun/feature_store/feature_vector.py:903-1067).  The reference reads the online (NoSQL) store once per entity row
through a storey graph; here the vector's online rows live in HBM (`b2s_table_*`, include/b200serve.h) and a batch of
keys is resolved by one kernel launch.  The feature-store control plane (vector definition, targets, stats
calculation jobs) is out of scope: a `FeatureVector` is built from an in-memory frame and registered by uri.

Values are float32 on the device; impute values (constants or "$mean"-style statistics) are rounded to float32 when the
service is initialised, and the stats table computed here is float32-valued, so `get()` returns the same numbers on
both sides of the C-ABI.  String entity keys are hashed to 64 bits (FNV-1a, collisions among the table's keys are
rejected when the table is built; an unknown key matching a stored hash has probability ~ n_keys / 2^64).
"""

import ctypes as C

import numpy as np

from .. import _native as nat
from ..serving.resolve import MLRunInvalidArgumentError

_REGISTRY = {}


def register_feature_vector(uri, vector):
    _REGISTRY[uri] = vector


def get_feature_vector(uri, project=None):
    """mlrun.feature_store.get_feature_vector (feature_store/api.py:73-97) over the process-local registry"""
    try:
        return _REGISTRY[uri]
    except KeyError:
        raise MLRunInvalidArgumentError(f"feature vector {uri!r} is not registered (register_feature_vector)")


def _hash_strings(values):
    data = [str(v).encode() for v in values]
    offsets = np.zeros(len(data) + 1, dtype=np.int64)
    np.cumsum([len(d) for d in data], out=offsets[1:])
    keys = np.empty(len(data), dtype=np.int64)
    nat.check(nat.load().b2s_hash_strings(b"".join(data), offsets.ctypes.data_as(C.POINTER(C.c_int64)), len(data),
                                          keys.ctypes.data_as(C.POINTER(C.c_int64))))
    return keys


class DeviceTable:
    """thin wrapper over b2s_table_* (64-bit keys -> rows of float32 features, optional impute vector)"""

    def __init__(self, keys, values, impute=None):
        nat.init()
        self._lib = nat.load()
        keys = np.ascontiguousarray(keys, dtype=np.int64)
        values = np.ascontiguousarray(values, dtype=np.float32)
        self.n_keys, self.n_feat = values.shape
        imp = None if impute is None else np.ascontiguousarray(impute, dtype=np.float32)
        self._h = C.c_void_p()
        nat.check(self._lib.b2s_table_create(keys.ctypes.data_as(C.POINTER(C.c_int64)), len(keys), nat._p(values, C.c_float),
                                             self.n_feat, nat._p(imp, C.c_float), C.byref(self._h)))

    def lookup(self, keys, with_stats=False):
        keys = np.ascontiguousarray(keys, dtype=np.int64)
        rows = np.empty((len(keys), self.n_feat), dtype=np.float32)
        found = np.empty(len(keys), dtype=np.int32)
        stats = nat.Stats()
        nat.check(self._lib.b2s_table_lookup_host(self._h, keys.ctypes.data_as(C.POINTER(C.c_int64)), len(keys),
                                                  nat._p(rows, C.c_float), nat._p(found, C.c_int32), C.byref(stats)))
        return (rows, found.astype(bool), stats.as_dict()) if with_stats else (rows, found.astype(bool))

    def enrich(self, plan, keys, with_stats=False):
        """keys -> gather -> `plan` (a finalized DevicePlan over this table's features) -> (outputs, status) on the host, in
        one C call (b2s_table_enrich_host); status carries ROW_UNKNOWN_KEY for keys that are not in the table.  The two
        arrays share one pinned block of the pool (PCIe-speed D2H, no second copy); it is reused once they are collected."""
        keys = np.ascontiguousarray(keys, dtype=np.int64)
        n = len(keys)
        out_b = (n * plan.out_cols * 4 + 63) // 64 * 64
        block = nat.PINNED.take(out_b + n * 4) if n else None
        if block is not None:
            out = np.frombuffer(block, dtype=plan.out_dtype, count=n * plan.out_cols).reshape(n, plan.out_cols)
            status = np.frombuffer(block, dtype=np.int32, count=n, offset=out_b)
        else:
            out = np.empty((n, plan.out_cols), dtype=plan.out_dtype)
            status = np.empty(n, dtype=np.int32)
        stats = nat.Stats() if with_stats else None
        nat.check(self._lib.b2s_table_enrich_host(self._h, plan._h, keys.ctypes.data_as(C.POINTER(C.c_int64)), n, out.ctypes.data,
                                                  out.nbytes, nat._p(status, C.c_int32), C.byref(stats) if with_stats else None))
        return (out, status, stats.as_dict()) if with_stats else (out, status)

    def enrich_device(self, plan, d_keys, n, d_out, d_status=None, stream=None):
        """device keys -> outputs (+ status) in ONE launch: the scoring kernel gathers its rows from the table
        (b2s_table_enrich_device).  Returns False when the plan is not covered by the gather loader (nothing was launched:
        use lookup_device + plan.run_device)."""
        rc = self._lib.b2s_table_enrich_device(self._h, plan._h, d_keys, int(n), d_out, d_status, stream)
        if rc == -6:  # B2S_ERR_UNSUPPORTED
            return False
        nat.check(rc)
        return True

    def lookup_device(self, d_keys, n, d_rows, row_stride, d_found=None, stream=None):
        nat.check(self._lib.b2s_table_lookup_device(self._h, d_keys, int(n), d_rows, int(row_stride), d_found, stream))

    def time_device(self, d_key_ptrs, n, d_rows, row_stride, iters):
        arr = (C.c_void_p * len(d_key_ptrs))(*d_key_ptrs)
        ms = C.c_float()
        nat.check(self._lib.b2s_table_time_device(self._h, arr, len(d_key_ptrs), int(n), d_rows, int(row_stride), None,
                                                  int(iters), C.byref(ms)))
        return ms.value

    def close(self):
        if self._h:
            self._lib.b2s_table_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FeatureVector:
    """name, requested features, index (entity) keys, label column, and the online rows as a frame whose index (or
    `index_keys` columns) holds the entity keys"""

    def __init__(self, name, features, index_keys, frame, stats=None, label_column=None, with_indexes=False):
        self.name = name
        self.features = list(features)
        self.index_keys = list(index_keys)
        self.label_column = label_column
        self.with_indexes = with_indexes
        if all(k in frame.columns for k in self.index_keys):
            frame = frame.set_index(self.index_keys)
        self.frame = frame
        self._stats = stats

    def get_stats_table(self):
        """feature statistics (mean / min / max / std / count per feature), float32-valued"""
        if self._stats is None:
            import pandas as pd

            cols = [f for f in self.features if f != self.label_column]
            vals = self.frame[cols].to_numpy(dtype=np.float32).copy()
            vals[~np.isfinite(vals)] = np.nan  # statistics of the finite observations
            with np.errstate(all="ignore"):
                stats = {
                    "mean": np.nanmean(vals.astype(np.float64), axis=0).astype(np.float32),
                    "min": np.nanmin(vals, axis=0), "max": np.nanmax(vals, axis=0),
                    "std": np.nanstd(vals.astype(np.float64), axis=0, ddof=1).astype(np.float32),
                    "count": np.sum(~np.isnan(vals), axis=0).astype(np.float32),
                }
            self._stats = pd.DataFrame({k: [float(x) for x in v] for k, v in stats.items()}, index=cols)
        return self._stats

    def get_online_feature_service(self, impute_policy=None, **kwargs):
        svc = OnlineVectorService(self, impute_policy)
        svc.initialize()
        return svc


class OnlineVectorService:
    """feature_vector.py:903-1067 with the store read replaced by the device table"""

    def __init__(self, vector, impute_policy=None):
        self.vector = vector
        self.impute_policy = impute_policy or {}
        self._index_columns = vector.index_keys
        self._requested_columns = vector.features
        self._columns = [c for c in vector.features if c != vector.label_column]
        self._impute_values = {}
        self.table = None
        self._string_keys = False
        self._label_alive = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def status(self):
        return "ready"

    def initialize(self):
        """impute values (feature_vector.py:935-968), then the online rows go to the device"""
        self._impute_values = self._resolve_policy(dict(self.impute_policy)) if self.impute_policy else {}
        frame = self.vector.frame
        values = frame[self._columns].to_numpy(dtype=np.float32)
        keys = self._encode_keys(frame.index, build=True)
        # the label is not an online feature, but in the reference a truthy label keeps an all-zero row from being reported as
        # missing (the `any(data.values())` quirk): remember which entities have one
        label = self.vector.label_column
        self._label_alive = None
        if label and label in frame.columns:
            col = frame[label]
            self._label_alive = np.sort(keys[(col.notna() & col.map(bool)).to_numpy(dtype=bool)])  # a missing label is no label
        impute = np.array([self._impute_values.get(c, np.nan) for c in self._columns], dtype=np.float32)
        self.table = DeviceTable(keys, values, impute if self._impute_values else None)

    def _resolve_policy(self, policy):
        """{feature | "*": constant | "$stat"} -> {feature: float32 value held on the device} (feature_vector.py:935-968)"""
        stats = self.vector.get_stats_table()

        def value_of(feature, spec):
            v = stats.loc[feature, spec[1:]] if isinstance(spec, str) and spec.startswith("$") else spec
            v = float(np.float32(v))
            if not np.isfinite(v):
                raise MLRunInvalidArgumentError(f"impute value of feature {feature} is not finite ({v}): not held on the device")
            return v

        default = policy.pop("*", None)
        unknown = [f for f in policy if f not in self._columns]
        if unknown:
            raise MLRunInvalidArgumentError(f"feature {unknown[0]} in impute_policy but not in feature vector")
        values = {} if default is None else {f: value_of(f, default) for f in self._columns if f not in policy}
        values.update({f: value_of(f, spec) for f, spec in policy.items()})
        return values

    def _encode_keys(self, raw, build=False):
        """entity keys -> int64: integers as they are, everything else through the 64-bit string hash"""
        import pandas as pd

        if isinstance(raw, pd.MultiIndex) or (len(raw) and isinstance(raw[0], (tuple, list))):
            raw = [".".join(str(x) for x in k) for k in raw]
            strings = True
        else:
            arr = np.asarray(raw)
            strings = arr.dtype.kind not in "iu"
            raw = arr
        if build:
            self._string_keys = strings
        elif strings != self._string_keys:
            raw, strings = ([str(x) for x in raw], True) if self._string_keys else (raw, strings)
            if not self._string_keys:
                raise MLRunInvalidArgumentError("the online table has integer entity keys")
        if not strings:
            return np.asarray(raw, dtype=np.int64)
        keys = _hash_strings(raw)
        if build and len(np.unique(keys)) != len(keys):
            raise MLRunInvalidArgumentError("two entity keys share a 64-bit hash (or a key is duplicated)")
        return keys

    def _has_truthy_label(self, key):
        alive = self._label_alive
        if alive is None or not len(alive):
            return False
        j = int(np.searchsorted(alive, key))
        return j < len(alive) and alive[j] == key

    # ---- batched engine surface --------------------------------------------------------------------
    def get_matrix(self, keys):
        """entity keys (one per row; tuples for composite keys) -> ((B, F) float32 imputed rows, found mask)"""
        return self.table.lookup(self._encode_keys(keys))

    # ---- the reference's call ---------------------------------------------------------------------------
    def get(self, entity_rows, as_list=False):
        """feature_vector.py:975-1067"""
        if isinstance(entity_rows, dict):
            entity_rows = [entity_rows]
        if not entity_rows or not isinstance(entity_rows, list) or not isinstance(entity_rows[0], (list, dict)):
            raise MLRunInvalidArgumentError(
                f"input data is of type {type(entity_rows)}. must be a list of lists or list of dicts")
        idx = self._index_columns
        if isinstance(entity_rows[0], list):
            if not idx or len(entity_rows[0]) != len(idx):
                raise MLRunInvalidArgumentError("input list must be in the same size of the index_keys list")
            entity_rows = [{idx[i]: item[i] for i in range(len(idx))} for item in entity_rows]
        keys = [row[idx[0]] if len(idx) == 1 else tuple(row[k] for k in idx) for row in entity_rows]
        encoded = self._encode_keys(keys)
        rows, found = self.table.lookup(encoded)
        results = []
        for i, row in enumerate(entity_rows):
            if not found[i]:
                # the graph returned only the entity columns (:1030-1034) -- unless the row carried more than the keys
                if all(col in idx for col in row):
                    results.append(None)
                    continue
                vals = [None] * len(self._columns)
            else:
                vals = rows[i].tolist()  # a stored NaN without an impute value stays NaN (:1046-1052)
            data = dict(row)
            if not found[i] and self._impute_values:
                vals = [self._impute_values.get(c, v) for c, v in zip(self._columns, vals)]
            data.update(zip(self._columns, vals))
            if not self.vector.with_indexes:
                for name in self.vector.index_keys:
                    data.pop(name, None)
            if not any(data.values()) and not (found[i] and self._has_truthy_label(encoded[i])):
                data = None
            if as_list and data is not None:
                data = [data.get(key, None) for key in self._requested_columns if key != self.vector.label_column]
            results.append(data)
        return results

    def close(self):
        if self.table is not None:
            self.table.close()
            self.table = None
