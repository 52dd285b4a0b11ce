"""Body codec (host code of libb200serve.so): same values as json.loads + np.asarray, same text as json.dumps."""

import json
import math

import numpy as np
import pytest

from mlrun_b200.serving import codec


def _ref(body):
    return np.asarray(json.loads(body)["inputs"], dtype=np.float32)


def test_parse_matches_json_loads_on_random_bodies():
    rng = np.random.default_rng(11)
    for rows, cols in [(1, 1), (3, 7), (64, 64), (257, 5)]:
        X = rng.normal(size=(rows, cols)) * 10.0 ** rng.integers(-8, 9, size=(rows, cols))
        body = json.dumps({"id": "abc", "inputs": X.tolist(), "model": "m1"})
        got, (b, e) = codec.parse_inputs(body)
        np.testing.assert_array_equal(got, _ref(body))
        assert json.loads(body[b:e]) == X.tolist()
        X2, rest = codec.decode_body(body.encode())
        assert rest == {"id": "abc", "model": "m1"} and X2.shape == (rows, cols)


def test_number_grammar_and_rounding():
    texts = ["0", "-0", "5", "-17", "1.5", "0.1", "-2.5e-3", "1E5", "1e+5", "3.4028235e38", "3.4028236e38", "1e39", "-1e39",
             "1e-46", "4.9e-324", "1e999", "-1e999", "1e-999", "123456789012345678901234567890", "0.30000000000000004",
             "16777217", "1.00000005960464477539"]
    body = '{"inputs": [[' + ", ".join(texts) + "]]}"
    got, _ = codec.parse_inputs(body)
    with np.errstate(over="ignore"):
        want = _ref(body)
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))  # bit-exact, including -0.0
    # out-of-range floats written without an exponent, or with one that does not decide alone
    texts = ["0." + "0" * 400 + "1", "-0." + "0" * 400 + "1", "1" + "0" * 400 + ".5", "-1" + "0" * 400 + ".0", "1e-400", "-1e-400",
             "1e400", "0." + "0" * 10 + "1e-390", "1" + "0" * 20 + ".0e290", "0.001e-999", "123.456e-330", "0.0e999"]
    body = '{"inputs": [[' + ", ".join(texts) + "]]}"
    got, _ = codec.parse_inputs(body)
    with np.errstate(over="ignore"):
        want = _ref(body)
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    # an integer literal beyond the double range: np.asarray raises OverflowError, so the codec does not take the body
    with pytest.raises(OverflowError):
        _ref('{"inputs": [1' + "0" * 400 + "]}")
    with pytest.raises(codec.NotV2Matrix):
        codec.parse_inputs('{"inputs": [1' + "0" * 400 + "]}")
    body = '{"inputs": [NaN, Infinity, -Infinity, null, 2]}'
    got, _ = codec.parse_inputs(body)
    assert got.shape == (5, 1) and np.isnan(got[0, 0]) and got[1, 0] == np.inf and got[2, 0] == -np.inf and np.isnan(got[3, 0])


def test_layout_cases_and_whitespace():
    nl, tab = chr(10), chr(9)
    body = ('  {"a": {"inputs": [9]}, "s": "x\\"]\\\\", "inputs" :' + nl + " [ [1 , 2]," + tab + '[3,4] ] , "z": [1,{"q":[2]}]}')
    assert json.loads(body)["inputs"] == [[1, 2], [3, 4]]
    got, _ = codec.parse_inputs(body)
    np.testing.assert_array_equal(got, [[1, 2], [3, 4]])
    got, _ = codec.parse_inputs('{"inputs": []}')
    assert got.shape == (0, 0)
    got, _ = codec.parse_inputs('{"inputs": [[], []]}')
    assert got.shape == (2, 0)
    got, _ = codec.parse_inputs('{"inputs": [5]}')  # BASELINE configs[0] body
    assert got.tolist() == [[5.0]]


@pytest.mark.parametrize("body", ['{"inputs": [[1, 2], [3]]}', '{"inputs": [{"a": 1}]}', '{"inputs": ["x"]}', '{"x": 1}', "[1, 2]",
                                  '{"inputs": [1, [2]]}', '{"inputs": 5}'])
def test_bodies_for_the_per_event_path_are_reported(body):
    json.loads(body)  # valid JSON
    with pytest.raises(codec.NotV2Matrix):
        codec.parse_inputs(body)


@pytest.mark.parametrize("body", ['{"inputs": [[1, 2]', '{"inputs": [1,, 2]}', '{"inputs": [01]}', '{"inputs": [1.]}', '{inputs: [1]}'])
def test_malformed_json_is_an_error(body):
    with pytest.raises(json.JSONDecodeError):
        json.loads(body)
    with pytest.raises(Exception) as ei:
        codec.parse_inputs(body)
    assert not isinstance(ei.value, AssertionError)


def test_format_is_byte_identical_to_json_dumps():
    rng = np.random.default_rng(12)
    vals = np.concatenate([
        (rng.normal(size=3000) * 10.0 ** rng.integers(-12, 13, size=3000)).astype(np.float32),
        np.array([0.0, -0.0, 1.0, -1.0, 100000.0, 1e16, 9999999.0, 1e-4, 9.999e-5, 1e-5, 123456.789, 3.4028235e38, 1e-45, 0.1,
                  16777216.0, 1e15, 1e22], dtype=np.float32)])
    assert codec.format_outputs(vals) == json.dumps([float(v) for v in vals]).encode()
    m = vals[:3000].reshape(600, 5)
    assert codec.format_outputs(m) == json.dumps(m.astype(np.float64).tolist()).encode()
    labels = rng.integers(-5, 1000, size=777).astype(np.int32)
    assert codec.format_outputs(labels) == json.dumps(labels.tolist()).encode()
    specials = np.array([np.nan, np.inf, -np.inf], dtype=np.float32)
    assert codec.format_outputs(specials) == json.dumps([math.nan, math.inf, -math.inf]).encode()
    resp = {"id": "e1", "model_name": "ens", "outputs": None, "model_version": "v1"}
    text = codec.dumps_with_outputs(resp, codec.format_outputs(vals[:4]))
    assert text == json.dumps({**resp, "outputs": [float(v) for v in vals[:4]]}).encode()


def test_codec_property_fuzz():
    """hypothesis: any float64 matrix json.dumps can print parses to the same float32 bits, and any float32 vector
    prints to the same text as json.dumps"""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    floats = st.floats(allow_nan=True, allow_infinity=True, width=64)

    @settings(max_examples=150, deadline=None)
    @given(st.lists(st.lists(floats, min_size=3, max_size=3), min_size=1, max_size=6))
    def parse_roundtrip(rows):
        body = json.dumps({"id": "x", "inputs": rows})
        got, _ = codec.parse_inputs(body)
        with np.errstate(over="ignore"):
            want = np.asarray(json.loads(body)["inputs"], dtype=np.float32)  # what the reference's decode gives
        np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))

    @settings(max_examples=150, deadline=None)
    @given(st.lists(st.floats(width=32, allow_nan=True, allow_infinity=True), min_size=1, max_size=8))
    def print_identical(vals):
        a = np.array(vals, dtype=np.float32)
        assert codec.format_outputs(a) == json.dumps([float(v) for v in a]).encode()

    parse_roundtrip()
    print_identical()


def _big_rows(n_rows=3000, n_cols=40, seed=3):
    rng = np.random.default_rng(seed)
    return (rng.normal(size=(n_rows, n_cols)) * 10.0 ** rng.integers(-6, 7, size=(n_rows, n_cols))).tolist()


def test_large_matrices_take_the_threaded_parser_with_the_same_results():
    """bodies over 256 KB are split at the rows' closing brackets and parsed by several threads (b2s_codec.cpp
    parse_matrix_parallel): same float32 bits and the same extents as json.loads, whatever the layout between rows"""
    rows = _big_rows()
    rows[17][3], rows[2999][39], rows[1500][0] = math.nan, math.inf, -math.inf
    for sep, indent in ((", ", None), (",", None), (",\n   ", None), (", ", 1)):
        inner = sep.join(json.dumps(r) for r in rows) if indent is None else json.dumps(rows, indent=indent)[1:-1]
        body = '{"id": "big", "inputs": [' + inner + '], "model": "m"}'
        assert len(body) > (1 << 18)
        got, (b, e) = codec.parse_inputs(body)
        want = np.asarray(json.loads(body)["inputs"], dtype=np.float32)
        np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
        assert body[b] == "[" and body[e - 1] == "]" and json.loads(body[:b] + "null" + body[e:]) == {"id": "big", "inputs": None, "model": "m"}
    body = '{"inputs": [' + ", ".join(["[]"] * 80000) + "]}"  # many empty rows
    got, _ = codec.parse_inputs(body)
    assert got.shape == (80000, 0)
    body = '{"inputs": [' + ", ".join(["[null, 1]"] * 40000) + "]}"
    got, _ = codec.parse_inputs(body)
    assert got.shape == (40000, 2) and np.isnan(got[:, 0]).all() and (got[:, 1] == 1).all()


@pytest.mark.parametrize("spoil,error", [
    (lambda rows: rows[2000].pop(), codec.NotV2Matrix),                       # ragged row
    (lambda rows: rows[10].__setitem__(5, "x]y"), codec.NotV2Matrix),          # a string holding a bracket
    (lambda rows: rows[2999].__setitem__(0, [1.0]), codec.NotV2Matrix),        # nested list
    (lambda rows: rows[0].__setitem__(1, {"a": [1]}), codec.NotV2Matrix),      # dict in row 0
    (lambda rows: rows.__setitem__(1234, 5.0), codec.NotV2Matrix),             # scalar between rows
])
def test_large_bodies_the_threaded_parser_cannot_take_are_classified_by_the_sequential_one(spoil, error):
    rows = _big_rows()
    spoil(rows)
    body = json.dumps({"inputs": rows})
    json.loads(body)
    with pytest.raises(error):
        codec.parse_inputs(body)


def test_large_malformed_bodies_are_errors():
    text = json.dumps({"inputs": _big_rows()})
    for bad in (text[:-2], text.replace("], [", "] [", 1), text.replace("], [", "],, [", 1), text[: len(text) // 2]):
        with pytest.raises(json.JSONDecodeError):
            json.loads(bad)
        with pytest.raises(Exception) as ei:
            codec.parse_inputs(bad)
        assert not isinstance(ei.value, AssertionError)  # NativeError, or NotV2Matrix -> json.loads raises for the caller


def test_number_grammar_fuzz_against_json_loads():
    """hypothesis: tokens over the number alphabet -- accepted exactly when json.loads accepts them, with the same value"""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=600, deadline=None)
    @given(st.text(alphabet="0123456789.eE+-", min_size=1, max_size=12))
    def check(tok):
        body = '{"inputs": [[' + tok + ", 1]]}"
        try:
            want = json.loads(body)["inputs"][0][0]
        except json.JSONDecodeError:
            want = None
        if want is None:
            with pytest.raises(Exception) as ei:
                codec.parse_inputs(body)
            assert not isinstance(ei.value, AssertionError)
            return
        try:
            with np.errstate(over="ignore"):
                ref = np.asarray([want], dtype=np.float32)
        except OverflowError:  # an int beyond the double range
            with pytest.raises(codec.NotV2Matrix):
                codec.parse_inputs(body)
            return
        got, _ = codec.parse_inputs(body)
        np.testing.assert_array_equal(got[0, :1].view(np.uint32), ref.view(np.uint32))

    check()


def test_binary_bodies_round_trip_and_reject_malformed_ones():
    from mlrun_b200.serving import codec as bcodec

    X = np.random.default_rng(3).normal(size=(37, 5)).astype(np.float32)
    X[3, 2] = np.nan
    body = bcodec.encode_rows(X)
    assert len(body) == 16 + X.nbytes and body[:4] == b"B2S1"
    back = bcodec.decode_rows(body)
    assert back.dtype == np.float32 and np.array_equal(back.view(np.uint32), X.view(np.uint32))
    labels = np.arange(12, dtype=np.int32)
    got = bcodec.decode_rows(bcodec.encode_rows(labels))
    assert got.dtype == np.int32 and got.shape == (12, 1) and np.array_equal(got[:, 0], labels)
    for bad in (b"", b"B2S0" + body[4:], body[:-1], body + b"\0"):
        with pytest.raises(ValueError):
            bcodec.decode_rows(bad)


def test_binary_content_type_through_graph_server_run(monkeypatch):
    """GraphServer.run with content_type application/x-b2s-f32: rows in, result words out, no JSON; a flagged row is the
    event's 400 (the same contract as run_json); runs on the emulated plan (no GPU)"""
    from mlrun_b200 import api
    from mlrun_b200.serving import codec as bcodec
    from mlrun_b200.synthetic import tree_workload
    from oracle import batch as obatch
    from tests import emulated_plan

    emulated_plan.install(monkeypatch)
    wl = tree_workload(n_rows=64, n_feat=8, n_models=3, n_trees=5, depth=3, seed=2, n_fit=300)
    server = wl.build_server(api)
    event = api.MockEvent(body=bcodec.encode_rows(wl.X), path="/v2/models/infer", content_type=bcodec.BINARY_CONTENT_TYPE)
    resp = server.run(event)
    assert resp.status_code == 200 and resp.content_type == bcodec.BINARY_CONTENT_TYPE
    out = bcodec.decode_rows(resp.body)
    np.testing.assert_allclose(out[:, 0], obatch.tree_ensemble(wl)["out"], rtol=1e-5, atol=1e-5)
    bad = wl.X.copy()
    bad[5, 1] = np.inf
    resp = server.run(api.MockEvent(body=bcodec.encode_rows(bad), path="/v2/models/infer", content_type=bcodec.BINARY_CONTENT_TYPE))
    assert resp.status_code == 400 and "infinity" in resp.body
    resp = server.run(api.MockEvent(body=b"garbage", path="/v2/models/infer", content_type=bcodec.BINARY_CONTENT_TYPE))
    assert resp.status_code == 400 and "bad magic" in resp.body


def _parse_tokens(tokens):
    got, _ = codec.parse_inputs('{"inputs": [[' + ", ".join(tokens) + "]]}")
    return got[0]


def _float32_of(tokens):
    with np.errstate(over="ignore"):
        return np.asarray([float(t) for t in tokens], dtype=np.float64).astype(np.float32)


def test_fast_number_path_is_correctly_rounded_on_long_digit_strings():
    """The codec's fast path (b2s_codec.cpp parse_number_fast: <= 19 digits, |exponent| <= 27, one 80-bit multiply or divide,
    hand-over to the exact converter on a double's rounding boundary) gives float32(float(token)) bit for bit: random
    digit strings of every length in every notation"""
    import random

    rnd = random.Random(5)
    tokens = []
    for _ in range(60000):
        nd = rnd.randint(1, 21)
        digits = str(rnd.randint(1, 9)) + "".join(rnd.choice("0123456789") for _ in range(nd - 1))
        form = rnd.randint(0, 4)
        if form == 0:  # d.ddd
            cut = rnd.randint(1, nd)
            tok = digits[:cut] + ("." + digits[cut:] if cut < nd else "")
        elif form == 1:  # 0.000ddd
            tok = "0." + "0" * rnd.randint(0, 12) + digits
        elif form == 2:  # d.ddde+-xx
            tok = digits[0] + ("." + digits[1:] if nd > 1 else "") + rnd.choice("eE") + rnd.choice(["", "+", "-"]) + str(rnd.randint(0, 45))
        elif form == 3:  # integer
            tok = digits
        else:  # ddd.ddde-xx around the fast path's exponent limit
            cut = rnd.randint(1, nd)
            tok = digits[:cut] + "." + (digits[cut:] or "0") + "e" + str(rnd.randint(-30, 30))
        tokens.append(("-" if rnd.random() < 0.5 else "") + tok)
    got = _parse_tokens(tokens)
    want = _float32_of(tokens)
    bad = np.flatnonzero(got.view(np.uint32) != want.view(np.uint32))
    assert bad.size == 0, [(tokens[i], got[i], want[i]) for i in bad[:5]]


def test_numbers_next_to_rounding_boundaries():
    """tokens a hair above / below / on the midpoint of two neighbouring float32 values (where the float32 result depends on
    the double being the correctly rounded one) and of two neighbouring doubles (where the 80-bit product lands on the
    boundary pattern and the exact converter has to take over)"""
    from decimal import Decimal, getcontext

    getcontext().prec = 60
    rng = np.random.default_rng(9)
    tokens = []
    a = np.abs(rng.normal(size=4000).astype(np.float32) * np.float32(10.0) ** rng.integers(-6, 7, size=4000).astype(np.float32))
    a = a[np.isfinite(a) & (a > 0)]
    b = np.nextafter(a, np.float32(np.inf))
    for lo, hi in zip(a.tolist(), b.tolist()):
        mid = (Decimal(lo) + Decimal(hi)) / 2  # exact: a float32 midpoint is a double
        ulp_d = Decimal(float(np.nextafter(np.float64(float(mid)), np.inf))) - mid
        # +-0.5: the midpoint of the float32 midpoint and its neighbouring double; its nearest 19-digit decimal is often closer to
        # it than half a unit of the 80-bit format, so the product lands exactly on the boundary while the token is beside it
        for k in (Decimal("-0.6"), Decimal("-0.5"), Decimal("-0.4"), Decimal(0), Decimal("0.4"), Decimal("0.5"), Decimal("0.6")):
            x = mid + k * ulp_d
            for digits in (17, 18, 19):
                tokens.append(format(x, f".{digits - 1}e"))  # d.ddde-xx
                tokens.append(format(x.quantize(Decimal(1).scaleb(x.adjusted() - digits + 1)), "f"))  # plain notation
    d = np.abs(rng.normal(size=3000) * 10.0 ** rng.integers(-8, 9, size=3000))
    for lo in d.tolist():
        hi = float(np.nextafter(np.float64(lo), np.inf))
        mid = (Decimal(lo) + Decimal(hi)) / 2
        for digits in (17, 18, 19):
            q = Decimal(1).scaleb(mid.adjusted() - digits + 1)
            for x in (mid.quantize(q), mid.quantize(q) + q, mid.quantize(q) - q):
                tokens.append(format(x, "f"))
                tokens.append(format(x, f".{digits - 1}e"))
    got = _parse_tokens(tokens)
    want = _float32_of(tokens)
    bad = np.flatnonzero(got.view(np.uint32) != want.view(np.uint32))
    assert bad.size == 0, [(tokens[i], got[i], want[i]) for i in bad[:5]]


def test_threaded_parser_serves_concurrent_callers_and_survives_fork():
    """the worker pool takes one body at a time (other callers parse on their own thread) and a forked child builds its own"""
    import multiprocessing as mp
    import threading

    rows = _big_rows(n_rows=4000)
    body = json.dumps({"inputs": rows}).encode()
    want = np.asarray(rows, dtype=np.float32)
    results, errors = [None] * 6, []

    def one(i):
        try:
            for _ in range(3):
                results[i] = codec.parse_inputs(body)[0].copy()
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=one, args=(i,)) for i in range(6)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors
    for r in results:
        np.testing.assert_array_equal(r.view(np.uint32), want.view(np.uint32))
    ctx = mp.get_context("fork")  # the pool's threads do not exist in the child
    q = ctx.Queue()
    p = ctx.Process(target=_forked_parse, args=(body, q))
    p.start()
    shape = q.get(timeout=60)
    p.join(timeout=60)
    assert shape == want.shape and p.exitcode == 0


def _forked_parse(body, q):
    q.put(codec.parse_inputs(body)[0].shape)


def test_structural_fuzz_of_whole_bodies_against_json_loads():
    """seeded random V2 bodies -- random white space between every token, specials, exponents, 1 .. 12 000 rows -- and the same
    bodies with one to three characters deleted / inserted / replaced anywhere: `decode_body` (the entry `run_json` uses)
    accepts exactly the bodies json.loads accepts whose "inputs" is a numeric matrix, with the same float32 bits; everything
    else is NotV2Matrix (valid JSON, not a matrix) or an error (invalid JSON), never a silent acceptance"""
    import random

    from mlrun_b200 import _native as nat

    rnd = random.Random(77)
    spaces = [" ", "", "", "\n", "\t", "  ", "\r\n"]

    def num():
        r = rnd.random()
        if r < 0.3:
            return str(rnd.randint(-1000, 1000))
        if r < 0.8:
            return repr(rnd.uniform(-100, 100))
        if r < 0.85:
            return rnd.choice(["NaN", "Infinity", "-Infinity", "null"])
        return "%.3e" % rnd.uniform(-1e10, 1e10)

    def matrix(rows, cols):
        w = lambda: rnd.choice(spaces)  # noqa: E731
        return "[" + w() + ("," + w()).join("[" + w() + ("," + w()).join(num() + w() for _ in range(cols)) + "]" + w() for _ in range(rows)) + "]"

    def numeric(y):
        return (isinstance(y, (int, float)) and not isinstance(y, bool)) or y is None

    def is_matrix(v):
        if not isinstance(v, list):
            return False
        if v and all(isinstance(x, list) for x in v):
            return all(len(x) == len(v[0]) and all(numeric(y) for y in x) for x in v)
        return all(numeric(y) for y in v)

    counts = {"same": 0, "not_matrix": 0, "invalid": 0}
    for it in range(1500):
        rows, cols = rnd.choice([(1, 1), (2, 3), (5, 4), (40, 7), (7, 2)]) if it % 150 else (12000, 12)  # (the last: the pool's path)
        body = '{"id": "x", "inputs":' + rnd.choice(spaces) + matrix(rows, cols) + ', "z": [1, {"a": "]"}]}'
        if rnd.random() < 0.6:
            chars = list(body)
            for _ in range(rnd.randint(1, 3)):
                i = rnd.randrange(len(chars))
                op = rnd.random()
                if op < 0.4:
                    del chars[i]
                elif op < 0.8:
                    chars.insert(i, rnd.choice('[],"{}:.e-0 a'))
                else:
                    chars[i] = rnd.choice('[],"{}:.e-0 a')
            body = "".join(chars)
        try:
            ref = json.loads(body)
        except (json.JSONDecodeError, RecursionError):
            ref = Ellipsis
        try:
            got, _rest = codec.decode_body(body.encode())
            verdict = "ok"
        except codec.NotV2Matrix:
            verdict = "not_matrix"
        except (nat.NativeError, json.JSONDecodeError):
            verdict = "invalid"
        if ref is Ellipsis:
            assert verdict != "ok", body[:300]
            counts["invalid"] += 1
            continue
        if body.count('"inputs"') != 1:
            continue  # a corruption made a second member of that name: json.loads keeps the last, the codec reads the first
        if isinstance(ref, dict) and "inputs" in ref and is_matrix(ref["inputs"]):
            try:
                with np.errstate(over="ignore"):
                    want = np.asarray(ref["inputs"], dtype=np.float32)
            except (OverflowError, ValueError):
                continue
            assert verdict == "ok", (verdict, body[:300])
            want = want.reshape(-1, 1) if want.ndim == 1 else want
            if want.size or got.size:
                assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32)), body[:300]
            counts["same"] += 1
        else:
            assert verdict != "ok", body[:300]
            counts["not_matrix"] += 1
    assert counts["same"] > 500 and counts["invalid"] > 300 and counts["not_matrix"] > 5, counts
