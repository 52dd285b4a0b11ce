"""Structural comparison helpers: exact for ints/strings/structure, tolerance for floats."""

import math


def assert_same(got, want, path="", rtol=0.0, atol=0.0):
    if isinstance(want, dict):
        assert isinstance(got, dict), f"{path}: expected dict, got {type(got).__name__}"
        assert list(got.keys()) == list(want.keys()) or set(got.keys()) == set(want.keys()), (
            f"{path}: keys differ {sorted(map(str, got.keys()))} vs {sorted(map(str, want.keys()))}"
        )
        for k in want:
            assert_same(got[k], want[k], f"{path}/{k}", rtol, atol)
    elif isinstance(want, (list, tuple)):
        assert isinstance(got, (list, tuple)), f"{path}: expected list, got {type(got).__name__}"
        assert len(got) == len(want), f"{path}: length {len(got)} != {len(want)}"
        for i, (g, w) in enumerate(zip(got, want)):
            assert_same(g, w, f"{path}[{i}]", rtol, atol)
    elif isinstance(want, bool) or want is None or isinstance(want, str):
        assert got == want, f"{path}: {got!r} != {want!r}"
    elif isinstance(want, int):
        # ints are exact (and must not silently become floats with a different value)
        assert got == want, f"{path}: {got!r} != {want!r}"
    elif isinstance(want, float):
        assert isinstance(got, (int, float)), f"{path}: expected number, got {got!r}"
        if math.isnan(want):
            assert math.isnan(got), f"{path}: expected NaN, got {got!r}"
        elif rtol == 0.0 and atol == 0.0:
            assert got == want, f"{path}: {got!r} != {want!r}"
        else:
            assert abs(got - want) <= atol + rtol * abs(want), f"{path}: {got!r} !~ {want!r} (rtol={rtol}, atol={atol})"
    else:
        assert got == want, f"{path}: {got!r} != {want!r}"
