"""Differential check of the remaining feature steps against the REAL reference classes (build container only), per event:
DateExtractor (every part pandas offers that the reference lists, timestamps as strings in several formats, epoch ints,
datetime objects, with and without time zones, leap days, year ends, bad input), FeaturesetValidator / MinMaxValidator
(random min / max / severity rules against ints, floats, nan, None, strings, huge ints; column filters), SetEventMetadata
(id / key paths present and missing).  Results, printed reports and exceptions compared.

    python -m tests.golden.diff_steps_misc
"""
import contextlib
import datetime
import io
import os
import random
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import api_b200 as mine  # noqa: E402
from tests.golden import api_reference as ref  # noqa: E402
from tests.scenarios import _clean, _first_line  # noqa: E402

PARTS = ["asm8", "day_of_week", "day_of_year", "dayofweek", "dayofyear", "days_in_month", "daysinmonth", "freqstr", "is_leap_year",
         "is_month_end", "is_month_start", "is_quarter_end", "is_quarter_start", "is_year_end", "is_year_start", "quarter", "tz", "week",
         "weekofyear", "year", "month", "day", "hour", "minute", "second", "microsecond", "nanosecond", "fortnight"]


def captured(fn):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        try:
            res = ("ok", fn())
        except Exception as exc:  # noqa: BLE001
            res = ("exc", f"{type(exc).__name__}: {_first_line(exc)}")
    return res, buf.getvalue().splitlines()


def stamp(rnd):
    y, mo, d = rnd.choice([1969, 1970, 1999, 2000, 2020, 2021, 2024, 2038, 2100]), rnd.randint(1, 12), rnd.randint(1, 28)
    if rnd.random() < 0.1:
        mo, d = 2, 29 if y % 4 == 0 and (y % 100 != 0 or y % 400 == 0) else 28
    if rnd.random() < 0.1:
        mo, d = 12, 31
    h, mi, s = rnd.randint(0, 23), rnd.randint(0, 59), rnd.randint(0, 59)
    dt = datetime.datetime(y, mo, d, h, mi, s, rnd.choice([0, 0, 123456]))
    form = rnd.randint(0, 7)
    if form == 0:
        return dt.isoformat()
    if form == 1:
        return dt.strftime("%Y-%m-%d %H:%M:%S")
    if form == 2:
        return dt.strftime("%Y-%m-%d")
    if form == 3:
        return dt
    if form == 4:
        return dt.isoformat() + rnd.choice(["+00:00", "+02:00", "-05:30", "Z"])
    if form == 5:
        return int(dt.replace(tzinfo=datetime.timezone.utc).timestamp())
    if form == 6:
        return rnd.choice(["not a date", "", None, float("nan"), "2021-13-45", 1e30])
    return dt.strftime("%d/%m/%Y %H:%M")


def main():
    rnd = random.Random(8)
    n = 0
    for _ in range(1500):
        parts = rnd.sample(PARTS, rnd.randint(1, 4))
        col = rnd.choice([None, "timestamp", "when"])
        body = {rnd.choice(["timestamp", "when"]): stamp(rnd), "v": 1}
        out = []
        for api in (ref, mine):
            kw = {"parts": list(parts)}
            if col:
                kw["timestamp_col"] = col
            res, _lines = captured(lambda: _clean(api.DateExtractor(**kw).do(dict(body))))
            out.append(repr(res))
        n += 1
        if out[0] != out[1]:
            print("DIFF DateExtractor", parts, col, body)
            print("  ref :", out[0][:500])
            print("  mine:", out[1][:500])
            return 1
    values = [5, 0, -3, 2.5, float("nan"), None, "text", "7", 10**40, True, float("inf"), [1], 30, 30.0]
    for _ in range(1200):
        rules = {}
        for c in rnd.sample(["a", "b", "c", "d"], rnd.randint(1, 3)):
            kw = {}
            if rnd.random() < 0.8:
                kw["min"] = rnd.choice([0, 1, 30, -1.5])
            if rnd.random() < 0.8:
                kw["max"] = rnd.choice([5, 30, 100, 2.5])
            if rnd.random() < 0.6:
                kw["severity"] = rnd.choice(["info", "warning", "error"])
            rules[c] = kw
        columns = rnd.choice([None, None, ["a"], ["a", "c"], ["zz"]])
        body = {c: rnd.choice(values) for c in rnd.sample(["a", "b", "c", "d", "e"], rnd.randint(1, 5))}
        key = rnd.choice([None, "k1", 17])
        out = []
        for api in (ref, mine):
            def go(api=api):
                step = api.validator_step(rules, columns)
                ev = types.SimpleNamespace(body=dict(body), key=key)
                got = step.do(ev)
                return (got is ev, _clean(ev.body))
            res, lines = captured(go)
            out.append(repr((res, lines)))
        n += 1
        if out[0] != out[1]:
            print("DIFF validator", rules, columns, body, key)
            print("  ref :", out[0][:700])
            print("  mine:", out[1][:700])
            return 1
    for _ in range(300):
        id_path = rnd.choice([None, "id", "meta.id", "nope", "meta.nope"])
        key_path = rnd.choice([None, "k", "meta.k", "nope"])
        body = rnd.choice([{"id": "i1", "k": "key1", "meta": {"id": 7, "k": [1]}}, {"meta": {}}, {"id": None, "k": 0}, 5, None])
        out = []
        for api in (ref, mine):
            def go(api=api):
                step = api.SetEventMetadata(id_path=id_path, key_path=key_path)
                ev = types.SimpleNamespace(body=body if not isinstance(body, dict) else {**body}, id="orig", key="origk")
                step.post_init() if hasattr(step, "post_init") else None
                got = step.do(ev)
                return (got is ev, ev.id, ev.key)
            res, _lines = captured(go)
            out.append(repr(res))
        n += 1
        if out[0] != out[1]:
            print("DIFF SetEventMetadata", id_path, key_path, body)
            print("  ref :", out[0][:500])
            print("  mine:", out[1][:500])
            return 1
    print("identical on", n, "cases")
    return 0


if __name__ == "__main__":
    sys.exit(main())
