"""Generate tests/golden/scenarios.json by running tests/scenarios.py through the REAL reference.

Run here (the build container), from the repo root:   python -m tests.golden.gen_golden
Requires /root/reference (read-only).  The reference is imported in place with its missing
third-party dependencies mocked (tests/golden/_refshim.py); `storey` (the async engine) is one of
them, so scenarios flagged ASYNC are not generated -- they stay pinned by the literal expectations
copied from the reference's tests.  Nothing here is imported by the test-suite or the product.
"""

import json
import math
import os
import sys
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))


def _jsonable(o):
    if isinstance(o, float) and (math.isnan(o) or math.isinf(o)):
        return repr(o)
    return o


def main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from tests import scenarios
    from tests.golden import api_reference as api

    out, failed = {}, []
    for fn in scenarios.SCENARIOS:
        if getattr(fn, "ASYNC", False):
            continue
        try:
            res = fn(api)
            skips = [tuple(s) for s in getattr(fn, "GOLDEN_SKIP", [])]
            for path, want in getattr(fn, "EXPECT", {}).items():
                if any(path[: len(s)] == s for s in skips):
                    continue  # environment-dependent key (documented at the scenario)
                got = scenarios.dig(res, path)
                assert got == want, f"{fn.__name__}{path}: reference gave {got!r}, its own test expects {want!r}"
            out[fn.__name__] = res
            print(f"  ok   {fn.__name__}")
        except Exception:  # noqa: BLE001
            failed.append(fn.__name__)
            print(f"  FAIL {fn.__name__}\n{traceback.format_exc()}")
    path = os.path.join(HERE, "scenarios.json")
    with open(path, "w") as fp:
        json.dump(out, fp, indent=1, sort_keys=True, default=_jsonable)
    print(f"wrote {path}: {len(out)} scenarios, failed: {failed}")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
