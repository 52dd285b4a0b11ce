"""Body codec on this box's host cores: MB/s and ns per number of b2s_json_parse_inputs for 1 / 2 / 4 / 8 pool threads
(B2S_CODEC_THREADS is read once per process, so every setting runs in its own process), scalar vs AVX2 rows, and the
output printer.  Pure host code: no GPU is touched.   python profiles/lab/codec_probe.py [rows] [cols]"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def child(rows, cols):
    import ctypes as C

    from mlrun_b200 import _native as nat

    lib = nat.load()
    rng = np.random.default_rng(0)
    X = rng.normal(size=(rows, cols)).astype(np.float32)
    body = json.dumps({"inputs": X.astype(np.float64).tolist()}).encode()
    out = np.empty(rows * cols + 16, np.float32)
    r, c, b, e = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()

    def run():
        return lib.b2s_json_parse_inputs(body, len(body), out.ctypes.data_as(C.POINTER(C.c_float)), out.size, C.byref(r), C.byref(c),
                                         C.byref(b), C.byref(e))

    for _ in range(3):
        assert run() == 0
    best = 1e9
    for _ in range(15):
        t0 = time.perf_counter()
        run()
        best = min(best, time.perf_counter() - t0)
    assert np.array_equal(out[: rows * cols].reshape(rows, cols), X)
    print(json.dumps({"threads": os.environ.get("B2S_CODEC_THREADS"), "scalar": bool(os.environ.get("B2S_CODEC_SCALAR")),
                      "rows": rows, "cols": cols, "body_mb": round(len(body) / 1e6, 2), "best_ms": round(best * 1e3, 3),
                      "mb_per_s": round(len(body) / best / 1e6, 1), "ns_per_number": round(best * 1e9 / (rows * cols), 2),
                      "events_per_s": round(rows / best)}))


if __name__ == "__main__":
    if os.environ.get("B2S_CODEC_PROBE_CHILD"):
        child(int(sys.argv[1]), int(sys.argv[2]))
        sys.exit(0)
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    cols = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    print("host cpus:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)))
    for threads, scalar in ((1, True), (1, False), (2, False), (4, False), (8, False), (16, False)):
        env = dict(os.environ, B2S_CODEC_PROBE_CHILD="1", B2S_CODEC_THREADS=str(threads))
        if scalar:
            env["B2S_CODEC_SCALAR"] = "1"
        subprocess.run([sys.executable, os.path.abspath(__file__), str(rows), str(cols)], env=env, check=False)
