"""Run every differential check against the REAL reference (needs /root/reference) and print their verdict lines.

    python -m tests.golden.run_diffs
"""
import glob
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def main():
    failed = []
    for path in sorted(glob.glob(os.path.join(HERE, "diff_*.py"))):
        name = os.path.splitext(os.path.basename(path))[0]
        t0 = time.time()
        done = subprocess.run([sys.executable, "-m", f"tests.golden.{name}"], cwd=ROOT, capture_output=True, text=True)
        verdict = [ln for ln in done.stdout.splitlines() if ln.startswith(("identical", "DIFF", "BUILD DIFF", "the batched", "ingest_columns", "vote math", "get_in", "FeatureSet.ingest"))]
        print(f"{name:28s} rc={done.returncode} {time.time() - t0:5.1f}s  {' | '.join(verdict)[:200]}")
        if done.returncode != 0:
            failed.append(name)
    print("failed:", failed)
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
