"""mlrun_b200 -- a B200-native serving-graph engine behind the mlrun.serving plugin API.

Importing the package never touches CUDA; device work goes through `mlrun_b200._native`, which
raises loudly when `libb200serve.so` is missing (there is no CPU fallback on the hot path).
"""

__version__ = "0.1.0"
