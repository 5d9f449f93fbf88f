"""xmap_b200: B200-native AlterEgo-construction hot path of X-MAP.

Layout (only what the path needs):
  csrc/            CUDA kernels + the C ABI (include/xmap_b200.h) -> libxmap_b200.so
  _native.py       ctypes binding of that ABI (no CPU fallback)
  engine.py        array-level driver: layout, similarity + top-k, X-SIM, generation
  encode.py        id strings <-> dense indices, the reference's per-item string tests
  core/, utils/    host-side mirror of the reference interface (same class and
                   pipeline names as xmap.core.* / xmap.utils.assist)
  rdd.py           minimal local RDD used when no SparkContext is supplied
  synth.py         synthetic Zipf rating data of the BASELINE.json shapes
"""
__version__ = "0.1.0"
