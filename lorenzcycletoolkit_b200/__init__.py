"""B200-native Lorenz Energy Cycle engine: a drop-in for the hot path of
daniloceano/LorenzCycleToolkit (``src/analysis`` term classes and the
``src/frameworks`` fixed/moving drivers) backed by hand-written sm_100a CUDA
kernels behind the C ABI of ``include/lec_b200.h``.  No CPU fallback."""

__version__ = "0.1.0"
