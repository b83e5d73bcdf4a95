"""rasr_b200 -- Blackwell-native acoustic front-end and emission-score engine behind RASR's
Flow::Node and Mm::FeatureScorer interfaces.  The compute lives in csrc/ (CUDA, sm_100a) behind the
C-ABI declared in include/rasr_b200.h; this package is the thin host-side mirror used by the tests
and the bench."""
__version__ = "0.1.0"
