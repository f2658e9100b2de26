"""GPU drop-ins for `scenedetect.detectors` (same names, constructors and metric keys)."""

from .adaptive_detector import AdaptiveDetector
from .content_detector import ContentDetector
from .hash_detector import HashDetector
from .histogram_detector import HistogramDetector
from .threshold_detector import ThresholdDetector

__all__ = ["AdaptiveDetector", "ContentDetector", "HashDetector", "HistogramDetector", "ThresholdDetector"]
