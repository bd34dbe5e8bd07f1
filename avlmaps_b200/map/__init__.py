"""Drop-in surface for the hot-path classes of `avlmaps.map` (reference avlmaps/map/__init__.py:7-13).
AreaMap / SoundMap / VisualMap wrap third-party encoders (CLIP L/14, AudioCLIP, HLoc) and are not
part of the accelerated path; their similarity call sites map to engine.DeviceMap.scores/topk."""
from .map import Map
from .vlmap import VLMap
from .vlmap_builder import VLMapBuilder
from .vlmap_builder_multi_floor import VLMapBuilderMultiFloor
from .vlmap_multi_floor import VLMapMultiFloor
from .avlmap import AVLMap

__all__ = ["Map", "VLMap", "VLMapBuilder", "VLMapMultiFloor", "VLMapBuilderMultiFloor", "AVLMap"]
