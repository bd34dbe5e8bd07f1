"""Drop-in surface for the hot-path classes of `avlmaps.map` (reference avlmaps/map/__init__.py:7-13).
AreaMap / SoundMap keep the reference's method surface with the encoders (CLIP L/14, AudioCLIP) injected as
callables; their similarity call sites run through engine.DeviceMap.  VisualMap (HLoc localisation) is out of
scope."""
from .map import Map
from .vlmap import VLMap
from .vlmap_builder import VLMapBuilder
from .vlmap_builder_multi_floor import VLMapBuilderMultiFloor
from .vlmap_multi_floor import VLMapMultiFloor
from .area_map import AreaMap
from .sound_map import SoundMap
from .avlmap import AVLMap

__all__ = ["Map", "VLMap", "VLMapBuilder", "VLMapMultiFloor", "VLMapBuilderMultiFloor", "AreaMap", "SoundMap", "AVLMap"]
