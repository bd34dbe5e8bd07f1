"""avlmaps_b200: the two data-parallel hot paths of AVLMaps (map build, landmark index) as
hand-written sm_100a kernels behind the reference's `avlmaps.map` surface.

    from avlmaps_b200.map import VLMap, VLMapBuilder        # drop-in classes
    from avlmaps_b200.engine import DeviceMap, DeviceBuilder # the C-ABI handles

Importing this package does not need a GPU; any compute call without one raises (no CPU fallback)."""
__version__ = "0.1.0"
