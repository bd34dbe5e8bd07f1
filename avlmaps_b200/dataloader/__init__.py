"""Pose converter between the simulator / robot frame and the map grid (reference avlmaps/dataloader)."""
from .habitat_dataloader import VLMapsDataloaderHabitat

__all__ = ["VLMapsDataloaderHabitat"]
