"""`SoundMap` (reference avlmaps/map/sound_map.py:19-153): a database of AudioCLIP segment embeddings
(`audio_database[id] = {"audio_features": (1024,), "locations": [...]}`) queried with text or audio.
AudioCLIP itself (and ffmpeg / librosa) stays outside this engine:
    text_encoder(list of category names) -> (C, 1024) text features as `aclp(text=...)` returns them,
    audio_encoder(path, sample_rate)     -> (1024,) features of a query clip,
    logit_scale_at                       -> the model's log logit scale (audioclip.py:94-95).
The similarity is the scaled form `clamp(exp(logit_scale_at), 1, 100) * A @ T.T` (:108-109, 141-142), computed
by engine.DeviceMap.scores(scale=...), then the reference's argmax retrieval (:113, :131) and min-max (:151-152)."""
from __future__ import annotations

import os
import pickle
from pathlib import Path
from typing import Callable, List, Optional, Tuple

import numpy as np

from ..engine import DeviceMap
from .vlmap import find_similar_category_id


class SoundMap:
    def __init__(self, sound_categories: List[str], text_encoder: Callable, logit_scale_at: float,
                 audio_encoder: Optional[Callable] = None, audio_database: Optional[dict] = None,
                 difficulty_level: int = 1, is_ambiguous: bool = False, is_real: bool = False):
        self.sound_categories = list(sound_categories)
        self.text_encoder = text_encoder
        self.audio_encoder = audio_encoder
        self.logit_scale_at = float(logit_scale_at)
        self.difficulty_level = difficulty_level
        self.manual_str = "_manual" if is_ambiguous else ""
        self.is_real = is_real
        self._device_map: Optional[DeviceMap] = None
        self.audio_database = None
        if audio_database is not None:
            self.set_audio_database(audio_database)

    def set_audio_database(self, audio_database: dict) -> None:
        self.audio_database = audio_database
        if self._device_map is not None:
            self._device_map.close()
        self._device_map = DeviceMap(self.get_all_audio_features_and_locations()[0])

    def load_sound_map(self, data_dir: str):
        """Reference sound_map.py:72-84 (same pickle layout and file names)."""
        filename = "audio_data.pkl" if self.is_real else f"audio_data{self.manual_str}_{self.difficulty_level}.pkl"
        with open(Path(data_dir) / "audio_video" / filename, "rb") as f:
            self.set_audio_database(pickle.load(f))
        return self.audio_database

    def get_all_audio_features_and_locations(self) -> Tuple[np.ndarray, List[List[np.ndarray]]]:
        """Reference sound_map.py:86-97."""
        audio_features, feature_locations = [], []
        for id in range(len(self.audio_database.keys())):
            audio_features.append(np.asarray(self.audio_database[id]["audio_features"], np.float32).reshape(-1))
            feature_locations.append(self.audio_database[id]["locations"])
        return np.stack(audio_features, axis=0), feature_locations

    @property
    def scale_audio_text(self) -> float:
        """torch.clamp(logit_scale_at.exp(), min=1.0, max=100.0) (:108), evaluated in float32 like torch."""
        return float(np.clip(np.exp(np.float32(self.logit_scale_at)), np.float32(1.0), np.float32(100.0)))

    def _logits_audio_text(self) -> np.ndarray:
        text_features = np.ascontiguousarray(self.text_encoder(self.sound_categories), np.float32)
        scale = np.full((text_features.shape[0],), self.scale_audio_text, np.float32)
        return self._device_map.scores(text_features, scale=scale)   # (M, C)

    def get_pos(self, name: str):
        """Reference sound_map.py:102-120: the segment retrieved for a category = argmax over segments."""
        _, feature_locations = self.get_all_audio_features_and_locations()
        retrievals = np.argmax(self._logits_audio_text(), axis=0)
        cat_id = find_similar_category_id(name, self.sound_categories)
        return feature_locations[retrievals[cat_id]]

    def get_pos_with_audio(self, audio_path: str, sample_rate: int):
        """Reference sound_map.py:122-133: `audio_features @ query.T`, argmax."""
        if not os.path.exists(audio_path):
            return [], []
        if self.audio_encoder is None:
            raise RuntimeError("SoundMap.get_pos_with_audio needs audio_encoder=... (AudioCLIP is outside this engine)")
        query = np.asarray(self.audio_encoder(audio_path, sample_rate), np.float32).reshape((1, -1))
        _, feature_locations = self.get_all_audio_features_and_locations()
        idx, _ = self._device_map.topk(query, 1)
        return feature_locations[int(idx[0, 0])]

    def get_distribution_and_locations(self, name: str) -> Tuple[np.ndarray, List[np.ndarray]]:
        """Reference sound_map.py:135-153."""
        _, feature_locations = self.get_all_audio_features_and_locations()
        cat_id = find_similar_category_id(name, self.sound_categories)
        probabilities = self._logits_audio_text()[:, cat_id]
        probabilities = (probabilities - np.min(probabilities)) / (np.max(probabilities) - np.min(probabilities))
        return probabilities, feature_locations
