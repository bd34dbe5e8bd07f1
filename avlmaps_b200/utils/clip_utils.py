"""Query-side helpers with the reference's names (reference avlmaps/utils/clip_utils.py).

`get_lseg_score` keeps the reference signature and return contract ((N, C[+1]) float32 scores) but
the contraction runs on the B200 through the engine; the text encoder (CLIP) is the caller's model
object exactly as in the reference -- or pre-computed embeddings can be passed instead of names."""
from __future__ import annotations

from typing import List, Sequence, Union

import numpy as np


def _make_templates() -> List[str]:
    """The 63 prompt templates of the reference (clip_utils.py:10-74), generated from their pattern."""
    t = ["There is {} in the scene.", "There is the {} in the scene.", "a photo of {} in the scene.",
         "a photo of the {} in the scene.", "a photo of one {} in the scene."]
    t += ["I took a picture of of {}.", "I took a picture of of my {}.", "I took a picture of of the {}."]
    t += ["a photo of {}.", "a photo of my {}.", "a photo of the {}.", "a photo of one {}.", "a photo of many {}."]
    for adj in ("good", "bad"):
        t += [f"a {adj} photo of {{}}.", f"a {adj} photo of the {{}}."]
    for adj in ("nice", "cool", "weird", "small", "large", "clean", "dirty"):
        t += [f"a photo of a {adj} {{}}.", f"a photo of the {adj} {{}}."]
    for adj in ("bright", "dark"):
        t += [f"a {adj} photo of {{}}.", f"a {adj} photo of the {{}}."]
    t += ["a photo of a hard to see {}.", "a photo of the hard to see {}."]
    for adj in ("low resolution", "cropped", "close-up", "jpeg corrupted", "blurry", "pixelated"):
        t += [f"a {adj} photo of {{}}.", f"a {adj} photo of the {{}}."]
    t += ["a black and white photo of the {}.", "a black and white photo of {}."]
    for adj in ("plastic", "toy", "plushie", "cartoon"):
        t += [f"a {adj} {{}}.", f"the {adj} {{}}."]
    t += ["an embroidered {}.", "the embroidered {}.", "a painting of the {}.", "a painting of a {}."]
    return t


multiple_templates = _make_templates()


def get_text_feats(in_text: List[str], clip_model, clip_feat_dim: int, batch_size: int = 64) -> np.ndarray:
    """Reference clip_utils.py:133-149: tokenize -> encode_text -> L2-normalise rows -> numpy float32.
    `clip_model` may also be any callable `texts -> (len(texts), D)` array (tests, custom encoders)."""
    if callable(clip_model) and not hasattr(clip_model, "encode_text"):
        feats = np.asarray(clip_model(in_text), dtype=np.float32)
        return feats / np.linalg.norm(feats, axis=-1, keepdims=True)
    import clip  # the reference's dependency (openai/CLIP); absent in the build container
    import torch

    dev = next(clip_model.parameters()).device
    text_tokens = clip.tokenize(in_text).to(dev)
    text_feats = np.zeros((len(in_text), clip_feat_dim), dtype=np.float32)
    text_id = 0
    while text_id < len(text_tokens):
        bs = min(len(in_text) - text_id, batch_size)
        with torch.no_grad():
            batch_feats = clip_model.encode_text(text_tokens[text_id:text_id + bs]).float()
        batch_feats /= batch_feats.norm(dim=-1, keepdim=True)
        text_feats[text_id:text_id + bs, :] = np.float32(batch_feats.cpu())
        text_id += bs
    return text_feats


def get_text_feats_multiple_templates(in_text: List[str], clip_model, clip_feat_dim: int, batch_size: int = 64) -> np.ndarray:
    """Reference clip_utils.py:152-159: mean over the 63 templates, no re-normalisation."""
    mul_tmp = multiple_templates.copy()
    prompts = [x.format(lm) for lm in in_text for x in mul_tmp]
    text_feats = get_text_feats(prompts, clip_model, clip_feat_dim)
    text_feats = text_feats.reshape((-1, len(mul_tmp), text_feats.shape[-1]))
    return np.mean(text_feats, axis=1)


def landmark_text_feats(clip_model, landmarks: Sequence[str], clip_feat_dim: int, use_multiple_templates: bool,
                        avg_mode: int, add_other: bool):
    """The query matrix get_lseg_score multiplies with (clip_utils.py:213-225, 236).  Returns
    (text_feats (Q', D), landmarks_other, n_templates) where Q' = C*63 for avg_mode 1."""
    landmarks_other = list(landmarks)
    if add_other and landmarks_other[-1] != "other":
        landmarks_other = landmarks_other + ["other"]
    if use_multiple_templates:
        mul_tmp = multiple_templates.copy()
        prompts = [x.format(lm) for lm in landmarks_other for x in mul_tmp]
        text_feats = get_text_feats(prompts, clip_model, clip_feat_dim)
        if avg_mode == 0:
            text_feats = np.mean(text_feats.reshape((-1, len(mul_tmp), text_feats.shape[-1])), axis=1)
        return np.ascontiguousarray(text_feats, np.float32), landmarks_other, len(mul_tmp)
    return np.ascontiguousarray(get_text_feats(landmarks_other, clip_model, clip_feat_dim), np.float32), landmarks_other, 1


def get_lseg_score(clip_model, landmarks: list, lseg_map: Union[np.ndarray, "object"], clip_feat_dim: int,
                   use_multiple_templates: bool = False, avg_mode: int = 0, add_other: bool = True) -> np.ndarray:
    """Reference clip_utils.py:196-242, same arguments and (N, C[+1]) float32 result.  `lseg_map` may be
    a numpy array ((N, D) or (h, w, D)) -- uploaded for this call -- or an engine.DeviceMap that is
    already resident in HBM (what VLMap passes)."""
    from ..engine import DeviceMap

    text_feats, landmarks_other, n_tmp = landmark_text_feats(clip_model, landmarks, clip_feat_dim,
                                                             use_multiple_templates, avg_mode, add_other)
    own = not isinstance(lseg_map, DeviceMap)
    dmap = DeviceMap(np.asarray(lseg_map).reshape((-1, np.asarray(lseg_map).shape[-1]))) if own else lseg_map
    try:
        scores_list = dmap.scores(text_feats)
    finally:
        if own:
            dmap.close()
    if use_multiple_templates and avg_mode == 1:
        scores_list = np.mean(scores_list.reshape((-1, len(landmarks_other), n_tmp)), axis=2)
    return scores_list
