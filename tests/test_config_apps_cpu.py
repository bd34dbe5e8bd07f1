"""CPU: the Hydra-free config composer (avlmaps_b200/config.py) and the two application drop-ins
(avlmaps_b200/application), i.e. the callers of the hot path (reference application/create_map.py:7-17,
application/index_map.py:18-149, config/*.yaml).  No GPU: the engine calls are stubbed, what is tested is the
config semantics and the control flow around them."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

from avlmaps_b200.config import Config, ConfigError, compose

REF_CONFIG = Path("/root/reference/config")


def write_tree(root: Path) -> Path:
    """A config tree with the same shapes as the reference's (defaults list, groups, interpolation)."""
    (root / "data_paths").mkdir(parents=True)
    (root / "map_config").mkdir()
    (root / "params").mkdir()
    (root / "main.yaml").write_text(
        "defaults:\n  - data_paths: default\n  - map_config: vlmaps\n  - params: default\n  - _self_\n"
        "nav:\n  valid_range: 1\n  vis: False\nscene_id: 0\ndecay_rate: 0.01\n")
    (root / "self_first.yaml").write_text("defaults:\n  - _self_\n  - params: default\nparams:\n  gs: 7\n  own: 1\n")
    (root / "no_self.yaml").write_text("defaults:\n  - params: default\nparams:\n  gs: 7\n")
    (root / "data_paths" / "default.yaml").write_text('avlmaps_data_dir: "/data/a"\n')
    (root / "data_paths" / "lab.yaml").write_text('avlmaps_data_dir: "/data/lab"\n')
    (root / "map_config" / "vlmaps.yaml").write_text(
        "map_type: vlmap\npose_info:\n  pose_type: mobile_base\n  camera_height: ${params.camera_height}\n"
        "  base2cam_rot: [1, 0, 0, 0, -1, 0, 0, 0, -1]\n  base_forward_axis: [0, 0, -1]\n  base_left_axis: [-1, 0, 0]\n"
        "  base_up_axis: [0, 1, 0]\ncam_calib_mat: [540, 0, 540, 0, 540, 360, 0, 0, 1]\ngrid_size: ${params.gs}\ncell_size: ${params.cs}\n"
        "depth_sample_rate: 100\nlabel: \"grid ${params.gs} at ${params.cs} m\"\nnames:\n  - chair\n  - \"${params.extra}\"\n")
    (root / "params" / "default.yaml").write_text("gs: 1000\ncs: 0.05\ncamera_height: 1.5\nextra: wall\nchain: ${params.gs}\n")
    return root


def test_compose_defaults_interpolation_and_access(tmp_path):
    c = compose(write_tree(tmp_path), "main.yaml")
    assert isinstance(c, Config)
    assert c.map_config.grid_size == 1000 and isinstance(c.map_config.grid_size, int)      # typed, not a string
    assert c["map_config"]["cell_size"] == 0.05 and c.map_config.pose_info.camera_height == 1.5
    assert c.map_config.label == "grid 1000 at 0.05 m"                                      # embedded -> formatted
    assert c.map_config.names == ["chair", "wall"] and c.params.chain == 1000
    assert c.map_config.pose_info.base2cam_rot == [1, 0, 0, 0, -1, 0, 0, 0, -1]
    assert c.scene_id == 0 and c.nav.vis is False and c.data_paths.avlmaps_data_dir == "/data/a"
    assert "map_config" in c and "nope" not in c and c.get("nope", 3) == 3 and c.get("decay_rate") == 0.01
    with pytest.raises(AttributeError):
        c.map_config.nope
    with pytest.raises(KeyError):
        c["nope"]
    d = c.to_dict()
    assert d["map_config"]["grid_size"] == 1000 and isinstance(d["map_config"], dict)
    # the map classes read it both ways (map.py:23-24,60-66)
    from avlmaps_b200.map.map import Map, cfg_get

    assert cfg_get(c.map_config, "grid_size") == 1000
    m = Map(c.map_config)
    assert m.gs == 1000 and m.cs == 0.05 and m.base2cam_tf[1, 3] == 1.5 and m.base2cam_tf[1, 1] == -1


def test_self_position_in_the_defaults_list(tmp_path):
    root = write_tree(tmp_path)
    assert compose(root, "self_first.yaml").params.gs == 1000      # the group file merges AFTER _self_ and wins
    assert compose(root, "self_first.yaml").params.own == 1
    assert compose(root, "no_self.yaml").params.gs == 7            # no _self_: the primary file merges last


def test_overrides(tmp_path):
    root = write_tree(tmp_path)
    c = compose(root, "main", ["scene_id=3", "map_config.depth_sample_rate=1", "params.gs=256", "+extra.k=[1, 2]",
                               "data_paths=lab", "nav.vis=true", "~decay_rate", "data_paths.avlmaps_data_dir=/x y"])
    assert c.scene_id == 3 and c.map_config.depth_sample_rate == 1 and c.map_config.grid_size == 256
    assert c.extra.k == [1, 2] and c.nav.vis is True and "decay_rate" not in c
    assert c.data_paths.avlmaps_data_dir == "/x y"
    assert compose(root, "main", ["data_paths=lab"]).data_paths.avlmaps_data_dir == "/data/lab"
    for bad, msg in ((["nope.key=1"], "not in the config"), (["scene_id"], "key=value"), (["~nope"], "not in the config"),
                     (["data_paths=missing"], "does not exist"), (["scene_id=[1"], "not valid YAML")):
        with pytest.raises(ConfigError, match=msg):
            compose(root, "main", bad)


def test_errors(tmp_path):
    root = write_tree(tmp_path)
    with pytest.raises(ConfigError, match="does not exist"):
        compose(root, "absent.yaml")
    (root / "params" / "default.yaml").write_text("gs: ${params.cs}\ncs: ${params.gs}\ncamera_height: 1\nextra: w\nchain: 1\n")
    with pytest.raises(ConfigError, match="cycle"):
        compose(root, "main")
    (root / "params" / "default.yaml").write_text("gs: ${oc.env:HOME}\ncs: 1\ncamera_height: 1\nextra: w\nchain: 1\n")
    with pytest.raises(ConfigError, match="resolver"):
        compose(root, "main")
    (root / "params" / "default.yaml").write_text("gs: ${params.missing}\ncs: 1\ncamera_height: 1\nextra: w\nchain: 1\n")
    with pytest.raises(ConfigError, match="'missing' not found"):
        compose(root, "main")


@pytest.mark.skipif(not REF_CONFIG.is_dir(), reason="reference tree not present")
def test_the_reference_config_tree_composes():
    c = compose(REF_CONFIG, "map_creation_cfg.yaml", ["scene_id=2"])
    assert c.map_config.grid_size == 1000 and c.map_config.cell_size == 0.05           # ${params.gs}, ${params.cs}
    assert c.map_config.depth_sample_rate == 100 and c.map_config.pose_info.pose_type == "mobile_base"
    assert c.map_config.cam_calib_mat == [540, 0, 540, 0, 540, 360, 0, 0, 1]
    assert c.params.sim_setting.sensor_height == 1.5 and c.params.controller_config.gs == 1000
    assert c.scene_id == 2 and c.nav.tasks_per_scene == 20 and c.map_config.obstacle_names[0] == "wall"
    # map_indexing_cfg.yaml names a data_paths option (lab_new) that is not in the repository: Hydra fails there too
    with pytest.raises(ConfigError, match="lab_new"):
        compose(REF_CONFIG, "map_indexing_cfg.yaml")
    c = compose(REF_CONFIG, "map_indexing_cfg.yaml", ["data_paths=default"])
    assert c.decay_rate == 0.01 and c.image_query_cfg.resolution.w == 1080


# ------------------------------------------------------------------------------------------ applications
def make_dataset(tmp_path):
    root = write_tree(tmp_path / "config")
    data = tmp_path / "data"
    for s in ("b_scene", "a_scene", "c_scene"):
        (data / "avlmaps_dataset" / s).mkdir(parents=True)
    (data / "avlmaps_dataset" / "stray_file.txt").write_text("x")
    return root, data


def fake_feature_fn(rgb):
    return np.zeros((1, 4, 2, 2), np.float32)


def test_create_map_application_flow(tmp_path, monkeypatch):
    from avlmaps_b200.application import create_map
    from avlmaps_b200.map import AVLMap

    root, data = make_dataset(tmp_path)
    seen = {}
    monkeypatch.setattr(AVLMap, "create_map", lambda self, d: seen.update(scene=Path(d), fn=self.vlmap.feature_fn, gs=self.vlmap.gs) or True)
    rc = create_map.main(["--config-dir", str(root), "--config-name", "main.yaml", "--feature-fn",
                          "test_config_apps_cpu:fake_feature_fn", f"data_paths.avlmaps_data_dir={data}", "scene_id=1", "params.gs=64"])
    assert rc == 0 and seen["scene"].name == "b_scene" and seen["gs"] == 64          # sorted dirs, files ignored
    assert seen["fn"](None).shape == (1, 4, 2, 2)
    with pytest.raises(SystemExit, match="out of range"):
        create_map.main(["--config-dir", str(root), "--config-name", "main", f"data_paths.avlmaps_data_dir={data}", "scene_id=9"])
    with pytest.raises(SystemExit, match="not a directory"):
        create_map.main(["--config-dir", str(root), "--config-name", "main", "data_paths.avlmaps_data_dir=/nonexistent"])


class FakeVLMap:
    grid_pos = np.arange(30, dtype=np.int32).reshape(10, 3)

    def __init__(self):
        self.encoder = None
        self.clip_inited = False

    def load_map(self, d):
        self.loaded = Path(d)
        return (Path(d) / "vlmap").exists()

    def set_text_encoder(self, enc, dim):
        self.encoder = (enc, dim)

    def _init_clip(self):
        self.clip_inited = True


class FakeAVLMap:
    last = None

    def __init__(self, config, data_dir=""):
        self.config, self.vlmap, self.calls = config, FakeVLMap(), []
        FakeAVLMap.last = self

    area_map = None
    sound_map = None

    def index_object(self, name, decay_rate=0.1):
        self.calls.append((name, decay_rate))
        heat = np.zeros(10, np.float32)
        heat[len(name) % 10] = 1.0
        return heat

    def index_area(self, name, decay_rate=0.1):
        self.calls.append(("area:" + name, decay_rate))
        return np.linspace(0, 1, 10, dtype=np.float32)

    def index_sound(self, name, decay_rate=0.01):
        self.calls.append(("sound:" + name, decay_rate))
        return np.linspace(1, 0, 10, dtype=np.float32)

    def get_max_pos_3d(self, heat):
        return self.vlmap.grid_pos[int(np.argmax(heat))]


def fake_text_encoder(texts):
    return np.ones((len(texts), 8), np.float32)


def test_index_map_application_flow(tmp_path, monkeypatch, capsys):
    from avlmaps_b200.application import index_map

    root, data = make_dataset(tmp_path)
    monkeypatch.setattr(index_map, "AVLMap", FakeAVLMap)
    common = ["--config-dir", str(root), "--config-name", "main.yaml", f"data_paths.avlmaps_data_dir={data}"]
    # a scene without a map: load_map prints and returns False, like the reference (vlmap.py:53-55)
    assert index_map.main(common + ["--object", "sofa"]) == 1
    (data / "avlmaps_dataset" / "a_scene" / "vlmap").mkdir()
    out = tmp_path / "heat"
    rc = index_map.main(common + ["--object", "sofa", "--object", "potted plant", "--out", str(out), "--text-encoder",
                                  "test_config_apps_cpu:fake_text_encoder", "--clip-dim", "8"])
    a = FakeAVLMap.last
    assert rc == 0 and a.calls == [("sofa", 0.01), ("potted plant", 0.01)] and a.vlmap.encoder[1] == 8
    assert a.vlmap.loaded.name == "a_scene" and not a.vlmap.clip_inited
    assert np.load(out / "heat_potted_plant.npy").shape == (10,)
    assert "goal voxel (row, col, height) = [12, 13, 14]" in capsys.readouterr().out
    # the prompt loop of the reference (index_map.py:33-142): object, an unsupported modality, exit
    answers = iter(["1", "chair", "2", "5", "6"])
    rc = index_map.main(common + ["decay_rate=0.05"], input_fn=lambda prompt: next(answers))
    a = FakeAVLMap.last
    assert rc == 0 and a.calls == [("chair", 0.05)] and a.vlmap.clip_inited
    out_text = capsys.readouterr().out
    assert "needs --sound-text-encoder" in out_text
    # area and sound modalities once their encoders are named on the command line
    made = {}

    class FakeArea:
        def __init__(self, data_dir, text_encoder=None, clip_feat_dim=768):
            made["area"] = (Path(data_dir).name, text_encoder, clip_feat_dim)

        def load_map(self, d):
            made["area_loaded"] = Path(d).name

    class FakeSound:
        def __init__(self, cats, text_encoder, logit_scale_at, difficulty_level=1):
            made["sound"] = (cats, text_encoder, round(float(logit_scale_at), 3), difficulty_level)

        def load_sound_map(self, d):
            made["sound_loaded"] = Path(d).name

    monkeypatch.setattr(index_map, "AreaMap", FakeArea)
    monkeypatch.setattr(index_map, "SoundMap", FakeSound)
    rc = index_map.main(common + ["--area", "kitchen", "--sound", "door knock", "--object", "sofa",
                                  "--area-text-encoder", "test_config_apps_cpu:fake_text_encoder",
                                  "--sound-text-encoder", "test_config_apps_cpu:fake_text_encoder",
                                  "--sound-categories", "door knock, dog ,", "+sound_data_collect_params.difficulty=2"])
    a = FakeAVLMap.last
    assert rc == 0 and a.calls == [("sofa", 0.01), ("area:kitchen", 0.01), ("sound:door knock", 0.01)]
    assert (made["area"][0], made["area"][1].__name__, made["area"][2]) == ("a_scene", "fake_text_encoder", 768)
    assert made["area_loaded"] == "a_scene" and made["sound_loaded"] == "a_scene"
    assert (made["sound"][0], made["sound"][1].__name__) + made["sound"][2:] == (["door knock", "dog"], "fake_text_encoder", 4.605, 2)
    answers = iter(["3", "kitchen", "4", "6"])
    assert index_map.main(common, input_fn=lambda prompt: next(answers)) == 0
    assert "needs --area-text-encoder" in capsys.readouterr().out


def test_generate_obstacle_map_application_flow(tmp_path, monkeypatch, capsys):
    """generate_obstacle_map.py:19-33 -> Map.create / load_map / generate_obstacle_map on a map file written here
    (the device upload of load_map is stubbed: no GPU in this suite)."""
    import avlmaps_b200.map.vlmap as vlmap_mod
    from avlmaps_b200.application import generate_obstacle_map
    from avlmaps_b200.utils import mapping_utils

    root = write_tree(tmp_path / "config")
    scene = tmp_path / "data" / "vlmaps_dataset" / "scene0"
    (scene / "vlmap").mkdir(parents=True)
    gs, vh = 1000, 30
    occ = -np.ones((gs, gs, vh), np.int32)
    occ[500:510, 400:420, 3] = np.arange(200, dtype=np.int32).reshape(10, 20) + 1
    occ[505, 405, 3] = 0                      # voxel id 0 counts as free (map.py:93 tests `> 0`)
    occ[600, 600, 0] = 7                      # height 0 is outside (h_min, h_max): ignored
    pos = np.argwhere(occ >= 0).astype(np.int32)
    mapping_utils.save_3d_map(scene / "vlmap" / "vlmaps.h5df", np.zeros((pos.shape[0], 4), np.float32), pos,
                              np.ones(pos.shape[0], np.float32), occ, [0], np.zeros((pos.shape[0], 3), np.uint8))
    monkeypatch.setattr(vlmap_mod, "DeviceMap", lambda feat, operand="f16": None)
    out = tmp_path / "obs"
    rc = generate_obstacle_map.main(["--config-dir", str(root), "--config-name", "main.yaml", "--out", str(out),
                                     f"data_paths.avlmaps_data_dir={tmp_path / 'data'}"])
    assert rc == 0
    obs = np.load(out / "obstacles.npy")
    assert obs.shape == (10, 20) and obs.dtype == bool and int((obs == 0).sum()) == 199 and obs[5, 5]
    assert (out / "obstacles.png").exists()
    assert "rows 500..509, cols 400..419, 199 occupied cells" in capsys.readouterr().out
    with pytest.raises(SystemExit, match="not a directory"):
        generate_obstacle_map.main(["--config-dir", str(root), "--config-name", "main", "--dataset-dir-name", "nope",
                                    f"data_paths.avlmaps_data_dir={tmp_path / 'data'}"])
