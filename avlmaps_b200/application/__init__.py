"""Drop-ins for the reference's two command-line applications (reference application/create_map.py,
application/index_map.py) on the B200 engine, driven by the reference's own `config/` YAML tree
(avlmaps_b200.config.compose stands in for Hydra)."""
