"""Config-string model registry (reference: ``getattr(importlib.import_module(cfg["architecture"]), "model")``,
model/object_composer.py:53-54, model/nerf_models/ray_bending_style_nerf_model.py:36-37).

The reference's own dotted paths (``model.nerf_models.adain_style_nerf_model`` ...) resolve to the modules of this
package, so shipped YAML files work unchanged; fully qualified ``playableenvironments_b200.model...`` paths work too.
"""
import importlib

PACKAGE = "playableenvironments_b200"
_REPLACED_PREFIXES = ("model.nerf_models.", "model.object_composer", "model.positional_encoder",
                      "model.annealable_positional_encoder", "model.layers.adain", "utils.lib_3d.", "utils.tensor_")


def resolve(architecture: str):
    """Returns the python module implementing ``architecture``."""
    if architecture.startswith(PACKAGE + "."):
        return importlib.import_module(architecture)
    if architecture.startswith(_REPLACED_PREFIXES):
        return importlib.import_module(PACKAGE + "." + architecture)
    return importlib.import_module(architecture)


def build(architecture: str, *args, factory: str = "model"):
    return getattr(resolve(architecture), factory)(*args)
