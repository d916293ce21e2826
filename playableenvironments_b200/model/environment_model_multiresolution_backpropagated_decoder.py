"""``model.architecture: playableenvironments_b200.model.environment_model_multiresolution_backpropagated_decoder`` (face 1 of the boundary): the
reference's own environment model (model/environment_model_multiresolution_backpropagated_decoder.py of the upstream tree, which must be on PYTHONPATH --
encoders, decoder and trainers are out of scope and stay the reference's code) with its ObjectComposer replaced by the B200 render
path.  train.py / train_autoencoder.py / play.py resolve this module by its dotted path (train.py:33-34, play.py:133-134) and run
unchanged; only the YAML string is edited."""
from .environment_model_glue import build_environment_model


def model(config):
    return build_environment_model("model.environment_model_multiresolution_backpropagated_decoder", config)
