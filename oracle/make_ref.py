"""Recipe for ``oracle/_ref/reference_path.zip`` (git-ignored, travels to the GPU box with the snapshot): the import closure of the
UPSTREAM ``model.object_composer.ObjectComposer`` as an importable zip, so that ``bench.py --impl reference`` / ``cpu_baseline`` /
``gpu_eager_baseline`` time the reference's own code instead of the oracle port.  Run in the build container
(``/root/reference`` present; ``__graft_entry__.build()`` calls it); nothing is copied into the tracked tree.

Test infrastructure, like everything under ``oracle/``: the product never imports it."""
from __future__ import annotations

import collections
import collections.abc
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("PE_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref", "reference_path.zip")


def install_shims(cpu: bool):
    """Harness-side shims of SURVEY 8c (python >= 3.10, numpy >= 1.24; on the CPU also the hard-coded ``.cuda()`` calls)."""
    import numpy as np
    import torch
    collections.Sequence = collections.abc.Sequence
    np.bool = bool
    if cpu:
        torch.Tensor.cuda = lambda self, *a, **k: self
        orig = torch.randn_like
        torch.randn_like = lambda t, **k: orig(t)


def main() -> str:
    if not os.path.isdir(os.path.join(REFERENCE, "model")):
        raise SystemExit(f"{REFERENCE} is not the upstream tree")
    install_shims(cpu=True)
    sys.path.insert(0, REFERENCE)
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests", "golden"))
    import scenes
    from model.object_composer import ObjectComposer
    for name in ("tennis_small", "minecraft_small"):          # the architecture strings of the shipped configs pull in the sub-models
        ObjectComposer(scenes.SCENES[name]()[0])
    import utils.tensor_batchifier  # noqa: F401  (the chunking helper of the caller, environment_model.py:474-521)
    files = sorted({m.__file__ for m in list(sys.modules.values())
                    if getattr(m, "__file__", None) and os.path.abspath(m.__file__).startswith(REFERENCE + os.sep)})
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with zipfile.ZipFile(OUT, "w", zipfile.ZIP_DEFLATED) as z:
        for f in files:
            z.write(f, os.path.relpath(f, REFERENCE))
    print(f"{OUT}: {len(files)} modules of the upstream tree")
    return OUT


if __name__ == "__main__":
    main()
