"""Expose this package under the reference's import paths (``src.gan.models`` ...), so the bodies of
TRAIN_SCRIPT / EVAL_SCRIPT (train_gan.py:53-59, eval_gan.py:50-55) run unchanged on the B200 path."""
from __future__ import annotations

import sys
import types


def install_as_src(force: bool = False) -> None:
    from . import configs, gan_losses, gan_modules, gan_trainer, train_step

    if "src" in sys.modules and not force and not getattr(sys.modules["src"], "__wgg_dropin__", False):
        raise RuntimeError("a different `src` package is already imported; pass force=True to shadow it")
    table = {
        "src": None, "src.gan": None, "src.shared": None,
        "src.gan.models": gan_modules, "src.gan.losses": gan_losses, "src.gan.trainer": gan_trainer,
        "src.shared.config": configs, "src.shared.utils": train_step,
    }
    for name, target in table.items():
        if target is None:
            mod = types.ModuleType(name)
            mod.__path__ = []
            mod.__wgg_dropin__ = True
        else:
            mod = target
        sys.modules[name] = mod
    for name in table:
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, sys.modules[name])
    for pkg, mods in (("src.gan", (gan_modules, gan_losses, gan_trainer)), ("src.shared", (configs, train_step))):
        for m in mods:
            for k, v in vars(m).items():
                if not k.startswith("_"):
                    setattr(sys.modules[pkg], k, v)
