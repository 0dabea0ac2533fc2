"""Import the UNMODIFIED reference (edwarddgao/WordGesture-GAN).

Search order: $WGG_REFERENCE_ROOT, /root/reference (the build container), oracle/_ref/ - the copy that
oracle/build_ref.sh vendors there (git-ignored: reference sources never enter this repository's history; the
directory travels to the GPU box with the snapshot, so bench.py's reference arm, its reference_cuda leg and the
acceptance run execute the reference's OWN files there).  The reference's ``src/gan/__init__.py:21`` imports
``visualization`` which imports matplotlib (absent here), so stub modules are pre-inserted into ``sys.modules``
(SURVEY.md section 8c).  The reference files are never modified.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    for cand in (os.environ.get("WGG_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "src", "gan")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "gan"))


def _install_stubs():
    for name, attrs in (
        ("matplotlib", {}),
        ("matplotlib.pyplot", {"Axes": object, "Figure": object}),
        ("matplotlib.patches", {"Rectangle": object}),
    ):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(mod, k, v)
            sys.modules[name] = mod


def load_reference():
    """Returns a namespace with the reference's hot-path symbols."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from src.gan import models as ref_models
        from src.gan import losses as ref_losses
        from src.gan import trainer as ref_trainer
        from src.shared import config as ref_config
        from src.shared import utils as ref_utils
        from src.shared import keyboard as ref_keyboard
        from src.gan import evaluation as ref_evaluation
    ns = types.SimpleNamespace(
        models=ref_models,
        losses=ref_losses,
        trainer=ref_trainer,
        config=ref_config,
        utils=ref_utils,
        keyboard=ref_keyboard,
        evaluation=ref_evaluation,
        root=REFERENCE_ROOT,
    )
    return ns
