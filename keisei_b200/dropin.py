"""Install keisei_b200 behind the reference's hot-path seams without touching its sources.

    import keisei_b200.dropin as kb
    kb.install_into_reference()      # before KataGoTrainingLoop / the tests import the names
    ...
    kb.uninstall_from_reference()    # restores every patched attribute

What is swapped (and nothing else):
  * `keisei.training.model_registry._REGISTRY["se_resnet"]` (and `"resnet"`)  -> the B200 models; the reference's params
    dataclasses are kept, and the adapter classes are registered as virtual subclasses of the reference model classes,
    so `isinstance(model, keisei...SEResNetModel)` checks (reference tests/test_se_resnet.py:129) keep passing;
  * `KataGoPPOAlgorithm` / `KataGoRolloutBuffer` in `keisei.training.katago_ppo` and in every already-imported `keisei.*`
    module that bound those names (`katago_loop.py:47-51` imports them by name); `KataGoPPOParams` stays the reference's
    dataclass (`katago_loop.py:538-541` isinstance check) — the trainer only reads its fields;
  * `keisei.training.gae.compute_gae{,_padded,_gpu,_padded_gpu}` -> the kernel-backed functions of `keisei_b200.gae`
    (same signatures); `update()` resolves them through that module's attributes at call time, so a test or caller that
    patches `keisei.training.gae.compute_gae_padded` is still seen (reference tests/test_split_merge_gae_opt.py:336-372);
  * `keisei.training.katago_loop.split_merge_step` (when that module is importable: it pulls in the reference's whole
    loop) -> `keisei_b200.split_merge.split_merge_step`: same signature and result type, host-side partition, grouped
    CUDA-graph forwards for the learner + opponents (reference katago_loop.py:284-431, called at :1179-1198).
"""
from __future__ import annotations

import abc
import sys
from typing import Any

from . import gae as kb_gae
from . import katago_ppo as kb_ppo
from . import split_merge as kb_split_merge
from .models.resnet import ResNetModel, ResNetParams
from .models.se_resnet import SEResNetModel, SEResNetParams

_PATCHED: list[tuple[Any, str, Any]] = []     # (owner, attribute / key, original)
_REGISTRY_PATCHED: list[tuple[dict, str, Any]] = []
_GAE_NAMES = ("compute_gae", "compute_gae_padded", "compute_gae_gpu", "compute_gae_padded_gpu")
_TRAINER_NAMES = ("KataGoPPOAlgorithm", "KataGoRolloutBuffer")


def _adapter(ours: type, ours_params: type, ref_model_cls: type, name: str) -> type:
    """Subclass of the B200 model that accepts the REFERENCE's params object, keeps it as `.params`, and answers
    isinstance() against the reference class (virtual subclass: the reference models derive from abc.ABC)."""
    fields = tuple(ours_params.__dataclass_fields__)

    class _Adapter(ours):  # type: ignore[misc, valid-type]
        def __init__(self, params: Any) -> None:
            # validate through this package's dataclass, then keep the REFERENCE object as `.params`: same field names,
            # and it is what reference code reads back (checkpoint metadata, tests)
            super().__init__(ours_params(**{f: getattr(params, f) for f in fields if hasattr(params, f)}))
            self.params = params

    _Adapter.__name__ = _Adapter.__qualname__ = name
    for base in ref_model_cls.__mro__:
        if isinstance(base, abc.ABCMeta) and base is not abc.ABC:
            base.register(_Adapter)
    return _Adapter


def installed() -> bool:
    return bool(_PATCHED or _REGISTRY_PATCHED)


def install_into_reference(resnet: bool = True, split_merge: bool = True) -> None:
    """Idempotent. Needs an importable `keisei` (the reference); raises ImportError otherwise."""
    if installed():
        return
    import keisei.training.gae as ref_gae                      # noqa: PLC0415 — optional dependency
    import keisei.training.katago_ppo as ref_ppo               # noqa: PLC0415
    import keisei.training.model_registry as ref_reg           # noqa: PLC0415

    def swap_entry(arch: str, ours: type, ours_params: type) -> None:
        old = ref_reg._REGISTRY[arch]
        cls = _adapter(ours, ours_params, old.model_cls, old.model_cls.__name__)
        _REGISTRY_PATCHED.append((ref_reg._REGISTRY, arch, old))
        ref_reg._REGISTRY[arch] = ref_reg.ArchitectureSpec(cls, old.params_cls, old.contract, old.obs_channels)

    swap_entry("se_resnet", SEResNetModel, SEResNetParams)
    if resnet and "resnet" in ref_reg._REGISTRY:
        swap_entry("resnet", ResNetModel, ResNetParams)

    originals = {n: getattr(ref_ppo, n) for n in _TRAINER_NAMES}
    replacements = {n: getattr(kb_ppo, n) for n in _TRAINER_NAMES}
    for mod_name, mod in list(sys.modules.items()):
        if mod is None or not (mod_name == "keisei" or mod_name.startswith("keisei.")):
            continue
        for n in _TRAINER_NAMES:
            if getattr(mod, n, None) is originals[n]:
                _PATCHED.append((mod, n, originals[n]))
                setattr(mod, n, replacements[n])
    for n in _GAE_NAMES:
        _PATCHED.append((ref_gae, n, getattr(ref_gae, n)))
        setattr(ref_gae, n, getattr(kb_gae, n))
        if getattr(ref_ppo, n, None) is _PATCHED[-1][2]:        # katago_ppo.py:15 binds compute_gae_gpu by name
            _PATCHED.append((ref_ppo, n, getattr(ref_ppo, n)))
            setattr(ref_ppo, n, getattr(kb_gae, n))
    if split_merge:
        try:
            import keisei.training.katago_loop as ref_loop     # noqa: PLC0415 — needs the loop's own dependencies
        except ImportError:
            ref_loop = None
        if ref_loop is not None and hasattr(ref_loop, "split_merge_step"):
            _PATCHED.append((ref_loop, "split_merge_step", ref_loop.split_merge_step))
            ref_loop.split_merge_step = kb_split_merge.split_merge_step
            for n in _TRAINER_NAMES:                            # the loop module was imported just now: same swap as above
                if getattr(ref_loop, n, None) is originals[n]:
                    _PATCHED.append((ref_loop, n, originals[n]))
                    setattr(ref_loop, n, replacements[n])


def uninstall_from_reference() -> None:
    while _PATCHED:
        owner, name, orig = _PATCHED.pop()
        setattr(owner, name, orig)
    while _REGISTRY_PATCHED:
        reg, key, orig = _REGISTRY_PATCHED.pop()
        reg[key] = orig
