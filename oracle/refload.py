"""Import the UNMODIFIED reference on CPU: from /root/reference in the build container, from its staged copy oracle/_ref/
(oracle/vendor_reference.py) on the GPU box.

TEST INFRASTRUCTURE ONLY - used by oracle/make_golden.py and tests that cross-check the oracle
against the live reference; those tests skip when /root/reference is absent (the GPU box).

The reference's hot-path modules pull in packages that are not installable offline but are not
arithmetic (transforms3d, imgviz, trimesh, ballpark, matplotlib; for the trainer also
pytorch_lightning, hydra, omegaconf, torch_scatter, ...).  They get permissive stand-in modules.
The one arithmetic dependency, torch_efficient_distloss.eff_distloss (renderer:30,101), is bound
to the oracle's restatement (clift_oracle.distortion_loss) - see the "parity unpinned" note there.
Nothing under /root/reference is copied or modified.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

import torch

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # oracle/vendor_reference.py (git-ignored)


def _resolve_root() -> str:
    """CLIFT_REFERENCE_ROOT, else /root/reference (build container), else the byte-for-byte staged copy of the path's
    modules under oracle/_ref/ (what travels to the GPU box)."""
    env = os.environ.get("CLIFT_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/model/renderer"):
        return "/root/reference"
    return _STAGED


REFERENCE_ROOT = _resolve_root()

_STUB_MODULES = [
    "transforms3d", "transforms3d.euler", "transforms3d.axangles", "transforms3d.quaternions",
    "trimesh", "ballpark", "imgviz", "matplotlib", "matplotlib.pyplot",
    "pytorch_lightning", "pytorch_lightning.utilities", "pytorch_lightning.strategies",
    "pytorch_lightning.callbacks", "pytorch_lightning.loggers", "pytorch_lightning.loggers.logger",
    "hydra", "omegaconf", "torch_scatter", "randomname", "pyquaternion", "imageio", "png", "h5py",
    "hdbscan", "dvis", "quaternion", "wandb",
]


class _Stub(types.ModuleType):
    """Module whose every attribute exists: Capitalised names are empty classes (usable as base
    classes), other names are sub-stubs; calling a stub returns an identity decorator."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name[0].isupper():
            val = type(name, (), {"__init__": lambda self, *a, **k: None})
        else:
            val = _Stub(self.__name__ + "." + name)
        setattr(self, name, val)
        return val

    def __call__(self, *args, **kwargs):
        if len(args) == 1 and callable(args[0]) and not kwargs:
            return args[0]
        return lambda fn: fn


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model", "renderer"))


_loaded = None


def load():
    """Returns a namespace with the reference modules (ray, tensorf, renderer, loss, trainer)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    from oracle import clift_oracle
    for name in _STUB_MODULES:
        if name in sys.modules:
            continue
        try:
            __import__(name)
        except Exception:
            mod = _Stub(name)
            mod.__path__ = []
            sys.modules[name] = mod
    sys.modules["pytorch_lightning"].LightningModule = torch.nn.Module
    dl = types.ModuleType("torch_efficient_distloss")
    dl.eff_distloss = clift_oracle.distortion_loss
    sys.modules["torch_efficient_distloss"] = dl
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import util.ray as ray
    import util.metrics as metrics
    import model.radiance_field.tensoRF as tensorf
    import model.renderer.panopli_tensoRF_renderer as renderer
    import model.loss.loss as loss
    import trainer.train_panopli_tensorf as trainer
    _loaded = SimpleNamespace(ray=ray, tensorf=tensorf, renderer=renderer, loss=loss, trainer=trainer,
                              metrics=metrics)
    return _loaded


def build_model(params, grid_dim, num_classes, max_instances, slow_fast=True, semantic_softmax=True,
                pe_sem=0, pe_ins=0, sem_grid_comps=None, ins_grid_comps=None):
    """Reference TensorVMSplit built like trainer/train_panopli_tensorf.py:55-65, weights from ``params``."""
    ref = load()
    model = ref.tensorf.TensorVMSplit(
        list(grid_dim), num_semantics_comps=(sem_grid_comps or 32,) * 3, num_instance_comps=(ins_grid_comps or 32,) * 3,
        num_semantic_classes=num_classes,
        dim_feature_instance=2 * max_instances if slow_fast else max_instances,
        output_mlp_semantics=torch.nn.Softmax(dim=-1) if semantic_softmax else torch.nn.Identity(),
        use_semantic_mlp=not sem_grid_comps, use_instance_mlp=not ins_grid_comps, use_feature_reg=False,
        use_distilled_features_semantic=False, use_distilled_features_instance=False,
        pe_sem=pe_sem, pe_ins=pe_ins, slow_fast_mode=slow_fast, use_proj=False)
    missing, unexpected = model.load_state_dict({k: v.clone() for k, v in params.items()}, strict=True)
    assert not missing and not unexpected
    return model


def build_renderer(aabb, grid_dim, semantic_softmax=True, stop_semantic_grad=True, step_ratio=0.5):
    ref = load()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        r = ref.renderer.TensoRFRenderer(aabb.clone(), list(grid_dim), stop_semantic_grad=stop_semantic_grad,
                                         semantic_weight_mode="softmax" if semantic_softmax else "none",
                                         step_ratio=step_ratio)
    return r


def slow_fast_loss(model, labels, features, confidences, use_ema=True):
    """Calls TensoRFTrainer.calculate_instance_clustering_loss unbound (trainer:230-313)."""
    ref = load()
    holder = SimpleNamespace(
        instance_loss_mode="slow_fast", use_delta=False, temperature=100.0,
        config=SimpleNamespace(use_proj=False), model=model, device=torch.device("cpu"))
    if use_ema:
        holder.ema_update_slownet = lambda s, f, m: ref.trainer.TensoRFTrainer.ema_update_slownet(holder, s, f, m)
    else:
        holder.ema_update_slownet = lambda s, f, m: None
    return ref.trainer.TensoRFTrainer.calculate_instance_clustering_loss(holder, labels, features, confidences, None)
