"""Import shims that let the UNMODIFIED reference run on CPU in the build container.

TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py and the container-only pin tests).  The reference
lives read-only at /root/reference and does not exist on the GPU box; nothing on the product path, in the
``-m gpu`` tests, in ``smoke()`` or in ``bench.py`` imports this module.

Shims: stub modules for absent third-party imports, the removed numpy aliases
``np.int``/``np.float``, ``torch.cuda.FloatTensor`` on a CUDA-less box, and a seeded random-init VGG16 in place
of the ImageNet download.
"""
from __future__ import annotations

import functools
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("AESR_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "networks"))


_STUBS = ["skimage", "skimage.metrics", "skimage.measure", "SimpleITK", "batchgenerators",
          "batchgenerators.transforms", "batchgenerators.transforms.spatial_transforms",
          "batchgenerators.transforms.abstract_transforms", "matplotlib", "matplotlib.pyplot", "matplotlib.cm",
          "imageio", "gpustat", "nibabel", "torchsummary"]


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


def _stub_getattr(attr):
    if attr.startswith("__"):
        raise AttributeError(attr)
    return _Anything


def bootstrap() -> None:
    """Idempotent.  After this, ``import networks.acai_vanilla`` etc. resolve to the reference."""
    if getattr(bootstrap, "_done", False):
        return
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    np.int, np.float = int, float        # removed numpy aliases (never alias np.bool)
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            __import__(name)
            continue
        except Exception:
            pass
        mod = types.ModuleType(name)
        mod.__getattr__ = _stub_getattr                       # any public attribute -> inert object
        mod.__path__ = []
        sys.modules[name] = mod
    sys.modules["SimpleITK"].Image = type("Image", (), {})
    sys.modules["batchgenerators.transforms.abstract_transforms"].AbstractTransform = type("AbstractTransform", (), {})
    sys.modules["batchgenerators.transforms.abstract_transforms"].Compose = type("Compose", (), {})
    sys.modules["batchgenerators.transforms.spatial_transforms"].SpatialTransform = type("SpatialTransform", (), {})
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if not torch.cuda.is_available():
        torch.cuda.FloatTensor = lambda data, device=None: torch.FloatTensor(data)
    import torchvision
    import lpips.pretrained_networks as pn
    _tv_vgg16 = torchvision.models.vgg16        # pn.tv IS torchvision.models: capture before rebinding
    pn.tv.vgg16 = lambda pretrained=True, **k: _tv_vgg16(weights=None)
    bootstrap._done = True


def reference_args(dataset="ACDC", width=128, latent_width=32, latent=128, depth=32, lr=1e-5, ex_loss_weight1=0.05,
                   batch_size=12, epochs=10, device="cpu") -> dict:
    """CLI defaults (kwatsch/arguments.py) merged with NetworkConfig('ae_combined', dataset) the way
    train_cardiac_aesr.py:23-30 merges them."""
    bootstrap()
    from networks.net_config import NetworkConfig
    arch = NetworkConfig("ae_combined", dataset, ae_class="VanillaACAI").architecture
    args = dict(arch)
    args.update(dict(dataset=dataset, model="ae_combined", ae_class="VanillaACAI", width=width,
                     latent_width=latent_width, latent=latent, depth=depth, lr=lr, weight_decay=0.0, epochs=epochs,
                     device=device, gpu_ids=[0], ex_loss_weight1=ex_loss_weight1, use_percept_loss=False,
                     use_loss_annealing=False, get_masks=False, epoch_threshold=0, log_tensorboard=False,
                     image_mix_loss_func="perceptual", use_extra_latent_loss=False, batch_size=batch_size,
                     colors=1, n_res_block=None, use_batchnorm=True, use_sigmoid=True, use_laploss=False))
    return args


def make_reference_trainer(args: dict, model_seed: int = 892372, vgg_seed: int = 3, trainer="cardiac"):
    """Build the reference model + trainer on CPU: model under ``torch.manual_seed(model_seed)``, LPIPS VGG
    under ``torch.manual_seed(vgg_seed)``."""
    bootstrap()
    from networks.acai_vanilla import VanillaACAI
    if trainer == "cardiac":
        from kwatsch.cardiac.trainer_ae import AETrainerEndToEnd as T
    elif trainer == "brain":
        from kwatsch.brain.trainer_ae import AETrainerExtension1Brain as T
    else:
        from kwatsch.trainer_ae import AEBaseTrainer as T
    torch.manual_seed(model_seed)
    model = VanillaACAI(args)
    torch.manual_seed(vgg_seed)
    tr = T(args, model, eval_mode=False)
    return tr


class CpuEvalTrainer:
    """6-line trainer shim exposing encode/decode/predict in eval mode for the reference synthesis loops
    (BaseTrainer.decode hard-codes ``z.to('cuda')``, kwatsch/base_trainer.py:316-317)."""

    def __init__(self, model):
        self.model = model.eval()

    def encode(self, x, **kw):
        with torch.no_grad():
            return self.model.encode(x)

    def decode(self, z, **kw):
        with torch.no_grad():
            return self.model.decode(z)

    def predict(self, x, **kw):
        with torch.no_grad():
            return self.model(x)


def reference_generate_module():
    """generate_hr_volumes with its cuda-default ``latent_space_interp`` rebound to device='cpu'."""
    bootstrap()
    import generate_hr_volumes as ghv
    if not isinstance(ghv.latent_space_interp, functools.partial):
        ghv.latent_space_interp = functools.partial(ghv.latent_space_interp, device="cpu")
    return ghv


def reference_eval_common_module():
    bootstrap()
    import evaluate.common as ec
    if not isinstance(ec.latent_space_interp, functools.partial):
        ec.latent_space_interp = functools.partial(ec.latent_space_interp, device="cpu")
    ec.torch.cuda.empty_cache = lambda: None
    return ec


class cuda_to_cpu:
    """Context manager: on a CUDA-less box make ``tensor.to('cuda')`` a no-op, for the reference call sites that
    hard-code it (generate_hr_volumes.py:34, evaluate/common.py:182)."""

    def __enter__(self):
        self._orig = torch.Tensor.to
        if torch.cuda.is_available():
            return self
        orig = self._orig

        def to(t, *a, **k):
            if a and isinstance(a[0], str) and a[0].startswith("cuda"):
                return t
            return orig(t, *a, **k)
        torch.Tensor.to = to
        return self

    def __exit__(self, *exc):
        torch.Tensor.to = self._orig
        return False
