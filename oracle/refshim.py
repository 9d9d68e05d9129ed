"""Import shim that makes the UNMODIFIED reference importable in this container (TEST INFRASTRUCTURE ONLY).

Used by oracle/make_golden.py (and bench.py --impl reference when /root/reference is present) to run the real
reference modules: the diffusers fork vendored at /root/reference/mustango/diffusers/src/diffusers, the AudioLDM VAE /
HiFi-GAN under /root/reference/audioldm and /root/reference/models.py. Nothing is copied: the packages are registered
with their __path__ pointing into /root/reference and a few optional third-party modules that are not installed here
(soundfile, progressbar, librosa) are stubbed, exactly as described in SURVEY.md §0 F4.
"""
from __future__ import annotations

import os
import sys
import types

REF = os.environ.get("TANGO_REFERENCE", "/root/reference")
DIFFUSERS_SRC = os.path.join(REF, "mustango", "diffusers", "src", "diffusers")


def available() -> bool:
    return os.path.isdir(DIFFUSERS_SRC) and os.path.isfile(os.path.join(REF, "models.py"))


def _pkg(name: str, path: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


_installed = False


def install() -> None:
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    import torch  # noqa: F401
    import transformers  # noqa: F401  (must be imported before the stubs below)
    # resolve the lazy transformers symbols /root/reference/models.py imports while soundfile/librosa are still absent
    from transformers import AutoModel, AutoTokenizer, CLIPTextModel, CLIPTokenizer, T5EncoderModel  # noqa: F401
    import huggingface_hub
    import huggingface_hub.constants as hc
    if not hasattr(hc, "hf_cache_home"):
        hc.hf_cache_home = "/tmp/hf_cache"
    if not hasattr(huggingface_hub, "HfFolder"):
        huggingface_hub.HfFolder = type("HfFolder", (), {})
    if not hasattr(huggingface_hub, "cached_download"):
        huggingface_hub.cached_download = lambda *a, **k: None
    d = _pkg("diffusers", DIFFUSERS_SRC)
    d.__version__ = "0.15.0.dev0"
    # import the fork's modules BEFORE stubbing librosa (its import_utils probes find_spec("librosa"))
    import diffusers.models.unet_2d_condition  # noqa: F401
    import diffusers.schedulers.scheduling_ddim  # noqa: F401
    import diffusers.schedulers.scheduling_ddpm  # noqa: F401
    # the names /root/reference/models.py pulls from the package root (the fork's __init__ is skipped on purpose)
    d.DDPMScheduler = diffusers.schedulers.scheduling_ddpm.DDPMScheduler
    d.DDIMScheduler = diffusers.schedulers.scheduling_ddim.DDIMScheduler
    d.UNet2DConditionModel = diffusers.models.unet_2d_condition.UNet2DConditionModel
    import diffusers.models.autoencoder_kl
    d.AutoencoderKL = diffusers.models.autoencoder_kl.AutoencoderKL
    for nm in ("soundfile", "progressbar", "librosa", "librosa.util", "librosa.filters"):
        if nm not in sys.modules:
            sys.modules[nm] = types.ModuleType(nm)
    sys.modules["librosa"].util = sys.modules["librosa.util"]
    sys.modules["librosa"].filters = sys.modules["librosa.filters"]
    for fn in ("pad_center", "tiny"):
        setattr(sys.modules["librosa.util"], fn, lambda *a, **k: None)
    sys.modules["librosa.filters"].mel = lambda *a, **k: None
    _pkg("audioldm", os.path.join(REF, "audioldm"))
    if REF not in sys.path:
        sys.path.append(REF)
    _installed = True


def unet_class():
    install()
    from diffusers.models.unet_2d_condition import UNet2DConditionModel
    return UNet2DConditionModel


def schedulers():
    install()
    from diffusers.schedulers.scheduling_ddim import DDIMScheduler
    from diffusers.schedulers.scheduling_ddpm import DDPMScheduler
    return DDPMScheduler, DDIMScheduler


def autoencoder_class():
    install()
    from audioldm.variational_autoencoder.autoencoder import AutoencoderKL
    return AutoencoderKL


def audio_diffusion_module():
    """/root/reference/models.py as a module (AudioDiffusion.inference is driven unbound with a stub self)."""
    install()
    import importlib
    return importlib.import_module("models")
