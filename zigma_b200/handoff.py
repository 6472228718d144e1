"""What happens to the sampler's latents after the hot path (/root/reference/sample_acc.py:355-392):
VAE decode + conversion to uint8 pixels for the image metrics / PNG files.  The VAE itself is a third-party
``diffusers.AutoencoderKL`` (absent from the reference tree and from this image): any object whose ``decode(latents)`` returns
something with a ``.sample`` tensor (or a tensor) works.  Plain PyTorch on purpose -- it runs once per sampled batch, after 49
denoiser evaluations, and is not part of the path this repository accelerates."""
import torch

LATENT_SCALE = 0.18215        # the Stable-Diffusion VAE's latent scaling, hard-coded at sample_acc.py:364,371


def _decoded(vae, z):
    out = vae.decode(z)
    return out.sample if hasattr(out, "sample") else out


@torch.no_grad()
def decode_latents(latents, vae, video=False, from_flow=True):
    """sample_acc.py:362-379.  ``from_flow``: latents produced by the flow-matching sampler are divided by LATENT_SCALE first;
    ground-truth latents drawn straight from the LDM encoder are not (the reference's comment at :367).
    Images: (N, C, h, w) -> (N, 3, H, W).  Videos: the reference decodes ``latents[i]`` for every index of the FIRST axis and
    stacks the results along dim 1 -- reproduced as is."""
    z = latents / LATENT_SCALE if from_flow else latents
    if not video:
        return _decoded(vae, z)
    return torch.stack([_decoded(vae, z[i]) for i in range(len(z))], dim=1)


def to_uint8_pixels(images):
    """``samples_pil`` of sample_acc.py:318-320 followed by the uint8 cast of :385-386: clamp(127.5 x + 128, 0, 255)."""
    return torch.clamp(127.5 * images + 128.0, 0, 255).to(torch.uint8)


@torch.no_grad()
def sample_and_decode(model, vae, z0, num_steps=50, y=None, video=False):
    """The sampling job of sample_acc.py:355-386 for one batch on one GPU: fixed-grid Euler sampling on the engine (one CUDA
    graph replay for the whole loop), VAE decode, uint8 pixels.  Returns (latents, uint8 images)."""
    latents = model.sample_euler(z0, num_steps=num_steps, y=y)
    return latents, to_uint8_pixels(decode_latents(latents.float(), vae, video=video))
