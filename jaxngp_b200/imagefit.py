"""BASELINE config C1: the 2-D image fit of the reference (``models/imagefit.py:11-71`` ``ImageFitter``, hash-grid arm;
``app/imagefit.py:99-186`` training loop), rebuilt as a harness around the 2-D hash-grid kernels.

The reference's app no longer runs against its own encoder (it passes ``dim=`` and omits ``bound`` --
``models/imagefit.py:28-37`` vs ``models/encoders.py:58-83``; ``data.make_image_metadata`` raises), so this is the
model and step it describes, not a drop-in: uv in [0, 1]^2 -> HashGridEncoder(dim=2, L=16, T=2^20, F=2, N_min=16,
N_max=2^19) -> Dense(128) -> ReLU -> Dense(128) -> ReLU -> Dense(3) -> sigmoid (biases on, zero-initialised), MSE
against the pixel, ``optax.adam(lr, b1=0.9, b2=0.99, eps=1e-15)``.  The encoder's unit square is reached with
``pos = 2 uv - 1, bound = 1`` (``pos01 = uv``, encoders.py:87).

The encoder forward / gradient scatter are this package's kernels (csrc/hashgrid.cu, dim = 2: 4 corners per level);
the three small dense layers are library GEMMs and the optimizer is torch's Adam: C1 is a parity case for the 2-D
encoder, not a measured path.
"""
import torch

from . import encoders


def lecun_normal_(w: torch.Tensor, generator=None):
    """flax's ``nn.initializers.lecun_normal()`` for a Dense kernel [in, out]: truncated normal (+-2 sigma) with
    variance 1 / fan_in (imagefit.py:51)."""
    fan_in = w.shape[0]
    std = (1.0 / fan_in) ** 0.5 / 0.87962566103423978  # flax rescales by the std of the truncated unit normal
    torch.nn.init.trunc_normal_(w, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=generator)
    return w


class ImageFitter(torch.nn.Module):
    """``ImageFitter(encoding="hashgrid")`` of the reference.  Parameter names follow its flax tree:
    ``HashGridEncoder_0/latent codes stored on grid vertices``, ``linear1``, ``linear2``, ``color_predictor``
    (kernel [in, out] + bias)."""

    def __init__(self, T: int = 2 ** 20, N_max: int = 2 ** 19, device=None, generator=None):
        super().__init__()
        self.encoder = encoders.HashGridEncoder(L=16, T=T, F=2, N_min=16, N_max=N_max, dim=2, device=device,
                                                generator=generator)
        self.kernels = torch.nn.ParameterDict()
        self.biases = torch.nn.ParameterDict()
        for name, i, o in (("linear1", 32, 128), ("linear2", 128, 128), ("color_predictor", 128, 3)):
            self.kernels[name] = torch.nn.Parameter(lecun_normal_(torch.empty(i, o, device=device), generator))
            self.biases[name] = torch.nn.Parameter(torch.zeros(o, device=device))

    def mlp(self, enc: torch.Tensor) -> torch.Tensor:
        x = torch.relu(enc @ self.kernels["linear1"] + self.biases["linear1"])
        x = torch.relu(x @ self.kernels["linear2"] + self.biases["linear2"])
        return torch.sigmoid(x @ self.kernels["color_predictor"] + self.biases["color_predictor"])

    def forward(self, uv: torch.Tensor) -> torch.Tensor:
        """uv [..., 2] in [0, 1] -> rgb [..., 3] in [0, 1] (imagefit.py:16-71)."""
        if uv.shape[-1] != 2:
            raise AssertionError(f"uv must have a trailing dimension of 2, got {tuple(uv.shape)}")  # chex, imagefit.py:24
        flat = uv.reshape(-1, 2).to(torch.float32)
        enc, _ = self.encoder(flat * 2 - 1, 1.0)
        return self.mlp(enc).reshape(*uv.shape[:-1], 3)


def make_optimizer(model: ImageFitter, lr: float = 1e-3):
    """app/imagefit.py:131-141: optax.adam(lr, b1=0.9, b2=0.99, eps=1e-15) (eps outside the square root, no decay)."""
    return torch.optim.Adam(model.parameters(), lr=lr, betas=(0.9, 0.99), eps=1e-15)


def pixel_uv(idcs: torch.Tensor, width: int, height: int) -> torch.Tensor:
    """Pixel indices -> normalised coordinates in [0, 1] (utils/data.py ``make_image_metadata``: x / W, y / H)."""
    x, y = idcs % width, idcs // width
    return torch.stack([x.to(torch.float32) / width, y.to(torch.float32) / height], dim=-1)


def train_step(model: ImageFitter, optimizer, uv: torch.Tensor, rgb: torch.Tensor) -> torch.Tensor:
    """One batch of app/imagefit.py ``train_step``: mean squared error, Adam.  Returns the loss (device scalar)."""
    optimizer.zero_grad(set_to_none=True)
    loss = torch.square(model(uv) - rgb).mean()
    loss.backward()  # encoder: ngp_hashgrid_a1_backward scatters d_enc into the table gradient
    optimizer.step()
    return loss.detach()
