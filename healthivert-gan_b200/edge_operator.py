"""Drop-in mirror of the reference's ``models/edge_operator.py`` Sobel (EEM edge extractor)."""
import torch
import torch.nn as nn

from . import _lib
from ._lib import check, ptr


class Sobel(nn.Module):
    """reference models/edge_operator.py:29-49: replicate-pad, 2 fixed 3x3 filters, magnitude,
    clamp to <= 1.  The ``filter.weight`` parameter is kept for state_dict compatibility only."""

    def __init__(self, requires_grad=False):
        super().__init__()
        self.filter = nn.Conv2d(1, 2, kernel_size=3, stride=1, padding=0, bias=False)
        gx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]])
        gy = torch.tensor([[1.0, 2.0, 1.0], [0.0, 0.0, 0.0], [-1.0, -2.0, -1.0]])
        self.filter.weight = nn.Parameter(torch.stack([gx, gy])[:, None], requires_grad=requires_grad)

    @torch.no_grad()
    def forward(self, img):
        n, c, h, w = img.shape
        assert c == 1, "Sobel expects a single-channel image"
        img = img.to(torch.float32).contiguous()
        out = torch.empty_like(img)
        check(_lib.lib().hv_sobel(ptr(img), ptr(out), n, h, w, _lib.stream()))
        return out


@torch.no_grad()
def edge_mse_loss(fake_mask, real_mask):
    """800 * mse(Sobel(fake), Sobel(real)) without materialising the edge maps
    (reference models/pix2pix_model.py:109,:263-264,:349).  Returns (loss [1] f32, xor_count [1] i64)."""
    n, c, h, w = fake_mask.shape
    fake_mask = fake_mask.to(torch.float32).contiguous()
    real_mask = real_mask.to(torch.float32).contiguous()
    cnt = torch.empty(1, device=fake_mask.device, dtype=torch.int64)
    loss = torch.empty(1, device=fake_mask.device, dtype=torch.float32)
    check(_lib.lib().hv_edge_xor_loss(ptr(fake_mask), ptr(real_mask), ptr(cnt), ptr(loss), n * c, h, w, _lib.stream()))
    return loss, cnt
