"""VQ-W-Net measurement harness -- BENCH / TEST INFRASTRUCTURE, not part of the quantiser product.

BASELINE.json's metric has two halves: VQ lookups/s (the hot path itself) and VQ-W-Net train slices/s (the hot path
inside its caller).  The reference network (`src/networks/vqwnet.py:13-152`, blocks `src/networks/blocks.py:9-61`)
cannot travel to the GPU box, so this file restates its topology with stock torch layers (cuDNN convolutions, the
part SURVEY section 8 leaves to the library) around a pluggable quantiser:

    x -> U-Net #1 (4 residual down stages 1->64->128->256->512, bottleneck 1024, 4 nearest-upsample stages)
      -> quantiser (emb_dim = 64, K = 512, momentum 0.99, eps 1e-5)        <- the hot path
      -> U-Net #2 (64->64->128->256->512, bottleneck 1024, 4 up stages) -> 1x1 conv -> tanh

`forward` returns the reference's dict (`recon`, `embed`, `commit_loss`, `ids` transposed to (b,h,w) and made
1-based, `vqwnet.py:108-111,147-152`).  `tests/test_wnet_harness.py` checks, when /root/reference is present, that the
parameter shapes match the reference network one for one and that both produce the same output from the same weights.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch
from torch import nn

WIDTHS = (64, 128, 256, 512, 1024)


def _conv_in_relu(cin: int, cout: int):
    return [nn.Conv2d(cin, cout, 3, padding=1), nn.InstanceNorm2d(cout), nn.ReLU(inplace=True)]


class ConvPair(nn.Sequential):
    """(conv3x3 - InstanceNorm - ReLU) x 2   (blocks.py:39-61, use_output_act=True)"""

    def __init__(self, cin: int, cout: int):
        super().__init__(*_conv_in_relu(cin, cout), *_conv_in_relu(cout, cout))


class DownStage(nn.Module):
    """relu(ConvPair(x) + IN(conv1x1(x))) -> (maxpool2 of it, it)   (blocks.py:21-36)"""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.shortcut = nn.Sequential(nn.Conv2d(cin, cout, 1, bias=False), nn.InstanceNorm2d(cout))
        self.body = ConvPair(cin, cout)

    def forward(self, x):
        full = torch.relu(self.body(x) + self.shortcut(x))
        return nn.functional.max_pool2d(full, 2), full


class UpStage(nn.Module):
    """nearest x2, concatenate the skip, ConvPair   (blocks.py:9-18)"""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.body = ConvPair(cin, cout)

    def forward(self, x, skip):
        x = nn.functional.interpolate(x, scale_factor=2, mode="nearest")
        return self.body(torch.cat([x, skip], dim=1))


class UNetHalf(nn.Module):
    def __init__(self, cin: int, w: Sequence[int]):
        super().__init__()
        chans = [cin, w[0], w[1], w[2], w[3]]
        self.down = nn.ModuleList(DownStage(chans[i], chans[i + 1]) for i in range(4))
        self.bottom = ConvPair(w[3], w[4])
        # deepest first: (w3 + w4) -> w3, (w2 + w3) -> w2, (w1 + w2) -> w1, (w1 + w0) -> w0   (vqwnet.py:39-42)
        self.up = nn.ModuleList([UpStage(w[3] + w[4], w[3]), UpStage(w[2] + w[3], w[2]),
                                 UpStage(w[1] + w[2], w[1]), UpStage(w[1] + w[0], w[0])])

    def forward(self, x):
        skips = []
        for stage in self.down:
            x, s = stage(x)
            skips.append(s)
        x = self.bottom(x)
        for stage in self.up:
            x = stage(x, skips.pop())
        return x


class WNetHarness(nn.Module):
    """`make_vq(emb_dim, dict_size)` builds the quantiser: this repo's CUDA `VQ` on the GPU arm, the CPU oracle on the
    reference arm."""

    def __init__(self, make_vq: Callable[[int, int], nn.Module], channels: int = 1, widths: Sequence[int] = WIDTHS,
                 dict_size: int = 512):
        super().__init__()
        self.first = UNetHalf(channels, widths)
        self.vq = make_vq(widths[0], dict_size)
        self.second = UNetHalf(widths[0], widths)
        self.head = nn.Conv2d(widths[0], channels, 1)
        for m in self.modules():                                   # initialize.py 'kaiming' flavour; values do not matter here
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, a=0, mode="fan_in")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x):
        embed = self.first(x)
        q, commit_loss, ids = self.vq(embed)
        ids = torch.transpose(ids, 1, 2)
        ids += 1
        recon = torch.tanh(self.head(self.second(q)))
        return {"recon": recon, "embed": embed, "commit_loss": commit_loss, "ids": ids}


def reference_key_order(h: "WNetHarness"):
    """Harness parameters in the order the reference network registers its own (vqwnet.py:32-71): used by the test
    that loads one network's weights into the other."""
    def pair(p):       # blocks.DoubleConv: conv, IN, relu, conv, IN, relu
        return [p[0].weight, p[0].bias, p[3].weight, p[3].bias]

    def half(u):
        out = []
        for d in u.down:                                           # ResBlock: downsample(conv1x1), double_conv
            out += [d.shortcut[0].weight] + pair(d.body)
        out += pair(u.bottom)
        for s in u.up:
            out += pair(s.body)
        return out
    return half(h.first) + half(h.second) + [h.head.weight, h.head.bias]
