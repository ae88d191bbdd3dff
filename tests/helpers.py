"""Synthetic input generators shared by tests and bench."""
import torch


def clustered_boxes(g, n, n_centres, img_w=1000., img_h=600., jitter=6.0):
    c = torch.rand(n_centres, 2, generator=g) * torch.tensor([img_w, img_h])
    wh = torch.exp(torch.rand(n_centres, 2, generator=g) * 2.5 + 3.0)
    which = torch.randint(0, n_centres, (n,), generator=g)
    ctr = c[which] + torch.randn(n, 2, generator=g) * jitter
    sz = wh[which] * torch.exp(torch.randn(n, 2, generator=g) * 0.15)
    b = torch.cat([ctr - sz / 2, ctr + sz / 2], 1)
    b[:, 0::2] = b[:, 0::2].clamp(0, img_w)
    b[:, 1::2] = b[:, 1::2].clamp(0, img_h)
    return b


def rpn_like_rois(g, n, n_frames=1, img_w=1000., img_h=600.):
    """n proposals per frame: centres uniform, w/h log-uniform in [16,600]x[16,400], clipped."""
    out = []
    for f in range(n_frames):
        c = torch.rand(n, 2, generator=g) * torch.tensor([img_w, img_h])
        w = torch.exp(torch.rand(n, generator=g) * (torch.log(torch.tensor(600.)) - torch.log(torch.tensor(16.))) + torch.log(torch.tensor(16.)))
        h = torch.exp(torch.rand(n, generator=g) * (torch.log(torch.tensor(400.)) - torch.log(torch.tensor(16.))) + torch.log(torch.tensor(16.)))
        b = torch.stack([c[:, 0] - w / 2, c[:, 1] - h / 2, c[:, 0] + w / 2, c[:, 1] + h / 2], 1)
        b[:, 0::2] = b[:, 0::2].clamp(0, img_w)
        b[:, 1::2] = b[:, 1::2].clamp(0, img_h)
        out.append(torch.cat([torch.full((n, 1), float(f)), b], 1))
    return torch.cat(out, 0)
