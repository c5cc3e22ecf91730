"""Synthetic particle sets of the BASELINE.json configurations (harness code: torch on the GPU for the
random numbers and FFTs; nothing here is on the product path).  Used by bench.py, the config tests
and scripts/config_sweep.py so that all three see the same points for the same seed."""
import torch


def uniform(n, seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return torch.rand((n, 3), device=device, generator=g)


def zeldovich(side, seed, device, rms_cells=1.5, index=-2.0):
    """BASELINE config 4: lattice (i+0.5)/side displaced by psi = grad(inverse-laplacian(delta)),
    delta a Gaussian field with power-law spectrum P(k) ~ k^index, rms displacement `rms_cells` lattice
    cells; wrapped into [0, 1], float32, (side^3, 3)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    k1 = torch.fft.fftfreq(side, d=1.0 / side, device=device)
    kz = torch.fft.rfftfreq(side, d=1.0 / side, device=device)
    k2 = k1[:, None, None] ** 2 + k1[None, :, None] ** 2 + kz[None, None, :] ** 2
    k2[0, 0, 0] = 1.0
    amp = k2 ** (index / 4.0)  # sqrt(P(k)), P ~ k^index
    amp[0, 0, 0] = 0.0
    re = torch.randn(k2.shape, device=device, generator=g)
    im = torch.randn(k2.shape, device=device, generator=g)
    delta_k = torch.complex(re * amp, im * amp)
    del re, im, amp
    pos = torch.empty((side ** 3, 3), device=device)
    lattice = (torch.arange(side, device=device, dtype=torch.float32) + 0.5) / side
    disp = []
    for kk in (k1[:, None, None], k1[None, :, None], kz[None, None, :]):
        psi_k = 1j * kk * delta_k / k2
        disp.append(torch.fft.irfftn(psi_k, s=(side, side, side)))
        del psi_k
    scale = rms_cells / side / torch.sqrt(sum((d ** 2).mean() for d in disp) / 3.0)
    for axis in range(3):
        shape = [1, 1, 1]
        shape[axis] = side
        coord = lattice.view(shape) + disp[axis] * scale
        pos[:, axis] = torch.remainder(coord, 1.0).reshape(-1)
    pos.clamp_(0.0, 1.0)
    return pos
