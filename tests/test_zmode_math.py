"""CPU check of the arithmetic the z-mode kernel chain implements: packed real columns -> column FFT
(pass 1) -> separation of the two real columns from rows ky / Ny-ky, two half-length row FFTs and a radix-2 step (pass 2) ->
|F|^2 or F1 conj(F2) written to rows ky and -ky with the fftshift folded into the store index.  numpy stands in for the
device FFTs; the index logic (self-mirrored rows, shift on/off, conjugate mirror) is the kernels'."""
import numpy as np
import pytest


def zpack(x, wx):
    """pass 1 (ColsR2CPack, z mode): w_x / 2 applied to the real and imaginary part of the packed columns, then the column FFT."""
    z = 0.5 * (wx[0::2] * x[:, 0::2] + 1j * wx[1::2] * x[:, 1::2])
    return np.fft.fft(z, axis=0)


def row_spectrum(Z, ky):
    """pass 2 loads (rowsz_power_kernel / rowszx_kernel): A = Z[ky] + conj Z[Ny-ky], B = -i (Z[ky] - conj Z[Ny-ky])."""
    Ny, M = Z.shape
    za, zb = Z[ky], Z[(Ny - ky) % Ny]
    A = (za.real + zb.real) + 1j * (za.imag - zb.imag)
    B = (za.imag + zb.imag) + 1j * (zb.real - za.real)
    FA, FB = np.fft.fft(A), np.fft.fft(B)
    w = np.exp(-2j * np.pi * np.arange(M) / (2 * M))
    return np.concatenate([FA + w * FB, FA - w * FB])


def mirrored_store(rows, Ny, Nx, shift, mirror):
    """the store loop shared by both pass-2 kernels: row ky -> output rows ky and -ky"""
    sy, sx = (Ny // 2, Nx // 2) if shift else (0, 0)
    out = np.full((Ny, Nx), np.nan, dtype=rows[0].dtype)
    for ky, c in enumerate(rows):
        rd, rm = (ky + sy) % Ny, (Ny - ky + sy) % Ny
        self_ = ky == 0 or 2 * ky == Ny
        for kx in range(Nx):
            if not self_ or 2 * kx <= Nx:
                out[rd, (kx + sx) % Nx] = c[kx]
            if not self_ or (0 < kx and 2 * kx < Nx):
                out[rm, (Nx - kx + sx) % Nx] = mirror(c[kx])
    assert not np.isnan(out).any(), "every output cell is written"
    return out


@pytest.mark.parametrize("shift", [True, False])
@pytest.mark.parametrize("shape", [(16, 32), (8, 64), (32, 8)])
def test_zmode_power_and_cross(shape, shift):
    Ny, Nx = shape
    rng = np.random.default_rng(Ny * 100 + Nx)
    x1, x2 = rng.standard_normal((Ny, Nx)), rng.standard_normal((Ny, Nx))
    wy, wx = np.hanning(Ny + 1)[:-1] + 0.1, np.hanning(Nx + 1)[:-1] + 0.2
    Z1, Z2 = zpack(wy[:, None] * x1, wx), zpack(wy[:, None] * x2, wx)
    F1, F2 = np.fft.fft2(x1 * wy[:, None] * wx[None, :]), np.fft.fft2(x2 * wy[:, None] * wx[None, :])
    rows1 = [row_spectrum(Z1, ky) for ky in range(Ny // 2 + 1)]
    rows2 = [row_spectrum(Z2, ky) for ky in range(Ny // 2 + 1)]
    for ky in range(Ny // 2 + 1):
        np.testing.assert_allclose(rows1[ky], F1[ky], atol=1e-10)
    fs = np.fft.fftshift if shift else (lambda a: a)
    power = mirrored_store([np.abs(r) ** 2 for r in rows1], Ny, Nx, shift, lambda v: v)
    np.testing.assert_allclose(power, fs(np.abs(F1) ** 2), atol=1e-9)
    cross = mirrored_store([a * np.conj(b) for a, b in zip(rows1, rows2)], Ny, Nx, shift, np.conj)
    np.testing.assert_allclose(cross, fs(F1 * np.conj(F2)), atol=1e-9)
