"""Film -> display image, following Image.hs:302-327 (getPixel, gamma, clamp) and Spectrum.hs:162-168."""
import numpy as np


def film_to_rgb(film: np.ndarray) -> np.ndarray:
    """[H][W]{w, X*w, Y*w, Z*w} -> linear sRGB (getPixel with splat weight 0 contribution)."""
    w = film[..., 0:1]
    xyz = np.where(w != 0, film[..., 1:4] / np.where(w != 0, w, 1), 0).astype(np.float32)
    m = np.array([[3.240479, -1.537150, -0.498535], [-0.969256, 1.875991, 0.041556], [0.055648, -0.204043, 1.057311]], np.float32)
    return xyz @ m.T


def film_xyz(film: np.ndarray) -> np.ndarray:
    w = film[..., 0:1]
    return np.where(w != 0, film[..., 1:4] / np.where(w != 0, w, 1), 0).astype(np.float32)


def image_xyz(film: np.ndarray, splat: np.ndarray = None, splat_weight: float = 0.0) -> np.ndarray:
    """getPixel (Image.hs:303-315): splat_weight * splat + film / weight, per pixel, in XYZ"""
    out = film_xyz(film)
    if splat is not None:
        out = (np.float32(splat_weight) * splat + out).astype(np.float32)
    return out


def to_png_bytes(rgb: np.ndarray) -> bytes:
    """gamma 2.2 + clamp + 8 bit (Image.hs:317-327); minimal PNG writer (no external deps)."""
    import struct, zlib
    img = (np.clip(np.clip(rgb, 0, None) ** (1 / 2.2), 0, 1) * 255).round().astype(np.uint8)
    h, w, _ = img.shape
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))
    def chunk(t, d): return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b"")
