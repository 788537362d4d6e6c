"""Textures — host-side mirror of the reference's `texture.go`.

Only loading lives here (load-time, host).  Sampling (`Texture.Sample`,
texture.go:69-89) is part of the hot path and runs on the GPU
(`csrc/raster.cuh`, `sample_texture`).
"""
from __future__ import annotations

import numpy as np

from .utils import isPowerOfTwo

# texture.go:11-15
TextureTypeSolidColor = 0
TextureTypeImage = 1
TextureTypeImageFast = 2


class Texture:
    """texture.go:19-26.  `pixels` is (height, width, 4) uint8, premultiplied RGBA."""

    def __init__(self, typ=TextureTypeSolidColor, color=(0, 0, 0, 0), pixels=None, scale=1.0):
        self.typ = int(typ)
        self.color = tuple(int(c) & 0xFF for c in color)
        self.pixels = pixels
        self.width = 0 if pixels is None else int(pixels.shape[1])
        self.height = 0 if pixels is None else int(pixels.shape[0])
        self.widthF = np.float32(self.width)
        self.heightF = np.float32(self.height)
        self.scale = np.float32(scale)

    def SetScale(self, scale) -> None:
        """texture.go:65-67."""
        self.scale = np.float32(scale)


def NewColorTexture(c) -> Texture:
    """texture.go:28-33."""
    return Texture(TextureTypeSolidColor, color=c)


def premultiply_nrgba(rgba: np.ndarray) -> np.ndarray:
    """`color.RGBAModel.Convert` of an `image.NRGBA` pixel (texture.go:57).

    Go: r = R; r |= r<<8; r *= A; r /= 0xff; uint8(r >> 8); alpha kept.
    """
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    a = rgba[..., 3:4].astype(np.uint32)
    c = rgba[..., :3].astype(np.uint32)
    c = ((c * 0x101) * a // 0xFF) >> 8
    out = np.empty_like(rgba)
    out[..., :3] = c.astype(np.uint8)
    out[..., 3] = rgba[..., 3]
    return out


def NewImageTexture(img) -> Texture:
    """texture.go:35-63.  `img` is a PIL image or an (H, W, 4) non-premultiplied uint8 array."""
    if isinstance(img, np.ndarray):
        rgba = img
    else:
        rgba = np.asarray(img.convert("RGBA"), dtype=np.uint8)
    if rgba.ndim != 3 or rgba.shape[2] != 4:
        raise ValueError("NewImageTexture expects an RGBA image")
    height, width = rgba.shape[:2]
    typ = TextureTypeImage
    if isPowerOfTwo(width) and isPowerOfTwo(height):
        typ = TextureTypeImageFast
    return Texture(typ, pixels=premultiply_nrgba(rgba), scale=1.0)


def LoadTextureFile(filename: str, strict: bool = False) -> Texture:
    """texture.go:91-103 (decode via PIL instead of Go's image/*).

    Texel parity with the reference holds for what PIL and Go decode identically: 8-bit greyscale / RGB / RGBA PNGs
    (the reference's own `models/textures-16.png` is one).  JPEG (Go's image/jpeg uses another IDCT and YCbCr -> RGB
    conversion), 16-bit PNGs (Go narrows after the premultiply, PIL before) and paletted PNGs with transparency can
    differ by a few codes; they load with a warning, or raise with `strict=True`."""
    import warnings

    from PIL import Image

    with Image.open(filename) as im:
        im.load()
        exact = im.format == "PNG" and im.mode in ("L", "LA", "RGB", "RGBA")
        if not exact:
            msg = (f"{filename}: {im.format} / mode {im.mode} is outside the decode domain in which texels are identical to "
                   "the Go reference's (8-bit L / LA / RGB / RGBA PNG)")
            if strict:
                raise ValueError(msg)
            warnings.warn(msg)
        return NewImageTexture(im)
