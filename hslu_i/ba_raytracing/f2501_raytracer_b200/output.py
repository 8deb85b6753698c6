"""Output side of the host API: the step after the render path.

`FileOutput(path).render_buffer(buffer)` mirrors the reference's `FileOutput::render_buffer`
(src/output/file.rs:27-49): the u32 framebuffer (0xFFRRGGBB for hit pixels, 0 for untouched ones) is
written as an 8-bit RGB PNG, linear values, no gamma (src/output/file.rs:61-65).  Plain zlib + struct: no
imaging dependency.
"""
from __future__ import annotations

import struct
import zlib
from pathlib import Path

import numpy as np

import ctypes as C

from . import _lib as L
from .renderer import ImageBuffer


def get_feature_string(features: str = "") -> str:
    """Output::get_feature_string() (src/output/mod.rs:25-88) for a Cargo feature list."""
    out = C.create_string_buffer(512)
    L.check(L.load().rth_feature_string(features.encode(), 0, out, len(out)))
    return out.value.decode("utf-8")


def window_title(features: str = "") -> str:
    """The title WindowOutput::new sets: "Raytracer - <feature string>" (src/output/window.rs:58-64)."""
    out = C.create_string_buffer(512)
    L.check(L.load().rth_feature_string(features.encode(), 1, out, len(out)))
    return out.value.decode("utf-8")


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


class FileOutput:
    """FileOutput<W, H, P>: `render_buffer(&ImageBuffer)` writes `path` as PNG (ColorType::Rgb, BitDepth::Eight)."""

    def __init__(self, path: str | Path, features: str | None = None):
        """`features`: when given, the feature string is stored in the PNG as a tEXt chunk ("Title"), the file-side
        counterpart of the reference's window title (the reference's own PNG carries no metadata, and a tEXt chunk
        changes no pixel)."""
        self.path = Path(path)
        self.title = window_title(features) if features is not None else None

    def render_buffer(self, buffer: ImageBuffer) -> None:
        rgb = np.ascontiguousarray(buffer.rgb(), np.uint8)  # (H, W, 3)
        h, w, _ = rgb.shape
        raw = np.concatenate([np.zeros((h, 1), np.uint8), rgb.reshape(h, w * 3)], 1).tobytes()  # filter type 0 per scanline
        text = b""
        if self.title:  # tEXt is Latin-1: U+00D7 is representable
            text = _chunk(b"tEXt", b"Title\x00" + self.title.encode("latin-1", "replace"))
        png = (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + text +
               _chunk(b"IDAT", zlib.compress(raw, 6)) + _chunk(b"IEND", b""))
        self.path.write_bytes(png)
