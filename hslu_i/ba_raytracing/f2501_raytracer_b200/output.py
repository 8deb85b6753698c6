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

from .renderer import ImageBuffer


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


class FileOutput:
    """FileOutput<W, H, P>: `render_buffer(&ImageBuffer)` writes `path` as PNG (ColorType::Rgb, BitDepth::Eight)."""

    def __init__(self, path: str | Path):
        self.path = Path(path)

    def render_buffer(self, buffer: ImageBuffer) -> None:
        rgb = np.ascontiguousarray(buffer.rgb(), np.uint8)  # (H, W, 3)
        h, w, _ = rgb.shape
        raw = np.concatenate([np.zeros((h, 1), np.uint8), rgb.reshape(h, w * 3)], 1).tobytes()  # filter type 0 per scanline
        png = (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
               _chunk(b"IDAT", zlib.compress(raw, 6)) + _chunk(b"IEND", b""))
        self.path.write_bytes(png)
