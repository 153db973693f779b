"""Minimal binary-glTF (GLB) and PNG writers for synthetic scenes (tests and tools; pure data, no path
arithmetic).  The reader under test is yart_b200/host/glb.cpp."""
from __future__ import annotations

import json
import struct
import zlib

import numpy as np


def png_encode(img: np.ndarray, filters=None, palette=None, bit_depth=8) -> bytes:
    """img: (h, w) or (h, w, c) uint8 (uint16 for bit_depth 16); c in {1: grey, 2: grey+alpha, 3: RGB, 4: RGBA}.
    `filters`: per-row PNG filter types (0-4), default cycles through all five.  `palette`: (n, 3) uint8 →
    colour type 3 with img holding indices."""
    if img.ndim == 2:
        img = img[..., None]
    h, w, c = img.shape
    ctype = 3 if palette is not None else {1: 0, 2: 4, 3: 2, 4: 6}[c]
    if bit_depth == 16:
        raw_rows = img.astype(">u2").reshape(h, -1).view(np.uint8)
    else:
        raw_rows = img.astype(np.uint8).reshape(h, -1)
    bpp = c * (bit_depth // 8)
    out = bytearray()
    prev = np.zeros(raw_rows.shape[1], np.int32)
    for y in range(h):
        row = raw_rows[y].astype(np.int32)
        ft = (y % 5) if filters is None else filters[y % len(filters)]
        a = np.concatenate([np.zeros(bpp, np.int32), row[:-bpp]])
        b = prev
        cc = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])
        if ft == 0:
            pred = 0
        elif ft == 1:
            pred = a
        elif ft == 2:
            pred = b
        elif ft == 3:
            pred = (a + b) >> 1
        else:
            p = a + b - cc
            pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - cc)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, cc))
        out.append(ft)
        out += ((row - pred) & 0xff).astype(np.uint8).tobytes()
        prev = row

    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xffffffff)
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, bit_depth, ctype, 0, 0, 0))
    if palette is not None:
        data += chunk(b"PLTE", np.asarray(palette, np.uint8).tobytes())
    comp = zlib.compress(bytes(out), 6)
    half = len(comp) // 2
    data += chunk(b"IDAT", comp[:half]) + chunk(b"IDAT", comp[half:])  # split on purpose
    return data + chunk(b"IEND", b"")


class GlbBuilder:
    def __init__(self):
        self.bin = bytearray()
        self.j = {"asset": {"version": "2.0"}, "buffers": [{"byteLength": 0}], "bufferViews": [], "accessors": [],
                  "images": [], "textures": [], "materials": [], "meshes": [], "nodes": [], "scenes": [{"nodes": []}],
                  "scene": 0}

    def _view(self, data: bytes, stride=None) -> int:
        while len(self.bin) % 4:
            self.bin.append(0)
        v = {"buffer": 0, "byteOffset": len(self.bin), "byteLength": len(data)}
        if stride:
            v["byteStride"] = stride
        self.bin += data
        self.j["bufferViews"].append(v)
        return len(self.j["bufferViews"]) - 1

    def accessor(self, arr: np.ndarray, kind: str) -> int:
        arr = np.ascontiguousarray(arr)
        ct = {np.dtype("float32"): 5126, np.dtype("uint8"): 5121, np.dtype("uint16"): 5123, np.dtype("uint32"): 5125}[arr.dtype]
        self.j["accessors"].append({"bufferView": self._view(arr.tobytes()), "componentType": ct, "count": int(arr.shape[0]),
                                    "type": kind})
        return len(self.j["accessors"]) - 1

    def interleaved(self, pos, nrm, uv):
        """POSITION/NORMAL/TEXCOORD_0 interleaved in one strided view (exercises byteStride + byteOffset)."""
        n = len(pos)
        rec = np.concatenate([pos, nrm, uv], axis=1).astype(np.float32)
        v = self._view(rec.tobytes(), stride=32)
        ids = []
        for off, kind in ((0, "VEC3"), (12, "VEC3"), (24, "VEC2")):
            self.j["accessors"].append({"bufferView": v, "byteOffset": off, "componentType": 5126, "count": n, "type": kind})
            ids.append(len(self.j["accessors"]) - 1)
        return ids

    def texture(self, png: bytes) -> int:
        self.j["images"].append({"bufferView": self._view(png), "mimeType": "image/png"})
        self.j["textures"].append({"source": len(self.j["images"]) - 1})
        return len(self.j["textures"]) - 1

    def material(self, m: dict) -> int:
        self.j["materials"].append(m)
        return len(self.j["materials"]) - 1

    def mesh(self, primitives: list) -> int:
        self.j["meshes"].append({"primitives": primitives})
        return len(self.j["meshes"]) - 1

    def node(self, n: dict, root=False) -> int:
        self.j["nodes"].append(n)
        i = len(self.j["nodes"]) - 1
        if root:
            self.j["scenes"][0]["nodes"].append(i)
        return i

    def tobytes(self) -> bytes:
        while len(self.bin) % 4:
            self.bin.append(0)
        self.j["buffers"][0]["byteLength"] = len(self.bin)
        js = json.dumps(self.j, separators=(",", ":")).encode()
        js += b" " * (-len(js) % 4)
        total = 12 + 8 + len(js) + 8 + len(self.bin)
        return (struct.pack("<III", 0x46546C67, 2, total) + struct.pack("<II", len(js), 0x4E4F534A) + js +
                struct.pack("<II", len(self.bin), 0x004E4942) + bytes(self.bin))
