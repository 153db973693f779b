"""Generates the image fixtures of tests/test_image_decoders.py (committed next to this script).

Inputs: small PNG / JPEG / Radiance files written with Pillow, OpenCV and numpy — every container variant the decoders
of yart_b200/host/images.cpp handle.  Expected outputs: `<name>.expect.npy` = what the REFERENCE produces for that file
(oracle/_ref/oracle_ref texload → loadTexture<4>(…, NonColor, {0,1,2,3}), i.e. stb_image's RGBA8; hdrload →
loadTextureHDR), so the suite can run where /root/reference and the oracle are absent.

    python tests/golden/images/generate.py        (needs oracle/_ref/oracle_ref, Pillow, OpenCV)
"""
import io
import os
import struct
import subprocess
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
ORACLE = os.path.join(ROOT, "oracle", "_ref", "oracle_ref")


def picture(w, h, seed):
    """Smooth gradients + edges + noise: exercises every DCT coefficient and the chroma filters."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([127 + 120 * np.sin(x / 3.1 + seed) * np.cos(y / 4.7), 255 * ((x // 5 + y // 3) % 2), 255 * x / max(w - 1, 1)], -1)
    img += rng.normal(0, 18, img.shape)
    img[h // 3: h // 3 + 4, :, :] = (250, 10, 10)
    return np.clip(img, 0, 255).astype(np.uint8)


def png_raw(w, h, depth, ctype, rows, interlace=0, plte=None, trns=None):
    """Hand-assembled PNG (filter type 0..4 cycling per row) for the layouts Pillow will not write."""
    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xFFFFFFFF)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace))
    if plte is not None:
        out += chunk(b"PLTE", bytes(plte))
    if trns is not None:
        out += chunk(b"tRNS", bytes(trns))
    return out + chunk(b"IDAT", zlib.compress(rows, 9)) + chunk(b"IEND", b"")


def pack_rows(samples, depth, filt=True):
    """samples: (h, w, ch) integer array of `depth`-bit values → filtered scanlines (filter 0, or 1 = Sub on odd rows)."""
    h, w, ch = samples.shape
    out = bytearray()
    for yy in range(h):
        if depth == 16:
            row = samples[yy].astype(">u2").tobytes()
            bpp = 2 * ch
        elif depth == 8:
            row = samples[yy].astype(np.uint8).tobytes()
            bpp = ch
        else:
            bits = "".join(format(int(v), f"0{depth}b") for v in samples[yy].reshape(-1))
            bits += "0" * (-len(bits) % 8)
            row = bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8))
            bpp = 1
        if filt and yy % 2 == 1:
            f = bytes((row[i] - (row[i - bpp] if i >= bpp else 0)) & 255 for i in range(len(row)))
            out += b"\x01" + f
        else:
            out += b"\x00" + row
    return bytes(out)


def adam7_rows(samples, depth):
    passes = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]
    out = b""
    for x0, y0, dx, dy in passes:
        sub = samples[y0::dy, x0::dx]
        if sub.shape[0] and sub.shape[1]:
            out += pack_rows(sub, depth)
    return out


def radiance(rgbe, rle):
    h, w, _ = rgbe.shape
    out = b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0\n\n" + f"-Y {h} +X {w}\n".encode()
    for yy in range(h):
        if not rle:
            out += rgbe[yy].tobytes()
            continue
        out += bytes([2, 2, w >> 8, w & 255])
        for k in range(4):
            ch = rgbe[yy, :, k]
            i = 0
            while i < w:
                run = 1
                while i + run < w and run < 127 and ch[i + run] == ch[i]:
                    run += 1
                if run >= 3:
                    out += bytes([128 + run, int(ch[i])])
                    i += run
                else:
                    n = min(100, w - i)
                    out += bytes([n]) + ch[i:i + n].tobytes()
                    i += n
    return out


def main():
    from PIL import Image
    import cv2
    files = {}
    pic = picture(37, 29, 1)  # odd sizes: partial MCUs on both axes
    for name, kw in (("jpeg_444_q90", dict(quality=90, subsampling=0)), ("jpeg_422_q75", dict(quality=75, subsampling=1)),
                     ("jpeg_420_q60", dict(quality=60, subsampling=2)), ("jpeg_420_q30_optimized", dict(quality=30, subsampling=2, optimize=True)),
                     ("jpeg_progressive_420", dict(quality=80, subsampling=2, progressive=True)),
                     ("jpeg_progressive_444", dict(quality=95, subsampling=0, progressive=True))):
        b = io.BytesIO()
        Image.fromarray(pic).save(b, "JPEG", **kw)
        files[name + ".jpg"] = b.getvalue()
    b = io.BytesIO()
    Image.fromarray(pic[..., 0]).save(b, "JPEG", quality=85)
    files["jpeg_grey.jpg"] = b.getvalue()
    b = io.BytesIO()
    Image.fromarray(pic[..., 0]).save(b, "JPEG", quality=85, progressive=True)
    files["jpeg_grey_progressive.jpg"] = b.getvalue()
    big = picture(96, 64, 2)
    ok, enc = cv2.imencode(".jpg", big[..., ::-1], [cv2.IMWRITE_JPEG_QUALITY, 70, cv2.IMWRITE_JPEG_RST_INTERVAL, 3])
    files["jpeg_restart_interval.jpg"] = enc.tobytes()
    ok, enc = cv2.imencode(".jpg", big[..., ::-1], [cv2.IMWRITE_JPEG_QUALITY, 88, cv2.IMWRITE_JPEG_PROGRESSIVE, 1, cv2.IMWRITE_JPEG_RST_INTERVAL, 2])
    files["jpeg_progressive_restart.jpg"] = enc.tobytes()
    try:
        ok, enc = cv2.imencode(".jpg", big[..., ::-1], [cv2.IMWRITE_JPEG_QUALITY, 80, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411])
        files["jpeg_411.jpg"] = enc.tobytes()
        ok, enc = cv2.imencode(".jpg", big[..., ::-1], [cv2.IMWRITE_JPEG_QUALITY, 80, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440])
        files["jpeg_440.jpg"] = enc.tobytes()
    except Exception as e:  # noqa: BLE001
        print("no sampling-factor control in this OpenCV:", e)
    b = io.BytesIO()
    Image.fromarray(picture(1, 1, 3)).save(b, "JPEG", quality=90, subsampling=2)
    files["jpeg_1x1.jpg"] = b.getvalue()
    b = io.BytesIO()
    Image.fromarray(picture(8, 3, 4)).save(b, "JPEG", quality=90, subsampling=2)
    files["jpeg_8x3.jpg"] = b.getvalue()

    # PNG: Pillow for the common layouts, hand-assembled for the rest
    rgba = np.concatenate([pic, (picture(37, 29, 5)[..., :1])], -1)
    for name, arr, kw in (("png_rgb8", pic, {}), ("png_rgba8", rgba, {}), ("png_grey8", pic[..., 1], {}),
                          ("png_rgb8_interlaced_pil", pic, {})):
        b = io.BytesIO()
        Image.fromarray(arr).save(b, "PNG", **kw)
        files[name + ".png"] = b.getvalue()
    rng = np.random.default_rng(9)
    w, h = 21, 13
    files["png_rgb8_adam7.png"] = png_raw(w, h, 8, 2, adam7_rows(picture(w, h, 6).astype(np.int64), 8), interlace=1)
    files["png_rgba16_adam7.png"] = png_raw(w, h, 16, 6, adam7_rows(rng.integers(0, 65536, (h, w, 4)), 16), interlace=1)
    files["png_grey16.png"] = png_raw(w, h, 16, 0, pack_rows(rng.integers(0, 65536, (h, w, 1)), 16))
    files["png_greyalpha8.png"] = png_raw(w, h, 8, 4, pack_rows(rng.integers(0, 256, (h, w, 2)), 8))
    for d in (1, 2, 4):
        files[f"png_grey{d}.png"] = png_raw(w, h, d, 0, pack_rows(rng.integers(0, 1 << d, (h, w, 1)), d, filt=False))
        files[f"png_grey{d}_adam7.png"] = png_raw(w, h, d, 0, adam7_rows(rng.integers(0, 1 << d, (h, w, 1)), d), interlace=1)
        pal = rng.integers(0, 256, (1 << d) * 3).tolist()
        files[f"png_palette{d}.png"] = png_raw(w, h, d, 3, pack_rows(rng.integers(0, 1 << d, (h, w, 1)), d, filt=False), plte=pal)
    pal = rng.integers(0, 256, 256 * 3).tolist()
    files["png_palette8_trns.png"] = png_raw(w, h, 8, 3, pack_rows(rng.integers(0, 256, (h, w, 1)), 8), plte=pal, trns=rng.integers(0, 256, 100).tolist())
    key_img = rng.integers(0, 4, (h, w, 3)) * 60
    files["png_rgb8_colourkey.png"] = png_raw(w, h, 8, 2, pack_rows(key_img, 8), trns=[0, 60, 0, 120, 0, 0])
    files["png_grey4_colourkey.png"] = png_raw(w, h, 4, 0, pack_rows(rng.integers(0, 16, (h, w, 1)), 4, filt=False), trns=[0, 7])
    files["png_grey16_colourkey.png"] = png_raw(w, h, 16, 0, pack_rows(rng.integers(0, 3, (h, w, 1)) * 0x1234, 16), trns=[0x12, 0x34])

    # Radiance: run-length encoded, flat, and narrower than the RLE minimum
    rgbe = rng.integers(0, 256, (9, 40, 4)).astype(np.uint8)
    rgbe[..., 3] = rng.integers(120, 140, (9, 40))
    rgbe[2, 5:30, :] = rgbe[2, 5, :]  # long runs
    rgbe[4, :, 3] = 0                 # zero exponent → black
    files["radiance_rle.hdr"] = radiance(rgbe, True)
    files["radiance_flat.hdr"] = radiance(rgbe, False)
    files["radiance_narrow.hdr"] = radiance(rgbe[:, :5], False)

    for name, data in files.items():
        path = os.path.join(HERE, name)
        open(path, "wb").write(data)
        out = os.path.join(HERE, "_tmp.bin")
        if name.endswith(".hdr"):
            subprocess.run([ORACLE, "hdrload", path, out], check=True)
            raw = open(out, "rb").read()
            ww, hh = struct.unpack_from("<II", raw, 0)
            expect = np.frombuffer(raw, np.float32, ww * hh * 3, 8).reshape(hh, ww, 3)
        else:
            subprocess.run([ORACLE, "texload", path, "2", "4", "0,1,2,3", out], check=True)
            raw = open(out, "rb").read()
            ww, hh, _ = struct.unpack_from("<III", raw, 0)
            expect = np.frombuffer(raw, np.uint8, ww * hh * 4, 12).reshape(hh, ww, 4)
        os.unlink(out)
        np.save(os.path.join(HERE, name + ".expect.npy"), expect)
        print(f"{name:34s} {len(data):6d} B → {expect.shape}")


if __name__ == "__main__":
    sys.exit(main())
