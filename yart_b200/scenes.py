"""Synthetic scene generators of the shapes BASELINE.json names, written as ".ysc" files
(layout: yart_b200/host/scene_desc.hpp).  Pure data generation (numpy), no path arithmetic.

The scene description mirrors the arguments of the reference's scene API
(ParametricBSDF ctor: reference src/bsdf/parametric.hpp:15-36; Mesh: src/core/mesh.hpp:54-61;
Node: src/core/scene.hpp:11-64; lights: src/core/light.hpp:76-171), so the same file feeds the
reference (through oracle/ref_driver.cpp) and the CUDA path (through ys_scene_load).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

LINEAR, SRGB, NONCOLOR = 0, 1, 2
AREA, IMAGE_INF, UNIFORM_INF = 0, 1, 2
IDENTITY = np.eye(4, dtype=np.float32)


@dataclass
class Texture:
    data: np.ndarray  # (h, w, c) uint8, or (h, w, 3) float32
    type: int = LINEAR


@dataclass
class Material:
    base: tuple = (1.0, 1.0, 1.0)
    base_tex: int = -1
    mr_tex: int = -1
    trans_tex: int = -1
    normal_tex: int = -1
    cc_tex: int = -1
    emis_tex: int = -1
    metallic: float = 0.0
    roughness: float = 0.0
    transmission: float = 0.0
    ior: float = 1.5
    anisotropic: float = 0.0
    aniso_rotation: float = 0.0
    clearcoat: float = 0.0
    clearcoat_roughness: float = 0.0
    emission: tuple = (0.0, 0.0, 0.0)
    normal_scale: float = 1.0
    thin: int = 0
    volume_color: tuple = (1.0, 1.0, 1.0)
    volume_density: float = 0.0

    def pack(self) -> bytes:
        return struct.pack(
            "<3f6i8f3ff i3ff".replace(" ", ""),
            *self.base, self.base_tex, self.mr_tex, self.trans_tex, self.normal_tex, self.cc_tex, self.emis_tex,
            self.metallic, self.roughness, self.transmission, self.ior, self.anisotropic, self.aniso_rotation,
            self.clearcoat, self.clearcoat_roughness, *self.emission, self.normal_scale, self.thin,
            *self.volume_color, self.volume_density)


@dataclass
class Mesh:
    positions: np.ndarray  # (n,3) f32
    normals: np.ndarray  # (n,3)
    tangents: np.ndarray  # (n,4)
    uvs: np.ndarray  # (n,2)
    faces: np.ndarray  # (m,4) u32: i0 i1 i2 material
    light_idx: np.ndarray = None  # (m,) i32

    def __post_init__(self):
        self.positions = np.ascontiguousarray(self.positions, np.float32).reshape(-1, 3)
        n = len(self.positions)
        self.normals = np.ascontiguousarray(self.normals, np.float32).reshape(n, 3)
        self.tangents = (np.zeros((n, 4), np.float32) if self.tangents is None
                         else np.ascontiguousarray(self.tangents, np.float32).reshape(n, 4))
        self.uvs = (np.zeros((n, 2), np.float32) if self.uvs is None
                    else np.ascontiguousarray(self.uvs, np.float32).reshape(n, 2))
        self.faces = np.ascontiguousarray(self.faces, np.uint32).reshape(-1, 4)
        if self.light_idx is None:
            self.light_idx = np.full(len(self.faces), -1, np.int32)
        self.light_idx = np.ascontiguousarray(self.light_idx, np.int32)

    def vertex_data(self) -> np.ndarray:
        return np.ascontiguousarray(np.concatenate([self.normals, self.tangents, self.uvs], axis=1), np.float32)


@dataclass
class Node:
    parent: int = -1
    mesh: int = -1
    transform: np.ndarray = None  # 4x4 row-major or None (default Transform)


@dataclass
class Light:
    type: int = AREA
    mesh: int = -1
    tri: int = -1
    emission: tuple = (0.0, 0.0, 0.0)
    transform: np.ndarray = None
    two_sided: int = 0
    scene_radius: float = 100.0
    hdr_tex: int = -1


@dataclass
class Scene:
    textures: list = field(default_factory=list)
    materials: list = field(default_factory=list)
    meshes: list = field(default_factory=list)
    nodes: list = field(default_factory=list)
    lights: list = field(default_factory=list)
    # suggested camera / settings (not part of the .ysc file)
    camera: dict = field(default_factory=dict)

    def n_tris(self) -> int:
        return sum(len(m.faces) for m in self.meshes)

    def add_area_lights(self, mesh_idx: int, node_transform=None):
        """One AreaLight per emissive-material triangle of a mesh, light indices appended globally
        (what reference src/gltf/gltf.cpp:299-314 does for a single emissive node)."""
        m = self.meshes[mesh_idx]
        for t, f in enumerate(m.faces):
            em = self.materials[int(f[3])].emission
            if em[0] * em[0] + em[1] * em[1] + em[2] * em[2] > 0:
                m.light_idx[t] = len(self.lights)
                self.lights.append(Light(AREA, mesh_idx, t, tuple(em), node_transform))

    def write(self, path: str):
        with open(path, "wb") as f:
            f.write(b"YSC1")
            f.write(struct.pack("<I", len(self.textures)))
            for t in self.textures:
                d = t.data
                h, w, c = d.shape
                is_float = 1 if d.dtype == np.float32 else 0
                f.write(struct.pack("<5I", c, is_float, t.type, w, h))
                f.write(np.ascontiguousarray(d).tobytes())
            f.write(struct.pack("<I", len(self.materials)))
            for m in self.materials:
                f.write(m.pack())
            f.write(struct.pack("<I", len(self.meshes)))
            for m in self.meshes:
                f.write(struct.pack("<II", len(m.positions), len(m.faces)))
                f.write(m.positions.tobytes())
                f.write(m.vertex_data().tobytes())
                f.write(m.faces.tobytes())
                f.write(m.light_idx.tobytes())
            f.write(struct.pack("<I", len(self.nodes)))
            for n in self.nodes:
                has = 0 if n.transform is None else 1
                mat = IDENTITY if n.transform is None else np.asarray(n.transform, np.float32)
                f.write(struct.pack("<3i", n.parent, n.mesh, has))
                f.write(np.ascontiguousarray(mat, np.float32).tobytes())
            f.write(struct.pack("<I", len(self.lights)))
            for l in self.lights:
                has = 0 if l.transform is None else 1
                mat = IDENTITY if l.transform is None else np.asarray(l.transform, np.float32)
                f.write(struct.pack("<3i3fi", l.type, l.mesh, l.tri, *l.emission, has))
                f.write(np.ascontiguousarray(mat, np.float32).tobytes())
                f.write(struct.pack("<ifi", l.two_sided, l.scene_radius, l.hdr_tex))


# ------------------------------------------------------------------------------------------
# geometry helpers
# ------------------------------------------------------------------------------------------
def translation(x, y, z):
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = (x, y, z)
    return m


def rotation_y(deg):
    a = np.float32(np.deg2rad(deg))
    c, s = np.cos(a), np.sin(a)
    m = np.eye(4, dtype=np.float32)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    return m


def scaling(s):
    m = np.eye(4, dtype=np.float32)
    m[0, 0] = m[1, 1] = m[2, 2] = s
    return m


class MeshBuilder:
    """Accumulates flat-shaded quads/triangles with per-vertex normal, tangent, uv."""

    def __init__(self):
        self.p, self.n, self.t, self.uv, self.f = [], [], [], [], []

    def tri(self, a, b, c, mat, uvs=((0, 0), (1, 0), (0, 1)), normal=None):
        a, b, c = (np.asarray(v, np.float64) for v in (a, b, c))
        nrm = np.cross(b - a, c - a)
        ln = np.linalg.norm(nrm)
        nrm = nrm / ln if ln > 0 else np.array([0.0, 1.0, 0.0])
        if normal is not None:
            nrm = np.asarray(normal, np.float64)
        tg = b - a
        lt = np.linalg.norm(tg)
        tg = tg / lt if lt > 0 else np.array([1.0, 0.0, 0.0])
        i = len(self.p)
        for v, uv in zip((a, b, c), uvs):
            self.p.append(v)
            self.n.append(nrm)
            self.t.append((*tg, 1.0))
            self.uv.append(uv)
        self.f.append((i, i + 1, i + 2, mat))

    def quad(self, a, b, c, d, mat, uv_scale=1.0):
        """a,b,c,d counter-clockwise seen from the front."""
        s = uv_scale
        self.tri(a, b, c, mat, ((0, 0), (s, 0), (s, s)))
        self.tri(a, c, d, mat, ((0, 0), (s, s), (0, s)))

    def box(self, lo, hi, mat, xf=None):
        lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
        c = np.array([[lo[0], lo[1], lo[2]], [hi[0], lo[1], lo[2]], [hi[0], hi[1], lo[2]], [lo[0], hi[1], lo[2]],
                      [lo[0], lo[1], hi[2]], [hi[0], lo[1], hi[2]], [hi[0], hi[1], hi[2]], [lo[0], hi[1], hi[2]]])
        if xf is not None:
            c = (np.asarray(xf, np.float64)[:3, :3] @ c.T).T + np.asarray(xf, np.float64)[:3, 3]
        for q in ((4, 5, 6, 7), (1, 0, 3, 2), (5, 1, 2, 6), (0, 4, 7, 3), (7, 6, 2, 3), (0, 1, 5, 4)):
            self.quad(c[q[0]], c[q[1]], c[q[2]], c[q[3]], mat)

    def build(self) -> Mesh:
        return Mesh(np.array(self.p), np.array(self.n), np.array(self.t), np.array(self.uv),
                    np.array(self.f, np.uint32))


# ------------------------------------------------------------------------------------------
# C1: procedural Cornell box (36 triangles, one quad area light, diffuse + glossy + metal)
# ------------------------------------------------------------------------------------------
def cornell(light_transform=True) -> Scene:
    s = Scene()
    s.materials = [
        Material(base=(0.73, 0.73, 0.73), roughness=1.0),  # 0 white
        Material(base=(0.65, 0.05, 0.05), roughness=1.0),  # 1 red
        Material(base=(0.12, 0.45, 0.15), roughness=1.0),  # 2 green
        Material(base=(0.8, 0.8, 0.8), roughness=0.2),  # 3 glossy dielectric-coated diffuse
        Material(base=(0.9, 0.7, 0.3), roughness=0.3, metallic=1.0),  # 4 rough metal
        Material(base=(1.0, 1.0, 1.0), roughness=1.0, emission=(17.0, 12.0, 4.0)),  # 5 light
    ]
    b = MeshBuilder()
    b.quad((-5, 0, 5), (5, 0, 5), (5, 0, -5), (-5, 0, -5), 0)  # floor (normal +y)
    b.quad((-5, 10, -5), (5, 10, -5), (5, 10, 5), (-5, 10, 5), 0)  # ceiling (normal -y)
    b.quad((-5, 0, -5), (5, 0, -5), (5, 10, -5), (-5, 10, -5), 0)  # back (normal +z)
    b.quad((-5, 0, 5), (-5, 0, -5), (-5, 10, -5), (-5, 10, 5), 1)  # left, red (normal +x)
    b.quad((5, 0, -5), (5, 0, 5), (5, 10, 5), (5, 10, -5), 2)  # right, green (normal -x)
    b.box((-1.5, 0, -1.5), (1.5, 3, 1.5), 3, translation(1.6, 0, 1.5) @ rotation_y(-18))  # short box, glossy
    b.box((-1.5, 0, -1.5), (1.5, 6, 1.5), 4, translation(-1.7, 0, -1.6) @ rotation_y(20))  # tall box, metal
    room = b.build()
    lb = MeshBuilder()
    lb.quad((-1.5, 0, -1.5), (1.5, 0, -1.5), (1.5, 0, 1.5), (-1.5, 0, 1.5), 5)  # faces down (normal -y)
    light = lb.build()
    s.meshes = [room, light]
    lxf = translation(0.0, 9.99, 0.0) if light_transform else None
    if not light_transform:
        light.positions[:, 1] += np.float32(9.99)
    s.nodes = [Node(-1, -1), Node(0, 0), Node(0, 1, lxf)]
    s.add_area_lights(1, lxf)
    s.camera = dict(pos=(0.0, 5.0, 15.0), target=(0.0, 5.0, 0.0), focal=35.0, fnum=0.0, exposure=0.0,
                    w=512, h=512, spp=16)
    return s


# ------------------------------------------------------------------------------------------
# C2: random triangle soup (traversal / intersection microbench), one big quad area light
# ------------------------------------------------------------------------------------------
def soup(n_tris=1_000_000, seed=1234, with_light=True) -> Scene:
    rng = np.random.default_rng(seed)
    s = Scene()
    s.materials = [Material(base=(0.7, 0.7, 0.7), roughness=1.0),
                   Material(base=(1.0, 1.0, 1.0), roughness=1.0, emission=(10.0, 10.0, 10.0))]
    centres = rng.uniform(-10.0, 10.0, (n_tris, 1, 3))
    ext = 5.0 * 2.0 / np.cbrt(n_tris)
    verts = (centres + rng.uniform(-ext, ext, (n_tris, 3, 3))).astype(np.float32)
    e1 = verts[:, 1] - verts[:, 0]
    e2 = verts[:, 2] - verts[:, 0]
    nrm = np.cross(e1, e2)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-20)
    normals = np.repeat(nrm[:, None, :], 3, axis=1).reshape(-1, 3)
    faces = np.zeros((n_tris, 4), np.uint32)
    faces[:, 0] = np.arange(n_tris) * 3
    faces[:, 1] = faces[:, 0] + 1
    faces[:, 2] = faces[:, 0] + 2
    s.meshes = [Mesh(verts.reshape(-1, 3), normals, None, None, faces)]
    s.nodes = [Node(-1, -1), Node(0, 0)]
    if with_light:
        lb = MeshBuilder()
        lb.quad((-30, 25, -30), (30, 25, -30), (30, 25, 30), (-30, 25, 30), 1)  # faces down
        s.meshes.append(lb.build())
        s.nodes.append(Node(0, 1))
        s.add_area_lights(1, None)
    s.camera = dict(pos=(0.0, 0.0, 40.0), target=(0.0, 0.0, 0.0), focal=35.0, fnum=0.0, exposure=0.0,
                    w=1920, h=1080, spp=1, maxdepth=1)
    return s


def degenerate_soup(n_tris=6000, seed=7) -> Scene:
    """Builder stress: clusters of identical triangles (zero centroid extent in nodes of more than a leaf's size),
    centroids on a coarse grid (many elements exactly on split planes), zero-area triangles, extents from 1e-4 to 1e3,
    signed zeros in the vertex data."""
    rng = np.random.default_rng(seed)
    s = Scene()
    s.materials = [Material(base=(0.7, 0.7, 0.7), roughness=1.0),
                   Material(base=(1.0, 1.0, 1.0), roughness=1.0, emission=(10.0, 10.0, 10.0))]
    kinds = rng.integers(0, 5, n_tris)
    centres = np.round(rng.uniform(-8.0, 8.0, (n_tris, 1, 3)) * 2.0) / 2.0  # half-unit grid
    ext = np.where(kinds == 3, 1e-4, np.where(kinds == 4, 50.0, 0.3))[:, None, None]
    verts = centres + rng.uniform(-1.0, 1.0, (n_tris, 3, 3)) * ext
    dup = np.nonzero(kinds == 1)[0]
    if len(dup) > 1:  # runs of copies of one triangle: 120 copies at most, so some runs exceed kSmallSpan and kMaxLeaf
        for start in range(0, len(dup), 120):
            verts[dup[start:start + 120]] = verts[dup[start]]
    flat = np.nonzero(kinds == 2)[0]
    verts[flat, 2] = verts[flat, 1]  # zero-area triangles
    verts[rng.integers(0, n_tris, n_tris // 50), rng.integers(0, 3, n_tris // 50), rng.integers(0, 3, n_tris // 50)] = -0.0
    verts = verts.astype(np.float32)
    nrm = np.zeros((n_tris * 3, 3), np.float32)
    nrm[:, 1] = 1.0
    faces = np.zeros((n_tris, 4), np.uint32)
    faces[:, 0] = np.arange(n_tris) * 3
    faces[:, 1] = faces[:, 0] + 1
    faces[:, 2] = faces[:, 0] + 2
    s.meshes = [Mesh(verts.reshape(-1, 3), nrm, None, None, faces)]
    s.nodes = [Node(-1, -1), Node(0, 0)]
    lb = MeshBuilder()
    lb.quad((-30, 25, -30), (30, 25, -30), (30, 25, 30), (-30, 25, 30), 1)
    s.meshes.append(lb.build())
    s.nodes.append(Node(0, 1))
    s.add_area_lights(1, None)
    s.camera = dict(pos=(0.0, 0.0, 40.0), target=(0.0, 0.0, 0.0), focal=35.0, fnum=0.0, exposure=0.0, w=64, h=64, spp=1, maxdepth=1)
    return s


# ------------------------------------------------------------------------------------------
# tiny two-quad scene (smoke tests / SURVEY Appendix C style)
# ------------------------------------------------------------------------------------------
def two_quads() -> Scene:
    s = Scene()
    s.materials = [Material(base=(0.8, 0.8, 0.8), roughness=1.0),
                   Material(base=(0.8, 0.3, 0.2), roughness=0.4),
                   Material(base=(1, 1, 1), roughness=1.0, emission=(8.0, 8.0, 8.0))]
    b = MeshBuilder()
    b.quad((-6, 0, 6), (6, 0, 6), (6, 0, -6), (-6, 0, -6), 0)
    b.quad((-4, 0, -3), (4, 0, -3), (4, 8, -3), (-4, 8, -3), 1)
    lb = MeshBuilder()
    lb.quad((-2, 9, -2), (2, 9, -2), (2, 9, 2), (-2, 9, 2), 2)
    s.meshes = [b.build(), lb.build()]
    s.nodes = [Node(-1, -1), Node(0, 0), Node(0, 1)]
    s.add_area_lights(1, None)
    s.camera = dict(pos=(0.0, 5.0, 15.0), target=(0.0, 5.0, 0.0), focal=35.0, fnum=0.0, exposure=0.0,
                    w=64, h=64, spp=16)
    return s


# ------------------------------------------------------------------------------------------
# procedural textures
# ------------------------------------------------------------------------------------------
def _noise_tex(rng, w, h, c, lo=0, hi=255):
    """Smooth-ish random 8-bit texture (low-res noise upsampled + detail)."""
    cw, ch = max(2, w // 8), max(2, h // 8)
    coarse = rng.uniform(lo, hi, (ch, cw, c))
    ys = (np.arange(h) * ch // h)[:, None]
    xs = (np.arange(w) * cw // w)[None, :]
    img = coarse[ys, xs] + rng.uniform(-12, 12, (h, w, c))
    return np.clip(img, 0, 255).astype(np.uint8)


def sky_hdr(w=64, h=64, seed=5, sun=4000.0) -> np.ndarray:
    """Octahedral-layout HDR sky: vertical gradient + a small hot sun patch (float32, (h,w,3))."""
    rng = np.random.default_rng(seed)
    v = np.linspace(0, 1, h, dtype=np.float32)[:, None, None]
    img = (0.2 + 0.8 * v) * np.array([0.5, 0.7, 1.0], np.float32)[None, None, :]
    img = img * (1.0 + 0.1 * rng.uniform(-1, 1, (h, w, 1)).astype(np.float32))
    cy, cx, r = int(h * 0.62), int(w * 0.33), max(1, w // 32)
    img[cy - r:cy + r, cx - r:cx + r] = np.array([sun, sun * 0.9, sun * 0.7], np.float32)
    return np.ascontiguousarray(img, np.float32)


# ------------------------------------------------------------------------------------------
# material zoo: every ParametricBSDF feature + every light type (function-level KATs, small renders)
# ------------------------------------------------------------------------------------------
def material_zoo(env=True) -> Scene:
    rng = np.random.default_rng(11)
    s = Scene()
    base_rgba = _noise_tex(rng, 32, 16, 4, 40, 255)
    base_rgba[::3, ::2, 3] = 90  # alpha cut-outs
    base_opaque = _noise_tex(rng, 16, 16, 4, 60, 255)
    base_opaque[..., 3] = 255
    s.textures = [
        Texture(base_rgba, SRGB),  # 0 base + alpha
        Texture(base_opaque, SRGB),  # 1 base, opaque
        Texture(_noise_tex(rng, 16, 32, 2, 20, 255), NONCOLOR),  # 2 roughness/metallic
        Texture(_noise_tex(rng, 16, 16, 1, 0, 255), NONCOLOR),  # 3 transmission
        Texture(np.clip(_noise_tex(rng, 32, 32, 3, 90, 165).astype(np.int32) + np.array([0, 0, 80]), 0, 255).astype(np.uint8), NONCOLOR),  # 4 normal
        Texture(_noise_tex(rng, 8, 8, 1, 100, 255), NONCOLOR),  # 5 clearcoat
        Texture(_noise_tex(rng, 8, 16, 3, 0, 255), SRGB),  # 6 emission
        Texture(sky_hdr(32, 32, 5, 300.0), LINEAR),  # 7 env
    ]
    s.materials = [
        Material(base=(0.8, 0.7, 0.6), roughness=1.0),  # 0 diffuse
        Material(base=(0.8, 0.6, 0.4), metallic=0.3, roughness=0.4, clearcoat=0.5, clearcoat_roughness=0.1),  # 1 SURVEY App. C
        Material(base=(0.95, 0.8, 0.4), metallic=1.0, roughness=0.25),  # 2 metal
        Material(base=(0.9, 0.9, 0.9), metallic=1.0, roughness=0.01),  # 3 mirror (smooth → specular)
        Material(base=(0.9, 0.95, 1.0), transmission=1.0, roughness=0.0, ior=1.5, thin=0,
                 volume_color=(0.6, 0.8, 0.9), volume_density=0.7),  # 4 solid smooth glass + volume
        Material(base=(0.9, 0.95, 1.0), transmission=1.0, roughness=0.35, ior=1.33, thin=0),  # 5 rough solid glass
        Material(base=(1.0, 0.9, 0.8), transmission=0.8, roughness=0.2, ior=1.5, thin=1),  # 6 thin rough glass
        Material(base=(1.0, 1.0, 1.0), transmission=1.0, roughness=0.0, ior=1.5, thin=1),  # 7 thin smooth (transparent to NEE)
        Material(base=(0.7, 0.2, 0.2), metallic=0.8, roughness=0.3, anisotropic=0.8, aniso_rotation=0.6),  # 8 aniso
        Material(base=(1, 1, 1), base_tex=0, roughness=0.6),  # 9 alpha-tested textured
        Material(base=(1, 1, 1), base_tex=1, mr_tex=2, normal_tex=4, roughness=1.0, metallic=1.0),  # 10 full PBR textured
        Material(base=(0.6, 0.6, 0.9), trans_tex=3, transmission=1.0, roughness=0.15, thin=1),  # 11 transmission tex
        Material(base=(0.1, 0.3, 0.8), roughness=0.5, clearcoat=1.0, clearcoat_roughness=0.03, cc_tex=5),  # 12 car paint
        Material(base=(1, 1, 1), roughness=1.0, emission=(6.0, 5.0, 3.0), emis_tex=6),  # 13 textured emitter
        Material(base=(1, 1, 1), roughness=1.0, emission=(12.0, 12.0, 12.0)),  # 14 emitter
        Material(base=(0.5, 0.5, 0.5), roughness=0.05, clearcoat=0.7, clearcoat_roughness=0.0),  # 15 smooth coat
        Material(base=(0.3, 0.8, 0.3), roughness=0.0),  # 16 smooth glossy (specular branch)
    ]
    # geometry: a floor, a back wall split in material patches, a few boxes, two emissive quads
    b = MeshBuilder()
    b.quad((-8, 0, 8), (8, 0, 8), (8, 0, -8), (-8, 0, -8), 10, uv_scale=3.0)
    mats = [0, 1, 2, 3, 8, 9, 12, 15, 16, 11, 6, 7]
    for i, m in enumerate(mats):
        x0 = -8 + i * (16 / len(mats))
        x1 = x0 + 16 / len(mats)
        b.quad((x0, 0, -6), (x1, 0, -6), (x1, 7, -6), (x0, 7, -6), m, uv_scale=2.0)
    b.box((-1.2, 0, -1.2), (1.2, 2.4, 1.2), 4, translation(-3.5, 0.01, 0.5) @ rotation_y(25))
    b.box((-1.0, 0, -1.0), (1.0, 2.0, 1.0), 5, translation(0.0, 0.01, 1.5) @ rotation_y(-15))
    b.box((-1.0, 0, -1.0), (1.0, 3.0, 1.0), 9, translation(3.5, 0.01, 0.0) @ rotation_y(40))
    b.quad((-2.5, 0.5, 3.0), (2.5, 0.5, 3.0), (2.5, 3.0, 3.0), (-2.5, 3.0, 3.0), 7)  # thin pane in front
    room = b.build()
    lb = MeshBuilder()
    lb.quad((-2, 0, -1), (2, 0, -1), (2, 0, 1), (-2, 0, 1), 14)
    lb.quad((-1, -0.5, 2), (1, -0.5, 2), (1, -0.5, 3), (-1, -0.5, 3), 13, uv_scale=1.0)
    lights = lb.build()
    s.meshes = [room, lights]
    lxf = translation(0.5, 8.0, -1.0) @ rotation_y(30)
    inner = translation(0.2, 0.0, -0.3) @ scaling(1.1)
    # two-level graph: root → group(transform) → room ; root → lights(transform)
    s.nodes = [Node(-1, -1), Node(0, -1, inner), Node(1, 0, rotation_y(5)), Node(0, 1, lxf)]
    s.add_area_lights(1, lxf)
    if env:
        s.lights.append(Light(IMAGE_INF, hdr_tex=7, scene_radius=60.0, transform=rotation_y(40)))
    s.lights.append(Light(UNIFORM_INF, emission=(0.05, 0.05, 0.08), scene_radius=60.0))
    s.camera = dict(pos=(0.5, 4.0, 13.0), target=(0.0, 2.5, 0.0), focal=35.0, fnum=4.0, exposure=1.0,
                    w=96, h=64, spp=16)
    return s


# ------------------------------------------------------------------------------------------
# parametric surfaces (vectorised) for the large synthetic scenes
# ------------------------------------------------------------------------------------------
def surface_mesh(fn, nu: int, nv: int, mat: int, uv_scale=(1.0, 1.0), flip=False):
    """Tessellates p = fn(u, v), u,v in [0,1], into nu x nv quads (2 triangles each).
    Returns (positions, normals, tangents(4), uvs, faces) with smooth normals from finite differences."""
    u = np.linspace(0.0, 1.0, nu + 1)
    v = np.linspace(0.0, 1.0, nv + 1)
    U, V = np.meshgrid(u, v, indexing="xy")  # (nv+1, nu+1)
    P = fn(U, V)  # (nv+1, nu+1, 3)
    e = 1e-4
    dU = (fn(np.clip(U + e, 0, 1), V) - fn(np.clip(U - e, 0, 1), V))
    dV = (fn(U, np.clip(V + e, 0, 1)) - fn(U, np.clip(V - e, 0, 1)))
    N = np.cross(dU, dV)
    ln = np.linalg.norm(N, axis=-1, keepdims=True)
    N = np.where(ln > 1e-20, N / np.maximum(ln, 1e-20), np.array([0.0, 1.0, 0.0]))
    T = dU / np.maximum(np.linalg.norm(dU, axis=-1, keepdims=True), 1e-20)
    if flip:
        N = -N
    idx = (np.arange(nv + 1)[:, None] * (nu + 1) + np.arange(nu + 1)[None, :])
    a, b, c, d = idx[:-1, :-1], idx[:-1, 1:], idx[1:, 1:], idx[1:, :-1]
    if flip:
        f1 = np.stack([a, c, b], -1)
        f2 = np.stack([a, d, c], -1)
    else:
        f1 = np.stack([a, b, c], -1)
        f2 = np.stack([a, c, d], -1)
    faces = np.concatenate([f1.reshape(-1, 3), f2.reshape(-1, 3)], 0)
    faces = np.concatenate([faces, np.full((len(faces), 1), mat)], 1).astype(np.uint32)
    uv = np.stack([U * uv_scale[0], V * uv_scale[1]], -1)
    tg = np.concatenate([T, np.ones(T.shape[:-1] + (1,))], -1)
    return (P.reshape(-1, 3).astype(np.float32), N.reshape(-1, 3).astype(np.float32),
            tg.reshape(-1, 4).astype(np.float32), uv.reshape(-1, 2).astype(np.float32), faces)


class BigMesh:
    """Concatenates surface patches into one Mesh (one BVH, like a merged glTF mesh, gltf.cpp:178-270)."""

    def __init__(self):
        self.parts = []
        self.nv = 0

    def add(self, part):
        p, n, t, uv, f = part
        f = f.copy()
        f[:, :3] += self.nv
        self.nv += len(p)
        self.parts.append((p, n, t, uv, f))

    def n_tris(self):
        return sum(len(x[4]) for x in self.parts)

    def build(self) -> Mesh:
        cat = lambda k: np.concatenate([x[k] for x in self.parts], 0)
        return Mesh(cat(0), cat(1), cat(2), cat(3), cat(4))


def _plane(origin, du, dv):
    o, du, dv = (np.asarray(x, np.float64) for x in (origin, du, dv))
    return lambda U, V: o + U[..., None] * du + V[..., None] * dv


def _cylinder(cx, cz, r, y0, y1, flute=0.0, nfl=16):
    def fn(U, V):
        ang = 2 * np.pi * U
        rr = r * (1.0 + flute * np.cos(nfl * ang)) * (1.0 + 0.08 * np.cos(np.pi * V) ** 8)
        return np.stack([cx + rr * np.cos(ang), y0 + (y1 - y0) * V, cz - rr * np.sin(ang)], -1)
    return fn


def _arch(x0, x1, z, y, r_tube):
    """Half torus spanning x0..x1 at height y (an arch between two columns)."""
    R = 0.5 * (x1 - x0)
    cx = 0.5 * (x0 + x1)

    def fn(U, V):
        a = np.pi * U  # along the arch
        b = 2 * np.pi * V  # around the tube
        rr = R + r_tube * np.cos(b)
        return np.stack([cx - rr * np.cos(a), y + rr * np.sin(a), z + r_tube * np.sin(b)], -1)
    return fn


def _pbr_material_set(rng, tex_res, n_sets, textures, alpha_every=0):
    """n_sets textured PBR materials (base sRGB RGBA, MR, normal); returns material list."""
    mats = []
    for k in range(n_sets):
        base = _noise_tex(rng, tex_res, tex_res, 4, 50, 240)
        base[..., :3] = (base[..., :3] * rng.uniform(0.6, 1.0, 3)).astype(np.uint8)
        base[..., 3] = 255
        if alpha_every and k % alpha_every == alpha_every - 1:
            yy, xx = np.mgrid[0:tex_res, 0:tex_res]
            holes = ((xx // max(1, tex_res // 8) + yy // max(1, tex_res // 8)) % 3 == 0)
            base[holes, 3] = 0
        mr = _noise_tex(rng, tex_res, tex_res, 2, 60, 255)
        mr[..., 1] = (mr[..., 1] * 0.15).astype(np.uint8)  # mostly dielectric
        nrm = _noise_tex(rng, tex_res, tex_res, 3, 100, 155)
        nrm[..., 2] = 235
        i0 = len(textures)
        textures += [Texture(base, SRGB), Texture(mr, NONCOLOR), Texture(nrm, NONCOLOR)]
        mats.append(Material(base=(1, 1, 1), base_tex=i0, mr_tex=i0 + 1, normal_tex=i0 + 2, roughness=1.0, metallic=1.0))
    return mats


# ------------------------------------------------------------------------------------------
# C3: Sponza-shaped atrium, textured PBR + normal maps, lit only by an HDR environment map
# ------------------------------------------------------------------------------------------
def sponza(n_tris=260_000, tex_res=1024, env_res=2048, n_materials=24, seed=21) -> Scene:
    rng = np.random.default_rng(seed)
    s = Scene()
    s.materials = _pbr_material_set(rng, tex_res, n_materials, s.textures, alpha_every=8)
    # tessellation budget: d scales every patch; triangles grow ~ d^2
    d = max(1, int(round(np.sqrt(n_tris / 1716.0))))
    bm = BigMesh()
    L, Wd, Hh = 30.0, 12.0, 14.0  # atrium length (x), width (z), height
    m = lambda k: k % n_materials
    bm.add(surface_mesh(lambda U, V: _plane((-L / 2, 0, Wd / 2), (L, 0, 0), (0, 0, -Wd))(U, V) +
                        np.stack([0 * U, 0.02 * np.sin(40 * U) * np.sin(30 * V), 0 * U], -1), 12 * d, 6 * d, m(0), (8, 4)))
    bm.add(surface_mesh(_plane((-L / 2, 0, -Wd / 2), (L, 0, 0), (0, Hh, 0)), 10 * d, 5 * d, m(1), (6, 3)))  # back wall
    bm.add(surface_mesh(_plane((L / 2, 0, Wd / 2), (-L, 0, 0), (0, Hh, 0)), 10 * d, 5 * d, m(2), (6, 3)))  # front wall
    bm.add(surface_mesh(_plane((-L / 2, 0, Wd / 2), (0, 0, -Wd), (0, Hh, 0)), 5 * d, 5 * d, m(3), (3, 3)))  # left
    bm.add(surface_mesh(_plane((L / 2, 0, -Wd / 2), (0, 0, Wd), (0, Hh, 0)), 5 * d, 5 * d, m(4), (3, 3)))  # right
    ncol = 8
    xs = np.linspace(-L / 2 + 3, L / 2 - 3, ncol)
    for row, z in enumerate((-Wd / 2 + 2.5, Wd / 2 - 2.5)):
        for i, x in enumerate(xs):
            bm.add(surface_mesh(_cylinder(x, z, 0.45, 0.0, 6.0, 0.04), 6 * d, 3 * d, m(5 + (i + row) % 6), (2, 4)))
            bm.add(surface_mesh(_cylinder(x, z, 0.38, 7.0, 11.5, 0.0), 4 * d, 2 * d, m(11 + (i + row) % 4), (2, 3)))
            if i + 1 < ncol:
                bm.add(surface_mesh(_arch(x, xs[i + 1], z, 6.0, 0.3), 5 * d, 2 * d, m(15 + i % 3), (3, 1)))
        # gallery floor slab along the row
        zz = z + (1.2 if row == 0 else -1.2)
        bm.add(surface_mesh(_plane((-L / 2, 7.0, zz - 1.4), (L, 0, 0), (0, 0, 2.8)), 8 * d, 2 * d, m(18 + row), (8, 1),
                            flip=True))
    # hanging drapes with alpha cut-outs (material k % 8 == 7 carries the holes)
    for i, x in enumerate(xs[1::2]):
        def drape(U, V, x=x, i=i):
            return np.stack([x + 0.15 * np.sin(6 * np.pi * V + i), 10.5 - 5.0 * V, -1.5 + 3.0 * U + 0 * V], -1)
        bm.add(surface_mesh(drape, 3 * d, 4 * d, 7 if n_materials > 7 else 0, (2, 3)))
    s.meshes = [bm.build()]
    s.nodes = [Node(-1, -1), Node(0, 0)]
    s.textures.append(Texture(sky_hdr(env_res, env_res, 5, 6000.0), LINEAR))
    s.lights = [Light(IMAGE_INF, hdr_tex=len(s.textures) - 1, scene_radius=100.0)]
    # src/main.cpp:32-34, 69-72 style: f/4, exposure set for the env brightness
    s.camera = dict(pos=(-12.0, 2.0, 0.5), target=(8.0, 5.0, -0.5), focal=24.0, fnum=4.0, exposure=0.0,
                    w=1920, h=1080, spp=1024)
    return s


# ------------------------------------------------------------------------------------------
# C4: McLaren-shaped scene: metallic paint + clearcoat, chrome, thin and solid glass (+ volume),
#     emissive headlights, ground, HDR environment
# ------------------------------------------------------------------------------------------
def mclaren(n_tris=2_000_000, env_res=2048, seed=33) -> Scene:
    rng = np.random.default_rng(seed)
    s = Scene()
    s.materials = [
        Material(base=(0.85, 0.25, 0.02), metallic=0.9, roughness=0.35, clearcoat=1.0, clearcoat_roughness=0.03),  # 0 paint
        Material(base=(0.95, 0.95, 0.95), metallic=1.0, roughness=0.05),  # 1 chrome
        Material(base=(0.03, 0.03, 0.03), roughness=0.8),  # 2 tyre
        Material(base=(0.9, 0.95, 1.0), transmission=1.0, roughness=0.0, ior=1.5, thin=1),  # 3 window (thin glass)
        Material(base=(1.0, 1.0, 1.0), transmission=1.0, roughness=0.05, ior=1.5, thin=0,
                 volume_color=(0.9, 0.6, 0.3), volume_density=1.5),  # 4 solid lens with Beer-Lambert volume
        Material(base=(1, 1, 1), roughness=1.0, emission=(40.0, 38.0, 30.0)),  # 5 headlight emitter
        Material(base=(0.35, 0.35, 0.36), roughness=0.6),  # 6 asphalt
        Material(base=(0.05, 0.05, 0.06), metallic=1.0, roughness=0.3, anisotropic=0.7, aniso_rotation=0.3),  # 7 brushed trim
    ]
    d = max(1, int(round(np.sqrt(n_tris / 2905.0))))
    bm = BigMesh()

    def body(U, V):
        th, ph = np.pi * V, 2 * np.pi * U
        bump = 1.0 + 0.03 * np.sin(9 * ph) * np.sin(7 * th) + 0.15 * np.exp(-((V - 0.35) / 0.12) ** 2) * (np.cos(ph) > 0)
        x = 2.3 * np.sin(th) * np.cos(ph) * bump
        y = 0.62 + 0.55 * np.cos(th) * bump
        z = 1.0 * np.sin(th) * np.sin(ph) * bump
        return np.stack([x, np.maximum(y, 0.18), z], -1)
    bm.add(surface_mesh(body, 40 * d, 20 * d, 0))

    def canopy(U, V):
        th, ph = 0.5 * np.pi * V, 2 * np.pi * U
        return np.stack([-0.2 + 0.9 * np.sin(th) * np.cos(ph), 1.0 + 0.42 * np.cos(th), 0.62 * np.sin(th) * np.sin(ph)], -1)
    bm.add(surface_mesh(canopy, 16 * d, 8 * d, 3))

    for sx in (-1.45, 1.45):
        for sz in (-1.02, 1.02):
            def tyre(U, V, sx=sx, sz=sz):
                a, b = 2 * np.pi * U, 2 * np.pi * V
                rr = 0.36 + 0.12 * np.cos(b)
                return np.stack([sx + rr * np.cos(a), 0.48 + rr * np.sin(a), sz + 0.13 * np.sin(b)], -1)
            bm.add(surface_mesh(tyre, 12 * d, 6 * d, 2))

            def rim(U, V, sx=sx, sz=sz):
                a = 2 * np.pi * U
                rr = 0.26 * V * (1.0 + 0.12 * np.cos(5 * a))
                return np.stack([sx + rr * np.cos(a), 0.48 + rr * np.sin(a), sz + np.sign(sz) * (0.14 - 0.05 * V) + 0 * a], -1)
            bm.add(surface_mesh(rim, 10 * d, 3 * d, 1, flip=sz < 0))
    for sz in (-0.55, 0.55):
        def lens(U, V, sz=sz):
            th, ph = np.pi * V, 2 * np.pi * U
            return np.stack([2.05 + 0.16 * np.sin(th) * np.cos(ph), 0.62 + 0.1 * np.cos(th), sz + 0.2 * np.sin(th) * np.sin(ph)], -1)
        bm.add(surface_mesh(lens, 8 * d, 4 * d, 4))
    bm.add(surface_mesh(_plane((-2.2, 0.75, -0.9), (0.0, 0, 1.8), (0.5, 0.05, 0.0)), 4 * d, 1 * d, 7))  # rear trim
    bm.add(surface_mesh(_plane((-12, 0, 8), (24, 0, 0), (0, 0, -16)), 8 * d, 6 * d, 6, (12, 8)))  # ground
    car = bm.build()
    lb = MeshBuilder()
    for sz in (-0.55, 0.55):  # emitters inside the lenses, facing +x
        lb.quad((2.0, 0.56, sz - 0.08), (2.0, 0.56, sz + 0.08), (2.0, 0.68, sz + 0.08), (2.0, 0.68, sz - 0.08), 5)
    lamps = lb.build()
    s.meshes = [car, lamps]
    s.nodes = [Node(-1, -1), Node(0, 0), Node(0, 1)]
    s.add_area_lights(1, None)
    s.textures.append(Texture(sky_hdr(env_res, env_res, 9, 3000.0), LINEAR))
    s.lights.append(Light(IMAGE_INF, hdr_tex=len(s.textures) - 1, scene_radius=100.0))
    s.camera = dict(pos=(5.5, 1.6, 4.2), target=(0.0, 0.6, 0.0), focal=35.0, fnum=4.0, exposure=0.0,
                    w=1920, h=1080, spp=1024)
    return s


# ------------------------------------------------------------------------------------------
# randomised small scenes (parity fuzzing): every material feature, nested transforms, all light types
# ------------------------------------------------------------------------------------------
def random_scene(seed: int) -> Scene:
    rng = np.random.default_rng(1000 + seed)
    s = Scene()
    n_tex = 6
    for k in range(n_tex):
        w, h = int(rng.integers(2, 24)), int(rng.integers(2, 24))
        s.textures.append(Texture(_noise_tex(rng, w, h, 4, 0, 255), SRGB))  # 4k+0 base (with alpha)
        s.textures[-1].data[rng.uniform(size=(h, w)) < 0.3, 3] = int(rng.integers(0, 255))
        s.textures.append(Texture(_noise_tex(rng, w, h, 2, 0, 255), NONCOLOR))  # mr
        s.textures.append(Texture(_noise_tex(rng, w, h, 1, 0, 255), NONCOLOR))  # mono (transmission / clearcoat)
        s.textures.append(Texture(_noise_tex(rng, w, h, 3, 60, 200), NONCOLOR))  # normal / emission rgb
    def pick(kind, p=0.4):
        return int(rng.integers(0, n_tex)) * 4 + kind if rng.uniform() < p else -1
    n_mat = int(rng.integers(3, 9))
    for k in range(n_mat):
        emissive = rng.uniform() < 0.25
        s.materials.append(Material(
            base=tuple(rng.uniform(0.05, 1.0, 3)), base_tex=pick(0), mr_tex=pick(1), trans_tex=pick(2, 0.2), normal_tex=pick(3, 0.3),
            cc_tex=pick(2, 0.2), emis_tex=pick(3, 0.5) if emissive else -1,
            metallic=float(rng.choice([0.0, 1.0, rng.uniform()])), roughness=float(rng.choice([0.0, 1.0, rng.uniform(0.02, 1.0)])),
            transmission=float(rng.choice([0.0, 0.0, 1.0, rng.uniform()])), ior=float(rng.uniform(1.05, 2.2)),
            anisotropic=float(rng.choice([0.0, rng.uniform()])), aniso_rotation=float(rng.uniform(0, 3.0)),
            clearcoat=float(rng.choice([0.0, 0.0, 1.0, rng.uniform()])), clearcoat_roughness=float(rng.choice([0.0, rng.uniform(0, 0.5)])),
            emission=tuple(rng.uniform(0.5, 20.0, 3)) if emissive else (0.0, 0.0, 0.0), thin=int(rng.integers(0, 2)),
            volume_color=tuple(rng.uniform(0.2, 1.0, 3)), volume_density=float(rng.uniform(0, 2.0))))
    if not any(m.emission[0] > 0 for m in s.materials):
        s.materials[-1].emission = (9.0, 8.0, 7.0)
    n_mesh = int(rng.integers(2, 5))
    for mi in range(n_mesh):
        b = MeshBuilder()
        for _ in range(int(rng.integers(1, 5))):
            c = rng.uniform(-3, 3, 3)
            e = rng.uniform(0.3, 2.5, 3)
            mat = int(rng.integers(0, n_mat))
            if rng.uniform() < 0.5:
                b.box(c - e / 2, c + e / 2, mat, rotation_y(float(rng.uniform(0, 90))))
            else:
                b.quad(c + (-e[0], 0, e[2]), c + (e[0], 0, e[2]), c + (e[0], e[1] * 0.2, -e[2]), c + (-e[0], 0, -e[2]), mat,
                       uv_scale=float(rng.uniform(0.5, 3.0)))
        if rng.uniform() < 0.3:  # a degenerate (zero-area) triangle and a sliver
            b.tri((0, 0, 0), (1, 1, 1), (2, 2, 2), int(rng.integers(0, n_mat)))
            b.tri((0, 0, 0), (1, 0, 0), (2, 1e-7, 0), int(rng.integers(0, n_mat)))
        s.meshes.append(b.build())
    b = MeshBuilder()  # a big floor so most paths keep bouncing
    b.quad((-9, -3.2, 9), (9, -3.2, 9), (9, -3.2, -9), (-9, -3.2, -9), 0, uv_scale=4.0)
    s.meshes.append(b.build())

    def xf():
        m = translation(*rng.uniform(-1.5, 1.5, 3)) @ rotation_y(float(rng.uniform(0, 360)))
        if rng.uniform() < 0.5:
            m = m @ scaling(float(rng.uniform(0.6, 1.6)))
        if rng.uniform() < 0.3:  # non-uniform scale + shear-free x rotation
            sc = np.eye(4, dtype=np.float32)
            sc[0, 0], sc[1, 1] = rng.uniform(0.5, 1.5), rng.uniform(0.5, 1.5)
            m = m @ sc
        return m.astype(np.float32)
    s.nodes = [Node(-1, -1, xf() if rng.uniform() < 0.3 else None)]
    emissive_mesh_nodes = {}
    for mi in range(len(s.meshes)):
        parent = int(rng.integers(0, len(s.nodes)))
        if rng.uniform() < 0.4:  # an intermediate group node
            s.nodes.append(Node(parent, -1, xf()))
            parent = len(s.nodes) - 1
        t = xf() if rng.uniform() < 0.7 else None
        s.nodes.append(Node(parent, mi, t))
        emissive_mesh_nodes[mi] = len(s.nodes) - 1
    # area lights: only for meshes whose node chain is a single transform under an identity root, the subset in
    # which the reference's light placement is well defined; others stay non-light emitters
    for mi, ni in emissive_mesh_nodes.items():
        n = s.nodes[ni]
        if n.parent == 0 and s.nodes[0].transform is None:
            s.add_area_lights(mi, n.transform)
    if rng.uniform() < 0.6:
        s.textures.append(Texture(sky_hdr(int(rng.integers(4, 40)), int(rng.integers(4, 40)), seed, float(rng.uniform(5, 500))), LINEAR))
        s.lights.append(Light(IMAGE_INF, hdr_tex=len(s.textures) - 1, scene_radius=float(rng.uniform(20, 200)),
                              transform=rotation_y(float(rng.uniform(0, 360))) if rng.uniform() < 0.5 else None))
    if rng.uniform() < 0.4 or not s.lights:
        s.lights.append(Light(UNIFORM_INF, emission=tuple(rng.uniform(0.05, 1.0, 3)), scene_radius=50.0))
    s.camera = dict(pos=tuple(rng.uniform(-1, 1, 3) + (0, 1.5, 11)), target=tuple(rng.uniform(-1, 1, 3)), focal=float(rng.uniform(20, 50)),
                    fnum=float(rng.choice([0.0, 1.4, 4.0])), exposure=float(rng.uniform(-1, 1)), sides=int(rng.choice([0, 0, 5, 8])),
                    w=40, h=24, spp=int(rng.choice([3, 4, 5, 8, 12, 16])))
    return s
