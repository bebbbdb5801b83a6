"""Seeded synthetic stand-ins for the reference's missing asset pack.

`scenes/media/` is git-ignored in the reference (.gitignore:7) and absent from the tree, and there is
no network: `media/globe.png`, `media/crytek-sponza/sponza.obj` and `media/hairball/hairball.obj`
cannot be obtained.  Every number reported on "sponza" / "hairball" / "globe" is therefore measured on
the stand-ins generated here; generator, seed, triangle count and bounds are part of the
result's label.

All generators are pure numpy, deterministic for a given seed, and return data in the form the
reference's loaders would hand on (`ObjData` before the /4 of loader3d.rs:669; textures as decoded
uint8 images before the y-flip of texture2d.rs:99-107).
"""
import numpy as np

from .loader3d import MtlMaterial, ObjData


# ---------------------------------------------------------------------------------------------
# media/globe.png stand-in (scenes/basic_materials.mtl:20)
# ---------------------------------------------------------------------------------------------
def globe_texture(width=2048, height=1024, seed=0):
    """Lat-long checker + smooth gradient + low-frequency 'continents', RGB uint8 (H,W,3)."""
    rng = np.random.default_rng(seed)
    u = (np.arange(width, dtype=np.float64) + 0.5) / width
    v = (np.arange(height, dtype=np.float64) + 0.5) / height
    U, V = np.meshgrid(u, v)
    checker = ((np.floor(U * 24) + np.floor(V * 12)) % 2).astype(np.float64)
    land = np.zeros_like(U)
    for _ in range(12):
        fu, fv = rng.integers(1, 7), rng.integers(1, 5)
        ph1, ph2 = rng.uniform(0, 2 * np.pi, 2)
        land += rng.uniform(0.3, 1.0) * np.sin(2 * np.pi * fu * U + ph1) * np.sin(np.pi * fv * V + ph2)
    land = (land > 0.35).astype(np.float64)
    r = 0.15 + 0.55 * land + 0.2 * checker * (1 - land)
    g = 0.25 + 0.45 * land * (1 - V) + 0.25 * U * (1 - land)
    b = 0.75 * (1 - land) + 0.15 * checker + 0.1 * V
    img = np.stack([r, g, b], axis=-1)
    return (np.clip(img, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)


def _pattern_texture(size, seed, kind):
    """Small procedural diffuse textures for the atrium materials (RGB uint8)."""
    rng = np.random.default_rng(seed)
    u = (np.arange(size) + 0.5) / size
    U, V = np.meshgrid(u, u)
    base = rng.uniform(0.35, 0.9, 3)
    if kind == "brick":
        row = np.floor(V * 16)
        mortar = ((V * 16) % 1 < 0.08) | (((U * 8 + 0.5 * (row % 2)) % 1) < 0.04)
        shade = 0.8 + 0.2 * np.sin(37.0 * row + 11.0 * np.floor(U * 8 + 0.5 * (row % 2)))
        img = base[None, None, :] * shade[:, :, None]
        img[mortar] = 0.82
    elif kind == "tile":
        edge = ((U * 8) % 1 < 0.05) | ((V * 8) % 1 < 0.05)
        alt = ((np.floor(U * 8) + np.floor(V * 8)) % 2)[:, :, None]
        img = base[None, None, :] * (0.7 + 0.3 * alt)
        img[edge] = 0.3
    else:  # cloth
        weave = 0.75 + 0.25 * np.sin(2 * np.pi * 64 * U) * np.sin(2 * np.pi * 64 * V)
        stripes = 0.7 + 0.3 * (np.floor(U * 6) % 2)
        img = base[None, None, :] * (weave * stripes)[:, :, None]
    return (np.clip(img, 0, 1) * 255.0 + 0.5).astype(np.uint8)


def _leaf_mask(size, seed):
    """Opacity map (single channel uint8): random discs = opaque leaves on a transparent ground."""
    rng = np.random.default_rng(seed)
    u = (np.arange(size) + 0.5) / size
    U, V = np.meshgrid(u, u)
    m = np.zeros((size, size), dtype=bool)
    for _ in range(60):
        cx, cy, r = rng.uniform(0, 1), rng.uniform(0, 1), rng.uniform(0.04, 0.11)
        du = np.minimum(np.abs(U - cx), 1 - np.abs(U - cx))
        dv = np.minimum(np.abs(V - cy), 1 - np.abs(V - cy))
        m |= (du * du + dv * dv) < r * r
    return (m * 255).astype(np.uint8)


# ---------------------------------------------------------------------------------------------
# mesh building blocks: all return (verts (N,3) f64, uvs (N,2) f64, faces (F,3) int64)
# ---------------------------------------------------------------------------------------------
def _grid(origin, du, dv, nu, nv, uv_scale=(1.0, 1.0), displace=None):
    """(nu x nv) quad grid spanned by du, dv from origin -> 2*nu*nv triangles."""
    s = np.linspace(0.0, 1.0, nu + 1)
    t = np.linspace(0.0, 1.0, nv + 1)
    S, T = np.meshgrid(s, t, indexing="ij")
    P = (np.asarray(origin, float)[None, None, :] + S[:, :, None] * np.asarray(du, float)[None, None, :] +
         T[:, :, None] * np.asarray(dv, float)[None, None, :])
    if displace is not None:
        P = P + displace(S, T)
    uv = np.stack([S * uv_scale[0], T * uv_scale[1]], axis=-1)
    idx = np.arange((nu + 1) * (nv + 1)).reshape(nu + 1, nv + 1)
    a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    faces = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)])
    return P.reshape(-1, 3), uv.reshape(-1, 2), faces


def _revolve(center, profile_r, profile_y, nseg):
    """Surface of revolution about +Y through `center`: len(profile)-1 rings x nseg quads."""
    ang = np.linspace(0.0, 2 * np.pi, nseg + 1)
    R = np.asarray(profile_r, float)[:, None]
    Y = np.asarray(profile_y, float)[:, None]
    X = center[0] + R * np.cos(ang)[None, :]
    Z = center[2] + R * np.sin(ang)[None, :]
    P = np.stack([X, center[1] + np.broadcast_to(Y, X.shape), Z], -1)
    nr = len(profile_r)
    uv = np.stack(np.meshgrid(np.linspace(0, 1, nr), np.linspace(0, 4, nseg + 1), indexing="ij"), -1)[..., ::-1]
    idx = np.arange(nr * (nseg + 1)).reshape(nr, nseg + 1)
    a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    faces = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)])
    return P.reshape(-1, 3), uv.reshape(-1, 2), faces


def _arch(p0, p1, y_spring, rise, thickness, depth_vec, nseg):
    """Half-ellipse arch band between column tops p0,p1 (x/z), extruded along depth_vec."""
    t = np.linspace(0.0, np.pi, nseg + 1)
    mid = 0.5 * (np.asarray(p0, float) + np.asarray(p1, float))
    half = 0.5 * (np.asarray(p1, float) - np.asarray(p0, float))
    span = -np.cos(t)
    up = np.sin(t)
    inner = mid[None, :] + span[:, None] * half[None, :]
    inner = np.stack([inner[:, 0], y_spring + rise * up, inner[:, 1]], -1)
    outer = inner.copy()
    outer[:, 1] = y_spring + rise + thickness
    dv = np.asarray(depth_vec, float)
    rows = [outer, inner, inner + dv[None, :], outer + dv[None, :]]
    P = np.stack(rows, 0)  # (4, nseg+1, 3)
    uv = np.stack(np.meshgrid(np.linspace(0, 1, 4), np.linspace(0, 2, nseg + 1), indexing="ij"), -1)
    idx = np.arange(4 * (nseg + 1)).reshape(4, nseg + 1)
    a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    faces = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)])
    return P.reshape(-1, 3), uv.reshape(-1, 2), faces


class _MeshBuilder:
    def __init__(self):
        self.v, self.uv, self.groups = [], [], []
        self.nv = 0

    def add(self, name, mat, parts):
        faces = []
        for P, UV, F in parts:
            faces.append(F + self.nv)
            self.v.append(P)
            self.uv.append(UV)
            self.nv += len(P)
        self.groups.append((name, np.concatenate(faces).astype(np.uint32), mat))

    def tri_count(self):
        return int(sum(len(g[1]) for g in self.groups))

    def finish(self):
        return ObjData(np.concatenate(self.v).astype(np.float32), np.concatenate(self.uv).astype(np.float32),
                       self.groups)


def _mtl(name, seed, tex=None, opacity=None, alpha=1.0, ns=60.0):
    rng = np.random.default_rng(seed)
    m = MtlMaterial(name)
    kd = rng.uniform(0.45, 0.95, 3)
    m.ambiant = tuple(float(np.float32(x)) for x in 0.12 * kd)
    m.diffuse = tuple(float(np.float32(x)) for x in kd)
    m.specular = tuple(float(np.float32(x)) for x in rng.uniform(0.05, 0.4, 3))
    m.shininess = float(ns)
    m.alpha = float(alpha)
    m.diffuse_texture = tex
    m.opacity_map = opacity
    return m


SPONZA_TARGET_TRIS = 262144


def sponza_standin(seed=0, target_tris=SPONZA_TARGET_TRIS, lod=1):
    """Procedural atrium standing in for media/crytek-sponza/sponza.obj.

    OBJ-space bounds ~ (+-1900, 0..1500, +-1200) (the loader divides by 4), two colonnade storeys with
    arches on both long sides, tessellated floor / walls / gallery slabs, hanging drapes, vases and
    two alpha-masked foliage groups (map_d).  Returns (ObjData, textures) where `textures` maps the
    MTL texture paths to decoded uint8 images.  Triangle count == target_tris exactly (the floor
    tessellation absorbs the remainder).  `lod` > 1 divides every fixed tessellation count (small test
    scenes: lod=4 leaves ~15 k fixed triangles).
    """
    def q(n):
        return max(2, int(n) // int(lod))

    rng = np.random.default_rng(seed)
    textures = {}
    for i, kind in enumerate(["brick", "tile", "cloth", "brick", "tile", "cloth"]):
        textures["textures/std_%d.png" % i] = _pattern_texture(256, seed * 101 + i, kind)
    textures["textures/leaf_mask_0.png"] = _leaf_mask(256, seed * 101 + 50)
    textures["textures/leaf_mask_1.png"] = _leaf_mask(256, seed * 101 + 51)

    mb = _MeshBuilder()
    X0, X1, Z0, Z1, H = -1900.0, 1900.0, -1200.0, 1200.0, 1500.0
    ZI = 620.0       # inner edge of the galleries (columns stand here)
    Y1, Y2 = 520.0, 1040.0  # gallery floor heights

    # walls (4 groups, textured)
    mb.add("wall_north", _mtl("wall_n", 1, "textures/std_0.png"), [_grid((X0, 0, Z1), (X1 - X0, 0, 0), (0, H, 0), q(96), q(40), (12, 5))])
    mb.add("wall_south", _mtl("wall_s", 2, "textures/std_0.png"), [_grid((X1, 0, Z0), (X0 - X1, 0, 0), (0, H, 0), q(96), q(40), (12, 5))])
    mb.add("wall_east", _mtl("wall_e", 3, "textures/std_3.png"), [_grid((X1, 0, Z1), (0, 0, Z0 - Z1), (0, H, 0), q(64), q(40), (8, 5))])
    mb.add("wall_west", _mtl("wall_w", 4, "textures/std_3.png"), [_grid((X0, 0, Z0), (0, 0, Z1 - Z0), (0, H, 0), q(64), q(40), (8, 5))])
    # ceiling ring (open roof in the middle, closed above galleries)
    mb.add("roof", _mtl("roof", 5), [
        _grid((X0, H, ZI), (X1 - X0, 0, 0), (0, 0, Z1 - ZI), q(64), q(12)),
        _grid((X0, H, Z0), (X1 - X0, 0, 0), (0, 0, -ZI - Z0), q(64), q(12)),
    ])
    # gallery slabs (two storeys x two sides), each a top and a bottom sheet
    slabs = []
    for y in (Y1, Y2):
        for z0, z1 in ((ZI, Z1), (Z0, -ZI)):
            slabs.append(_grid((X0, y, z0), (X1 - X0, 0, 0), (0, 0, z1 - z0), q(80), q(12), (10, 2)))
            slabs.append(_grid((X0, y - 40.0, z0), (0, 0, z1 - z0), (X1 - X0, 0, 0), q(12), q(80), (2, 10)))
    mb.add("gallery_slabs", _mtl("slabs", 6, "textures/std_4.png"), slabs)

    # columns: 3 storeys x 2 sides x ncol, surfaces of revolution with entasis + capital
    ncol = 14
    xs = np.linspace(X0 + 160.0, X1 - 160.0, ncol)
    prof_t = np.linspace(0, 1, q(16) + 1)
    for storey, (yb, yt) in enumerate(((0.0, Y1 - 40.0), (Y1, Y2 - 40.0), (Y2, H))):
        cols = []
        rad = 46.0 - 6.0 * storey
        for side in (-1.0, 1.0):
            for x in xs:
                r = rad * (1.0 - 0.12 * prof_t + 0.35 * np.exp(-((prof_t - 1.0) / 0.05) ** 2) +
                           0.3 * np.exp(-(prof_t / 0.04) ** 2))
                cols.append(_revolve((x, yb, side * ZI), r, prof_t * (yt - yb), q(24)))
        mb.add("columns_%d" % storey, _mtl("col%d" % storey, 10 + storey, "textures/std_%d.png" % (1 if storey else 3), ns=80.0), cols)

    # arches between neighbouring columns, each storey / side
    for storey, ytop in enumerate((Y1 - 40.0, Y2 - 40.0)):
        arches = []
        for side in (-1.0, 1.0):
            for i in range(ncol - 1):
                arches.append(_arch((xs[i], side * ZI - 30.0), (xs[i + 1], side * ZI - 30.0), ytop - 170.0, 150.0, 20.0,
                                    (0.0, 0.0, 60.0), q(32)))
        mb.add("arches_%d" % storey, _mtl("arch%d" % storey, 20 + storey, "textures/std_0.png"), arches)

    # drapes: wavy cloth sheets hanging between first-storey columns (4 materials)
    for k in range(4):
        drapes = []
        for i in range(k, ncol - 1, 4):
            for side in (-1.0, 1.0):
                ph = rng.uniform(0, 2 * np.pi)

                def disp(S, T, ph=ph, side=side):
                    w = 28.0 * np.sin(6 * np.pi * S + ph) * (0.3 + 0.7 * T)
                    return np.stack([np.zeros_like(S), -18.0 * np.sin(np.pi * S) * T, side * w], -1)

                drapes.append(_grid((xs[i] + 40.0, Y1 - 60.0, side * (ZI - 70.0)), (xs[i + 1] - xs[i] - 80.0, 0, 0),
                                    (0, -330.0, 0), q(32), q(24), (3, 3), disp))
        mb.add("drape_%d" % k, _mtl("drape%d" % k, 30 + k, "textures/std_%d.png" % (2 if k % 2 else 5), ns=20.0), drapes)

    # vases on the floor (2 materials)
    for k in range(2):
        vases = []
        pt = np.linspace(0, 1, q(16) + 1)
        for x in xs[k::2]:
            for side in (-1.0, 1.0):
                r = 30.0 + 45.0 * np.sin(np.pi * pt) ** 2 * (1.0 - 0.4 * pt)
                vases.append(_revolve((x, 0.0, side * (ZI - 200.0)), r, pt * 170.0, q(32)))
        mb.add("vases_%d" % k, _mtl("vase%d" % k, 40 + k, ns=120.0), vases)

    # alpha-masked foliage: crossed quads, densely tessellated, two groups with map_d
    for k in range(2):
        leaves = []
        for x in xs[k::2]:
            for side in (-1.0, 1.0):
                c = np.array([x, 170.0, side * (ZI - 200.0)])
                for a in (0.0, np.pi / 3, 2 * np.pi / 3):
                    d = np.array([np.cos(a), 0.0, np.sin(a)]) * 240.0
                    leaves.append(_grid(c - 0.5 * d, d, (0, 260.0, 0), q(10), q(10), (2, 2)))
        mb.add("foliage_%d" % k, _mtl("leaf%d" % k, 50 + k, "textures/std_%d.png" % (1 + k), "textures/leaf_mask_%d.png" % k), leaves)

    # a lion-head-like bumpy relief on the far wall + a few banners to diversify materials
    def bump(S, T):
        return np.stack([-(60.0 * np.exp(-(((S - 0.5) / 0.22) ** 2 + ((T - 0.5) / 0.22) ** 2)) *
                           (1.0 + 0.3 * np.sin(24 * S) * np.sin(24 * T))), np.zeros_like(S), np.zeros_like(S)], -1)

    mb.add("relief", _mtl("relief", 60, ns=40.0), [_grid((X1 - 5.0, 300.0, -250.0), (0, 0, 500.0), (0, 500.0, 0), q(96), q(96), (1, 1), bump)])
    for k in range(3):
        zb = -300.0 + 300.0 * k
        mb.add("banner_%d" % k, _mtl("banner%d" % k, 70 + k, "textures/std_2.png", ns=15.0),
               [_grid((X0 + 400.0 + 900.0 * k, 1400.0, zb - 60.0), (0, 0, 120.0), (0, -620.0, 0), q(16), q(64), (1, 4),
                      lambda S, T: np.stack([14.0 * np.sin(4 * np.pi * T + S), np.zeros_like(S), np.zeros_like(S)], -1))])

    # floor last: its tessellation absorbs the remaining triangle budget exactly
    remaining = target_tris - mb.tri_count()
    if remaining < 2000:
        raise ValueError("target_tris too small for the fixed geometry (%d used)" % mb.tri_count())
    nq = remaining // 2
    nu = max(1, int(np.sqrt(nq * (X1 - X0) / (Z1 - Z0))))
    nv = max(1, nq // nu)
    parts = [_grid((X0, 0.0, Z0), (0, 0, Z1 - Z0), (X1 - X0, 0, 0), nv, nu, (12, 19))]
    left = remaining - 2 * nu * nv
    if left % 2:
        # an odd remainder cannot come from quads: one extra degenerate-free triangle under the floor
        P = np.array([[X0, -1.0, Z0], [X0 + 10.0, -1.0, Z0], [X0, -1.0, Z0 + 10.0]])
        parts.append((P, np.zeros((3, 2)), np.array([[0, 1, 2]])))
        left -= 1
    if left > 0:
        parts.append(_grid((X0, -2.0, Z0), (0, 0, Z1 - Z0), (X1 - X0, 0, 0), left // 2, 1))
    mb.add("floor", _mtl("floor", 7, "textures/std_1.png", ns=90.0), parts)
    od = mb.finish()
    assert mb.tri_count() == target_tris, (mb.tri_count(), target_tris)
    return od, textures


HAIRBALL_TARGET_TRIS = 2880000


def hairball_standin(seed=0, target_tris=HAIRBALL_TARGET_TRIS, segments=80, radius=4.4, width=0.012):
    """Thin triangle ribbons random-walking inside a sphere, standing in for media/hairball/hairball.obj
    (real asset: ~2.88 M triangles, one group, no uvs).  OBJ-space radius ~4.4 (x1/4 at load)."""
    rng = np.random.default_rng(seed)
    per_strand = 2 * segments
    strands = target_tris // per_strand
    if strands * per_strand != target_tris:
        raise ValueError("target_tris must be a multiple of 2*segments")
    # start points in a small core, initial directions outward, then a correlated random walk that is
    # softly pulled back inside the sphere
    d = rng.normal(size=(strands, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p = d * rng.uniform(0.1, 0.5, (strands, 1)) * radius
    step = radius * 1.6 / segments
    pts = np.empty((strands, segments + 1, 3))
    pts[:, 0] = p
    for s in range(segments):
        d = d + 0.35 * rng.normal(size=(strands, 3))
        r = np.linalg.norm(p, axis=1, keepdims=True)
        d = d - np.clip((r - 0.8 * radius) / (0.2 * radius), 0.0, 4.0) * 0.6 * (p / np.maximum(r, 1e-9))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        p = p + step * d
        pts[:, s + 1] = p
    tang = np.gradient(pts, axis=1)
    ref = rng.normal(size=(strands, 1, 3))
    side = np.cross(tang, ref)
    side /= np.maximum(np.linalg.norm(side, axis=2, keepdims=True), 1e-12)
    w = width * radius / 4.4
    verts = np.stack([pts - 0.5 * w * side, pts + 0.5 * w * side], axis=2)  # (S, L+1, 2, 3)
    idx = np.arange(strands * (segments + 1) * 2, dtype=np.int64).reshape(strands, segments + 1, 2)
    a, b = idx[:, :-1, 0], idx[:, :-1, 1]
    c, e = idx[:, 1:, 1], idx[:, 1:, 0]
    faces = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, e], -1).reshape(-1, 3)])
    return ObjData(verts.reshape(-1, 3).astype(np.float32), None, [("hairball", faces.astype(np.uint32), None)])
