"""Twin of src/obj.rs: the reference's "simplistic obj loader" (§8f rank 3: asset ingest -> flattened scene).

Semantics kept from the reference (file:line = /root/reference/src/obj.rs):
  * one shared vertex / uv pool for the whole file, de-duplicated on the (v, vt, vn) index triple
    (`reformat`, :327-397); every non-empty group becomes one mesh over that pool;
  * groups: the file starts in group `basename`; `g a b` selects/creates "basename/a b" (:310-325);
    a second `usemtl` inside one group opens "auto_generated_group_/<group id><material>" (:149-160);
  * faces are triangulated on the fly as a fan whose pivot is `g[len - i]` (:232-239) — the first vertex
    for quads, NOT for pentagons and beyond (the reference mis-triangulates those; replicated);
  * 1-based and negative (relative) indices (:250-272); if ANY face vertex lacks a vt (vn) index, uvs
    (normals) are dropped for the whole file (:241-247, :104-118) and uvs become zeros (:383);
  * vertex normals are parsed but never reach the renderer, so they are not returned.
Deviation: the reference iterates its groups in HashMap order (unspecified); this twin uses
first-appearance order.
"""
import os

import numpy as np

from .loader3d import ObjData, SceneFileError, parse_mtl

_MAX = 2 ** 31 - 1


def parse_file(path, mtl_base_dir, basename=""):
    with open(path, "r", errors="replace") as f:
        text = f.read()

    def read_text(name):
        with open(os.path.join(mtl_base_dir, name), "r", errors="replace") as g:
            return g.read()

    return parse(text, read_text, basename)


def parse(string, read_mtl=None, basename=""):
    """obj::parse (:62-120).  `read_mtl(filename) -> text` serves `mtllib` lines (missing files only warn, :183)."""
    coords, uvs = [], []
    n_normals = 0
    groups = {basename: 0}
    groups_ids = [[]]
    curr_group = 0
    ignore_normals = ignore_uvs = False
    mtllib, group2mtl = {}, {}
    curr_mtl = None

    def parse_g(words, prefix):
        suffix = " ".join(words)
        name = prefix if not suffix else "%s/%s" % (prefix, suffix)
        if name not in groups:
            groups_ids.append([])
            groups[name] = len(groups_ids) - 1
        return groups[name]

    for l, line in enumerate(string.splitlines()):
        words = line.split()
        if not words or words[0].startswith("#"):
            continue
        tag, rest = words[0], words[1:]
        if tag == "v":
            if len(rest) < 3:
                raise SceneFileError("At line %d: 3 components were expected, found %d." % (l, len(rest)))
            coords.append((np.float32(rest[0]), np.float32(rest[1]), np.float32(rest[2])))
        elif tag == "vn":
            if not ignore_normals:
                n_normals += 1
        elif tag == "vt":
            if not ignore_uvs:
                if len(rest) < 2:
                    raise SceneFileError("At line %d: at least 2 components were expected, found %d." % (l, len(rest)))
                uvs.append((np.float32(rest[0]), np.float32(rest[1])))
        elif tag == "f":
            g = groups_ids[curr_group]
            i = 0
            for word in rest:
                ids = [_MAX, _MAX, _MAX]
                for k, w in enumerate(word.split("/")[:3]):
                    if k == 0 or w:
                        try:
                            ids[k] = int(w) - 1
                        except ValueError as e:
                            raise SceneFileError("At line %d: failed to parse `%s' as a i32: %s" % (l, w, e))
                if i > 2:  # on-the-fly fan triangulation (:232-239)
                    p1, p2 = g[len(g) - i], g[len(g) - 1]
                    g.append(p1)
                    g.append(p2)
                if ids[1] == _MAX:
                    ignore_uvs = True
                if ids[2] == _MAX:
                    ignore_normals = True
                x = len(coords) + ids[0] + 1 if ids[0] < 0 else ids[0]
                y = len(uvs) + ids[1] + 1 if ids[1] < 0 else ids[1]
                z = n_normals + ids[2] + 1 if ids[2] < 0 else ids[2]
                g.append((x, y, z))
                i += 1
            if i < 2 and g:  # not enough vertices: repeat the last one (:279-284)
                for _ in range(3 - i):
                    g.append(g[-1])
        elif tag == "g":
            curr_group = parse_g(rest, basename)
            if curr_mtl is not None:
                group2mtl[curr_group] = curr_mtl
        elif tag == "mtllib":
            if read_mtl is not None:
                try:
                    for m in parse_mtl(read_mtl(" ".join(rest))):
                        mtllib[m.name] = m
                except OSError:
                    pass  # the reference only warns (:183)
        elif tag == "usemtl":
            mname = " ".join(rest)
            if mname != "None":
                m = mtllib.get(mname)
                if m is None:
                    curr_mtl = None
                elif curr_group not in group2mtl:
                    group2mtl[curr_group] = m
                    curr_mtl = m
                else:  # several usemtl in one group: auto-generated group (:149-160)
                    curr_group = parse_g((str(curr_group) + mname).split(), "auto_generated_group_")
                    group2mtl[curr_group] = m
                    curr_mtl = m
            else:
                curr_mtl = None
    return _reformat(coords, None if ignore_uvs else uvs, groups_ids, groups, group2mtl)


def _reformat(coords, uvs, groups_ids, groups, group2mtl):
    """reformat (:327-397): de-duplicate on the index triple, build per-group face lists."""
    vt2id = {}
    resc, resu = [], ([] if uvs is not None else None)
    out_groups = []
    for name, gi in groups.items():
        ids = []
        for point in groups_ids[gi]:
            key = point
            idx = vt2id.get(key)
            if idx is None:
                idx = len(resc)
                if point[0] < 0 or point[0] >= len(coords):
                    raise SceneFileError("face references vertex %d of %d" % (point[0] + 1, len(coords)))
                resc.append(coords[point[0]])
                if resu is not None:
                    if point[1] < 0 or point[1] >= len(uvs):
                        raise SceneFileError("face references texture coordinate %d of %d" % (point[1] + 1, len(uvs)))
                    resu.append(uvs[point[1]])
                vt2id[key] = idx
            ids.append(idx)
        if len(ids) % 3 != 0:
            raise SceneFileError("group %r: face index count %d is not a multiple of 3 (assert at src/obj.rs:370)" % (name, len(ids)))
        faces = np.asarray(ids, dtype=np.uint32).reshape(-1, 3)
        if len(faces):
            out_groups.append((name, faces, group2mtl.get(gi)))
    c = np.asarray(resc, dtype=np.float32).reshape(-1, 3)
    u = None if resu is None else np.asarray(resu, dtype=np.float32).reshape(-1, 2)
    return ObjData(c, u, out_groups)


def write_obj(od, mtl_name=None):
    """Serialise ObjData as Wavefront OBJ text (+ MTL text) — used to push the synthetic stand-ins through
    the same text -> parser -> flattened-scene path a real asset takes."""
    out = []
    mtl = []
    if mtl_name:
        out.append("mtllib %s" % mtl_name)
    for p in od.coords:
        out.append("v %r %r %r" % (float(p[0]), float(p[1]), float(p[2])))
    if od.uvs is not None:
        for t in od.uvs:
            out.append("vt %r %r" % (float(t[0]), float(t[1])))
    seen = set()
    for name, faces, m in od.groups:
        # `usemtl` first: a `g` line hands the current material to the new group (src/obj.rs:95), so this order
        # keeps one group per (name, material) when the text is parsed back
        out.append("usemtl %s" % (m.name if m is not None else "None"))
        out.append("g %s" % name)
        if m is not None:
            if m.name not in seen:
                seen.add(m.name)
                mtl.append("newmtl %s" % m.name)
                mtl.append("Ka %r %r %r" % tuple(float(x) for x in m.ambiant))
                mtl.append("Kd %r %r %r" % tuple(float(x) for x in m.diffuse))
                mtl.append("Ks %r %r %r" % tuple(float(x) for x in m.specular))
                mtl.append("Ns %r" % float(m.shininess))
                mtl.append("d %r" % float(m.alpha))
                if m.diffuse_texture:
                    mtl.append("map_Kd %s" % m.diffuse_texture)
                if m.opacity_map:
                    mtl.append("map_d %s" % m.opacity_map)
        fmt = "f %d/%d %d/%d %d/%d" if od.uvs is not None else "f %d %d %d"
        for f in faces:
            a, b, c = int(f[0]) + 1, int(f[1]) + 1, int(f[2]) + 1
            out.append(fmt % ((a, a, b, b, c, c) if od.uvs is not None else (a, b, c)))
    return "\n".join(out) + "\n", "\n".join(mtl) + "\n"
