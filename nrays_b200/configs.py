"""The BASELINE.json configurations as concrete inputs.

Scene content (camera, light, geometry, materials) restates the values of the reference's scene
files, cited per config; resolution and `aa` are harness overrides (the files ship other values:
SURVEY F7).  The `.scene` text is generated from those values and fed through the loader twin, so
the parser is on the path of every config.  Assets are the seeded stand-ins of assets.py.
"""
from . import assets
from .loader3d import AssetResolver, load_scene, parse

# scenes/basic_materials.mtl — values restated (Ns / Ka / Kd / d / map_Kd per material)
_BASIC_MATERIALS = [
    ("blue", dict(Ns=100, Ka=(0, 0, 0.1), Kd=(0, 0, 1))),
    ("red", dict(Ns=100, Ka=(0.1, 0, 0), Kd=(1, 0, 0))),
    ("green", dict(Ns=100, Ka=(0, 0.1, 0), Kd=(0, 1, 0))),
    ("globe", dict(Ns=100, Ka=(1, 1, 1), Kd=(1, 1, 1), map_Kd="media/globe.png")),
    ("bright_red", dict(Ns=100, Ka=(1, 0, 0), Kd=(1, 0, 0))),
    ("transparent_green", dict(d=0.2, Ns=100, Ka=(0, 0.1, 0), Kd=(0, 1, 0))),
    # the reference's names and colours disagree for red/blue (basic_materials.mtl:34-44); kept as is
    ("transparent_red", dict(d=0.2, Ns=100, Ka=(0, 0, 0.1), Kd=(0, 0, 1))),
    ("transparent_blue", dict(d=0.2, Ns=100, Ka=(0.1, 0, 0), Kd=(1, 0, 0))),
    ("transparent_default", dict(d=0.2, Ns=100, Ka=(0.1, 0.1, 0.1), Kd=(1, 1, 1), Ks=(1, 1, 1))),
]


def _mtl_text(materials):
    out = []
    for name, f in materials:
        out.append("newmtl %s" % name)
        for k in ("d", "Ns"):
            if k in f:
                out.append("%s %s" % (k, f[k]))
        for k in ("Ka", "Kd", "Ks"):
            if k in f:
                out.append("%s %s %s %s" % ((k,) + tuple(float(x) for x in f[k])))
        if "map_Kd" in f:
            out.append("map_Kd %s" % f["map_Kd"])
        out.append("")
    return "\n".join(out)


def _scene_text(camera, lights, geoms, mtllib=None):
    out = []
    if mtllib:
        out.append("mtllib %s" % mtllib)
    out += ["camera", "  output out.png", "  resolution %d %d" % tuple(camera["resolution"]),
            "  eye %s %s %s" % tuple(camera["eye"]), "  at %s %s %s" % tuple(camera["at"]),
            "  fovy %s" % camera["fovy"]]
    for l in lights:
        out += ["light", "  pos %s %s %s" % tuple(l["pos"]), "  color %s %s %s" % tuple(l.get("color", (1.0, 1.0, 1.0)))]
        if "radius" in l:
            out.append("  radius %s" % l["radius"])
        if "nsample" in l:
            out.append("  nsample %s" % l["nsample"])
    for g in geoms:
        out.append("geometry")
        out.append("  %s %s" % (g["shape"][0], " ".join(str(x) for x in g["shape"][1:])))
        out.append("  pos %s %s %s" % tuple(g["pos"]))
        out.append("  angle %s %s %s" % tuple(g.get("angle", (0.0, 0.0, 0.0))))
        out.append("  material %s" % g["material"])
        if "refl" in g:
            out.append("  refl %s %s" % tuple(g["refl"]))
        if "refr" in g:
            out.append("  refr %s" % g["refr"])
    return "\n".join(out) + "\n"


def _basic_resolver(seed=0, globe_size=(2048, 1024)):
    r = AssetResolver()
    r.files["basic_materials.mtl"] = _mtl_text(_BASIC_MATERIALS)
    r.textures["media/globe.png"] = assets.globe_texture(globe_size[0], globe_size[1], seed)
    return r


def primitives_text(light_radius=0.1):
    """scenes/primitives.scene:3-53: ball / box / cone / cylinder / reflective plane, one area light."""
    return _scene_text(
        dict(resolution=(4096, 4096), eye=(0.0, 5.0, -20.0), at=(0.0, 0.0, 0.0), fovy=45.0),
        [dict(pos=(0.0, 0.0, 0.0), radius=light_radius, nsample=10)],
        [
            dict(shape=("ball", 1.0), pos=(-2.1, 0.0, 0.0), material="default", refl=(0.0, 0.0), refr=1.5),
            dict(shape=("box", 1.0, 1.0, 1.0), pos=(2.1, 0.0, 0.0), material="transparent_red", refl=(0.0, 0.0), refr=1.5),
            dict(shape=("cone", 1.0, 1.0), pos=(0.0, -2.1, 0.0), material="transparent_blue", refl=(0.0, 0.0), refr=1.5),
            dict(shape=("cylinder", 1.0, 1.0), pos=(0.0, 2.1, 1.0), material="transparent_green", refl=(0.0, 0.0), refr=1.5),
            dict(shape=("plane", 0.0, 1.0, 0.0), pos=(0.0, -3.0, 0.0), material="default", refl=(0.2, 0.5)),
        ],
        mtllib="basic_materials.mtl")


def balls_text():
    """scenes/balls.scene:3-33: three unit balls (uvs / normals / globe), refl (0.2, 0.2), point light."""
    return _scene_text(
        dict(resolution=(1024, 1024), eye=(0.0, 5.0, -10.0), at=(0.0, 0.0, 0.0), fovy=45.0),
        [dict(pos=(0.0, 10.0, 0.0))],
        [
            dict(shape=("ball", 1.0), pos=(-2.1, 0.0, 0.0), material="uvs", refl=(0.2, 0.2)),
            dict(shape=("ball", 1.0), pos=(2.1, 0.0, 0.0), material="normals", refl=(0.2, 0.2)),
            dict(shape=("ball", 1.0), pos=(0.0, 0.0, 0.0), material="globe", refl=(0.2, 0.2)),
        ],
        mtllib="basic_materials.mtl")


def sponza_text():
    """scenes/crytek_sponza.scene:1-17: eye = light = (-250, 50, 0), at (0, 50, 0), fovy 45, refl 0."""
    return _scene_text(
        dict(resolution=(1024, 1024), eye=(-250.0, 50.0, 0.0), at=(0.0, 50.0, 0.0), fovy=45.0),
        [dict(pos=(-250.0, 50.0, 0.0))],
        [dict(shape=("obj", "media/crytek-sponza/sponza.obj", "media/crytek-sponza"), pos=(0.0, 0.0, 0.0),
              material="default", refl=(0.0, 0.0))])


def hairball_text():
    """scenes/hairball.scene:1-17: eye = light = (0, 0.2, -5), fovy 25, node pos (0, 0.1, 0), angle (0, 0.1, 0)."""
    return _scene_text(
        dict(resolution=(1024, 1024), eye=(0.0, 0.2, -5.0), at=(0.0, 0.2, 0.0), fovy=25.0),
        [dict(pos=(0.0, 0.2, -5.0))],
        [dict(shape=("obj", "media/hairball/hairball.obj", "media/hairball"), pos=(0.0, 0.1, 0.0), angle=(0.0, 0.1, 0.0),
              material="default", refl=(0.0, 0.0))])


def sponza_resolver(seed=0, target_tris=assets.SPONZA_TARGET_TRIS, lod=1):
    r = AssetResolver()
    od, tex = assets.sponza_standin(seed, target_tris, lod)
    r.objs["media/crytek-sponza/sponza.obj"] = od
    for k, v in tex.items():
        r.textures["media/crytek-sponza/" + k] = v
    return r


def hairball_resolver(seed=0, target_tris=assets.HAIRBALL_TARGET_TRIS, segments=80):
    r = AssetResolver()
    r.objs["media/hairball/hairball.obj"] = assets.hairball_standin(seed, target_tris, segments)
    return r


# name -> (text fn, resolver fn, width, height, spp, window) — BASELINE.md §3
CONFIGS = {
    "C1": dict(name="primitives.scene 256x256 1spp", text=primitives_text, resolver=_basic_resolver,
               width=256, height=256, spp=1, window=0.0),
    "C2": dict(name="balls.scene 1024x1024 4spp (synthetic globe texture)", text=balls_text, resolver=_basic_resolver,
               width=1024, height=1024, spp=4, window=1.0),
    "C3": dict(name="crytek_sponza.scene 1920x1080 4spp (synthetic stand-in mesh, 262144 tris, seed 0)",
               text=sponza_text, resolver=sponza_resolver, width=1920, height=1080, spp=4, window=1.0),
    "C4": dict(name="hairball.scene 1920x1080 8spp (synthetic stand-in mesh, 2880000 tris, seed 0)",
               text=hairball_text, resolver=hairball_resolver, width=1920, height=1080, spp=8, window=1.0),
    "C5": dict(name="crytek_sponza.scene 3840x2160 16spp tile-sharded (synthetic stand-in mesh)",
               text=sponza_text, resolver=sponza_resolver, width=3840, height=2160, spp=16, window=1.0),
}


def build(config, device=0, upload=True, **resolver_kw):
    """Returns (scene, camera, cfg) for a BASELINE config id ("C1".."C5")."""
    cfg = CONFIGS[config]
    scene, cameras = load_scene(cfg["text"](), cfg["resolver"](**resolver_kw), device=device, upload=upload)
    return scene, cameras[0], cfg


def build_flat(config, **resolver_kw):
    """Flattened tables only (no device) — what tests hand to the oracle."""
    return build(config, upload=False, **resolver_kw)
