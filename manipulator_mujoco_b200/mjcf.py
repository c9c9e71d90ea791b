"""MJCF -> flat model constants for the UR5e + Hand-E planner scene.

The reference planner builds its model with ``mujoco.MjModel.from_xml_path`` followed by
``mjx.put_model`` (reference ``sampling_based_planner/mjx_planner.py:100-105``).  Neither package
exists in this environment, so this module is a small, self-contained MJCF reader that produces
exactly the constants the rollout needs (body tree, inertias, joints, collision geoms, the static
candidate-pair list, solver options and the compile-time ``invweight0`` / ``meaninertia``
statistics).  It supports the MJCF subset used by the planner scene
(``ur5e_hande_mjx/scene.xml`` + ``ur5e_1_robotiq_hande.xml`` + ``objects.xml``): ``<include>``,
nested ``<default>`` classes with ``childclass``, hinge and free joints, plane / box / capsule
collision geoms, mesh geoms (binary STL, only as a source of inertia), explicit ``<inertial>``,
sites, ``gravcomp``, ``<option>`` flags.

MuJoCo compile rules restated here (all "MJX-recall"):
  * quaternions are normalised; bodies without ``<inertial>`` get mass/inertia from *all* their
    geoms at density 1000 (meshes by exact signed-volume integration);
  * a geom pair is a collision candidate when the contype/conaffinity masks match, the two geoms do
    not belong to the same weld body, and their weld bodies are not parent/child (unless one of
    them is the world);
  * MJX groups candidate pairs by geom-type pair (types sorted ascending) and gives every pair a
    fixed number of contact slots: plane-capsule 2, capsule-capsule 1, capsule-box 2, plane-box 4,
    box-box 4.

Nothing here touches the GPU; the result is a plain dict of numpy arrays (``ModelConsts``) that is
(a) serialised to ``assets/scene_a.json`` so the GPU box never needs the XML, (b) flattened into the
``cemk_model`` C struct (``include/cemk.h``) for the CUDA library and (c) handed unchanged to the
CPU oracle used by the tests.
"""
from __future__ import annotations

import json
import os
import struct
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

# geom type codes: MuJoCo's mjtGeom ordering (pairs are sorted by it)
GEOM_PLANE, GEOM_SPHERE, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = 0, 2, 3, 5, 6, 7
_GEOM_TYPES = {"plane": GEOM_PLANE, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE, "cylinder": GEOM_CYLINDER,
               "box": GEOM_BOX, "mesh": GEOM_MESH}
JNT_FREE, JNT_SLIDE, JNT_HINGE = 0, 2, 3

# contact slots per pair function (MJX collision_driver ncon table)
PAIR_SLOTS = {(GEOM_PLANE, GEOM_CAPSULE): 2, (GEOM_PLANE, GEOM_BOX): 4,
              (GEOM_CAPSULE, GEOM_CAPSULE): 1, (GEOM_CAPSULE, GEOM_BOX): 2,
              (GEOM_BOX, GEOM_BOX): 4}


# --------------------------------------------------------------------------------------------
# small math helpers (float64, host only)
# --------------------------------------------------------------------------------------------
def _vec(s, n=None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=np.float64)
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and v.size != n:
        raise ValueError(f"expected {n} numbers, got {s!r}")
    return v


def quat_normalize(q):
    q = np.asarray(q, dtype=np.float64)
    return q / np.linalg.norm(q)


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw])


def _stl_mass_properties(path):
    """Exact signed-volume integration of a closed binary STL: (volume, com, inertia about com)."""
    with open(path, "rb") as f:
        raw = f.read()
    ntri = struct.unpack_from("<I", raw, 80)[0]
    if len(raw) < 84 + 50 * ntri:
        raise ValueError(f"{path}: not a binary STL")
    rec = np.frombuffer(raw, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]),
                        count=ntri, offset=84)
    v = rec["v"].astype(np.float64)
    a, b, c = v[:, 0], v[:, 1], v[:, 2]
    vol6 = np.einsum("ij,ij->i", a, np.cross(b, c))          # 6 * signed tetra volume (apex 0)
    vol = vol6.sum() / 6.0
    com = ((a + b + c) * vol6[:, None]).sum(0) / 24.0 / vol
    # second moment  int x x^T dV  of a tetra (0,a,b,c) = V/20 * (aa^T + bb^T + cc^T + s s^T)
    s = a + b + c
    cov = (np.einsum("i,ij,ik->jk", vol6, a, a) + np.einsum("i,ij,ik->jk", vol6, b, b)
           + np.einsum("i,ij,ik->jk", vol6, c, c) + np.einsum("i,ij,ik->jk", vol6, s, s)) / 120.0
    cov_c = cov - vol * np.outer(com, com)
    inertia = np.trace(cov_c) * np.eye(3) - cov_c
    if vol < 0:                                              # inward-facing normals
        vol, inertia = -vol, -inertia
    return vol, com, inertia


def _primitive_mass_properties(gtype, size, density):
    """(mass, inertia diag in geom frame) for box / capsule / sphere / cylinder (MuJoCo mjCGeom::SetInertia)."""
    if gtype == GEOM_BOX:
        sx, sy, sz = size[:3]
        m = 8 * sx * sy * sz * density
        return m, np.array([m * (sy * sy + sz * sz) / 3, m * (sx * sx + sz * sz) / 3,
                            m * (sx * sx + sy * sy) / 3])
    if gtype == GEOM_CAPSULE:
        r, h = size[0], 2 * size[1]
        vs, vc = 4.0 / 3.0 * np.pi * r ** 3, np.pi * r * r * h
        ms, mc = density * vs, density * vc
        ixx = mc * (3 * r * r + h * h) / 12 + 0.4 * ms * r * r + ms * h * (3 * r + 2 * h) / 8
        izz = mc * r * r / 2 + 0.4 * ms * r * r
        return ms + mc, np.array([ixx, ixx, izz])
    if gtype == GEOM_SPHERE:
        r = size[0]
        m = density * 4.0 / 3.0 * np.pi * r ** 3
        return m, np.full(3, 0.4 * m * r * r)
    if gtype == GEOM_CYLINDER:
        r, h = size[0], 2 * size[1]
        m = density * np.pi * r * r * h
        ixx = m * (3 * r * r + h * h) / 12
        return m, np.array([ixx, ixx, m * r * r / 2])
    return 0.0, np.zeros(3)


# --------------------------------------------------------------------------------------------
# XML handling: includes + default classes
# --------------------------------------------------------------------------------------------
def _load_xml(path):
    tree = ET.parse(path)
    root = tree.getroot()
    _expand_includes(root, os.path.dirname(os.path.abspath(path)))
    return root


def _expand_includes(elem, base):
    i = 0
    while i < len(elem):
        ch = elem[i]
        if ch.tag == "include":
            sub = ET.parse(os.path.join(base, ch.get("file"))).getroot()
            _expand_includes(sub, base)
            elem.remove(ch)
            for k, sc in enumerate(list(sub)):
                elem.insert(i + k, sc)
            i += len(sub)
        else:
            _expand_includes(ch, base)
            i += 1


def _collect_defaults(root):
    """class name -> {tag: attr dict}, with nested classes inheriting from their parent."""
    classes = {"main": {}}

    def walk(d, parent):
        name = d.get("class") or "main"
        cur = {t: dict(a) for t, a in classes.get(parent, {}).items()} if parent else {}
        cur.update({t: dict(a) for t, a in classes.get(name, {}).items()})
        for ch in d:
            if ch.tag == "default":
                continue
            merged = dict(cur.get(ch.tag, {}))
            merged.update(ch.attrib)
            cur[ch.tag] = merged
        classes[name] = cur
        for ch in d:
            if ch.tag == "default":
                walk(ch, name)

    for d in root.findall("default"):
        walk(d, None)
    return classes


def _attrs(elem, classes, childclass):
    cls = elem.get("class") or childclass or "main"
    out = dict(classes.get(cls, {}).get(elem.tag, {}))
    out.update(elem.attrib)
    return out


# --------------------------------------------------------------------------------------------
# the compiled model
# --------------------------------------------------------------------------------------------
@dataclass
class ModelConsts:
    """Flat, numpy-only description of the compiled scene (see module docstring)."""
    d: dict = field(default_factory=dict)

    def __getattr__(self, k):
        try:
            return self.__dict__["d"][k]
        except KeyError as e:
            raise AttributeError(k) from e

    # ---- lookups used by the planner boundary (reference mjx_planner.py:113,120-121) ----
    def body_id(self, name):
        return self.body_names.index(name)

    def geom_id(self, name):
        return self.geom_names.index(name)

    def site_id(self, name):
        return self.site_names.index(name)

    def to_json(self, path):
        ser = {}
        for k, v in self.d.items():
            if isinstance(v, np.ndarray):
                ser[k] = {"dtype": str(v.dtype), "shape": list(v.shape), "data": v.ravel().tolist()}
            else:
                ser[k] = v
        with open(path, "w") as f:
            json.dump(ser, f, indent=None, separators=(",", ":"))

    @staticmethod
    def from_json(path):
        with open(path) as f:
            ser = json.load(f)
        d = {}
        for k, v in ser.items():
            if isinstance(v, dict) and "dtype" in v:
                d[k] = np.array(v["data"], dtype=v["dtype"]).reshape(v["shape"])
            else:
                d[k] = v
        return ModelConsts(d)


def compile_mjcf(xml_path) -> ModelConsts:
    root = _load_xml(xml_path)
    classes = _collect_defaults(root)
    comp = {}
    for c in root.findall("compiler"):
        comp.update(c.attrib)
    meshdir = os.path.join(os.path.dirname(os.path.abspath(xml_path)), comp.get("meshdir", ""))
    if comp.get("angle", "degree") != "radian":
        raise NotImplementedError("only angle=radian models are supported")
    autolimits = comp.get("autolimits", "true") == "true"

    opt = {"timestep": 0.002, "iterations": 100, "ls_iterations": 50, "tolerance": 1e-8,
           "ls_tolerance": 0.01, "impratio": 1.0, "gravity": [0.0, 0.0, -9.81],
           "eulerdamp": 1, "actuation": 1, "integrator": "Euler"}
    for o in root.findall("option"):
        for k, v in o.attrib.items():
            if k in ("iterations", "ls_iterations"):
                opt[k] = int(v)
            elif k == "gravity":
                opt[k] = [float(x) for x in v.split()]
            elif k == "integrator":
                opt[k] = v
            else:
                opt[k] = float(v)
        for fl in o.findall("flag"):
            for k, v in fl.attrib.items():
                opt[k] = 0 if v == "disable" else 1
    # (the integrator is recorded, not judged: the loader accepts every scene of the reference -- dual_arm_scene.xml asks for
    #  implicitfast -- and kmodel.build_kmodel refuses what the rollout kernel does not implement)

    meshes = {}
    for a in root.findall("asset"):
        for m in a.findall("mesh"):
            name = m.get("name") or os.path.splitext(os.path.basename(m.get("file")))[0]
            meshes[name] = os.path.join(meshdir, m.get("file"))

    bodies, joints, geoms, sites = [], [], [], []

    def add_body(elem, parent, childclass):
        bid = len(bodies)
        if elem.tag == "worldbody":
            b = dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]),
                     gravcomp=0.0, inertial=None, joints=[], geoms=[])
        else:
            cc = elem.get("childclass") or childclass
            childclass = cc
            b = dict(name=elem.get("name"), parent=parent,
                     pos=_vec(elem.get("pos"), 3, [0, 0, 0]),
                     quat=quat_normalize(_vec(elem.get("quat"), 4, [1, 0, 0, 0])),
                     gravcomp=float(elem.get("gravcomp", 0.0)), inertial=None, joints=[], geoms=[])
        bodies.append(b)
        for ch in elem:
            if ch.tag == "inertial":
                a = _attrs(ch, classes, childclass)
                if "fullinertia" in a:
                    raise NotImplementedError("fullinertia")
                b["inertial"] = dict(mass=float(a["mass"]), pos=_vec(a.get("pos"), 3, [0, 0, 0]),
                                     quat=quat_normalize(_vec(a.get("quat"), 4, [1, 0, 0, 0])),
                                     diag=_vec(a["diaginertia"], 3))
            elif ch.tag in ("joint", "freejoint"):
                a = _attrs(ch, classes, childclass)
                jt = "free" if ch.tag == "freejoint" else a.get("type", "hinge")
                if jt not in ("free", "hinge", "slide"):
                    raise NotImplementedError(f"joint type {jt}")
                rng = _vec(a.get("range"), 2, [0, 0])
                limited = a.get("limited", "auto")
                lim = (limited == "true") or (limited == "auto" and autolimits and "range" in a)
                if jt == "free":
                    lim = False
                joints.append(dict(name=a.get("name"), type={"free": JNT_FREE, "slide": JNT_SLIDE, "hinge": JNT_HINGE}[jt],
                                   body=bid, axis=_vec(a.get("axis"), 3, [0, 0, 1]),
                                   pos=_vec(a.get("pos"), 3, [0, 0, 0]), range=rng, limited=lim,
                                   armature=float(a.get("armature", 0.0)),
                                   damping=float(a.get("damping", 0.0)),
                                   margin=float(a.get("margin", 0.0))))
                b["joints"].append(len(joints) - 1)
            elif ch.tag == "geom":
                a = _attrs(ch, classes, childclass)
                gt = a.get("type", "mesh" if "mesh" in a else "sphere")
                g = dict(name=a.get("name"), type=_GEOM_TYPES.get(gt, -1), typename=gt, body=bid,
                         pos=_vec(a.get("pos"), 3, [0, 0, 0]),
                         quat=quat_normalize(_vec(a.get("quat"), 4, [1, 0, 0, 0])),
                         size=np.zeros(3),
                         contype=int(a.get("contype", 1)), conaffinity=int(a.get("conaffinity", 1)),
                         density=float(a.get("density", 1000.0)), mesh=a.get("mesh"),
                         friction=_vec(a.get("friction"), None, [1, 0.005, 0.0001]),
                         solref=_vec(a.get("solref"), 2, [0.02, 1]),
                         solimp=np.array([0.9, 0.95, 0.001, 0.5, 2]),
                         margin=float(a.get("margin", 0.0)), gap=float(a.get("gap", 0.0)),
                         condim=int(a.get("condim", 3)))
                if a.get("size"):
                    sz = _vec(a.get("size"))
                    g["size"] = np.concatenate([sz, np.zeros(3)])[:3]
                if a.get("solimp"):
                    si = _vec(a.get("solimp"))
                    g["solimp"] = np.concatenate([si, [0.9, 0.95, 0.001, 0.5, 2][si.size:]])
                if "fromto" in a:
                    raise NotImplementedError("geom fromto")
                geoms.append(g)
                b["geoms"].append(len(geoms) - 1)
            elif ch.tag == "site":
                a = _attrs(ch, classes, childclass)
                sites.append(dict(name=a.get("name"), body=bid, pos=_vec(a.get("pos"), 3, [0, 0, 0]),
                                  quat=quat_normalize(_vec(a.get("quat"), 4, [1, 0, 0, 0]))))
            elif ch.tag == "body":
                add_body(ch, bid, childclass)

    wb = root.findall("worldbody")
    if len(wb) != 1:
        # several <worldbody> blocks (from includes) are merged by MuJoCo; do the same
        merged = ET.Element("worldbody")
        for w in wb:
            merged.extend(list(w))
        wb = [merged]
    add_body(wb[0], 0, None)

    # MuJoCo numbers geoms / sites / joints by owning body (world's own geoms first)
    def _by_body(items, key):
        order = sorted(range(len(items)), key=lambda i: (items[i]["body"], i))
        remap = {old: new for new, old in enumerate(order)}
        for b in bodies:
            b[key] = [remap[i] for i in b[key]] if key in b else []
        return [items[i] for i in order]

    geoms = _by_body(geoms, "geoms")
    joints = _by_body(joints, "joints")
    sites = _by_body(sites, "sites")
    nbody, njnt, ngeom = len(bodies), len(joints), len(geoms)

    # ---- joint / dof addressing ----
    qposadr, dofadr, nq, nv = [], [], 0, 0
    for j in joints:
        qposadr.append(nq)
        dofadr.append(nv)
        nq += 7 if j["type"] == JNT_FREE else 1
        nv += 6 if j["type"] == JNT_FREE else 1
    body_dofnum = np.zeros(nbody, dtype=np.int32)
    body_dofadr = np.full(nbody, -1, dtype=np.int32)
    body_jntadr = np.full(nbody, -1, dtype=np.int32)
    for bi, b in enumerate(bodies):
        if len(b["joints"]) > 1:
            raise NotImplementedError("more than one joint per body")
        if b["joints"]:
            j = b["joints"][0]
            body_jntadr[bi] = j
            body_dofadr[bi] = dofadr[j]
            body_dofnum[bi] = 6 if joints[j]["type"] == JNT_FREE else 1

    # weld body = nearest ancestor-or-self that has a joint, else world
    body_weldid = np.zeros(nbody, dtype=np.int32)
    body_rootid = np.zeros(nbody, dtype=np.int32)
    for bi in range(1, nbody):
        p = bodies[bi]["parent"]
        body_weldid[bi] = bi if bodies[bi]["joints"] else body_weldid[p]
        body_rootid[bi] = bi if p == 0 else body_rootid[p]
    movable = np.array([body_weldid[bi] != 0 for bi in range(nbody)])

    # ---- inertias ----
    body_mass = np.zeros(nbody)
    body_ipos = np.zeros((nbody, 3))
    body_inertia = np.zeros((nbody, 3, 3))        # about the body COM, in body-frame axes
    for bi, b in enumerate(bodies):
        if bi == 0:
            continue
        if b["inertial"] is not None:
            it = b["inertial"]
            R = quat_to_mat(it["quat"])
            body_mass[bi] = it["mass"]
            body_ipos[bi] = it["pos"]
            body_inertia[bi] = R @ np.diag(it["diag"]) @ R.T
            continue
        parts = []
        for gi in b["geoms"]:
            g = geoms[gi]
            Rg = quat_to_mat(g["quat"])
            if g["type"] == GEOM_MESH:
                # static bodies never enter the dynamics; skip their (possibly missing) meshes
                if not movable[bi]:
                    continue
                path = meshes.get(g["mesh"])
                if path is None or not os.path.exists(path):
                    raise FileNotFoundError(f"mesh {g['mesh']} needed for the inertia of {b['name']}")
                if not path.lower().endswith(".stl"):
                    raise NotImplementedError("only STL meshes can provide inertia")
                vol, com, I = _stl_mass_properties(path)
                m = g["density"] * vol
                parts.append((m, g["pos"] + Rg @ com, Rg @ (g["density"] * I) @ Rg.T))
            elif g["type"] in (GEOM_BOX, GEOM_CAPSULE, GEOM_SPHERE, GEOM_CYLINDER):
                m, diag = _primitive_mass_properties(g["type"], g["size"], g["density"])
                parts.append((m, g["pos"].copy(), Rg @ np.diag(diag) @ Rg.T))
        if parts:
            M = sum(p[0] for p in parts)
            com = sum(p[0] * p[1] for p in parts) / M
            I = np.zeros((3, 3))
            for m, c, Ic in parts:
                d = c - com
                I += Ic + m * (d @ d * np.eye(3) - np.outer(d, d))
            body_mass[bi], body_ipos[bi], body_inertia[bi] = M, com, I

    # ---- collision geoms and the static candidate pair list ----
    # <contact><exclude body1=.. body2=../> (the planner scene ships them commented out, scene.xml:27-45)
    names = [b["name"] for b in bodies]
    excl = []
    for c in root.findall("contact"):
        for e in c.findall("exclude"):
            excl.append((names.index(e.get("body1")), names.index(e.get("body2"))))
    for g in geoms:
        if (g["contype"] or g["conaffinity"]) and g["type"] not in (GEOM_PLANE, GEOM_CAPSULE, GEOM_BOX, GEOM_CYLINDER):
            raise NotImplementedError(f"collision geom type {g['typename']}")
    unknown_pairs = []
    pair_geom, pair_type, pair_nslot, pair_slotadr, collides = _candidate_pairs(
        [g["type"] for g in geoms], [g["body"] for g in geoms], [g["contype"] for g in geoms], [g["conaffinity"] for g in geoms],
        body_weldid, [b["parent"] for b in bodies], excl, unknown=unknown_pairs)

    # ---- equality constraints, tendons, actuators, keyframes: recorded for the loader's callers (scene_mjx.xml couples the
    # two Hand-E fingers with a joint equality and a fixed tendon; the dual-arm scene has 12 position servos).  The rollout
    # kernel implements none of them: kmodel.build_kmodel refuses a model that has any.
    jname = {j["name"]: i for i, j in enumerate(joints)}
    eqs, tendons, acts, keys = [], [], [], []
    for e in root.findall("equality"):
        for c in e:
            a = _attrs(c, classes, None)
            if c.tag != "joint":
                raise NotImplementedError(f"equality type {c.tag}")
            eqs.append(dict(type="joint", joint1=jname[a["joint1"]], joint2=jname[a["joint2"]] if "joint2" in a else -1,
                            polycoef=np.concatenate([_vec(a.get("polycoef"), None, [0, 1, 0, 0, 0]), np.zeros(5)])[:5],
                            solref=_vec(a.get("solref"), 2, [0.02, 1]), solimp=_vec(a.get("solimp"), None, [0.9, 0.95, 0.001, 0.5, 2]),
                            active=a.get("active", "true") == "true"))
    for t in root.findall("tendon"):
        for c in t:
            if c.tag != "fixed":
                raise NotImplementedError(f"tendon type {c.tag}")
            tendons.append(dict(name=c.get("name"), joints=[jname[w.get("joint")] for w in c.findall("joint")],
                                coefs=[float(w.get("coef")) for w in c.findall("joint")]))
    tname = {t["name"]: i for i, t in enumerate(tendons)}
    for ac in root.findall("actuator"):
        for c in ac:
            a = _attrs(c, classes, None)
            acts.append(dict(name=a.get("name"), kind=c.tag, joint=jname.get(a.get("joint"), -1), tendon=tname.get(a.get("tendon"), -1),
                             gaintype=a.get("gaintype", "fixed"), biastype=a.get("biastype", "none"),
                             gainprm=np.concatenate([_vec(a.get("gainprm"), None, [1.0]), np.zeros(3)])[:3],
                             biasprm=np.concatenate([_vec(a.get("biasprm"), None, [0.0]), np.zeros(3)])[:3],
                             ctrlrange=_vec(a.get("ctrlrange"), 2, [0, 0]), forcerange=_vec(a.get("forcerange"), 2, [0, 0])))
    for kf in root.findall("keyframe"):
        for c in kf.findall("key"):
            keys.append(dict(name=c.get("name"), qpos=_vec(c.get("qpos"), None, []), ctrl=_vec(c.get("ctrl"), None, [])))
    col = [gi for gi in range(ngeom) if collides[gi]]

    d = dict(
        xml=os.path.basename(xml_path), nq=nq, nv=nv, nbody=nbody, njnt=njnt, ngeom=ngeom,
        opt=opt,
        body_names=[b["name"] for b in bodies],
        body_parent=np.array([b["parent"] for b in bodies], dtype=np.int32),
        body_pos=np.array([b["pos"] for b in bodies]),
        body_quat=np.array([b["quat"] for b in bodies]),
        body_mass=body_mass, body_ipos=body_ipos, body_inertia=body_inertia,
        body_gravcomp=np.array([b["gravcomp"] for b in bodies]),
        body_weldid=body_weldid, body_rootid=body_rootid,
        body_jntadr=body_jntadr, body_dofadr=body_dofadr, body_dofnum=body_dofnum,
        jnt_names=[j["name"] for j in joints],
        jnt_type=np.array([j["type"] for j in joints], dtype=np.int32),
        jnt_body=np.array([j["body"] for j in joints], dtype=np.int32),
        jnt_axis=np.array([j["axis"] / np.linalg.norm(j["axis"]) for j in joints]).reshape(-1, 3),
        jnt_pos=np.array([j["pos"] for j in joints]).reshape(-1, 3),
        jnt_range=np.array([j["range"] for j in joints]).reshape(-1, 2),
        jnt_limited=np.array([j["limited"] for j in joints], dtype=np.int32),
        jnt_armature=np.array([j["armature"] for j in joints]),
        jnt_damping=np.array([j["damping"] for j in joints]),
        jnt_margin=np.array([j["margin"] for j in joints]),
        jnt_qposadr=np.array(qposadr, dtype=np.int32), jnt_dofadr=np.array(dofadr, dtype=np.int32),
        geom_names=[g["name"] for g in geoms],
        geom_type=np.array([g["type"] for g in geoms], dtype=np.int32),
        geom_body=np.array([g["body"] for g in geoms], dtype=np.int32),
        geom_pos=np.array([g["pos"] for g in geoms]),
        geom_quat=np.array([g["quat"] for g in geoms]),
        geom_size=np.array([g["size"] for g in geoms]),
        geom_friction=np.array([g["friction"][:3] for g in geoms]),
        geom_solref=np.array([g["solref"] for g in geoms]),
        geom_solimp=np.array([g["solimp"] for g in geoms]),
        geom_margin=np.array([g["margin"] for g in geoms]),
        geom_condim=np.array([g["condim"] for g in geoms], dtype=np.int32),
        geom_collides=np.array([gi in col for gi in range(ngeom)], dtype=np.int32),
        site_names=[s["name"] for s in sites],
        site_body=np.array([s["body"] for s in sites], dtype=np.int32),
        site_pos=np.array([s["pos"] for s in sites]).reshape(-1, 3),
        pair_geom=pair_geom, pair_type=pair_type, pair_nslot=pair_nslot, pair_slotadr=pair_slotadr,
        ncon=int(pair_nslot.sum()),
        pair_unknown=np.array(unknown_pairs, dtype=np.int32).reshape(-1, 2),
        nexclude=len(excl),
        neq=len(eqs), eq_joint1=np.array([e["joint1"] for e in eqs], dtype=np.int32), eq_joint2=np.array([e["joint2"] for e in eqs], dtype=np.int32),
        eq_polycoef=np.array([e["polycoef"] for e in eqs]).reshape(-1, 5), eq_active=np.array([e["active"] for e in eqs], dtype=np.int32),
        eq_solref=np.array([e["solref"] for e in eqs]).reshape(-1, 2),
        ntendon=len(tendons), tendon_names=[t["name"] for t in tendons], tendon_joints=[t["joints"] for t in tendons],
        tendon_coefs=[t["coefs"] for t in tendons],
        nu=len(acts), actuator_names=[a["name"] for a in acts], actuator_joint=np.array([a["joint"] for a in acts], dtype=np.int32),
        actuator_tendon=np.array([a["tendon"] for a in acts], dtype=np.int32),
        actuator_gainprm=np.array([a["gainprm"] for a in acts]).reshape(-1, 3), actuator_biasprm=np.array([a["biasprm"] for a in acts]).reshape(-1, 3),
        actuator_ctrlrange=np.array([a["ctrlrange"] for a in acts]).reshape(-1, 2),
        actuator_forcerange=np.array([a["forcerange"] for a in acts]).reshape(-1, 2),
        key_names=[k["name"] for k in keys], key_qpos=[list(map(float, k["qpos"])) for k in keys],
    )
    # qpos0: hinge / slide 0, free joint = body pose
    qpos0 = np.zeros(nq)
    for ji, j in enumerate(joints):
        if j["type"] == JNT_FREE:
            b = bodies[j["body"]]
            qpos0[qposadr[ji]:qposadr[ji] + 3] = b["pos"]
            qpos0[qposadr[ji] + 3:qposadr[ji] + 7] = b["quat"]
    d["qpos0"] = qpos0
    mc = ModelConsts(d)
    _set_const(mc)
    return mc


# --------------------------------------------------------------------------------------------
# compile-time statistics at qpos0 (MuJoCo mj_setConst):  dof_invweight0, body_invweight0, meaninertia
# --------------------------------------------------------------------------------------------
def host_kinematics(mc: ModelConsts, qpos):
    """World poses of bodies (xpos, xmat) for a configuration; float64, host-side helper.

    Used for the compile-time statistics and for the ``data`` stand-in of the planner boundary
    (``mpc_planner.py:113-121`` reads ``data.site_xpos`` / ``data.xquat`` after ``mj_forward``).
    """
    nb = mc.nbody
    xpos = np.zeros((nb, 3))
    xquat = np.zeros((nb, 4))
    xquat[0] = [1, 0, 0, 0]
    for b in range(1, nb):
        p = mc.body_parent[b]
        j = mc.body_jntadr[b]
        if j >= 0 and mc.jnt_type[j] == JNT_FREE:
            a = mc.jnt_qposadr[j]
            xpos[b] = qpos[a:a + 3]
            xquat[b] = quat_normalize(qpos[a + 3:a + 7])
            continue
        Rp = quat_to_mat(xquat[p])
        xpos[b] = xpos[p] + Rp @ mc.body_pos[b]
        xquat[b] = quat_mul(xquat[p], mc.body_quat[b])
        if j >= 0:
            ang = qpos[mc.jnt_qposadr[j]]
            ax = mc.jnt_axis[j]
            if mc.jnt_type[j] == JNT_SLIDE:
                xpos[b] = xpos[b] + quat_to_mat(xquat[b]) @ ax * ang
                continue
            # joint anchor offset jnt_pos is zero in the supported scenes
            if np.any(mc.jnt_pos[j] != 0):
                raise NotImplementedError("non-zero joint pos")
            qj = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * ax])
            xquat[b] = quat_mul(xquat[b], qj)
    xmat = np.array([quat_to_mat(q) for q in xquat])
    return xpos, xquat, xmat


def host_mass_matrix(mc: ModelConsts, qpos):
    """Dense joint-space inertia M(q) (with armature) by summing J^T diag(m, I) J over bodies."""
    xpos, xquat, xmat = host_kinematics(mc, qpos)
    nv = mc.nv
    M = np.zeros((nv, nv))
    jacs = {}
    for b in range(1, mc.nbody):
        if mc.body_mass[b] == 0:
            continue
        com = xpos[b] + xmat[b] @ mc.body_ipos[b]
        jp, jr = host_jac(mc, xpos, xmat, b, com)
        Iw = xmat[b] @ mc.body_inertia[b] @ xmat[b].T
        M += mc.body_mass[b] * jp.T @ jp + jr.T @ Iw @ jr
        jacs[b] = (jp, jr)
    for j in range(mc.njnt):
        n = 6 if mc.jnt_type[j] == JNT_FREE else 1
        for k in range(n):
            M[mc.jnt_dofadr[j] + k, mc.jnt_dofadr[j] + k] += mc.jnt_armature[j]
    return M, jacs, (xpos, xquat, xmat)


def host_jac(mc: ModelConsts, xpos, xmat, body, point):
    """Translational / rotational Jacobian (3 x nv each) of a world point attached to ``body``."""
    nv = mc.nv
    jp, jr = np.zeros((3, nv)), np.zeros((3, nv))
    b = body
    while b != 0:
        j = mc.body_jntadr[b]
        if j >= 0:
            da = mc.jnt_dofadr[j]
            if mc.jnt_type[j] == JNT_HINGE:
                ax = xmat[b] @ mc.jnt_axis[j]
                jr[:, da] = ax
                jp[:, da] = np.cross(ax, point - xpos[b])
            elif mc.jnt_type[j] == JNT_SLIDE:
                jp[:, da] = xmat[b] @ mc.jnt_axis[j]
            else:
                jp[:, da:da + 3] = np.eye(3)
                for k in range(3):
                    ax = xmat[b][:, k]
                    jr[:, da + 3 + k] = ax
                    jp[:, da + 3 + k] = np.cross(ax, point - xpos[b])
        b = mc.body_parent[b]
    return jp, jr


def _set_const(mc: ModelConsts):
    M, jacs, (xpos, xquat, xmat) = host_mass_matrix(mc, mc.qpos0)
    Minv = np.linalg.inv(M)
    nv = mc.nv
    dof_invweight0 = np.diag(Minv).copy()
    for j in range(mc.njnt):
        if mc.jnt_type[j] == JNT_FREE:
            a = mc.jnt_dofadr[j]
            dof_invweight0[a:a + 3] = dof_invweight0[a:a + 3].mean()
            dof_invweight0[a + 3:a + 6] = dof_invweight0[a + 3:a + 6].mean()
    body_invweight0 = np.zeros((mc.nbody, 2))
    for b in range(1, mc.nbody):
        if mc.body_weldid[b] == 0:
            continue
        com = xpos[b] + xmat[b] @ mc.body_ipos[b]
        jp, jr = host_jac(mc, xpos, xmat, b, com)
        body_invweight0[b, 0] = np.trace(jp @ Minv @ jp.T) / 3
        body_invweight0[b, 1] = np.trace(jr @ Minv @ jr.T) / 3
    mc.d["dof_invweight0"] = dof_invweight0
    mc.d["body_invweight0"] = body_invweight0
    mc.d["meaninertia"] = float(np.mean(np.diag(M))) if nv else 1.0


def exclude_body_pairs(mc: ModelConsts, body_pairs) -> ModelConsts:
    """Copy of a compiled model with the candidate pairs between the given body-name pairs removed,
    i.e. what `<contact><exclude body1=... body2=.../>` does at compile time."""
    ex = set()
    for a, b in body_pairs:
        i, j = mc.body_id(a), mc.body_id(b)
        ex.add((min(i, j), max(i, j)))
    keep = [k for k, (g1, g2) in enumerate(mc.pair_geom)
            if (min(mc.geom_body[g1], mc.geom_body[g2]), max(mc.geom_body[g1], mc.geom_body[g2])) not in ex]
    d = dict(mc.d)
    d["pair_geom"], d["pair_type"], d["pair_nslot"] = mc.pair_geom[keep], mc.pair_type[keep], mc.pair_nslot[keep]
    d["pair_slotadr"] = np.concatenate([[0], np.cumsum(d["pair_nslot"])[:-1]]).astype(np.int32)
    d["ncon"] = int(d["pair_nslot"].sum())
    return ModelConsts(d)


def _candidate_pairs(geom_type, geom_body, contype, conaffinity, body_weldid, body_parent, excluded=(), unknown=None):
    """MJX's static candidate pair list (collision_driver: same weld body / parent-child weld bodies / contype-conaffinity
    filters; pairs grouped by geom-type pair, geom1 = the lower type): pair_geom, pair_type, pair_nslot, pair_slotadr, collides.
    Pairs whose type combination has no slot count in PAIR_SLOTS (cylinders: the dual-arm scene) raise, or -- when the caller
    passes a list as ``unknown`` -- are appended to it as (geom1, geom2) and left out of the slot table."""
    col = [g for g in range(len(geom_type)) if (contype[g] or conaffinity[g]) and geom_type[g] in (GEOM_PLANE, GEOM_CAPSULE, GEOM_BOX, GEOM_CYLINDER)]
    for g in range(len(geom_type)):
        if (contype[g] or conaffinity[g]) and g not in col:
            raise NotImplementedError(f"collision geom type {int(geom_type[g])}")
    excl = {(min(a, b), max(a, b)) for a, b in excluded}
    pairs = []
    for ia, ga in enumerate(col):
        for gb in col[ia + 1:]:
            if not ((contype[ga] & conaffinity[gb]) or (contype[gb] & conaffinity[ga])):
                continue
            w1, w2 = body_weldid[geom_body[ga]], body_weldid[geom_body[gb]]
            if w1 == w2:
                continue
            p1, p2 = body_weldid[body_parent[w1]], body_weldid[body_parent[w2]]
            if w1 != 0 and w2 != 0 and (p1 == w2 or p2 == w1):
                continue
            if (min(geom_body[ga], geom_body[gb]), max(geom_body[ga], geom_body[gb])) in excl:
                continue
            g1, g2 = (ga, gb) if geom_type[ga] <= geom_type[gb] else (gb, ga)
            key = (int(geom_type[g1]), int(geom_type[g2]))
            if key not in PAIR_SLOTS:
                if unknown is None:
                    raise NotImplementedError(f"collision pair types {key}")
                unknown.append((g1, g2))
                continue
            pairs.append((key, g1, g2))
    pairs.sort(key=lambda p: (p[0], p[1], p[2]))
    pair_geom = np.array([[p[1], p[2]] for p in pairs], dtype=np.int32).reshape(-1, 2)
    pair_type = np.array([[p[0][0], p[0][1]] for p in pairs], dtype=np.int32).reshape(-1, 2)
    pair_nslot = np.array([PAIR_SLOTS[p[0]] for p in pairs], dtype=np.int32)
    pair_slotadr = np.concatenate([[0], np.cumsum(pair_nslot)[:-1]]).astype(np.int32)
    collides = np.array([g in col for g in range(len(geom_type))], dtype=np.int32)
    return pair_geom, pair_type, pair_nslot, pair_slotadr, collides


def from_mjmodel(m) -> ModelConsts:
    """Import a compiled ``mujoco.MjModel`` instead of compiling the MJCF here: the same
    ModelConsts, taken from MuJoCo's own compiler output, including its mesh inertias, ``body_invweight0`` /
    ``dof_invweight0`` and ``stat.meaninertia``.  Only attribute access on ``m`` is used (``mujoco`` itself is not
    imported), so the day ``mujoco.mjx`` is installable this is what makes tests/test_against_mjx.py compare like with like.
    The candidate pair list is rebuilt with MJX's rules (``_candidate_pairs``); excludes come from ``m.exclude_signature``."""
    nbody, njnt, ngeom, nsite = int(m.nbody), int(m.njnt), int(m.ngeom), int(m.nsite)
    A = lambda x, dt=np.float64: np.array(x, dtype=dt)
    name = lambda kind, i: getattr(m, kind)(i).name
    jnt_type = A(m.jnt_type, np.int32)
    if not np.all(np.isin(jnt_type, (JNT_FREE, JNT_SLIDE, JNT_HINGE))):
        raise NotImplementedError("only hinge, slide and free joints are supported")
    jnt_body = A(m.jnt_bodyid, np.int32)
    jnt_dofadr = A(m.jnt_dofadr, np.int32)
    body_jntadr = A(m.body_jntadr, np.int32)
    if np.any(A(m.body_jntnum, np.int32) > 1):
        raise NotImplementedError("more than one joint per body")
    body_iquat = A(m.body_iquat)
    body_inertia = np.zeros((nbody, 3, 3))
    for b in range(nbody):                                   # MuJoCo stores the principal inertia and its frame
        R = quat_to_mat(body_iquat[b])
        body_inertia[b] = R @ np.diag(A(m.body_inertia)[b]) @ R.T
    geom_type = A(m.geom_type, np.int32)
    geom_body = A(m.geom_bodyid, np.int32)
    body_weldid, body_parent = A(m.body_weldid, np.int32), A(m.body_parentid, np.int32)
    excluded = [(int(s) >> 16, int(s) & 0xFFFF) for s in np.atleast_1d(getattr(m, "exclude_signature", []))]
    pair_geom, pair_type, pair_nslot, pair_slotadr, collides = _candidate_pairs(
        geom_type, geom_body, A(m.geom_contype, np.int32), A(m.geom_conaffinity, np.int32), body_weldid, body_parent, excluded)
    flags = int(m.opt.disableflags)
    opt = {"timestep": float(m.opt.timestep), "iterations": int(m.opt.iterations), "ls_iterations": int(m.opt.ls_iterations),
           "tolerance": float(m.opt.tolerance), "ls_tolerance": float(m.opt.ls_tolerance), "impratio": float(m.opt.impratio),
           "gravity": [float(g) for g in m.opt.gravity],
           "eulerdamp": 0 if flags & (1 << 14) else 1,            # mjDSBL_EULERDAMP
           "actuation": 0 if flags & (1 << 10) else 1,            # mjDSBL_ACTUATION
           "integrator": {0: "Euler", 1: "RK4", 2: "implicit", 3: "implicitfast"}.get(int(m.opt.integrator), "?")}
    d = dict(
        xml="<mujoco.MjModel>", nq=int(m.nq), nv=int(m.nv), nbody=nbody, njnt=njnt, ngeom=ngeom, opt=opt,
        body_names=[name("body", i) for i in range(nbody)],
        body_parent=body_parent, body_pos=A(m.body_pos), body_quat=A(m.body_quat),
        body_mass=A(m.body_mass), body_ipos=A(m.body_ipos), body_inertia=body_inertia,
        body_gravcomp=A(m.body_gravcomp), body_weldid=body_weldid, body_rootid=A(m.body_rootid, np.int32),
        body_jntadr=body_jntadr, body_dofadr=A(m.body_dofadr, np.int32), body_dofnum=A(m.body_dofnum, np.int32),
        jnt_names=[name("joint", i) for i in range(njnt)],
        jnt_type=jnt_type, jnt_body=jnt_body, jnt_axis=A(m.jnt_axis).reshape(-1, 3), jnt_pos=A(m.jnt_pos).reshape(-1, 3),
        jnt_range=A(m.jnt_range).reshape(-1, 2), jnt_limited=A(m.jnt_limited, np.int32),
        jnt_armature=A(m.dof_armature)[jnt_dofadr], jnt_damping=A(m.dof_damping)[jnt_dofadr], jnt_margin=A(m.jnt_margin),
        jnt_qposadr=A(m.jnt_qposadr, np.int32), jnt_dofadr=jnt_dofadr,
        geom_names=[name("geom", i) for i in range(ngeom)],
        geom_type=geom_type, geom_body=geom_body, geom_pos=A(m.geom_pos), geom_quat=A(m.geom_quat), geom_size=A(m.geom_size),
        geom_friction=A(m.geom_friction), geom_solref=A(m.geom_solref), geom_solimp=A(m.geom_solimp), geom_margin=A(m.geom_margin),
        geom_condim=A(m.geom_condim, np.int32), geom_collides=collides,
        site_names=[name("site", i) for i in range(nsite)], site_body=A(m.site_bodyid, np.int32), site_pos=A(m.site_pos).reshape(-1, 3),
        pair_geom=pair_geom, pair_type=pair_type, pair_nslot=pair_nslot, pair_slotadr=pair_slotadr, ncon=int(pair_nslot.sum()),
        qpos0=A(m.qpos0), dof_invweight0=A(m.dof_invweight0), body_invweight0=A(m.body_invweight0).reshape(nbody, 2),
        meaninertia=float(m.stat.meaninertia),
        neq=int(getattr(m, "neq", 0)), ntendon=int(getattr(m, "ntendon", 0)), nu=int(getattr(m, "nu", 0)),
    )
    return ModelConsts(d)


ModelConsts.from_mjmodel = staticmethod(from_mjmodel)

DEFAULT_ASSET = os.path.join(os.path.dirname(__file__), "assets", "scene_a.json")


def load_model(model_path=None) -> ModelConsts:
    """Scene constants: compile ``model_path`` (an MJCF file) or load the shipped scene-A table."""
    if model_path is None:
        return ModelConsts.from_json(DEFAULT_ASSET)
    if model_path.endswith(".json"):
        return ModelConsts.from_json(model_path)
    return compile_mjcf(model_path)
